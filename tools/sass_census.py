"""SASS opcode census of libphk.so: per kernel, the Blackwell-specific mnemonics that prove tcgen05 / TMEM / TMA / cluster
use (B200_PROFILING.md): UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG
(cp.async.bulk.tensor load / store), UTCBAR (tcgen05.commit), SYNCS (mbarrier), UCGABAR (cluster barrier), HMMA (mma.sync),
LDSM (ldmatrix), MUFU.   usage: python tools/sass_census.py [libphk.so] > profiles/r02/sass_census.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "phenaki_pytorch_b200/libphk.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "HMMA", "LDSM",
        "MUFU", "REDUX", "ATOM", "RED"]
fn, census, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn).replace("phk::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        census[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        total[fn] += 1
        for k in KEYS:
            if op.startswith(k):
                census[fn][k] += 1
                break
print(f"# {lib}: {len(census)} kernels; columns = static SASS instruction counts")
print(f"{'kernel':72s} {'insts':>7s}  " + "  ".join(f"{k}" for k in KEYS))
for fn, c in sorted(census.items(), key=lambda kv: -(kv[1]['UTCHMMA'] + kv[1]['UTCHMMA.2CTA'] + kv[1]['UTMALDG'] + kv[1]['HMMA'])):
    if not any(c[k] for k in KEYS[:12]):
        continue
    print(f"{fn[:72]:72s} {total[fn]:7d}  " + "  ".join(f"{c[k]:{len(k)}d}" for k in KEYS))
