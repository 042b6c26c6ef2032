"""Compact per-kernel table from an ncu report (--set full): python tools/ncu_summary.py report.ncu-rep [out.csv]
One row per distinct (kernel, grid): duration, DRAM read/write bytes, L2->SM bytes, tensor-pipe / SM / L2 utilisation,
issue activity, registers.  Runs `ncu -i ... --page raw --csv` (ncu must be on PATH)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm"),
     ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_%"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
     ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_insts")]
col = {h: i for i, h in enumerate(hdr)}
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}


def val(r, m):
    if m not in col or r[col[m]] in ("", "no data", "n/a"):
        return ""
    v, u = float(r[col[m]].replace(",", "")), units[col[m]]
    if m.startswith("dram__bytes") or m.startswith("l1tex__m_xbar"):
        return f"{v * SC.get(u, 1) / 1e6:.3f} MB"
    if m == "gpu__time_duration.sum":
        return f"{v * SC.get(u, 1):.2f} us"
    return f"{v:.1f}" if "%" in dict(M)[m] else f"{v:.0f}"


STALL = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h]


def stalls(r):
    """top warp-stall reasons (cycles a warp waits in that state per issued instruction), WarpStateStats section"""
    vals = []
    for h in STALL:
        try:
            vals.append((float(r[col[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per_")[0]))
        except (ValueError, IndexError):
            pass
    return " ".join(f"{n}={v:.1f}" for v, n in sorted(vals, reverse=True)[:5])


def dram_mb(r):
    t = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if m in col and r[col[m]] not in ("", "no data", "n/a"):
            t += float(r[col[m]].replace(",", "")) * SC.get(units[col[m]], 1) / 1e6
    return t


seen, out, traffic = {}, [], {}
for r in rows[2:]:
    key = (r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("phk::", "").replace("<unnamed>::", ""), r[col["Grid Size"]])
    seen[key] = seen.get(key, 0) + 1
    traffic[key] = traffic.get(key, 0.0) + dram_mb(r)
    if seen[key] > 1:
        continue
    out.append([key[0], key[1]] + [val(r, m) for m, _ in M] + [stalls(r)])
for row in out:  # how many launches of this (kernel, grid) the capture holds and their mean DRAM traffic (read + write)
    k = (row[0], row[1])
    row += [seen[k], f"{traffic[k] / seen[k]:.3f} MB"]
w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
w.writerow(["kernel", "grid"] + [n for _, n in M] + ["top_stalls_cycles_per_issue", "launches_captured", "dram_rd_wr_mean"])
w.writerows(out)
