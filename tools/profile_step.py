"""Runs a few steps of one workload for ncu (launch list / --set full captures).
usage: python tools/profile_step.py {encode|decode|maskgit} {f32|bf16} [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import phenaki_pytorch_b200 as P  # noqa: E402
from phenaki_pytorch_b200 import _lib as L  # noqa: E402

what, prec = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prec = L.PREC_BF16 if prec == "bf16" else L.PREC_F32
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if what == "encode":
    model = P.CViViT(**bench.CFG2).to(dev).eval()
    model.precision = prec
    video = torch.randn(bench.VIDEO, device=dev)
    for _ in range(steps):
        model(video, return_only_codebook_ids=True)
elif what == "decode":
    model = P.CViViT(**bench.CFG2).to(dev).eval()
    model.precision = prec
    ids = torch.randint(0, 65536, (8, 9, 8, 8), device=dev)
    for _ in range(steps):
        model.decode_from_codebook_indices(ids)
else:
    cv = P.CViViT(**bench.CFG2).to(dev)
    mg = P.MaskGit(**bench.CFG3).to(dev)
    mg.precision = prec
    ph = P.Phenaki(cvivit=cv, maskgit=mg, steps=steps, text_embed_dim=768)
    ctx = torch.randn(4, 16, 768, device=dev)
    ph.sample_token_ids(num_tokens=576, patch_shape=(9, 8, 8), batch_size=4, text_embeds=ctx, cond_scale=3.0)
torch.cuda.synchronize()
