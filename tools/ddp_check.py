"""Two-or-more-GPU check of the data-parallel training step (torchrun, NCCL): the gradient all-reduce launched slice by
slice on a side stream while the backward still runs (phk_train_set_progress_events) must give the same averaged gradients
as one all-reduce of the whole bucket after the step, and both must equal the mean of the per-rank gradients.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ddp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import phenaki_pytorch_b200 as P  # noqa: E402
from phenaki_pytorch_b200 import _lib as L  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)  # same weights on every rank
cv = P.CViViT(dim=64, codebook_size=256, image_size=(16, 24), patch_size=(8, 8), temporal_patch_size=3, spatial_depth=1,
              temporal_depth=1, dim_head=32, heads=2, use_vgg_and_gan=False).to(dev)
mg = P.MaskGit(dim=256, num_tokens=256, max_seq_len=64, heads=4, dim_head=64, depth=3, dim_context=48).to(dev)
mg.precision = L.PREC_BF16
ph = P.Phenaki(cvivit=cv, maskgit=mg, steps=6, text_embed_dim=48).to(dev).train()
g = torch.Generator().manual_seed(100 + rank)  # different data per rank
ids = torch.randint(0, 256, (3, 3, 2, 3), generator=g).to(dev)
ctx = torch.randn((3, 5, 48), generator=g).to(dev)
draws = {"rand_step": torch.randint(0, 6, (3,), generator=g), "perm": torch.rand((3, 18), generator=g)}


def grads(sync, overlap):
    ph.sync_gradients = sync
    if not overlap:
        mg._overlap_plan = lambda *a, **k: None
    elif "_overlap_plan" in mg.__dict__:
        del mg.__dict__["_overlap_plan"]
    for p in mg.parameters():
        p.grad = None
    loss = ph(video_codebook_ids=ids, text_embeds=ctx, draw_fn=lambda shape, tag: draws[tag])
    loss.backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in mg.named_parameters() if p.grad is not None}


local_g = grads(False, False)
mean_g = {}
for k, v in local_g.items():
    t = v.clone()
    dist.all_reduce(t)
    mean_g[k] = t / world
plain = grads(True, False)
over = grads(True, True)
used_overlap = getattr(mg, "_overlap_cache", None) is not None
worst_a = worst_b = 0.0
for k in mean_g:
    if mean_g[k].numel() == 0:
        continue
    scale = max(mean_g[k].abs().max().item(), 1e-12)
    worst_a = max(worst_a, (plain[k] - mean_g[k]).abs().max().item() / scale if scale > 1e-7 else 0.0)
    worst_b = max(worst_b, (over[k] - mean_g[k]).abs().max().item() / scale if scale > 1e-7 else 0.0)
ok = worst_a < 1e-5 and worst_b < 1e-5 and used_overlap
res = torch.tensor([float(ok)], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DDP_CHECK world={world} overlap_used={used_overlap} worst_rel_err whole-bucket={worst_a:.2e} sliced-overlapped={worst_b:.2e} "
          f"{'OK' if res.item() == 1.0 else 'FAILED'}")
dist.destroy_process_group()
sys.exit(0 if res.item() == 1.0 else 1)
