"""N>1 host logic on CPU: two `gloo` ranks shard a batch the way bench.py / a data-parallel caller does (contiguous
video shards, no data-path collective), each runs the encode on its shard (here: the CPU oracle stands in for the device
call -- the sharding code under test is backend-independent), ids are gathered and must equal the single-process
result; the timing reduction must be the max over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phenaki_pytorch_b200 import sharding as S
from tests import cases as C


def test_shard_range_partitions_every_batch():
    for n in range(0, 19):
        for w in (1, 2, 3, 4, 8):
            spans = [S.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))           # contiguous, no overlap
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n_videos, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        torch.set_num_threads(2)
        from oracle import phenaki_oracle as O
        import phenaki_pytorch_b200 as P
        case = C.CVIVIT_CASES["cfg1"]
        torch.manual_seed(case["seed"])
        sd = P.CViViT(**case["ctor"]).state_dict()                              # replicated weights
        shape = (n_videos,) + tuple(case["video"][1:])
        video = C.seeded_randn(shape, case["video_seed"])                       # the global batch
        assert S.world() == (rank, world_size)
        mine = S.shard_batch(video)
        lo, hi = S.shard_range(n_videos, rank, world_size)
        assert mine.shape[0] == hi - lo
        with torch.no_grad():
            if mine.shape[0]:
                ids = O.cvivit_codebook_ids(mine, sd, (64, 64), (16, 16))
            else:
                ids = torch.empty((0, 3, 4, 4), dtype=torch.int64)
        full = S.gather_batch(ids, n_videos)
        slowest = S.max_over_ranks(10.0 + rank)
        if rank == 0:
            with torch.no_grad():
                ref = O.cvivit_codebook_ids(video, sd, (64, 64), (16, 16))
            torch.save(dict(full=full, ref=ref, slowest=slowest, seeds=[S.rank_seed(7, r) for r in range(world_size)]), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_videos", [4, 3, 1])
def test_two_rank_gloo_encode_matches_single_process(tmp_path, n_videos):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_videos, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["full"].shape == r["ref"].shape == (n_videos, 3, 4, 4)
    assert torch.equal(r["full"], r["ref"])          # integer ids: bit-exact, whatever the shard boundaries
    assert r["slowest"] == 11.0
    assert r["seeds"] == [7, 8]


def test_bench_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` launched the way the driver launches N>1 (torchrun, 2 ranks, CPU only here): rank 0
    alone runs and prints ONE JSON line with the contract's keys; the other rank exits 0 without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1"]
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cvivit_encode_frames_per_s" and d["n_gpus"] == 2
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def _train_worker(rank, world_size, port, out):
    """Data-parallel training step on two gloo ranks: each rank runs the (emulated, CPU-executed) CUDA training step on
    its contiguous batch shard; backward() all-reduces the flat gradient bucket."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        torch.set_num_threads(2)
        import phenaki_pytorch_b200 as P
        from oracle import phenaki_oracle as O
        from tests import emu_runtime
        emu_runtime.route_product_to_emulator(emu_runtime.build_emu())
        case = C.TRAIN_CASES["with_critic"]
        torch.manual_seed(case["seed"])
        cvivit, maskgit = P.CViViT(**C.SAMPLE_CVIVIT), P.MaskGit(**case["maskgit"])       # replicated weights
        phenaki = P.Phenaki(cvivit=cvivit, maskgit=maskgit, steps=case["steps"],
                            text_embed_dim=case["maskgit"]["dim_context"]).train()
        ids, ctx = C.train_inputs(case)
        b, n = ids.shape[0], ids[0].numel()
        torch.manual_seed(case["noise_seed"])
        rand_step, u = O.train_draws(b, n, case["steps"])
        lo, hi = S.shard_range(b, rank, world_size)
        draws = {"rand_step": rand_step[lo:hi], "perm": u[lo:hi]}
        loss = phenaki(video_codebook_ids=S.shard_batch(ids), text_embeds=S.shard_batch(ctx),
                       draw_fn=lambda shape, tag: draws[tag])
        loss.backward()
        if rank == 0:
            # reference: DDP semantics = mean over ranks of each rank's own gradient (autograd through the oracle)
            ref = None
            for r in range(world_size):
                a, z = S.shard_range(b, r, world_size)
                sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v)
                      for k, v in maskgit.state_dict().items()}
                tm = O.train_token_mask(rand_step[a:z], u[a:z], case["steps"])
                O.maskgit_train_loss(ids[a:z].reshape(z - a, n), sd, tm, video_patch_shape=case["patch_shape"],
                                     heads=case["maskgit"]["heads"], context=ctx[a:z],
                                     text_mask=torch.any(ctx[a:z] != 0, dim=-1)).backward()
                g = {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
                ref = g if ref is None else {k: ref[k] + g[k] for k in g}
            ref = {k: v / world_size for k, v in ref.items()}
            got = {k: p.grad for k, p in maskgit.named_parameters() if p.grad is not None}
            torch.save(dict(ref=ref, got=got), out)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_training_step_averages_the_gradient_bucket(tmp_path):
    out = str(tmp_path / "train.pt")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert set(r["got"]) <= set(r["ref"]) and len(r["got"]) > 40  # the oracle also differentiates the beta buffers
    for k in r["got"]:
        ref = r["ref"][k]
        if ref.numel() == 0:
            continue
        scale = max(ref.abs().max().item(), 1e-12)
        torch.testing.assert_close(r["got"][k], ref, rtol=1e-3, atol=1e-4 * scale + 1e-7, msg=lambda m, k=k: f"{k}: {m}")
