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
