"""world_size-2 gloo test (CPU) of the multi-GPU host logic: block exchange and the
region-major segment plan K3 consumes.  No kernels involved."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pangenome_b200 import multigpu
    n_sub, part_cap = 4, 8
    # bucket (owner d, region b) of rank r holds c = 1 + (r + d + b) % 5 records tagged (r, d, b, i)
    counts = torch.zeros(world, n_sub, dtype=torch.int64)
    rec = torch.full((world, n_sub, part_cap, 2), -1, dtype=torch.int64)
    for d in range(world):
        for b in range(n_sub):
            c = 1 + (rank + d + b) % 5
            counts[d, b] = c
            for i in range(c):
                rec[d, b, i, 0] = rank * 1000 + d * 100 + b * 10 + i
                rec[d, b, i, 1] = 7
    recv_counts = multigpu.exchange_blocks(counts, world)
    recv = multigpu.exchange_blocks(rec.view(world, -1), world).view(-1, 2)
    seg_off, seg_cnt = multigpu.segment_plan(recv_counts, part_cap)
    got = []
    for o, c in zip(seg_off.tolist(), seg_cnt.tolist()):
        got.append(recv[o:o + c, 0].tolist())
    want = []
    for b in range(n_sub):              # region-major: all sources of region 0, then region 1, ...
        for s in range(world):
            c = 1 + (s + rank + b) % 5
            want.append([s * 1000 + rank * 100 + b * 10 + i for i in range(c)])
    out[rank] = (got == want)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_and_segment_plan_gloo():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, 29571, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)) and len(out) == world


def test_log2_exact():
    from pangenome_b200 import multigpu
    assert [multigpu.log2_exact(n) for n in (1, 2, 4, 8)] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        multigpu.log2_exact(3)


def _gv_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pangenome_b200 import multigpu
    t = torch.arange(rank * 10, rank * 10 + rank + 1, dtype=torch.int64)       # rank r contributes r+1 items
    cat, sizes = multigpu.gather_varlen(t, world)
    out[rank] = (cat.tolist() == [0, 10, 11] and sizes == [1, 2])
    e, _ = multigpu.gather_varlen(torch.zeros(0, dtype=torch.int64), world)
    out[rank] = out[rank] and e.numel() == 0
    dist.barrier()
    dist.destroy_process_group()


def test_gather_varlen_gloo():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_gv_worker, args=(world, 29573, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)) and len(out) == world
