"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous slices, per-rank filtering, host
gather; the result must equal the single-rank run.  The per-rank compute here is the oracle (there
is no GPU in this test) - what is under test is the partitioning and the gather, which is all the
N>1 path adds (no collective on the data path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest as cf
from asif_b200.sharding import all_slices, slice_bounds


def test_slice_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 1000, 10_000_001):
        for w in (1, 2, 3, 4, 8):
            sl = all_slices(n, w)
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert max(b - a for a, b in sl) - min(b - a for a, b in sl) <= -(-n // w)
    with pytest.raises(ValueError):
        slice_bounds(10, 2, 2)


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, cf.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyref
    O = pyref.OracleLib()
    x, ud = cf.c2_inputs(n, seed=4242)
    lo, hi = slice_bounds(n, world, rank)
    u, relax, rc = O.filter_batch(2, x[lo:hi], ud[lo:hi], cf.C2_TB_OPTS)
    per = -(-n // world)
    pad_u = torch.zeros(per, dtype=torch.float64)
    pad_rc = torch.zeros(per, dtype=torch.int32)
    pad_u[: hi - lo] = torch.from_numpy(u[:, 0].copy())
    pad_rc[: hi - lo] = torch.from_numpy(rc.copy())
    gu = [torch.zeros(per, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    grc = [torch.zeros(per, dtype=torch.int32) for _ in range(world)] if rank == 0 else None
    dist.gather(pad_u, gu, dst=0)
    dist.gather(pad_rc, grc, dst=0)
    # the timing reduction bench.py uses: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert t.item() == world
        U = torch.cat([g[: b - a] for g, (a, b) in zip(gu, all_slices(n, world))]).numpy()
        RC = torch.cat([g[: b - a] for g, (a, b) in zip(grc, all_slices(n, world))]).numpy()
        np.savez(out_path, u=U, rc=RC)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(oracle, tmp_path):
    n, world = 301, 2
    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    g = np.load(out)
    x, ud = cf.c2_inputs(n, seed=4242)
    u, relax, rc = oracle.filter_batch(2, x, ud, cf.C2_TB_OPTS)
    assert np.array_equal(g["rc"], rc)
    assert np.array_equal(g["u"], u[:, 0])
