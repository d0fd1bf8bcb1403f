"""N>1 host logic on CPU: world_size-2 gloo processes exercise the strip plan, the halo exchange ordering (with
2 ranks both neighbours are the same peer), all-reduce and the strip all-gather used for coarse-level
agglomeration.  The CUDA side of the same code path is covered by tests/test_gpu_dist.py (self-neighbour on one GPU, 2 ranks when two are visible)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mg2d
from importlib import import_module

dmod = import_module("2d_multigrid_b200.dist")


def test_plan_strips():
    p = mg2d.make_params(4096, 0.0, nlevels=4, block=4, n_null=8, smoother="rbgs")
    assert dmod.plan_strips(p, 8, 32) == [(True, 512), (True, 128), (True, 32), (False, 64), (False, 16)]
    assert dmod.plan_strips(p, 2, 32) == [(True, 2048), (True, 512), (True, 128), (True, 32), (False, 16)]
    q = mg2d.make_params(64, 0.0, nlevels=2, block=2, smoother="rbgs")
    assert dmod.plan_strips(q, 2, 8) == [(True, 32), (True, 16), (True, 8)]
    assert dmod.plan_strips(q, 2, 32)[1:] == [(False, 32), (False, 16)]            # replicated from level 1 on
    assert dmod.plan_strips(q, 3, 8)[0] == (False, 64)                              # 64 rows do not split in 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = dmod.Comm(world, rank)
        Lx, Ly, n = 6, 4, 2
        # global field value = global row index * 100 + x * 10 + dof
        y0 = rank * Ly
        rows = torch.arange(y0, y0 + Ly, dtype=torch.float64).repeat_interleave(Lx)
        t = (rows * 100)[:, None] + (torch.arange(Lx, dtype=torch.float64).repeat(Ly) * 10)[:, None] + torch.arange(n, dtype=torch.float64)[None]
        t = t.to(torch.complex128)
        lo, hi = comm.exchange_rows(t, Lx, Ly, n)
        Ltot = world * Ly
        want_lo = ((y0 - 1) % Ltot) * 100
        want_hi = ((y0 + Ly) % Ltot) * 100
        ok = bool(torch.all(lo.real[:, 0] == want_lo + torch.arange(Lx) * 10)) and bool(torch.all(hi.real[:, 0] == want_hi + torch.arange(Lx) * 10))
        # batched
        tb = torch.stack([t, t + 1000])
        lo2, hi2 = comm.exchange_rows(tb, Lx, Ly, n, nvec=2)
        ok = ok and lo2.shape == (2, Lx, n) and bool(torch.all(lo2[1].real[:, 0] == want_lo + 1000 + torch.arange(Lx) * 10))
        ok = ok and bool(torch.all(hi2[0].real[:, 1] == want_hi + 1 + torch.arange(Lx) * 10))
        s = torch.tensor([float(rank + 1), 2.0], dtype=torch.float64)
        comm.allreduce(s)
        ok = ok and s.tolist() == [sum(range(1, world + 1)), 2.0 * world]
        m = torch.tensor([float(rank)], dtype=torch.float64)
        comm.allreduce(m, "max")
        ok = ok and m.item() == world - 1
        full = torch.zeros((world * Ly * Lx, n), dtype=torch.complex128)
        comm.allgather(full, t)
        ok = ok and bool(torch.all(full.real[:, 0].reshape(world * Ly, Lx)[:, 0] == torch.arange(world * Ly, dtype=torch.float64) * 100))
        ok = ok and dmod.bcast_float(comm, 3.5 if rank == 0 else -1.0) == 3.5
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_exchange(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
