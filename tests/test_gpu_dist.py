"""Strip-decomposed (multi-GPU) solver against the single-GPU solver on the same links: hierarchy, iteration count, solution.

Two ways to run the strip code:
  * world = 1, the rank is its own neighbour (Comm.single): every strip code path -- IPC slab, halo slots, the standalone
    exchange kernel, the halo push fused into the smoother kernels, the in-kernel cross-rank reductions -- runs on ONE GPU
    (this is what the driver's single-GPU test box executes);
  * world = 2, spawned here when two GPUs are visible (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`).
SURVEY 4: "same problem on 1 vs 2/4/8 ranks must give identical iteration counts"."""
import os
import sys

import numpy as np
import pytest
import torch

import mg2d
from importlib import import_module

pytestmark = pytest.mark.gpu
dmod = import_module("2d_multigrid_b200.dist")


def _problem(L, dev):
    U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=20, seed=1234, device=dev)
    nl = 2 if L <= 256 else 3
    p = mg2d.make_params(L, -0.03, nlevels=nl, block=4, n_null=8, n_smooth=3, n_pre=0, n_post=([4, 2] + [8] * nl)[:nl + 1],
                         smoother="rbgs", null_iters=40, tol=1e-10, max_iters=100)
    rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=f"cuda:{dev}")
    rhs[L // 2 + (L // 2) * L, 0] = 1.0
    return U, p, rhs


def _compare(ref, dmg, x_ref, i_ref, x, info, L):
    worst = 0.0
    for a, b in zip(dmg.LVL[1:], ref.LVL[1:]):
        Db = b.D[a.y0 * a.L:(a.y0 + a.Ly) * a.L] if a.distributed else b.D
        worst = max(worst, float((a.D - Db).abs().max() / b.D.abs().max()))
    lv0 = dmg.LVL[0]
    dx = float((x - x_ref[lv0.y0 * L:(lv0.y0 + lv0.Ly) * L]).abs().max() / x_ref.abs().max())
    return worst, dx


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("use_graph", [False, True])
def test_strip_code_with_self_neighbour_equals_single_gpu(fused, use_graph):
    L = 128
    U, p, rhs = _problem(L, 0)
    ref = mg2d.setup(U, p, init="device")
    x_ref, i_ref = mg2d.solve(ref, rhs=rhs, tol=1e-10, outer="gcr")
    x_ref = x_ref.clone()
    comm = dmod.Comm.single(mg2d.Context(0), torch.device("cuda", 0))
    comm.fused = fused
    dmg = dmod.setup(U, p, comm, min_rows=16)         # levels 128, 32 as strips (halo machinery), level 8 replicated
    assert [d for d, _ in dmg.plan] == [True, True, False]
    x, info = mg2d.solve(dmg, rhs=rhs, tol=1e-10, outer="gcr", use_graph=use_graph)
    worst, dx = _compare(ref, dmg, x_ref, i_ref, x, info, L)
    assert worst < 1e-12, worst
    assert info["converged"] and info["iters"] == i_ref["iters"] and info["executed_iters"] == info["iters"]
    assert dx < 1e-9, dx
    assert info["true_resnorm"] < 1e-10
    assert comm.p2p_errors() == 0
    # the complex64 preconditioner copy runs through the same strip code
    xm, im = mg2d.solve(dmg, rhs=rhs, tol=1e-10, outer="gcr", use_graph=use_graph, precond_dtype="complex64")
    assert im["converged"] and im["true_resnorm"] < 1e-10 and abs(im["iters"] - i_ref["iters"]) <= 1
    # stationary cycle (f_perform_MG) on strips, residual norm reduced inside the kernel
    dmg.LVL[0].phi.zero_()
    xs, is_ = mg2d.solve(dmg, rhs=rhs, tol=1e-6, max_iters=60)
    ref.LVL[0].phi.zero_()
    xr, ir = mg2d.solve(ref, rhs=rhs, tol=1e-6, max_iters=60)
    assert is_["iters"] == ir["iters"] and is_["converged"] == ir["converged"]
    ref.close(); dmg.close()


def test_ntl_cycle_on_strips_with_replicated_copy_levels():
    """f_MG_ntl (S6/modules_main.h:386-439) on the strip code: the quadrant-shifted copies live on the two coarsest levels,
    which the plan replicates; fine levels are strips.  Same iteration count and solution as the single-GPU NTL solve; a plan
    that would keep the copy level striped is refused."""
    L = 128
    U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=20, seed=1234, device=0)
    p = mg2d.make_params(L, -0.03, nlevels=3, block=4, n_null=8, n_smooth=2, smoother="rbgs", null_iters=40, tol=1e-10, max_iters=100,
                         ntl=True, n_copies=4)
    rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device="cuda:0")
    rhs[L // 2 + (L // 2) * L, 0] = 1.0
    ref = mg2d.setup(U, p, init="device")
    x_ref, i_ref = mg2d.solve(ref, rhs=rhs, tol=1e-10)
    x_ref = x_ref.clone()
    assert i_ref["converged"] and len(i_ref["ntl_weights"]) == i_ref["iters"]
    comm = dmod.Comm.single(mg2d.Context(0), torch.device("cuda", 0))
    dmg = dmod.setup(U, p, comm, min_rows=16)
    assert [d for d, _ in dmg.plan] == [True, True, False, False]
    for use_graph in (False, True):
        x, info = mg2d.solve(dmg, rhs=rhs, tol=1e-10, use_graph=use_graph)
        assert info["converged"] and info["iters"] == i_ref["iters"]
        assert float((x - x_ref).abs().max() / x_ref.abs().max()) < 1e-9
        wg, wr = np.array(info["ntl_weights"]), np.array(i_ref["ntl_weights"])
        assert np.allclose(wg[:5], wr[:5], rtol=1e-6, atol=1e-9)         # (the min-res weights of late cycles are fixed by residuals
        assert np.allclose(wg, wr, rtol=1e-3, atol=1e-4)                 #  of order 1e-9: their last digits follow rounding)
    assert comm.p2p_errors() == 0
    with pytest.raises(NotImplementedError):
        dmod.DistMG(mg2d.make_params(L, -0.03, nlevels=2, block=4, n_null=8, smoother="rbgs", ntl=True), comm, min_rows=16)
    ref.close(); dmg.close()


def _gs_problem(L, dev):
    U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=20, seed=1234, device=dev)
    p = mg2d.make_params(L, 0.02, nlevels=2, block=2, n_null=2, n_smooth=2, smoother="gs", null_iters=16, tol=1e-10, max_iters=200)
    rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=f"cuda:{dev}")
    rhs[L // 2 + (L // 2) * L, 0] = 1.0
    return U, p, rhs


def test_lexicographic_gs_on_a_strip_with_self_neighbour():
    """The reference's own smoother (gs_flag = 1, S6/level.h:100-128) through the strip kernel mg2d_relax_gs_strip with the rank
    as its own neighbour: global fronts, the periodic upper neighbour of the last row delivered through the halo buffer and its
    progress flag.  One sweep is bit-identical to the single-GPU wavefront kernel; the whole solve needs the same iterations."""
    L = 64
    U, p, rhs = _gs_problem(L, 0)
    ref = mg2d.setup(U, p, init="device")
    comm = dmod.Comm.single(mg2d.Context(0), torch.device("cuda", 0))
    dmg = dmod.setup(U, p, comm, min_rows=8)
    assert [d for d, _ in dmg.plan] == [True, True, True]
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    for lvl in (0, 1):
        a, b = ref.LVL[lvl], dmg.LVL[lvl]
        assert float((a.D - b.D).abs().max()) < 1e-12 * float(a.D.abs().max())
        b.D.copy_(a.D); b.D0inv = None; a.D0inv = None
        v = torch.randn((a.S, a.n, 2), generator=g, dtype=torch.float64, device="cuda")
        v = torch.view_as_complex(v).contiguous()
        rr = torch.view_as_complex(torch.randn((a.S, a.n, 2), generator=g, dtype=torch.float64, device="cuda")).contiguous()
        va, vb = v.clone(), v.clone()
        a.relax(2, phi=va, r=rr, smoother="gs")
        b.relax(2, phi=vb, r=rr, smoother="gs")
        assert torch.equal(va, vb), lvl
    ref2 = mg2d.setup(U, p, init="device")
    x_ref, i_ref = mg2d.solve(ref2, rhs=rhs, tol=1e-10)
    dmg2 = dmod.setup(U, p, comm, min_rows=8)
    x, info = mg2d.solve(dmg2, rhs=rhs, tol=1e-10)
    assert info["converged"] and info["iters"] == i_ref["iters"]
    assert float((x - x_ref).abs().max() / x_ref.abs().max()) < 1e-9
    assert comm.p2p_errors() == 0


def _gs_worker(rank, world, port, L, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    comm = dmod.init(world, rank, rank)
    U, p, rhs = _gs_problem(L, rank)
    res = {}
    if rank == 0:
        ref = mg2d.setup(U, p, init="device")
        x_ref, i_ref = mg2d.solve(ref, rhs=rhs, tol=1e-10)
        x_ref = x_ref.clone()
    dmg = dmod.setup(U, p, comm, min_rows=8)
    x, info = mg2d.solve(dmg, rhs=dmg.scatter_field(rhs), tol=1e-10)
    if rank == 0:
        lv0 = dmg.LVL[0]
        dx = float((x - x_ref[lv0.y0 * L:(lv0.y0 + lv0.Ly) * L]).abs().max() / x_ref.abs().max())
        res = dict(dx=dx, iters=info["iters"], ref_iters=i_ref["iters"], conv=info["converged"], errors=comm.p2p_errors(),
                   plan=[d for d, _ in dmg.plan])
        out.put(res)
    torch.cuda.synchronize()
    torch.distributed.barrier()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_lexicographic_gs_on_two_ranks_equals_single_gpu():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_gs_worker, args=(r, 2, 29537, 64, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    r = out.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
    assert all(r["plan"]) and r["conv"] and r["iters"] == r["ref_iters"] and r["dx"] < 1e-9 and r["errors"] == 0, r


def _worker(rank, world, port, L, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    comm = dmod.init(world, rank, rank)
    U, p, rhs = _problem(L, rank)
    res = {}
    if rank == 0:
        ref = mg2d.setup(U, p, init="device")
        x_ref, i_ref = mg2d.solve(ref, rhs=rhs, tol=1e-10, outer="gcr")
        x_ref = x_ref.clone()
    for use_graph in (False, True):
        dmg = dmod.setup(U, p, comm, min_rows=16)
        x, info = mg2d.solve(dmg, rhs=dmg.scatter_field(rhs), tol=1e-10, outer="gcr", use_graph=use_graph)
        x2, info = mg2d.solve(dmg, rhs=dmg.scatter_field(rhs), tol=1e-10, outer="gcr", use_graph=use_graph)
        if rank == 0:
            worst, dx = _compare(ref, dmg, x_ref, i_ref, x2, info, L)
            res[use_graph] = dict(worst=worst, dx=dx, iters=info["iters"], ref_iters=i_ref["iters"], true=info["true_resnorm"],
                                  executed=info["executed_iters"], errors=comm.p2p_errors(), plan=[d for d, _ in dmg.plan])
        torch.distributed.barrier()
    if rank == 0:
        out.put(res)
    torch.cuda.synchronize()
    torch.distributed.barrier()
    os._exit(0)          # NCCL teardown under live graphs hangs


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_single_gpu():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, 256, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = out.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
    for ug, r in res.items():
        assert r["plan"][0] and r["worst"] < 1e-12 and r["dx"] < 1e-9 and r["true"] < 1e-10, (ug, r)
        assert r["iters"] == r["ref_iters"] == r["executed"] and r["errors"] == 0, (ug, r)
