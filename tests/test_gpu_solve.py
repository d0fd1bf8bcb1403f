"""End-to-end parity: the GPU solver against the oracle on the BASELINE configs -- identical outer iteration
counts, residual histories, near-null vectors and solutions -- plus size-independent properties at full size."""
import numpy as np
import pytest
import torch

import mg2d
from oracle import mg_oracle as O
from oracle import scalar_s2 as S2

pytestmark = pytest.mark.gpu


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def hist_close(hg, ho, rtol=1e-7, floor=5e-15):
    """Residual histories agree to rtol, down to the rounding floor of the residual itself: a relative norm of
    1e-13 is only defined to ~eps/1e-13 ~ 1e-3 relative, i.e. ~1e-16 absolute."""
    return all(abs(a - b) <= rtol * b + floor for a, b in zip(hg, ho))


def _run_both(L, m, U, stencil="wilson", nlevels=2, n_smooth=3, smoother="gs", ntl=False, n_copies=4, null_iters=40,
              max_iters=300, block=2, n_null=None, tol=1e-13, **kw):
    nds = None if n_null is None else (2 * n_null if stencil == "wilson" else n_null)
    po = O.Params(L=L, num_iters=n_smooth, block=block, m=m, nlevels=nlevels, stencil=stencil, smoother=smoother,
                  t_flag=int(ntl), n_copies=n_copies, null_iters=null_iters, max_iters=max_iters, n_dof_scale=nds,
                  res_threshold=tol)
    LVLo, NTLo, io = O.run_reference_flow(po, U)
    p = mg2d.make_params(L, m, stencil=stencil, nlevels=nlevels, block=block, n_null=n_null, n_smooth=n_smooth,
                         smoother=smoother, ntl=ntl, n_copies=n_copies, null_iters=null_iters, max_iters=max_iters, tol=tol, **kw)
    mg, ig = mg2d.run_reference_flow(p, T(U))
    return LVLo, io, mg, ig


# config 2: 2D U(1) Wilson 64x64, adaptive 3-level MG (levels 64/32/16), fp64, the reference's GS smoother
def test_config2_wilson64_three_level_gs():
    L = 64
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 32.0, sweeps=30))
    LVLo, io, mg, ig = _run_both(L, 0.02, U, nlevels=2, smoother="gs", null_iters=60)
    assert io["converged"] and ig["converged"]
    assert ig["iters"] == io["iters"]                                   # identical outer iteration count
    assert ig["resnorms"][-1] < 1e-13
    assert hist_close(ig["resnorms"], io["resnorms"])
    assert rel(mg.LVL[0].phi, LVLo[0].phi) < 1e-9
    for l in range(2):
        assert rel(mg.LVL[l].phi_null, LVLo[l].phi_null) < 1e-8


@pytest.mark.parametrize("smoother", ["jacobi", "rbgs", "mr"])
def test_wilson32_other_smoothers(smoother):
    L = 32
    U = O.gauge_gaussian(L, 0.3)
    LVLo, io, mg, ig = _run_both(L, 0.05, U, nlevels=2, smoother=smoother, n_smooth=3, max_iters=400)
    assert ig["iters"] == io["iters"] and ig["converged"] == io["converged"]
    assert hist_close(ig["resnorms"], io["resnorms"])


def test_laplace32_gauged():
    L = 32
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=30))
    LVLo, io, mg, ig = _run_both(L, 0.05, U, stencil="laplace", nlevels=2, smoother="gs")
    assert ig["converged"] and ig["iters"] == io["iters"]
    assert rel(mg.LVL[0].phi, LVLo[0].phi) < 1e-9


# config 3 (scaled to what the oracle finishes in seconds): non-telescoping cycle with min-res weights
@pytest.mark.parametrize("stencil,n_copies", [("wilson", 4), ("wilson", 2), ("laplace", 4), ("wilson", 1)])
def test_ntl_minres(stencil, n_copies):
    L = 32
    U = O.gauge_gaussian(L, 0.3)
    LVLo, io, mg, ig = _run_both(L, 0.05, U, stencil=stencil, nlevels=3, smoother="gs", ntl=True, n_copies=n_copies)
    assert ig["converged"] and ig["iters"] == io["iters"]
    # the first cycles (before chaotic amplification of rounding) agree to high accuracy, weights included
    for a, b in zip(ig["resnorms"][:5], io["resnorms"][:5]):
        assert abs(a / b - 1) < 1e-8
    assert np.max(np.abs(ig["ntl_weights"][0][:n_copies] - io["ntl_weights"][0][:n_copies])) < 1e-8


# config 3 at full size (2D Wilson 256x256 near-critical, non-telescoping cycle, the reference's GS smoother): the numpy
# oracle needs minutes for this, so the GPU run is checked through size-independent properties instead
def test_config3_wilson256_ntl_fullsize():
    L = 256
    th = mg2d.gauge.quenched_phases(L, 32.0, sweeps=40, device="cuda")
    U = torch.exp(1j * th).to(torch.complex128)
    p = mg2d.make_params(L, -0.01, nlevels=3, block=2, n_smooth=3, smoother="gs", ntl=True, n_copies=4, tol=1e-10, max_iters=400)
    mg = mg2d.setup(U, p)                                                 # reference-compatible random start (S6/mgrid_ntl.cpp:38-48)
    assert max(mg.info["ortho_worst"]) < 1e-12                            # f_check_ortho on every level and every NTL copy
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    lo = p.nlevels - 1
    for q in range(4):                                                    # tests 1 and 2 of S6/tests.h for each copy / quadrant
        nt, bot = mg.NTL[lo][q], mg.NTL[p.nlevels][q]
        vc = torch.randn((bot.S, bot.n), generator=g, dtype=torch.float64, device="cuda").to(torch.complex128)
        f1 = torch.zeros((nt.S, nt.n), dtype=torch.complex128, device="cuda")
        nt.prolongation(f1, vc, q + 1)
        c1 = torch.empty_like(vc)
        nt.restriction(c1, f1, q + 1)
        assert float((c1 - vc).abs().max()) < 1e-12
        f2 = torch.empty_like(f1)
        mg.LVL[lo].apply_D(f2, f1)
        nt.restriction(c1, f2, q + 1)
        c2 = torch.empty_like(vc)
        bot.apply_D(c2, vc)
        assert float((c1 - c2).abs().max()) < 1e-11
    x, info = mg2d.solve(mg)
    assert info["converged"] and info["resnorms"][-1] < 1e-10
    assert mg.LVL[0].get_residue_mag() < 1e-10                            # true residual, recomputed
    w = np.array(info["ntl_weights"][-1])
    assert np.all(np.isfinite(w)) and abs(w.sum()) > 0.05                 # min-res weights are in use


def test_config3_matches_reference_run():
    """BASELINE configs[2] at its stated size against the reference ITSELF: tests/golden/big_s6_wilson256_ntl4_nearcrit.npz was
    written by the unmodified S6 program (`./a.out 256 3 2 1 -0.01 3 1 4`, 102 cycles, 6 min on one core; make_golden.py
    --large).  Same links, same mt19937 start, same lexicographic GS / 500-sweep near-null vectors / 4-copy min-res cycle on
    the GPU: identical iteration count, printed residual history, NTL weights and final phi."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "big_s6_wilson256_ntl4_nearcrit.npz"))
    L, beta, sweeps, seed = int(z["gauge"][0]), float(z["gauge"][1]), int(z["gauge"][2]), int(z["gauge"][3])
    theta = O.gauge_quenched_phases(L, beta, sweeps=sweeps, seed=seed)
    assert float(np.sum(theta * np.arange(1, theta.size + 1).reshape(theta.shape))) == pytest.approx(float(z["theta_checksum"]), rel=1e-12)
    a = [str(x) for x in z["argv"]]
    p = mg2d.make_params(L, float(a[4]), nlevels=int(a[5]), block=int(a[2]), n_smooth=int(a[1]), smoother="gs", ntl=True, n_copies=int(a[7]))
    mg, info = mg2d.run_reference_flow(p, T(O.gauge_from_phases(theta)))
    assert info["converged"] and info["iters"] == int(z["iters"]) == 102
    printed = z["resmag"]
    for k in range(len(printed)):
        assert abs(info["resnorms"][k] - printed[k]) <= 5e-4 * printed[k] + 5e-15, (k, info["resnorms"][k], printed[k])
    assert rel(mg.LVL[0].phi, z["phi_final"]) < 1e-9
    assert np.max(np.abs(mg.LVL[0].phi_null[:4096].cpu().numpy() - z["null0_head"])) < 1e-8
    w = np.array(info["ntl_weights"])
    assert np.max(np.abs(w[:5] - z["ntl_weights"][:5])) < 2e-3 * np.max(np.abs(z["ntl_weights"][:5]))


# config 4 shape at a size the oracle can follow: 8 null vectors, 4x4 aggregates
def test_config4_shape_small():
    L = 64
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=30))
    LVLo, io, mg, ig = _run_both(L, 0.0, U, nlevels=2, block=4, n_null=8, smoother="rbgs", n_smooth=2, null_iters=40, tol=1e-10)
    assert ig["converged"] and ig["iters"] == io["iters"]
    assert rel(mg.LVL[0].phi, LVLo[0].phi) < 1e-8


@pytest.mark.parametrize("post", [4, [4, 2, 8]])
def test_v04_cycle_chiral_transfer_parity(post):
    """The bench cycle shape at a size the oracle follows: V(0,4) red-black cycle inside FGCR(8), 4x4 aggregates,
    8 null vectors, chirality-compacted transfers and the overwrite-prolongation shortcut: same iteration count
    and solution as the oracle (which uses the dense projector and no shortcut)."""
    L = 64
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=30))
    b = np.zeros((L * L, 2), dtype=complex)
    b[L // 2 + (L // 2) * L, 0] = 1.0
    po = O.Params(L=L, num_iters=4, n_pre=0, n_post=post, block=4, m=-0.02, nlevels=2, null_iters=40, smoother="rbgs", n_dof_scale=16)
    LVLo, NTLo = O.build_reference_problem(po, U)
    O.compute_near_null(LVLo, NTLo, po, 1)
    xo, io = O.gcr_MG(LVLo, NTLo, po, b, tol=1e-10, restart=8)
    p = mg2d.make_params(L, -0.02, nlevels=2, block=4, n_null=8, n_smooth=4, n_pre=0, n_post=post, smoother="rbgs", null_iters=40)
    mg = mg2d.setup(T(U), p)
    assert mg.LVL[0].phi_null_c is not None and mg.LVL[0].matrix_free
    for use_graph in (False, True):
        x, ig = mg2d.solve(mg, rhs=T(b), tol=1e-10, outer="gcr", restart=8, use_graph=use_graph)
        assert ig["iters"] == io["iters"] and ig["true_resnorm"] < 1e-10
        assert hist_close(ig["resnorms"], io["resnorms"], rtol=1e-6)
        assert rel(x, xo) < 1e-7
    # chiral kernels == dense kernels
    lv = mg.LVL[0]
    rng = np.random.default_rng(1)
    vf = T(rng.normal(size=(L * L, 2)) + 1j * rng.normal(size=(L * L, 2)))
    c1, c2 = torch.empty((256, 16), dtype=torch.complex128, device="cuda"), torch.empty((256, 16), dtype=torch.complex128, device="cuda")
    lv.restriction(c1, vf, 1)
    pc, lv.phi_null_c = lv.phi_null_c, None
    lv.restriction(c2, vf, 1)
    f1, f2, f3 = vf.clone(), vf.clone(), torch.full_like(vf, 7.0)
    lv.prolongation(f2, c1.clone(), 1)
    lv.phi_null_c = pc
    lv.prolongation(f1, c1.clone(), 1)
    lv.prolongation(f3, c1.clone(), 1, accumulate=False)
    assert float((c1 - c2).abs().max()) < 1e-13 and float((f1 - f2).abs().max()) < 1e-13
    assert float((f3 - (f1 - vf)).abs().max()) < 1e-13


def test_lowrank_level1_solve_parity():
    """The bench cycle with the level-1 sweeps on the rank-4 factors of the hopping blocks (mg2d_relax_rb_lr): same
    iteration count, residual history and solution as the oracle (dense blocks) and as the dense GPU path, eager and from
    the iteration graphs, complex128 and with the complex64 preconditioner copy."""
    L = 64
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=30))
    b = np.zeros((L * L, 2), dtype=complex)
    b[L // 2 + (L // 2) * L, 0] = 1.0
    po = O.Params(L=L, num_iters=4, n_pre=0, n_post=[4, 2, 8], block=4, m=-0.02, nlevels=2, null_iters=40, smoother="rbgs", n_dof_scale=16)
    LVLo, NTLo = O.build_reference_problem(po, U)
    O.compute_near_null(LVLo, NTLo, po, 1)
    xo, io = O.gcr_MG(LVLo, NTLo, po, b, tol=1e-10, restart=8)
    p = mg2d.make_params(L, -0.02, nlevels=2, block=4, n_null=8, n_smooth=4, n_pre=0, n_post=[4, 2, 8], smoother="rbgs", null_iters=40)
    mg = mg2d.MG(p)
    mg.persistent_sites = 0                     # (16^2 sites would otherwise take the one-launch dense path)
    mg.init_reference_fields()
    mg.set_gauge(T(U))
    mg2d.compute_near_null(mg)
    assert mg.LVL[1].lr_rank == 4 and mg.LVL[1].F is not None and mg.LVL[2].lr_rank == 0
    res = {}
    for lowrank in (True, False):
        mg.lowrank = lowrank
        for use_graph in (False, True):
            x, ig = mg2d.solve(mg, rhs=T(b), tol=1e-10, outer="gcr", restart=8, use_graph=use_graph)
            assert ig["iters"] == io["iters"] and ig["true_resnorm"] < 1e-10
            assert hist_close(ig["resnorms"], io["resnorms"], rtol=1e-6)
            assert rel(x, xo) < 1e-7
            res[(lowrank, use_graph)] = x.clone()
    assert float((res[(True, False)] - res[(False, False)]).abs().max()) < 1e-9 * float(res[(False, False)].abs().max())
    mg.lowrank = True
    x32, i32 = mg2d.solve(mg, rhs=T(b), tol=1e-10, outer="gcr", restart=8, precond_dtype="complex64")
    assert mg.info["single"].LVL[1].F is not None and mg.info["single"].LVL[1].F.dtype == torch.complex64
    assert i32["true_resnorm"] < 1e-10 and abs(i32["iters"] - io["iters"]) <= 1 and rel(x32, xo) < 1e-7


def test_gcr_outer_and_graph():
    L = 32
    U = O.gauge_gaussian(L, 0.3)
    b = np.zeros((L * L, 2), dtype=complex)
    b[L // 2 + (L // 2) * L, 0] = 1.0
    po = O.Params(L=L, num_iters=2, block=2, m=0.02, nlevels=2, null_iters=40, smoother="rbgs")
    LVLo, NTLo = O.build_reference_problem(po, U)
    O.compute_near_null(LVLo, NTLo, po, 1)
    xo, io = O.gcr_MG(LVLo, NTLo, po, b, tol=1e-10, restart=4)
    p = mg2d.make_params(L, 0.02, nlevels=2, n_smooth=2, smoother="rbgs", null_iters=40)
    mg = mg2d.setup(T(U), p)
    assert io["iters"] > 4                                                       # (the restart cycle wraps at least once)
    xs = {}
    for lazy in (True, False):       # raw stored directions + one solution update per restart cycle vs the textbook updates
        mg.lazy_gcr = lazy
        for use_graph in (False, True):
            x, ig = mg2d.solve(mg, rhs=T(b), tol=1e-10, outer="gcr", restart=4, use_graph=use_graph, check_every=3)
            assert ig["iters"] == io["iters"] and ig["resnorms"][io["iters"] - 1] < 1e-10      # identical outer iteration count
            assert all(r >= 1e-10 for r in ig["resnorms"][:io["iters"] - 1])          # same first crossing
            assert ig["true_resnorm"] < 1e-10                                        # final TRUE residual
            assert rel(x, xo) < 1e-6
            x1, i1 = mg2d.solve(mg, rhs=T(b), tol=1e-10, outer="gcr", restart=4, use_graph=use_graph, check_every=1)
            assert i1["iters"] == i1["executed_iters"] == io["iters"] and hist_close(i1["resnorms"], io["resnorms"], rtol=1e-6)
            assert rel(x1, xo) < 1e-7 and i1["true_resnorm"] < 1e-10
            xs[(lazy, use_graph)] = x1.clone()
    assert float((xs[(True, True)] - xs[(False, False)]).abs().max()) < 1e-9 * float(xs[(False, False)].abs().max())
    mg.lazy_gcr = True
    # stationary cycle from a graph equals the eager cycle
    m1 = mg2d.setup(T(U), p)
    x1, i1 = mg2d.solve(m1)
    m2 = mg2d.setup(T(U), p)
    x2, i2 = mg2d.solve(m2, use_graph=True)
    assert i1["iters"] == i2["iters"] and float((x1 - x2).abs().max()) < 1e-10


def test_kcycle_parity():
    """K-cycle (2 FGCR steps per coarse level instead of one recursive visit) against the oracle's K-cycle:
    identical iteration counts and residual histories, stationary and as the preconditioner of the outer FGCR, eager and
    from the iteration graphs."""
    L = 32
    U = O.gauge_gaussian(L, 0.3)
    b = np.zeros((L * L, 2), dtype=complex)
    b[L // 2 + (L // 2) * L, 0] = 1.0
    po = O.Params(L=L, num_iters=2, block=2, m=-0.02, nlevels=3, null_iters=40, smoother="rbgs", cycle="K")
    LVLo, NTLo = O.build_reference_problem(po, U)
    O.compute_near_null(LVLo, NTLo, po, 1)
    xo, io = O.gcr_MG(LVLo, NTLo, po, b, tol=1e-10, restart=8)
    pv = O.Params(L=L, num_iters=2, block=2, m=-0.02, nlevels=3, null_iters=40, smoother="rbgs")
    LVLv, NTLv = O.build_reference_problem(pv, U)
    O.compute_near_null(LVLv, NTLv, pv, 1)
    _, iv = O.gcr_MG(LVLv, NTLv, pv, b, tol=1e-10, restart=8)
    assert io["iters"] < iv["iters"]                                  # the K-cycle is the stronger preconditioner
    p = mg2d.make_params(L, -0.02, nlevels=3, n_smooth=2, smoother="rbgs", null_iters=40, cycle="K")
    mg = mg2d.setup(T(U), p)
    for use_graph in (False, True):
        x, ig = mg2d.solve(mg, rhs=T(b), tol=1e-10, outer="gcr", restart=8, use_graph=use_graph)
        assert ig["iters"] == io["iters"] and hist_close(ig["resnorms"], io["resnorms"], rtol=1e-6)
        assert ig["true_resnorm"] < 1e-10 and rel(x, xo) < 1e-7
    # stationary K-cycle iteration (reference-compatible random start)
    LVLs, NTLs = O.build_reference_problem(po, U)
    O.compute_near_null(LVLs, NTLs, po, 1)
    po.res_threshold, po.max_iters = 1e-10, 300
    is_ = O.perform_MG(LVLs, NTLs, po)
    m2 = mg2d.setup(T(U), p)
    x2, i2 = mg2d.solve(m2, tol=1e-10, max_iters=300)
    assert i2["iters"] == is_["iters"] and hist_close(i2["resnorms"], is_["resnorms"], rtol=1e-6)


def test_per_level_block_sizes():
    """Per-level aggregate sizes (S5L/setup.h:2-10 `block_x[level]`): 4x4 aggregates on the fine lattice, 2x2 above
    (levels 32/8/4) against the oracle with the same list: coarse operators, iteration count and solution."""
    L = 32
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=30))
    po = O.Params(L=L, num_iters=3, block=[4, 2], m=0.0, nlevels=2, null_iters=40, smoother="rbgs", res_threshold=1e-12,
                  n_dof_scale=8)
    assert po.size == [32, 8, 4]
    LVLo, NTLo, io = O.run_reference_flow(po, U)
    p = mg2d.make_params(L, 0.0, nlevels=2, block=[4, 2], n_null=4, n_smooth=3, smoother="rbgs", null_iters=40, tol=1e-12)
    assert p.size == [32, 8, 4] and p.blocks == [4, 2]
    mg, ig = mg2d.run_reference_flow(p, T(U))
    for l in (1, 2):
        assert rel(mg2d.D_to_reference_layout(mg.LVL[l].D), LVLo[l].D) < 1e-8
    assert ig["converged"] and ig["iters"] == io["iters"]
    assert hist_close(ig["resnorms"], io["resnorms"], rtol=1e-6)
    assert rel(mg.LVL[0].phi, LVLo[0].phi) < 1e-8
    with pytest.raises(ValueError):
        mg2d.make_params(L, 0.0, nlevels=2, block=[4])               # one entry per coarsening step
    with pytest.raises(ValueError):
        mg2d.make_params(L, 0.0, nlevels=2, block=[3, 2])            # must divide the lattice


def test_error_spectrum_on_device_matches_oracle_history():
    """SURVEY 8f N4 (NB/2_spectral_analysis_solution.ipynb cell 5): the 2D FFT of phi_k - phi* per iteration, computed on the
    device from the solver's own recorded iterates, against numpy's FFT of the oracle's iterates of the same run; and the
    smoothing property it is there to show: relaxation alone damps the high-frequency error faster than the low-frequency one."""
    L = 16
    U = O.gauge_gaussian(L, 0.3)
    po = O.Params(L=L, num_iters=2, block=2, m=0.1, nlevels=2, null_iters=40, res_threshold=1e-12, max_iters=100)
    LVLo, NTLo = O.build_reference_problem(po, U)
    O.compute_near_null(LVLo, NTLo, po, 1)
    io = O.perform_MG(LVLo, NTLo, po, record_phi=True)
    p = mg2d.make_params(L, 0.1, nlevels=2, n_smooth=2, null_iters=40, tol=1e-12, max_iters=100)
    mg, ig = mg2d.run_reference_flow(p, T(U), record_phi=True)
    assert ig["iters"] == io["iters"]
    star_g, star_o = mg.LVL[0].phi, LVLo[0].phi
    for k in (0, 1, 3):
        sg = mg2d.diagnostics.error_spectrum(ig["phi_hist"][k].cuda(), star_g, L)
        eo = (io["phi_hist"][k] - star_o).reshape(L, L, 2).transpose(2, 0, 1)
        so = np.abs(np.fft.fft2(eo))
        assert sg.is_cuda and rel(sg, so) < 1e-8
    # the notebooks' summary (largest low- / high-frequency error amplitude) after 6 Gauss-Seidel sweeps on D e = 0 from a random
    # error: device relaxation + device FFT against the oracle's relaxation + numpy FFT
    import copy
    lv = mg.LVL[0]
    rng = np.random.default_rng(5)
    e0 = rng.normal(size=(L * L, 2)) + 1j * rng.normal(size=(L * L, 2))
    e = T(e0)
    zero = torch.zeros_like(e)
    lv.relax(6, phi=e, r=None, smoother="gs")
    lo1, hi1 = mg2d.diagnostics.mode_amplitudes(e, zero, L)
    o2 = copy.deepcopy(LVLo[0])
    o2.phi, o2.r = e0.copy(), np.zeros_like(e0)
    o2.relax(L, 6, 1)
    so = np.abs(np.fft.fft2(o2.phi.reshape(L, L, 2).transpose(2, 0, 1)))
    k = np.abs(np.fft.fftfreq(L, d=1.0 / L))
    low = (k[:, None] <= L // 4) & (k[None, :] <= L // 4)
    assert abs(lo1 - so[:, low].max()) < 1e-9 * lo1 and abs(hi1 - so[:, ~low].max()) < 1e-9 * hi1
    lo0, hi0 = mg2d.diagnostics.mode_amplitudes(T(e0), zero, L)
    assert lo1 < lo0 and hi1 < hi0           # relaxation damps both bands


def test_critical_mass_estimate_against_dense_spectrum():
    """critical.estimate_critical_mass (shifted inverse iteration with MG-preconditioned FGCR solves, gamma5-symmetric
    quotient) against the dense spectrum of the oracle's 1152 x 1152 Wilson matrix D(0): D(0) + m becomes singular where m
    crosses minus the smallest REAL eigenvalue.  The links carry one unit of topological charge (constant flux + noise), which
    gives D(0) an isolated real mode -- as the large quenched lattices of the bench have, and small ones usually do not."""
    from importlib import import_module
    critical = import_module("2d_multigrid_b200.critical")
    L = 24
    rng = np.random.default_rng(3)
    F = 2 * np.pi / (L * L)
    s = np.arange(L * L)
    x, y = s % L, s // L
    th = np.zeros((L * L, 2))
    th[:, 0] = -F * y
    th[:, 1] = np.where(y == L - 1, F * L * x, 0.0)
    th += 0.15 * rng.normal(size=th.shape)
    U = np.exp(1j * th)
    po = O.Params(L=L, num_iters=1, block=2, m=0.0, nlevels=1)
    lv = O.Level()
    lv.compute_lvl0_matrix(U, po)
    S = L * L
    A = np.zeros((2 * S, 2 * S), dtype=complex)
    e = np.zeros((S, 2), dtype=complex)
    for col in range(2 * S):
        e.reshape(-1)[col] = 1.0
        A[:, col] = lv.apply_D(e, L).reshape(-1)
        e.reshape(-1)[col] = 0.0
    ev = np.linalg.eigvals(A)
    lam_real = np.sort(ev[np.abs(ev.imag) < 1e-9].real)[0]
    assert 0.0 < lam_real < 0.1
    factory = lambda m: mg2d.make_params(L, m, nlevels=2, block=2, n_null=4, n_smooth=3, smoother="rbgs", null_iters=40, tol=1e-10)
    mcrit, hist = critical.estimate_critical_mass(T(U), factory)
    info = critical.estimate_critical_mass.info
    assert abs(mcrit + lam_real) < 1e-4, (mcrit, lam_real, info)
    assert info["last_change"] < 2e-5 and info["quotient"] == "gamma5" and abs(info["imag"]) < 1e-8


def test_complex64_and_mixed_precision_solves():
    """complex64 hierarchy (true residual ~1e-6 class) and the mixed-precision solve: complex64 V-cycle inside the
    complex128 FGCR must reach the same 1e-10 TRUE residual (fp64 check) as the all-complex128 solve."""
    L = 64
    U = T(O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=30)))
    b = torch.zeros((L * L, 2), dtype=torch.complex128, device="cuda")
    b[L // 2 + (L // 2) * L, 0] = 1.0
    p = mg2d.make_params(L, 0.0, nlevels=2, block=4, n_null=4, n_smooth=3, smoother="rbgs", null_iters=40, tol=1e-10)
    mg = mg2d.setup(U, p)
    x64, i64 = mg2d.solve(mg, rhs=b, tol=1e-10, outer="gcr", use_graph=True)
    xm, im = mg2d.solve(mg, rhs=b, tol=1e-10, outer="gcr", use_graph=True, precond_dtype="complex64")
    assert i64["converged"] and im["converged"] and im["true_resnorm"] < 1e-10
    assert abs(im["iters"] - i64["iters"]) <= 2
    chk = torch.empty_like(xm)
    mg.LVL[0].apply_D(chk, xm)
    assert float(torch.linalg.vector_norm(chk - b)) < 1e-10
    assert float((xm - x64).abs().max()) < 1e-8
    # half-precision operator storage in the preconditioner (16 coarse dof so that the half kernel is used)
    p16 = mg2d.make_params(L, 0.0, nlevels=2, block=4, n_null=8, n_smooth=3, smoother="rbgs", null_iters=40, tol=1e-10)
    mg16 = mg2d.setup(U, p16)
    xa, ia = mg2d.solve(mg16, rhs=b, tol=1e-10, outer="gcr")
    xh, ih = mg2d.solve(mg16, rhs=b, tol=1e-10, outer="gcr", precond_dtype="complex64+half")
    sh = mg16.info["single"]
    # level 1 streams its complex64 low-rank factors (fewer bytes than half-precision dense blocks), level 2 the half blocks
    assert sh.LVL[1].Dh is None and sh.LVL[1].F is not None and sh.LVL[2].Dh is not None and sh.use_half
    assert ih["converged"] and ih["true_resnorm"] < 1e-10 and abs(ih["iters"] - ia["iters"]) <= 3
    assert float((xh - xa).abs().max()) < 1e-8
    # the half kernel itself against the complex64 kernel on the same data (difference = half rounding of D only)
    l1 = sh.LVL[2]
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    ph = torch.randn((l1.S, l1.n), generator=g, dtype=torch.float32, device="cuda").to(torch.complex64)
    l1.r.copy_(torch.randn((l1.S, l1.n), generator=g, dtype=torch.float32, device="cuda").to(torch.complex64))
    pa, pb = ph.clone(), ph.clone()
    l1.relax(1, phi=pa, smoother="rbgs")
    dh, l1.Dh = l1.Dh, None
    l1.relax(1, phi=pb, smoother="rbgs")
    l1.Dh = dh
    assert float((pa - pb).abs().max()) < 5e-3 * float(pb.abs().max())
    # stand-alone complex64 hierarchy: converges to single-precision accuracy
    p32 = mg2d.make_params(L, 0.0, nlevels=2, block=4, n_null=4, n_smooth=3, smoother="rbgs", null_iters=40, tol=1e-5, dtype="complex64")
    m32 = mg2d.setup(U, p32)
    x32, i32 = mg2d.solve(m32, rhs=b, tol=1e-5, outer="gcr")
    assert i32["converged"]
    assert float((x32.to(torch.complex128) - x64).abs().max()) < 1e-4 * float(x64.abs().max())


def test_supplied_null_vectors_and_public_api():
    """setup(U, params, null_vectors=...) and the per-function API of SURVEY 8(b)."""
    L = 16
    U = O.gauge_gaussian(L, 0.3)
    po = O.Params(L=L, num_iters=2, block=2, m=0.05, nlevels=2, null_iters=24)
    LVLo, NTLo = O.build_reference_problem(po, U)
    O.compute_near_null(LVLo, NTLo, po, 1)
    p = mg2d.make_params(L, 0.05, nlevels=2, n_smooth=2, null_iters=24, matrix_free=False)
    mg = mg2d.setup(T(U), p, null_vectors=[T(LVLo[0].phi_null), T(LVLo[1].phi_null)])
    for l in (1, 2):
        assert rel(mg2d.D_to_reference_layout(mg.LVL[l].D), LVLo[l].D) < 1e-10
    rng = np.random.default_rng(0)
    v = rng.normal(size=(L * L, 2)) + 0j
    out = torch.empty_like(T(v))
    mg2d.apply_D(out, T(v), 0, mg)
    assert rel(out, LVLo[0].apply_D(v, L)) < 1e-12
    assert abs(mg2d.residue_mag(0, mg) / LVLo[0].get_residue_mag(L) - 1) < 1e-12
    mg2d.relax(0, 2, mg, 1)
    LVLo[0].relax(L, 2, 1)
    assert rel(mg.LVL[0].phi, LVLo[0].phi) < 1e-11
    vc = torch.zeros((64, 4), dtype=torch.complex128, device="cuda")
    mg2d.restriction(vc, mg.LVL[0].phi, 0, mg, 1)
    assert rel(vc, LVLo[0].restriction(LVLo[0].phi, 0, po, 1)) < 1e-11


# config 1: real scalar Laplace 32x32 (and the 64x64 golden sweep) on the GPU
@pytest.mark.parametrize("args,want", [((32, 0.1, 1, 3, 0), None), ((32, 0.2, 2, 20, 0), None), ((64, 0.1, 3, 3, 0), 19),
                                       ((64, 0.2, 4, 3, 0), 7), ((64, 0.08, 3, 3, 0), 29), ((64, 0.1, 3, 3, 1), None),
                                       ((32, 0.3, 0, 4, 0), None), ((16, 0.3, 2, 2, 1), None)])
def test_config1_scalar_laplace(args, want):
    L, m, nl, ni, tf = args
    it_o, phi_o, hist_o = S2.solve(L, m, nl, ni, tf)
    it_g, phi_g, hist_g = mg2d.solve_scalar(L, m, nl, ni, tf)
    assert it_g == it_o                                                 # identical iteration count
    if want is not None:
        assert it_g == want                                             # NB/2c...:587-598 golden numbers
    assert rel(phi_g, phi_o) < 1e-12
    assert abs(hist_g[0] / hist_o[0] - 1) < 1e-12


@pytest.mark.parametrize("L,m,nl,ni,want", [(32, 0.1, 2, 3, 231), (32, 0.05, 3, 20, 31)])
def test_config1_s1_variant(L, m, nl, ni, want):
    """BASELINE configs[0] literally (code/1_laplace_scalar/2D_laplace_Mgrid.cpp): `want` is what the reference binary
    prints ("Loop breaks at iteration N"), reproduced in tests/test_oracle_scalar.py::test_s1_variant_vs_binary."""
    it_o, phi_o, _ = S2.solve_s1(L, m, nl, ni)
    it_g, phi_g, _ = mg2d.solve_scalar_s1(L, m, nl, ni)
    assert it_o == want and it_g == want
    assert rel(phi_g, phi_o) < 1e-12


# ---- full-size properties (no oracle run possible at these sizes) ----------------------------------------
def test_fullsize_properties_1024():
    """config 4 size: linearity and gamma5-hermiticity of the matrix-free D, P P^dagger = 1, Galerkin identity,
    gauge covariance, and a solve that reaches 1e-10 (true residual)."""
    L = 1024
    th = mg2d.gauge.quenched_phases(L, 6.0, sweeps=20, device="cuda")
    U = torch.exp(1j * th).to(torch.complex128)
    p = mg2d.make_params(L, -0.02, nlevels=4, block=4, n_null=8, n_smooth=4, smoother="rbgs", null_iters=60, tol=1e-10, max_iters=200)
    mg = mg2d.setup(U, p, init="device")
    lv = mg.LVL[0]
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64, device="cuda") + 1j * torch.randn(*s, generator=g, dtype=torch.float64, device="cuda")
    v, w = rnd(L * L, 2), rnd(L * L, 2)
    Dv, Dw, Dvw = torch.empty_like(v), torch.empty_like(v), torch.empty_like(v)
    lv.apply_D(Dv, v); lv.apply_D(Dw, w)
    comb = v * (0.3 - 0.7j) + w
    lv.apply_D(Dvw, comb)
    assert float((Dvw - (Dv * (0.3 - 0.7j) + Dw)).abs().max()) < 1e-11          # linearity
    g5 = torch.tensor([1.0, -1.0], dtype=torch.complex128, device="cuda")
    lhs = torch.vdot((g5 * Dv).reshape(-1), w.reshape(-1))                      # <g5 D v, w>
    rhs = torch.vdot(v.reshape(-1), (g5 * Dw).reshape(-1))                      # <v, g5 D w>
    assert abs(lhs - rhs) / abs(lhs) < 1e-12                                     # D^dagger = g5 D g5
    # gauge covariance of the kernel
    om = torch.exp(1j * torch.rand(L * L, generator=g, dtype=torch.float64, device="cuda") * 6.283)
    omx = torch.roll(om.reshape(L, L), -1, 1).reshape(-1)
    omy = torch.roll(om.reshape(L, L), -1, 0).reshape(-1)
    U2 = torch.stack([om * U[:, 0] * omx.conj(), om * U[:, 1] * omy.conj()], 1).contiguous()
    lv2 = mg2d.MG(mg2d.make_params(L, -0.02, nlevels=0, smoother="rbgs")).LVL[0]
    lv2.compute_lvl0_matrix(U2, store=False)          # links only: matrix-free operator
    out2 = torch.empty_like(v)
    lv2.apply_D(out2, om[:, None] * v)
    assert float((out2 - om[:, None] * Dv).abs().max()) < 1e-11
    # test1 / test2 on the two finest interfaces
    for l in (0, 1):
        a, bl = mg.LVL[l], mg.LVL[l + 1]
        vc = rnd(bl.S, bl.n)
        f1 = torch.zeros((a.S, a.n), dtype=torch.complex128, device="cuda")
        a.prolongation(f1, vc, 1)
        c1 = torch.empty_like(vc)
        a.restriction(c1, f1, 1)
        assert float((c1 - vc).abs().max()) < 1e-11
        f2 = torch.empty_like(f1)
        a.apply_D(f2, f1)
        a.restriction(c1, f2, 1)
        c2 = torch.empty_like(vc)
        bl.apply_D(c2, vc)
        assert float((c1 - c2).abs().max()) < 1e-10
    rhs_ = torch.zeros((L * L, 2), dtype=torch.complex128, device="cuda")
    rhs_[L // 2 + (L // 2) * L, 0] = 1.0
    x, info = mg2d.solve(mg, rhs=rhs_, tol=1e-10, outer="gcr", use_graph=True, check_every=4)
    assert info["converged"] and info["true_resnorm"] < 1e-10
    chk = torch.empty_like(x)
    lv.apply_D(chk, x)
    assert float(torch.linalg.vector_norm(chk - rhs_)) < 1e-10
