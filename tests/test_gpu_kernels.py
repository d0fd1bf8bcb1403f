"""GPU parity tests proper: every kernel behind the C ABI against the numpy oracle on identical seeded inputs.
fp64 tolerance: 1e-12 relative (BASELINE.json north_star: "D.x within 1e-12 relative (fp64)"); complex64
variants 2e-5."""
import copy

import numpy as np
import pytest
import torch

import mg2d
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def crand(rng, *s):
    return rng.normal(size=s) + 1j * rng.normal(size=s)


def _pair(L, m, stencil="wilson", nlevels=2, block=2, n_null=None, **kw):
    """oracle levels + GPU MG with identical reference-compatible initial data."""
    nds = None if n_null is None else (2 * n_null if stencil == "wilson" else n_null)
    po = O.Params(L=L, num_iters=2, block=block, m=m, nlevels=nlevels, stencil=stencil, n_dof_scale=nds, null_iters=24, **kw)
    U = O.gauge_gaussian(L, width=0.4, seed=9)
    LVLo, NTLo = O.build_reference_problem(po, U)
    p = mg2d.make_params(L, m, stencil=stencil, nlevels=nlevels, block=block, n_null=n_null, n_smooth=2, null_iters=24,
                         ntl=bool(kw.get("t_flag", 0)), n_copies=kw.get("n_copies", 4), matrix_free=False)
    mg = mg2d.MG(p)
    mg.init_reference_fields()
    mg.set_gauge(T(U))
    return po, LVLo, NTLo, p, mg, U


@pytest.mark.parametrize("L", [4, 6, 34, 64, 130, 256])
def test_wilson_apply_matrix_free(L):
    """mg2d_wilson_apply vs Level.apply_D; ragged sizes exercise the partial-warp / partial-strip paths."""
    rng = np.random.default_rng(L)
    po, LVLo, _, p, mg, U = _pair(L, -0.03, nlevels=1)
    v = crand(rng, L * L, 2)
    want = LVLo[0].apply_D(v, L)
    lv = mg.LVL[0]
    lv.matrix_free = True
    out = torch.empty_like(T(v))
    lv.apply_D(out, T(v))
    assert rel(out, want) < TOL
    # residual + fused reductions
    d = lv.dots("t")
    lv.phi.copy_(T(v))
    lv._stencil(out, lv.phi, lv.r, 1, d)
    r_want = LVLo[0].r - want
    assert rel(out, r_want) < TOL
    dd = d[:4].cpu().numpy()
    assert abs(dd[0] / np.sum(np.abs(r_want) ** 2) - 1) < TOL
    assert abs((dd[1] + 1j * dd[2]) / np.vdot(r_want, v) - 1) < 1e-10
    assert abs(dd[3] / np.sum(np.abs(LVLo[0].r) ** 2) - 1) < TOL


def test_wilson_apply_complex64():
    L = 64
    rng = np.random.default_rng(1)
    po, LVLo, _, _, _, U = _pair(L, 0.01, nlevels=1)
    p = mg2d.make_params(L, 0.01, nlevels=1, dtype="complex64", matrix_free=False)
    mg = mg2d.MG(p)
    mg.init_reference_fields()
    mg.set_gauge(T(U))
    v = crand(rng, L * L, 2)
    want = LVLo[0].apply_D(v, L)
    out = torch.empty((L * L, 2), dtype=torch.complex64, device="cuda")
    mg.LVL[0].apply_D(out, T(v).to(torch.complex64))
    assert rel(out, want) < 2e-5
    mg.LVL[0].matrix_free = True
    mg.LVL[0].apply_D(out, T(v).to(torch.complex64))
    assert rel(out, want) < 2e-5


@pytest.mark.parametrize("stencil", ["wilson", "laplace"])
def test_lvl0_matrix_and_stencil(stencil):
    L = 12
    rng = np.random.default_rng(2)
    po, LVLo, _, p, mg, U = _pair(L, 0.05, stencil=stencil, nlevels=1)
    assert rel(mg2d.D_to_reference_layout(mg.LVL[0].D), LVLo[0].D) < 1e-15
    n = po.n_dof[0]
    v = crand(rng, L * L, n)
    out = torch.empty_like(T(v))
    mg.LVL[0].apply_D(out, T(v))
    assert rel(out, LVLo[0].apply_D(v, L)) < TOL
    assert abs(mg.LVL[0].get_residue_mag() / LVLo[0].get_residue_mag(L) - 1) < TOL


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32])
def test_block_stencil_all_sizes(n):
    """generic n x n block stencil: apply, residual, D0 inverse, Jacobi, lexicographic GS, red-black GS."""
    L = 10
    rng = np.random.default_rng(n)
    S = L * L
    Dref = crand(rng, S, 5, n, n) * 0.2
    Dref[:, 0] += 3.0 * np.eye(n)
    lvo = O.Level()
    lvo.D, lvo.phi, lvo.r = Dref, crand(rng, S, n), crand(rng, S, n)
    p = mg2d.make_params(L, 0.1, nlevels=0, matrix_free=False)
    mg = mg2d.MG(p)
    lv = mg.LVL[0]
    lv.n, lv.L, lv.S = n, L, S
    lv.D = mg2d.D_to_reference_layout(T(Dref))
    lv.phi, lv.r = T(lvo.phi), T(lvo.r)
    v = crand(rng, S, n)
    out = torch.empty_like(T(v))
    lv.apply_D(out, T(v))
    assert rel(out, lvo.apply_D(v, L)) < TOL
    rt = torch.empty_like(out)
    lv.residue(rt)
    assert rel(rt, lvo.residue(L)) < TOL
    lv._ensure_D0inv()
    assert rel(lv.D0inv.transpose(-1, -2), np.linalg.inv(Dref[:, 0])) < 1e-11
    for sm, fn in (("jacobi", lambda o: o.relax(L, 2, 0)), ("gs", lambda o: o.relax(L, 2, 1)), ("rbgs", lambda o: o.relax_rb(L, 2)),
                   ("mr", lambda o: o.relax_mr(L, 2))):
        o2 = copy.deepcopy(lvo)
        fn(o2)
        for premul in ((True, False) if sm == "rbgs" else (True,)):     # rbgs: pre-multiplied blocks -D0^-1 D_k and plain blocks
            mg.premul = premul
            phi0 = lv.phi.clone()
            lv.relax(2, smoother=sm)
            assert rel(lv.phi, o2.phi) < 1e-11, (sm, premul)
            lv.phi.copy_(phi0)
    mg.premul = True
    if n <= 16:
        lv._ensure_M()
        Mref = -np.einsum("sil,sklj->skij", np.linalg.inv(Dref[:, 0]), Dref[:, 1:])
        assert rel(lv.M.transpose(-1, -2), Mref) < 1e-11
    # batched red-black path (4 vectors share one stream of the operator) == vector-at-a-time path
    for nvec in (4, 8):
        V = T(crand(rng, nvec, S, n))
        Vb = V.clone()
        lv.relax(2, phi=Vb, r=None, smoother="rbgs")
        for k in range(nvec):
            one = V[k].clone()
            lv.relax(2, phi=one, r=None, smoother="rbgs")
            assert float((Vb[k] - one).abs().max()) < 1e-12 * max(1.0, float(one.abs().max())), (n, nvec, k)


@pytest.mark.parametrize("n,L", [(2, 10), (8, 34), (16, 34), (16, 64)])
def test_persistent_sweeps_kernel(n, L):
    """mg2d_relax_rb_pm_sweeps (all red-black sweeps of one relax call in ONE cooperative launch, grid barriers between the
    half sweeps) against the launch-per-half-sweep kernel (bit-identical: same arithmetic order) and the oracle's relax_rb."""
    rng = np.random.default_rng(100 + n)
    S = L * L
    Dref = crand(rng, S, 5, n, n) * 0.2
    Dref[:, 0] += 3.0 * np.eye(n)
    lvo = O.Level()
    lvo.D, lvo.phi, lvo.r = Dref, crand(rng, S, n), crand(rng, S, n)
    p = mg2d.make_params(L, 0.1, nlevels=0, matrix_free=False)
    mg = mg2d.MG(p)
    lv = mg.LVL[0]
    lv.n, lv.L, lv.S = n, L, S
    lv.D = mg2d.D_to_reference_layout(T(Dref))
    for with_r in (True, False):
        for nsw in (1, 3):
            o2 = copy.deepcopy(lvo)
            if not with_r:
                o2.r = np.zeros_like(o2.r)
            o2.relax_rb(L, nsw)
            out = []
            for persistent in (8192, 0):
                mg.persistent_sites = persistent
                phi = T(lvo.phi)
                n0 = mg.ctx.launches
                lv.relax(nsw, phi=phi, r=T(lvo.r) if with_r else None, smoother="rbgs")
                out.append((phi, mg.ctx.launches - n0))
            assert rel(out[0][0], o2.phi) < 1e-11, (with_r, nsw)
            assert torch.equal(out[0][0], out[1][0]), (with_r, nsw)
            assert out[0][1] < out[1][1] or nsw == 1


def test_relax_matrix_free_and_batched():
    L = 16
    po, LVLo, _, p, mg, U = _pair(L, 0.02, nlevels=1)
    lv = mg.LVL[0]
    for sm, fn in (("rbgs", lambda o: o.relax_rb(L, 3)), ("mr", lambda o: o.relax_mr(L, 3))):
        o2 = copy.deepcopy(LVLo[0])
        fn(o2)
        for mf in (False, True):
            lv.matrix_free = mf
            phi0 = lv.phi.clone()
            lv.relax(3, smoother=sm)
            assert rel(lv.phi, o2.phi) < 1e-11, (sm, mf)
            lv.phi.copy_(phi0)
    lv.matrix_free = False
    # batched vectors with r = 0 (near-null relaxation) equal one-at-a-time relaxation
    rng = np.random.default_rng(3)
    V = T(crand(rng, 3, L * L, 2))
    for sm in ("gs", "jacobi", "rbgs", "mr"):
        Vb = V.clone()
        lv.relax(2, phi=Vb, r=None, smoother=sm)
        for k in range(3):
            o2 = copy.deepcopy(LVLo[0])
            o2.phi, o2.r = V[k].cpu().numpy().copy(), np.zeros((L * L, 2), dtype=complex)
            {"gs": lambda: o2.relax(L, 2, 1), "jacobi": lambda: o2.relax(L, 2, 0), "rbgs": lambda: o2.relax_rb(L, 2),
             "mr": lambda: o2.relax_mr(L, 2)}[sm]()
            assert rel(Vb[k], o2.phi) < 1e-11, (sm, k)


@pytest.mark.parametrize("stencil,block,n_null", [("wilson", 2, None), ("laplace", 2, None), ("wilson", 4, 8), ("wilson", 2, 4), ("laplace", 4, 4)])
def test_setup_and_transfer(stencil, block, n_null):
    """near_null, norm_nn, ortho x2, check_ortho, coarse matrix, restrict / prolong (all quadrants) and the
    reference's tests 1 (P P^dagger = 1) and 2 (Galerkin identity) evaluated on GPU data."""
    L = 16
    rng = np.random.default_rng(4)
    po, LVLo, NTLo, p, mg, U = _pair(L, 0.05, stencil=stencil, nlevels=2 if block == 2 else 1, block=block, n_null=n_null)
    O.compute_near_null(LVLo, NTLo, po, 1)
    mg2d.compute_near_null(mg)
    assert max(mg.info["ortho_worst"]) < TOL                              # f_check_ortho, S6/near_null.h:205
    for l in range(po.nlevels):
        assert rel(mg.LVL[l].phi_null, LVLo[l].phi_null) < 1e-10
        assert rel(mg2d.D_to_reference_layout(mg.LVL[l + 1].D), LVLo[l + 1].D) < 1e-10
        nf, nc = po.n_dof[l], po.n_dof[l + 1]
        Sf, Sc = po.size[l] ** 2, po.size[l + 1] ** 2
        vf, vc = crand(rng, Sf, nf), crand(rng, Sc, nc)
        for quad in (1, 2, 3, 4):
            rc = torch.empty_like(T(vc))
            mg.LVL[l].restriction(rc, T(vf), quad)
            assert rel(rc, LVLo[l].restriction(vf, l, po, quad)) < TOL
            fo = vf.copy()
            LVLo[l].prolongation(fo, vc, l + 1, po, quad)
            fg = T(vf).clone()
            cg = T(vc).clone()
            mg.LVL[l].prolongation(fg, cg, quad, zero_vc=True)
            assert rel(fg, fo) < TOL and float(cg.abs().max()) == 0.0
        # test1 / test2 of S6/tests.h on the GPU
        f1 = torch.zeros((Sf, nf), dtype=torch.complex128, device="cuda")
        mg.LVL[l].prolongation(f1, T(vc), 1)
        c1 = torch.empty_like(T(vc))
        mg.LVL[l].restriction(c1, f1, 1)
        assert float((c1 - T(vc)).abs().max()) < 1e-12
        f2 = torch.empty_like(f1)
        mg.LVL[l].apply_D(f2, f1)
        mg.LVL[l].restriction(c1, f2, 1)
        c2 = torch.empty_like(c1)
        mg.LVL[l + 1].apply_D(c2, T(vc))
        assert float((c1 - c2).abs().max()) < 1e-11


@pytest.mark.parametrize("stencil,block,n_null", [("wilson", 4, 8), ("wilson", 2, 4), ("laplace", 4, 16)])
def test_lowrank_hop_factors_and_sweep(stencil, block, n_null):
    """First coarse level: the Galerkin hopping blocks (S6/modules_main.h:148-155) have rank <= block because every fine
    hopping term has rank one (S6/level.h:139-172).  mg2d_hop_factors must reproduce the dense blocks of
    mg2d_coarse_matrix (sum_b A_q B_q^dagger = D_k), and the red-black sweep on the packed factors (mg2d_relax_rb_lr) must
    equal the oracle's relax_rb and the dense pre-multiplied kernel, with and without a right-hand side."""
    L = 16
    rng = np.random.default_rng(8)
    po, LVLo, NTLo, p, mg, U = _pair(L, 0.05, stencil=stencil, nlevels=1, block=block, n_null=n_null)
    O.compute_near_null(LVLo, NTLo, po, 1)
    mg.persistent_sites = 0
    lv = mg.LVL[1]
    P = mg.LVL[0].phi_null
    mg2d.compute_near_null(mg)
    assert lv.lr_rank == block and int(mg.status[1].item()) == 0
    # factors (recomputed here: compute_near_null already packed and dropped them)
    p_lo, p_hi = mg.LVL[0]._halo(mg.LVL[0].phi_null, 1, lv.n * mg.LVL[0].n)
    lv.hop_factors(mg.LVL[0], mg.LVL[0].phi_null, p_lo, p_hi)
    A, B = lv._lr_AB
    Sc, nc = lv.S, lv.n
    A4, B4 = A.reshape(Sc, 4, block, nc), B.reshape(Sc, 4, block, nc)
    rec = torch.einsum("skbi,skbj->skij", A4, B4.conj())
    assert rel(rec, LVLo[1].D[:, 1:]) < 1e-10
    assert rel(rec, mg2d.D_to_reference_layout(lv.D)[:, 1:].cpu().numpy()) < 1e-13
    lv.F = None
    assert lv._ensure_F() and lv._lr_AB is None
    for with_r in (True, False):
        phi0, r0 = crand(rng, Sc, nc), crand(rng, Sc, nc)
        o2 = copy.deepcopy(LVLo[1])
        o2.phi, o2.r = phi0.copy(), (r0.copy() if with_r else np.zeros_like(r0))
        o2.relax_rb(po.size[1], 3)
        out = {}
        for lowrank in (True, False):
            mg.lowrank = lowrank
            phi = T(phi0)
            lv.relax(3, phi=phi, r=T(r0) if with_r else None, smoother="rbgs")
            out[lowrank] = phi
        mg.lowrank = True
        assert rel(out[True], o2.phi) < 1e-11 and rel(out[False], o2.phi) < 1e-11
        assert float((out[True] - out[False]).abs().max()) < 1e-12 * float(out[False].abs().max())
    assert lv.M is not None         # the dense path above rebuilt its blocks lazily
    # batches of 4 vectors share one stream of the factors (near-null relaxation): equal to one vector at a time, bit for bit
    for nvec in (4, 8):
        V = T(crand(rng, nvec, Sc, nc))
        Vb = V.clone()
        lv.relax(2, phi=Vb, r=None, smoother="rbgs")
        for k in range(nvec):
            one = V[k].clone()
            lv.relax(2, phi=one, r=None, smoother="rbgs")
            assert torch.equal(Vb[k], one), (nvec, k)


def test_lowrank_not_used_when_fine_hops_are_not_rank_one():
    """A fine operator whose hopping blocks have full rank (random 2x2 blocks): mg2d_hop_factors flags it and the setup
    keeps the dense blocks."""
    L = 16
    rng = np.random.default_rng(9)
    po, LVLo, NTLo, p, mg, U = _pair(L, 0.05, nlevels=1, block=4, n_null=8)
    D = crand(rng, L * L, 5, 2, 2) * 0.2
    D[:, 0] += 3.0 * np.eye(2)
    mg.LVL[0].D = T(D)
    mg2d.compute_near_null(mg)
    assert mg.LVL[1].lr_rank == 0 and mg.LVL[1].F is None and int(mg.status[1].item()) == 0


def test_minres_pieces():
    """Gram matrix (mg2d_cdot_batch), column-pivoted QR solve (mg2d_minres_solve), f_scale_phi."""
    rng = np.random.default_rng(6)
    p = mg2d.make_params(8, 0.1, nlevels=0, matrix_free=False)
    mg = mg2d.MG(p)
    n = 300
    X, Y = crand(rng, 4, n), crand(rng, 3, n)
    out = torch.zeros(64, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    Xd, Yd = T(X), T(Y)          # keep the device tensors alive across the asynchronous call
    mg.ctx.call("mg2d_cdot_batch", Xd.data_ptr(), n, 4, Yd.data_ptr(), n, 3, n, 0, out.data_ptr(), s)
    got = out[:24].cpu().numpy()
    want = np.conj(X) @ Y.T
    assert np.max(np.abs((got[0::2] + 1j * got[1::2]).reshape(4, 3) - want)) < 1e-11
    for k in (1, 2, 3, 4):
        A, b = crand(rng, k, k), crand(rng, k)
        g = T(np.stack([A.real, A.imag], -1).reshape(-1))
        sv = T(np.stack([b.real, b.imag], -1).reshape(-1))
        a = torch.zeros(8, dtype=torch.float64, device="cuda")
        mg.ctx.call("mg2d_minres_solve", g.data_ptr(), sv.data_ptr(), k, a.data_ptr(), s)
        av = a.cpu().numpy()
        assert np.max(np.abs(av[0:2 * k:2] + 1j * av[1:2 * k:2] - O.colpiv_householder_qr_solve(A, b))) < 1e-11


def test_error_behaviour():
    """The ABI reports errors by return code + message; the python layer raises MG2DError (no exit(), no crash)."""
    p = mg2d.make_params(8, 0.1, nlevels=1, matrix_free=False)
    mg = mg2d.MG(p)
    mg.init_reference_fields()
    mg.set_gauge(T(O.gauge_cold(8)))
    v = mg.LVL[0].new_field()
    with pytest.raises(ValueError):
        mg.LVL[0].apply_D(v, v)                                         # aliasing
    s = torch.cuda.current_stream().cuda_stream
    with pytest.raises(mg2d.MG2DError, match="n_dof"):
        mg.ctx.call("mg2d_stencil_apply", v.data_ptr(), mg.LVL[0].phi.data_ptr(), v.data_ptr(), v.data_ptr(),
                    mg.LVL[0].D.data_ptr(), None, 3, 8, 8, 0, 0, 1, 128, 128, None, s)
    with pytest.raises(mg2d.MG2DError, match="geometry"):
        mg.ctx.call("mg2d_restrict", v.data_ptr(), v.data_ptr(), v.data_ptr(), 2, 4, 9, 9, 2, 1, 0, s)
    with pytest.raises(mg2d.MG2DError):
        mg.ctx.call("mg2d_wilson_apply", None, v.data_ptr(), v.data_ptr(), v.data_ptr(), v.data_ptr(), v.data_ptr(),
                    None, 0.1, 8, 8, 0, 0, None, s)
    assert mg.ctx.launches > 0


@pytest.mark.parametrize("L", [4, 6, 34, 64, 130, 256])
def test_wilson_two_colour_sweep(L):
    """mg2d_wilson_relax_rb2 (both colours in one pass, out of place) vs oracle Level.relax_rb and vs the two half-sweep
    launches it replaces; ragged sizes exercise partial tiles / chunks and the periodic wrap of the two-row halos."""
    rng = np.random.default_rng(L)
    po, LVLo, _, p, mg, U = _pair(L, -0.03, nlevels=1)
    lv = mg.LVL[0]
    lv.matrix_free = True
    phi0 = lv.phi.clone()
    for nsweep in (1, 2, 3):
        o2 = copy.deepcopy(LVLo[0])
        o2.relax_rb(L, nsweep)
        mg.two_colour = True
        lv.phi.copy_(phi0)
        lv.relax(nsweep, smoother="rbgs")
        assert rel(lv.phi, o2.phi) < 1e-11, nsweep
        got = lv.phi.clone()
        mg.two_colour = False
        lv.phi.copy_(phi0)
        lv.relax(nsweep, smoother="rbgs")
        assert rel(got, lv.phi.cpu().numpy()) < 1e-13, nsweep
    # r = 0 (near-null relaxation), batched
    V = T(crand(rng, 2, L * L, 2))
    mg.two_colour = True
    Va = V.clone(); lv.relax(2, phi=Va, r=None, smoother="rbgs")
    mg.two_colour = False
    Vb = V.clone(); lv.relax(2, phi=Vb, r=None, smoother="rbgs")
    assert rel(Va, Vb.cpu().numpy()) < 1e-13
    mg.two_colour = True


def test_wilson_two_colour_sweep_complex64():
    L = 64
    U = O.gauge_gaussian(L, width=0.4, seed=9)
    outs = {}
    for two in (True, False):
        p = mg2d.make_params(L, 0.05, nlevels=0, smoother="rbgs", dtype="complex64")
        mg = mg2d.MG(p)
        mg.init_reference_fields()
        mg.LVL[0].compute_lvl0_matrix(T(U).to(torch.complex64), store=False)
        mg.two_colour = two
        mg.LVL[0].relax(3)
        outs[two] = mg.LVL[0].phi.cpu().numpy()
    assert np.max(np.abs(outs[True] - outs[False])) < 2e-5 * np.max(np.abs(outs[False]))


def test_counter_rng_and_gauge_kernels():
    """mg2d_fill_uniform == oracle counter_uniform bit for bit (any offset); mg2d_gauge_metropolis == the oracle's counter
    Metropolis draw for draw; mg2d_plaquette == Gauge::f_plaquette."""
    p = mg2d.make_params(16, 0.1, nlevels=0, matrix_free=False)
    mg = mg2d.MG(p)
    out = torch.empty(1000, dtype=torch.complex128, device="cuda")
    for seed, stream, off in ((4302529, 0, 0), (7, 3, 123456789012), (2**40 + 5, 17, 999)):
        mg.ctx.call("mg2d_fill_uniform", out.data_ptr(), 1000, off, seed, stream, -np.pi, np.pi, mg.dcode, torch.cuda.current_stream().cuda_stream)
        want = O.counter_uniform(seed, stream, off, 1000, -np.pi, np.pi)
        got = out.cpu().numpy()
        assert np.array_equal(got.real, want) and not got.imag.any()
    for L, beta, sweeps in ((16, 6.0, 12), (32, 32.0, 8)):
        Ud, th = mg2d.gauge.quenched_links_device(L, beta, sweeps=sweeps, seed=1234, return_phases=True)
        want = O.gauge_quenched_phases_counter(L, beta, sweeps=sweeps, seed=1234)
        assert np.max(np.abs(th.cpu().numpy() - want)) < 1e-12
        Uo = O.gauge_from_phases(want)
        assert rel(Ud, Uo) < 1e-14
        pd, po_ = mg2d.gauge.plaquette_device(Ud, L), O.plaquette(Uo, L)
        assert abs(pd - po_) < 1e-13
        assert 0.3 < pd.real < 1.0


def test_device_init_fields_match_oracle():
    """MG.init_fields (counter-seeded near-null start) == oracle build_device_problem, and the resulting hierarchy too."""
    L = 32
    U = O.gauge_gaussian(L, width=0.4, seed=9)
    po = O.Params(L=L, num_iters=2, block=4, m=0.01, nlevels=1, stencil="wilson", smoother="rbgs", n_dof_scale=8, null_iters=12)
    LVLo, NTLo = O.build_device_problem(po, U)
    p = mg2d.make_params(L, 0.01, nlevels=1, block=4, n_null=4, n_smooth=2, smoother="rbgs", null_iters=12)
    mg = mg2d.MG(p)
    mg.init_fields()
    assert np.array_equal(mg.LVL[0].phi_null.cpu().numpy(), LVLo[0].phi_null)
    mg.set_gauge(T(U))
    mg2d.compute_near_null(mg)
    O.compute_near_null(LVLo, NTLo, po, 1)
    assert rel(mg.LVL[0].phi_null, LVLo[0].phi_null) < 1e-10
    assert rel(mg2d.D_to_reference_layout(mg.LVL[1].D), LVLo[1].D) < 1e-10
