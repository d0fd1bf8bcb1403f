"""The numpy restatement of S6 against everything the reference offers to pin it without Eigen:
its in-run property tests (S6/tests.h), the gauge-covariance test (S3/5_mg_without_quadrants/tests.h:186-484),
the analytic free Wilson spectrum (NB/spectrum_calc/1_compute_spectrum.ipynb cells 28-30), and the libstdc++
random stream (S6/mgrid_ntl.cpp:35-36)."""
import numpy as np
import pytest

from oracle import mg_oracle as O

EPS = 1.0e-12   # Epsilon of S6/tests.h:10


def test_rng_matches_libstdcxx():
    # first five values of std::mt19937(4302529) + uniform_real_distribution(-pi,pi), printed by a C++ program
    want = [-1.0164261620102244, -2.8518551777285293, 1.7189784040649814, 0.75964941503876515, 0.070352200777714202]
    got = O.StdMT19937(4302529).uniform_pm_pi(5)
    assert np.array_equal(got, np.array(want))


def test_free_wilson_spectrum():
    """lambda(k) = (2+m) + cos kx + cos ky +- i sqrt(sin^2 kx + sin^2 ky) for the +1/2(1-+gamma) hopping
    convention of S6/level.h:167-170; minimum |lambda| = |m| at k = (pi, pi)."""
    L, m = 8, 0.1
    p = O.Params(L=L, num_iters=1, block=2, m=m, nlevels=1)
    lv = O.Level()
    lv.compute_lvl0_matrix(O.gauge_cold(L), p)
    ev = np.linalg.eigvals(O.dense_matrix(lv, L, 2))
    k = 2 * np.pi * np.arange(L) / L
    kx, ky = np.meshgrid(k, k)
    root = np.sqrt(np.sin(kx) ** 2 + np.sin(ky) ** 2)
    want = np.concatenate([((2 + m) + np.cos(kx) + np.cos(ky) + s * 1j * root).ravel() for s in (1, -1)])
    for w in want:
        assert np.min(np.abs(ev - w)) < 1e-10
    assert abs(np.min(np.abs(ev)) - m) < 1e-10


@pytest.mark.parametrize("stencil,m,ntl", [("wilson", 0.05, 0), ("laplace", 0.05, 0), ("wilson", 0.05, 1), ("laplace", 0.1, 1)])
def test_reference_property_tests(stencil, m, ntl):
    """f_MG_tests (S6/tests.h:250-295): tests 1-4 on every level, NTL copies with their own quadrant."""
    L = 16
    p = O.Params(L=L, num_iters=2, block=2, m=m, nlevels=2, stencil=stencil, null_iters=40, t_flag=ntl, n_copies=4)
    LVL, NTL = O.build_reference_problem(p, O.gauge_gaussian(L))
    O.compute_near_null(LVL, NTL, p, 1)
    rng = np.random.default_rng(3)
    for lvl in range(p.nlevels + 1):
        S, n = p.size[lvl] ** 2, p.n_dof[lvl]
        vec = rng.uniform(-np.pi, np.pi, (S, n)) + 1j * rng.uniform(-np.pi, np.pi, (S, n))
        if ntl and lvl == p.nlevels:
            for q in range(p.n_copies):
                assert O.test1_restriction_prolongation(NTL[lvl - 1][q], vec, lvl - 1, p, q + 1) < EPS
                assert O.test2_D(vec, NTL[lvl][q], LVL[lvl - 1], NTL[lvl - 1][q], lvl - 1, p, q + 1) < EPS
                assert O.test3_hermiticity(NTL[lvl][q], lvl, p) < EPS
                assert O.test4_hermiticity_full(NTL[lvl][q], vec, lvl, p) < 1e-10
        else:
            if lvl > 0:
                assert O.test1_restriction_prolongation(LVL[lvl - 1], vec, lvl - 1, p, 1) < EPS
                assert O.test2_D(vec, LVL[lvl], LVL[lvl - 1], LVL[lvl - 1], lvl - 1, p, 1) < EPS
            assert O.test3_hermiticity(LVL[lvl], lvl, p) < EPS
            assert O.test4_hermiticity_full(LVL[lvl], vec, lvl, p) < 1e-10
        if lvl < p.nlevels:
            assert LVL[lvl].check_ortho(lvl, 1, p) < EPS


@pytest.mark.parametrize("stencil", ["wilson", "laplace"])
def test_gauge_covariance(stencil):
    """U' = Omega(x) U_mu(x) Omega(x+mu)^dagger  =>  D'(Omega v) = Omega (D v), |r| invariant, relaxation covariant
    (S3/5_mg_without_quadrants/tests.h:186-204, 233-484)."""
    L = 8
    rng = np.random.default_rng(5)
    U = O.gauge_gaussian(L)
    om = np.exp(1j * rng.uniform(-np.pi, np.pi, L * L))
    xp, _, yp, _ = O.neighbours(L)
    U2 = np.stack([om * U[:, 0] * np.conj(om[xp]), om * U[:, 1] * np.conj(om[yp])], axis=1)
    p = O.Params(L=L, num_iters=1, block=2, m=0.1, nlevels=1, stencil=stencil)
    a, b = O.Level(), O.Level()
    a.compute_lvl0_matrix(U, p)
    b.compute_lvl0_matrix(U2, p)
    n = p.n_dof[0]
    v = rng.normal(size=(L * L, n)) + 1j * rng.normal(size=(L * L, n))
    assert np.max(np.abs(b.apply_D(om[:, None] * v, L) - om[:, None] * a.apply_D(v, L))) < EPS
    a.phi, a.r = v.copy(), rng.normal(size=(L * L, n)) + 0j
    b.phi, b.r = om[:, None] * a.phi, om[:, None] * a.r
    assert abs(a.get_residue_mag(L) - b.get_residue_mag(L)) < EPS
    a.relax(L, 5, 1)
    b.relax(L, 5, 1)
    assert np.max(np.abs(b.phi - om[:, None] * a.phi)) < 1e-11


def test_wavefront_is_lexicographic_gs():
    """The anti-diagonal evaluation equals the literal `for x: for y:` loop of S6/level.h:113-123, bit for bit."""
    L = 6
    p = O.Params(L=L, num_iters=1, block=2, m=0.2, nlevels=1)
    LVL, _ = O.build_reference_problem(p, O.gauge_gaussian(L))
    lv = LVL[0]
    phi = lv.phi.copy()
    for _ in range(2):
        for x in range(L):
            for y in range(L):
                s = x + y * L
                acc = (lv.D[s, 1] @ phi[(x + 1) % L + y * L] + lv.D[s, 2] @ phi[(x - 1 + L) % L + y * L]
                       + lv.D[s, 3] @ phi[x + ((y + 1) % L) * L] + lv.D[s, 4] @ phi[x + ((y - 1 + L) % L) * L] - lv.r[s])
                phi[s] = (-1.0 * np.linalg.inv(lv.D[s, 0])) @ acc
    lv.relax(L, 2, 1)
    assert np.max(np.abs(lv.phi - phi)) < 1e-14


def test_colpiv_qr_solve():
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 4):
        A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        b = rng.normal(size=n) + 1j * rng.normal(size=n)
        assert np.max(np.abs(O.colpiv_householder_qr_solve(A, b) - np.linalg.solve(A, b))) < 1e-12
    A = np.ones((3, 3), dtype=complex)     # rank 1: minimum-structure solution, no NaN
    x = O.colpiv_householder_qr_solve(A, np.ones(3))
    assert np.all(np.isfinite(x)) and abs(np.sum(x) - 1.0) < 1e-12


def test_solve_converges_and_quirks():
    """main() flow: random start on all levels, source entry r[2+2L][0] = 5, iters = iter+1 (S6/level.h:57,
    S6/modules_main.h:467-475)."""
    L = 16
    p = O.Params(L=L, num_iters=3, block=2, m=0.05, nlevels=2, null_iters=40)
    LVL, NTL, info = O.run_reference_flow(p, O.gauge_gaussian(L))
    assert info["converged"] and info["iters"] == len(info["resnorms"]) and info["resnorms"][-1] < 1e-13
    assert LVL[0].r[2 + 2 * L, 0] == 5.0
    with pytest.raises(ValueError):
        O.Params(L=16, num_iters=1, block=2, m=0.1, nlevels=1, t_flag=1)     # S6/params.h:52-55
    with pytest.raises(ValueError):
        O.Params(L=16, num_iters=1, block=2, m=0.1, nlevels=5)               # S6/params.h:100-106


def test_gcr_and_smoothers_agree_on_solution():
    L = 16
    U = O.gauge_gaussian(L)
    b = np.zeros((L * L, 2), dtype=complex)
    b[5, 0] = 1.0
    xs = []
    for sm in ("gs", "rbgs", "mr", "jacobi"):
        p = O.Params(L=L, num_iters=3, block=2, m=0.1, nlevels=2, null_iters=40, smoother=sm)
        LVL, NTL = O.build_reference_problem(p, U)
        O.compute_near_null(LVL, NTL, p, 1)
        x, info = O.gcr_MG(LVL, NTL, p, b, tol=1e-12)
        assert info["converged"]
        xs.append(x)
    for x in xs[1:]:
        assert np.max(np.abs(x - xs[0])) < 1e-9


def test_kcycle_and_per_level_blocks():
    """ours, on top of the reference flow: the K-cycle is a stronger preconditioner than the V-cycle and converges to the
    same solution; a per-level block list equal to the scalar block size is the same hierarchy; an uneven list coarsens
    as listed (S5L/setup.h:2-10)."""
    L = 16
    U = O.gauge_gaussian(L, 0.3)
    b = np.zeros((L * L, 2), dtype=complex)
    b[5, 0] = 1.0
    res = {}
    for cyc in ("V", "K"):
        p = O.Params(L=L, num_iters=2, block=2, m=-0.02, nlevels=3, null_iters=40, smoother="rbgs", cycle=cyc)
        LVL, NTL = O.build_reference_problem(p, U)
        O.compute_near_null(LVL, NTL, p, 1)
        res[cyc] = O.gcr_MG(LVL, NTL, p, b, tol=1e-11, restart=8)
    assert res["K"][1]["converged"] and res["K"][1]["iters"] <= res["V"][1]["iters"]
    assert np.max(np.abs(res["K"][0] - res["V"][0])) < 1e-8 * np.max(np.abs(res["V"][0]))
    pa = O.Params(L=L, num_iters=2, block=2, m=0.05, nlevels=2, null_iters=20)
    pb = O.Params(L=L, num_iters=2, block=[2, 2], m=0.05, nlevels=2, null_iters=20)
    _, _, ia = O.run_reference_flow(pa, U)
    _, _, ib = O.run_reference_flow(pb, U)
    assert ia["resnorms"] == ib["resnorms"]
    pc = O.Params(L=L, num_iters=2, block=[4, 2], m=0.05, nlevels=2, null_iters=20)
    assert pc.size == [16, 4, 2]
    LVLc, _, ic = O.run_reference_flow(pc, U)
    assert ic["converged"] and LVLc[1].D.shape[0] == 16 and LVLc[2].D.shape[0] == 4


def test_c_port_matches_numpy_oracle():
    """oracle/c_port (plain C + OpenMP restatement of the solve loop, used as the CPU baseline of bench.py) against the
    numpy oracle on the bench's cycle shape: same iteration count, residual history and solution."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from oracle import c_port
    L = 32
    U = O.gauge_from_phases(O.gauge_quenched_phases(L, 6.0, sweeps=20))
    p = O.Params(L=L, num_iters=4, n_pre=0, n_post=[4, 2], block=4, m=-0.02, nlevels=1, smoother="rbgs", n_dof_scale=16, null_iters=20)
    LVL, NTL = O.build_reference_problem(p, U)
    O.compute_near_null(LVL, NTL, p, 1)
    b = np.zeros((L * L, 2), dtype=complex)
    b[L // 2 + (L // 2) * L, 0] = 1.0
    import copy
    xo, io = O.gcr_MG(copy.deepcopy(LVL), NTL, p, b, tol=1e-10, restart=8)
    xc, ic = c_port.gcr_solve(LVL, p, b, tol=1e-10, restart=8)
    assert ic["iters"] == io["iters"] and ic["converged"]
    assert max(abs(a / c - 1) for a, c in zip(ic["resnorms"], io["resnorms"])) < 1e-8
    assert np.max(np.abs(xc - xo)) < 1e-12 * np.max(np.abs(xo))
    # pre-smoothing path too
    p2 = O.Params(L=L, num_iters=2, block=4, m=-0.02, nlevels=1, smoother="rbgs", n_dof_scale=16, null_iters=20)
    xo2, io2 = O.gcr_MG(copy.deepcopy(LVL), NTL, p2, b, tol=1e-10, restart=4)
    xc2, ic2 = c_port.gcr_solve(LVL, p2, b, tol=1e-10, restart=4)
    assert ic2["iters"] == io2["iters"] and np.max(np.abs(xc2 - xo2)) < 1e-12 * np.max(np.abs(xo2))
