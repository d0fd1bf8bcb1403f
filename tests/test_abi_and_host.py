"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mg2d.h
declares (no compute call is made), the ctypes table matches the header, and the host-side parameter logic
follows S6/params.h."""
import ctypes
import os
import re

import pytest

import mg2d
from mg2d import _lib


def _header_functions(repo_root):
    txt = open(os.path.join(repo_root, "include", "mg2d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mg2d_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(repo_root):
    names = _header_functions(repo_root)
    assert len(names) >= 35
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libmg2d_sm100.so is not built: run __graft_entry__.build()")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.mg2d_version() >= 100


def test_ctypes_table_matches_header(repo_root):
    names = set(_header_functions(repo_root))
    bound = set(_lib.SIGNATURES) | set(_lib.PLAIN)
    assert names == bound, names ^ bound
    # argument counts agree with the prototypes
    txt = open(os.path.join(repo_root, "include", "mg2d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        proto = re.search(name + r"\s*\((.*?)\)\s*;", txt, flags=re.S).group(1)
        assert len(proto.split(",")) == len(args) + 1, name


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mg2d.MG2DError):
        mg2d.MG(mg2d.make_params(16, 0.1))
    with pytest.raises(mg2d.MG2DError):
        mg2d.ScalarMG(16, 0.1, 1)


def test_params_follow_reference():
    p = mg2d.make_params(64, -0.015, nlevels=3, block=2)
    assert p.size == [64, 32, 16, 8] and p.n_dof == [2, 4, 4, 4]           # S6/params.h:75,115,121
    assert abs(p.diag - (2.0 - 0.015)) < 1e-15                              # 1/scale[0], S6/params.h:76
    q = mg2d.make_params(32, 0.1, stencil="laplace", nlevels=2)
    assert q.n_dof == [1, 2, 2] and abs(q.diag + 4.1) < 1e-15               # S6/params.h:81-82, level.h:148
    c4 = mg2d.make_params(1024, 0.0, nlevels=4, block=4, n_null=8, smoother="rbgs")
    assert c4.size == [1024, 256, 64, 16, 4] and c4.n_dof == [2, 16, 16, 16, 16] and c4.matrix_free
    with pytest.raises(ValueError):
        mg2d.make_params(16, 0.1, nlevels=1, ntl=True)                      # S6/params.h:52-55
    with pytest.raises(ValueError):
        mg2d.make_params(16, 0.1, nlevels=5)                                # S6/params.h:100-106
    with pytest.raises(ValueError):
        mg2d.make_params(16, 0.1, stencil="staggered")                      # S6/params.h:84-86
    a = mg2d.from_argv(["32", "3", "2", "1", "-0.015", "4", "1", "4"])      # S6/params.h:42-50
    assert (a.L, a.n_smooth, a.block, a.mass, a.nlevels, a.ntl, a.n_copies) == (32, 3, 2, -0.015, 4, True, 4)


def test_rng_product_copy_matches_oracle():
    import numpy as np
    from oracle import mg_oracle as O
    assert np.array_equal(mg2d.StdMT19937(7).uniform_pm_pi(64), O.StdMT19937(7).uniform_pm_pi(64))


def test_phase_file_roundtrip(tmp_path):
    import numpy as np
    L = 6
    th = np.random.default_rng(0).uniform(-3, 3, (L * L, 2))
    f = tmp_path / "phase_6_b6.0.dat"
    mg2d.gauge.write_phase_file(str(f), th, L)
    assert np.allclose(mg2d.gauge.read_phase_file(str(f), L), th, atol=1e-15)
    lines = open(f).read().split()
    # x outer, y inner, dir inner (S6/gauge.h:103-107): second record is (x=0,y=0,dir=1), third (x=0,y=1,dir=0)
    assert abs(float(lines[1]) - th[0, 1]) < 1e-15 and abs(float(lines[2]) - th[L, 0]) < 1e-15


def test_error_spectrum_diagnostics():
    """A single plane wave error shows up in exactly one FFT bin (NB/2_spectral_analysis_solution.ipynb cell 5)."""
    import numpy as np
    import torch
    L, kx, ky = 16, 3, 5
    s = np.arange(L * L)
    x, y = s % L, s // L
    err = np.exp(2j * np.pi * (kx * x + ky * y) / L)
    phi_star = torch.zeros((L * L, 2), dtype=torch.complex128)
    phi = phi_star.clone()
    phi[:, 1] = torch.as_tensor(err)
    spec = mg2d.diagnostics.error_spectrum(phi, phi_star, L)
    assert spec.shape == (2, L, L)
    assert abs(float(spec[1, ky, kx]) - L * L) < 1e-9 and float(spec[0].max()) == 0.0
    spec[1, ky, kx] = 0
    assert float(spec.max()) < 1e-9
    lo, hi = mg2d.diagnostics.mode_amplitudes(phi, phi_star, L)
    assert hi > 100 and lo < 1e-9        # (3,5): ky > L/4 -> a high-frequency mode


def test_gamma5_quotient_is_second_order():
    """critical._quotients (host logic of the critical-mass estimate): D is gamma5-hermitian, so for a right eigenvector x the
    left one is gamma5 x and <g5 x, D x>/<g5 x, x> has a second-order error in a perturbation of x where the plain Rayleigh
    quotient of this non-normal matrix has a first-order one.  Dense 8^2 Wilson matrix of the oracle, links with one unit of
    flux (an isolated real mode)."""
    import numpy as np
    import torch
    from importlib import import_module
    from oracle import mg_oracle as O
    critical = import_module("2d_multigrid_b200.critical")
    L = 8
    rng = np.random.default_rng(1)
    F = 2 * np.pi / (L * L)
    s = np.arange(L * L)
    x, y = s % L, s // L
    th = np.zeros((L * L, 2))
    th[:, 0] = -F * y
    th[:, 1] = np.where(y == L - 1, F * L * x, 0.0)
    th += 0.1 * rng.normal(size=th.shape)
    lv = O.Level()
    lv.compute_lvl0_matrix(np.exp(1j * th), O.Params(L=L, num_iters=1, block=2, m=0.0, nlevels=1))
    S = L * L
    A = np.zeros((2 * S, 2 * S), dtype=complex)
    e = np.zeros((S, 2), dtype=complex)
    for col in range(2 * S):
        e.reshape(-1)[col] = 1.0
        A[:, col] = lv.apply_D(e, L).reshape(-1)
        e.reshape(-1)[col] = 0.0
    g5 = np.tile([1.0, -1.0], S)
    assert np.abs(g5[:, None] * A * g5[None, :] - A.conj().T).max() < 1e-13          # gamma5 D gamma5 = D^dagger
    w, V = np.linalg.eig(A)
    k = np.argmin(np.where(np.abs(w.imag) < 1e-9, w.real, np.inf))
    lam, v = w[k].real, V[:, k]
    errs = []
    d = (rng.normal(size=2 * S) + 1j * rng.normal(size=2 * S)) / np.sqrt(2 * S)
    for eps in (1e-2, 1e-3):
        xp = v + eps * d
        xt = torch.as_tensor(xp.reshape(S, 2))
        Dx = torch.as_tensor((A @ xp).reshape(S, 2))
        q5, rq = critical._quotients(xt, Dx, True)
        errs.append((abs(q5 - lam), abs(rq - lam)))
    assert errs[1][0] < 2e-2 * errs[0][0] + 1e-14        # x10 smaller perturbation -> ~x100 smaller error
    assert errs[1][0] < 0.1 * errs[1][1]                  # and far better than the plain Rayleigh quotient
    q5, rq = critical._quotients(torch.as_tensor(v.reshape(S, 2)), torch.as_tensor((A @ v).reshape(S, 2)), True)
    assert abs(q5 - lam) < 1e-12
