"""The drop-in boundary exercised from plain C: tests/abi/abi_consumer.c includes include/mg2d.h and nothing else of this
repository, links against libmg2d_sm100.so and calls the entry points that replace Level::f_apply_D / f_residue
(S6/level.h:251-265, 61-77) and Near_null::f_restriction (S6/near_null.h:217-240) on cudaMalloc'ed buffers.
CPU: the program compiles as C99 with -Wall -Wextra -Werror and links; the ctypes table agrees with the header TYPE by
TYPE.  GPU: its results equal the oracle's."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from mg2d import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_consumer(tmp_path):
    exe = str(tmp_path / "abi_consumer")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", f"-I{CUDA}/include", f"-I{ROOT}/include",
           os.path.join(HERE, "abi", "abi_consumer.c"), "-o", exe, f"-L{CUDA}/lib64", "-lcudart", f"-L{libdir}", "-lmg2d_sm100",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{CUDA}/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_c_consumer_compiles_and_links(tmp_path):
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libmg2d_sm100.so is not built: run __graft_entry__.build()")
    exe = build_consumer(tmp_path)
    assert os.path.exists(exe)
    # every symbol the C program needs resolves against the library
    need = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    assert {"mg2d_create", "mg2d_wilson_apply", "mg2d_stencil_apply", "mg2d_restrict"} <= set(re.findall(r"\b(mg2d_\w+)", need))


def _kind(t) -> str:
    import ctypes as C
    if t in (C.c_void_p, C.c_char_p) or t.__name__.startswith("LP_"):
        return "ptr"
    if t is C.c_double:
        return "double"
    if t is C.c_int:
        return "int"
    if t is C.c_longlong:          # (an alias of c_long on LP64; what matters is 8 bytes, signed)
        return "ll"
    if t is C.c_ulonglong:
        return "ull"
    raise KeyError(t)


def _c_kind(decl: str) -> str:
    d = decl.strip()
    if "*" in d:
        return "ptr"
    d = re.sub(r"\b(const|struct)\b", "", d)
    d = re.sub(r"\b[A-Za-z_]\w*$", "", d.strip()).strip() or d.strip()      # drop the parameter name
    return {"int": "int", "double": "double", "long long": "ll", "unsigned long long": "ull"}[d]


def test_ctypes_argument_types_match_header():
    """Not just the argument COUNT: a wrong int / long long / double in _lib.SIGNATURES would corrupt the call."""
    txt = open(os.path.join(ROOT, "include", "mg2d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    checked = 0
    for name, args in _lib.SIGNATURES.items():
        proto = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", txt, flags=re.S).group(1)
        params = [a for a in proto.split(",")]
        assert _c_kind(params[0]) == "ptr"                                # mg2d_ctx*
        got = [_c_kind(a) for a in params[1:]]
        want = [_kind(t) for t in args]
        assert got == want, (name, got, want)
        checked += 1
    assert checked >= 45


@pytest.mark.gpu
def test_c_consumer_results_equal_oracle(tmp_path):
    from oracle import mg_oracle as O
    exe = build_consumer(tmp_path)
    L, n, nc, block, mass = 24, 4, 8, 4, -0.02
    rng = np.random.default_rng(11)
    cr = lambda *s: rng.normal(size=s) + 1j * rng.normal(size=s)
    U = O.gauge_gaussian(L, 0.4, seed=3)
    v, w = cr(L * L, 2), cr(L * L, n)
    Dref = cr(L * L, 5, n, n)
    P = cr(L * L, nc, n)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(struct.pack("<iiiid", L, n, nc, block, mass))
        for a in (U, v, np.ascontiguousarray(Dref.transpose(0, 1, 3, 2)), w, P):      # D: column-major blocks on the device
            f.write(np.ascontiguousarray(a, dtype=np.complex128).tobytes())
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "abi_consumer ok" in r.stdout, r.stderr
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.uint8)
    S, Sc = L * L, (L // block) ** 2
    off = 0

    def take(count, dtype):
        nonlocal off
        nb = count * np.dtype(dtype).itemsize
        a = raw[off:off + nb].view(dtype)
        off += nb
        return a
    Dv = take(S * 2, np.complex128).reshape(S, 2)
    res = take(S * 2, np.complex128).reshape(S, 2)
    dots = take(4, np.float64)
    Dw = take(S * n, np.complex128).reshape(S, n)
    Pw = take(Sc * nc, np.complex128).reshape(Sc, nc)
    po = O.Params(L=L, num_iters=1, block=block, m=mass, nlevels=1, n_dof_scale=nc)
    lv = O.Level()
    lv.compute_lvl0_matrix(U, po)
    want_Dv = lv.apply_D(v, L)
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    assert rel(Dv, want_Dv) < 1e-12
    want_res = v - lv.apply_D(want_Dv, L)
    assert rel(res, want_res) < 1e-12
    assert abs(dots[0] - np.sum(np.abs(want_res) ** 2)) < 1e-10 * dots[0] and abs(dots[3] - np.sum(np.abs(v) ** 2)) < 1e-10 * dots[3]
    assert abs((dots[1] + 1j * dots[2]) - np.vdot(want_res, want_Dv)) < 1e-9 * abs(np.vdot(want_res, want_Dv)) + 1e-9
    lw = O.Level()
    lw.D = Dref
    assert rel(Dw, lw.apply_D(w, L)) < 1e-12
    lp = O.Level()
    lp.phi_null = P
    pw = O.Params(L=L, num_iters=1, block=block, m=mass, nlevels=1, n_dof_scale=nc)
    pw.n_dof = [n, nc]
    assert rel(Pw, lp.restriction(w, 0, pw, 1)) < 1e-12
