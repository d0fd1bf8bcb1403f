"""ctypes glue for oracle/c_port/mg_port.c (plain C + OpenMP restatement of the solve loop) -- TEST INFRASTRUCTURE.
The hierarchy comes from the numpy oracle (oracle.mg_oracle); this module only runs the timed solve."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(os.path.dirname(_HERE), "_ref")
_SO = os.path.join(_OUT, "libmgport.so")


class _Level(C.Structure):
    _fields_ = [("L", C.c_int), ("n", C.c_int), ("nc", C.c_int), ("D", C.c_void_p), ("mD0inv", C.c_void_p), ("P", C.c_void_p),
                ("phi", C.c_void_p), ("r", C.c_void_p), ("tmp", C.c_void_p)]


def build() -> str:
    src = os.path.join(_HERE, "mg_port.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(_OUT, exist_ok=True)
        # baseline x86-64 ISA on purpose: the built library travels with the repo snapshot to a different host CPU
        subprocess.check_call(["gcc", "-O3", "-fcx-limited-range", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def gcr_solve(LVL, p, b, tol=1e-10, max_iters=1000, restart=8):
    """Same contract as oracle.mg_oracle.gcr_MG for smoother 'rbgs', quad 1, telescoping cycle.  Returns (x, info)."""
    assert p.smoother == "rbgs" and p.t_flag == 0 and p.quad == 1
    lib = C.CDLL(build())
    lib.mgport_gcr_solve.restype = C.c_int
    assert lib.mgport_level_size() == C.sizeof(_Level)
    keep, levels = [], (_Level * (p.nlevels + 1))()
    for l, lv in enumerate(LVL):
        S, n = p.size[l] ** 2, p.n_dof[l]
        D = np.ascontiguousarray(lv.D, dtype=np.complex128)
        mD0inv = np.ascontiguousarray(-np.linalg.inv(lv.D[:, 0]), dtype=np.complex128)
        P = np.ascontiguousarray(lv.phi_null, dtype=np.complex128) if l < p.nlevels else None
        phi, r, tmp = (np.zeros((S, n), dtype=np.complex128) for _ in range(3))
        keep += [D, mD0inv, P, phi, r, tmp]
        levels[l] = _Level(p.size[l], n, p.n_dof[l + 1] if l < p.nlevels else 0, D.ctypes.data, mD0inv.ctypes.data,
                           P.ctypes.data if P is not None else None, phi.ctypes.data, r.ctypes.data, tmp.ctypes.data)
    pre = (C.c_int * (p.nlevels + 1))(*p.pre)
    post = (C.c_int * (p.nlevels + 1))(*p.post)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.zeros_like(b)
    res = np.zeros(max_iters, dtype=np.float64)
    import time
    t0 = time.perf_counter()
    it = lib.mgport_gcr_solve(C.c_int(p.nlevels), levels, pre, post, C.c_int(p.block), C.c_void_p(b.ctypes.data),
                              C.c_void_p(x.ctypes.data), C.c_double(tol), C.c_int(max_iters), C.c_int(restart),
                              C.c_void_p(res.ctypes.data))
    info = {"iters": int(it), "resnorms": res[:it].tolist(), "converged": bool(it > 0 and res[it - 1] < tol),
            "seconds": time.perf_counter() - t0}
    return x, info
