"""numpy restatement of the reference's real scalar Laplace geometric multigrid (BASELINE config 1)
-- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows `S2/` = /root/reference/code/2_scalar_2d_nontelescoping/telescoping_2d_laplace_Mgrid.cpp.
Pinned by the reference's golden iteration counts (NB/2c_analysis_mass_variation_non-telescoping.ipynb:555-598)
and by the S2 binary itself (oracle/Makefile -> oracle/_ref/s2_mgrid).

  a_l = 2^l, scale_l = 1/(4 + m^2 a_l^2)                                  S2:216,228-232
  relax   phi(s) = scale*(sum 4 nbrs - b(s) a^2), lexicographic GS (x outer)   S2:46-72
  residue b - (1/a^2)(sum 4 nbrs - phi/scale)                              S2:30-35
  project res_c = 1/4 sum over the quadrant's 2x2 block of the residual    S2:74-110
  interp  phi_f(4 sites) += phi_c ; phi_c = 0                              S2:112-143
  stop    sum|res| < 1e-13, prints 0-based iter                            S2:318-325
"""
from __future__ import annotations

import math

import numpy as np

from .mg_oracle import fronts, neighbours


class S2Params:
    def __init__(self, L: int, m: float, nlevels: int):
        max_levels = int(math.log2(L)) - 1  # S2:218
        if nlevels > max_levels:
            raise ValueError("Too many levels")
        self.L, self.m, self.nlevels = L, m, nlevels
        self.size = [L]
        self.a = [1.0]
        self.scale = [1.0 / (4.0 + m * m)]
        for _ in range(1, nlevels + 1):
            self.size.append(self.size[-1] // 2)
            self.a.append(2.0 * self.a[-1])
            self.scale.append(1.0 / (4 + m * m * self.a[-1] * self.a[-1]))


def residue(phi, b, level, p):
    L = p.size[level]
    xp, xm, yp, ym = neighbours(L)
    return b - (1.0 / p.a[level] ** 2) * (phi[xp] + phi[xm] + phi[yp] + phi[ym] - phi / p.scale[level])


def get_residue_mag(phi, b, level, p) -> float:
    return float(np.sum(np.abs(residue(phi, b, level, p))))


def relax(phi, res, lev, num_iter, p, gs_flag=1):
    L, a, sc = p.size[lev], p.a[lev], p.scale[lev]
    xp, xm, yp, ym = neighbours(L)
    if gs_flag == 1:
        fr = fronts(L)
        for _ in range(num_iter):
            for idx in fr:
                phi[idx] = sc * (phi[xp[idx]] + phi[xm[idx]] + phi[yp[idx]] + phi[ym[idx]] - res[idx] * a * a)
    else:
        for _ in range(num_iter):
            phi[:] = sc * (phi[xp] + phi[xm] + phi[yp] + phi[ym] - res * a * a)


def _quad_sites(L, Lc, quad):
    X = np.arange(Lc * Lc)
    x, y = X % Lc, X // Lc
    xa, ya = 2 * x, 2 * y
    sx = 1 if quad in (1, 4) else -1
    sy = 1 if quad in (1, 2) else -1
    xb, yb = (2 * x + sx + L) % L, (2 * y + sy + L) % L
    return xa + ya * L, xa + yb * L, xb + ya * L, xb + yb * L


def projection(res_f, phi, level, p, quad):
    L, Lc = p.size[level], p.size[level + 1]
    rt = residue(phi, res_f, level, p)
    s0, s1, s2, s3 = _quad_sites(L, Lc, quad)
    return 0.25 * (rt[s0] + rt[s1] + rt[s2] + rt[s3])


def interpolate(phi_f, phi_c, lev, p, quad):
    Lc, L = p.size[lev], p.size[lev - 1]
    for s in _quad_sites(L, Lc, quad):
        phi_f[s] += phi_c
    phi_c[:] = 0.0


def solve_s1(L, m, nlevels, num_iters, max_iters=10000, res_threshold=1.0e-14):
    """main() of S1 = code/1_laplace_scalar/2D_laplace_Mgrid.cpp:111-215 (BASELINE configs[0] names this file).
    Same operators as S2 (relax :51-66, f_projection :68-91, f_interpolate :93-109); differences: four sources
    r[0][0][0]=1, r[0][1][0]=2, r[0][2][2]=5, r[0][3][3]=7.5 (:163, arrays are [x][y]), threshold 1e-14 (:122),
    and the way up starts at level nlevels-1 (:181): the coarsest level receives a residual but is never relaxed.
    Returns (iter [0-based, as printed by :189], phi_0, history)."""
    p = S2Params(L, m, nlevels)
    phi = [np.zeros(p.size[i] ** 2) for i in range(nlevels + 1)]
    r = [np.zeros(p.size[i] ** 2) for i in range(nlevels + 1)]
    for (x, y), v in (((0, 0), 1.0), ((1, 0), 2.0), ((2, 2), 5.0), ((3, 3), 7.5)):
        r[0][x + y * L] = v
    hist = []
    for it in range(max_iters):
        for lvl in range(nlevels):
            relax(phi[lvl], r[lvl], lvl, num_iters, p)
            r[lvl + 1] = projection(r[lvl], phi[lvl], lvl, p, 1)
        for lvl in range(nlevels - 1, -1, -1):
            relax(phi[lvl], r[lvl], lvl, num_iters, p)
            if lvl > 0:
                interpolate(phi[lvl - 1], phi[lvl], lvl, p, 1)
        resmag = get_residue_mag(phi[0], r[0], 0, p)
        hist.append(resmag)
        if resmag < res_threshold:
            return it, phi[0], hist
        if resmag > 1e6:
            break
    return -1, phi[0], hist


def solve(L, m, nlevels, num_iters, t_flag=0, max_iters=5000, res_threshold=1.0e-13, n_copies=2):
    """main() of S2 (:178-347).  Returns (iter [0-based, as printed], phi_0, residual history)."""
    p = S2Params(L, m, nlevels)
    phi = [np.zeros(p.size[i] ** 2) for i in range(nlevels + 1)]
    r = [np.zeros(p.size[i] ** 2) for i in range(nlevels + 1)]
    js = p.size[nlevels] ** 2
    phi_tel = [np.zeros(js) for _ in range(4)]
    r_tel = [np.zeros(js) for _ in range(4)]
    r[0][L // 2 + (L // 2) * L] = 1.0 * p.scale[0]
    hist = []
    for it in range(max_iters):
        if nlevels > 0:
            for lvl in range(nlevels):
                relax(phi[lvl], r[lvl], lvl, num_iters, p)
                if lvl == nlevels - 1 and t_flag == 1:
                    for i in range(4):
                        r_tel[i] = projection(r[lvl], phi[lvl], lvl, p, i + 1)
                else:
                    r[lvl + 1] = projection(r[lvl], phi[lvl], lvl, p, 1)
            for lvl in range(nlevels, -1, -1):
                if lvl == nlevels and t_flag == 1:
                    for i in range(4):
                        phi_tel[i][:] = 0.0
                    for i in range(n_copies):
                        relax(phi_tel[i], r_tel[i], lvl, num_iters, p)
                        if lvl > 0:
                            interpolate(phi[lvl - 1], phi_tel[i], lvl, p, i + 1)
                    phi[lvl - 1] /= float(n_copies)
                else:
                    relax(phi[lvl], r[lvl], lvl, num_iters, p)
                    if lvl > 0:
                        interpolate(phi[lvl - 1], phi[lvl], lvl, p, 1)
        else:
            relax(phi[0], r[0], 0, num_iters, p)
        resmag = get_residue_mag(phi[0], r[0], 0, p)
        hist.append(resmag)
        if resmag < res_threshold:
            return it, phi[0], hist
        if resmag > 1e6:
            break
    return -1, phi[0], hist
