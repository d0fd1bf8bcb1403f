"""Real scalar Laplace geometric multigrid (BASELINE config 1) on the GPU.

Mirrors main() of S2 = code/2_scalar_2d_nontelescoping/telescoping_2d_laplace_Mgrid.cpp (:178-347; the same
operators as code/1_laplace_scalar/2D_laplace_Mgrid.cpp:25-106): CLI `./a.out L m nlevels num_iters t_flag`,
a_l = 2^l, scale_l = 1/(4 + m^2 a_l^2), source b[L/2 + (L/2)L] = scale_0, phi = 0, lexicographic GS,
piecewise-constant restriction (x 1/4) / injection prolongation, optional non-telescoping average of
n_copies = 2 quadrant copies, stop on sum|res| < 1e-13 (prints the 0-based iteration).
"""
from __future__ import annotations

import math

import torch

from ._lib import Context, MG2DError


def _s():
    return torch.cuda.current_stream().cuda_stream


class ScalarMG:
    def __init__(self, L: int, m: float, nlevels: int, device: int | None = None):
        if not torch.cuda.is_available():
            raise MG2DError("2d_multigrid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        max_levels = int(math.log2(L)) - 1                                    # S2:218
        if nlevels > max_levels:
            raise ValueError(f"Too many levels {nlevels}. Can only have {max_levels} levels for lattice of size {L}")
        self.dev_index = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", self.dev_index)
        self.ctx = Context(self.dev_index)
        self.L, self.m, self.nlevels = L, m, nlevels
        self.size, self.a, self.scale = [L], [1.0], [1.0 / (4.0 + m * m)]
        for _ in range(1, nlevels + 1):
            self.size.append(self.size[-1] // 2)
            self.a.append(2.0 * self.a[-1])
            self.scale.append(1.0 / (4 + m * m * self.a[-1] * self.a[-1]))
        z = lambda n: torch.zeros(n * n, dtype=torch.float64, device=self.device)
        self.phi = [z(s) for s in self.size]
        self.r = [z(s) for s in self.size]
        js = self.size[nlevels]
        self.phi_tel = [z(js) for _ in range(4)]
        self.r_tel = [z(js) for _ in range(4)]
        self.out = torch.zeros(4, dtype=torch.float64, device=self.device)

    # the four operators of S2
    def relax(self, phi, b, lev, num_iter, gs_flag=1):
        self.ctx.call("mg2d_s2_relax", phi.data_ptr(), b.data_ptr(), self.size[lev], self.scale[lev], self.a[lev],
                      num_iter, gs_flag, _s())

    def projection(self, res_c, res_f, phi, level, quad):
        self.ctx.call("mg2d_s2_project", res_c.data_ptr(), res_f.data_ptr(), phi.data_ptr(), self.size[level],
                      self.scale[level], self.a[level], quad, _s())

    def interpolate(self, phi_f, phi_c, lev, quad):
        self.ctx.call("mg2d_s2_interpolate", phi_f.data_ptr(), phi_c.data_ptr(), self.size[lev], quad, _s())

    def get_residue_mag(self, level=0) -> float:
        self.ctx.call("mg2d_s2_residue_mag", self.phi[level].data_ptr(), self.r[level].data_ptr(), self.size[level],
                      self.scale[level], self.a[level], self.out.data_ptr(), _s())
        return float(self.out[0].item())

    def cycle(self, num_iters: int, t_flag: int = 0, n_copies: int = 2):
        """One pass of the loop body S2:277-314."""
        nl, phi, r = self.nlevels, self.phi, self.r
        if nl == 0:
            self.relax(phi[0], r[0], 0, num_iters)
            return
        for lvl in range(nl):
            self.relax(phi[lvl], r[lvl], lvl, num_iters)
            if lvl == nl - 1 and t_flag == 1:
                for i in range(4):
                    self.projection(self.r_tel[i], r[lvl], phi[lvl], lvl, i + 1)
            else:
                self.projection(r[lvl + 1], r[lvl], phi[lvl], lvl, 1)
        for lvl in range(nl, -1, -1):
            if lvl == nl and t_flag == 1:
                for i in range(4):
                    self.phi_tel[i].zero_()
                for i in range(n_copies):
                    self.relax(self.phi_tel[i], self.r_tel[i], lvl, num_iters)
                    self.interpolate(phi[lvl - 1], self.phi_tel[i], lvl, i + 1)
                self.ctx.call("mg2d_s2_scale", phi[lvl - 1].data_ptr(), 1.0 / n_copies, phi[lvl - 1].numel(), _s())
            else:
                self.relax(phi[lvl], r[lvl], lvl, num_iters)
                if lvl > 0:
                    self.interpolate(phi[lvl - 1], phi[lvl], lvl, 1)


def solve_scalar(L: int, m: float, nlevels: int, num_iters: int, t_flag: int = 0, max_iters: int = 5000,
                 res_threshold: float = 1.0e-13, n_copies: int = 2, device: int | None = None):
    """main() of S2: returns (iter [0-based, as the reference prints], phi_0 tensor, residual history)."""
    mg = ScalarMG(L, m, nlevels, device)
    mg.r[0][L // 2 + (L // 2) * L] = 1.0 * mg.scale[0]
    hist = []
    for it in range(max_iters):
        mg.cycle(num_iters, t_flag, n_copies)
        resmag = mg.get_residue_mag(0)
        hist.append(resmag)
        if resmag < res_threshold:
            return it, mg.phi[0], hist
        if resmag > 1e6:
            break
    return -1, mg.phi[0], hist


def solve_scalar_s1(L: int, m: float, nlevels: int, num_iters: int, max_iters: int = 10000, res_threshold: float = 1.0e-14,
                    device: int | None = None):
    """main() of code/1_laplace_scalar/2D_laplace_Mgrid.cpp:111-215 (BASELINE configs[0]): the S2 operators with four
    sources (:163), threshold 1e-14 (:122) and a way up that starts at level nlevels-1 (:181), so the coarsest level
    is projected to but never relaxed.  Returns (iter [0-based, as printed by :189], phi_0, residual history)."""
    mg = ScalarMG(L, m, nlevels, device)
    for (x, y), v in (((0, 0), 1.0), ((1, 0), 2.0), ((2, 2), 5.0), ((3, 3), 7.5)):
        mg.r[0][x + y * L] = v
    hist = []
    for it in range(max_iters):
        for lvl in range(nlevels):
            mg.relax(mg.phi[lvl], mg.r[lvl], lvl, num_iters)
            mg.projection(mg.r[lvl + 1], mg.r[lvl], mg.phi[lvl], lvl, 1)
        for lvl in range(nlevels - 1, -1, -1):
            mg.relax(mg.phi[lvl], mg.r[lvl], lvl, num_iters)
            if lvl > 0:
                mg.interpolate(mg.phi[lvl - 1], mg.phi[lvl], lvl, 1)
        resmag = mg.get_residue_mag(0)
        hist.append(resmag)
        if resmag < res_threshold:
            return it, mg.phi[0], hist
        if resmag > 1e6:
            break
    return -1, mg.phi[0], hist
