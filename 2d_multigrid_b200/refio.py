"""The reference's text file formats, so that data can move between the reference program and
this package in both directions:

  Near-null_L{L}_blk{block}_ndof{n_dof_scale}.txt   f_write_near_null / f_read_near_null, S6/modules_main.h:39-79
      one complex per line "%20.25e+i%20.25e", levels 0..nlevels-1 concatenated, per level: site j, row d1, column d2
  ../gauge_config_files/phase_{L}_b{beta}.dat        f_read_gauge_heatbath, S6/gauge.h:88-110 (see gauge.py)
  results_phi.txt                                    Level::f_write_op, S6/level.h:287-300: "iter," then
      "%20.25e+i%20.25e," per dof, sites x OUTER / y inner
  results_gen_scaling.txt                            S6/modules_main.h:472: L num_iters m block_x block_y n_dof_scale nlevels iters

With gen_null = 0 (argv[4]) the reference reads the near-null file instead of relaxing 500 sweeps, so vectors produced on
the GPU can be handed to the unmodified reference program, and vice versa.
"""
from __future__ import annotations

import numpy as np


def near_null_filename(L: int, block: int, n_dof_scale: int) -> str:
    return "Near-null_L%d_blk%d_ndof%d.txt" % (L, block, n_dof_scale)      # S6/modules_main.h:43,66


def write_near_null(path: str, nulls) -> None:
    """nulls: list over levels of phi_null[L_l^2, nc, nf] (numpy or torch)."""
    with open(path, "w") as f:
        for P in nulls:
            P = np.asarray(P.cpu() if hasattr(P, "cpu") else P).reshape(-1)
            f.write("".join("%20.25e+i%20.25e\n" % (z.real, z.imag) for z in P))


def read_near_null(path: str, size, n_dof):
    """size / n_dof: the per-level lists of the parameter object.  Returns the list of phi_null arrays."""
    vals = []
    with open(path) as f:
        for line in f:
            a, b = line.strip().split("+i")
            vals.append(float(a) + 1j * float(b))
    vals = np.array(vals, dtype=np.complex128)
    out, off = [], 0
    for lvl in range(len(size) - 1):
        n = size[lvl] ** 2 * n_dof[lvl + 1] * n_dof[lvl]
        out.append(vals[off:off + n].reshape(size[lvl] ** 2, n_dof[lvl + 1], n_dof[lvl]))
        off += n
    if off != len(vals):
        raise ValueError("near-null file does not match the level sizes")
    return out


def write_results_phi_row(f, it: int, phi, L: int) -> None:
    """One row of results_phi.txt (S6/level.h:287-300): sites x outer, y inner."""
    v = np.asarray(phi.cpu() if hasattr(phi, "cpu") else phi).reshape(L, L, -1).transpose(1, 0, 2).reshape(-1)
    f.write("%d," % it + "".join("%20.25e+i%20.25e," % (z.real, z.imag) for z in v) + "\n")


def gen_scaling_row(L, num_iters, m, block, n_dof_scale, nlevels, iters) -> str:
    return "%d\t%d\t%f\t%d\t%d\t%d\t%d\t%d\n" % (L, num_iters, m, block, block, n_dof_scale, nlevels, iters)


def write_ntl_weights_row(f, it: int, a, total_copies: int = 4) -> None:
    """One row of results_NTL_weights.txt (f_write_NTL_weights, S6/modules_indiv.h:137-143): total_copies = 4 entries."""
    a = list(np.asarray(a).reshape(-1)) + [0j] * total_copies
    f.write("%d," % it + "".join("%.4e+i%.4e," % (complex(z).real, complex(z).imag) for z in a[:total_copies]) + "\n")


class ResultWriters:
    """The reference's per-iteration result files (S6/params.h:89-97 open them, f_perform_MG writes them,
    S6/modules_main.h:446-458, 469-475), so that its analysis notebooks (NB/7a, NB/8a) can read a GPU run unchanged:
        results_phi.txt              row "iter+1, phi_0(x outer, y inner, dof)" at the START of iteration `iter` and once
                                     more after convergence                                  (Level::f_write_op)
        results_res_lvl-%d.txt       the residual r - D phi of every level, same rows       (Level::f_write_residue; with
                                     t_flag the lowest level is NTL copy 0, modules_main.h:452-454)
        results_NTL_weights.txt      row "iter, a_0..a_3" after every non-telescoping cycle  (f_write_NTL_weights)
    The reference writes every iteration (write_interval = 1, S6/params.h:65), which costs a device->host copy of every
    level per iteration; `stride` k writes every k-th iteration (and always the final state) and keeps the solve resident.
    Use as  perform_MG(mg, on_iteration=w.on_iteration)  then  w.finish(mg, info)."""

    def __init__(self, directory: str, p, stride: int = 1):
        import os
        self.p, self.stride = p, max(int(stride), 1)
        self.f_phi = open(os.path.join(directory, "results_phi.txt"), "w")
        self.f_w = open(os.path.join(directory, "results_NTL_weights.txt"), "w")
        self.f_res = [open(os.path.join(directory, "results_res_lvl-%d.txt" % lvl), "w") for lvl in range(p.nlevels + 1)]
        self.rows = 0

    def _state_rows(self, label: int, mg, final: bool):
        p = self.p
        write_results_phi_row(self.f_phi, label, mg.LVL[0].phi, p.size[0])
        for lvl in range(p.nlevels + 1):
            lv = mg.NTL[lvl][0] if (p.ntl and lvl == p.nlevels and not final) else mg.LVL[lvl]
            if lv.D is None and not lv.matrix_free or lv.phi is None or lv.r is None:
                continue
            rt = lv.work("rtemp")
            lv.residue(rt)
            write_results_phi_row(self.f_res[lvl], label, rt, p.size[lvl])
        self.rows += 1

    def on_iteration(self, it: int, mg):
        if it % self.stride == 0:
            self._state_rows(it + 1, mg, final=False)

    def finish(self, mg, info):
        for k, w in enumerate(info.get("ntl_weights", [])):
            write_ntl_weights_row(self.f_w, k, w)
        if info.get("converged"):
            self._state_rows(info["iters"], mg, final=True)
        for f in [self.f_phi, self.f_w] + self.f_res:
            f.close()
