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
