"""The reference's command line on the GPU path:

    python -m mg2d L num_iters block gen_null m nlevels t_flag n_copies [--stencil wilson|laplace] [--beta 32.0]

Same 8 positional arguments as `./a.out` (S6/params.h:42-50), same inputs (`../gauge_config_files/phase_{L}_b{beta}.dat`,
S6/gauge.h:44; `Near-null_L*_blk*_ndof*.txt` when gen_null = 0, S6/modules_main.h:39-60) and the same outputs in the
working directory: `results_gen_scaling.txt` (appended row, S6/modules_main.h:472), `results_phi.txt`,
`results_res_lvl-%d.txt` and `results_NTL_weights.txt` (one row per iteration as the reference writes them,
S6/modules_main.h:446-458, S6/level.h:266-300, S6/modules_indiv.h:137-143; --write-interval k thins them out),
`Near-null_*.txt` (when gen_null = 1) and the "Ans" line on stdout.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch


def main(argv=None):
    import mg2d
    ap = argparse.ArgumentParser(prog="python -m mg2d")
    ap.add_argument("args", nargs=8, help="L num_iters block gen_null m nlevels t_flag n_copies")
    ap.add_argument("--stencil", default="wilson", choices=["wilson", "laplace"])
    ap.add_argument("--beta", type=float, default=32.0)
    ap.add_argument("--write-interval", type=int, default=1,
                    help="write results_phi / results_res_lvl-* rows every k-th iteration (reference: 1, S6/params.h:65); "
                         "0: only the final phi row")
    ns = ap.parse_args(argv)
    gen_null = int(ns.args[3])
    p = mg2d.from_argv(ns.args, stencil=ns.stencil)
    fname = "../gauge_config_files/phase_%d_b%0.1f.dat" % (p.L, ns.beta)                    # S6/gauge.h:44
    if not os.path.exists(fname):
        print("\nCannot find file " + fname)                                              # S6/gauge.h:98-101
        return 1
    theta = mg2d.gauge.read_phase_file(fname, p.L)
    U = torch.as_tensor(mg2d.gauge.from_phases(theta)).cuda()
    print("\nPlaquette", mg2d.gauge.plaquette(U, p.L))
    nn_file = mg2d.refio.near_null_filename(p.L, p.block, p.n_dof_scale)
    nulls = None
    if gen_null == 0 and p.nlevels > 0:
        if not os.path.exists(nn_file):
            print("\nCannot find file " + nn_file)                                        # S6/modules_main.h:50-53
            return 1
        nulls = [torch.as_tensor(a) for a in mg2d.refio.read_near_null(nn_file, p.size, p.n_dof)]
    mg = mg2d.setup(U, p, null_vectors=nulls, init="reference")
    if gen_null == 1 and p.nlevels > 0:
        mg2d.refio.write_near_null(nn_file, [mg.LVL[l].phi_null for l in range(p.nlevels)])
    writers = mg2d.refio.ResultWriters(".", p, stride=ns.write_interval) if ns.write_interval > 0 else None
    info = mg2d.perform_MG(mg, on_iteration=None if writers is None else writers.on_iteration)
    x = mg.LVL[0].phi
    if info["converged"]:
        print("\nLoop breaks at iteration %d with residue %e < %e" % (info["iters"], info["resnorms"][-1], p.tol))
        print("\nL %d\tm %f\tnlevels %d\tnum_per_level %d\tAns %d" % (p.L, p.mass, p.nlevels, p.n_smooth, info["iters"]))
        with open("results_gen_scaling.txt", "a") as f:
            f.write(mg2d.refio.gen_scaling_row(p.L, p.n_smooth, p.mass, p.block, p.n_dof_scale, p.nlevels, info["iters"]))
    elif info["diverged"]:
        print("\nDiverging. Residue %g at iteration %d" % (info["resnorms"][-1], info["iters"]))
    if writers is not None:
        writers.finish(mg, info)
    else:
        with open("results_phi.txt", "w") as f:
            mg2d.refio.write_results_phi_row(f, info["iters"], x, p.L)
    return 0


if __name__ == "__main__":
    sys.exit(main())
