#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/s6_*.npz by RUNNING THE REFERENCE.

The reference's own solver source (S6 = /root/reference/code/6_ntl-mg_new_code/3_combining_laplace_and_wilson/
mgrid_ntl.cpp, unmodified) is compiled by oracle/Makefile against oracle/eigen_shim (our stand-in for the absent
Eigen headers) into oracle/_ref/s6_mgrid_ntl[_laplace] and run here with the reference CLI
    ./a.out L num_iters block gen_null m nlevels t_flag n_copies                 (S6/params.h:42-50)
in a scratch directory holding ../gauge_config_files/phase_{L}_b32.0.dat (the file S6/gauge.h:44 demands; the
reference ships none, so the phases come from oracle.mg_oracle.gauge_quenched_phases with the seed recorded in
the fixture).  Each fixture stores the inputs and what the reference wrote:
  argv, stencil, theta[L*L,2]                       inputs
  iters                                             "Ans %d" (S6/modules_main.h:471)
  resmag[k]                                         residual printed at the start of iteration k+1 (6 digits)
  phi_final[L*L,n], phi_after_1[L*L,n]              results_phi.txt (S6/level.h:287-300; x outer, y inner)
  null0[L*L,nc,nf]                                  Near-null_L*_blk*_ndof*.txt, level 0 (S6/modules_main.h:62-79)
  ntl_weights[k,4]                                  results_NTL_weights.txt (4 significant digits)
Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mg_oracle as O  # noqa: E402

CASES = [
    # name, stencil, L, num_iters, block, m, nlevels, t_flag, n_copies, gauge sweeps
    ("s6_wilson16_2lvl", "wilson", 16, 3, 2, 0.05, 2, 0, 1, 30),
    ("s6_wilson16_1lvl", "wilson", 16, 3, 2, 0.05, 1, 0, 1, 30),
    ("s6_wilson16_ntl4", "wilson", 16, 3, 2, 0.05, 2, 1, 4, 30),
    ("s6_wilson16_ntl2", "wilson", 16, 3, 2, 0.05, 2, 1, 2, 30),
    ("s6_wilson32_3lvl", "wilson", 32, 3, 2, -0.005, 3, 0, 1, 50),
    ("s6_laplace16_2lvl", "laplace", 16, 3, 2, 0.05, 2, 0, 1, 30),
    ("s6_laplace16_ntl4", "laplace", 16, 3, 2, 0.05, 2, 1, 4, 30),
    ("s6_wilson16_relax_only", "wilson", 16, 10, 2, 0.3, 0, 0, 1, 30),
    ("s6_wilson16_blk4", "wilson", 16, 3, 4, 0.05, 1, 0, 1, 30),
    ("s6_wilson16_ntl3", "wilson", 16, 3, 2, 0.05, 2, 1, 3, 30),
    ("s6_wilson32_negmass_2lvl", "wilson", 32, 4, 2, -0.01, 2, 0, 1, 50),
    ("s6_laplace16_blk4", "laplace", 16, 3, 4, 0.1, 1, 0, 1, 30),
]


# Full-size cases (BASELINE configs[2]: Wilson 256^2 near-critical, non-telescoping): the reference needs ~6 min and
# writes ~2.4 GB of per-iteration text, so only the LAST row of results_phi.txt and the head of the near-null file are kept.
# m_crit of this configuration is -0.0163 (smallest real part of spec D(0), scipy eigs): m = -0.01 is 0.006 above it.
LARGE_CASES = [
    ("big_s6_wilson256_ntl4_nearcrit", "wilson", 256, 3, 2, -0.01, 3, 1, 4, 50),
]


def last_line(path):
    with open(path, "rb") as f:
        f.seek(0, 2)
        end = f.tell()
        pos = end - 2
        while pos > 0:
            f.seek(pos)
            if f.read(1) == b"\n":
                break
            pos -= 1
        f.seek(pos + 1 if pos > 0 else 0)
        return f.read().decode()


def make_large(reuse=None):
    """reuse: a directory that already holds run/ of the case (skips the 6-minute reference run)."""
    for name, stencil, L, ni, blk, m, nl, tf, nco, sweeps in LARGE_CASES:
        exe = os.path.join(ROOT, "oracle", "_ref", "s6_mgrid_ntl")
        theta = O.gauge_quenched_phases(L, 32.0, sweeps=sweeps, seed=1234)
        argv = [str(L), str(ni), str(blk), "1", repr(m), str(nl), str(tf), str(nco)]
        d = reuse or tempfile.mkdtemp()
        run = os.path.join(d, "run")
        if not reuse:
            os.makedirs(run)
            os.makedirs(os.path.join(d, "gauge_config_files"))
            O.write_phase_file(os.path.join(d, "gauge_config_files", f"phase_{L}_b32.0.dat"), theta, L)
            with open(os.path.join(run, "out.txt"), "w") as fo:
                subprocess.run([exe] + argv, cwd=run, stdout=fo, timeout=3600)
        out = open(os.path.join(run, "out.txt")).read()
        iters = int(re.search(r"Ans (\d+)", out).group(1))
        resmag = [float(x) for x in re.findall(r"At iteration \d+, the mag residue is (\S+)", out)][1:]
        f = last_line(os.path.join(run, "results_phi.txt")).rstrip().rstrip(",").split(",")
        phi = np.array([complex(float(x.split("+i")[0]), float(x.split("+i")[1])) for x in f[1:]])
        phi = phi.reshape(L, L, 2).transpose(1, 0, 2).reshape(L * L, 2)
        head = []
        with open(os.path.join(run, f"Near-null_L{L}_blk{blk}_ndof4.txt")) as fn:
            for _ in range(4096 * 4 * 2):
                a, b = fn.readline().strip().split("+i")
                head.append(float(a) + 1j * float(b))
        rows = open(os.path.join(run, "results_NTL_weights.txt")).read().strip().split("\n")
        w = np.array([[complex(float(x.split("+i")[0]), float(x.split("+i")[1])) for x in r.rstrip(",").split(",")[1:]] for r in rows])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), argv=np.array(argv), stencil=stencil,
                            gauge=np.array([L, 32.0, sweeps, 1234]), theta_checksum=float(np.sum(theta * np.arange(1, theta.size + 1).reshape(theta.shape))),
                            iters=iters, resmag=np.array(resmag), phi_final=phi, last_row_iter=int(f[0]),
                            null0_head=np.array(head).reshape(4096, 4, 2), ntl_weights=w)
        print(f"{name}: iters {iters} final printed residual {resmag[-1]}")


# Cycle pinned without the near-null generation in the loop: the reference is run twice in the same directory, first with
# gen_null = 1 (writes Near-null_*.txt, S6/modules_main.h:62-79), then with gen_null = 0 (reads it back, :39-60).  The
# fixture keeps the file's vectors (25 digits), so oracle / GPU start from the reference's OWN near-null vectors and the
# residual history can be demanded to 1e-9: it is recomputed at full precision from results_res_lvl-0.txt (the residual
# vector the reference writes at the start of every iteration, S6/level.h:266-285) instead of the 6-digit printout.
GN0_CASES = [
    ("gn0_s6_laplace16_2lvl", "laplace", 16, 3, 2, 0.05, 2, 0, 1, 30),
    ("gn0_s6_laplace16_ntl4", "laplace", 16, 3, 2, 0.05, 2, 1, 4, 30),
    ("gn0_s6_wilson16_ntl4", "wilson", 16, 3, 2, 0.05, 2, 1, 4, 30),
]


def make_gn0():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    for name, stencil, L, ni, blk, m, nl, tf, nco, sweeps in GN0_CASES:
        exe = os.path.join(ROOT, "oracle", "_ref", "s6_mgrid_ntl" + ("_laplace" if stencil == "laplace" else ""))
        theta = O.gauge_quenched_phases(L, 32.0, sweeps=sweeps, seed=1234)
        n0, nc = (2, 4) if stencil == "wilson" else (1, 2)
        with tempfile.TemporaryDirectory() as d:
            run = os.path.join(d, "run")
            os.makedirs(run)
            os.makedirs(os.path.join(d, "gauge_config_files"))
            O.write_phase_file(os.path.join(d, "gauge_config_files", f"phase_{L}_b32.0.dat"), theta, L)
            argv = [str(L), str(ni), str(blk), "1", repr(m), str(nl), str(tf), str(nco)]
            subprocess.run([exe] + argv, cwd=run, capture_output=True, text=True, timeout=600)
            argv[3] = "0"
            out = subprocess.run([exe] + argv, cwd=run, capture_output=True, text=True, timeout=600).stdout
            iters = int(re.search(r"Ans (\d+)", out).group(1))
            vals = []
            with open(os.path.join(run, f"Near-null_L{L}_blk{blk}_ndof{nc}.txt")) as f:
                for line in f:
                    a, b = line.strip().split("+i")
                    vals.append(float(a) + 1j * float(b))
            vals = np.array(vals)
            nulls, off, size, ndof = {}, 0, [L // blk ** k for k in range(nl + 1)], [n0] + [nc] * nl
            for lvl in range(nl):
                cnt = size[lvl] ** 2 * ndof[lvl + 1] * ndof[lvl]
                nulls[f"null{lvl}"] = vals[off:off + cnt].reshape(size[lvl] ** 2, ndof[lvl + 1], ndof[lvl])
                off += cnt
            assert off == len(vals)
            # the second run appended to the result files of the first: keep the rows of the LAST run (iteration labels restart at 1)
            rows = parse_cplx_rows(os.path.join(run, "results_res_lvl-0.txt"), L, n0)
            start = max(i for i, (lab, _) in enumerate(rows) if lab == 1)
            rows = rows[start:]
            res_norm = np.array([np.sqrt(np.sum(np.abs(v) ** 2)) for _, v in rows])
            phis = parse_cplx_rows(os.path.join(run, "results_phi.txt"), L, n0)
            w = np.zeros((0, 4), dtype=complex)
            if tf:
                wr = open(os.path.join(run, "results_NTL_weights.txt")).read().strip().split("\n")
                wr = [r for r in wr if r]
                w = np.array([[complex(float(x.split("+i")[0]), float(x.split("+i")[1])) for x in r.rstrip(",").split(",")[1:]] for r in wr])
                w = w[-iters:]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), argv=np.array(argv), stencil=stencil, theta=theta,
                            iters=iters, res_norm=res_norm, phi_final=phis[-1][1], ntl_weights=w, **nulls)
        print(f"{name}: iters {iters} rows {len(res_norm)} final |r| {res_norm[-1]:.3e}")


def parse_cplx_rows(path, L, n):
    rows = open(path).read().strip().split("\n")
    out = []
    for row in rows:
        f = row.rstrip(",").split(",")
        v = np.array([complex(float(x.split("+i")[0]), float(x.split("+i")[1])) for x in f[1:]])
        out.append((int(f[0]), v.reshape(L, L, n).transpose(1, 0, 2).reshape(L * L, n)))   # file: x outer, y inner
    return out


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    for name, stencil, L, ni, blk, m, nl, tf, nco, sweeps in CASES:
        exe = os.path.join(ROOT, "oracle", "_ref", "s6_mgrid_ntl" + ("_laplace" if stencil == "laplace" else ""))
        theta = O.gauge_quenched_phases(L, 32.0, sweeps=sweeps, seed=1234)
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "run"))
            os.makedirs(os.path.join(d, "gauge_config_files"))
            O.write_phase_file(os.path.join(d, "gauge_config_files", f"phase_{L}_b32.0.dat"), theta, L)
            argv = [str(L), str(ni), str(blk), "1", repr(m), str(nl), str(tf), str(nco)]
            out = subprocess.run([exe] + argv, cwd=os.path.join(d, "run"), capture_output=True, text=True, timeout=600).stdout
            iters = int(re.search(r"Ans (\d+)", out).group(1))
            resmag = [float(x) for x in re.findall(r"At iteration \d+, the mag residue is (\S+)", out)][1:]
            n0 = 2 if stencil == "wilson" else 1
            nc = 4 if stencil == "wilson" else 2
            phis = parse_cplx_rows(os.path.join(d, "run", "results_phi.txt"), L, n0)
            null0 = np.zeros((0,), dtype=complex)
            if nl > 0:
                vals = []
                with open(os.path.join(d, "run", f"Near-null_L{L}_blk{blk}_ndof{nc}.txt")) as f:
                    for line in f:
                        a, b = line.strip().split("+i")
                        vals.append(float(a) + 1j * float(b))
                null0 = np.array(vals[: L * L * nc * n0]).reshape(L * L, nc, n0)
            w = np.zeros((0, 4), dtype=complex)
            if tf:
                rows = open(os.path.join(d, "run", "results_NTL_weights.txt")).read().strip().split("\n")
                w = np.array([[complex(float(x.split("+i")[0]), float(x.split("+i")[1])) for x in r.rstrip(",").split(",")[1:]] for r in rows])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), argv=np.array(argv), stencil=stencil,
                            theta=theta, iters=iters, resmag=np.array(resmag), phi_final=phis[-1][1],
                            phi_after_1=phis[1][1] if len(phis) > 2 else phis[-1][1], null0=null0, ntl_weights=w)
        print(f"{name}: iters {iters} final printed residual {resmag[-1] if resmag else None}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--gn0":
        make_gn0()
    elif len(sys.argv) > 1 and sys.argv[1] == "--large":
        make_large(sys.argv[2] if len(sys.argv) > 2 else None)
    else:
        main()
