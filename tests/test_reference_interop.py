"""Drop-in interoperability with the UNMODIFIED reference program (oracle/_ref/s6_mgrid_ntl, built from /root/reference
against oracle/eigen_shim): near-null vectors written in the reference's own file format are consumed by the reference
(gen_null = 0, S6/modules_main.h:39-60,189) and it converges in the same number of iterations as our solver run with the
same vectors; and the reference's own near-null file is read back and reproduces its run."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import mg2d
from oracle import mg_oracle as O

L, M, NL = 16, 0.05, 2


def _exe(repo_root):
    path = os.path.join(repo_root, "oracle", "_ref", "s6_mgrid_ntl")
    if not os.path.exists(path) and os.path.isdir("/root/reference"):
        subprocess.call(["make", "-C", os.path.join(repo_root, "oracle"), "_ref/s6_mgrid_ntl"])
    return path if os.path.exists(path) else None


def _run_reference(exe, theta, gen_null, nulls=None):
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "run"))
        os.makedirs(os.path.join(d, "gauge_config_files"))
        mg2d.gauge.write_phase_file(os.path.join(d, "gauge_config_files", f"phase_{L}_b32.0.dat"), theta, L)
        nn = os.path.join(d, "run", mg2d.refio.near_null_filename(L, 2, 4))
        if nulls is not None:
            mg2d.refio.write_near_null(nn, nulls)
        out = subprocess.run([exe, str(L), "3", "2", str(gen_null), repr(M), str(NL), "0", "1"], cwd=os.path.join(d, "run"),
                             capture_output=True, text=True, timeout=300).stdout
        iters = int(re.search(r"Ans (\d+)", out).group(1))
        written = mg2d.refio.read_near_null(nn, [16, 8, 4], [2, 4, 4])
    return iters, written


def test_reference_consumes_our_near_null_vectors(repo_root):
    exe = _exe(repo_root)
    if exe is None:
        pytest.skip("oracle/_ref/s6_mgrid_ntl not built (needs /root/reference)")
    theta = O.gauge_quenched_phases(L, 32.0, sweeps=30, seed=77)
    U = O.gauge_from_phases(theta)
    p = O.Params(L=L, num_iters=3, block=2, m=M, nlevels=NL)
    LVL, NTL, info = O.run_reference_flow(p, U)                       # our (oracle) setup + solve
    nulls = [LVL[l].phi_null for l in range(NL)]
    it_ref, _ = _run_reference(exe, theta, 0, nulls)                  # the reference, fed with those vectors
    # our run from the same supplied vectors (gen_null = 0 path: norm, ortho x2, Galerkin, solve)
    LVL2, NTL2 = O.build_reference_problem(p, U)
    for l in range(NL):
        LVL2[l].phi_null = nulls[l].copy()
    O.compute_near_null(LVL2, NTL2, p, 1, gen_null=0)
    info2 = O.perform_MG(LVL2, NTL2, p)
    assert it_ref == info2["iters"] == info["iters"]


def test_read_reference_near_null_file(repo_root):
    exe = _exe(repo_root)
    if exe is None:
        pytest.skip("oracle/_ref/s6_mgrid_ntl not built (needs /root/reference)")
    theta = O.gauge_quenched_phases(L, 32.0, sweeps=30, seed=78)
    it_ref, written = _run_reference(exe, theta, 1)                   # the reference generates and writes its vectors
    p = O.Params(L=L, num_iters=3, block=2, m=M, nlevels=NL)
    LVL, NTL = O.build_reference_problem(p, O.gauge_from_phases(theta))
    for l in range(NL):
        LVL[l].phi_null = written[l].copy()
    O.compute_near_null(LVL, NTL, p, 1, gen_null=0)
    info = O.perform_MG(LVL, NTL, p)
    assert info["iters"] == it_ref


@pytest.mark.gpu
def test_reference_consumes_gpu_near_null_vectors(repo_root):
    """The GPU-generated near-null vectors, written in the reference format, drive the unmodified reference program to the
    same iteration count as the GPU solve."""
    import torch
    exe = _exe(repo_root)
    if exe is None:
        pytest.skip("oracle/_ref/s6_mgrid_ntl not present on this box")
    theta = O.gauge_quenched_phases(L, 32.0, sweeps=30, seed=79)
    p = mg2d.make_params(L, M, nlevels=NL, block=2, n_smooth=3, smoother="gs")
    mg, info = mg2d.run_reference_flow(p, torch.as_tensor(O.gauge_from_phases(theta)).cuda())
    it_ref, _ = _run_reference(exe, theta, 0, [mg.LVL[l].phi_null for l in range(NL)])
    assert info["converged"] and it_ref == info["iters"]


@pytest.mark.gpu
def test_cli_drop_in(repo_root):
    """`python mg2d.py L num_iters block gen_null m nlevels t_flag n_copies` in a directory laid out like the reference's
    (../gauge_config_files/phase_L_b32.0.dat) prints the same "Ans" as the reference program and writes the same files."""
    import sys
    exe = _exe(repo_root)
    theta = O.gauge_quenched_phases(L, 32.0, sweeps=30, seed=80)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "run"))
        os.makedirs(os.path.join(d, "gauge_config_files"))
        mg2d.gauge.write_phase_file(os.path.join(d, "gauge_config_files", f"phase_{L}_b32.0.dat"), theta, L)
        argv = [str(L), "3", "2", "1", repr(M), str(NL), "0", "1"]
        out = subprocess.run([sys.executable, os.path.join(repo_root, "mg2d.py")] + argv, cwd=os.path.join(d, "run"),
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-1500:]
        ans = int(re.search(r"Ans (\d+)", out.stdout).group(1))
        assert os.path.exists(os.path.join(d, "run", "results_phi.txt"))
        assert os.path.exists(os.path.join(d, "run", mg2d.refio.near_null_filename(L, 2, 4)))
        row = open(os.path.join(d, "run", "results_gen_scaling.txt")).read().split()
        assert int(row[-1]) == ans and int(row[0]) == L

        def rows(name):          # [(label, values)] of a per-iteration result file (S6/level.h:266-300)
            out_ = []
            for line in open(os.path.join(d, "run", name)).read().strip().split("\n"):
                f = line.rstrip(",").split(",")
                out_.append((int(f[0]), np.array([complex(float(x.split("+i")[0]), float(x.split("+i")[1])) for x in f[1:]])))
            return out_
        # per-iteration writers: one row per iteration (label iter+1, written at its start) plus the final state
        names = ["results_phi.txt"] + ["results_res_lvl-%d.txt" % lvl for lvl in range(NL + 1)]
        mine = {n: rows(n) for n in names}
        for n in names:
            assert len(mine[n]) == ans + 1 and [lab for lab, _ in mine[n]] == list(range(1, ans + 1)) + [ans], n
        assert len(mine["results_res_lvl-0.txt"][0][1]) == L * L * 2 and len(mine["results_res_lvl-%d.txt" % NL][0][1]) == (L >> NL) ** 2 * 4
        if exe is not None:
            ref = subprocess.run([exe] + argv, cwd=os.path.join(d, "run"), capture_output=True, text=True, timeout=300).stdout
            assert int(re.search(r"Ans (\d+)", ref).group(1)) == ans
            for n in names:      # the reference has just overwritten the files: same rows, same numbers
                theirs = rows(n)
                assert [lab for lab, _ in theirs] == [lab for lab, _ in mine[n]], n
                for k in (0, 1, min(5, ans - 1), ans):
                    a, b = mine[n][k][1], theirs[k][1]
                    assert np.max(np.abs(a - b)) <= 1e-6 * np.max(np.abs(b)) + 1e-13, (n, k)
