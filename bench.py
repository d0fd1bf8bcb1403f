#!/usr/bin/env python
"""bench.py -- 2D Wilson adaptive-MG time-to-solution (1e-10) and D-apply HBM GB/s on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (reference arm: the CPU path, see below)

A "step" is one full solve of D x = b (point source, x0 = 0) to |r|/|b| < 1e-10 on the workload lattice with
the hierarchy already set up and all inputs resident in HBM.  `value` is the time to solution in ms
(higher_is_better = false); `e2e` is the same solve through the public API with the right-hand side in pinned
HOST memory and the solution copied back to the host inside the timed region.

Workload (config.workload): BASELINE.json configs[4]/[3] -- 2D U(1) Wilson, quenched beta=6 links generated on
the device, mass = m_crit + 1e-3 (near-critical; m_crit located by MG inverse iteration), adaptive MG with
8 null vectors per chirality-pair (16 coarse dof), 4x4 aggregates, levels L/4^k down to 16, red-black GS
smoother (no pre-smoothing; 4 / 2 / 8 / 8 ... post-smoothing sweeps on levels 0 / 1 / 2 / deeper: measured fastest),
flexible GCR(8) outer iteration,
complex128.

Reference arm / cpu_baseline: the reference (single-threaded C++/Eigen) cannot be built for this workload --
Eigen is absent and its coarse dof count is hard-wired to 4 -- so the CPU side is the oracle PORT running the SAME
algorithm on a bounded sample: the hierarchy of a 256^2 lattice of the same shape is set up by the numpy oracle (not
timed), full solves are run by the plain C + OpenMP restatement of the solve loop (oracle/c_port, all host threads) and
the cost is scaled per site and per iteration to the workload; the JSON says so in `sample`.  The reference's own
binary IS timed on the configuration it can run (BASELINE configs[1], key `config2_vs_reference_binary`).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1.0e-10
METRIC = "wilson_mg_time_to_solution_1e-10"
UNIT = "ms"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed region (the B200_PROFILING.md clocks line), read through NVML
    in-process: forking nvidia-smi from a process that holds a 45 GB CUDA context perturbs the launch thread."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indices follow CUDA_VISIBLE_DEVICES-less enumeration; map through the UUID of the torch device
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None or self.h is None:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, rs))
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(r[0] for r in self.rows)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = sorted({n for r in self.rows for n, b in bits.items() if r[2] & b})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[1] for r in self.rows), default=None),
                "reasons": reasons, "samples": len(self.rows)}


def workload_params(mg2d, L, mass, **kw):
    nlevels = max(1, int(round(math.log(L / 16, 4))))
    return mg2d.make_params(L, mass, nlevels=nlevels, block=4, n_null=8, n_smooth=4, n_pre=0, n_post=post_sweeps(nlevels),
                            smoother="rbgs", null_iters=100, tol=TOL, max_iters=500, **kw)


def post_sweeps(nlevels):
    """Cycle shape measured fastest at 4096^2 (tools/tune.py): no pre-smoothing, 4 red-black sweeps on the fine lattice,
    2 on level 1 (whose 16x16-block operator dominates the traffic of a cycle), 8 on the cheap deeper levels."""
    return ([4, 2] + [8] * nlevels)[:nlevels + 1]


# ---------------------------------------------------------------------------------------------------------
CPU_SAMPLE_L = int(os.environ.get("MG2D_CPU_SAMPLE_L", "256"))   # lattice of the bounded CPU sample (workload's hierarchy shape)
CPU_SAMPLE_ITERS = 6      # outer iterations timed per sample (~7 s of numpy work at 256^2)


def oracle_setup(L_cpu: int):
    """Hierarchy of the numpy oracle on a bounded sample of the workload: same shape (block 4, 16 coarse dof,
    rbgs post-only 4/2/8.., GCR(8)), lattice L_cpu.  Not timed."""
    import copy
    import numpy as np
    from oracle import mg_oracle as O
    nlevels = max(1, int(round(math.log(L_cpu / 16, 4))))
    th = O.gauge_quenched_phases(L_cpu, 6.0, sweeps=20, seed=1234)
    U = O.gauge_from_phases(th)
    po = O.Params(L=L_cpu, num_iters=4, n_pre=0, n_post=post_sweeps(nlevels), block=4, m=-0.05, nlevels=nlevels,
                  stencil="wilson", smoother="rbgs", n_dof_scale=16, null_iters=20)
    LVL, NTL = O.build_reference_problem(po, U)
    O.compute_near_null(LVL, NTL, po, 1)
    b = np.zeros((L_cpu * L_cpu, 2), dtype=complex)
    b[L_cpu // 2 + (L_cpu // 2) * L_cpu, 0] = 1.0
    return {"O": O, "po": po, "LVL": LVL, "NTL": NTL, "b": b, "L": L_cpu, "copy": copy}


def oracle_solve(st, budget_s: float = 8.0):
    """CPU sample: full solves (to 1e-10) of the bounded-sample problem with the C + OpenMP port of the solve loop
    (oracle/c_port, all host threads), repeated until `budget_s` seconds of CPU work are accumulated; falls back to the
    numpy oracle if the C port cannot be built.  Returns (seconds per site*iteration, iterations per solve, seconds, kind)."""
    try:
        from oracle import c_port
        c_port.build()
        tot_s, tot_it, reps = 0.0, 0, 0
        while tot_s < budget_s and reps < 200:
            _, info = c_port.gcr_solve(st["LVL"], st["po"], st["b"], tol=TOL, max_iters=200, restart=8)
            tot_s += info["seconds"]
            tot_it += info["iters"]
            reps += 1
        return tot_s / (tot_it * st["L"] ** 2), tot_it // reps, tot_s, f"C+OpenMP port (oracle/c_port), {reps} full solves"
    except Exception as e:      # no compiler on the box: numpy port
        LVL = st["copy"].deepcopy(st["LVL"])
        t0 = time.perf_counter()
        _, info = st["O"].gcr_MG(LVL, st["NTL"], st["po"], st["b"], tol=TOL, max_iters=CPU_SAMPLE_ITERS, restart=8)
        dt = time.perf_counter() - t0
        return dt / (info["iters"] * st["L"] ** 2), info["iters"], dt, f"numpy port ({type(e).__name__}: C port unavailable), {info['iters']} iterations"


def run_reference_arm(args):
    """--impl reference: the CPU path (oracle port, see module docstring) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = args.L
    cores = os.cpu_count() or 1
    iters_gpu = args.ref_iters
    st = oracle_setup(CPU_SAMPLE_L)
    for _ in range(args.warmup):
        oracle_solve(st, 1.0)
    vals = []
    for _ in range(args.steps):
        per_site_iter, it_done, dt, how = oracle_solve(st, 6.0)
        vals.append(per_site_iter * L * L * iters_gpu * 1e3)
    v = sum(vals) / len(vals)
    sample = (f"{how} on a {CPU_SAMPLE_L}^2 lattice with the workload's hierarchy shape (block 4, 16 coarse dof, red-black post-smoothing "
              f"4/2/8.., FGCR(8), {it_done} iterations per solve), ~6 s of CPU work per step on {cores} threads; scaled per site and "
              f"per iteration to {L}^2 x {iters_gpu} iterations (extrapolated)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": f"wilson{L}_adaptive_mg_near_critical", "L": L, "tol": TOL, "iters": iters_gpu},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def reference_binary_config2(mg2d=None):
    """BASELINE configs[1] (2D U(1) Wilson 64x64, adaptive 3-level MG, the reference's Gauss-Seidel smoother, fp64) run by
    the reference's OWN program -- the unmodified S6 source compiled against oracle/eigen_shim (oracle/_ref/s6_mgrid_ntl),
    wall time of the whole process (setup, 500-sweep near-null generation, solve, its per-iteration text output) -- and by
    this package on the same links with the same algorithm (exact lexicographic GS, stationary cycle, tol 1e-13).
    Returns None when the binary is not present."""
    import re
    import subprocess
    import tempfile
    import numpy as np
    from oracle import mg_oracle as O
    exe = os.path.join(ROOT, "oracle", "_ref", "s6_mgrid_ntl")
    if not os.path.exists(exe):
        return None
    L, m, nl = 64, -0.01, 2
    theta = O.gauge_quenched_phases(L, 32.0, sweeps=30, seed=1234)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "run"))
        os.makedirs(os.path.join(d, "gauge_config_files"))
        O.write_phase_file(os.path.join(d, "gauge_config_files", f"phase_{L}_b32.0.dat"), theta, L)
        t0 = time.perf_counter()
        out = subprocess.run([exe, str(L), "3", "2", "1", repr(m), str(nl), "0", "1"], cwd=os.path.join(d, "run"),
                             capture_output=True, text=True, timeout=900).stdout
        t_ref = time.perf_counter() - t0
    it_ref = int(re.search(r"Ans (\d+)", out).group(1))
    res = {"config": "wilson64_adaptive_3level_gs (BASELINE configs[1]): L=64 m=-0.01 beta=32 block 2, 3 GS sweeps, tol 1e-13",
           "reference_binary_s": t_ref, "reference_iters": it_ref, "cores": 1,
           "note": "reference = unmodified S6 source + oracle/eigen_shim (Eigen absent); wall time of the whole program"}
    if mg2d is not None:
        import torch
        U = torch.as_tensor(O.gauge_from_phases(theta)).cuda()
        p = mg2d.make_params(L, m, nlevels=nl, block=2, n_smooth=3, smoother="gs")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mg = mg2d.setup(U, p, init="reference")
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        x, info = mg2d.solve(mg)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res.update({"gpu_setup_s": t1 - t0, "gpu_solve_s": t2 - t1, "gpu_iters": info["iters"],
                    "gpu_final_residual": info["resnorms"][-1], "iters_identical": info["iters"] == it_ref})
        mg.close()
    return res


# ---------------------------------------------------------------------------------------------------------
def time_kernel(fn, reps, flush=None):
    import torch
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--delta", type=float, default=1e-3, help="mass offset above the estimated critical mass")
    ap.add_argument("--ref-iters", type=int, default=20,
                    help="outer iterations of the workload solve (20 measured by the GPU arm at 4096^2) used to scale the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket ONE extra solve (+ one D-apply) with cudaProfilerStart/Stop for `ncu --profile-from-start off`; "
                         "numbers printed by such a run are not bench values")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import mg2d
    from importlib import import_module
    critical = import_module("2d_multigrid_b200.critical")
    dist_mod = import_module("2d_multigrid_b200.dist")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = dist_mod.init(world, rank, local) if world > 1 else None
    L = args.L
    pk, pk_src = peaks()

    def log(msg):
        if rank == 0:
            print(f"# [{time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    # ---- inputs (not timed): links, critical mass, hierarchy ---------------------------------------------
    t0 = time.time()
    U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=60, seed=1234, device=local)     # Metropolis + plaquette kernels (N1)
    plaq = mg2d.gauge.plaquette_device(U, L).real
    if rank == 0:
        print(f"# links {L}^2 beta=6 plaquette {plaq:.4f} ({time.time()-t0:.1f}s)", file=sys.stderr)
    t0 = time.time()
    if comm is None:
        mcrit, _ = critical.estimate_critical_mass(U, lambda m: workload_params(mg2d, L, m), iters=4, refine=3)
    else:
        mcrit = dist_mod.bcast_float(comm, critical.estimate_critical_mass(
            U, lambda m: workload_params(mg2d, L, m), iters=4, refine=3)[0] if rank == 0 else 0.0)
    mass = mcrit + args.delta
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_crit = time.time() - t0
    p = workload_params(mg2d, L, mass)
    t0 = time.time()
    mg = mg2d.setup(U, p, init="device") if comm is None else dist_mod.setup(U, p, comm)
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    if rank == 0:
        print(f"# m_crit~{mcrit:.5f} ({t_crit:.1f}s)  mass {mass:.5f}  setup {t_setup:.2f}s levels {p.size} dof {p.n_dof}", file=sys.stderr)

    lv0 = mg.LVL[0]
    rhs_host = torch.zeros((L * L, 2), dtype=torch.complex128).pin_memory()
    rhs_host[L // 2 + (L // 2) * L, 0] = 1.0
    rhs = mg.scatter_field(rhs_host) if comm is not None else rhs_host.to(dev)
    x_host = torch.empty((L * L, 2), dtype=torch.complex128).pin_memory()

    def barrier():
        if comm is not None:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def one_solve():
        return mg2d.solve(mg, rhs=rhs, tol=TOL, outer="gcr", restart=8, use_graph=True, check_every=1)

    # ---- device-resident steps ------------------------------------------------------------------------
    log("warm-up solves")
    for _ in range(max(args.warmup, 3)):
        x, info = one_solve()
    barrier()
    log(f"timed solves (iters {info['iters']}, true residual {info.get('true_resnorm')})")
    n0 = mg.launches
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        x, info = one_solve()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    launches = mg.launches - n0
    if comm is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())

    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        one_solve()
        tmp_a, tmp_b = lv0.new_field(), lv0.new_field()
        lv0.apply_D(tmp_b, tmp_a)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        del tmp_a, tmp_b

    log(f"c128 {ms:.1f} ms; mixed-precision leg")
    # ---- the same solve with the V-cycle preconditioner in complex64 (outer GCR / residual stay complex128) ----
    def mixed_solve():
        return mg2d.solve(mg, rhs=rhs, tol=TOL, outer="gcr", restart=8, use_graph=True, check_every=1, precond_dtype="complex64")
    for _ in range(3):
        xm, info_m = mixed_solve()
    barrier()
    e0.record()
    for _ in range(args.steps):
        xm, info_m = mixed_solve()
    e1.record()
    barrier()
    ms_mixed = e0.elapsed_time(e1) / args.steps
    if comm is not None:
        t = torch.tensor([ms_mixed], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_mixed = float(t.item())

    log(f"mixed {ms_mixed:.1f} ms; e2e leg")
    # ---- ... and with that copy's coarse operators stored in half precision (fp32 arithmetic) -------------------
    def half_solve():
        return mg2d.solve(mg, rhs=rhs, tol=TOL, outer="gcr", restart=8, use_graph=True, check_every=1, precond_dtype="complex64+half")
    for _ in range(3):
        xh, info_h = half_solve()
    barrier()
    e0.record()
    for _ in range(args.steps):
        xh, info_h = half_solve()
    e1.record()
    barrier()
    ms_half = e0.elapsed_time(e1) / args.steps
    if comm is not None:
        t = torch.tensor([ms_half], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_half = float(t.item())
    log(f"c64+half {ms_half:.1f} ms ({info_h['iters']} iters, true residual {info_h.get('true_resnorm')})")

    # ---- end to end: host rhs -> H2D -> solve -> D2H solution -----------------------------------------------
    def e2e_solve():
        r = mg.scatter_field(rhs_host) if comm is not None else rhs_host.to(dev, non_blocking=True)
        xx, inf = mg2d.solve(mg, rhs=r, tol=TOL, outer="gcr", restart=8, use_graph=True, check_every=1)
        if comm is not None:
            mg.gather_field(xx, x_host)
        else:
            x_host.copy_(xx, non_blocking=True)
        torch.cuda.synchronize()
        return inf
    e2e_solve()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_solve()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if comm is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    nbytes = L * L * 2 * 16

    log("e2e done; kernel timings")
    # ---- per-kernel rooflines (every rank runs them: strip-level kernels exchange halos; rank 0 reports) -------
    flush = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 256 MiB > 126 MB L2
    hbm = float(pk["hbm_gbs"])
    kern = []
    a, b_ = lv0.new_field(), lv0.new_field()
    a.normal_()
    S0 = lv0.S            # local sites (strip)
    t_apply = time_kernel(lambda: lv0.apply_D(b_, a), 20, flush)
    kern.append(("wilson_march_kernel<double> (D-apply, matrix-free)", 96.0 * S0, t_apply))
    lv0.phi.copy_(a)
    t_rb0 = time_kernel(lambda: lv0.relax(2, smoother="rbgs"), 10, flush) / 2
    kern.append(("wilson_rb2_kernel<double> (one red-black GS sweep, both colours in one pass, level 0)", 128.0 * S0, t_rb0))
    if p.nlevels >= 1:
        l1 = mg.LVL[1]
        n1 = l1.n
        big = l1.S * n1 * n1 * 96 >= 2e8            # operator >> L2: no explicit flush needed
        t_rb1a = time_kernel(lambda: l1.relax(1, smoother="rbgs"), 10, None if big else flush)
        t_rb1b = time_kernel(lambda: l1.relax(3, smoother="rbgs"), 6, None if big else flush)
        t_rb1 = (t_rb1b - t_rb1a) / 2
        kern.append((f"stencil_rb_pm_kernel<double,{n1},1,2> x2 (one red-black GS sweep on pre-multiplied blocks, level 1)",
                     (4 * n1 * n1 + 3 * n1) * 16.0 * l1.S, t_rb1))
        kern.append((f"stencil_rb_pm_kernel<double,{n1},1,1> x2 (first sweep of a relax call: also forms c = D0^-1 r, level 1)",
                     (5 * n1 * n1 + 4 * n1) * 16.0 * l1.S, t_rb1a))
        c, d_ = l1.new_field(), l1.new_field()
        c.normal_()
        t_ap1 = time_kernel(lambda: l1.apply_D(d_, c), 10, flush if l1.S * n1 * n1 * 80 < 2e8 else None)
        kern.append((f"stencil_kernel<double,{n1}> (coarse D-apply, level 1)", (5 * n1 * n1 + 2 * n1) * 16.0 * l1.S, t_ap1))
        t_res = time_kernel(lambda: lv0.restriction(l1.r, a, 1), 10, flush)
        chiral = lv0.phi_null_c is not None          # compacted projector: nc x nf/2 per fine site
        pb = n1 * (1 if chiral else 2)
        kern.append((f"restrict{'_chiral' if chiral else ''}_kernel<double,2,{n1}> (level 0->1)", ((pb + 2) * 16.0 + n1 * 16.0 / 16) * S0, t_res))
        t_pro = time_kernel(lambda: lv0.prolongation(b_, l1.phi, 1), 10, flush)
        kern.append((f"prolong{'_chiral' if chiral else ''}_kernel<double,2,{n1}> (level 1->0, accumulate)", ((pb + 2 * 2) * 16.0 + n1 * 16.0 / 16) * S0, t_pro))
    barrier()
    if rank != 0:
        _finish(comm)
        return
    table = [{"kernel": k, "bytes": by, "ms": t, "gbs": by / t / 1e6, "frac": by / t / 1e6 / hbm} for k, by, t in kern]
    # share of one V-cycle (nu = 4 pre + 4 post sweeps per level) taken by the level-1 smoother, from these timings
    dom = table[2] if len(table) > 2 else table[0]      # level-1 smoother sweep: the largest share of a solve
    dapply = table[0]
    # dram bytes per launch from the committed ncu capture (profiles/traffic_L1024.json) apply to the L=1024 workload
    traffic = None
    traffic_file = os.path.join(ROOT, "profiles", f"traffic_L{L}.json")
    if os.path.exists(traffic_file):
        try:
            tj = {k.replace(" ", ""): v for k, v in json.load(open(traffic_file)).items() if isinstance(v, (int, float))}
            traffic = tj.get(dom["kernel"].split(" ")[0].replace(" ", ""))
            if traffic is not None and " x2 " in dom["kernel"]:
                traffic *= 2          # the timed unit is a full sweep = two half-sweep launches
        except Exception:
            traffic = None

    # ---- CPU baseline (rank 0, N=1 only): bounded oracle sample, scaled ----------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        Lc = CPU_SAMPLE_L
        per_site_iter, it_done, t_cpu, how = oracle_solve(oracle_setup(Lc), 10.0)
        cpu = {"value": per_site_iter * L * L * info["iters"] * 1e3, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": (f"{how}, same algorithm on a {Lc}^2 lattice (block 4, 16 coarse dof, red-black post-smoothing 4/2/8.., FGCR(8), "
                          f"{it_done} iterations per solve), {t_cpu:.1f} s of CPU work; scaled per site x iteration to {L}^2 x "
                          f"{info['iters']} iterations (extrapolated)")}

    config2 = None
    if not args.no_cpu_baseline and world == 1:
        try:
            config2 = reference_binary_config2(mg2d)
        except Exception as e:      # an extra, never allowed to break the bench line
            config2 = {"error": f"{type(e).__name__}: {e}"[:200]}

    line = {
        "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic",
        "config": {"workload": f"wilson{L}_adaptive_mg_near_critical", "L": L, "beta": 6.0, "hbm_gb_in_use": round(torch.cuda.max_memory_allocated() / 1e9, 1), "plaquette": plaq, "mass": mass,
                   "m_crit_est": mcrit, "delta": args.delta, "levels": p.size, "n_dof": p.n_dof, "block": 4, "n_null": 8,
                   "smoother": "rbgs, pre 0, post " + str(p.post), "outer": "fgcr(8)", "tol": TOL, "iters": info["iters"],
                   "final_true_residual": info.get("true_resnorm"), "converged": info["converged"],
                   "setup_s": t_setup, "mcrit_s": t_crit, "cache": "working set >> L2 (126 MB); kernel timings flush L2 with a 256 MiB write",
                   "parallelism": f"strip{world}"},
        "clocks": clocks,
        "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
        "gpu_launches": launches,
        "roofline": {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["gbs"], "peak": hbm, "unit": "GB/s",
                     "frac": dom["frac"], "traffic": traffic, "peak_source": pk_src},
        "dapply": {"kernel": dapply["kernel"], "bound": "hbm", "achieved": dapply["gbs"], "peak": hbm, "unit": "GB/s",
                   "frac": dapply["frac"], "bytes_per_site": 96, "us": dapply["ms"] * 1e3, "target_frac": 0.70},
        "mixed_precision": {"value": ms_mixed, "unit": UNIT, "iters": info_m["iters"], "final_true_residual": info_m.get("true_resnorm"),
                            "note": "same solve, V-cycle preconditioner on a complex64 copy of the hierarchy; outer FGCR, residual and the 1e-10 test in complex128"},
        "mixed_precision_half": {"value": ms_half, "unit": UNIT, "iters": info_h["iters"], "final_true_residual": info_h.get("true_resnorm"),
                                 "note": "as mixed_precision, with the coarse operators of the complex64 copy stored as __half2 (fp32 arithmetic)"},
        "kernels": table,
        "cpu_baseline": cpu,
        "config2_vs_reference_binary": config2,
    }
    print(json.dumps(line), flush=True)
    _finish(comm)


def _finish(comm):
    """Leave without tearing NCCL down under live CUDA graphs that captured its kernels (that hangs): all ranks
    synchronise, flush and exit."""
    if comm is None:
        return
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
