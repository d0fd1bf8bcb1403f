"""How well is the critical mass known?  Prints the inverse-iteration history and the bench solve at m_crit + delta."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d, bench
from importlib import import_module
critical = import_module("2d_multigrid_b200.critical")
torch.cuda.set_device(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=60, seed=1234, device=0)
t0 = time.time()
mcrit, hist = critical.estimate_critical_mass(U, lambda m: bench.workload_params(mg2d, L, m), verbose=True)
print(f"m_crit {mcrit:.7f}  ({time.time() - t0:.1f} s)  {critical.estimate_critical_mass.info}")
for delta in (1e-3,):
    p = bench.workload_params(mg2d, L, mcrit + delta)
    mg = mg2d.setup(U, p, init="device")
    rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device="cuda"); rhs[L // 2 + (L // 2) * L, 0] = 1.0
    for _ in range(2):
        x, info = mg2d.solve(mg, rhs=rhs, tol=1e-10, outer="gcr", restart=8, use_graph=True)
    torch.cuda.synchronize(); t0 = time.time()
    x, info = mg2d.solve(mg, rhs=rhs, tol=1e-10, outer="gcr", restart=8, use_graph=True)
    torch.cuda.synchronize()
    print(f"delta {delta:g}: mass {mcrit + delta:.6f}  iters {info['iters']}  {1e3 * (time.time() - t0):.1f} ms  true {info['true_resnorm']:.2e}")
    mg.close()
