"""GPU exploration: convergence / timing of the adaptive MG for the BASELINE configs."""
import os, sys, time, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d
from mg2d import gauge
from importlib import import_module
critical = import_module("2d_multigrid_b200.critical")
torch.cuda.set_device(0); dev = torch.device("cuda", 0)

def timed(f):
    torch.cuda.synchronize(); t = time.time(); r = f(); torch.cuda.synchronize(); return r, time.time() - t

for L, beta in ((256, 6.0), (1024, 6.0), (1024, 32.0)):
    th, tg = timed(lambda: gauge.quenched_phases(L, beta, sweeps=200, device="cuda"))
    U = torch.exp(1j * th).to(torch.complex128)
    print(f"\n=== L={L} beta={beta} plaq={gauge.plaquette(U, L).real:.4f} gen {tg:.1f}s")
    blk, nn = 4, 8
    nl = {256: 3, 1024: 4}[L]
    fac = lambda m: mg2d.make_params(L, m, nlevels=nl, block=blk, n_null=nn, n_smooth=4, smoother="mr", null_iters=200, tol=1e-10, max_iters=300)
    (mc, hist), tc = timed(lambda: critical.estimate_critical_mass(U, fac, verbose=True))
    print(f"m_crit ~ {mc:.6f}  ({tc:.1f}s)")
    for delta in (1e-2, 1e-3):
        for (nsm, nulli, nlv) in ((2, 200, nl), (4, 200, nl), (4, 500, nl), (4, 200, nl - 1), (8, 200, nl)):
            p = mg2d.make_params(L, mc + delta, nlevels=nlv, block=blk, n_null=nn, n_smooth=nsm, smoother="mr", null_iters=nulli, tol=1e-10, max_iters=300)
            mg, ts = timed(lambda: mg2d.setup(U, p, init="device"))
            rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=dev); rhs[L // 2 + (L // 2) * L, 0] = 1.0
            (x, info), tsol = timed(lambda: mg2d.solve(mg, rhs=rhs, tol=1e-10, check_every=4))
            print(f"delta={delta:g} nsm={nsm} null={nulli} nlev={nlv}: setup {ts:.2f}s solve {tsol*1e3:.0f} ms iters {info['iters']} conv {info['converged']} res {info['resnorms'][-1]:.2e}")
            del mg
