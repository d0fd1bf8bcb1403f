"""rbgs / gcr / cuda-graph parity vs oracle + convergence exploration with the rbgs smoother."""
import os, sys, time, copy
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d
from mg2d import gauge
from oracle import mg_oracle as O
from importlib import import_module
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
def T(a): return torch.as_tensor(np.ascontiguousarray(a)).to(dev)
def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
def timed(f):
    torch.cuda.synchronize(); t = time.time(); r = f(); torch.cuda.synchronize(); return r, time.time() - t

L = 32; m = 0.02
th = O.gauge_quenched_phases(L, 6.0, sweeps=20); U = O.gauge_from_phases(th)
for mf in (False, True):
    po = O.Params(L=L, num_iters=2, block=2, m=m, nlevels=2, stencil="wilson", null_iters=40, smoother="rbgs", max_iters=200)
    LVLo, NTLo, io = O.run_reference_flow(po, U)
    p = mg2d.make_params(L, m, nlevels=2, n_smooth=2, smoother="rbgs", null_iters=40, max_iters=200, matrix_free=mf)
    mgg, ig = mg2d.run_reference_flow(p, T(U))
    k = min(len(io["resnorms"]), len(ig["resnorms"]))
    print(f"rbgs mf={mf}: oracle {io['iters']} gpu {ig['iters']} dev {max(abs(a/b-1) for a,b in zip(ig['resnorms'][:k], io['resnorms'][:k])):.1e} phi {rel(mgg.LVL[0].phi, LVLo[0].phi):.1e}")
# graph vs eager
p = mg2d.make_params(L, m, nlevels=2, n_smooth=2, smoother="rbgs", null_iters=40, max_iters=200)
mg1 = mg2d.setup(T(U), p); x1, i1 = mg2d.solve(mg1)
mg2 = mg2d.setup(T(U), p); x2, i2 = mg2d.solve(mg2, use_graph=True, check_every=4)
print("graph vs eager iters", i1["iters"], i2["iters"], "phi diff", rel(x2, x1.cpu().numpy()))
# gcr parity
b = np.zeros((L * L, 2), dtype=complex); b[L // 2 + L // 2 * L, 0] = 1
for sm in ("rbgs", "mr"):
    po = O.Params(L=L, num_iters=2, block=2, m=m, nlevels=2, stencil="wilson", null_iters=40, smoother=sm)
    LVLo, NTLo = O.build_reference_problem(po, U); O.compute_near_null(LVLo, NTLo, po, 1)
    xo, io = O.gcr_MG(LVLo, NTLo, po, b, tol=1e-10, restart=4)
    p = mg2d.make_params(L, m, nlevels=2, n_smooth=2, smoother=sm, null_iters=40)
    mgg = mg2d.setup(T(U), p)
    for ug in (False, True):
        xg, ig = mg2d.solve(mgg, rhs=T(b), tol=1e-10, outer="gcr", restart=4, use_graph=ug)
        k = min(len(io["resnorms"]), len(ig["resnorms"]))
        print(f"gcr {sm} graph={ug}: oracle {io['iters']} gpu {ig['iters']} true {ig['true_resnorm']:.2e} dev {max(abs(a/b_-1) for a,b_ in zip(ig['resnorms'][:k], io['resnorms'][:k])):.1e} x {rel(xg, xo):.1e}")

critical = import_module("2d_multigrid_b200.critical")
for L, beta in ((256, 6.0), (1024, 6.0)):
    thg = gauge.quenched_phases(L, beta, sweeps=200, device="cuda"); Ug = torch.exp(1j * thg).to(torch.complex128)
    nl = {256: 3, 1024: 4}[L]
    fac = lambda mm: mg2d.make_params(L, mm, nlevels=nl, block=4, n_null=8, n_smooth=4, smoother="rbgs", null_iters=100)
    (mc, hist), tc = timed(lambda: critical.estimate_critical_mass(Ug, fac, verbose=True))
    print(f"=== L={L} beta={beta} m_crit ~ {mc:.6f} ({tc:.1f}s)")
    rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=dev); rhs[L // 2 + (L // 2) * L, 0] = 1.0
    for delta in (1e-2, 1e-3):
        for nsm, nulli, nn, blk, nlv in ((2, 100, 8, 4, nl), (4, 100, 8, 4, nl), (4, 500, 8, 4, nl), (2, 100, 4, 4, nl), (2, 100, 2, 2, 2 * nl - 1)):
            p = mg2d.make_params(L, mc + delta, nlevels=nlv, block=blk, n_null=nn, n_smooth=nsm, smoother="rbgs", null_iters=nulli, tol=1e-10, max_iters=400)
            mgg, ts = timed(lambda: mg2d.setup(Ug, p, init="device"))
            for outer in ("stationary", "gcr"):
                (x, info), tsol = timed(lambda: mg2d.solve(mgg, rhs=rhs, tol=1e-10, check_every=4, outer=outer, use_graph=True))
                (x, info), tsol = timed(lambda: mg2d.solve(mgg, rhs=rhs, tol=1e-10, check_every=4, outer=outer, use_graph=True))
                print(f"delta={delta:g} nsm={nsm} null={nulli} nn={nn} blk={blk} nlev={nlv} {outer}: setup {ts:.2f}s solve {tsol*1e3:.0f} ms iters {info['iters']} conv {info['converged']} res {info['resnorms'][-1]:.2e}", flush=True)
            del mgg
