"""Developer smoke script (GPU): every kernel against the numpy oracle, printing the worst deviation."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d
from oracle import mg_oracle as O

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
def T(a): return torch.as_tensor(np.ascontiguousarray(a)).to(dev)
def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
rng = np.random.default_rng(7)
def crand(*s): return rng.normal(size=s) + 1j * rng.normal(size=s)

L = 32; m = -0.01
th = O.gauge_quenched_phases(L, 32.0, sweeps=20); U = O.gauge_from_phases(th)
po = O.Params(L=L, num_iters=3, block=2, m=m, nlevels=2, stencil="wilson", null_iters=40)
p = mg2d.make_params(L, m, nlevels=2, block=2, n_smooth=3, smoother="gs", null_iters=40)

# ---- oracle reference flow pieces
LVLo, NTLo = O.build_reference_problem(po, U)
mg = mg2d.MG(p); mg.init_reference_fields(); mg.set_gauge(T(U))
print("init phi", rel(mg.LVL[0].phi, LVLo[0].phi), "r", rel(mg.LVL[0].r, LVLo[0].r), "P", rel(mg.LVL[0].phi_null, LVLo[0].phi_null))
print("lvl0 D", rel(mg2d.D_to_reference_layout(mg.LVL[0].D), LVLo[0].D))
v = crand(L * L, 2); vo = LVLo[0].apply_D(v, L)
out = torch.empty_like(T(v)); mg.LVL[0].apply_D(out, T(v)); print("stencil apply n=2", rel(out, vo))
mg.LVL[0].matrix_free = True
out2 = torch.empty_like(out); mg.LVL[0].apply_D(out2, T(v)); print("wilson apply", rel(out2, vo))
d = mg.LVL[0].dots("t"); mg.LVL[0]._stencil(out2, T(v), None, 0, d)
dd = d[:4].cpu().numpy(); print("wilson dots", dd[0] / np.sum(np.abs(vo) ** 2) - 1, (dd[1] + 1j * dd[2]) / np.vdot(vo, v) - 1)
print("resmag wilson", mg.LVL[0].get_residue_mag() / LVLo[0].get_residue_mag(L) - 1)
mg.LVL[0].matrix_free = False
print("resmag stencil", mg.LVL[0].get_residue_mag() / LVLo[0].get_residue_mag(L) - 1)
# c64
p64 = mg2d.make_params(L, m, nlevels=2, dtype="complex64"); mg64 = mg2d.MG(p64); mg64.init_reference_fields(); mg64.set_gauge(T(U))
o64 = torch.empty((L * L, 2), dtype=torch.complex64, device=dev)
mg64.LVL[0].apply_D(o64, T(v).to(torch.complex64)); print("c64 stencil", rel(o64, vo))
mg64.LVL[0].matrix_free = True; mg64.LVL[0].apply_D(o64, T(v).to(torch.complex64)); print("c64 wilson", rel(o64, vo))

# ---- relaxations
import copy
for sm, gs in (("gs", 1), ("jacobi", 0)):
    lo = copy.deepcopy(LVLo[0]); lo.relax(L, 2, gs)
    phi0 = mg.LVL[0].phi.clone(); mg.LVL[0].relax(2, gs); print("relax", sm, rel(mg.LVL[0].phi, lo.phi)); mg.LVL[0].phi.copy_(phi0)
lo = copy.deepcopy(LVLo[0]); lo.relax_mr(L, 3)
phi0 = mg.LVL[0].phi.clone(); mg.LVL[0].relax(3, smoother="mr"); print("relax mr", rel(mg.LVL[0].phi, lo.phi)); mg.LVL[0].phi.copy_(phi0)
mg.LVL[0].matrix_free = True
mg.LVL[0].relax(3, smoother="mr"); print("relax mr (matrix-free)", rel(mg.LVL[0].phi, lo.phi)); mg.LVL[0].phi.copy_(phi0)
mg.LVL[0].matrix_free = False

# ---- setup
t0 = time.time(); O.compute_near_null(LVLo, NTLo, po, 1); print("oracle setup s", time.time() - t0)
t0 = time.time(); mg2d.compute_near_null(mg); torch.cuda.synchronize(); print("gpu setup s", time.time() - t0, mg.info)
for l in range(2):
    print("P", l, rel(mg.LVL[l].phi_null, LVLo[l].phi_null), "D", l + 1, rel(mg2d.D_to_reference_layout(mg.LVL[l + 1].D), LVLo[l + 1].D))
for l in range(2):
    nf, nc = po.n_dof[l], po.n_dof[l + 1]
    vf = crand(po.size[l] ** 2, nf); vc = crand(po.size[l + 1] ** 2, nc)
    for quad in (1, 2, 3, 4):
        ro = LVLo[l].restriction(vf, l, po, quad)
        rc = torch.empty_like(T(vc)); mg.LVL[l].restriction(rc, T(vf), quad)
        fo = vf.copy(); LVLo[l].prolongation(fo, vc, l + 1, po, quad)
        fg = T(vf).clone(); mg.LVL[l].prolongation(fg, T(vc), quad)
        print("lvl", l, "quad", quad, "restrict", rel(rc, ro), "prolong", rel(fg, fo))
    vv = crand(po.size[l + 1] ** 2, nc); oo = LVLo[l + 1].apply_D(vv, po.size[l + 1])
    og = torch.empty_like(T(vv)); mg.LVL[l + 1].apply_D(og, T(vv)); print("coarse apply", l + 1, rel(og, oo))

# ---- full solves
for sm in ("gs", "mr", "jacobi"):
    for ntl in (False, True):
        if sm == "jacobi" and ntl: continue
        nl = 3 if ntl else 2
        po2 = O.Params(L=L, num_iters=3, block=2, m=m, nlevels=nl, stencil="wilson", null_iters=40, smoother=sm, t_flag=int(ntl), n_copies=4, max_iters=400)
        t0 = time.time(); _, _, io = O.run_reference_flow(po2, U); to = time.time() - t0
        p2 = mg2d.make_params(L, m, nlevels=nl, block=2, n_smooth=3, smoother=sm, null_iters=40, ntl=ntl, n_copies=4, max_iters=400, matrix_free=False)
        t0 = time.time(); mgg, ig = mg2d.run_reference_flow(p2, T(U)); tg = time.time() - t0
        k = min(len(io["resnorms"]), len(ig["resnorms"]))
        dev_r = max(abs(a / b - 1) for a, b in zip(ig["resnorms"][:k], io["resnorms"][:k]))
        print(f"solve {sm} ntl={ntl}: oracle iters {io['iters']} ({to:.1f}s) gpu iters {ig['iters']} ({tg:.1f}s) conv {ig['converged']} max resnorm rel dev {dev_r:.2e}")
        if ntl: print("   weights", io["ntl_weights"][-1], ig["ntl_weights"][-1])

# ---- laplace
pl = O.Params(L=L, num_iters=3, block=2, m=0.05, nlevels=2, stencil="laplace", null_iters=40, max_iters=400)
_, _, io = O.run_reference_flow(pl, U)
p3 = mg2d.make_params(L, 0.05, stencil="laplace", nlevels=2, n_smooth=3, null_iters=40, max_iters=400)
_, ig = mg2d.run_reference_flow(p3, T(U)); print("laplace solve iters", io["iters"], ig["iters"])

# ---- timing of the Wilson apply
for Lb in (1024, 4096):
    pb = mg2d.make_params(Lb, 0.01, nlevels=0, smoother="mr")
    mb = mg2d.MG(pb)
    Ub = torch.exp(1j * torch.randn(Lb * Lb, 2, dtype=torch.float64, device=dev) * 0.2).to(torch.complex128)
    lv = mb.LVL[0]; lv.compute_lvl0_matrix(Ub, store=False)
    a = torch.randn(Lb * Lb, 2, dtype=torch.complex128, device=dev); b = torch.empty_like(a)
    for _ in range(3): lv.apply_D(b, a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lv.apply_D(b, a); lv.apply_D(a, b)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 40
    print(f"wilson apply L={Lb}: {ms*1e3:.1f} us, {Lb*Lb*96/ms/1e6:.0f} GB/s")
print("launches", mg.ctx.launches)
