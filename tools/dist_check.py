"""2+ GPU check (run under torch.distributed.run): the strip-decomposed solver against the single-GPU solver
on the same links and the same near-null vectors -- coarse operators, iteration count, solution."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d
from importlib import import_module
dmod = import_module("2d_multigrid_b200.dist")

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
comm = dmod.init(world, rank, local)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
min_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16
nl = 2 if L <= 256 else 3
th = mg2d.gauge.quenched_phases(L, 6.0, sweeps=20, device=str(dev)); U = torch.exp(1j * th).to(torch.complex128)
p = mg2d.make_params(L, -0.03, nlevels=nl, block=4, n_null=8, n_smooth=3, smoother="rbgs", null_iters=40, tol=1e-10, max_iters=100)
ref = mg2d.setup(U, p, init="device")                      # every rank builds the same single-GPU hierarchy
rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=dev); rhs[L // 2 + (L // 2) * L, 0] = 1.0
x_ref, i_ref = mg2d.solve(ref, rhs=rhs, tol=1e-10, outer="gcr")
nulls = [lv.phi_null for lv in ref.LVL[:-1]]
dmg = dmod.DistMG(p, comm, min_rows=min_rows)
dmg.init_fields(); dmg.set_gauge(U)
for lv, P in zip(dmg.LVL, nulls):
    lv.phi_null = (P[lv.y0 * lv.L:(lv.y0 + lv.Ly) * lv.L] if lv.distributed else P).contiguous().clone()
mg2d.compute_near_null(dmg, 1, gen_null=0)
worst = 0.0
for a, b in zip(dmg.LVL[1:], ref.LVL[1:]):
    Db = b.D[a.y0 * a.L:(a.y0 + a.Ly) * a.L] if a.distributed else b.D
    worst = max(worst, float((a.D - Db).abs().max()))
x, info = mg2d.solve(dmg, rhs=dmg.scatter_field(rhs), tol=1e-10, outer="gcr")
lv0 = dmg.LVL[0]
dx = float((x - x_ref[lv0.y0 * L:(lv0.y0 + lv0.Ly) * L]).abs().max())
t = torch.tensor([worst, dx], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"L={L} world={world} plan={dmg.plan} coarse-op diff {t[0].item():.2e} iters single {i_ref['iters']} dist {info['iters']} "
          f"true res {info['true_resnorm']:.2e} x diff {t[1].item():.2e}", flush=True)
# generated-null-vector path + timing
dm2 = dmod.setup(U, p, comm)
torch.cuda.synchronize(); dist.barrier()
for ug in (False, True):
    try:
        x2, i2 = mg2d.solve(dm2, rhs=dm2.scatter_field(rhs), tol=1e-10, outer="gcr", use_graph=ug, check_every=4)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
        x2, i2 = mg2d.solve(dm2, rhs=dm2.scatter_field(rhs), tol=1e-10, outer="gcr", use_graph=ug, check_every=4)
        torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t0
        if rank == 0:
            print(f"  own setup, graph={ug}: iters {i2['iters']} conv {i2['converged']} true {i2['true_resnorm']:.2e} solve {dt*1e3:.1f} ms", flush=True)
    except Exception as e:
        if rank == 0:
            print(f"  graph={ug} failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
        break
if rank == 0:
    print(f"  halo mode {dmod.HALO_MODE} p2p active {comm.p2p is not None} p2p timeouts {comm.p2p_errors(dev)}", flush=True)
t0 = time.time(); x1, i1 = mg2d.solve(ref, rhs=rhs, tol=1e-10, outer="gcr", use_graph=True, check_every=4); torch.cuda.synchronize()
t0 = time.time(); x1, i1 = mg2d.solve(ref, rhs=rhs, tol=1e-10, outer="gcr", use_graph=True, check_every=4); torch.cuda.synchronize()
if rank == 0:
    print(f"  single GPU graph solve {1e3*(time.time()-t0):.1f} ms iters {i1['iters']}", flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)   # NCCL teardown under live graphs hangs
