"""Where does an iteration spend its time INSIDE the CUDA graphs (where per-launch events are not available)?  Ablation: the
same hierarchy, FGCR iterations replayed from graphs with the post-smoothing of one level after the other switched off; the
differences of the per-iteration times are the in-graph costs of those levels (launch gaps, halo waits and all).
    python tools/ablate.py [L]                                   one GPU
    torchrun --nproc-per-node N tools/ablate.py [L]              strips (max over ranks)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d, bench
from importlib import import_module
dmod = import_module("2d_multigrid_b200.dist")
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
comm = dmod.init(world, rank, local) if world > 1 else None
if world == 1 and os.environ.get("MG2D_SELF"):      # the strip code on ONE GPU: the rank is its own neighbour
    comm = dmod.Comm.single(mg2d.Context(local), dev, slab_bytes=256 << 20)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=60, seed=1234, device=local)
p = bench.workload_params(mg2d, L, float(os.environ.get("MG2D_MASS", "-0.06369")))
mg = mg2d.setup(U, p, init="device") if comm is None else dmod.setup(U, p, comm)
rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=dev); rhs[L // 2 + (L // 2) * L, 0] = 1.0
if comm is not None:
    rhs = mg.scatter_field(rhs)
NIT = 16
base = list(p.post)


def timed(post, lazy=True):
    mg.p.post = list(post)
    mg.lazy_gcr = lazy
    kw = dict(rhs=rhs, tol=1e-300, max_iters=NIT, outer="gcr", restart=8, use_graph=True, check_every=1)
    mg2d.solve(mg, **kw)                      # capture + warm
    mg2d.solve(mg, **kw)
    if comm is not None and world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    for _ in range(3):
        done += mg2d.solve(mg, **kw)[1]["executed_iters"]
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / max(done, 1)], dtype=torch.float64, device=dev)
    if comm is not None and world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


rows = [("full cycle " + str(base), base)]
cur = list(base)
for lvl in range(len(base) - 1, -1, -1):
    cur = list(cur); cur[lvl] = 0
    rows.append((f"post sweeps of level {lvl}.. off", cur))
prev = None
for name, post in rows:
    t = timed(post)
    if rank == 0:
        print(f"world {world} L {L}  {name:44s} {t:8.3f} ms/iteration" + ("" if prev is None else f"   (that level: {prev - t:7.3f} ms)"), flush=True)
    prev = t
t = timed(base, lazy=False)
if rank == 0:
    print(f"world {world} L {L}  full cycle, textbook FGCR updates (not lazy)    {t:8.3f} ms/iteration", flush=True)
    print("(the last ablation row = transfers + D-apply + FGCR passes; solves do not converge with sweeps off -- timing only)")
if comm is not None and world > 1:
    torch.cuda.synchronize(); torch.distributed.barrier(); sys.stdout.flush(); os._exit(0)
