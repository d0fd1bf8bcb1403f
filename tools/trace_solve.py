"""Where does a solve spend its device time?  Runs the bench workload eagerly (no graphs) with every entry point of the C ABI
bracketed by CUDA events (MG2D trace mode) and prints device milliseconds per entry point and per iteration.
    python tools/trace_solve.py [L]                                 one GPU
    torchrun --nproc-per-node N tools/trace_solve.py [L]            strips (rank 0 reports; times include waiting for peers)"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d, bench
from importlib import import_module
dmod = import_module("2d_multigrid_b200.dist")
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
comm = dmod.init(world, rank, local) if world > 1 else None
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
U = mg2d.gauge.quenched_links_device(L, 6.0, sweeps=60, seed=1234, device=local)
mass = float(os.environ.get("MG2D_MASS", "-0.06369"))
p = bench.workload_params(mg2d, L, mass)
mg = mg2d.setup(U, p, init="device") if comm is None else dmod.setup(U, p, comm)
rhs = torch.zeros((L * L, 2), dtype=torch.complex128, device=dev); rhs[L // 2 + (L // 2) * L, 0] = 1.0
if comm is not None:
    rhs = mg.scatter_field(rhs)
for _ in range(2):
    x, info = mg2d.solve(mg, rhs=rhs, tol=1e-10, outer="gcr", restart=8, use_graph=False)
torch.cuda.synchronize()
if comm is not None:
    torch.distributed.barrier()
mg2d._lib.trace_begin()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
x, info = mg2d.solve(mg, rhs=rhs, tol=1e-10, outer="gcr", restart=8, use_graph=False)
e1.record()
rep = mg2d._lib.trace_report()
tot = e0.elapsed_time(e1)
if rank == 0:
    it = info["iters"]
    print(f"L={L} world={world} iters {it} eager solve {tot:.1f} ms = {tot/it:.3f} ms/iteration; sum of traced launches {sum(t for _, t in rep.values()):.1f} ms")
    for name, (c, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:28s} {c:6d} calls {t:9.2f} ms  {t/it:8.3f} ms/iter  {1e3*t/c:8.1f} us/call")
if comm is not None:
    torch.cuda.synchronize(); torch.distributed.barrier(); sys.stdout.flush(); os._exit(0)
