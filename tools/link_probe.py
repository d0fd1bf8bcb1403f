"""What does the strip (halo-linked) code path cost per kernel, on ONE GPU?  A 16-dof level of L x Ly sites with random blocks is
relaxed with 8 red-black sweeps (16 half-sweep kernels) replayed from a CUDA graph, (a) as a whole periodic lattice, (b) as a strip
that is its own neighbour with the fused push/wait, (c) the same without push/wait (MG2D_LINK_DEBUG=nopush is read at import, so
run the script twice for that).  Prints microseconds per half sweep including the gaps between kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg2d
from importlib import import_module
dmod = import_module("2d_multigrid_b200.dist")
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
comm = dmod.Comm.single(mg2d.Context(0), dev, slab_bytes=64 << 20)
NSW = 8


def probe(L, Ly, strip, lowrank=False):
    p = mg2d.make_params(L, 0.1, nlevels=0, matrix_free=False, smoother="rbgs")
    mg = dmod.DistMG(p, comm, min_rows=1) if strip else mg2d.MG(p)
    mg.persistent_sites = 0
    lv = mg.LVL[0]
    n = 16
    lv.n, lv.L, lv.Ly, lv.S = n, L, Ly, L * Ly
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    D = torch.randn((lv.S, 5, n, n, 2), generator=g, dtype=torch.float64, device=dev) * 0.05
    D[:, 0, :, :, 0] += 3.0 * torch.eye(n, device=dev, dtype=torch.float64)
    lv.D = torch.view_as_complex(D).contiguous()
    lv.phi = torch.view_as_complex(torch.randn((lv.S, n, 2), generator=g, dtype=torch.float64, device=dev)).contiguous()
    lv.r = torch.view_as_complex(torch.randn((lv.S, n, 2), generator=g, dtype=torch.float64, device=dev)).contiguous()
    lv.matrix_free = False
    lv.relax(NSW, smoother="rbgs")
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        lv.relax(NSW, smoother="rbgs")
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        lv.relax(NSW, smoother="rbgs")
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 10 / (2 * NSW)
    mb = lv.S / 2 * (4 * n * n + 3 * n) * 16 / 1e6
    print(f"L {L:4d} x Ly {Ly:4d}  {'strip (self-neighbour)' if strip else 'whole lattice         '} {us:8.1f} us per half sweep  ({mb:7.1f} MB -> {mb / us * 1e-3:5.2f} TB/s)"
          f"   MG2D_LINK_DEBUG={os.environ.get('MG2D_LINK_DEBUG', '-')} MG2D_PUBLISH={os.environ.get('MG2D_PUBLISH', '-')}", flush=True)
    mg.close()


for L, Ly in ((256, 256), (256, 128), (256, 32), (1024, 128)):
    probe(L, Ly, False)
    probe(L, Ly, True)
