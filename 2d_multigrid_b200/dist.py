"""Multi-GPU domain decomposition: 1-D strips in y over torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference is single-process (no MPI/NCCL anywhere); this is new.  Site index is
s = x + y*L with x fastest (S6/level.h:69-75), so a strip of rows is contiguous and its two halo rows are
contiguous messages:

  * rank p owns rows [p*L/P, (p+1)*L/P) of every DISTRIBUTED level; before each stencil application / red-black
    half sweep it sends its last row to rank p+1 (their `lo` halo) and its first row to rank p-1 (their `hi`
    halo) -- one grouped ncclSend/ncclRecv pair per neighbour, periodic (rank 0 <-> P-1).  The kernels take the
    halo rows as separate pointers (include/mg2d.h), so no padded copies of the fields exist.
  * reductions (residual norms, MR / GCR coefficients, near-null norms) are 1-40 doubles: ncclAllReduce.
  * aggregates never straddle a strip (rows per rank divisible by the block), so restriction, prolongation and
    the per-aggregate orthonormalisation are communication-free; the Galerkin product needs one halo row of P.
  * coarse-level agglomeration: once a level has fewer than `min_rows` rows per rank (halo latency would
    dominate), it and all coarser levels are REPLICATED: the restricted residual strips are all-gathered and
    every rank redundantly runs the (tiny) coarse part of the cycle, then prolongs from its own rows.  No
    scatter, no idle ranks, bit-identical coarse fields on every rank.

The exchange helpers work on CPU tensors with the gloo backend as well, which is how tests/ cover the N>1
host logic without GPUs.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from .mg import MG
from .params import MGParams

MIN_ROWS = int(os.environ.get("MG2D_MIN_ROWS", "32"))
HALO_MODE = os.environ.get("MG2D_HALO", "p2p")      # 'p2p': NVLink peer stores from our own kernels; 'nccl': send/recv
FUSED = os.environ.get("MG2D_FUSED", "1") != "0"    # p2p only: halo push fused into the smoother kernels, reductions
                                                    # summed over ranks inside the producing kernel
LINK_DEBUG = os.environ.get("MG2D_LINK_DEBUG", "")   # 'nowait' / 'nopush': tools/ablate.py timing experiments (wrong results)
SLAB_BYTES = int(os.environ.get("MG2D_P2P_SLAB_MB", "256")) << 20
SLOT_REGION = 1 << 16                                # 1024 halo slots of 64 bytes at the start of the slab
MAILBOX_OFF = 1 << 16                                # all-reduce mailbox (XRedArea) of the rank
DATA_OFF = 1 << 20                                   # halo row buffers


class _Ptr:
    """Minimal stand-in for a tensor where only the address is needed."""

    def __init__(self, ptr: int):
        self._p = ptr

    def data_ptr(self) -> int:
        return self._p


def plan_strips(p: MGParams, world: int, min_rows: int = MIN_ROWS):
    """For every level: (distributed?, rows per rank).  Distribution stops at the first level that cannot be cut
    into >= min_rows rows per rank in whole aggregates; that level and all coarser ones are replicated."""
    plan, alive = [], True
    for lvl, L in enumerate(p.size):
        rows = L // world
        ok = alive and L % world == 0 and rows >= min_rows and (lvl == p.nlevels or rows % p.blocks[lvl] == 0)
        alive = ok
        plan.append((ok, rows if ok else L))
    return plan


class Comm:
    """The ranks of one node as strip neighbours.  With the 'p2p' halo mode every rank owns a CUDA-IPC slab that all
    peers map: 64-byte halo slots, the all-reduce mailbox and, per exchanged field, two pairs (A/B) of halo row buffers.
      * exchange_rows: ONE standalone kernel (mg2d_halo_exchange) per exchange, buffers A;
      * fused_link: the descriptor a smoother kernel takes to push its boundary rows itself (alternating B, A, ...)
        and to wait for the neighbours' rows in its boundary work only -- no exchange launch between half sweeps;
      * reductions: summed over the ranks inside the producing kernel (mg2d_comm_reduce) or by mg2d_allreduce.
    world = 1 (`Comm.single`) makes the rank its own neighbour: the same code paths run on one GPU (tests)."""

    def __init__(self, world: int, rank: int, group=None, backend: str | None = None):
        self.world, self.rank, self.group = world, rank, group
        self.prev, self.next = (rank - 1) % world, (rank + 1) % world
        self._halo = {}
        self.backend = backend if backend is not None else dist.get_backend(group)
        self.p2p = None            # set by enable_p2p()
        self.fused = False
        self.xdesc = None

    @classmethod
    def single(cls, ctx, device, slab_bytes: int = 64 << 20):
        """One rank that is its own strip neighbour (periodic wrap through the halo machinery)."""
        c = cls(1, 0, backend="self")
        c.enable_p2p(ctx, device, slab_bytes)
        return c

    # ---- peer-to-peer path: CUDA-IPC slabs ---------------------------------------------------------------------
    def enable_p2p(self, ctx, device, slab_bytes: int = SLAB_BYTES):
        """Allocate this rank's slab, trade IPC handles, map every peer's slab, build the reduction descriptor."""
        if self.world > 1:
            # CUDA IPC maps memory between processes of ONE node: refuse clearly instead of failing inside cudaIpcOpenMemHandle
            import socket
            hosts = [None] * self.world
            dist.all_gather_object(hosts, socket.gethostname(), group=self.group)
            if len(set(hosts)) != 1:
                from ._lib import MG2DError
                raise MG2DError(f"the peer-to-peer halo path needs all ranks on one node (got {sorted(set(hosts))}); "
                                "use MG2D_HALO=nccl for strips across nodes")
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        ctx.call("mg2d_ipc_alloc", slab_bytes, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p))
        base = {self.rank: ptr.value}
        if self.world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=self.group)
            for peer in range(self.world):
                if peer == self.rank:
                    continue
                hb = (ctypes.c_ubyte * 64)(*allh[peer].cpu().tolist())
                pp = ctypes.c_void_p()
                ctx.call("mg2d_ipc_open", ctypes.cast(hb, ctypes.c_void_p), ctypes.byref(pp))
                base[peer] = pp.value
            dist.barrier(group=self.group)
        self.p2p = {"ctx": ctx, "base": base, "slots": 0, "off": DATA_OFF, "size": slab_bytes, "keys": {}}
        if self.world <= 8:
            boxes = (ctypes.c_void_p * self.world)(*[base[q] + MAILBOX_OFF for q in range(self.world)])
            desc = ctypes.c_void_p()
            ctx.call("mg2d_comm_create", self.world, self.rank, boxes, ctypes.byref(desc))
            self.xdesc = desc.value
            self.fused = FUSED

    def attach(self, ctx):
        """Let another handle of this rank (the complex64 shadow hierarchy) use the same reduction descriptor."""
        if self.xdesc is not None:
            ctx.call("mg2d_comm_attach", self.xdesc)

    def _entry(self, key, nvec: int, row: int):
        """Slot + halo row buffers (lo/hi x A/B) of one exchanged field; identical offsets on every rank."""
        st = self.p2p
        k = (key, nvec, row)
        if k not in st["keys"]:
            need = ((nvec * row + 255) // 256) * 256
            if st["off"] + 4 * need > st["size"] or (st["slots"] + 1) * 64 > SLOT_REGION:
                raise MemoryError("P2P halo slab exhausted: raise MG2D_P2P_SLAB_MB")
            o = st["off"]
            st["keys"][k] = {"slot": st["slots"] * 64, "lo": (o, o + 2 * need), "hi": (o + need, o + 3 * need)}
            st["slots"] += 1
            st["off"] += 4 * need
        return st["keys"][k]

    def _p2p_exchange(self, t, Lx, Ly, width, nvec, key, depth=1):
        from .mg import _stream
        st = self.p2p
        es = t.element_size()
        row = Lx * width * es * depth         # `depth` boundary rows are one contiguous piece
        e = self._entry(key, nvec, row)
        b = st["base"]
        me, pv, nx = b[self.rank], b[self.prev], b[self.next]
        stride = Ly * Lx * width * es
        st["ctx"].call("mg2d_halo_exchange", t.data_ptr(), t.data_ptr() + stride - row, stride, row, nvec,
                       nx + e["lo"][0], pv + e["hi"][0], me + e["slot"], pv + e["slot"], nx + e["slot"], _stream())
        return _Ptr(me + e["lo"][0]), _Ptr(me + e["hi"][0])

    def fused_link(self, t, Lx, Ly, width, nvec, key, depth: int, phase: int, push: bool):
        """For the phase-th kernel after a standalone exchange of the same field (phase 0 reads buffers A):
        returns (mg2d_halo_link, lo pointer, hi pointer).  The kernel reads its halos from buffers phase % 2 and, when
        `push`, stores its boundary rows into the neighbours' buffers (phase + 1) % 2."""
        from ._lib import HaloLink
        st = self.p2p
        row = Lx * width * t.element_size() * depth
        e = self._entry(key, nvec, row)
        b = st["base"]
        me, pv, nx = b[self.rank], b[self.prev], b[self.next]
        rd, wr = phase & 1, (phase + 1) & 1
        if LINK_DEBUG == "nopush":      # timing experiments only (results are wrong): kernels neither push nor wait
            push = False
        link = HaloLink(me + e["slot"], pv + e["slot"], nx + e["slot"],
                        (nx + e["lo"][wr]) if push else None, (pv + e["hi"][wr]) if push else None,
                        {"nowait": 0, "nopush": 0}.get(LINK_DEBUG, 1))
        return link, me + e["lo"][rd], me + e["hi"][rd]

    def gs_links(self, t, Lx, width, key):
        """Pointers for the strip lexicographic Gauss-Seidel sweep (mg2d_relax_gs_strip) on the field whose halo buffers belong to
        `key` (the key of the preceding exchange_rows): progress slots (mine, next rank's, last rank's), the next rank's lo
        buffer and the last rank's hi buffer."""
        st = self.p2p
        row = Lx * width * t.element_size()
        e = self._entry(key, 1, row)
        gs = self._entry(("gs",) + tuple(key), 1, 16)
        b = st["base"]
        last = self.world - 1
        return {"slot_mine": b[self.rank] + gs["slot"], "slot_next": b[self.next] + gs["slot"], "slot_last": b[last] + gs["slot"],
                "push_next_lo": None if self.rank == last else b[self.next] + e["lo"][0],
                "push_last_hi": (b[last] + e["hi"][0]) if self.rank == 0 else None,
                "first": int(self.rank == 0), "last": int(self.rank == last)}

    def check_errors(self):
        """Raise if any exchange / reduction of this rank timed out waiting for a peer (synchronises)."""
        if self.p2p is None:
            return
        from ._lib import MG2DError
        st = self.p2p
        n = ctypes.c_longlong(0)
        torch.cuda.synchronize()
        st["ctx"].call("mg2d_halo_errors", st["base"][self.rank], st["slots"], ctypes.byref(n))
        m = ctypes.c_longlong(0)
        if self.xdesc is not None:
            st["ctx"].call("mg2d_comm_error", self.xdesc, ctypes.byref(m))
        if n.value or m.value:
            raise MG2DError(f"rank {self.rank}: {n.value} halo exchange(s) / {m.value} reduction(s) timed out waiting for a peer; "
                            "the fields of this solve are not valid")

    def p2p_errors(self, device=None) -> int:
        """Number of exchanges that timed out waiting for a neighbour (0 = healthy)."""
        if self.p2p is None:
            return 0
        n = ctypes.c_longlong(0)
        torch.cuda.synchronize()
        self.p2p["ctx"].call("mg2d_halo_errors", self.p2p["base"][self.rank], self.p2p["slots"], ctypes.byref(n))
        return int(n.value)

    def exchange_rows(self, t: torch.Tensor, Lx: int, Ly: int, width: int, nvec: int = 1, key=None, as_tensor: bool = False,
                      depth: int = 1):
        """Returns (lo, hi): the `depth` rows below local row 0 (last rows of rank-1) and above the last local row
        (first rows of rank+1), periodic.  t: [Ly*Lx, width...] or a batch [nvec, Ly*Lx, width]."""
        if self.p2p is not None and not as_tensor and t.is_cuda and (Lx * width * t.element_size()) % 16 == 0:
            return self._p2p_exchange(t, Lx, Ly, width, nvec, key, depth)
        if nvec == 1:
            first, last = t[:depth * Lx], t[(Ly - depth) * Lx:Ly * Lx]
        else:
            first, last = t[:, :depth * Lx].contiguous(), t[:, (Ly - depth) * Lx:Ly * Lx].contiguous()
        k = (key, tuple(first.shape), first.dtype)
        if k not in self._halo:
            self._halo[k] = (torch.empty_like(first), torch.empty_like(first))
        lo, hi = self._halo[k]
        if self.world == 1:
            lo.copy_(last); hi.copy_(first)
        elif self.backend == "nccl":
            # order matters when prev == next (2 ranks): my last row is the peer's lo, my first row its hi
            ops = [dist.P2POp(dist.isend, last, self.next, self.group), dist.P2POp(dist.isend, first, self.prev, self.group),
                   dist.P2POp(dist.irecv, lo, self.prev, self.group), dist.P2POp(dist.irecv, hi, self.next, self.group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        else:
            reqs = [dist.isend(last.contiguous(), self.next, group=self.group, tag=1), dist.isend(first.contiguous(), self.prev, group=self.group, tag=2),
                    dist.irecv(lo, self.prev, group=self.group, tag=1), dist.irecv(hi, self.next, group=self.group, tag=2)]
            for req in reqs:
                req.wait()
        return lo, hi

    def allreduce(self, t: torch.Tensor, op: str = "sum"):
        if self.world == 1:
            return
        if (self.fused and op == "sum" and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= 64):
            from .mg import _stream
            self.p2p["ctx"].call("mg2d_allreduce", t.data_ptr(), t.numel(), _stream())     # our own mailbox kernel
            return
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=self.group)

    def allgather(self, full: torch.Tensor, strip: torch.Tensor):
        """full = concatenation of every rank's strip in rank order (strips are contiguous row blocks)."""
        if self.world == 1:
            full.reshape(-1).copy_(strip.reshape(-1))
        elif self.backend == "nccl":
            dist.all_gather_into_tensor(full.reshape(-1), strip.reshape(-1).contiguous(), group=self.group)
        else:
            self._allgather_list(full, strip)

    def _allgather_list(self, full, strip):
        parts = [torch.empty_like(strip) for _ in range(self.world)]
        dist.all_gather(parts, strip.contiguous(), group=self.group)
        full.reshape(-1).copy_(torch.cat([q.reshape(-1) for q in parts]))


class DistMG(MG):
    """MG whose finest levels are strip-decomposed over the ranks of `comm` (see module docstring)."""

    def __init__(self, params: MGParams, comm: Comm, device: int | None = None, min_rows: int = MIN_ROWS, plan=None):
        super().__init__(params, device)
        self.comm = comm
        self.min_rows = min_rows
        self.plan = plan_strips(params, comm.world, min_rows) if plan is None else plan
        if params.ntl and self.plan[params.nlevels - 1][0]:
            # f_MG_ntl (S6/modules_main.h:386-439) shifts the aggregates of level nlevels-1 by one site per quadrant
            # (f_get_base_site, S6/modules_indiv.h:6-14): on a strip they would straddle the cut.  Supported whenever that level is
            # one of the replicated (agglomerated) ones -- then every rank runs the four copies locally, as on one GPU.
            raise NotImplementedError(f"non-telescoping cycle: level {params.nlevels - 1} (where the shifted copies live) must be "
                                      f"replicated, but {self.plan[params.nlevels - 1][1]} rows per rank keep it striped; "
                                      "raise min_rows / MG2D_MIN_ROWS or use fewer ranks")
        if not self.plan[0][0]:
            raise ValueError(f"lattice {params.L} cannot be cut into {comm.world} strips of >= {min_rows} rows in whole aggregates")
        for lv, (d, rows) in zip(self.LVL, self.plan):
            if d:
                lv.set_strip(comm.rank * rows, rows)
        if HALO_MODE == "p2p" and comm.p2p is None and comm.backend == "nccl" and comm.world > 1:
            comm.enable_p2p(self.ctx, self.device)
        comm.attach(self.ctx)

    # field movement between a full host/device field and the strips
    def scatter_field(self, full: torch.Tensor) -> torch.Tensor:
        lv = self.LVL[0]
        return full[lv.y0 * lv.L:(lv.y0 + lv.Ly) * lv.L].to(self.device, non_blocking=True).to(self.tdtype)

    def gather_field(self, strip: torch.Tensor, host_full: torch.Tensor):
        """Each rank writes its rows of the solution into its host buffer (the distributed end-to-end result)."""
        lv = self.LVL[0]
        host_full[lv.y0 * lv.L:(lv.y0 + lv.Ly) * lv.L].copy_(strip, non_blocking=True)


def init(world: int, rank: int, local: int, backend: str = "nccl") -> Comm:
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, world_size=world, rank=rank, **kw)
    return Comm(world, rank)


def bcast_float(comm: Comm, v: float, src: int = 0) -> float:
    dev = "cuda" if comm.backend == "nccl" else "cpu"
    t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
    dist.broadcast(t, src, group=comm.group)
    return float(t.item())


def setup(U, params: MGParams, comm: Comm, init_fields: str = "device", null_vectors=None, min_rows: int = MIN_ROWS) -> DistMG:
    """Distributed counterpart of the package-level setup(): U is the full link field (every rank keeps its rows)."""
    from .mg import compute_near_null
    mg = DistMG(params, comm, min_rows=min_rows)
    if init_fields == "device":
        mg.init_fields()
    else:
        raise ValueError("distributed setup draws its near-null seeds on the device (init_fields='device')")
    mg.set_gauge(U)
    gen_null = 1
    if null_vectors is not None:
        for lv, P in zip(mg.LVL, null_vectors):
            P = torch.as_tensor(P).to(mg.tdtype)
            if lv.distributed and P.shape[0] == lv.L * lv.L:
                P = P[lv.y0 * lv.L:(lv.y0 + lv.Ly) * lv.L]
            lv.phi_null = P.to(mg.device).contiguous().clone()
        gen_null = 0
    if params.nlevels > 0:
        compute_near_null(mg, params.quad, gen_null)
    elif params.matrix_free:
        mg.LVL[0].D = None
        mg.LVL[0].matrix_free = True
    return mg
