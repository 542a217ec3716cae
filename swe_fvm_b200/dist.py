"""Domain decomposition for the multi-GPU time step: one process per GPU, one sub-mesh per rank.

Every rank owns a set of cells plus a halo deep enough (HALO_LAYERS vertex-adjacent rings) that
one exchange of cell states per RK stage suffices: the new state of an owned cell depends on
reconstructions of its ring-2 cells (fluxes of the neighbours' edges enter their draining dt),
which depend on ring-3 states and, through the node maxima of the part-wet pass, on every cell
sharing a node with those. Halo cells are recomputed redundantly with the SAME
kernels and the SAME global edge orientation, reductions are min/max only, so owned cells come
out bit-identical to the single-GPU run for any number of ranks.

Per step: one halo exchange per stage (send/recv over NCCL -> NVLink; gloo in the CPU tests) and
one scalar min all-reduce for the CFL dt. The backend-specific part is a small "local solver"
object (GpuLocal below; the CPU tests plug the oracle in), everything else is shared.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import StructTriangMesh, TriangMesh

HALO_LAYERS = 4       # vertex-adjacent rings (>= N1(V(N2)) of SURVEY §8e)
HALO_ROWS = 3         # rows of squares for the structured strip decomposition


@dataclass
class Decomposition:
    mesh: TriangMesh                    # local sub-mesh (owned + halo cells)
    rank: int
    world: int
    owned: np.ndarray                   # bool (nt_local)
    global_cells: np.ndarray | None     # int64 (nt_local) or None for strips (implicit)
    peers: list = field(default_factory=list)   # [(peer, send_local_ids, recv_local_ids)] sorted by peer
    # local cell range [a, b) whose reconstruction stencil touches no halo cell (None: unknown); the
    # halo exchange of the previous stage is overlapped with the reconstruction of this range
    interior: tuple | None = None

    @property
    def n_owned(self) -> int:
        return int(self.owned.sum())

    def cfl_edge_mask(self) -> np.ndarray:
        """Edges that touch an owned cell: the only ones whose CFL candidate is valid and needed."""
        et = self.mesh.edge_elements
        m = self.owned[et[:, 0]].copy()
        has = et[:, 1] >= 0
        m[has] |= self.owned[et[has, 1]]
        return m.astype(np.uint8)

    def cell_classes(self) -> np.ndarray:
        """Ordering class per local cell for swe_create_classes: 1 = the reconstruction stencil
        (the cell and its three edge neighbours) contains a halo cell, 0 = it does not, so class 0
        can be reconstructed while the halo exchange of the previous stage is still in flight."""
        halo = np.zeros(self.mesh.nt, dtype=bool)
        if self.peers:
            halo[self.recv_list()] = True
        tt = self.mesh.element_neighbours
        dep = halo.copy()
        for k in range(3):
            j = tt[:, k]
            ok = j >= 0
            dep[ok] |= halo[j[ok]]
        return dep.astype(np.uint8)

    def send_list(self) -> np.ndarray:
        return np.concatenate([p[1] for p in self.peers]) if self.peers else np.zeros(0, np.int64)

    def recv_list(self) -> np.ndarray:
        return np.concatenate([p[2] for p in self.peers]) if self.peers else np.zeros(0, np.int64)


def strip_rows(nj: int, world: int):
    """Row range [j0, j1) of squares owned by each rank (as even as possible)."""
    base, rem = divmod(nj, world)
    out, j = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((j, j + n))
        j += n
    return out


def decompose_strips(ni: int, nj: int, h: float, rank: int, world: int, halo_rows: int = HALO_ROWS) -> Decomposition:
    """Structured strips: rank r owns rows [j0, j1) of the global StructTriangMesh(ni, nj, h) and
    builds its block directly (no global mesh), halo_rows rows of squares on each open side. Local
    cell order = global order restricted, node coordinates bitwise equal to the global mesh's."""
    rows = strip_rows(nj, world)
    j0, j1 = rows[rank]
    if j1 - j0 < halo_rows and world > 1:
        raise ValueError("strip thinner than the halo")
    lo, hi = max(0, j0 - halo_rows), min(nj, j1 + halo_rows)
    mesh = StructTriangMesh(ni, hi - lo, h, 0, lo)
    cells_per_row = 4 * ni
    owned = np.zeros(mesh.nt, dtype=bool)
    owned[(j0 - lo) * cells_per_row:(j1 - lo) * cells_per_row] = True

    def row_cells(ja, jb):  # local ids of global rows [ja, jb)
        return np.arange((ja - lo) * cells_per_row, (jb - lo) * cells_per_row, dtype=np.int64)

    peers = []
    if rank > 0:  # lower neighbour: it needs my first halo_rows rows, I need its last halo_rows rows
        pj0, pj1 = rows[rank - 1]
        peers.append((rank - 1, row_cells(j0, min(j1, j0 + halo_rows)), row_cells(max(pj0, j0 - halo_rows), j0)))
    if rank < world - 1:
        pj0, pj1 = rows[rank + 1]
        peers.append((rank + 1, row_cells(max(j0, j1 - halo_rows), j1), row_cells(j1, min(pj1, j1 + halo_rows))))
    # cells of the halo rows and of the owned row next to them read halo states in K1
    a = (halo_rows + 1) * cells_per_row if rank > 0 else 0
    b = mesh.nt - ((halo_rows + 1) * cells_per_row if rank < world - 1 else 0)
    interior = (a, b) if (world > 1 and b > a) else None
    return Decomposition(mesh, rank, world, owned, None, peers, interior)


def decompose_general(global_mesh: TriangMesh, part: np.ndarray, rank: int, world: int,
                      layers: int = HALO_LAYERS, all_gather_object=None) -> Decomposition:
    """Any mesh / any partition vector. Send lists are the peers' receive lists: every rank
    publishes, per owner, the global ids of the halo cells it wants (all_gather_object)."""
    sub = global_mesh.extract(part, rank, layers)
    gc = np.array(sub.global_cells, dtype=np.int64)
    owner = np.array(sub.cell_owner, dtype=np.int32)
    owned = owner == rank
    want = {int(r): gc[owner == r] for r in np.unique(owner) if r != rank}  # sorted by global id
    if world == 1:
        return Decomposition(sub, rank, world, owned, gc, [])
    if all_gather_object is None:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, want)
    else:
        gathered = all_gather_object(want)
    g2l = {int(g): l for l, g in enumerate(gc)}
    peers = []
    for peer in range(world):
        if peer == rank:
            continue
        recv_g = want.get(peer, np.zeros(0, np.int64))
        send_g = gathered[peer].get(rank, np.zeros(0, np.int64))
        if len(recv_g) == 0 and len(send_g) == 0:
            continue
        send_l = np.array([g2l[int(g)] for g in send_g], dtype=np.int64)
        recv_l = np.array([g2l[int(g)] for g in recv_g], dtype=np.int64)
        peers.append((peer, send_l, recv_l))
    return Decomposition(sub, rank, world, owned, gc, peers)


class HaloExchanger:
    """Packs owned boundary cells, exchanges with every peer, unpacks into the halo cells.
    `local` supplies pack/unpack and a buffer allocator; tensors may be CPU (gloo) or CUDA (NCCL)."""

    def __init__(self, dec: Decomposition, local, transport: str = "nccl"):
        self.dec, self.local = dec, local
        self.nsend = sum(len(p[1]) for p in dec.peers)
        self.nrecv = sum(len(p[2]) for p in dec.peers)
        local.set_halo_lists(dec.send_list(), dec.recv_list())
        self.bytes_per_exchange = 8 * 3 * (self.nsend + self.nrecv)
        self.transport = transport if (dec.peers and getattr(local, "supports_p2p", False)) else "nccl"
        if transport == "p2p" and self.transport != "p2p" and dec.peers:
            raise ValueError("peer-memory halo transport needs a GPU local solver")
        if self.transport == "p2p":
            self._setup_p2p()
        else:
            self.sendbuf = local.alloc(3 * max(self.nsend, 1))
            self.recvbuf = local.alloc(3 * max(self.nrecv, 1))

    def _setup_p2p(self):
        """Peer-memory transport: every rank allocates its receive buffers + flags, publishes their
        CUDA-IPC handles and its receive-segment table; every sender maps the peer's buffers and
        learns where its segment goes. Data then moves by NVLink stores from the pack kernel."""
        import torch.distributed as dist
        dec, L = self.dec, self.local
        handles = L.p2p_alloc(len(dec.peers))
        table, ro = {}, 0
        for slot, (peer, s, r) in enumerate(dec.peers):
            table[int(peer)] = (ro, len(r), slot)
            ro += len(r)
        gathered = [None] * dec.world
        dist.all_gather_object(gathered, {"handles": handles, "table": table})
        so = 0
        for peer, s, r in dec.peers:
            off, cnt, slot = gathered[peer]["table"][dec.rank]
            if cnt != len(s):
                raise RuntimeError("halo lists of neighbouring ranks disagree")
            L.p2p_connect(so, len(s), gathered[peer]["handles"], off, slot)
            so += len(s)
        dist.barrier()

    def _post(self):
        if self.transport == "p2p":
            self.local.p2p_push()   # pack kernels store into the peers' buffers, then publish a flag
            self.local.p2p_pull()   # wait for the peers' flags, unpack
            return
        import torch.distributed as dist
        self.local.pack(self.sendbuf)
        ops, so, ro = [], 0, 0
        for peer, s, r in self.dec.peers:
            if len(s):
                ops.append(dist.P2POp(dist.isend, self.sendbuf[3 * so:3 * (so + len(s))], peer))
            if len(r):
                ops.append(dist.P2POp(dist.irecv, self.recvbuf[3 * ro:3 * (ro + len(r))], peer))
            so += len(s)
            ro += len(r)
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.local.unpack(self.recvbuf)

    def exchange(self):
        """Blocking (in stream order) exchange on the solver's own stream."""
        if self.dec.peers:
            self._post()

    def start(self):
        """Asynchronous exchange on a side stream (CUDA only): pack -> send/recv -> unpack run
        concurrently with whatever the main stream does next; finish() orders the main stream
        after it."""
        if not self.dec.peers:
            return
        self.local.begin_side_stream()
        try:
            self._post()
        finally:
            self.local.end_side_stream()

    def finish(self):
        if self.dec.peers:
            self.local.wait_side_stream()


class DistributedSolver:
    """Solvers::{Euler,SSPRK2,SSPRK3} across ranks: the per-stage pieces of the C-ABI plus one
    halo exchange per stage and one min all-reduce of min_len_to_wavespeed per step."""

    STAGES = {
        0: [(0.0, 1.0, 1.0)],
        1: [(0.0, 1.0, 1.0), (0.5, 0.5, 0.5)],
        2: [(0.0, 1.0, 1.0), (0.75, 0.25, 0.25), (1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0)],
    }

    def __init__(self, dec: Decomposition, local, overlap: bool | None = None, transport: str = "nccl"):
        self.dec, self.local = dec, local
        self.halo = HaloExchanger(dec, local, transport)
        local.set_cfl_edge_mask(dec.cfl_edge_mask())
        self.exchanges = 0
        self.allreduces = 0
        can = bool(dec.peers) and getattr(local, "supports_overlap", False) and getattr(local, "has_classes", False)
        self.overlap = can if overlap is None else (overlap and can)
        self._pending = False

    def _interface_values(self):
        L = self.local
        if not self._pending:
            L.compute_interface_values()
            return
        L.compute_interface_values_class(0, True, False)   # overlaps the halo exchange in flight
        self.halo.finish()
        self._pending = False
        L.compute_interface_values_class(1, False, True)

    def step(self, scheme: int, dt: float | None):
        """dt None: adaptive, dt = 0.15 * global min_len of the previous step's last stage
        (kept on the device; prime it with local.set_dt(dt0))."""
        import torch.distributed as dist
        L = self.local
        stages = self.STAGES[scheme]
        for k, (a0, a1, coef) in enumerate(stages):
            self._interface_values()
            L.compute_fluxes()
            if k == len(stages) - 1 and self.dec.world > 1:
                # global CFL min: only the NEXT step's dt needs it, so on the GPU it runs on the side
                # stream while this stage's draining-dt / update kernels execute
                if getattr(L, "supports_overlap", False):
                    L.begin_side_stream()
                    try:
                        dist.all_reduce(L.min_len_tensor(), op=dist.ReduceOp.MIN)
                    finally:
                        L.end_side_stream(allreduce=True)
                    self._ar_pending = True
                else:
                    dist.all_reduce(L.min_len_tensor(), op=dist.ReduceOp.MIN)
                self.allreduces += 1
            if k == 0 and len(stages) > 1:
                L.save_state()
            L.stage_update(a0, a1, coef, dt)
            if self.overlap:
                self.halo.start()
                self._pending = True
            else:
                self.halo.exchange()
            self.exchanges += 1
        if getattr(self, "_ar_pending", False):
            L.wait_side_stream(allreduce=True)
            self._ar_pending = False
        L.advance_dt(dt)

    def finish(self):
        """Order the main stream after an exchange that is still in flight."""
        if self._pending:
            self.halo.finish()
            self._pending = False

    def run(self, scheme: int, nsteps: int, dt: float | None, dt0: float = 0.0):
        if dt is None:
            self.local.set_dt(dt0)
        for _ in range(nsteps):
            self.step(scheme, dt)
        self.finish()


class GpuLocal:
    """The device context of this rank, driven through the per-stage C-ABI entry points."""

    supports_overlap = True
    supports_p2p = True

    def p2p_alloc(self, npeers: int) -> bytes:
        import ctypes as C
        buf = (C.c_uint8 * 192)()
        self.sd._call("swe_halo_p2p_alloc", int(npeers), buf)
        return bytes(buf)

    def p2p_connect(self, send_start, send_count, peer_handles: bytes, dst_offset, my_slot):
        import ctypes as C
        buf = (C.c_uint8 * 192).from_buffer_copy(peer_handles)
        self.sd._call("swe_halo_p2p_connect", int(send_start), int(send_count), buf, int(dst_offset), int(my_slot))

    def p2p_push(self):
        self.sd._call("swe_halo_p2p_push")

    def p2p_pull(self):
        self.sd._call("swe_halo_p2p_pull")

    def p2p_error(self) -> bool:
        from . import capi
        return capi.lib().swe_halo_p2p_error(self.sd._ctx) == 1

    def __init__(self, sd, has_classes: bool = False):
        import torch
        self.sd = sd
        self.has_classes = has_classes  # sd was created with cell_class = dec.cell_classes()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._minlen = None
        self._main = torch.cuda.current_stream()
        self._side = None
        self._ev_main = torch.cuda.Event()
        self._ev_side = torch.cuda.Event()
        self._ev_ar = torch.cuda.Event()
        self._ctx_mgr = None
        sd.set_stream(self._main.cuda_stream)

    # -- side stream for the overlapped halo exchange --
    def begin_side_stream(self):
        if self._side is None:
            self._side = self.torch.cuda.Stream()
        self._ev_main.record(self._main)
        self._side.wait_event(self._ev_main)
        self._ctx_mgr = self.torch.cuda.stream(self._side)
        self._ctx_mgr.__enter__()
        self.sd.set_stream(self._side.cuda_stream)

    def end_side_stream(self, allreduce: bool = False):
        (self._ev_ar if allreduce else self._ev_side).record(self._side)
        self.sd.set_stream(self._main.cuda_stream)
        self._ctx_mgr.__exit__(None, None, None)
        self._ctx_mgr = None

    def wait_side_stream(self, allreduce: bool = False):
        self._main.wait_event(self._ev_ar if allreduce else self._ev_side)

    def compute_interface_values_class(self, cls, begin, finish):
        self.sd._call("swe_compute_interface_values_class", int(cls), int(begin), int(finish))

    def alloc(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device=self.device)

    def set_halo_lists(self, send, recv):
        import ctypes as C
        s = np.ascontiguousarray(send, dtype=np.int64)
        r = np.ascontiguousarray(recv, dtype=np.int64)
        self.sd._call("swe_halo_set_lists", len(s), s.ctypes.data_as(C.POINTER(C.c_int64)), len(r),
                      r.ctypes.data_as(C.POINTER(C.c_int64)))

    def set_cfl_edge_mask(self, mask):
        import ctypes as C
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        self.sd._call("swe_set_cfl_edge_mask", m.ctypes.data_as(C.POINTER(C.c_uint8)))

    def pack(self, buf):
        import ctypes as C
        self.sd._call("swe_halo_pack", C.c_void_p(buf.data_ptr()))

    def unpack(self, buf):
        import ctypes as C
        self.sd._call("swe_halo_unpack", C.c_void_p(buf.data_ptr()))

    def min_len_tensor(self):
        """torch view of the device scalar min_len_to_wavespeed (all-reduced in place)."""
        if self._minlen is None:
            import ctypes as C
            p = C.c_void_p()
            self.sd._call("swe_min_len_device_ptr", C.byref(p))

            class _Ptr:
                pass
            o = _Ptr()
            o.__cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (p.value, False), "version": 2}
            self._minlen = self.torch.as_tensor(o, device=self.device)
        return self._minlen

    def compute_interface_values(self):
        self.sd.ComputeInterfaceValues()

    def compute_fluxes(self):
        self.sd.ComputeFluxes()

    def save_state(self):
        self.sd._call("swe_save_state")

    def stage_update(self, a0, a1, coef, dt):
        if dt is None:
            self.sd._call("swe_stage_update_dev", a0, a1, coef)
        else:
            self.sd._call("swe_stage_update", a0, a1, coef * dt)

    def set_dt(self, dt):
        self.sd._call("swe_set_dt", float(dt))

    def advance_dt(self, dt):
        self.sd._call("swe_advance_dt", 1 if dt is None else 0, 0.0 if dt is None else float(dt))
