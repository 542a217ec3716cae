"""Multi-GPU time step: thin ctypes binding of the swe_dist_* C-ABI (include/swe_b200.h).

Everything multi-GPU — decomposition, halo lists, per-rank device contexts, the peer-memory halo
exchange fused into the stage, the global CFL minimum — lives in C++/CUDA behind the C-ABI
(csrc/distplan.cpp, csrc/swe_dist.cuh). Python only launches: it builds a plan, hands the library one
bootstrap primitive (an all-gather of a small blob, here torch.distributed) and calls run/step.

    plan = Plan.struct(rank, world, ni, nj, h)          # or Plan.from_mesh(rank, world, mesh, part)
    ds = DistSolver(plan, device=local_rank)            # collective
    ds.sd.SetVolField(v0_local); ds.exchange()
    ds.run("ssprk2", nsteps, dt=0.0, dt0=1e-3); ds.synchronize()
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .mesh import TriangMesh

HALO_LAYERS = 4   # vertex-adjacent rings of the general decomposition (distplan.cpp)
HALO_ROWS = 3     # rows of squares of the structured strip decomposition


def strip_rows(nj: int, world: int):
    """Row range [j0, j1) of squares owned by each rank (same rule as distplan.cpp strip_rows)."""
    base, rem = divmod(nj, world)
    out, j = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((j, j + n))
        j += n
    return out


class Plan:
    """Host-side decomposition of one rank (swe_dist_plan): local sub-mesh = owned + halo cells."""

    def __init__(self, handle):
        self._h = handle
        l = capi.lib()
        self.mesh = TriangMesh(C.c_void_p(l.swe_dist_plan_local_mesh(self._h)), owner=False)
        nt, ne = self.mesh.nt, self.mesh.ne
        self.n_owned = int(l.swe_dist_plan_owned_count(self._h))
        self.owned = np.ctypeslib.as_array(l.swe_dist_plan_owned(self._h), shape=(nt,)).astype(bool)
        self.global_cells = np.ctypeslib.as_array(l.swe_dist_plan_global_cells(self._h), shape=(nt,)).copy()
        self.classes = np.ctypeslib.as_array(l.swe_dist_plan_classes(self._h), shape=(nt,)).copy()
        self.cfl_mask = np.ctypeslib.as_array(l.swe_dist_plan_cfl_mask(self._h), shape=(ne,)).copy()
        self.peers = []  # [(peer rank, send local ids, recv local ids)]
        for k in range(l.swe_dist_plan_npeers(self._h)):
            pr, ns, nr = C.c_int32(), C.c_int64(), C.c_int64()
            sp, rp = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
            capi.check(l.swe_dist_plan_peer(self._h, k, C.byref(pr), C.byref(ns), C.byref(sp), C.byref(nr), C.byref(rp)))
            send = np.ctypeslib.as_array(sp, shape=(ns.value,)).copy() if ns.value else np.zeros(0, np.int64)
            recv = np.ctypeslib.as_array(rp, shape=(nr.value,)).copy() if nr.value else np.zeros(0, np.int64)
            self.peers.append((int(pr.value), send, recv))

    @classmethod
    def struct(cls, rank: int, world: int, ni: int, nj: int, h: float) -> "Plan":
        hd = C.c_void_p()
        capi.check(capi.lib().swe_dist_plan_struct(C.byref(hd), rank, world, ni, nj, float(h)))
        p = cls(hd)
        p.rank, p.world = rank, world
        return p

    @classmethod
    def from_mesh(cls, rank: int, world: int, mesh: TriangMesh, part: np.ndarray) -> "Plan":
        part = np.ascontiguousarray(part, dtype=np.int32)
        hd = C.c_void_p()
        capi.check(capi.lib().swe_dist_plan_mesh(C.byref(hd), rank, world, mesh._h, part.ctypes.data_as(C.POINTER(C.c_int32))))
        p = cls(hd)
        p.rank, p.world = rank, world
        return p

    def send_list(self) -> np.ndarray:
        return np.concatenate([p[1] for p in self.peers]) if self.peers else np.zeros(0, np.int64)

    def recv_list(self) -> np.ndarray:
        return np.concatenate([p[2] for p in self.peers]) if self.peers else np.zeros(0, np.int64)

    def release(self):
        """The handle now belongs to a swe_dist (swe_dist_create takes ownership)."""
        h, self._h = self._h, None
        return h

    def __del__(self):
        try:
            if self._h:
                self.mesh._h = None
                capi.lib().swe_dist_plan_free(self._h)
                self._h = None
        except Exception:
            pass


def torch_allgather(group=None):
    """The bootstrap primitive for one-process-per-GPU jobs: all-gather of a byte blob over
    torch.distributed (device tensors under NCCL, host tensors under gloo)."""
    import torch
    import torch.distributed as dist

    def fn(_user, send, recv, nbytes):
        try:
            world = dist.get_world_size(group)
            src = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,))
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.from_numpy(src.copy()).to(dev)
            out = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, t, group=group)
            dst = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(world * nbytes,))
            dst[:] = out.cpu().numpy()
            return 0
        except Exception as e:  # never raise through the C frame
            print(f"[swe_fvm_b200.dist] bootstrap all-gather failed: {e}", flush=True)
            return 1
    return capi.ALLGATHER_FN(fn)


class _DistBase:
    def _check(self, rc, d=None):
        if rc != capi.OK:
            msg = capi.lib().swe_dist_last_error(d)
            raise capi.SweError(rc, msg.decode() if msg else "unknown error")


class DistSolver(_DistBase):
    """One rank of a one-process-per-GPU job (swe_dist_create). `sd` is a SpaceDisc view of the rank's
    device context in LOCAL numbering (owned + halo cells) for state transfer, taps and device-side cases."""

    def __init__(self, plan: Plan, device: int = 0, flux="hllc", wavespeed="einfeldt", cor=0.0, tau=0.0, reorder=True,
                 overlap=True, allgather=None, wait_timeout_s: float = 0.0):
        from .solver import SpaceDisc
        self.plan = plan
        self.rank, self.world = plan.rank, plan.world
        self._cb = allgather if allgather is not None else (torch_allgather() if plan.world > 1 else capi.ALLGATHER_FN(0))
        cfg = capi.DistConfig(device=device, reorder=int(reorder), overlap=int(overlap), cor=cor, tau=tau,
                              wait_timeout_s=wait_timeout_s, allgather=self._cb, user=None)
        self._d = C.c_void_p()
        self._check(capi.lib().swe_dist_create(C.byref(self._d), plan._h, C.byref(cfg)))
        plan.release()
        self.sd = SpaceDisc.from_ctx(capi.lib().swe_dist_ctx(self._d), plan.mesh, flux, wavespeed, cor, tau)

    def exchange(self):
        self._check(capi.lib().swe_dist_exchange(self._d), self._d)

    def step(self, scheme, dt: float):
        sc = capi.SCHEMES[scheme.lower()] if isinstance(scheme, str) else int(scheme)
        self._check(capi.lib().swe_dist_step(self._d, sc, self.sd.flux, self.sd.wavespeed, float(dt)), self._d)

    def run(self, scheme, nsteps: int, dt: float = 0.0, dt0: float = 0.0):
        sc = capi.SCHEMES[scheme.lower()] if isinstance(scheme, str) else int(scheme)
        self._check(capi.lib().swe_dist_run(self._d, sc, self.sd.flux, self.sd.wavespeed, int(nsteps), float(dt), float(dt0)), self._d)

    def synchronize(self):
        self._check(capi.lib().swe_dist_synchronize(self._d), self._d)

    def submit_step_host(self, in_ptr: int, out_ptr: int, scheme, dt: float):
        sc = capi.SCHEMES[scheme.lower()] if isinstance(scheme, str) else int(scheme)
        self._check(capi.lib().swe_dist_submit_step_host(
            self._d, C.cast(C.c_void_p(in_ptr), C.POINTER(C.c_double)), C.cast(C.c_void_p(out_ptr), C.POINTER(C.c_double)),
            sc, self.sd.flux, self.sd.wavespeed, float(dt)), self._d)

    def wait_host(self):
        self._check(capi.lib().swe_dist_wait_host(self._d), self._d)

    def cfl_dt(self) -> float:
        v = C.c_double()
        self._check(capi.lib().swe_dist_cfl_dt(self._d, C.byref(v)), self._d)
        return v.value

    def state_hash(self) -> int:
        """This rank's share of the order-independent state hash (sum the ranks' values mod 2**64)."""
        v = C.c_uint64()
        self._check(capi.lib().swe_dist_state_hash(self._d, C.byref(v)), self._d)
        return int(v.value)

    def owned_state(self):
        """(global ids, (n_owned, 3) states) of this rank's owned cells."""
        self.synchronize()
        st = self.sd.GetVolField()
        return self.plan.global_cells[self.plan.owned], st[self.plan.owned]

    def close(self):
        if getattr(self, "_d", None):
            self.sd._ctx = None  # owned by the swe_dist
            capi.lib().swe_dist_destroy(self._d)
            self._d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistGroup(_DistBase):
    """All ranks inside ONE process (swe_dist_group_create): one host thread drives `world` GPUs —
    or, with every rank on the same device, exercises the complete transport (pack-and-signal,
    wait-and-unpack, class-split launches, peer-memory min) on a single GPU."""

    def __init__(self, plans, devices, flux="hllc", wavespeed="einfeldt", cor=0.0, tau=0.0, reorder=True, overlap=True,
                 wait_timeout_s: float = 0.0):
        from .solver import SpaceDisc
        self.plans = plans
        self.world = len(plans)
        cfg = capi.DistConfig(device=0, reorder=int(reorder), overlap=int(overlap), cor=cor, tau=tau,
                              wait_timeout_s=wait_timeout_s, allgather=capi.ALLGATHER_FN(0), user=None)
        arr_p = (C.c_void_p * self.world)(*[p._h for p in plans])
        arr_d = (C.c_int32 * self.world)(*devices)
        self._ds = (C.c_void_p * self.world)()
        self._check(capi.lib().swe_dist_group_create(self._ds, arr_p, arr_d, self.world, C.byref(cfg)))
        for p in plans:
            p.release()
        self.sds = [SpaceDisc.from_ctx(capi.lib().swe_dist_ctx(self._ds[r]), plans[r].mesh, flux, wavespeed, cor, tau)
                    for r in range(self.world)]

    def exchange(self):
        for r in range(self.world):
            self._check(capi.lib().swe_dist_exchange(self._ds[r]), self._ds[r])

    def run(self, scheme, nsteps: int, dt: float = 0.0, dt0: float = 0.0):
        sc = capi.SCHEMES[scheme.lower()] if isinstance(scheme, str) else int(scheme)
        sd = self.sds[0]
        rc = capi.lib().swe_dist_group_run(self._ds, self.world, sc, sd.flux, sd.wavespeed, int(nsteps), float(dt), float(dt0))
        if rc != capi.OK:
            msgs = [capi.lib().swe_dist_last_error(self._ds[r]).decode() for r in range(self.world)]
            raise capi.SweError(rc, "; ".join(m for m in msgs if m))

    def synchronize(self):
        for r in range(self.world):
            self._check(capi.lib().swe_dist_synchronize(self._ds[r]), self._ds[r])

    def cfl_dt(self, r: int = 0) -> float:
        v = C.c_double()
        self._check(capi.lib().swe_dist_cfl_dt(self._ds[r], C.byref(v)), self._ds[r])
        return v.value

    def state_hash(self) -> int:
        tot = 0
        for r in range(self.world):
            v = C.c_uint64()
            self._check(capi.lib().swe_dist_state_hash(self._ds[r], C.byref(v)), self._ds[r])
            tot = (tot + int(v.value)) % (1 << 64)
        return tot

    def owned_states(self):
        self.synchronize()
        out = []
        for r in range(self.world):
            st = self.sds[r].GetVolField()
            out.append((self.plans[r].global_cells[self.plans[r].owned], st[self.plans[r].owned]))
        return out

    def close(self):
        if getattr(self, "_ds", None) is not None:
            for r in range(self.world):
                self.sds[r]._ctx = None
                if self._ds[r]:
                    capi.lib().swe_dist_destroy(self._ds[r])
            self._ds = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
