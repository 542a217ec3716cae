"""Host-side mirror of the reference's SpaceDisc / TimeDisc / Solvers API over the C-ABI.

    sd = SpaceDisc("hllc", "einfeldt", mesh, v0, cor=0, tau=0)   # include/SpaceDisc.h:24
    td = TimeDisc(sd)                                              # include/TimeDisc.h:6
    Solvers.SSPRK2(td, dt)                                         # include/Solvers.h:7
    td.CFLdt()                                                     # include/TimeDisc.h:13

The reference plugs the flux in as a std::function (include/SpaceDisc.h:22); a std::function
cannot run on the device, so the known fluxers are selected by name from the compile-time
registry {HLL, HLLC} x {Rusanov, Davis, Einfeldt} (include/Fluxes.h, src/Fluxes.cpp).
All state lives in HBM; numpy arrays cross the boundary only in set/get calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .mesh import TriangMesh


class SpaceDisc:
    def __init__(self, flux: str | int, wavespeed: str | int, mesh: TriangMesh, v0: np.ndarray | None = None,
                 cor: float = 0.0, tau: float = 0.0, device: int = 0, reorder: bool = False, taps: bool = False,
                 cell_class: np.ndarray | None = None):
        self.flux = capi.FLUXES[flux.lower()] if isinstance(flux, str) else int(flux)
        self.wavespeed = capi.WAVESPEEDS[wavespeed.lower()] if isinstance(wavespeed, str) else int(wavespeed)
        self.mesh = mesh
        self.nt, self.ne, self.nn = mesh.nt, mesh.ne, mesh.nn
        self.cor, self.tau = float(cor), float(tau)
        self._ctx = C.c_void_p()
        cm = mesh.c_mesh(cor, tau)
        if cell_class is None:
            capi.check(capi.lib().swe_create(C.byref(self._ctx), C.byref(cm), device, int(reorder)))
        else:
            cc = np.ascontiguousarray(cell_class, dtype=np.uint8)
            if cc.shape != (self.nt,):
                raise ValueError("cell_class must have one entry per cell")
            capi.check(capi.lib().swe_create_classes(C.byref(self._ctx), C.byref(cm), device, int(reorder),
                                                     cc.ctypes.data_as(C.POINTER(C.c_uint8))))
        if taps:
            self._call("swe_enable_taps", 1)
        if v0 is not None:
            self.SetVolField(v0)

    @classmethod
    def from_ctx(cls, ctx_ptr, mesh, flux="hllc", wavespeed="einfeldt", cor=0.0, tau=0.0) -> "SpaceDisc":
        """View of a device context owned by someone else (a swe_dist rank); never destroys it."""
        self = cls.__new__(cls)
        self.flux = capi.FLUXES[flux.lower()] if isinstance(flux, str) else int(flux)
        self.wavespeed = capi.WAVESPEEDS[wavespeed.lower()] if isinstance(wavespeed, str) else int(wavespeed)
        self.mesh = mesh
        self.nt, self.ne, self.nn = mesh.nt, mesh.ne, mesh.nn
        self.cor, self.tau = float(cor), float(tau)
        self._ctx = C.c_void_p(ctx_ptr)
        self._borrowed = True
        return self

    def state_hash(self) -> int:
        v = C.c_uint64()
        self._call("swe_state_hash", C.byref(v))
        return int(v.value)

    # -- plumbing --
    def _call(self, name, *args):
        capi.check(getattr(capi.lib(), name)(self._ctx, *args), self._ctx)

    def _get(self, name, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        ptr = out.ctypes.data_as(C.POINTER(C.c_double if dtype == np.float64 else C.c_int8))
        self._call(name, ptr)
        return out

    def set_stream(self, cuda_stream_ptr: int):
        self._call("swe_set_stream", C.c_void_p(cuda_stream_ptr))

    def synchronize(self):
        self._call("swe_synchronize")

    # -- reference accessors --
    def SetVolField(self, prim):
        p = capi.as_f64(prim, (self.nt, 3))
        self._call("swe_set_state", capi.dptr(p))

    def GetVolField(self) -> np.ndarray:
        """(nt, 3) primitive (w, u, v), the layout of Storage<3>."""
        return self._get("swe_get_state", (self.nt, 3))

    def set_state_async(self, pinned_ptr: int):
        self._call("swe_set_state_async", C.cast(C.c_void_p(pinned_ptr), C.POINTER(C.c_double)))

    def get_state_async(self, pinned_ptr: int):
        self._call("swe_get_state_async", C.cast(C.c_void_p(pinned_ptr), C.POINTER(C.c_double)))

    def submit_step_host(self, in_ptr: int, out_ptr: int, scheme, dt: float):
        """Host-buffer pipeline (swe_submit_step_host): upload in_ptr -> one step -> download into out_ptr, asynchronous."""
        sc = capi.SCHEMES[scheme.lower()] if isinstance(scheme, str) else int(scheme)
        self._call("swe_submit_step_host", C.cast(C.c_void_p(in_ptr), C.POINTER(C.c_double)),
                   C.cast(C.c_void_p(out_ptr), C.POINTER(C.c_double)), sc, self.flux, self.wavespeed, float(dt))

    def wait_host(self):
        self._call("swe_wait_host")

    def ComputeInterfaceValues(self):
        self._call("swe_compute_interface_values")

    def ComputeFluxes(self):
        self._call("swe_compute_fluxes", self.flux, self.wavespeed)

    def GetEdgField(self) -> np.ndarray:
        return self._get("swe_get_edge_states", (2 * self.ne, 3))

    def GetSrcField(self) -> np.ndarray:
        return self._get("swe_get_sources", (2 * self.ne, 3))

    def GetFluxes(self) -> np.ndarray:
        return self._get("swe_get_fluxes", (self.ne, 3))

    def GetMinLenToWavespeed(self) -> float:
        v = C.c_double()
        self._call("swe_get_min_len_to_wavespeed", C.byref(v))
        return v.value

    def GetCor(self):
        return self.cor

    def GetTau(self):
        return self.tau

    def node_max_w(self) -> np.ndarray:
        return self._get("swe_get_node_max_w", (self.nn,))

    def draining_dt(self) -> np.ndarray:
        return self._get("swe_get_draining_dt", (self.nt,))

    def cell_class(self) -> np.ndarray:
        return self._get("swe_get_cell_class", (self.nt,), dtype=np.int8)

    BRANCHES = ("pw1_submerged", "pw1_cbrt", "pw1_bisection", "fw_dry_neighbour", "fw_partwet_neighbour", "fw_vertex_zeroed",
                "fw_tvd_off", "pw2_to_pw1", "pw2_one_wet", "pw2_three_wet", "pw2_two_wet", "pw2_two_wet_fallback")

    @staticmethod
    def registered_fluxes() -> dict:
        """{name: id} of the compile-time flux registry (csrc/swe_flux_registry.cuh)."""
        l = capi.lib()
        return {l.swe_fluxer_name(k).decode(): int(l.swe_fluxer_id(k)) for k in range(l.swe_fluxer_count())}

    def set_fluxer(self, name_or_id):
        """Select a registered flux by name or id for every following step (overrides flux / wavespeed)."""
        fid = name_or_id if isinstance(name_or_id, int) else capi.lib().swe_fluxer_find(str(name_or_id).encode())
        if fid < 0 and not isinstance(name_or_id, int):
            raise KeyError(f"flux {name_or_id!r} is not registered: {sorted(self.registered_fluxes())}")
        self._call("swe_set_fluxer", int(fid))

    def set_option(self, key: str, value: int):
        """Semantic-decision switches: recon (0 repaired | 1 as written | 2 first order), pw2, roe_fix, cfl_abs."""
        self._call("swe_set_option", key.encode(), int(value))

    def get_option(self, key: str) -> int:
        v = C.c_int32()
        self._call("swe_get_option", key.encode(), C.byref(v))
        return v.value

    def branch_counts(self) -> dict:
        """How many cells took each reconstruction branch in the last ComputeInterfaceValues (taps on)."""
        out = np.zeros(12, dtype=np.int64)
        self._call("swe_get_branch_counts", out.ctypes.data_as(C.POINTER(C.c_int64)))
        return dict(zip(self.BRANCHES, out.tolist()))

    def diagnostics(self) -> dict:
        out = self._get("swe_diagnostics", (6,))
        return dict(mass=out[0], kinetic=out[1], potential=out[2], vmax=out[3], hmin=out[4], wet_cells=int(out[5]))

    # -- analytic cases on the device (initial state / error norms without host loops) --
    def set_case_bathymetry(self, case):
        self._call("swe_case_set_bathymetry_device", C.byref(case.c))

    def set_case_state(self, case, quad_n: int = 4, t: float = 0.0):
        self._call("swe_case_initial_state_device", C.byref(case.c), int(quad_n), float(t))

    def case_l2_error(self, case, t: float) -> np.ndarray:
        """L2 error of (h, hu, hv) against the exact solution at time t."""
        out = np.empty(3)
        self._call("swe_case_l2_error", C.byref(case.c), float(t), capi.dptr(out))
        return out

    def time(self) -> float:
        v = C.c_double()
        self._call("swe_get_time", C.byref(v))
        return v.value

    # -- per-cell accessors of the reference API --
    def classify(self) -> np.ndarray:
        """IsDryCell / IsPartWetCell / IsFullWetCell of the current state: 0 / 1 / 2 per cell."""
        return self._get("swe_classify", (self.nt,), dtype=np.int8)

    def IsDryCell(self, i: int) -> bool:
        return int(self.classify()[i]) == 0

    def IsPartWetCell(self, i: int) -> bool:
        return int(self.classify()[i]) == 1

    def IsFullWetCell(self, i: int) -> bool:
        return int(self.classify()[i]) == 2

    def rhs(self, dt: float) -> np.ndarray:
        """TimeDisc::RHS(i, dt) of every cell, (nt, 3), for the last interface values / fluxes."""
        out = np.empty((self.nt, 3))
        self._call("swe_compute_rhs", float(dt), capi.dptr(out))
        return out

    # -- checkpoint / restart (binary, in the C-ABI; the reference only has text dumps, examples/Main.cpp:65-73) --
    def save_checkpoint(self, path: str):
        self._call("swe_checkpoint_save", str(path).encode())

    def load_checkpoint(self, path: str) -> float:
        """Restores state, time, dt and min_len_to_wavespeed; returns the simulated time stored with it."""
        self._call("swe_checkpoint_load", str(path).encode())
        return self.time()

    def kernel_timing(self, enable: bool = True):
        self._call("swe_kernel_timing", int(enable))

    def kernel_times(self) -> dict:
        """{kernel name: (total ms, launches)} recorded since kernel_timing(True)."""
        ms = np.zeros(16)
        cnt = np.zeros(16, dtype=np.int64)
        names = (C.c_char_p * 16)()
        k = capi.lib().swe_kernel_times(self._ctx, 16, capi.dptr(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64)), names)
        if k < 0:
            capi.check(k, self._ctx)
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(k)}

    def launch_count(self) -> int:
        return int(capi.lib().swe_launch_count(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None):
            if not getattr(self, "_borrowed", False):
                capi.lib().swe_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TimeDisc:
    def __init__(self, sd: SpaceDisc | None = None):
        self._sd = sd

    def GetSpaceDisc(self) -> SpaceDisc:
        return self._sd

    def SetSpaceDisc(self, sd: SpaceDisc):
        self._sd = sd

    def CFLdt(self) -> float:
        v = C.c_double()
        self._sd._call("swe_cfl_dt", C.byref(v))
        return v.value

    def RHS(self, i: int, dt: float) -> np.ndarray:
        """TimeDisc::RHS(i, dt) (include/TimeDisc.h:15); whole-array form: sd.rhs(dt)."""
        return self._sd.rhs(dt)[i]

    def ComputeDrainingDt(self, i: int) -> float:
        return float(self._sd.draining_dt()[i]) if i >= 0 else float("inf")


class Solvers:
    """Solvers::Euler / SSPRK2 / SSPRK3 (src/Solvers.cpp): one step of size dt on the device."""

    @staticmethod
    def _step(td: TimeDisc, scheme: int, dt: float):
        sd = td.GetSpaceDisc()
        sd._call("swe_step", scheme, sd.flux, sd.wavespeed, float(dt))

    @staticmethod
    def Euler(td: TimeDisc, dt: float):
        Solvers._step(td, capi.EULER, dt)

    @staticmethod
    def SSPRK2(td: TimeDisc, dt: float):
        Solvers._step(td, capi.SSPRK2, dt)

    @staticmethod
    def SSPRK3(td: TimeDisc, dt: float):
        Solvers._step(td, capi.SSPRK3, dt)

    @staticmethod
    def run(td: TimeDisc, scheme: str | int, nsteps: int, dt: float = 0.0, dt0: float = 0.0):
        """nsteps steps without host synchronisation; dt <= 0: dt = CFLdt() of the previous step."""
        sd = td.GetSpaceDisc()
        sc = capi.SCHEMES[scheme.lower()] if isinstance(scheme, str) else int(scheme)
        sd._call("swe_run", sc, sd.flux, sd.wavespeed, int(nsteps), float(dt), float(dt0))
