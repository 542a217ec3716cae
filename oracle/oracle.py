"""ctypes wrapper of the CPU oracle (oracle/libswe_oracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswe_oracle.so")
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "swe_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libswe_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.oracle_create.restype = C.c_void_p
        l.oracle_create.argtypes = [C.c_int64, C.c_int64, C.c_int64, _D, _I64, _I64, _I64, _I64, _I64, C.c_double, C.c_double]
        l.oracle_destroy.argtypes = [C.c_void_p]
        l.oracle_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.oracle_get_threads.argtypes = [C.c_void_p]
        l.oracle_set_cfl_edge_mask.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        l.oracle_set_state.argtypes = [C.c_void_p, _D]
        l.oracle_get_state.argtypes = [C.c_void_p, _D]
        l.oracle_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        l.oracle_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double]
        l.oracle_compute_interface_values.argtypes = [C.c_void_p]
        l.oracle_compute_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.oracle_stage_update.argtypes = [C.c_void_p, _D, C.c_double, C.c_double, C.c_double, C.c_int]
        l.oracle_min_len_to_wavespeed.restype = C.c_double
        l.oracle_min_len_to_wavespeed.argtypes = [C.c_void_p]
        l.oracle_cfl_dt.restype = C.c_double
        l.oracle_cfl_dt.argtypes = [C.c_void_p]
        for name in ("edge_states", "sources", "fluxes", "node_max_w", "draining_dt"):
            getattr(l, "oracle_get_" + name).argtypes = [C.c_void_p, _D]
        l.oracle_get_branch_counts.argtypes = [C.c_void_p, _I64]
        l.oracle_get_cell_class.argtypes = [C.c_void_p, C.POINTER(C.c_int8)]
        l.oracle_get_geometry.argtypes = [C.c_void_p, _D, _D, _D, _D, _D, _D]
        l.oracle_diagnostics.argtypes = [C.c_void_p, _D]
        l.oracle_cbrt.restype = C.c_double
        l.oracle_cbrt.argtypes = [C.c_double]
        l.oracle_ilog2_trunc.restype = C.c_int
        l.oracle_ilog2_trunc.argtypes = [C.c_double]
        l.oracle_bisection_cubic.restype = C.c_double
        l.oracle_bisection_cubic.argtypes = [C.c_double] * 5
        l.oracle_bisection_cubic2.restype = C.c_double
        l.oracle_bisection_cubic2.argtypes = [C.c_double] * 5 + [C.c_int]
        l.oracle_rhs.argtypes = [C.c_void_p, C.c_int64, C.c_double, _D]
        l.oracle_edge_flux.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, _D, _D]
        l.oracle_wavespeeds.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _D]
        l.oracle_get_draining_dt_live.argtypes = [C.c_void_p, _D]
        l.oracle_assign_cons.argtypes = [C.c_void_p, C.c_int64, _D]
        l.oracle_get_cons.argtypes = [C.c_void_p, C.c_int64, _D]
        l.oracle_gradient.argtypes = [_D, _D]
        l.oracle_elem_flux.argtypes = [_D, _D, _D]
        l.oracle_reconstruct.argtypes = [C.c_void_p, C.c_int, C.c_int64, _D, _D]
        l.oracle_muscl_at_point.argtypes = [C.c_void_p, C.c_int64, _D, _D, _D, _D]
        l.oracle_set_node_max_w.argtypes = [C.c_void_p, _D]
        l.oracle_struct_mesh_sizes.argtypes = [C.c_int64, C.c_int64, _I64, _I64, _I64]
        l.oracle_struct_mesh.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_int64, C.c_int64, _D, _I64, _I64, _I64, _I64, _I64]
        l.oracle_case_eval.argtypes = [C.c_int, _D, C.c_double, C.c_double, C.c_double, _D]
        l.oracle_triang_average_poly.argtypes = [C.c_int, _D, _D, _D, _D, _D]
        l.oracle_case_set_bathymetry.argtypes = [C.c_int, _D, C.c_int64, _D]
        l.oracle_case_initial_state.argtypes = [C.c_int, _D, C.c_int64, _D, _I64, C.c_int, C.c_double, _D]
        _lib = l
    return _lib


def _d(a):
    return a.ctypes.data_as(_D)


def _i(a):
    return a.ctypes.data_as(_I64)


class Oracle:
    """CPU restatement of SpaceDisc + TimeDisc + Solvers on a mesh given exactly like the
    reference's Topology/Domain (numpy arrays: geometry (nn,3), edge_nodes (ne,2), ...)."""

    def __init__(self, mesh, cor: float = 0.0, tau: float = 0.0, **options):
        l = lib()
        self.nn, self.ne, self.nt = mesh.nn, mesh.ne, mesh.nt
        g = np.ascontiguousarray(mesh.geometry, dtype=np.float64)
        arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in
                (mesh.edge_nodes, mesh.edge_elements, mesh.element_nodes, mesh.element_edges, mesh.element_neighbours)]
        self._h = l.oracle_create(self.nn, self.ne, self.nt, _d(g), *[_i(a) for a in arrs], cor, tau)
        for k, v in options.items():
            self.set_option(k, v)

    def set_option(self, key: str, value: int):
        if lib().oracle_set_option(self._h, key.encode(), int(value)) != 0:
            raise KeyError(key)

    @property
    def threads(self) -> int:
        return lib().oracle_get_threads(self._h)

    def set_cfl_edge_mask(self, mask):
        if mask is None:
            lib().oracle_set_cfl_edge_mask(self._h, None)
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            lib().oracle_set_cfl_edge_mask(self._h, m.ctypes.data_as(C.POINTER(C.c_uint8)))

    def set_state(self, prim):
        p = np.ascontiguousarray(prim, dtype=np.float64)
        assert p.shape == (self.nt, 3)
        lib().oracle_set_state(self._h, _d(p))

    def get_state(self):
        out = np.empty((self.nt, 3))
        lib().oracle_get_state(self._h, _d(out))
        return out

    def step(self, scheme=1, flux=1, ws=2, dt=1e-3):
        lib().oracle_step(self._h, scheme, flux, ws, dt)

    def run(self, scheme, flux, ws, nsteps, dt, dt0=0.0):
        lib().oracle_run(self._h, scheme, flux, ws, nsteps, dt, dt0)

    def compute_interface_values(self):
        lib().oracle_compute_interface_values(self._h)

    def compute_fluxes(self, flux=1, ws=2):
        lib().oracle_compute_fluxes(self._h, flux, ws)

    def stage_update(self, U0, a0, a1, dts, plain_sum):
        u0 = None if U0 is None else np.ascontiguousarray(U0, dtype=np.float64)
        lib().oracle_stage_update(self._h, None if u0 is None else _d(u0), a0, a1, dts, int(plain_sum))

    def min_len_to_wavespeed(self):
        return lib().oracle_min_len_to_wavespeed(self._h)

    def cfl_dt(self):
        return lib().oracle_cfl_dt(self._h)

    def _get(self, name, shape):
        out = np.empty(shape)
        getattr(lib(), "oracle_get_" + name)(self._h, _d(out))
        return out

    def edge_states(self):
        return self._get("edge_states", (2 * self.ne, 3))

    def sources(self):
        return self._get("sources", (2 * self.ne, 3))

    def fluxes(self):
        return self._get("fluxes", (self.ne, 3))

    def node_max_w(self):
        return self._get("node_max_w", (self.nn,))

    def draining_dt(self):
        return self._get("draining_dt", (self.nt,))

    def draining_dt_live(self):
        return self._get("draining_dt_live", (self.nt,))

    def rhs(self, i, dt):
        out = np.empty(3)
        lib().oracle_rhs(self._h, i, dt, _d(out))
        return out

    def edge_flux(self, flux, ws, e, with_r=True):
        F = np.empty(3)
        r = C.c_double(1.0)
        lib().oracle_edge_flux(self._h, flux, ws, e, _d(F), C.byref(r) if with_r else None)
        return F, r.value

    def wavespeeds(self, ws, ul, hl, ur, hr):
        a = np.empty(2)
        lib().oracle_wavespeeds(self._h, ws, ul, hl, ur, hr, _d(a))
        return a

    def assign_cons(self, i, cons3):
        v = np.ascontiguousarray(cons3, dtype=np.float64)
        lib().oracle_assign_cons(self._h, i, _d(v))

    def get_cons(self, i):
        out = np.empty(3)
        lib().oracle_get_cons(self._h, i, _d(out))
        return out

    BRANCHES = ("pw1_submerged", "pw1_cbrt", "pw1_bisection", "fw_dry_neighbour", "fw_partwet_neighbour", "fw_vertex_zeroed",
                "fw_tvd_off", "pw2_to_pw1", "pw2_one_wet", "pw2_three_wet", "pw2_two_wet", "pw2_two_wet_fallback")

    def branch_counts(self):
        """Branch-hit counters of the last compute_interface_values (see swe_oracle.cpp `branch`)."""
        out = np.zeros(12, dtype=np.int64)
        lib().oracle_get_branch_counts(self._h, _i(out))
        return dict(zip(self.BRANCHES, out.tolist()))

    def cell_class(self):
        out = np.empty(self.nt, dtype=np.int8)
        lib().oracle_get_cell_class(self._h, out.ctypes.data_as(C.POINTER(C.c_int8)))
        return out

    def geometry(self):
        T, E = np.empty((self.nt, 3)), np.empty((self.ne, 3))
        L, A = np.empty(self.ne), np.empty(self.nt)
        n0, sl = np.empty((self.ne, 2)), np.empty((self.nt, 2))
        lib().oracle_get_geometry(self._h, _d(T), _d(E), _d(L), _d(A), _d(n0), _d(sl))
        return dict(T=T, E=E, L=L, A=A, n0=n0, slope=sl)

    def diagnostics(self):
        out = np.empty(6)
        lib().oracle_diagnostics(self._h, _d(out))
        return out

    def reconstruct(self, kind: int, i: int):
        o, G = np.empty(3), np.empty((3, 2))
        lib().oracle_reconstruct(self._h, kind, i, _d(o), _d(G))
        return o, G

    def muscl_at_point(self, i, o, G, pt):
        o = np.ascontiguousarray(o, dtype=np.float64)
        G = np.ascontiguousarray(G, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        out = np.empty(3)
        lib().oracle_muscl_at_point(self._h, i, _d(o), _d(G), _d(pt), _d(out))
        return out

    def set_node_max_w(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        lib().oracle_set_node_max_w(self._h, _d(v))

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass


class OracleStructMesh:
    """StructTriangMesh(ni, nj, h) generated by the oracle itself (closed form, conventions of SURVEY App. B):
    same attributes as the product's mesh classes, no product code involved."""

    def __init__(self, ni: int, nj: int, h: float, i0: int = 0, j0: int = 0):
        l = lib()
        nn, ne, nt = C.c_int64(), C.c_int64(), C.c_int64()
        l.oracle_struct_mesh_sizes(ni, nj, C.byref(nn), C.byref(ne), C.byref(nt))
        self.nn, self.ne, self.nt = nn.value, ne.value, nt.value
        self.ni, self.nj, self.h = ni, nj, float(h)
        self.geometry = np.empty((self.nn, 3))
        self.edge_nodes = np.empty((self.ne, 2), dtype=np.int64)
        self.edge_elements = np.empty((self.ne, 2), dtype=np.int64)
        self.element_nodes = np.empty((self.nt, 3), dtype=np.int64)
        self.element_edges = np.empty((self.nt, 3), dtype=np.int64)
        self.element_neighbours = np.empty((self.nt, 3), dtype=np.int64)
        l.oracle_struct_mesh(ni, nj, float(h), i0, j0, _d(self.geometry), _i(self.edge_nodes), _i(self.edge_elements),
                             _i(self.element_nodes), _i(self.element_edges), _i(self.element_neighbours))


class OracleCase:
    """Analytic cases of examples/Tests.h restated for the oracle: bathymetry, exact solution, cell-average IC."""

    KINDS = {"lake_at_rest": 0, "classic_thacker": 1, "gauss_wave": 2, "fully_wet": 3, "bowl_hump": 4}

    def __init__(self, kind: str, mid_x: float, mid_y: float, length: float, cor=0.0, tau=0.0, delta=1.0, H0=0.5, p0=0.0,
                 q0=0.0, level=0.0, amp=None):
        self.kind = self.KINDS[kind]
        if amp is None:
            amp = 0.05 if kind == "fully_wet" else 0.0
        self.par = np.array([mid_x, mid_y, length, cor, tau, delta, H0, p0, q0, level, amp], dtype=np.float64)

    def eval(self, x, y, t=0.0):
        out = np.empty(4)
        lib().oracle_case_eval(self.kind, _d(self.par), x, y, t, _d(out))
        return out

    def set_bathymetry(self, mesh):
        g = mesh.geometry
        assert g.flags.c_contiguous
        lib().oracle_case_set_bathymetry(self.kind, _d(self.par), mesh.nn, _d(g))

    def initial_state(self, mesh, quad_n=4, t=0.0):
        prim = np.empty((mesh.nt, 3))
        g = np.ascontiguousarray(mesh.geometry, dtype=np.float64)
        tp = np.ascontiguousarray(mesh.element_nodes, dtype=np.int64)
        lib().oracle_case_initial_state(self.kind, _d(self.par), mesh.nt, _d(g), _i(tp), int(quad_n), float(t), _d(prim))
        return prim
