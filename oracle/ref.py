"""ctypes wrapper of oracle/_ref/libswe_ref_*.so: the UPSTREAM solver sources compiled here
(oracle/Makefile.ref, against the Eigen subset shim) behind oracle/ref_driver.cpp.
TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's reference arm, never by the product.

Two builds: "aswritten" (upstream + the two repairs HEAD cannot run without, S1 and S11) and
"repaired" (+ S2, S3). /root/reference exists only in the build container; on the GPU box the
prebuilt .so files that travelled with the working tree are used as they are."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SWE_REF_ROOT", "/root/reference")
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_libs: dict = {}
FN3 = C.CFUNCTYPE(None, _D, _D, C.c_void_p)


def lib_path(variant: str) -> str:
    return os.path.join(_HERE, "_ref", f"libswe_ref_{variant}.so")


def sources_present() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "MUSCLObject.cpp"))


def available(variant: str = "aswritten") -> bool:
    return os.path.exists(lib_path(variant)) or sources_present()


def build(force: bool = False) -> None:
    """Compile both variants when the upstream tree is present (no-op otherwise)."""
    if not sources_present():
        return
    deps = [os.path.join(_HERE, "ref_driver.cpp"), os.path.join(_HERE, "eigen_shim", "Eigen", "Dense"), os.path.join(_HERE, "Makefile.ref")]
    deps += [os.path.join(_HERE, "ref_patches", f) for f in os.listdir(os.path.join(_HERE, "ref_patches"))]
    newest = max(os.path.getmtime(d) for d in deps)
    stale = any(not os.path.exists(lib_path(v)) or os.path.getmtime(lib_path(v)) < newest for v in ("aswritten", "repaired"))
    if force or stale:
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call(["make", "-f", os.path.join(_HERE, "Makefile.ref"), "-j2", f"REF={REF_ROOT}", f"CXX={cxx}"] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)


def lib(variant: str = "aswritten") -> C.CDLL:
    if variant not in _libs:
        if not os.path.exists(lib_path(variant)):
            build()
        l = C.CDLL(lib_path(variant))
        l.ref_create.restype = C.c_void_p
        l.ref_create.argtypes = [C.c_int64, C.c_int64, C.c_int64, _D, _I64, _I64, _I64, _I64, _I64, C.c_double, C.c_double]
        l.ref_destroy.argtypes = [C.c_void_p]
        l.ref_last_error.restype = C.c_char_p
        l.ref_last_error.argtypes = [C.c_void_p]
        l.ref_set_state.argtypes = [C.c_void_p, _D]
        l.ref_get_state.argtypes = [C.c_void_p, _D]
        l.ref_assign_prim.argtypes = [C.c_void_p, C.c_int64, _D]
        l.ref_assign_cons.argtypes = [C.c_void_p, C.c_int64, _D]
        l.ref_get_cons.argtypes = [C.c_void_p, C.c_int64, _D]
        l.ref_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        l.ref_compute_interface_values.argtypes = [C.c_void_p]
        l.ref_compute_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.ref_stage_update_snapshot.argtypes = [C.c_void_p, _D, C.c_double, C.c_double, C.c_double, C.c_int]
        l.ref_min_len_to_wavespeed.restype = C.c_double
        l.ref_min_len_to_wavespeed.argtypes = [C.c_void_p]
        l.ref_cfl_dt.restype = C.c_double
        l.ref_cfl_dt.argtypes = [C.c_void_p]
        for name in ("edge_states", "sources", "fluxes", "node_max_w", "draining_dt"):
            getattr(l, "ref_get_" + name).argtypes = [C.c_void_p, _D]
        l.ref_set_node_max_w.argtypes = [C.c_void_p, _D]
        l.ref_rhs.argtypes = [C.c_void_p, C.c_int64, C.c_double, _D]
        l.ref_get_cell_class.argtypes = [C.c_void_p, C.POINTER(C.c_int8)]
        l.ref_is_part_wet.argtypes = [C.c_void_p, C.c_int64]
        l.ref_get_geometry.argtypes = [C.c_void_p, _D, _D, _D, _D, _D, _D]
        l.ref_norm.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _D]
        l.ref_tang.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _D]
        l.ref_bisection_cubic.restype = C.c_double
        l.ref_bisection_cubic.argtypes = [C.c_double] * 5
        l.ref_gradient.argtypes = [_D, _D]
        l.ref_elem_flux.argtypes = [_D, _D, _D]
        l.ref_wavespeeds.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _D]
        l.ref_len.restype = C.c_double
        l.ref_len.argtypes = [_D, _D]
        l.ref_det.restype = C.c_double
        l.ref_det.argtypes = [_D, _D]
        l.ref_triang_area.restype = C.c_double
        l.ref_triang_area.argtypes = [_D, _D, _D]
        l.ref_reconstruct.argtypes = [C.c_void_p, C.c_int, C.c_int64, _D, _D]
        l.ref_muscl_at_point.argtypes = [C.c_void_p, C.c_int64, _D, _D, _D, _D, _D]
        l.ref_edge_flux.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, _D, _D]
        l.ref_triang_average3.argtypes = [C.c_int, _D, _D, _D, FN3, C.c_void_p, _D]
        l.ref_patches.restype = C.c_char_p
        _libs[variant] = l
    return _libs[variant]


def _d(a):
    return a.ctypes.data_as(_D)


def _i(a):
    return a.ctypes.data_as(_I64)


class Ref:
    """Upstream Topology + Domain + SpaceDisc + TimeDisc on a mesh given as numpy arrays
    (same constructor and method names as oracle.oracle.Oracle)."""

    def __init__(self, mesh, cor: float = 0.0, tau: float = 0.0, variant: str = "aswritten"):
        self.l = lib(variant)
        self.variant = variant
        self.nn, self.ne, self.nt = mesh.nn, mesh.ne, mesh.nt
        g = np.ascontiguousarray(mesh.geometry, dtype=np.float64)
        arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in
                (mesh.edge_nodes, mesh.edge_elements, mesh.element_nodes, mesh.element_edges, mesh.element_neighbours)]
        self._h = self.l.ref_create(self.nn, self.ne, self.nt, _d(g), *[_i(a) for a in arrs], cor, tau)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.l.ref_last_error(self._h).decode())

    def patches(self) -> str:
        return self.l.ref_patches().decode()

    def set_state(self, prim):
        p = np.ascontiguousarray(prim, dtype=np.float64)
        assert p.shape == (self.nt, 3)
        self.l.ref_set_state(self._h, _d(p))

    def get_state(self):
        out = np.empty((self.nt, 3))
        self.l.ref_get_state(self._h, _d(out))
        return out

    def assign_prim(self, i, prim3):
        v = np.ascontiguousarray(prim3, dtype=np.float64)
        self.l.ref_assign_prim(self._h, i, _d(v))

    def assign_cons(self, i, cons3):
        v = np.ascontiguousarray(cons3, dtype=np.float64)
        self.l.ref_assign_cons(self._h, i, _d(v))

    def get_cons(self, i):
        out = np.empty(3)
        self.l.ref_get_cons(self._h, i, _d(out))
        return out

    def step(self, scheme=1, flux=1, ws=2, dt=1e-3):
        self._chk(self.l.ref_step(self._h, scheme, flux, ws, dt))

    def compute_interface_values(self):
        self._chk(self.l.ref_compute_interface_values(self._h))

    def compute_fluxes(self, flux=1, ws=2):
        self._chk(self.l.ref_compute_fluxes(self._h, flux, ws))

    def stage_update_snapshot(self, U0, a0, a1, dts, plain_sum):
        u0 = None if U0 is None else np.ascontiguousarray(U0, dtype=np.float64)
        self._chk(self.l.ref_stage_update_snapshot(self._h, None if u0 is None else _d(u0), a0, a1, dts, int(plain_sum)))

    def min_len_to_wavespeed(self):
        return self.l.ref_min_len_to_wavespeed(self._h)

    def cfl_dt(self):
        return self.l.ref_cfl_dt(self._h)

    def _get(self, name, shape):
        out = np.empty(shape)
        getattr(self.l, "ref_get_" + name)(self._h, _d(out))
        return out

    def edge_states(self):
        return self._get("edge_states", (2 * self.ne, 3))

    def sources(self):
        return self._get("sources", (2 * self.ne, 3))

    def fluxes(self):
        return self._get("fluxes", (self.ne, 3))

    def node_max_w(self):
        return self._get("node_max_w", (self.nn,))

    def set_node_max_w(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.l.ref_set_node_max_w(self._h, _d(v))

    def draining_dt(self):
        return self._get("draining_dt", (self.nt,))

    def rhs(self, i, dt):
        out = np.empty(3)
        self.l.ref_rhs(self._h, i, dt, _d(out))
        return out

    def cell_class(self):
        out = np.empty(self.nt, dtype=np.int8)
        self.l.ref_get_cell_class(self._h, out.ctypes.data_as(C.POINTER(C.c_int8)))
        return out

    def geometry(self):
        T, E = np.empty((self.nt, 3)), np.empty((self.ne, 3))
        L, A = np.empty(self.ne), np.empty(self.nt)
        n0, sl = np.empty((self.ne, 2)), np.empty((self.nt, 2))
        self.l.ref_get_geometry(self._h, _d(T), _d(E), _d(L), _d(A), _d(n0), _d(sl))
        return dict(T=T, E=E, L=L, A=A, n0=n0, slope=sl)

    def norm(self, e, t):
        out = np.empty(2)
        self.l.ref_norm(self._h, e, t, _d(out))
        return out

    def reconstruct(self, kind: int, i: int):
        o, G = np.empty(3), np.empty((3, 2))
        self._chk(self.l.ref_reconstruct(self._h, kind, i, _d(o), _d(G)))
        return o, G

    def muscl_at_point(self, i, o, G, pt):
        o = np.ascontiguousarray(o, dtype=np.float64)
        G = np.ascontiguousarray(G, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        out, g = np.empty(3), np.empty(2)
        self.l.ref_muscl_at_point(self._h, i, _d(o), _d(G), _d(pt), _d(out), _d(g))
        return out, g

    def edge_flux(self, flux, ws, e, with_r=True):
        F = np.empty(3)
        r = C.c_double(1.0)
        self._chk(self.l.ref_edge_flux(self._h, flux, ws, e, _d(F), C.byref(r) if with_r else None))
        return F, r.value

    def __del__(self):
        try:
            if self._h:
                self.l.ref_destroy(self._h)
                self._h = None
        except Exception:
            pass


# unit-level helpers (no context)
def gradient(P9, variant="aswritten"):
    P9 = np.ascontiguousarray(P9, dtype=np.float64)
    g = np.empty(2)
    lib(variant).ref_gradient(_d(P9), _d(g))
    return g


def bisection_cubic(d, c, b, lo, hi, variant="aswritten"):
    return lib(variant).ref_bisection_cubic(d, c, b, lo, hi)


def elem_flux(n2, U3, variant="aswritten"):
    n2 = np.ascontiguousarray(n2, dtype=np.float64)
    U3 = np.ascontiguousarray(U3, dtype=np.float64)
    F = np.empty(3)
    lib(variant).ref_elem_flux(_d(n2), _d(U3), _d(F))
    return F


def wavespeeds(ws, ul, hl, ur, hr, variant="aswritten"):
    a = np.empty(2)
    lib(variant).ref_wavespeeds(ws, ul, hl, ur, hr, _d(a))
    return a


def triang_average3(n, p0, p1, p2, fn, variant="aswritten"):
    """TriangAverage<3,n> of the Python callable fn(point3) -> 3 values."""
    def cb(pt, out, _user):
        v = fn(np.array([pt[0], pt[1], pt[2]]))
        out[0], out[1], out[2] = float(v[0]), float(v[1]), float(v[2])
    p0, p1, p2 = (np.ascontiguousarray(p, dtype=np.float64) for p in (p0, p1, p2))
    out = np.empty(3)
    rc = lib(variant).ref_triang_average3(n, _d(p0), _d(p1), _d(p2), FN3(cb), None, _d(out))
    if rc != 0:
        raise ValueError(f"TriangAverage<3,{n}> is not instantiated in ref_driver.cpp")
    return out
