"""Host-side meshes and analytic cases — Python mirror of the reference's TriangMesh /
StructTriangMesh (include/TriangMesh.h, include/StructTriangMesh.h) and Test hierarchy
(examples/Tests.h), implemented in C++ behind the C-ABI (csrc/hostmesh.cpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class TriangMesh:
    """Owning triangular mesh: node geometry (x, y, b) and the five Topology incidence arrays
    (include/TriangMesh.h:27-33) as numpy views into C++-owned memory."""

    def __init__(self, handle, owner=True):
        self._h = handle
        self._owner = owner  # False: a view of a mesh owned by someone else (e.g. a decomposition plan)
        self._view = capi.MeshStruct()
        capi.check(capi.lib().swe_hostmesh_view(self._h, C.byref(self._view)))
        v = self._view
        self.nn, self.ne, self.nt = int(v.nn), int(v.ne), int(v.nt)

        def arr(ptr, n, dtype):
            ctype = C.c_double if dtype == np.float64 else C.c_int64
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))

        # geometry: 3 x nn column-major -> expose as (nn, 3) rows (x, y, b), writable
        self.geometry = arr(capi.lib().swe_hostmesh_geometry(self._h), 3 * self.nn, np.float64).reshape(self.nn, 3)
        self.edge_nodes = arr(v.edge_nodes, 2 * self.ne, np.int64).reshape(self.ne, 2)
        self.edge_elements = arr(v.edge_elements, 2 * self.ne, np.int64).reshape(self.ne, 2)
        self.element_nodes = arr(v.element_nodes, 3 * self.nt, np.int64).reshape(self.nt, 3)
        self.element_edges = arr(v.element_edges, 3 * self.nt, np.int64).reshape(self.nt, 3)
        self.element_neighbours = arr(v.element_neighbours, 3 * self.nt, np.int64).reshape(self.nt, 3)

    # -- constructors ------------------------------------------------------------------
    @classmethod
    def from_gmsh(cls, path: str) -> "TriangMesh":
        h = C.c_void_p()
        capi.check(capi.lib().swe_hostmesh_gmsh(C.byref(h), path.encode()))
        return cls(h)

    @classmethod
    def from_triangles(cls, xy, tri, boundary=None) -> "TriangMesh":
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        tri = np.ascontiguousarray(tri, dtype=np.int64)
        nb = 0 if boundary is None else len(boundary)
        b = None if boundary is None else np.ascontiguousarray(boundary, dtype=np.int64)
        h = C.c_void_p()
        capi.check(capi.lib().swe_hostmesh_from_triangles(
            C.byref(h), len(xy), capi.dptr(xy), len(tri), tri.ctypes.data_as(C.POINTER(C.c_int64)),
            nb, None if b is None else b.ctypes.data_as(C.POINTER(C.c_int64))))
        return cls(h)

    def refine(self) -> "TriangMesh":
        """Uniform 1 -> 4 refinement (new nodes at edge midpoints)."""
        h = C.c_void_p()
        capi.check(capi.lib().swe_hostmesh_refine(C.byref(h), self._h))
        return TriangMesh(h)

    def extract(self, part: np.ndarray, rank: int, layers: int) -> "TriangMesh":
        """Sub-mesh of the cells with part == rank plus `layers` vertex-adjacent halo rings."""
        part = np.ascontiguousarray(part, dtype=np.int32)
        h = C.c_void_p()
        capi.check(capi.lib().swe_hostmesh_extract(
            C.byref(h), self._h, part.ctypes.data_as(C.POINTER(C.c_int32)), rank, layers))
        sub = TriangMesh(h)
        gc = capi.lib().swe_hostmesh_global_cells(sub._h)
        ow = capi.lib().swe_hostmesh_cell_owner(sub._h)
        sub.global_cells = np.ctypeslib.as_array(gc, shape=(sub.nt,))
        sub.cell_owner = np.ctypeslib.as_array(ow, shape=(sub.nt,))
        return sub

    def partition_rcb(self, nparts: int) -> np.ndarray:
        part = np.empty(self.nt, dtype=np.int32)
        capi.check(capi.lib().swe_partition_rcb(self._h, nparts, part.ctypes.data_as(C.POINTER(C.c_int32))))
        return part

    # -- helpers -----------------------------------------------------------------------
    def c_mesh(self, cor: float = 0.0, tau: float = 0.0) -> capi.MeshStruct:
        m = capi.MeshStruct()
        C.memmove(C.byref(m), C.byref(self._view), C.sizeof(m))
        m.cor, m.tau = cor, tau
        return m

    def centroids(self) -> np.ndarray:
        """Domain::T (src/Bathymetry.cpp:24-27): (nt, 3) centroid x, y and cell bed b_i."""
        p = self.geometry[self.element_nodes]  # (nt, 3 nodes, 3 coords)
        third = 1.0 / 3.0
        return p[:, 0] * third + p[:, 1] * third + p[:, 2] * third

    def areas(self) -> np.ndarray:
        p = self.geometry[self.element_nodes]
        a, b = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        return 0.5 * np.abs(a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1])

    def __del__(self):
        try:
            if self._h and self._owner:
                capi.lib().swe_hostmesh_free(self._h)
            self._h = None
        except Exception:
            pass


class StructTriangMesh(TriangMesh):
    """StructTriangMesh(ni, nj, h) (include/StructTriangMesh.h:4-15): [0, ni*h] x [0, nj*h],
    4 triangles (Bottom/Right/Top/Left) per square. i0/j0 place the block inside a larger grid."""

    def __init__(self, ni: int, nj: int, h: float, i0: int = 0, j0: int = 0):
        hd = C.c_void_p()
        capi.check(capi.lib().swe_hostmesh_struct(C.byref(hd), ni, nj, float(h), i0, j0))
        super().__init__(hd)
        self.ni, self.nj, self.h = ni, nj, float(h)

    def Ni(self):
        return self.ni

    def Nj(self):
        return self.nj


class Case:
    """Analytic test case (examples/Tests.h): bathymetry b(x, y) and exact (h, u, v)(x, y, t)."""

    KINDS = {
        "lake_at_rest": capi.CASE_LAKE_AT_REST,
        "classic_thacker": capi.CASE_CLASSIC_THACKER,
        "gauss_wave": capi.CASE_GAUSS_WAVE,
        "fully_wet": capi.CASE_FULLY_WET,
        "bowl_hump": capi.CASE_BOWL_HUMP,
    }

    def __init__(self, kind: str, mid_x: float, mid_y: float, length: float, **params):
        self.c = capi.CaseStruct()
        capi.lib().swe_case_defaults(C.byref(self.c), self.KINDS[kind], mid_x, mid_y, length)
        for k, v in params.items():
            if not hasattr(self.c, k):
                raise KeyError(k)
            setattr(self.c, k, v)
        self.kind = kind

    def eval(self, x: float, y: float, t: float = 0.0):
        """-> (b, h, u, v) of the exact solution."""
        out = np.empty(4)
        capi.check(capi.lib().swe_case_eval(C.byref(self.c), x, y, t, capi.dptr(out)))
        return out

    def set_bathymetry(self, mesh: TriangMesh) -> None:
        capi.check(capi.lib().swe_case_set_bathymetry(C.byref(self.c), mesh._h))

    def initial_state(self, mesh: TriangMesh, quad_n: int = 4, t: float = 0.0) -> np.ndarray:
        """(nt, 3) primitive cell state (w, u, v), cf. examples/Main.cpp:211-223."""
        prim = np.empty((mesh.nt, 3))
        capi.check(capi.lib().swe_case_initial_state(C.byref(self.c), mesh._h, quad_n, t, capi.dptr(prim)))
        return prim
