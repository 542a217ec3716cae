"""ctypes binding of the C-ABI declared in include/swe_b200.h (libswe_b200.so).

This is the only way Python reaches the solver: plain pointers and sizes, numpy arrays for
host buffers, integers for device addresses. There is no CPU fallback: if the shared library is
missing, or no CUDA device is present, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SWE_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("SWE_B200_LIB") or os.path.join(_HERE, "libswe_b200.so")

OK = 0
EULER, SSPRK2, SSPRK3 = 0, 1, 2
HLL, HLLC = 0, 1
RUSANOV, DAVIS, EINFELDT = 0, 1, 2
CASE_LAKE_AT_REST, CASE_CLASSIC_THACKER, CASE_GAUSS_WAVE, CASE_FULLY_WET, CASE_BOWL_HUMP = range(5)

SCHEMES = {"euler": EULER, "ssprk2": SSPRK2, "ssprk3": SSPRK3}
FLUXES = {"hll": HLL, "hllc": HLLC}
WAVESPEEDS = {"rusanov": RUSANOV, "davis": DAVIS, "einfeldt": EINFELDT}


class SweError(RuntimeError):
    """Raised for any non-zero status returned by the C-ABI."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[swe_b200 status {status}] {message}")
        self.status = status


class MeshStruct(C.Structure):  # struct swe_mesh
    _fields_ = [
        ("nn", C.c_int64), ("ne", C.c_int64), ("nt", C.c_int64),
        ("geometry", C.POINTER(C.c_double)),
        ("edge_nodes", C.POINTER(C.c_int64)),
        ("edge_elements", C.POINTER(C.c_int64)),
        ("element_nodes", C.POINTER(C.c_int64)),
        ("element_edges", C.POINTER(C.c_int64)),
        ("element_neighbours", C.POINTER(C.c_int64)),
        ("cor", C.c_double), ("tau", C.c_double),
    ]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)  # swe_allgather_fn


class DistConfig(C.Structure):  # struct swe_dist_config
    _fields_ = [
        ("device", C.c_int32), ("reorder", C.c_int32), ("overlap", C.c_int32), ("reserved", C.c_int32),
        ("cor", C.c_double), ("tau", C.c_double), ("wait_timeout_s", C.c_double),
        ("allgather", ALLGATHER_FN), ("user", C.c_void_p),
    ]


class CaseStruct(C.Structure):  # struct swe_case
    _fields_ = [
        ("kind", C.c_int32),
        ("mid_x", C.c_double), ("mid_y", C.c_double), ("length", C.c_double),
        ("cor", C.c_double), ("tau", C.c_double), ("delta", C.c_double),
        ("H0", C.c_double), ("p0", C.c_double), ("q0", C.c_double),
        ("level", C.c_double), ("amp", C.c_double),
    ]


# every symbol include/swe_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_I32 = C.POINTER(C.c_int32)
_U8 = C.POINTER(C.c_uint8)
_I8 = C.POINTER(C.c_int8)
SYMBOLS = {
    "swe_create": (C.c_int, [C.POINTER(_P), C.POINTER(MeshStruct), C.c_int, C.c_int]),
    "swe_create_classes": (C.c_int, [C.POINTER(_P), C.POINTER(MeshStruct), C.c_int, C.c_int, _U8]),
    "swe_destroy": (None, [_P]),
    "swe_last_error": (C.c_char_p, [_P]),
    "swe_set_stream": (C.c_int, [_P, _P]),
    "swe_synchronize": (C.c_int, [_P]),
    "swe_set_state": (C.c_int, [_P, _D]),
    "swe_get_state": (C.c_int, [_P, _D]),
    "swe_set_state_async": (C.c_int, [_P, _D]),
    "swe_get_state_async": (C.c_int, [_P, _D]),
    "swe_submit_step_host": (C.c_int, [_P, _D, _D, C.c_int, C.c_int, C.c_int, C.c_double]),
    "swe_wait_host": (C.c_int, [_P]),
    "swe_dist_submit_step_host": (C.c_int, [_P, _D, _D, C.c_int, C.c_int, C.c_int, C.c_double]),
    "swe_dist_wait_host": (C.c_int, [_P]),
    "swe_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double]),
    "swe_run": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double]),
    "swe_cfl_dt": (C.c_int, [_P, _D]),
    "swe_get_min_len_to_wavespeed": (C.c_int, [_P, _D]),
    "swe_get_time": (C.c_int, [_P, _D]),
    "swe_launch_count": (C.c_int64, [_P]),
    "swe_kernel_timing": (C.c_int, [_P, C.c_int]),
    "swe_kernel_times": (C.c_int, [_P, C.c_int32, _D, _I64, C.POINTER(C.c_char_p)]),
    "swe_compute_interface_values": (C.c_int, [_P]),
    "swe_compute_fluxes": (C.c_int, [_P, C.c_int, C.c_int]),
    "swe_compute_interface_values_range": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "swe_compute_interface_values_class": (C.c_int, [_P, C.c_int32, C.c_int, C.c_int]),
    "swe_save_state": (C.c_int, [_P]),
    "swe_stage_update": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "swe_stage_update_dev": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "swe_set_dt": (C.c_int, [_P, C.c_double]),
    "swe_advance_dt": (C.c_int, [_P, C.c_int, C.c_double]),
    "swe_enable_taps": (C.c_int, [_P, C.c_int]),
    "swe_fluxer_count": (C.c_int32, []),
    "swe_fluxer_name": (C.c_char_p, [C.c_int32]),
    "swe_fluxer_id": (C.c_int32, [C.c_int32]),
    "swe_fluxer_find": (C.c_int32, [C.c_char_p]),
    "swe_set_fluxer": (C.c_int, [_P, C.c_int32]),
    "swe_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "swe_get_option": (C.c_int, [_P, C.c_char_p, _I32]),
    "swe_get_branch_counts": (C.c_int, [_P, _I64]),
    "swe_get_edge_states": (C.c_int, [_P, _D]),
    "swe_get_sources": (C.c_int, [_P, _D]),
    "swe_get_fluxes": (C.c_int, [_P, _D]),
    "swe_get_node_max_w": (C.c_int, [_P, _D]),
    "swe_get_draining_dt": (C.c_int, [_P, _D]),
    "swe_get_cell_class": (C.c_int, [_P, _I8]),
    "swe_classify": (C.c_int, [_P, _I8]),
    "swe_compute_rhs": (C.c_int, [_P, C.c_double, _D]),
    "swe_set_time": (C.c_int, [_P, C.c_double]),
    "swe_get_dt": (C.c_int, [_P, _D]),
    "swe_checkpoint_save": (C.c_int, [_P, C.c_char_p]),
    "swe_checkpoint_load": (C.c_int, [_P, C.c_char_p]),
    "swe_diagnostics": (C.c_int, [_P, _D]),
    "swe_set_cfl_edge_mask": (C.c_int, [_P, _U8]),
    "swe_halo_set_lists": (C.c_int, [_P, C.c_int64, _I64, C.c_int64, _I64]),
    "swe_halo_pack": (C.c_int, [_P, _P]),
    "swe_halo_unpack": (C.c_int, [_P, _P]),
    "swe_halo_p2p_alloc": (C.c_int, [_P, C.c_int32, _U8]),
    "swe_halo_p2p_connect": (C.c_int, [_P, C.c_int64, C.c_int64, _U8, C.c_int64, C.c_int32]),
    "swe_halo_p2p_push": (C.c_int, [_P]),
    "swe_halo_p2p_pull": (C.c_int, [_P]),
    "swe_halo_p2p_error": (C.c_int, [_P]),
    "swe_set_min_len_to_wavespeed": (C.c_int, [_P, C.c_double]),
    "swe_min_len_device_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "swe_dist_plan_struct": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_double]),
    "swe_dist_plan_mesh": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, _P, _I32]),
    "swe_dist_plan_free": (None, [_P]),
    "swe_dist_plan_local_mesh": (_P, [_P]),
    "swe_dist_plan_owned_count": (C.c_int64, [_P]),
    "swe_dist_plan_owned": (_U8, [_P]),
    "swe_dist_plan_global_cells": (_I64, [_P]),
    "swe_dist_plan_classes": (_U8, [_P]),
    "swe_dist_plan_cfl_mask": (_U8, [_P]),
    "swe_dist_plan_npeers": (C.c_int32, [_P]),
    "swe_dist_plan_peer": (C.c_int, [_P, C.c_int32, _I32, _I64, C.POINTER(_I64), _I64, C.POINTER(_I64)]),
    "swe_dist_create": (C.c_int, [C.POINTER(_P), _P, C.POINTER(DistConfig)]),
    "swe_dist_group_create": (C.c_int, [C.POINTER(_P), C.POINTER(_P), _I32, C.c_int32, C.POINTER(DistConfig)]),
    "swe_dist_destroy": (None, [_P]),
    "swe_dist_last_error": (C.c_char_p, [_P]),
    "swe_dist_ctx": (_P, [_P]),
    "swe_dist_get_plan": (_P, [_P]),
    "swe_dist_exchange": (C.c_int, [_P]),
    "swe_dist_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double]),
    "swe_dist_run": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double]),
    "swe_dist_group_run": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double]),
    "swe_dist_synchronize": (C.c_int, [_P]),
    "swe_dist_cfl_dt": (C.c_int, [_P, _D]),
    "swe_dist_state_hash": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "swe_state_hash": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "swe_dist_get_owned_state": (C.c_int, [_P, _D]),
    "swe_dist_set_state_global": (C.c_int, [_P, _D]),
    "swe_hostmesh_struct": (C.c_int, [C.POINTER(_P), C.c_int64, C.c_int64, C.c_double, C.c_int64, C.c_int64]),
    "swe_hostmesh_gmsh": (C.c_int, [C.POINTER(_P), C.c_char_p]),
    "swe_hostmesh_from_triangles": (C.c_int, [C.POINTER(_P), C.c_int64, _D, C.c_int64, _I64, C.c_int64, _I64]),
    "swe_hostmesh_refine": (C.c_int, [C.POINTER(_P), _P]),
    "swe_hostmesh_free": (None, [_P]),
    "swe_hostmesh_view": (C.c_int, [_P, C.POINTER(MeshStruct)]),
    "swe_hostmesh_geometry": (_D, [_P]),
    "swe_hostmesh_extract": (C.c_int, [C.POINTER(_P), _P, _I32, C.c_int32, C.c_int32]),
    "swe_hostmesh_global_cells": (_I64, [_P]),
    "swe_hostmesh_cell_owner": (_I32, [_P]),
    "swe_partition_rcb": (C.c_int, [_P, C.c_int32, _I32]),
    "swe_case_defaults": (None, [C.POINTER(CaseStruct), C.c_int32, C.c_double, C.c_double, C.c_double]),
    "swe_case_eval": (C.c_int, [C.POINTER(CaseStruct), C.c_double, C.c_double, C.c_double, _D]),
    "swe_case_set_bathymetry": (C.c_int, [C.POINTER(CaseStruct), _P]),
    "swe_case_initial_state": (C.c_int, [C.POINTER(CaseStruct), _P, C.c_int32, C.c_double, _D]),
    "swe_case_set_bathymetry_device": (C.c_int, [_P, C.POINTER(CaseStruct)]),
    "swe_case_initial_state_device": (C.c_int, [_P, C.POINTER(CaseStruct), C.c_int32, C.c_double]),
    "swe_case_l2_error": (C.c_int, [_P, C.POINTER(CaseStruct), C.c_double, _D]),
    "swe_version": (C.c_char_p, []),
    "swe_device_count": (C.c_int32, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load libswe_b200.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, ctx=None) -> None:
    if status != OK:
        msg = lib().swe_last_error(ctx)
        raise SweError(status, msg.decode() if msg else "unknown error")


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_D)


def as_f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a
