"""swe_fvm_b200 — B200-native explicit finite-volume shallow-water time step (SWE_FVM hot path).

Python host-side mirror of the reference's C++ API over the C-ABI of include/swe_b200.h.
"""
from .capi import (EULER, SSPRK2, SSPRK3, HLL, HLLC, RUSANOV, DAVIS, EINFELDT, SweError)  # noqa: F401
from .mesh import TriangMesh, StructTriangMesh, Case  # noqa: F401
