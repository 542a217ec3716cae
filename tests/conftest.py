import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build (or reuse) the in-tree shared libraries once per session."""
    import __graft_entry__ as g
    g.build()


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


def make_case(kind, n=None, quad_n=4, mesh=None, **params):
    """(mesh, case, v0) for a structured [0,4]^2 domain (or the given mesh)."""
    from swe_fvm_b200 import Case, StructTriangMesh
    if mesh is None:
        mesh = StructTriangMesh(n, n, 4.0 / n)
        case = Case(kind, 2.0, 2.0, 4.0, **params)
    else:
        case = Case(kind, 4.0, 4.0, 8.0, **params)
    case.set_bathymetry(mesh)
    return mesh, case, case.initial_state(mesh, quad_n=quad_n)


def rel_l2(a, b):
    d = np.linalg.norm(np.asarray(a) - np.asarray(b))
    n = np.linalg.norm(np.asarray(b))
    return d / n if n > 0 else d
