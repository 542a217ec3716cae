"""The C-ABI shared library: loads, exports every symbol include/swe_b200.h declares, and the
compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "swe_b200.h")).read()
    return sorted(set(re.findall(r"SWE_API\s+[\w\s\*]+?\b(swe_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from swe_fvm_b200 import capi
    lib = C.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding table covers exactly the header
    assert sorted(capi.SYMBOLS) == names


def test_version_and_error_strings():
    from swe_fvm_b200 import capi
    assert b"sm_100a" in capi.lib().swe_version()


def test_host_api_argument_validation():
    from swe_fvm_b200 import StructTriangMesh, SweError, capi
    with pytest.raises(SweError):
        StructTriangMesh(0, 4, 1.0)
    with pytest.raises(SweError):
        StructTriangMesh(4, 4, -1.0)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from swe_fvm_b200 import StructTriangMesh, SweError
    from swe_fvm_b200.solver import SpaceDisc
    m = StructTriangMesh(4, 4, 1.0)
    with pytest.raises(SweError) as ei:
        SpaceDisc("hllc", "einfeldt", m)
    assert ei.value.status == -2 and "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_meshes_before_touching_the_gpu():
    from swe_fvm_b200 import StructTriangMesh, SweError, capi
    m = StructTriangMesh(3, 3, 1.0)
    cm = m.c_mesh()
    bad = np.array(m.edge_elements, dtype=np.int64).copy()
    bad[0, 1] = -2 if bad[0, 1] < 0 else bad[0, 1]  # FREE_FLOW is not implemented upstream
    first_wall = int(np.nonzero(m.edge_elements[:, 1] < 0)[0][0])
    bad[first_wall, 1] = -2
    cm.edge_elements = bad.ctypes.data_as(C.POINTER(C.c_int64))
    ctx = C.c_void_p()
    rc = capi.lib().swe_create(C.byref(ctx), C.byref(cm), 0, 0)
    assert rc == -1
    assert b"SOLID_WALL" in capi.lib().swe_last_error(None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no product module may import, load or link it."""
    pkg = os.path.join(ROOT, "swe_fvm_b200")
    pat = re.compile(r"(from\s+oracle|import\s+oracle|libswe_oracle|oracle/|oracle\.oracle|swe_oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(txt), (f, pat.search(txt).group(0))
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_header_is_plain_c(tmp_path):
    """include/swe_b200.h must be consumable from C (the boundary is a C ABI, not a C++ one)."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "swe_b200.h"\n'
                   'int main(void) { swe_mesh m; swe_case c; swe_ctx *ctx = 0; (void)m; (void)c; (void)ctx;\n'
                   '  return (int)(SWE_OK + SWE_SSPRK2 - SWE_SSPRK2 + SWE_HLLC - SWE_HLLC); }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
