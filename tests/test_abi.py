"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports
every symbol include/fus_b200.h declares (no compute calls here)."""

import os
import re

from fenicsx_fus_gpu_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fus_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fus_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.lib()
    assert lib.fus_abi_version() == 1
    names = _declared()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(_lib.exported_symbols()) == _declared()


def test_error_reporting_without_gpu():
    lib = _lib.lib()
    # argument validation happens before any CUDA call
    rc = lib.fus_stiffness_f64(None, None, None, None, None, None, 10, 9, 0, None)
    assert rc == 100001  # FUS_ERR_BAD_DEGREE
    assert b"degree" in lib.fus_last_error()
    rc = lib.fus_mass_f32(None, None, None, None, None, -1, 8, None)
    assert rc == 100002  # FUS_ERR_BAD_ARGUMENT
    assert lib.fus_axpy_f64(1.0, None, None, 0, None) == 0  # empty vector: no launch


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "fenicsx_fus_gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libfus_oracle" not in txt and "fus_oracle" not in txt, f
