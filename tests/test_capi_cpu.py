"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads, exports
every symbol include/vhp.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import visibility_heuristic_path_planner_b200 as vhp
    return vhp.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vhp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vhp_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version(lib):
    assert lib.vhp_abi_version() == 1
    assert b"sm_100a" in lib.vhp_version_string()


def test_no_cpu_fallback(lib):
    """Without a device every compute entry point must fail loudly."""
    import visibility_heuristic_path_planner_b200 as vhp
    if lib.vhp_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vhp.VhpError) as e:
        vhp.Context(0)
    assert e.value.status == -2  # VHP_ERR_NO_DEVICE


def test_product_does_not_touch_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "visibility_heuristic_path_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in src and "vhp_oracle" not in src, f
                assert "libvhp_ref" not in src and "/root/reference" not in src, f


def test_headers_compile_standalone(tmp_path):
    """include/vhp.h is plain C99 (MATLAB loadlibrary, cgo, ctypes generators read it);
    include/vhp_solver.hpp is the C++20 drop-in for the reference's classes."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "vhp.h"\nint main(void) { return vhp_abi_version() > 0 ? 0 : 1; }\n')
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "vhp_solver.hpp"\nint main() { return 0; }\n')
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                    "-I", inc, str(c)], check=True)
    subprocess.run(["g++", "-std=c++20", "-Wall", "-fsyntax-only", "-I", inc, str(cpp)], check=True)


def test_packed_handle_api_without_a_handle(lib):
    """The accessors of the packed handle take NULL; expanding without a handle is an argument error."""
    import numpy as np
    assert lib.vhp_packed_pairs(None) == 0 and lib.vhp_packed_bytes(None) == 0 and lib.vhp_packed_pair_bytes(None) == 0
    out = np.zeros(4, np.float32)
    assert lib.vhp_packed_expand(None, 0, 1, out.ctypes.data, 1) == -1
    lib.vhp_packed_destroy(None)
