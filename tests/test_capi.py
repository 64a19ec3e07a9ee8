"""The C-ABI library loads and exports exactly what include/cbn_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from continuousbayesiannetwork_b200.build import build_native

    build_native()
    from continuousbayesiannetwork_b200 import _native

    return _native


def _header_functions():
    src = open(os.path.join(ROOT, "include", "cbn_b200.h")).read()
    return sorted(set(re.findall(r"CBN_API\s+[\w\s\*]+?\b(cbn_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(native):
    lib = native.lib()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(native.SIGNATURES) == declared, "ctypes SIGNATURES and the header disagree"
    assert lib.cbn_abi_version() == 2


def test_struct_layouts_match_the_header(native, tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "cbn_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(cbn_family), sizeof(cbn_contract), sizeof(cbn_gather_table), sizeof(cbn_row_input), sizeof(cbn_row_step));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(native.Family), C.sizeof(native.Contract), C.sizeof(native.GatherTable), C.sizeof(native.RowInput),
                     C.sizeof(native.RowStep)]


def test_errors_without_a_gpu_are_loud(native):
    import torch

    lib = native.lib()
    h = C.c_void_p()
    if not torch.cuda.is_available():
        rc = lib.cbn_ctx_create(0, C.byref(h))
        assert rc == native.ERR_CUDA
        assert b"no CUDA device" in lib.cbn_last_error(None)
        with pytest.raises(RuntimeError):
            native.context_for("cuda")
    with pytest.raises(RuntimeError):
        native.context_for("cpu")
    # NULL context -> invalid argument, message retrievable
    assert lib.cbn_count_run(None, None, None, 0, 0, None, None) == native.ERR_INVALID
    assert b"ctx is NULL" in lib.cbn_last_error(None)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or execute it."""
    pkg = os.path.join(ROOT, "continuousbayesiannetwork_b200")
    pat = re.compile(r"^\s*(import|from)\s+oracle\b|libcbn_oracle|build_oracle|cbn_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_missing_library_is_an_import_error(native, monkeypatch):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libcbn_b200.so")
    with pytest.raises(ImportError):
        native.lib()
