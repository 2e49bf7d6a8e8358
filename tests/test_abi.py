"""CPU-only checks of the drop-in boundary: the library loads without a GPU and exports exactly what include/*.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

import poulpy_b200 as pb
from poulpy_b200 import hal


def _declared():
    src = open(hal.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pb.lib()
    names = _declared()
    assert len(names) >= 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    out = subprocess.check_output(["nm", "-D", "--defined-only", hal.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r" T (pgb_[a-z0-9_]+)", out)))
    assert exported == names, set(exported) ^ set(names)


def test_no_torch_types_in_abi():
    src = re.sub(r"/\*.*?\*/", "", open(hal.HEADER_PATH).read(), flags=re.S)
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_module_new_fails_loudly_without_gpu():
    lib = pb.lib()
    if lib.pgb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PoulpyError) as e:
        pb.Module(1024, pb.NTT120)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under poulpy_b200/ may import, include or link it."""
    root = os.path.dirname(hal.HEADER_PATH)
    pkg = os.path.join(os.path.dirname(root), "poulpy_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "poulpy_oracle" not in txt and "liboracle" not in txt, f
    out = subprocess.check_output(["ldd", hal.LIB_PATH]).decode()
    assert "oracle" not in out


def test_struct_layouts_match_header():
    assert C.sizeof(hal._VZ) == 40 and C.sizeof(hal._PP) == 24 and C.sizeof(hal._PM) == 48 and C.sizeof(hal._BT) == 32


def test_rust_ffi_is_generated_from_the_header():
    """rust/poulpy-gpu-b200/src/ffi.rs (the `extern "C"` block of the backend crate of INTEGRATION.md) is generated from the header: the
    committed file must equal what scripts/gen_rust_ffi.py produces, and it binds every exported symbol."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(hal.HEADER_PATH))
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(root, "scripts", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    src, fns = gen.generate()
    assert open(gen.OUT).read() == src, "run python scripts/gen_rust_ffi.py"
    assert sorted(n for n, _, _ in fns) == _declared()
