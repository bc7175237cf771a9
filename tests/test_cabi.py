"""The C-ABI library loads on a machine without a GPU and exports every symbol include/pandaseq_b200.h declares.
No compute calls here."""
import ctypes
import os
import re

import pytest

import pandaseq_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pandaseq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline[^{]*\{.*?\n\}", "", text, flags=re.S)
    funcs = set(re.findall(r"\b((?:panda|pb)_[a-z0-9_]+)\s*\(", text))
    text = text.replace('extern "C"', "")
    data = set(re.findall(r"\bextern\s+[^;({]*?\b((?:panda|pb)_[a-z0-9_]+)\s*;", text))
    typedef_fp = set(re.findall(r"\(\*\s*(\w+)\s*\)", text))
    return (funcs | data) - typedef_fp


def test_library_loads_without_gpu(built):
    L = pb.lib()
    assert L.panda_max_len() == 450


def test_every_declared_symbol_is_exported(built):
    L = ctypes.CDLL(pb.LIB_PATH)
    missing = [s for s in sorted(declared_symbols()) if not hasattr(L, s)]
    assert not missing, missing
    assert len(declared_symbols()) > 70


def test_no_oracle_linkage(built):
    """The product must not link or load anything under oracle/."""
    import subprocess
    out = subprocess.run(["ldd", pb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "pandaseq_ref" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", pb.LIB_PATH], capture_output=True, text=True).stdout
    assert "po_" not in syms and "ref_assemble" not in syms


def test_compute_fails_loudly_without_gpu(built):
    L = pb.lib()
    if L.pb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PandaseqError):
        pb.Context(0)
    L.panda_assembler_new.restype = ctypes.c_void_p
    L.panda_assembler_new.argtypes = [ctypes.c_void_p] * 4
    assert L.panda_assembler_new(None, None, None, None) is None      # no device, no assembler: there is no CPU path


def test_c_demo_compiles_and_links_without_gpu(built, tmp_path):
    """The drop-in demo builds against the public header alone (it is run in the GPU tests)."""
    import subprocess
    exe = str(tmp_path / "dropin_demo")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "dropin_demo.c"), "-L", os.path.join(ROOT, "pandaseq_b200"), "-lpandaseq_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "pandaseq_b200"), "-o", exe], check=True)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 2          # usage
