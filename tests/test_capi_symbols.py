"""The C-ABI library loads on a CPU-only box and exports every symbol include/nsb200.h declares; without a
CUDA device every entry point that needs one fails loudly (no CPU fallback)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nsb = importlib.import_module("3d_navier_stokes_b200")
capi = importlib.import_module("3d_navier_stokes_b200.capi")

pytestmark = pytest.mark.skipif(not os.path.exists(capi.lib_path()), reason="libnsb200.so not built")


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 25
    bound = {s[0] for s in capi.SYMBOLS}
    for n in names:
        assert hasattr(lib.dll, n), "libnsb200.so does not export %s" % n
        assert n in bound, "capi.py does not bind %s" % n
    assert lib.nsb200_version().startswith(b"nsb200")


def test_header_compiles_as_c():
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "nsb200.h"\nint main(void){ return nsb200_version() == 0; }\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", c, "-o", os.path.join(d, "t.o")], check=True)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nsb.Solver(32)


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "3d_navier_stokes_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dp, f)).read()
                assert "ns_oracle" not in txt and "ref_lib" not in txt and "oracle/" not in txt, f
