"""The C-ABI library loads without a GPU, exports every symbol include/h263cu.h declares,
and its device entry points fail loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from h263_rs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "h263cu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(h263cu_[a-z0-9_]+)\s*\(", text))
    names |= set(re.findall(r"extern\s+const\s+\w+\s+(h263cu_[a-z0-9_]+)\s*\[", text))
    return names


def test_header_and_loader_agree():
    assert declared_symbols() == set(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in sorted(declared_symbols()):
        assert hasattr(L, name), name


def test_generator_library_is_separate_from_the_product():
    """The stream generator is its own library (include/h263synth.h): it exports what its header declares, the
    product library does not carry it, and generating streams maps no product code (bench.py --impl reference gets
    its input this way)."""
    import subprocess
    import sys

    from h263_rs_b200 import synth

    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "h263synth.h")).read(), flags=re.S)
    names = set(re.findall(r"\b(h263cu_synth_[a-z0-9_]+)\s*\(", text))
    assert names == set(synth.SYMBOLS)
    for n in names:
        assert hasattr(synth.lib(), n) and not hasattr(_lib.lib(), n)
    code = ("import sys; sys.path.insert(0, %r); from h263_rs_b200 import synth; p = synth.make_stream(176, 144, 2, 1); "
            "m = open('/proc/self/maps').read(); assert 'libh263synth.so' in m and 'libh263cu.so' not in m and len(p) == 2" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_struct_layouts():
    assert C.sizeof(_lib.Pic) == 32 and C.sizeof(_lib.Mb) == 24
    assert _lib.Mb.u.offset == 16 and _lib.Mb.nev.offset == 10 and _lib.Pic.first_mb.offset == 16


def test_strerror_and_classification():
    L = _lib.lib()
    assert L.h263cu_is_eof_error(-16) and not L.h263cu_is_eof_error(-3)
    assert L.h263cu_is_macroblock_error(-3) and L.h263cu_is_macroblock_error(-4) and not L.h263cu_is_macroblock_error(-5)
    assert L.h263cu_is_gob_error(-11)
    for code in list(range(-17, 1)) + [-100, -101, -102, -103, -104, -105, -106]:
        assert L.h263cu_strerror(code)
    table = (C.c_uint8 * 32).in_dll(L, "h263cu_quant_to_strength")
    assert list(table)[:8] == [0, 1, 1, 2, 2, 3, 3, 4] and table[31] == 12


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    L = _lib.lib()
    assert L.h263cu_device_count() == 0
    err = C.c_int(0)
    assert not L.h263cu_create(0, 1, 176, 144, 0, C.byref(err))
    assert err.value == _lib.ERR_NO_DEVICE
    y = np.zeros(16, np.uint8)
    out = np.zeros(64, np.uint8)
    assert L.h263cu_yuv420_to_rgba(y.ctypes.data, y.ctypes.data, y.ctypes.data, 16, 4, out.ctypes.data) == _lib.ERR_NO_DEVICE
    assert L.h263cu_deblock(y.ctypes.data, 16, 4, 3, out.ctypes.data) == _lib.ERR_NO_DEVICE


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (Python, C++, CUDA) names it, and the shipped
    library neither links nor embeds it."""
    import os
    import subprocess

    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".inc", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                for needle in ("oracle/", "oracle_lib", "liboracle", "import oracle", "h263_oracle", "orc_"):
                    assert needle not in text, (os.path.join(root, f), needle)
    needed = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed.lower()
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in syms
