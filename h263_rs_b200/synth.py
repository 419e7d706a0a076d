"""Synthetic Sorenson-flavour stream generator: ctypes front for libh263synth.so (include/h263synth.h).

The generator is test / benchmark tooling in its own library: importing this module does not load the product
library (bench.py's CPU reference arm gets its streams from here without mapping any product code)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libh263synth.so")


class SynthParams(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("n_pictures", C.c_uint32), ("seed", C.c_uint64),
        ("flavour", C.c_uint32), ("version", C.c_uint32), ("intra_period", C.c_uint32), ("deblock_flag", C.c_uint32),
        ("qp_min", C.c_uint32), ("qp_max", C.c_uint32), ("pct_uncoded", C.c_uint32), ("pct_intra", C.c_uint32),
        ("pct_fourmv", C.c_uint32), ("pct_dquant", C.c_uint32), ("pct_cbp_inter", C.c_uint32),
        ("pct_cbp_intra", C.c_uint32), ("mean_events_x10", C.c_uint32), ("pct_escape", C.c_uint32),
        ("permille_overflow", C.c_uint32), ("mv_mode", C.c_uint32), ("truncate_permille", C.c_uint32),
        ("pct_disposable", C.c_uint32), ("reserved", C.c_uint32 * 3),
    ]


SYMBOLS = ["h263cu_synth_default_params", "h263cu_synth_stream"]
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libh263synth.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.h263cu_synth_default_params.argtypes = [C.POINTER(SynthParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    L.h263cu_synth_default_params.restype = None
    L.h263cu_synth_stream.restype = C.c_int64
    L.h263cu_synth_stream.argtypes = [C.POINTER(SynthParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def default_params(width, height, n_pictures, seed, **overrides):
    p = SynthParams()
    lib().h263cu_synth_default_params(C.byref(p), width, height, n_pictures, seed)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def params_dict(p):
    return {f: getattr(p, f) for f, _ in p._fields_ if f != "reserved"}


def make_stream(width, height, n_pictures, seed, **overrides):
    """Returns a list of `bytes`, one packet per picture."""
    blob, off, ln = make_stream_blob(default_params(width, height, n_pictures, seed, **overrides))
    return [bytes(blob[int(o) : int(o) + int(l)]) for o, l in zip(off, ln)]


def make_stream_blob(p):
    """Returns (blob uint8 array, pkt_off uint64 array, pkt_len uint32 array)."""
    L = lib()
    n = p.n_pictures
    off = np.zeros(n, np.uint64)
    ln = np.zeros(n, np.uint32)
    need = int(L.h263cu_synth_stream(C.byref(p), None, 0, None, None))
    if need < 0:
        raise ValueError("h263cu_synth_stream: bad parameters (%d)" % need)
    blob = np.zeros(max(need, 1), np.uint8)
    got = L.h263cu_synth_stream(C.byref(p), blob.ctypes.data, blob.size, off.ctypes.data, ln.ctypes.data)
    assert got == need
    return blob[:need], off, ln
