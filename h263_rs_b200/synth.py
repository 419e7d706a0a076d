"""Synthetic Sorenson-flavour stream generator (ctypes front for h263cu_synth_stream)."""
import ctypes as C

import numpy as np

from . import _lib


def default_params(width, height, n_pictures, seed, **overrides):
    p = _lib.SynthParams()
    _lib.lib().h263cu_synth_default_params(C.byref(p), width, height, n_pictures, seed)
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
    L = _lib.lib()
    n = p.n_pictures
    off = np.zeros(n, np.uint64)
    ln = np.zeros(n, np.uint32)
    need = L.h263cu_synth_stream(C.byref(p), None, 0, None, None)
    _lib.check(int(need))
    blob = np.zeros(max(int(need), 1), np.uint8)
    got = L.h263cu_synth_stream(C.byref(p), blob.ctypes.data, blob.size, off.ctypes.data, ln.ctypes.data)
    assert got == need
    return blob[: int(need)], off, ln
