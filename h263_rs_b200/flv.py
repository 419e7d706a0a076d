"""FLV container feed (ctypes front for h263cu_flv_scan / h263cu_flv_mux): the caller side of the
reference's `H263Reader::from_source(&packet[..])` -- one reader per FLV video tag, Sorenson Spark
being FLV video codec 2 with one picture per tag."""
import ctypes as C

import numpy as np

from . import _lib

PACKET_DTYPE = np.dtype([("offset", "<u8"), ("size", "<u4"), ("timestamp_ms", "<u4"), ("frame_type", "u1"),
                         ("codec_id", "u1"), ("reserved", "<u2"), ("reserved2", "<u4")])
assert PACKET_DTYPE.itemsize == 24

FRAME_KEY, FRAME_INTER, FRAME_DISPOSABLE_INTER = 1, 2, 3


def scan(data):
    """List the H.263 picture packets of an FLV byte stream (bytes or uint8 array, not copied).
    Returns (packets structured array, number of other tags skipped)."""
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    L = _lib.lib()
    other = C.c_uint32(0)
    n = int(L.h263cu_flv_scan(buf.ctypes.data, buf.size, None, 0, C.byref(other)))
    _lib.check(n)
    out = np.zeros(n, PACKET_DTYPE)
    got = int(L.h263cu_flv_scan(buf.ctypes.data, buf.size, out.ctypes.data, n, C.byref(other)))
    assert got == n
    return out, int(other.value)


def packets(data):
    """The picture packets of an FLV byte stream as a list of `bytes` (what
    `H263State.decode_next_picture` / `BatchDecoder.decode_step` take)."""
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    pk, _ = scan(buf)
    return [bytes(buf[int(p["offset"]) : int(p["offset"]) + int(p["size"])]) for p in pk]


def mux(picture_packets, ms_per_picture=40, filler_every=0, frame_types=None):
    """Wrap a list of picture packets into an FLV byte stream (one video tag per picture)."""
    n = len(picture_packets)
    lens = np.array([len(p) for p in picture_packets], np.uint32)
    offs = np.zeros(n, np.uint64)
    if n > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bytes(p) for p in picture_packets) or b"\0", np.uint8)
    ft = None if frame_types is None else np.ascontiguousarray(frame_types, np.uint8)
    L = _lib.lib()
    args = (blob.ctypes.data, offs.ctypes.data, lens.ctypes.data, None if ft is None else ft.ctypes.data, n,
            ms_per_picture, filler_every)
    need = int(L.h263cu_flv_mux(*args, None, 0))
    _lib.check(need)
    out = np.zeros(need, np.uint8)
    got = int(L.h263cu_flv_mux(*args, out.ctypes.data, out.size))
    assert got == need
    return out.tobytes()
