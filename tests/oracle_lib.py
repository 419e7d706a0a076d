"""ctypes binding of the CPU oracle (oracle/liboracle_h263.so).

TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle_h263.so")

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
i16p = C.POINTER(C.c_int16)
f32p = C.POINTER(C.c_float)

ERR_NAMES = {
    0: "Ok", 1: "InternalDecoderError", 2: "MiddleOfBitstream", 3: "InvalidMacroblockHeader",
    4: "InvalidMacroblockCodedBits", 5: "InvalidIntraDc", 6: "InvalidShortCoefficient",
    7: "InvalidLongCoefficient", 8: "InvalidMvd", 9: "InvalidPType", 10: "InvalidPlusPType",
    11: "InvalidGobHeader", 12: "InvalidBitstream", 13: "PictureFormatMissing",
    14: "PictureFormatInvalid", 15: "UncodedIFrameBlocks", 16: "UnhandledIoError(UnexpectedEof)",
    17: "UnimplementedDecoding", 100: "ReferenceWouldAbort",
}


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "h263_oracle.cpp")
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.orc_state_new.restype = C.c_void_p
    L.orc_state_new.argtypes = [C.c_int]
    L.orc_state_free.argtypes = [C.c_void_p]
    L.orc_decode_next_picture.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.orc_last_picture_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 7
    L.orc_last_picture_yuv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_state_set_trace.argtypes = [C.c_void_p, C.c_int]
    L.orc_trace_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_trace_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 8
    L.orc_yuv420_to_rgba.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    L.orc_yuv420_to_rgba.restype = None
    L.orc_deblock.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    L.orc_deblock.restype = None
    L.orc_deblock_process.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_deblock_process.restype = None
    L.orc_inverse_rle.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_idct_block.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.orc_idct_block.restype = None
    L.orc_idct_1d.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_idct_1d.restype = None
    L.orc_gather_block.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]
    L.orc_gather_block.restype = None
    L.orc_read_vlc.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.orc_read_bits.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_int64)]
    L.orc_recognize_start_code.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
    L.orc_decode_block.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int]
    L.orc_bench_decode.restype = C.c_double
    L.orc_bench_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_batch_new.restype = C.c_void_p
    L.orc_batch_new.argtypes = [C.c_int, C.c_int]
    L.orc_batch_free.argtypes = [C.c_void_p]
    L.orc_batch_step.restype = C.c_double
    L.orc_batch_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_batch_stream_yuv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle error %d (%s)" % (code, ERR_NAMES.get(code, "?")))
        self.code = code


class OracleState:
    """Mirror of h263::H263State (decoder/state.rs) backed by the oracle."""

    SORENSON = 1

    def __init__(self, options=1, trace=False):
        self.L = lib()
        self.h = self.L.orc_state_new(options)
        if trace:
            self.L.orc_state_set_trace(self.h, 1)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_state_free(self.h)
            self.h = None

    def decode_next_picture(self, packet: bytes):
        e = self.L.orc_decode_next_picture(self.h, packet, len(packet))
        if e:
            raise OracleError(e)

    def info(self):
        v = [C.c_int() for _ in range(7)]
        if self.L.orc_last_picture_info(self.h, *[C.byref(x) for x in v]):
            return None
        k = ("width", "height", "tr", "ptype", "quant", "deblock", "version")
        return dict(zip(k, [x.value for x in v]))

    def yuv(self):
        i = self.info()
        w, h = i["width"], i["height"]
        cw, ch = (w + 1) // 2, (h + 1) // 2
        y = np.empty(w * h, np.uint8)
        cb = np.empty(cw * ch, np.uint8)
        cr = np.empty(cw * ch, np.uint8)
        self.L.orc_last_picture_yuv(self.h, _ptr(y), _ptr(cb), _ptr(cr))
        return y, cb, cr

    def trace(self):
        n, ne = C.c_int(), C.c_int()
        self.L.orc_trace_counts(self.h, C.byref(n), C.byref(ne))
        n, ne = n.value, ne.value
        t = dict(
            mb_type=np.empty(n, np.int8), coded=np.empty(n, np.int8), quant=np.empty(n, np.uint8),
            mv=np.empty(n * 8, np.int16), intradc=np.empty(n * 6, np.int16), nev=np.empty(n * 6, np.uint8),
            run=np.empty(ne, np.uint8), level=np.empty(ne, np.int16),
        )
        self.L.orc_trace_copy(self.h, *[_ptr(t[k]) for k in ("mb_type", "coded", "quant", "mv", "intradc", "nev",
                                                             "run", "level")])
        t["mv"] = t["mv"].reshape(n, 4, 2)
        t["intradc"] = t["intradc"].reshape(n, 6)
        t["nev"] = t["nev"].reshape(n, 6)
        return t


def yuv420_to_rgba(y, cb, cr, width):
    y = np.ascontiguousarray(y, np.uint8)
    cb = np.ascontiguousarray(cb, np.uint8)
    cr = np.ascontiguousarray(cr, np.uint8)
    out = np.empty(y.size * 4, np.uint8)
    lib().orc_yuv420_to_rgba(_ptr(y), _ptr(cb), _ptr(cr), y.size, width, _ptr(out))
    return out


def deblock(data, width, strength):
    data = np.ascontiguousarray(data, np.uint8)
    out = np.empty(data.size, np.uint8)
    lib().orc_deblock(_ptr(data), data.size, width, strength, _ptr(out))
    return out


def deblock_process(abcd, strength, simd):
    a = np.array(abcd, np.uint8)
    lib().orc_deblock_process(_ptr(a), strength, int(simd))
    return [int(x) for x in a]


def inverse_rle(intradc_code, runs, levels, quant):
    r = np.array(runs, np.uint8)
    l = np.array(levels, np.int16)
    out = np.zeros(64, np.float32)
    cls = lib().orc_inverse_rle(-1 if intradc_code is None else intradc_code, len(r), _ptr(r), _ptr(l), quant,
                                _ptr(out))
    return cls, out.reshape(8, 8)


def idct_block(cls, coefs, pixels):
    c = np.ascontiguousarray(coefs, np.float32).reshape(64)
    p = np.ascontiguousarray(pixels, np.uint8).reshape(64).copy()
    lib().orc_idct_block(cls, _ptr(c), _ptr(p))
    return p.reshape(8, 8)


def idct_1d(v):
    a = np.ascontiguousarray(v, np.float32)
    o = np.empty(8, np.float32)
    lib().orc_idct_1d(_ptr(a), _ptr(o))
    return o


def gather_block(src, pos, mv, dst):
    h, w = src.shape
    s = np.ascontiguousarray(src, np.uint8)
    d = np.ascontiguousarray(dst, np.uint8).copy()
    lib().orc_gather_block(_ptr(s), w, h, pos[0], pos[1], mv[0], mv[1], _ptr(d))
    return d
