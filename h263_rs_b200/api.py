"""Host-side mirror of the reference's public API, backed by libh263cu.so.

Names, argument meaning and error behaviour follow the reference so that the parity
tests read like the reference's own tests:

    h263::H263State::{new, decode_next_picture, get_last_picture, is_sorenson}   state.rs:40-141
    h263::DecodedPicture::{as_yuv, as_luma, as_chroma_b, as_chroma_r, ...}         picture.rs:60-142
    yuv::bt601::yuv420_to_rgba(y, chroma_b, chroma_r, y_width) -> Vec<u8>          bt601.rs:105
    deblock::deblock::deblock(data, width, strength) -> Vec<u8>, QUANT_TO_STRENGTH deblock.rs:5-8,305

plus `BatchDecoder`, the batched form the GPU path is built for (one picture for each of
many independent streams per step).  All compute runs on the GPU through the C ABI; there
is no CPU fallback (errors surface as H263Error).
"""
import ctypes as C

import numpy as np

from . import _lib, frontend
from ._lib import H263Error, check

SORENSON_SPARK_BITSTREAM = _lib.OPT_SORENSON
USE_SCALABILITY_MODE = 2


def _quant_to_strength():
    arr = (C.c_uint8 * 32).in_dll(_lib.lib(), "h263cu_quant_to_strength")
    return [int(v) for v in arr]


class _LazyTable:
    def __init__(self):
        self._v = None

    def _get(self):
        if self._v is None:
            self._v = _quant_to_strength()
        return self._v

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return 32

    def __iter__(self):
        return iter(self._get())


QUANT_TO_STRENGTH = _LazyTable()


def yuv420_to_rgba(y, chroma_b, chroma_r, y_width):
    """yuv::bt601::yuv420_to_rgba: planar YUV 4:2:0 -> interleaved RGBA8888 (GPU)."""
    y = np.ascontiguousarray(y, np.uint8).reshape(-1)
    cb = np.ascontiguousarray(chroma_b, np.uint8).reshape(-1)
    cr = np.ascontiguousarray(chroma_r, np.uint8).reshape(-1)
    out = np.empty(y.size * 4, np.uint8)
    if y.size == 0:
        return out
    check(_lib.lib().h263cu_yuv420_to_rgba(y.ctypes.data, cb.ctypes.data, cr.ctypes.data, y.size, y_width,
                                           out.ctypes.data))
    return out


def deblock(data, width, strength):
    """deblock::deblock::deblock: Annex-J-style post filter on one plane (GPU)."""
    data = np.ascontiguousarray(data, np.uint8).reshape(-1)
    out = np.empty(data.size, np.uint8)
    if data.size == 0:
        return out
    check(_lib.lib().h263cu_deblock(data.ctypes.data, data.size, width, strength, out.ctypes.data))
    return out


class DecodedPicture:
    """Mirror of h263::DecodedPicture: header fields + tight row-major u8 planes."""

    def __init__(self, width, height, picture_type, quantizer, temporal_reference, y, cb, cr):
        self.width, self.height = width, height
        self.picture_type = picture_type
        self.quantizer = quantizer
        self.temporal_reference = temporal_reference
        self._y, self._cb, self._cr = y, cb, cr

    def as_yuv(self):
        return self._y, self._cb, self._cr

    def as_luma(self):
        return self._y

    def as_chroma_b(self):
        return self._cb

    def as_chroma_r(self):
        return self._cr

    def format(self):
        """The picture's dimensions (DecodedPicture::format, picture.rs:60-66) as (width, height)."""
        return self.width, self.height

    def as_header(self):
        """The header fields the reference keeps with a decoded picture (DecodedPicture::as_header)."""
        return {"width": self.width, "height": self.height, "picture_type": self.picture_type,
                "quantizer": self.quantizer, "temporal_reference": self.temporal_reference}

    def luma_samples_per_row(self):
        return self.width

    def chroma_samples_per_row(self):
        return (self.width + 1) // 2


class Context:
    """Thin RAII wrapper of h263cu_ctx."""

    def __init__(self, device, max_streams, max_width, max_height, _borrowed=None):
        self.L = _lib.lib()
        self._owned = _borrowed is None
        if _borrowed is None:
            err = C.c_int(0)
            self.h = self.L.h263cu_create(device, max_streams, max_width, max_height, 0, C.byref(err))
            if not self.h:
                raise H263Error(err.value)
        else:
            self.h = _borrowed  # a context that belongs to a DeviceGroup
        self.max_streams, self.max_width, self.max_height = max_streams, max_width, max_height

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.L.h263cu_destroy(self.h)
            self.h = None

    def readback_wait(self, age=0):
        check(self.L.h263cu_readback_wait(self.h, age))

    def __del__(self):
        self.close()

    def submit_step(self, pics, mbs, events, out_flags):
        check(self.L.h263cu_submit_step(self.h, pics.ctypes.data, len(pics), mbs.ctypes.data, len(mbs),
                                        events.ctypes.data if len(events) else None, len(events), out_flags))

    def step_upload(self, pics, mbs, events):
        err = C.c_int(0)
        s = self.L.h263cu_step_upload(self.h, pics.ctypes.data, len(pics), mbs.ctypes.data, len(mbs),
                                      events.ctypes.data if len(events) else None, len(events), C.byref(err))
        if not s:
            raise H263Error(err.value)
        return s

    def step_run(self, step, out_flags):
        check(self.L.h263cu_step_run(self.h, step, out_flags))

    def step_free(self, step):
        self.L.h263cu_step_free(self.h, step)

    def graph_build(self, steps, out_flags):
        """n resident steps as one CUDA graph (h263cu_graph_build); the steps must outlive it."""
        arr = (C.c_void_p * len(steps))(*steps)
        err = C.c_int(0)
        g = self.L.h263cu_graph_build(self.h, arr, len(steps), out_flags, C.byref(err))
        if not g:
            raise H263Error(err.value)
        return g

    def graph_launch(self, graph):
        check(self.L.h263cu_graph_launch(self.h, graph))

    def graph_free(self, graph):
        self.L.h263cu_graph_free(self.h, graph)

    def sync(self):
        check(self.L.h263cu_sync(self.h))

    def stream_info(self, stream):
        v = [C.c_uint32() for _ in range(5)]
        check(self.L.h263cu_stream_info(self.h, stream, *[C.byref(x) for x in v]))
        return dict(zip(("width", "height", "pic_type", "pquant", "tr"), [x.value for x in v]))

    def read_yuv(self, stream):
        i = self.stream_info(stream)
        w, h = i["width"], i["height"]
        cw, ch = (w + 1) // 2, (h + 1) // 2
        y = np.empty(w * h, np.uint8)
        cb = np.empty(cw * ch, np.uint8)
        cr = np.empty(cw * ch, np.uint8)
        check(self.L.h263cu_read_yuv(self.h, stream, y.ctypes.data, cb.ctypes.data, cr.ctypes.data))
        return y, cb, cr

    def read_rgba(self, stream):
        i = self.stream_info(stream)
        out = np.empty(i["width"] * i["height"] * 4, np.uint8)
        check(self.L.h263cu_read_rgba(self.h, stream, out.ctypes.data))
        return out

    def checksums(self, streams):
        s = np.ascontiguousarray(streams, np.uint32)
        out = np.zeros((len(s), 4), np.uint64)
        check(self.L.h263cu_checksums(self.h, s.ctypes.data, len(s), out.ctypes.data))
        return out

    def timer_start(self):
        check(self.L.h263cu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.L.h263cu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.h263cu_launch_count(self.h))

    def tiled_launch_count(self):
        return int(self.L.h263cu_tiled_launch_count(self.h))

    def profile_enable(self, on=True):
        check(self.L.h263cu_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms = (C.c_double * 2)()
        n = (C.c_uint64 * 2)()
        check(self.L.h263cu_profile_read(self.h, ms, n))
        return dict(recon_ms=ms[0], recon_launches=int(n[0]), deblock_ms=ms[1], deblock_launches=int(n[1]))


class H263State:
    """Mirror of h263::H263State for ONE stream: host parse + GPU reconstruction.

    `decode_next_picture(packet)` takes the bytes of one picture (the reference reads them
    through `H263Reader::from_source(&packet[..])`; Sorenson pictures end at EOF, so one
    reader per packet is the only usable form)."""

    def __init__(self, decoder_options=SORENSON_SPARK_BITSTREAM, device=0, deblock=False, pipelined=False):
        """pipelined=True: decode_next_picture returns as soon as the picture is queued (parse done, upload + kernels +
        RGBA read-back in flight); get_last_rgba / get_last_picture wait for it.  A caller that hands in packet t + 1
        before it consumes picture t overlaps the host parse with the device work (two pinned RGBA buffers alternate)."""
        self.decoder_options = decoder_options
        self.device = device
        self.pipelined = pipelined
        self._ring, self._ring_size, self._ring_pos, self._pending = [None, None], 0, 0, False
        self._last_dims = (0, 0)
        self._views, self._views_size = [None, None], -1
        self.parser = frontend.Parser(decoder_options)
        self.ctx = None
        self.out_flags = _lib.OUT_RGBA | (_lib.OUT_DEBLOCK if deblock else 0)
        self._has_picture = False
        # the one-element argument arrays of h263cu_decode_step; the pinned RGBA ring is self._ring
        self._rgba_view = None
        self._one_parser = (C.c_void_p * 1)(self.parser.h)
        self._one_packet = (C.c_void_p * 1)()
        self._one_len = (C.c_size_t * 1)()
        self._one_id = np.zeros(1, np.uint32)
        self._one_err = np.zeros(1, np.int32)

    def is_sorenson(self):
        return bool(self.decoder_options & SORENSON_SPARK_BITSTREAM)

    def decode_next_picture(self, packet, _force_new_ctx=False):
        """One call into h263cu_decode_step for the one stream: parse into the context's pinned staging, upload,
        reconstruction and the RGBA read-back into this state's pinned buffer, then wait.  Transactional like the
        reference (state.rs:120-137): a packet that fails to parse raises and leaves parser and stream untouched."""
        data = packet if isinstance(packet, bytes) else bytes(packet)
        buf = np.frombuffer(data, np.uint8)
        # A picture larger than the context needs a bigger one.  It replaces the old context only once the packet has
        # decoded (a size change can only succeed on an I picture, which needs no reference planes): a packet that
        # fails leaves parser, context and last picture as they were (state.rs:120-137).  With a context in place the
        # header is not peeked first: decode_step reports a picture that does not fit (H263CU_ERR_CAPACITY) and the
        # call is repeated with a larger context.
        ctx = self.ctx
        if ctx is not None and not _force_new_ctx:
            w, h = self._last_dims
        else:
            hdr = frontend.peek_picture(data, self.decoder_options)
            w, h = int(hdr["width"]), int(hdr["height"])
            if w == 0 or h == 0:
                w = h = 16  # no usable size in the header: the parse inside decode_step reports the reference's error
            if ctx is None or w > ctx.max_width or h > ctx.max_height:
                ctx = Context(self.device, 1, max(w, 16), max(h, 16))
        L = _lib.lib()
        size = w * h * 4
        if size > self._ring_size:
            # both pinned buffers grow together; nothing may still be copying into the old ones
            if self.ctx is not None:
                self.ctx.sync()
            for k in range(2):
                if self._ring[k]:
                    L.h263cu_free_pinned(self._ring[k])
                self._ring[k] = L.h263cu_alloc_pinned(size)
                if not self._ring[k]:
                    self._ring_size = 0
                    raise MemoryError
            self._ring_size = size
            self._views_size = -1
        pos = self._ring_pos ^ 1  # the buffer that does not hold the last picture
        nd = C.c_uint32(0)
        self._one_packet[0], self._one_len[0], self._one_err[0] = buf.ctypes.data, buf.size, 0
        _lib.check(L.h263cu_decode_step(ctx.h, self._one_parser, self._one_packet, self._one_len, self._one_id.ctypes.data, 1, 1,
                                        self.out_flags, self._ring[pos], 0, self._one_err.ctypes.data, C.byref(nd)))
        if self._one_err[0]:
            if self._one_err[0] == _lib.ERR_CAPACITY and not _force_new_ctx:
                return self.decode_next_picture(data, _force_new_ctx=True)  # larger than the context: peek and grow
            raise _lib.H263Error(int(self._one_err[0]))
        d = L.h263cu_stream_dims(ctx.h, 0)  # the size the parser found (it may differ from the previous picture's)
        if d != (w << 16 | h):
            w, h = d >> 16, d & 0xFFFF
            size = w * h * 4
            assert size <= self._ring_size  # a picture that fits the context fits the buffers made with it
        self._last_dims = (w, h)
        if ctx is not self.ctx and self.ctx is not None:
            self.ctx.sync()  # the old context may still be copying the previous picture back
        self.ctx = ctx
        self._ring_pos = pos
        if self._views_size != size:  # numpy views of the two pinned buffers, made once per picture size
            self._views = [np.ctypeslib.as_array(C.cast(self._ring[k], C.POINTER(C.c_uint8)), shape=(size,)) for k in range(2)]
            self._views_size = size
        self._rgba_view = self._views[pos]
        self._has_picture = True
        self._pending = True
        if not self.pipelined:
            self._wait()

    def _wait(self):
        """Wait for the picture queued last: its planes are reconstructed and its RGBA has arrived in host memory."""
        if self._pending:
            self.ctx.readback_wait(0)
            self._pending = False

    def __del__(self):
        try:
            if getattr(self, "ctx", None) is not None:
                self.ctx.sync()
            for k in range(2):
                if self._ring[k]:
                    _lib.lib().h263cu_free_pinned(self._ring[k])
                    self._ring[k] = None
        except Exception:
            pass

    def get_last_picture(self):
        if not self._has_picture:
            return None
        self._wait()
        i = self.ctx.stream_info(0)
        y, cb, cr = self.ctx.read_yuv(0)
        return DecodedPicture(i["width"], i["height"], i["pic_type"], i["pquant"], i["tr"], y, cb, cr)

    def get_reference_picture(self):
        """state.rs:72-78.  The reference picture is the last non-disposable picture; disposable
        pictures do not decode in the reference (macroblock.rs:461-465), so it is the last picture."""
        return self.get_last_picture()

    def parse_picture(self, packet):
        """Header-only peek (H263State::parse_picture, state.rs:102-111): no decoder state changes."""
        return frontend.peek_picture(bytes(packet), self.decoder_options)

    def cleanup_buffers(self):
        """state.rs:81-98 drops every picture except the last and the reference one.  The device context
        never holds more than those two plane slots per stream, so there is nothing to free."""
        return None

    def get_last_rgba(self, copy=True):
        """RGBA of the last picture (fused yuv420_to_rgba, or deblock + yuv420_to_rgba).  copy=False returns a view of
        the pinned read-back buffer, valid until the picture after next is queued (two buffers alternate)."""
        if not self._has_picture:
            return None
        self._wait()
        return self._rgba_view.copy() if copy else self._rgba_view


class BatchDecoder:
    """n independent streams decoded in lock step: one picture per stream per step."""

    def __init__(self, n_streams, max_width, max_height, decoder_options=SORENSON_SPARK_BITSTREAM, device=0,
                 threads=0, ctx=None):
        self.n = n_streams
        self.parsers = [frontend.Parser(decoder_options) for _ in range(n_streams)]
        self.ctx = ctx if ctx is not None else Context(device, n_streams, max_width, max_height)
        self.threads = threads

    def parse_step(self, packets, stream_ids=None):
        ids = np.arange(self.n, dtype=np.uint32) if stream_ids is None else np.asarray(stream_ids, np.uint32)
        parsers = [self.parsers[int(s)] for s in ids]
        return frontend.parse_step(parsers, packets, ids, self.threads)

    def decode_step(self, packets, out_flags=_lib.OUT_RGBA, stream_ids=None, host_rgba=None, rgba_stride=0):
        """One decode_next_picture per stream through h263cu_decode_step: threaded parse into the
        context's pinned staging, asynchronous upload + reconstruction (+ RGBA read-back into the
        pinned buffer `host_rgba` when given).  Returns the per-picture error codes."""
        plan = self.plan_step(packets, stream_ids)
        return self.decode_planned(plan, out_flags, host_rgba, rgba_stride)

    def plan_step(self, packets, stream_ids=None):
        """The pointer arrays h263cu_decode_step takes, built once for a list of packets (bytes or
        uint8 arrays, kept alive by the plan)."""
        n = len(packets)
        ids = np.arange(self.n, dtype=np.uint32)[:n] if stream_ids is None else np.ascontiguousarray(stream_ids, np.uint32)
        bufs = [np.frombuffer(pk, np.uint8) if not isinstance(pk, np.ndarray) else pk for pk in packets]
        return {
            "n": n, "bufs": bufs, "ids": ids,
            # an id beyond the decoder's streams is passed on as it is: the library refuses the step
            "parsers": (C.c_void_p * n)(*[self.parsers[min(int(s), self.n - 1)].h for s in ids]),
            "packets": (C.c_void_p * n)(*[b.ctypes.data for b in bufs]),
            "lens": (C.c_size_t * n)(*[b.size for b in bufs]),
            "errs": np.zeros(n, np.int32),
        }

    def decode_planned(self, plan, out_flags=_lib.OUT_RGBA, host_rgba=None, rgba_stride=0):
        nd = C.c_uint32(0)
        _lib.check(_lib.lib().h263cu_decode_step(
            self.ctx.h, plan["parsers"], plan["packets"], plan["lens"], plan["ids"].ctypes.data, plan["n"], self.threads,
            out_flags, host_rgba, rgba_stride, plan["errs"].ctypes.data, C.byref(nd)))
        return plan["errs"]


class DeviceGroup:
    """Several GPUs driven from ONE process: one context per device, one shared pool of parser threads
    (h263cu_group_*).  Global stream s lives on device s % n_devices, slot s // n_devices; with a host
    buffer the RGBA of stream s lands at index (s % n_devices) * streams_per_device + s // n_devices."""

    def __init__(self, devices, streams_per_device, max_width, max_height, decoder_options=SORENSON_SPARK_BITSTREAM, threads=0):
        self.L = _lib.lib()
        self.devices = list(devices)
        self.streams_per_device = streams_per_device
        dv = (C.c_int * len(self.devices))(*self.devices)
        err = C.c_int(0)
        self.h = self.L.h263cu_group_create(dv, len(self.devices), streams_per_device, max_width, max_height, threads, C.byref(err))
        if not self.h:
            raise H263Error(err.value)
        self.n = streams_per_device * len(self.devices)
        self.parsers = [frontend.Parser(decoder_options) for _ in range(self.n)]
        self.ctxs = [Context(d, streams_per_device, max_width, max_height, _borrowed=self.L.h263cu_group_ctx(self.h, k))
                     for k, d in enumerate(self.devices)]

    def close(self):
        if getattr(self, "h", None):
            for c in self.ctxs:
                c.close()
            self.L.h263cu_group_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def position(self, stream):
        """Index of a global stream's picture in the host RGBA buffer (device-major)."""
        nd = len(self.devices)
        return (stream % nd) * self.streams_per_device + stream // nd

    def where(self, stream):
        """(context, local slot) of a global stream."""
        nd = len(self.devices)
        return self.ctxs[stream % nd], stream // nd

    def plan_step(self, packets, stream_ids=None):
        n = len(packets)
        ids = np.arange(n, dtype=np.uint32) if stream_ids is None else np.ascontiguousarray(stream_ids, np.uint32)
        bufs = [np.frombuffer(pk, np.uint8) if not isinstance(pk, np.ndarray) else pk for pk in packets]
        return {
            "n": n, "bufs": bufs, "ids": ids,
            "parsers": (C.c_void_p * n)(*[self.parsers[min(int(s), self.n - 1)].h for s in ids]),
            "packets": (C.c_void_p * n)(*[b.ctypes.data for b in bufs]),
            "lens": (C.c_size_t * n)(*[b.size for b in bufs]),
            "errs": np.zeros(n, np.int32),
        }

    def decode_planned(self, plan, out_flags=_lib.OUT_RGBA, host_rgba=None, rgba_stride=0):
        nd = C.c_uint32(0)
        _lib.check(self.L.h263cu_group_decode_step(self.h, plan["parsers"], plan["packets"], plan["lens"], plan["ids"].ctypes.data,
                                                   plan["n"], out_flags, host_rgba, rgba_stride, plan["errs"].ctypes.data, C.byref(nd)))
        return plan["errs"]

    def decode_step(self, packets, out_flags=_lib.OUT_RGBA, stream_ids=None, host_rgba=None, rgba_stride=0):
        return self.decode_planned(self.plan_step(packets, stream_ids), out_flags, host_rgba, rgba_stride)

    def sync(self):
        check(self.L.h263cu_group_sync(self.h))
