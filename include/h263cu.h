/*
 * h263cu.h -- C ABI of libh263cu.so, the B200-native reconstruction path for
 * ruffle-rs/h263-rs streams (H.263 / Sorenson Spark).
 *
 * The reference is a pure-Rust library without an FFI layer; the boundary it offers is
 * its public Rust API.  Each entry point below names the reference interface it stands
 * in for (paths relative to the reference repository).  A Rust `-sys` crate binds these
 * symbols 1:1; the façade crates on top keep the reference's own
 * names (`H263State::decode_next_picture`, `yuv::bt601::yuv420_to_rgba`,
 * `deblock::deblock::deblock`).
 *
 * Division of labour (BASELINE.json north_star):
 *   host  : serial bitstream / VLC parse, one parser per stream, threaded across streams
 *           (h263cu_parser_*, h263cu_parse_step).  Output = compact side info:
 *           h263cu_pic + h263cu_mb records + 16-bit run/level event units.
 *   device: everything after the parse -- inverse RLE + dequantisation, block
 *           classification, f32 IDCT (reference operation order, no FMA), motion
 *           compensation with edge clamping, residual add + clamp, optional deblocking
 *           post-filter, BT.601 YUV420 -> RGBA.  (h263cu_step_*, h263cu_submit_step)
 * There is no CPU fallback: device entry points fail with H263CU_ERR_NO_DEVICE /
 * H263CU_ERR_CUDA when no usable GPU is present.
 *
 * Threading: a context is externally synchronised (one submitting thread per context);
 * parsers are independent objects; h263cu_parse_step runs its own worker threads.
 * Ownership: the caller owns every host buffer; the context owns all device memory.
 * Errors: 0 = success, negative = failure, never aborts.
 */
#ifndef H263CU_H
#define H263CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes --------------------------------------------------------------------
 * -1..-17 are the reference's h263::Error variants in declaration order
 * (h263/src/error.rs:6-57). */
#define H263CU_OK 0
#define H263CU_ERR_INTERNAL_DECODER_ERROR (-1)
#define H263CU_ERR_MIDDLE_OF_BITSTREAM (-2)
#define H263CU_ERR_INVALID_MACROBLOCK_HEADER (-3)
#define H263CU_ERR_INVALID_MACROBLOCK_CODED_BITS (-4)
#define H263CU_ERR_INVALID_INTRA_DC (-5)
#define H263CU_ERR_INVALID_SHORT_COEFFICIENT (-6)
#define H263CU_ERR_INVALID_LONG_COEFFICIENT (-7)
#define H263CU_ERR_INVALID_MVD (-8)
#define H263CU_ERR_INVALID_PTYPE (-9)
#define H263CU_ERR_INVALID_PLUSPTYPE (-10)
#define H263CU_ERR_INVALID_GOB_HEADER (-11)
#define H263CU_ERR_INVALID_BITSTREAM (-12)
#define H263CU_ERR_PICTURE_FORMAT_MISSING (-13)
#define H263CU_ERR_PICTURE_FORMAT_INVALID (-14)
#define H263CU_ERR_UNCODED_IFRAME_BLOCKS (-15)
#define H263CU_ERR_UNHANDLED_IO_ERROR (-16) /* in-memory packets: UnexpectedEof */
#define H263CU_ERR_UNIMPLEMENTED_DECODING (-17)
/* library errors */
#define H263CU_ERR_BAD_ARGUMENT (-100)
#define H263CU_ERR_CUDA (-101)
#define H263CU_ERR_NO_DEVICE (-102)
#define H263CU_ERR_CAPACITY (-103)              /* stream / MB / event capacity exceeded */
#define H263CU_ERR_REFERENCE_WOULD_ABORT (-104) /* input on which the reference panics (abort) */
#define H263CU_ERR_NO_PICTURE (-105)            /* stream has not decoded a picture yet */
#define H263CU_ERR_OUT_OF_MEMORY (-106)

/* error.rs:66-93 classification helpers */
int h263cu_is_eof_error(int err);
int h263cu_is_macroblock_error(int err);
int h263cu_is_gob_error(int err);
const char* h263cu_strerror(int err);
int h263cu_version(void);

/* DecoderOption (h263/src/decoder/types.rs:3-18) */
#define H263CU_OPT_SORENSON_SPARK_BITSTREAM 1u
#define H263CU_OPT_USE_SCALABILITY_MODE 2u
/* EXTENSION, beyond the reference: decode Sorenson disposable P pictures (picture type 2, FLV frame type 3) like P
 * pictures that are shown but never become a reference.  The reference fails them with UnimplementedDecoding
 * (macroblock.rs:461-465), and so does this library unless the bit is set at h263cu_parser_create. */
#define H263CU_OPT_DECODE_DISPOSABLE 0x100u

/* ---- side-info format (host -> device) ------------------------------------------------
 * One time step = one picture for each of n streams.  All three arrays live in
 * caller-owned (ideally pinned) memory. */

#define H263CU_PIC_I 0
#define H263CU_PIC_P 1
#define H263CU_PIC_DISPOSABLE_P 2
#define H263CU_PIC_OTHER 3 /* PB / reserved types: only all-uncoded pictures decode */

#define H263CU_PICFLAG_DEBLOCK 1u   /* Sorenson DeblockingFlag (advisory, picture.rs:319-325) */
#define H263CU_PICFLAG_HAS_INTER 2u /* at least one MB needs the reference picture */
/* Every motion vector component of the picture lies in [-32, 31] half-pel units.  Always true for a stream decoded by
 * the front end: halfpel_decode (mvd_pred.rs:70-117) wraps into that range, and the extended range is unreachable
 * because the running options stay empty (state.rs:152-155).  When every picture of a step carries the flag the
 * device runs the reconstruction kernel without the clamped per-sample prediction path.  A producer that sets the
 * flag on a picture with longer vectors gets them clamped to the range. */
#define H263CU_PICFLAG_MV_IN_RANGE 4u
/* The picture is a disposable P picture decoded under H263CU_OPT_DECODE_DISPOSABLE: it becomes the stream's last
 * picture but not the reference of the next one (the previous reference stays in place). */
#define H263CU_PICFLAG_DISPOSABLE 8u

typedef struct h263cu_pic { /* 32 bytes */
    uint32_t stream;        /* stream slot inside the context, < max_streams */
    uint16_t width, height; /* true luma dimensions (not MB rounded) */
    uint8_t mb_w, mb_h;     /* ceil(width/16), ceil(height/16) (state.rs:173-174); <= 255 */
    uint8_t pic_type;       /* H263CU_PIC_* */
    uint8_t pquant;         /* PQUANT of the header; selects the deblock strength */
    uint8_t flags;          /* H263CU_PICFLAG_* */
    uint8_t version;        /* Sorenson version field (0xFF for baseline H.263) */
    uint16_t temporal_reference;
    uint32_t first_mb;      /* index of the picture's first record in the step's mb array */
    uint32_t n_mbs;         /* mb_w * mb_h */
    uint32_t first_event;   /* index of the picture's first unit in the step's event array */
    uint32_t n_event_units; /* 16-bit units used by this picture */
} h263cu_pic;

#define H263CU_MB_INTER 1u  /* motion compensated from the reference (MacroblockType::is_inter) */
#define H263CU_MB_WIDE 2u   /* this MB's events use the 2-unit wide form */
#define H263CU_MB_FOURMV 4u /* four motion vectors (informational; mv[] always holds 4) */
#define H263CU_MB_CODED 8u  /* COD=0 (informational) */

typedef struct h263cu_mb { /* 24 bytes, raster order inside a picture */
    uint32_t ev_off;       /* first event unit of the MB, relative to pic.first_event */
    uint16_t pic;          /* index into the step's pic array */
    uint8_t mbx, mby;      /* macroblock coordinates */
    uint8_t flags;         /* H263CU_MB_* */
    uint8_t quant;         /* in-force quantiser 1..31 (state.rs:226-227) */
    uint8_t nev[6];        /* events per block, order Y0 Y1 Y2 Y3 Cb Cr (state.rs:287-381) */
    union {
        int8_t mv[4][2];    /* INTER: decoded half-pel vectors (x, y) per luma block */
        uint8_t intradc[6]; /* INTRA: raw 8-bit INTRADC codes (0xFF => 1024, else code*8);
                               0 = block dropped (zig-zag overflow, rle.rs:125-127) */
    } u;
} h263cu_mb;

/* Event unit, narrow form (default): bits 15..10 = RUN, bits 9..0 = LEVEL as a signed
 * 10-bit integer (never 0).  Wide form (H263CU_MB_WIDE, needed when some |LEVEL| of the MB
 * exceeds the 10-bit range, i.e. Sorenson 11-bit escapes): two units per event, first =
 * RUN, second = LEVEL as int16.  The device accumulates RUN into zig-zag positions,
 * dequantises and classifies (rle.rs:82-172).  Precondition: the events of one block
 * never push the zig-zag index past 63 (the front end drops such blocks, as the
 * reference does). */
typedef uint16_t h263cu_event;

/* ---- host front end (replaces H263Reader + the serial loop of
 *      H263State::decode_next_picture, h263/src/decoder/state.rs:142-427) ---------------- */
typedef struct h263cu_parser h263cu_parser;

/* H263State::new (state.rs:42-50): one parser per stream */
h263cu_parser* h263cu_parser_create(uint32_t decoder_options);
void h263cu_parser_destroy(h263cu_parser*);
/* The DecoderOption bits the parser was created with */
uint32_t h263cu_parser_options(const h263cu_parser*);
/* Seek: discard all stream state; the next picture must be an I picture (state.rs:134-137) */
void h263cu_parser_reset(h263cu_parser*);

/* H263State::parse_picture (state.rs:102-111): header only; fills width/height/mb_w/mb_h/
 * pic_type/pquant/flags/version/temporal_reference/n_mbs of *pic. Stateless. */
int h263cu_peek_picture(uint32_t decoder_options, const uint8_t* data, size_t len, h263cu_pic* pic);

/* Serial parse of one packet = one picture (one H263Reader::from_source(&packet[..]) per
 * picture).  Writes *pic, pic->n_mbs records to mbs[0..] and the event units to events[0..];
 * pic->first_mb / first_event are set to mb_base / ev_base and every record's `pic` field
 * to pic_index, so that many pictures can be packed into one step.  Transactional like the
 * reference (state.rs:120-137): on error the parser state is unchanged. */
int h263cu_parse_picture(h263cu_parser*, const uint8_t* data, size_t len, uint32_t stream,
                         uint16_t pic_index, uint32_t mb_base, uint32_t ev_base, h263cu_pic* pic,
                         h263cu_mb* mbs, uint32_t mb_cap, h263cu_event* events, uint32_t ev_cap);

/* Threaded parse of one time step: packet i belongs to parsers[i] / stream_ids[i] and
 * becomes picture i.  Pictures that fail to parse are reported in per_pic_err[i] (may be
 * NULL) and left out of the step (the remaining pictures are packed densely; *n_pics_out
 * counts them and pic_of_input[i] (may be NULL) gives the packed index or -1).
 * `threads` <= 0 selects std::thread::hardware_concurrency(). */
int h263cu_parse_step(h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                      const uint32_t* stream_ids, uint32_t n, int threads, h263cu_pic* pics,
                      h263cu_mb* mbs, uint32_t mb_cap, h263cu_event* events, uint32_t ev_cap,
                      uint32_t* n_pics_out, uint32_t* n_mbs_out, uint32_t* n_units_out,
                      int* per_pic_err, int32_t* pic_of_input);

/* ---- test hooks: the front end's bit reader, VLC tables and block decoder as plain calls, so that the
 *      reference's own parser known-answer tests (reader.rs:448-559, macroblock.rs:551-1010, block.rs:757-2124)
 *      can be replayed on the product code.  Tables: 0 MCBPC_I, 1 MCBPC_P, 2 CBPY, 3 MVD, 4 TCOEF;
 *      out4 = {kind (0 valid, 1 stuffing, 2 invalid, 3 escape), a, b, c} as in csrc/vlc_codes.inc. ---- */
int h263cu_test_read_bits(const uint8_t* data, size_t len, size_t* bitpos, int nbits, int is_signed, int peek,
                          int64_t* value);
/* recognize_start_code (reader.rs:240-258): *skipped = stuffing bits before the code, -1 when there is none */
int h263cu_test_start_code(const uint8_t* data, size_t len, size_t bitpos, int* skipped);
int h263cu_test_read_vlc(int table, const uint8_t* data, size_t len, size_t* bitpos, int* out4);
/* decode_block (block.rs:670-755): *intradc_code = -1 when absent; run/level hold up to 64 events */
int h263cu_test_decode_block(const uint8_t* data, size_t len, size_t* bitpos, uint32_t decoder_options, int version,
                             int is_intra, int tcoef_present, int* intradc_code, int* n_events, uint8_t* run,
                             int16_t* level, int* overflow);

/* ---- device context ---------------------------------------------------------------- */
typedef struct h263cu_ctx h263cu_ctx;
typedef struct h263cu_step h263cu_step;

#define H263CU_OUT_RGBA 1u    /* produce RGBA (yuv::bt601::yuv420_to_rgba) for every picture */
#define H263CU_OUT_DEBLOCK 2u /* RGBA from deblocked planes: deblock::deblock(plane, width,
                                 QUANT_TO_STRENGTH[pquant]) per plane; the reference frames
                                 stay un-deblocked (deblock/src/lib.rs:1-2) */

int h263cu_device_count(void);
/* Device memory for max_streams streams of up to max_width x max_height: two
 * reconstruction slots (current / reference) and two RGBA slots per stream. */
h263cu_ctx* h263cu_create(int device, uint32_t max_streams, uint32_t max_width, uint32_t max_height,
                          uint32_t flags, int* err);
void h263cu_destroy(h263cu_ctx*);
int h263cu_device_of(h263cu_ctx*);

/* Pinned host memory for side info / read-back buffers (cudaHostAlloc). */
void* h263cu_alloc_pinned(size_t bytes);
void h263cu_free_pinned(void* p);

/* Copy one step's side info into device memory (cudaMemcpyAsync on the context stream) and
 * keep it there: the "inputs resident in HBM" form used for kernel-only timing.
 * Side info handed in through h263cu_step_upload / h263cu_submit_step / h263cu_submit_step_readback need not come
 * from the library's parser, so it is checked on the host first: every record must lie in its picture's range, carry
 * that picture's index, sit at its raster position (mby * mb_w + mbx == index inside the picture) and keep its events
 * inside the picture's share of the event array, else H263CU_ERR_BAD_ARGUMENT.  H263CU_PICFLAG_HAS_INTER and
 * H263CU_PICFLAG_MV_IN_RANGE are derived from the records (set / cleared as needed), not taken on trust. */
h263cu_step* h263cu_step_upload(h263cu_ctx*, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs,
                                uint32_t n_mbs, const h263cu_event* events, uint32_t n_units, int* err);
void h263cu_step_free(h263cu_ctx*, h263cu_step*);
/* Reconstruct every picture of the step (async): the recon tail of decode_next_picture
 * (state.rs:419-485: gather + 3x idct_channel + reference bookkeeping) followed by the
 * requested outputs.  Each stream's "last picture" becomes the reference of its next one
 * (state.rs:72-78). */
int h263cu_step_run(h263cu_ctx*, h263cu_step*, uint32_t out_flags);
/* n resident steps as ONE CUDA graph.  A single stream's pictures are dependent launches of about ten microseconds
 * each (BASELINE.json configs[1]); captured into a graph they cost one launch call together.  The graph holds plane
 * addresses, so it is bound to the per-stream state it was built from: h263cu_graph_launch checks that every stream
 * the steps touch is where the capture assumed (plane slot, size, picture present), launches, and advances the
 * bookkeeping as the steps would have.  A graph over an even number of pictures per stream leaves the slots where they
 * started and can be launched again at once.  The steps must outlive the graph. */
typedef struct h263cu_graph h263cu_graph;
h263cu_graph* h263cu_graph_build(h263cu_ctx*, h263cu_step* const* steps, uint32_t n_steps, uint32_t out_flags, int* err);
int h263cu_graph_launch(h263cu_ctx*, h263cu_graph*);
void h263cu_graph_free(h263cu_ctx*, h263cu_graph*);
/* upload + run through an internal double-buffered ring (the streaming form) */
int h263cu_submit_step(h263cu_ctx*, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs,
                       uint32_t n_mbs, const h263cu_event* events, uint32_t n_units, uint32_t out_flags);
/* submit_step + asynchronous read-back of the RGBA pictures of this step into
 * host_rgba (pinned; picture i at i * 4*width*height when all pictures share one size,
 * else at rgba_offsets[i]); overlaps the copy with the next step's kernels.  Call
 * h263cu_sync before reading host_rgba. */
int h263cu_submit_step_readback(h263cu_ctx*, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs,
                                uint32_t n_mbs, const h263cu_event* events, uint32_t n_units,
                                uint32_t out_flags, uint8_t* host_rgba, const uint64_t* rgba_offsets);
/* Batched H263State::decode_next_picture (state.rs:138-489) from the bitstream: packet i is the next
 * picture of parsers[i] / stream_ids[i] (NULL: stream i).  The serial VLC parse runs on `threads`
 * host threads (h263cu_parse_step) into pinned staging owned by the context, the side info goes to
 * the device through the double-buffered ring, and the step is reconstructed asynchronously -- the
 * call returns as soon as the work is queued, so the parse of the next step overlaps this step's
 * kernels and copies.  Pictures that fail to parse are reported in per_pic_err[i] (may be NULL),
 * leave their stream untouched (the reference's transactional behaviour, state.rs:120-137) and are
 * left out of the step; *n_decoded (may be NULL) counts the rest.  A picture larger than the context is such a
 * per-picture error (H263CU_ERR_CAPACITY).  A stream id out of range or named twice fails the call as a whole
 * before anything is parsed, and the parsers advance only once the device stage has accepted the step: a call
 * that returns an error has changed no parser and no stream.  When host_rgba is not NULL the
 * RGBA picture of input i is copied to host_rgba + i * rgba_stride (tight rows of 4 * width bytes)
 * on the read-back stream; call h263cu_sync before reading it. */
int h263cu_decode_step(h263cu_ctx*, h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                       const uint32_t* stream_ids, uint32_t n, int threads, uint32_t out_flags, uint8_t* host_rgba,
                       uint64_t rgba_stride, int* per_pic_err, uint32_t* n_decoded);
int h263cu_sync(h263cu_ctx*);
/* Waits for the RGBA read-back of an earlier step only: age 0 = the step submitted last with a read-back, 1 = the one
 * before it (the read-back ring is two deep).  Lets a caller consume picture t while picture t + 1 is in flight: the
 * pipelined form of decode_next_picture + yuv420_to_rgba for ONE stream (api.H263State(pipelined=True)). */
int h263cu_readback_wait(h263cu_ctx*, uint32_t age);

/* ---- one process, several GPUs (streams shard by stream, no exchange between devices) ------------------------------
 * A group owns one context per listed device and feeds them all from ONE shared pool of parser threads: the pictures
 * of device d are parsed on all threads and queued on d, then the threads move on to d + 1 while d reconstructs and
 * copies back.  Global stream s lives on device s % n_devices in slot s / n_devices.  With host_rgba != NULL the RGBA
 * of stream s lands at host_rgba + ((s % n_devices) * streams_per_device + s / n_devices) * rgba_stride (device-major:
 * each device's pictures are contiguous).  Errors and the transactional behaviour are those of h263cu_decode_step. */
typedef struct h263cu_group h263cu_group;
h263cu_group* h263cu_group_create(const int* devices, uint32_t n_devices, uint32_t streams_per_device, uint32_t max_width,
                                  uint32_t max_height, int threads, int* err);
void h263cu_group_destroy(h263cu_group*);
uint32_t h263cu_group_size(const h263cu_group*);
h263cu_ctx* h263cu_group_ctx(h263cu_group*, uint32_t index);
int h263cu_group_decode_step(h263cu_group*, h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                             const uint32_t* stream_ids, uint32_t n, uint32_t out_flags, uint8_t* host_rgba, uint64_t rgba_stride,
                             int* per_pic_err, uint32_t* n_decoded);
int h263cu_group_sync(h263cu_group*);

/* DecodedPicture accessors (h263/src/decoder/picture.rs:60-142): tight row-major planes,
 * chroma = ceil(w/2) x ceil(h/2).  Synchronise the context first. */
int h263cu_stream_info(h263cu_ctx*, uint32_t stream, uint32_t* width, uint32_t* height, uint32_t* pic_type,
                       uint32_t* pquant, uint32_t* temporal_reference);
/* width << 16 | height of the stream's last picture, 0 when it has none: the cheap form of h263cu_stream_info */
uint32_t h263cu_stream_dims(h263cu_ctx*, uint32_t stream);
int h263cu_read_yuv(h263cu_ctx*, uint32_t stream, uint8_t* y, uint8_t* cb, uint8_t* cr);
/* RGBA of the stream's last picture (4*width*height bytes, R,G,B,A); requires that the
 * last step ran with H263CU_OUT_RGBA. */
int h263cu_read_rgba(h263cu_ctx*, uint32_t stream, uint8_t* rgba);
/* Position-weighted 64-bit checksums computed on the device over the tight planes /
 * RGBA of the last picture of each listed stream:
 *   sum_i (byte[i] + 1) * ((uint32)(i * 2654435761) | 1)   (mod 2^64)
 * out = n x {y, cb, cr, rgba}. */
int h263cu_checksums(h263cu_ctx*, const uint32_t* streams, uint32_t n, uint64_t* out4);

/* CUDA-event timing on the context stream (used by bench.py). */
int h263cu_timer_start(h263cu_ctx*);
int h263cu_timer_stop(h263cu_ctx*, float* milliseconds);
/* Number of kernel launches issued by this context so far. */
uint64_t h263cu_launch_count(h263cu_ctx*);
/* How many of the reconstruction launches took the tiled kernel (the fast path: references with a
 * replicated border; any picture size) rather than the generic warp-per-macroblock kernel. */
uint64_t h263cu_tiled_launch_count(h263cu_ctx*);
/* Host time spent inside h263cu_decode_step since the last reset: the bitstream parse, and everything else (staging,
 * driver calls).  Reported beside the device time (the parse is the host's share of decode_next_picture). */
int h263cu_host_times(h263cu_ctx*, double* parse_seconds, double* other_seconds, uint64_t* calls, int reset);
/* Per-kernel timing: when enabled, every recon / deblock launch is bracketed by CUDA events
 * on the launching stream.  h263cu_profile_read synchronises, accumulates the elapsed times
 * since the last read into ms[0] (recon) / ms[1] (deblock+rgba) and the launch counts into
 * launches[0..1], and resets. */
int h263cu_profile_enable(h263cu_ctx*, int enable);
int h263cu_profile_read(h263cu_ctx*, double* ms2, uint64_t* launches2);

/* ---- stateless drop-ins for the sibling crates (host buffers in, host buffers out) ---- */
/* yuv::bt601::yuv420_to_rgba(y, chroma_b, chroma_r, y_width) -> Vec<u8>
 * (yuv/src/bt601.rs:105-196).  rgba_out holds 4*y_len bytes. y_len == 0 is a no-op. */
int h263cu_yuv420_to_rgba(const uint8_t* y, const uint8_t* chroma_b, const uint8_t* chroma_r, size_t y_len,
                          size_t y_width, uint8_t* rgba_out);
/* deblock::deblock::deblock(data, width, strength) -> Vec<u8> (deblock/src/deblock.rs:305-315) */
int h263cu_deblock(const uint8_t* data, size_t len, size_t width, uint8_t strength, uint8_t* out);
/* deblock::deblock::QUANT_TO_STRENGTH (deblock/src/deblock.rs:5-8) */
extern const uint8_t h263cu_quant_to_strength[32];

/* ---- FLV container feed (the caller side of H263Reader::from_source(&packet[..]): one reader per FLV
 *      video tag; Sorenson Spark is FLV video codec 2, one picture per tag) ------------------------- */
typedef struct h263cu_flv_packet { /* 24 bytes */
    uint64_t offset;       /* of the picture packet inside the FLV buffer (past the 1-byte video header) */
    uint32_t size;         /* bytes of the picture packet */
    uint32_t timestamp_ms; /* 32-bit tag timestamp (extended byte included) */
    uint8_t frame_type;    /* 1 key, 2 inter, 3 disposable inter */
    uint8_t codec_id;      /* always 2 (Sorenson H.263): other video codecs are skipped */
    uint16_t reserved;
    uint32_t reserved2;
} h263cu_flv_packet;

/* Zero-copy scan of an FLV byte stream: lists its H.263 video packets in file order.  Returns how many
 * there are (fill up to `cap` of them; call with cap = 0 to count) or a negative error when the buffer
 * is not FLV.  A buffer that ends inside a tag (streaming input) ends the scan at the last complete tag.
 * Audio, script, other-codec and video-info tags are skipped and counted in *n_other_tags (may be NULL). */
int64_t h263cu_flv_scan(const uint8_t* data, size_t len, h263cu_flv_packet* out, size_t cap, uint32_t* n_other_tags);
/* The matching muxer for generated streams (tests, examples): wraps n picture packets (as laid out by
 * h263cu_synth_stream) into an FLV byte stream, one video tag per picture at i * ms_per_picture;
 * frame_types may be NULL (first = key, rest = inter); filler_every > 0 adds an audio and a script tag
 * before every filler_every-th picture.  Returns the bytes needed / written, or a negative error. */
int64_t h263cu_flv_mux(const uint8_t* packets, const uint64_t* pkt_off, const uint32_t* pkt_len, const uint8_t* frame_types,
                       uint32_t n, uint32_t ms_per_picture, uint32_t filler_every, uint8_t* out, size_t cap);

/* The synthetic stream generator lives in its own library (include/h263synth.h, libh263synth.so): it is test and
 * benchmark tooling, not part of the decode path. */

#ifdef __cplusplus
}
#endif
#endif /* H263CU_H */
