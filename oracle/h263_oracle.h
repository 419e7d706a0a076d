/*
 * h263_oracle.h -- C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a literal CPU restatement of the
 * reference decoder (ruffle-rs/h263-rs) and exists to CHECK the CUDA product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  Nothing under h263_rs_b200/ links,
 * includes or calls anything in oracle/.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - yuv420_to_rgba, deblock, bit reader, VLC tables, decode_block: PINNED by the
 *     reference's own unit-test vectors (the JSON fixtures under tests/golden/, extracted by
 *     tests/golden/extract_reference_kats.py).
 *   - inverse_rle, idct, gather, mv prediction, decode_next_picture: the reference
 *     has no tests or fixtures for these and cannot be built here (no Rust
 *     toolchain) => "parity unpinned" by the reference; mitigated by an
 *     independent numpy-f32 restatement (tests/np_restatement.py).
 *
 * Error codes are the reference's `h263::Error` variants (error.rs:6-57) numbered
 * 1..17 in declaration order; 0 = Ok.
 */
#ifndef H263_ORACLE_H
#define H263_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_OK = 0,
    ORC_ERR_INTERNAL_DECODER_ERROR = 1,
    ORC_ERR_MIDDLE_OF_BITSTREAM = 2,
    ORC_ERR_INVALID_MACROBLOCK_HEADER = 3,
    ORC_ERR_INVALID_MACROBLOCK_CODED_BITS = 4,
    ORC_ERR_INVALID_INTRA_DC = 5,
    ORC_ERR_INVALID_SHORT_COEFFICIENT = 6,
    ORC_ERR_INVALID_LONG_COEFFICIENT = 7,
    ORC_ERR_INVALID_MVD = 8,
    ORC_ERR_INVALID_PTYPE = 9,
    ORC_ERR_INVALID_PLUSPTYPE = 10,
    ORC_ERR_INVALID_GOB_HEADER = 11,
    ORC_ERR_INVALID_BITSTREAM = 12,
    ORC_ERR_PICTURE_FORMAT_MISSING = 13,
    ORC_ERR_PICTURE_FORMAT_INVALID = 14,
    ORC_ERR_UNCODED_IFRAME_BLOCKS = 15,
    ORC_ERR_UNHANDLED_IO_ERROR = 16, /* in-memory source: always UnexpectedEof */
    ORC_ERR_UNIMPLEMENTED_DECODING = 17,
    /* Situations in which the reference process would abort (panic="abort",
     * Cargo.toml:14-18) or index out of bounds; the oracle reports them instead. */
    ORC_ERR_REFERENCE_WOULD_ABORT = 100
};

/* DecoderOption bitflags (decoder/types.rs:3-18) */
#define ORC_OPT_SORENSON_SPARK_BITSTREAM 1
#define ORC_OPT_USE_SCALABILITY_MODE 2
/* EXTENSION, not reference behaviour: decode Sorenson disposable P pictures (type code 2) like P pictures that
 * never become a reference.  The reference fails them with UnimplementedDecoding (macroblock.rs:461-465) and, for
 * the all-uncoded ones it does accept, predicts the next picture from them (get_reference_picture returns the LAST
 * picture, state.rs:72-78).  With this bit the oracle predicts from the last NON-disposable picture instead; it is
 * the checker of the product's H263CU_OPT_DECODE_DISPOSABLE and is pinned by no reference vector. */
#define ORC_OPT_DECODE_DISPOSABLE 0x100

/* PictureTypeCode as reported by orc_last_picture_info */
#define ORC_PIC_I 0
#define ORC_PIC_P 1
#define ORC_PIC_DISPOSABLE_P 2
#define ORC_PIC_OTHER 3

typedef struct orc_state orc_state;

/* ---- H263State (decoder/state.rs) ---- */
orc_state* orc_state_new(int decoder_options);
void orc_state_free(orc_state*);
/* One packet == one H263Reader::from_source(&packet[..]) + decode_next_picture. */
int orc_decode_next_picture(orc_state*, const uint8_t* data, size_t len);
/* get_last_picture(): returns 0 if a picture exists, else -1. */
int orc_last_picture_info(orc_state*, int* width, int* height, int* temporal_reference,
                          int* picture_type, int* quantizer, int* deblock_flag, int* version);
int orc_last_picture_yuv(orc_state*, uint8_t* y, uint8_t* cb, uint8_t* cr);

/* Optional parse trace of the last successful decode (for parser cross-checks).
 * mb_type uses the numbering of vlc_codes.inc (0 Inter..5 Inter4Vq); uncoded MBs and
 * padded MBs are reported as Inter with coded=0. */
void orc_state_set_trace(orc_state*, int enable);
int orc_trace_counts(orc_state*, int* n_mbs, int* n_events);
/* Arrays sized from orc_trace_counts: mb_type[n], coded[n], quant[n], mv[n*8] (x,y per
 * block), intradc[n*6] (-1 = none, else the 8-bit code), nev[n*6], run[e], level[e]. */
int orc_trace_copy(orc_state*, int8_t* mb_type, int8_t* coded, uint8_t* quant, int16_t* mv,
                   int16_t* intradc, uint8_t* nev, uint8_t* run, int16_t* level);

/* ---- sibling crates ---- */
/* yuv::bt601::yuv420_to_rgba (yuv/src/bt601.rs:105-196). out holds 4*y_len bytes. */
void orc_yuv420_to_rgba(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, size_t y_len,
                        size_t y_width, uint8_t* out);
/* deblock::deblock::deblock (deblock/src/deblock.rs:305-315). */
void orc_deblock(const uint8_t* in, size_t len, size_t width, int strength, uint8_t* out);
/* scalar `process` (deblock.rs:29-42) when simd==0, `process_simd` lane semantics
 * (deblock.rs:99-127) when simd==1. abcd updated in place. */
void orc_deblock_process(uint8_t* abcd, int strength, int simd);
int orc_quant_to_strength(int quant);

/* ---- hot-path pieces, exposed for unit tests ---- */
/* inverse_rle (rle.rs:82-172). intradc_code < 0 => None. Returns class 0 Zero,1 Dc,
 * 2 Horiz, 3 Vert, 4 Full; coefs[64] row-major [y][x] (Dc: coefs[0]; Horiz: coefs[0..8]
 * = first row; Vert: coefs[8*k] = first column). */
int orc_inverse_rle(int intradc_code, int n_events, const uint8_t* run, const int16_t* level,
                    int quant, float* coefs);
/* idct_channel on a single 8x8 block (idct.rs:82-201); pixels is an 8x8 row-major u8
 * block that already holds the motion-compensated prediction. */
void orc_idct_block(int cls, const float* coefs, uint8_t* pixels);
/* idct_1d (idct.rs:52-65) */
void orc_idct_1d(const float* in, float* out);
/* gather_block (gather.rs:47-126): src plane (width x height) -> dst plane, same dims. */
void orc_gather_block(const uint8_t* src, int width, int height, int pos_x, int pos_y, int mv_x,
                      int mv_y, uint8_t* dst);
/* average_sum_of_mvs on one component (types.rs:759-768) */
int orc_average_sum_of_mvs(int sum);
/* halfpel_decode without UMV (mvd_pred.rs:70-117) and median (types.rs:772-798) */
int orc_halfpel_decode(int predictor, int mvd);
int orc_median_of(int a, int b, int c);

/* ---- parser pieces for the reference's known-answer tests ---- */
/* Tables: 0 MCBPC_I, 1 MCBPC_P, 2 CBPY, 3 MVD, 4 TCOEF. Reads one code starting at
 * *bitpos; returns 0 or an error; out = {kind, a, b, c} (see vlc_codes.inc). */
int orc_read_vlc(int table, const uint8_t* data, size_t len, size_t* bitpos, int* out4);
int orc_read_bits(const uint8_t* data, size_t len, size_t* bitpos, int nbits, int is_signed,
                  int peek, int64_t* value);
/* recognize_start_code (reader.rs:240-258): returns error or 0; *skipped = -1 for None. */
int orc_recognize_start_code(const uint8_t* data, size_t len, size_t bitpos, int in_error,
                             int* skipped);
/* decode_block (block.rs:670-755). Outputs intradc code (-1 none), events. */
int orc_decode_block(const uint8_t* data, size_t len, size_t* bitpos, int decoder_options,
                     int version, int is_intra, int tcoef_present, int* intradc_code,
                     int* n_events, uint8_t* run, int16_t* level, uint8_t* is_short, int cap);

/* ---- CPU baseline driver (bench.py cpu_baseline / --impl reference) ----
 * Decodes n_streams independent streams, one H263State per stream, `threads` worker
 * threads each taking whole streams (the reference's natural host model: one state per
 * stream per thread). Stream s owns pictures [pic_first[s], pic_first[s+1]) of the
 * packet table (pkt_off[i], pkt_len[i]) into `blob`. Every picture is decoded, optionally
 * deblocked with QUANT_TO_STRENGTH[PQUANT], and converted to RGBA.
 * Returns wall seconds (<0 on decode error); *pixels = luma pixels produced;
 * *checksum = position-weighted checksum of all RGBA bytes (order independent). */
double orc_bench_decode(const uint8_t* blob, const uint64_t* pkt_off, const uint32_t* pkt_len,
                        const uint32_t* pic_first, int n_streams, int decoder_options,
                        int do_deblock, int threads, uint64_t* pixels, uint64_t* checksum);

/* Step-wise form of the same baseline: n_streams persistent H263States; one call decodes
 * ONE picture for every stream (packet s = blob + pkt_off[s], pkt_len[s]) on `threads`
 * worker threads, deblocks (optional) and converts to RGBA.  Returns wall seconds (<0 on
 * error); adds the luma pixels produced to *pixels.  The worker threads live as long as the batch and every stream
 * keeps its RGBA buffer; when checksum is not NULL the RGBA checksums are added to *checksum AFTER the timed region
 * (the seconds returned cover parse + reconstruction + [deblock] + yuv420_to_rgba only). */
typedef struct orc_batch orc_batch;
orc_batch* orc_batch_new(int n_streams, int decoder_options);
void orc_batch_free(orc_batch*);
double orc_batch_step(orc_batch*, const uint8_t* blob, const uint64_t* pkt_off, const uint32_t* pkt_len,
                      int do_deblock, int threads, uint64_t* pixels, uint64_t* checksum);
/* planes / RGBA of stream s after the last step (for cross-checks) */
int orc_batch_stream_yuv(orc_batch*, int s, uint8_t* y, uint8_t* cb, uint8_t* cr);

#ifdef __cplusplus
}
#endif
#endif
