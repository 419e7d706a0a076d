/* abi_smoke.c -- the C ABI of libh263cu.so exercised from C (not through ctypes): the binding a C or Rust caller makes.
 *
 *   gcc -O1 -Wall -Wextra -Werror -I include -o /tmp/abi_smoke tests/native/abi_smoke.c -L h263_rs_b200 -lh263cu \
 *       -Wl,-rpath,$PWD/h263_rs_b200 -ldl
 *   /tmp/abi_smoke h263_rs_b200/libh263synth.so
 *
 * Host part (always): error classification, header peek, serial parse of a generated QCIF stream with the
 * transactional error behaviour, the parser test hooks.  Device part (when a GPU is present, else it checks that the
 * device entry points refuse loudly): decode the stream through h263cu_decode_step, read planes and RGBA back, check
 * the RGBA against the stateless h263cu_yuv420_to_rgba of the planes.  Prints "abi_smoke ok ..." and returns 0. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "h263cu.h"
#include "h263synth.h"

#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

typedef void (*defaults_fn)(h263cu_synth_params*, uint32_t, uint32_t, uint32_t, uint64_t);
typedef int64_t (*stream_fn)(const h263cu_synth_params*, uint8_t*, size_t, uint64_t*, uint32_t*);

int main(int argc, char** argv) {
    const char* synth_path = argc > 1 ? argv[1] : "h263_rs_b200/libh263synth.so";
    void* so = dlopen(synth_path, RTLD_NOW);
    CHECK(so != NULL);
    defaults_fn defaults = (defaults_fn)dlsym(so, "h263cu_synth_default_params");
    stream_fn gen = (stream_fn)dlsym(so, "h263cu_synth_stream");
    CHECK(defaults && gen);

    /* ---- a QCIF stream: 1 I + 3 P pictures ---- */
    enum { W = 176, H = 144, N = 4, MBS = 11 * 9 };
    h263cu_synth_params sp;
    defaults(&sp, W, H, N, 7);
    sp.mv_mode = 1;
    int64_t need = gen(&sp, NULL, 0, NULL, NULL);
    CHECK(need > 0);
    uint8_t* blob = (uint8_t*)malloc((size_t)need);
    uint64_t off[N];
    uint32_t len[N];
    CHECK(gen(&sp, blob, (size_t)need, off, len) == need);

    /* ---- host side ---- */
    CHECK(h263cu_version() >= 100);
    CHECK(h263cu_is_eof_error(H263CU_ERR_UNHANDLED_IO_ERROR) && !h263cu_is_eof_error(H263CU_ERR_INVALID_MVD));
    CHECK(h263cu_is_macroblock_error(H263CU_ERR_INVALID_MACROBLOCK_HEADER) && h263cu_is_gob_error(H263CU_ERR_INVALID_GOB_HEADER));
    CHECK(strlen(h263cu_strerror(H263CU_ERR_NO_DEVICE)) > 0);
    CHECK(h263cu_quant_to_strength[1] == 1 && h263cu_quant_to_strength[31] == 12);

    h263cu_pic hdr;
    CHECK(h263cu_peek_picture(H263CU_OPT_SORENSON_SPARK_BITSTREAM, blob + off[1], len[1], &hdr) == 0);
    CHECK(hdr.width == W && hdr.height == H && hdr.mb_w == 11 && hdr.mb_h == 9 && hdr.n_mbs == MBS && hdr.pic_type == H263CU_PIC_P);

    h263cu_parser* ps = h263cu_parser_create(H263CU_OPT_SORENSON_SPARK_BITSTREAM);
    CHECK(ps && h263cu_parser_options(ps) == H263CU_OPT_SORENSON_SPARK_BITSTREAM);
    static h263cu_mb mbs[MBS];
    static h263cu_event ev[1 << 16];
    h263cu_pic pic;
    /* a P picture first: no reference yet -> UncodedIFrameBlocks (gather.rs:149), and the parser has not moved */
    CHECK(h263cu_parse_picture(ps, blob + off[1], len[1], 0, 0, 0, 0, &pic, mbs, MBS, ev, 1 << 16) == H263CU_ERR_UNCODED_IFRAME_BLOCKS);
    uint32_t total_units = 0;
    for (int t = 0; t < N; t++) {
        CHECK(h263cu_parse_picture(ps, blob + off[t], len[t], 0, 0, 0, 0, &pic, mbs, MBS, ev, 1 << 16) == 0);
        CHECK(pic.n_mbs == MBS && pic.width == W && (pic.flags & H263CU_PICFLAG_MV_IN_RANGE));
        CHECK((t == 0) == !(pic.flags & H263CU_PICFLAG_HAS_INTER));
        for (uint32_t k = 0; k < MBS; k++) CHECK(mbs[k].mbx == k % 11 && mbs[k].mby == k / 11 && mbs[k].quant >= 1 && mbs[k].quant <= 31);
        total_units += pic.n_event_units;
    }
    CHECK(total_units > 0);
    CHECK(h263cu_parse_picture(ps, blob, 2, 0, 0, 0, 0, &pic, mbs, MBS, ev, 1 << 16) == H263CU_ERR_UNHANDLED_IO_ERROR);
    h263cu_parser_reset(ps);

    /* the parser hooks: the first 17 bits of a packet are the picture start code, 0x00 0x00 0x8x */
    size_t bitpos = 0;
    int64_t v = 0;
    CHECK(h263cu_test_read_bits(blob + off[0], len[0], &bitpos, 17, 0, 0, &v) == 0 && v == 1 && bitpos == 17);
    int skipped = -1;
    CHECK(h263cu_test_start_code(blob + off[0], len[0], 0, &skipped) == 0 && skipped == 0);

    /* ---- device side ---- */
    int err = 0;
    if (h263cu_device_count() == 0) {
        CHECK(h263cu_create(0, 1, W, H, 0, &err) == NULL && err == H263CU_ERR_NO_DEVICE);
        uint8_t y4[4] = {16, 16, 16, 16}, c1 = 128, out[16];
        CHECK(h263cu_yuv420_to_rgba(y4, &c1, &c1, 4, 2, out) == H263CU_ERR_NO_DEVICE);
        printf("abi_smoke ok (host part; no CUDA device: device entry points refuse with H263CU_ERR_NO_DEVICE)\n");
        return 0;
    }
    h263cu_ctx* ctx = h263cu_create(0, 1, W, H, 0, &err);
    CHECK(ctx && err == 0);
    uint8_t* rgba = (uint8_t*)h263cu_alloc_pinned((size_t)W * H * 4);
    CHECK(rgba != NULL);
    static uint8_t y[W * H], cb[W * H / 4], cr[W * H / 4], rgba2[W * H * 4];
    for (int t = 0; t < N; t++) {
        const uint8_t* packet = blob + off[t];
        size_t plen = len[t];
        uint32_t id = 0, n_decoded = 0;
        int perr = 0;
        CHECK(h263cu_decode_step(ctx, &ps, &packet, &plen, &id, 1, 1, H263CU_OUT_RGBA, rgba, (uint64_t)W * H * 4, &perr, &n_decoded) == 0);
        CHECK(perr == 0 && n_decoded == 1);
        CHECK(h263cu_readback_wait(ctx, 0) == 0);
        CHECK(h263cu_read_yuv(ctx, 0, y, cb, cr) == 0);
        /* the fused RGBA equals the stateless conversion of the planes (bt601.rs:105-196) */
        CHECK(h263cu_yuv420_to_rgba(y, cb, cr, (size_t)W * H, W, rgba2) == 0);
        CHECK(memcmp(rgba, rgba2, (size_t)W * H * 4) == 0);
    }
    uint32_t w = 0, h = 0, type = 0, q = 0, tr = 0;
    CHECK(h263cu_stream_info(ctx, 0, &w, &h, &type, &q, &tr) == 0 && w == W && h == H && type == H263CU_PIC_P);
    CHECK(h263cu_launch_count(ctx) == N && h263cu_tiled_launch_count(ctx) == N);
    h263cu_free_pinned(rgba);
    h263cu_destroy(ctx);
    h263cu_parser_destroy(ps);
    free(blob);
    printf("abi_smoke ok (host + device: %d pictures decoded, fused RGBA == yuv420_to_rgba(planes))\n", N);
    return 0;
}
