// single_stream_bench.cpp -- BASELINE.json configs[1] through the C ABI without an interpreter in the loop: ONE CIF
// stream, 300 pictures, each packet handed to h263cu_decode_step (host parse + upload + kernel + RGBA read-back), the way
// a native caller (the Rust facade, bindings/h263-rs) drives it.
//   synchronous : decode_step(t); readback_wait(0); touch the RGBA       -- what H263State::decode_next_picture + RGBA costs
//   pipelined   : decode_step(t + 1) is queued before picture t is consumed (readback_wait(1)); two pinned buffers
//   parse only  : h263cu_parse_picture alone, the host share of the above
// Prints one JSON object.  Built by h263_rs_b200/build.py (g++, links libh263cu.so; the generator comes from
// libh263synth.so through dlopen).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "../include/h263cu.h"
#include "../include/h263synth.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "h263_rs_b200";
    const int device = argc > 2 ? atoi(argv[2]) : 0;
    void* so = dlopen((dir + "/libh263synth.so").c_str(), RTLD_NOW);
    if (!so) return fprintf(stderr, "dlopen: %s\n", dlerror()), 1;
    auto defaults = (void (*)(h263cu_synth_params*, uint32_t, uint32_t, uint32_t, uint64_t))dlsym(so, "h263cu_synth_default_params");
    auto gen = (int64_t(*)(const h263cu_synth_params*, uint8_t*, size_t, uint64_t*, uint32_t*))dlsym(so, "h263cu_synth_stream");
    const int W = 352, H = 288, N = 300, WARM = 20;
    h263cu_synth_params sp;
    defaults(&sp, W, H, N, 2);
    sp.mv_mode = 1;
    int64_t need = gen(&sp, nullptr, 0, nullptr, nullptr);
    std::vector<uint8_t> blob((size_t)need);
    std::vector<uint64_t> off(N);
    std::vector<uint32_t> len(N);
    gen(&sp, blob.data(), blob.size(), off.data(), len.data());

    int err = 0;
    h263cu_ctx* ctx = h263cu_create(device, 1, W, H, 0, &err);
    if (!ctx) return fprintf(stderr, "h263cu_create: %s\n", h263cu_strerror(err)), 1;
    const size_t pic_bytes = (size_t)W * H * 4;
    uint8_t* host[2] = {(uint8_t*)h263cu_alloc_pinned(pic_bytes), (uint8_t*)h263cu_alloc_pinned(pic_bytes)};
    volatile uint64_t sink = 0;
    auto step = [&](h263cu_parser* ps, int t) {
        const uint8_t* pk = blob.data() + off[t];
        size_t l = len[t];
        uint32_t id = 0, nd = 0;
        int perr = 0;
        int e = h263cu_decode_step(ctx, &ps, &pk, &l, &id, 1, 1, H263CU_OUT_RGBA, host[t & 1], 0, &perr, &nd);
        if (e || perr) {
            fprintf(stderr, "decode_step picture %d: %d / %d\n", t, e, perr);
            exit(1);
        }
    };
    auto touch = [&](int t) { sink += host[t & 1][0] + host[t & 1][pic_bytes - 1] + host[t & 1][pic_bytes / 2]; };

    // synchronous
    h263cu_parser* ps = h263cu_parser_create(H263CU_OPT_SORENSON_SPARK_BITSTREAM);
    for (int t = 0; t < WARM; t++) step(ps, t), h263cu_readback_wait(ctx, 0);
    double t0 = now();
    for (int t = WARM; t < N; t++) {
        step(ps, t);
        h263cu_readback_wait(ctx, 0);
        touch(t);
    }
    const double sync_s = now() - t0;
    h263cu_sync(ctx);
    std::vector<uint8_t> last_sync(host[(N - 1) & 1], host[(N - 1) & 1] + pic_bytes);
    h263cu_parser_destroy(ps);

    // pipelined: the same stream again from its I picture
    ps = h263cu_parser_create(H263CU_OPT_SORENSON_SPARK_BITSTREAM);
    for (int t = 0; t < WARM; t++) step(ps, t);
    h263cu_readback_wait(ctx, 0);
    h263cu_host_times(ctx, nullptr, nullptr, nullptr, 1);
    t0 = now();
    for (int t = WARM; t < N; t++) {
        step(ps, t);                  // queued: parse of t done, device work of t in flight
        h263cu_readback_wait(ctx, 1);  // picture t - 1 has arrived
        touch(t - 1);
    }
    h263cu_readback_wait(ctx, 0);
    touch(N - 1);
    const double pipe_s = now() - t0;
    double hp = 0, ho = 0;
    uint64_t hc = 0;
    h263cu_host_times(ctx, &hp, &ho, &hc, 1);
    const bool same = memcmp(last_sync.data(), host[(N - 1) & 1], pic_bytes) == 0;
    h263cu_parser_destroy(ps);

    // parse alone
    ps = h263cu_parser_create(H263CU_OPT_SORENSON_SPARK_BITSTREAM);
    std::vector<h263cu_mb> mbs(396);
    std::vector<h263cu_event> ev(1 << 18);
    h263cu_pic pic;
    for (int t = 0; t < WARM; t++) h263cu_parse_picture(ps, blob.data() + off[t], len[t], 0, 0, 0, 0, &pic, mbs.data(), 396, ev.data(), (uint32_t)ev.size());
    t0 = now();
    for (int t = WARM; t < N; t++) h263cu_parse_picture(ps, blob.data() + off[t], len[t], 0, 0, 0, 0, &pic, mbs.data(), 396, ev.data(), (uint32_t)ev.size());
    const double parse_s = now() - t0;
    h263cu_parser_destroy(ps);

    const int n = N - WARM;
    printf("{\"pictures\": %d, \"synchronous_us_per_picture\": %.2f, \"synchronous_frames_per_s\": %.1f, \"pipelined_us_per_picture\": %.2f, "
           "\"pipelined_frames_per_s\": %.1f, \"parse_only_us_per_picture\": %.2f, \"pipelined_last_picture_matches_synchronous\": %s, "
           "\"pipelined_host_parse_us_per_picture\": %.2f, \"pipelined_host_other_us_per_picture\": %.2f, \"launches\": %llu}\n",
           n, sync_s / n * 1e6, n / sync_s, pipe_s / n * 1e6, n / pipe_s, parse_s / n * 1e6, same ? "true" : "false", hp / (double)hc * 1e6, ho / (double)hc * 1e6,
           (unsigned long long)h263cu_launch_count(ctx));
    h263cu_free_pinned(host[0]);
    h263cu_free_pinned(host[1]);
    h263cu_destroy(ctx);
    return (int)(sink & 0);
}
