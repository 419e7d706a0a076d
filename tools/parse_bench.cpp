// parse_bench.cpp -- single-thread timing of the host front end on generated CIF streams (no GPU, no CUDA):
//   g++ -O2 -std=c++17 -pthread -o /tmp/parse_bench tools/parse_bench.cpp h263_rs_b200/csrc/frontend.cpp -ldl
// The generator is loaded from libh263synth.so (dlopen) so that this links the front end alone.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>

#include "../include/h263cu.h"
#include "../include/h263synth.h"

int main(int argc, char** argv) {
    const int n_streams = argc > 1 ? atoi(argv[1]) : 8, n_pics = argc > 2 ? atoi(argv[2]) : 30, reps = argc > 3 ? atoi(argv[3]) : 20;
    const int mean_events_x10 = argc > 4 ? atoi(argv[4]) : 0, pct_cbp = argc > 5 ? atoi(argv[5]) : -1;
    void* so = dlopen("h263_rs_b200/libh263synth.so", RTLD_NOW);
    if (!so) return fprintf(stderr, "dlopen: %s\n", dlerror()), 1;
    auto defaults = (void (*)(h263cu_synth_params*, uint32_t, uint32_t, uint32_t, uint64_t))dlsym(so, "h263cu_synth_default_params");
    auto gen = (int64_t(*)(const h263cu_synth_params*, uint8_t*, size_t, uint64_t*, uint32_t*))dlsym(so, "h263cu_synth_stream");
    std::vector<std::vector<uint8_t>> blobs(n_streams);
    std::vector<std::vector<uint64_t>> offs(n_streams);
    std::vector<std::vector<uint32_t>> lens(n_streams);
    size_t bytes = 0;
    for (int s = 0; s < n_streams; s++) {
        h263cu_synth_params p;
        defaults(&p, 352, 288, n_pics, 1000 + s);
        p.mv_mode = s % 4 ? 0 : 1;
        if (mean_events_x10) p.mean_events_x10 = mean_events_x10;
        if (pct_cbp >= 0) p.pct_cbp_inter = p.pct_cbp_intra = pct_cbp;
        int64_t need = gen(&p, nullptr, 0, nullptr, nullptr);
        blobs[s].resize(need), offs[s].resize(n_pics), lens[s].resize(n_pics);
        gen(&p, blobs[s].data(), need, offs[s].data(), lens[s].data());
        bytes += need;
    }
    std::vector<h263cu_mb> mbs(396);
    std::vector<h263cu_event> ev(1 << 18);
    double best = 1e9;
    uint64_t units = 0;
    for (int r = 0; r < reps; r++) {
        units = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < n_streams; s++) {
            h263cu_parser* ps = h263cu_parser_create(1);
            for (int t = 0; t < n_pics; t++) {
                h263cu_pic pic;
                int e = h263cu_parse_picture(ps, blobs[s].data() + offs[s][t], lens[s][t], 0, 0, 0, 0, &pic, mbs.data(), 396, ev.data(), (uint32_t)ev.size());
                if (e) return fprintf(stderr, "parse error %d\n", e), 1;
                units += pic.n_event_units;
            }
            h263cu_parser_destroy(ps);
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
    }
    const double pics = (double)n_streams * n_pics;
    printf("%.1f us per CIF picture, %.1f ns per event unit, %.0f MB/s of bitstream (%.0f bytes, %.0f units per picture)\n", best / pics * 1e6,
           best / units * 1e9, bytes / best / 1e6, bytes / pics, units / pics);
    return 0;
}
