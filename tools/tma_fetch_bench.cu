// tma_fetch_bench.cu -- the prediction fetch of the recon kernel in isolation, two ways:
//   A  what recon_tile_kernel does: every lane loads the aligned words of its own rows (4-5 luma rows x 2-3 words,
//      2-3 chroma rows x 2-3 words, predicated on the half-pel bits like the kernel) straight from global memory and
//      funnel-shifts them into place;
//   B  what the north-star sketch suggests: the reference window of every macroblock is staged in shared memory by
//      TMA (cp.async.bulk.tensor.2d, completion on a per-warp mbarrier) and the lanes read it from there.
// Finding 1 (this file's first version): a box whose first column is not a multiple of 16 bytes raises "illegal
// instruction" at UTMALDG (compute-sanitizer, profiles/r01_tma_fetch.txt) -- TMA does not align byte-granular
// windows.  So the boxes start at the column rounded down to 16: 48 x 17 for luma, 32 x 9 per chroma plane, the lanes
// still funnel-shift, and a macroblock needs 896 + 2 x 384 = 1 664 bytes of staging (6.5 KB per warp of four).
// Same macroblock grid as the benchmark (1024 CIF pictures, padded planes, one warp per 4 macroblocks, 4 warps per
// CTA), vectors drawn like the benchmark's generator (synth.cpp mv_mode 0).  Each lane reduces what it fetched to one
// word; the two variants must produce the same words.  Shared memory per CTA is padded so that A runs at the real
// kernel's 8 CTAs per SM and B at a chosen residency (second argument; the real kernel's 24 KB plus the staging area
// leave 3 CTAs per SM, 4 if the staging aliases the buffers that are dead by then).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_fetch_bench tools/tma_fetch_bench.cu
//   /tmp/tma_fetch_bench [ctas_per_sm_A=8] [ctas_per_sm_B=4]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int STREAMS = 1024, MBW = 22, MBH = 18;
constexpr int PITCH_Y = 416, ROWS_Y = 320, PITCH_C = 208, ROWS_C = 160;  // padded CIF planes (context.cu)
constexpr int PAD_Y_COLS = 32, PAD_Y_ROWS = 16, PAD_C_COLS = 16, PAD_C_ROWS = 8;
constexpr int CTA_WARPS = 4, WARP_MBS = 4;

struct Rec {  // one macroblock: top-left source sample of its luma / chroma window in pool coordinates
    uint32_t yrow, ycol, crow, ccol;
    uint32_t hp;  // half-pel bits: luma x, y (bits 0, 1), chroma x, y (bits 2, 3)
    uint32_t pad[3];
};

__device__ __forceinline__ uint32_t fold_row(uint32_t w0, uint32_t w1, uint32_t w2, int sh) {
    const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
    return a0 ^ a1 ^ ((w2 >> sh) & 0xFFu);
}

// ---- A: direct aligned loads, the lane mapping of recon_tile_kernel phase 3 -------------------------------------
__global__ void __launch_bounds__(CTA_WARPS * 32) fetch_ldg(const Rec* __restrict__ recs, const uint8_t* __restrict__ y,
                                                           const uint8_t* __restrict__ cb, const uint8_t* __restrict__ cr,
                                                           uint32_t n_mbs, uint32_t* __restrict__ out) {
    extern __shared__ uint8_t pad[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * CTA_WARPS + warp;
    const int mbq = lane >> 3, rg = (lane >> 1) & 3, h = lane & 1;
    const uint32_t mb = tile * WARP_MBS + mbq;
    if (mb >= n_mbs) return;
    if (threadIdx.x == 9999) pad[0] = 1;
    const Rec r = recs[mb];
    uint32_t acc = 0;
    {
        const uint32_t col = r.ycol + 8 * h, a = col & 3;
        const bool third = a != 0 || (r.hp & 1u), extra = (r.hp & 2u) != 0;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(y + (size_t)(r.yrow + rg * 4) * PITCH_Y + (col - a));
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint32_t* q = p + k * (PITCH_Y / 4);
            if (k < 4 || extra) acc ^= fold_row(__ldg(q), __ldg(q + 1), third ? __ldg(q + 2) : 0u, a * 8);
        }
    }
    {
        const uint32_t a = r.ccol & 3;
        const bool third = a != 0 || (r.hp & 4u), extra = (r.hp & 8u) != 0;
        const uint32_t* p = reinterpret_cast<const uint32_t*>((h ? cr : cb) + (size_t)(r.crow + rg * 2) * PITCH_C + (r.ccol - a));
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t* q = p + k * (PITCH_C / 4);
            if (k < 2 || extra) acc ^= fold_row(__ldg(q), __ldg(q + 1), third ? __ldg(q + 2) : 0u, a * 8);
        }
    }
    out[(size_t)tile * 32 + lane] = acc;
}

// ---- C: direct loads with another division of the macroblock: a lane owns 4 luma columns x 8 rows and the 2 x 4
// chroma samples of BOTH planes under them (4 lanes side by side in a row): fewer distinct sectors per load pass
// (8 rows per pass instead of 16 for luma, 8 instead of 32 for chroma), more passes.  Throughput probe only (its
// reduction is over a different partition of the same bytes, so its words are not comparable with A's).
__global__ void __launch_bounds__(CTA_WARPS * 32) fetch_ldg_4x8(const Rec* __restrict__ recs, const uint8_t* __restrict__ y,
                                                               const uint8_t* __restrict__ cb, const uint8_t* __restrict__ cr,
                                                               uint32_t n_mbs, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * CTA_WARPS + warp;
    const int mbq = lane >> 3, half = (lane >> 2) & 1, q = lane & 3;
    const uint32_t mb = tile * WARP_MBS + mbq;
    if (mb >= n_mbs) return;
    const Rec r = recs[mb];
    uint32_t acc = 0;
    {
        const uint32_t col = r.ycol + 4 * q, a = col & 3;
        const bool second = a != 0 || (r.hp & 1u), extra = (r.hp & 2u) != 0;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(y + (size_t)(r.yrow + half * 8) * PITCH_Y + (col - a));
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const uint32_t* q0 = p + k * (PITCH_Y / 4);
            if (k < 8 || extra) {
                const uint32_t w0 = __ldg(q0), w1 = second ? __ldg(q0 + 1) : 0u;
                acc ^= __funnelshift_r(w0, w1, a * 8) ^ ((w1 >> (a * 8)) & 0xFFu);
            }
        }
    }
    {
        const uint32_t col = r.ccol + 2 * q, a = col & 3;
        const bool second = a + 2 + ((r.hp >> 2) & 1u) > 4, extra = (r.hp & 8u) != 0;
        const size_t off = (size_t)(r.crow + half * 4) * PITCH_C + (col - a);
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(cb + off);
        const uint32_t* pr = reinterpret_cast<const uint32_t*>(cr + off);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            if (k < 4 || extra) {
                const uint32_t b0 = __ldg(pb + k * (PITCH_C / 4)), r0 = __ldg(pr + k * (PITCH_C / 4));
                const uint32_t b1 = second ? __ldg(pb + k * (PITCH_C / 4) + 1) : 0u, r1 = second ? __ldg(pr + k * (PITCH_C / 4) + 1) : 0u;
                acc ^= (__funnelshift_r(b0, b1, a * 8) & 0xFFFFFFu) ^ (__funnelshift_r(r0, r1, a * 8) & 0xFFFFFFu);
            }
        }
    }
    out[(size_t)tile * 32 + lane] = acc;
}

// ---- B: TMA boxes into shared memory ----------------------------------------------------------------------------
constexpr int LUMA_BOX_W = 48, LUMA_BOX_H = 17, CHROMA_BOX_W = 32, CHROMA_BOX_H = 9;
constexpr int LUMA_BYTES = LUMA_BOX_W * LUMA_BOX_H;        // 816
constexpr int CHROMA_BYTES = CHROMA_BOX_W * CHROMA_BOX_H;  // 288
constexpr int LUMA_SLOT = 896, CHROMA_SLOT = 384;          // 128-byte aligned slots (cp.async.bulk.tensor destination alignment)
constexpr int MB_STAGE = LUMA_SLOT + 2 * CHROMA_SLOT;
constexpr int WARP_STAGE = WARP_MBS * MB_STAGE;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(CTA_WARPS * 32) fetch_tma(const Rec* __restrict__ recs, const __grid_constant__ CUtensorMap tm_y,
                                                           const __grid_constant__ CUtensorMap tm_cb,
                                                           const __grid_constant__ CUtensorMap tm_cr, uint32_t n_mbs,
                                                           uint32_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[CTA_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * CTA_WARPS + warp;
    const uint32_t mb0 = tile * WARP_MBS;
    if (mb0 >= n_mbs) return;
    const int n_w = min((uint32_t)WARP_MBS, n_mbs - mb0);
    uint8_t* stage = smem + warp * WARP_STAGE;
    const uint32_t bar = smem_u32(&bars[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_w * (LUMA_BYTES + 2 * CHROMA_BYTES)) : "memory");
    }
    __syncwarp();
    if (lane < n_w) {  // lane = macroblock: three boxes (each descriptor is addressed statically, as a kernel parameter)
        const Rec r = recs[mb0 + lane];
        const uint32_t dst = smem_u32(stage + lane * MB_STAGE);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(&tm_y), "r"((int)(r.ycol & ~15u)), "r"((int)r.yrow), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst + LUMA_SLOT),
                     "l"(&tm_cb), "r"((int)(r.ccol & ~15u)), "r"((int)r.crow), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst + LUMA_SLOT + CHROMA_SLOT),
                     "l"(&tm_cr), "r"((int)(r.ccol & ~15u)), "r"((int)r.crow), "r"(bar)
                     : "memory");
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar), "r"(0)
                : "memory");
        }
    }
    const int mbq = lane >> 3, rg = (lane >> 1) & 3, h = lane & 1;
    if (mbq >= n_w) return;
    const uint8_t* s = stage + mbq * MB_STAGE;
    const Rec r = recs[mb0 + mbq];
    uint32_t acc = 0;
    {
        const uint32_t o = (r.ycol & 15u) + 8 * h, a = o & 3;
        const bool third = a != 0 || (r.hp & 1u), extra = (r.hp & 2u) != 0;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(s + (rg * 4) * LUMA_BOX_W + (o - a));
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint32_t* q = p + k * (LUMA_BOX_W / 4);
            if (k < 4 || extra) acc ^= fold_row(q[0], q[1], third ? q[2] : 0u, a * 8);
        }
    }
    {
        const uint32_t o = r.ccol & 15u, a = o & 3;
        const bool third = a != 0 || (r.hp & 4u), extra = (r.hp & 8u) != 0;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(s + LUMA_SLOT + h * CHROMA_SLOT + (rg * 2) * CHROMA_BOX_W + (o - a));
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t* q = p + k * (CHROMA_BOX_W / 4);
            if (k < 2 || extra) acc ^= fold_row(q[0], q[1], third ? q[2] : 0u, a * 8);
        }
    }
    out[(size_t)tile * 32 + lane] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled enc, void* base, uint64_t pitch, uint64_t rows, uint32_t bw, uint32_t bh) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {pitch, rows}, strides[1] = {pitch};
    const cuuint32_t box[2] = {bw, bh}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(1);
    }
    return m;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const int ctas_a = argc > 1 ? atoi(argv[1]) : 8, ctas_b = argc > 2 ? atoi(argv[2]) : 4;
    const size_t ysz = (size_t)STREAMS * 2 * PITCH_Y * ROWS_Y, csz = (size_t)STREAMS * 2 * PITCH_C * ROWS_C;
    uint8_t *y, *cb, *cr;
    CK(cudaMalloc(&y, ysz));
    CK(cudaMalloc(&cb, csz));
    CK(cudaMalloc(&cr, csz));
    {
        std::vector<uint64_t> h(ysz / 8);
        for (auto& v : h) v = rnd();
        CK(cudaMemcpy(y, h.data(), ysz, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(cb, h.data(), csz, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(cr, h.data() + csz / 8, csz, cudaMemcpyHostToDevice));
    }
    const uint32_t n_mbs = STREAMS * MBW * MBH;
    std::vector<Rec> recs(n_mbs);
    for (uint32_t s = 0, i = 0; s < (uint32_t)STREAMS; s++)
        for (int my = 0; my < MBH; my++)
            for (int mx = 0; mx < MBW; mx++, i++) {
                auto r7 = []() { return (int)(rnd() % 7) - 3; };
                int mvx = r7() + r7(), mvy = r7() + r7();  // half-pel units, synth.cpp mv_mode 0
                if (rnd() % 100 < 8) mvx = (int)(rnd() % 64) - 32, mvy = (int)(rnd() % 64) - 32;
                const int cmx = (mvx >> 1) | (mvx & 1), cmy = (mvy >> 1) | (mvy & 1);  // chroma vector of a 1-vector macroblock
                const int slot = 0;
                Rec& r = recs[i];
                r.yrow = (s * 2 + slot) * ROWS_Y + PAD_Y_ROWS + my * 16 + (mvy >> 1);
                r.ycol = PAD_Y_COLS + mx * 16 + (mvx >> 1);
                r.crow = (s * 2 + slot) * ROWS_C + PAD_C_ROWS + my * 8 + (cmy >> 1);
                r.ccol = PAD_C_COLS + mx * 8 + (cmx >> 1);
                r.hp = (uint32_t)((mvx & 1) | ((mvy & 1) << 1) | ((cmx & 1) << 2) | ((cmy & 1) << 3));
            }
    Rec* d_recs;
    CK(cudaMalloc(&d_recs, recs.size() * sizeof(Rec)));
    CK(cudaMemcpy(d_recs, recs.data(), recs.size() * sizeof(Rec), cudaMemcpyHostToDevice));
    const uint32_t n_tiles = (n_mbs + WARP_MBS - 1) / WARP_MBS, grid = (n_tiles + CTA_WARPS - 1) / CTA_WARPS;
    uint32_t *out_a, *out_b, *out_c;
    CK(cudaMalloc(&out_c, (size_t)n_tiles * 32 * 4));
    CK(cudaMalloc(&out_a, (size_t)n_tiles * 32 * 4));
    CK(cudaMalloc(&out_b, (size_t)n_tiles * 32 * 4));
    CK(cudaMemset(out_a, 0, (size_t)n_tiles * 32 * 4));
    CK(cudaMemset(out_b, 0xFF, (size_t)n_tiles * 32 * 4));

    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    if (!enc) {
        fprintf(stderr, "no cuTensorMapEncodeTiled\n");
        return 1;
    }
    const CUtensorMap tm_y = make_map(enc, y, PITCH_Y, (uint64_t)STREAMS * 2 * ROWS_Y, LUMA_BOX_W, LUMA_BOX_H);
    const CUtensorMap tm_cb = make_map(enc, cb, PITCH_C, (uint64_t)STREAMS * 2 * ROWS_C, CHROMA_BOX_W, CHROMA_BOX_H);
    const CUtensorMap tm_cr = make_map(enc, cr, PITCH_C, (uint64_t)STREAMS * 2 * ROWS_C, CHROMA_BOX_W, CHROMA_BOX_H);

    // shared memory per CTA that yields the wanted number of resident CTAs (196 KB carve-out, 1 KB reserved per CTA)
    auto smem_for = [](int ctas) { return (size_t)(196 * 1024 / ctas - 1024) & ~(size_t)127; };
    const size_t smem_a = smem_for(ctas_a), smem_b = smem_for(ctas_b);
    if (smem_b < (size_t)CTA_WARPS * WARP_STAGE) {
        fprintf(stderr, "staging area does not fit %d CTAs per SM\n", ctas_b);
        return 1;
    }
    CK(cudaFuncSetAttribute(fetch_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    CK(cudaFuncSetAttribute(fetch_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int warm = 3, reps = 20;
    float ms_a = 0, ms_b = 0, ms_c = 0;
    CK(cudaFuncSetAttribute(fetch_ldg_4x8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    for (int v = 0; v < 3; v++) {
        for (int i = 0; i < warm + reps; i++) {
            if (i == warm) CK(cudaEventRecord(e0));
            if (v == 0)
                fetch_ldg<<<grid, CTA_WARPS * 32, smem_a>>>(d_recs, y, cb, cr, n_mbs, out_a);
            else if (v == 2)
                fetch_ldg_4x8<<<grid, CTA_WARPS * 32, smem_a>>>(d_recs, y, cb, cr, n_mbs, out_c);
            else
                fetch_tma<<<grid, CTA_WARPS * 32, smem_b>>>(d_recs, tm_y, tm_cb, tm_cr, n_mbs, out_b);
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(v == 0 ? &ms_a : (v == 1 ? &ms_b : &ms_c), e0, e1));
    }
    std::vector<uint32_t> ha((size_t)n_tiles * 32), hb((size_t)n_tiles * 32);
    CK(cudaMemcpy(ha.data(), out_a, ha.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), out_b, hb.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < ha.size(); i++) bad += ha[i] != hb[i];
    const double window_bytes = (double)n_mbs * (17 * 17 + 2 * 9 * 9);
    printf("prediction fetch, %u macroblocks (1024 CIF pictures), vectors as in the benchmark\n", n_mbs);
    printf("A  per-lane aligned LDG (recon_tile_kernel's pattern), %d CTAs/SM : %.1f us per launch  (%.0f GB/s of window bytes)\n", ctas_a,
           ms_a / reps * 1e3, window_bytes / (ms_a / reps * 1e-3) / 1e9);
    printf("B  TMA boxes (48x17 + 2 x 32x9 per macroblock, 16-byte aligned starts) + LDS, %d CTAs/SM : %.1f us per launch  (%.0f GB/s of window bytes)\n", ctas_b,
           ms_b / reps * 1e3, window_bytes / (ms_b / reps * 1e-3) / 1e9);
    printf("C  per-lane aligned LDG, lane = 4 columns x 8 rows + both chroma planes, %d CTAs/SM : %.1f us per launch  (%.0f GB/s of window bytes)\n",
           ctas_a, ms_c / reps * 1e3, window_bytes / (ms_c / reps * 1e-3) / 1e9);
    printf("results %s (%zu of %zu words differ)\n", bad ? "DIFFER" : "identical", bad, ha.size());
    return bad ? 2 : 0;
}
