// microbench.cu -- instruction-sequence variants for the epilogue of the recon kernel, timed
// in isolation (register-resident inputs, no memory traffic) to pick the cheapest exact form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/microbench tools/microbench.cu
// Every variant is checked against variant 0 on the same inputs before it is timed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../h263_rs_b200/csrc/device_math.cuh"

using namespace h263dev;

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return (w >> (8 * k)) & 0xFFu; }

// ---- RGBA, 4 pixels: yw = 4 luma bytes, cb2/cr2 = two chroma samples in the low two bytes ----
__device__ __forceinline__ uint4 rgba_v0(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    const ChromaTerms t0 = chroma_terms((int)(cb2 & 0xFF), (int)(cr2 & 0xFF));
    const ChromaTerms t1 = chroma_terms((int)((cb2 >> 8) & 0xFF), (int)((cr2 >> 8) & 0xFF));
    uint4 o;
    o.x = yuv_pixel((int)byte_of(yw, 0), t0);
    o.y = yuv_pixel((int)byte_of(yw, 1), t0);
    o.z = yuv_pixel((int)byte_of(yw, 2), t1);
    o.w = yuv_pixel((int)byte_of(yw, 3), t1);
    return o;
}

__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c) {
    uint32_t d;  // d = c[15:0] << 16 | sat_u8(a) << 8 | sat_u8(b)
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
struct CT {
    int r, g, b;
};
__device__ __forceinline__ CT cterms(int cb, int cr) {
    CT t;  // chroma terms with the -16*76309 luma offset and the rounding constant folded in
    t.r = cr * 104597 + (32768 - 128 * 104597 - 16 * 76309);
    t.g = cr * -53279 + (cb * -25675 + (32768 + 128 * 53279 + 128 * 25675 - 16 * 76309));
    t.b = cb * 132201 + (32768 - 128 * 132201 - 16 * 76309);
    return t;
}
__device__ __forceinline__ uint32_t px_v1(int y, const CT& t) {
    const int r = (y * 76309 + t.r) >> 16, g = (y * 76309 + t.g) >> 16, b = (y * 76309 + t.b) >> 16;
    return pack_sat(g, r, pack_sat(255, b, 0));
}
__device__ __forceinline__ uint4 rgba_v1(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    const CT t0 = cterms((int)(cb2 & 0xFF), (int)(cr2 & 0xFF));
    const CT t1 = cterms((int)((cb2 >> 8) & 0xFF), (int)((cr2 >> 8) & 0xFF));
    uint4 o;
    o.x = px_v1((int)__byte_perm(yw, 0, 0x4440), t0);
    o.y = px_v1((int)__byte_perm(yw, 0, 0x4441), t0);
    o.z = px_v1((int)__byte_perm(yw, 0, 0x4442), t1);
    o.w = px_v1((int)__byte_perm(yw, 0, 0x4443), t1);
    return o;
}
// v2: 64-bit multiply-add, take the high word: ((y << 16) * K + (t << 16)) >> 32 == (y*K + t) >> 16
__device__ __forceinline__ uint32_t px_v2(uint32_t y16, long long tr, long long tg, long long tb) {
    const int r = (int)(((long long)(int)y16 * 76309 + tr) >> 32);
    const int g = (int)(((long long)(int)y16 * 76309 + tg) >> 32);
    const int b = (int)(((long long)(int)y16 * 76309 + tb) >> 32);
    return pack_sat(g, r, pack_sat(255, b, 0));
}
__device__ __forceinline__ uint4 rgba_v2(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    const CT t0 = cterms((int)(cb2 & 0xFF), (int)(cr2 & 0xFF));
    const CT t1 = cterms((int)((cb2 >> 8) & 0xFF), (int)((cr2 >> 8) & 0xFF));
    const long long r0 = (long long)t0.r << 16, g0 = (long long)t0.g << 16, b0 = (long long)t0.b << 16;
    const long long r1 = (long long)t1.r << 16, g1 = (long long)t1.g << 16, b1 = (long long)t1.b << 16;
    uint4 o;
    o.x = px_v2(__byte_perm(yw, 0, 0x4044), r0, g0, b0);  // y << 16
    o.y = px_v2(__byte_perm(yw, 0, 0x4144), r0, g0, b0);
    o.z = px_v2(__byte_perm(yw, 0, 0x4244), r1, g1, b1);
    o.w = px_v2(__byte_perm(yw, 0, 0x4344), r1, g1, b1);
    return o;
}
// v3: high halves via PRMT, packed s16x2 clamp, PRMT packing (two pixels at a time)
__device__ __forceinline__ void px2_v3(int ya, int yb, const CT& t, uint32_t& pa, uint32_t& pb) {
    const int ra = ya * 76309 + t.r, ga = ya * 76309 + t.g, ba = ya * 76309 + t.b;
    const int rb = yb * 76309 + t.r, gb = yb * 76309 + t.g, bb = yb * 76309 + t.b;
    const uint32_t R = __vimin_s16x2_relu(__byte_perm(ra, rb, 0x7632), 0x00FF00FFu);
    const uint32_t G = __vimin_s16x2_relu(__byte_perm(ga, gb, 0x7632), 0x00FF00FFu);
    const uint32_t B = __vimin_s16x2_relu(__byte_perm(ba, bb, 0x7632), 0x00FF00FFu) | 0xFF00FF00u;
    const uint32_t RG = __byte_perm(R, G, 0x6240);  // R0 G0 R1 G1
    pa = __byte_perm(RG, B, 0x5410);
    pb = __byte_perm(RG, B, 0x7632);
}
__device__ __forceinline__ uint4 rgba_v3(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    const CT t0 = cterms((int)(cb2 & 0xFF), (int)(cr2 & 0xFF));
    const CT t1 = cterms((int)((cb2 >> 8) & 0xFF), (int)((cr2 >> 8) & 0xFF));
    uint4 o;
    px2_v3((int)__byte_perm(yw, 0, 0x4440), (int)__byte_perm(yw, 0, 0x4441), t0, o.x, o.y);
    px2_v3((int)__byte_perm(yw, 0, 0x4442), (int)__byte_perm(yw, 0, 0x4443), t1, o.z, o.w);
    return o;
}

// v4: mad.wide.s32 spelled in PTX so that it stays one IMAD.WIDE per channel
__device__ __forceinline__ int madwide_hi(int a, int b, long long c) {
    long long d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return (int)(d >> 32);
}
__device__ __forceinline__ uint32_t px_v4(uint32_t y16, long long tr, long long tg, long long tb) {
    return pack_sat(madwide_hi((int)y16, 76309, tg), madwide_hi((int)y16, 76309, tr), pack_sat(255, madwide_hi((int)y16, 76309, tb), 0));
}
__device__ __forceinline__ uint4 rgba_v4(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    const CT t0 = cterms((int)(cb2 & 0xFF), (int)(cr2 & 0xFF));
    const CT t1 = cterms((int)((cb2 >> 8) & 0xFF), (int)((cr2 >> 8) & 0xFF));
    const long long r0 = (long long)t0.r << 16, g0 = (long long)t0.g << 16, b0 = (long long)t0.b << 16;
    const long long r1 = (long long)t1.r << 16, g1 = (long long)t1.g << 16, b1 = (long long)t1.b << 16;
    uint4 o;
    o.x = px_v4(__byte_perm(yw, 0, 0x4044), r0, g0, b0);
    o.y = px_v4(__byte_perm(yw, 0, 0x4144), r0, g0, b0);
    o.z = px_v4(__byte_perm(yw, 0, 0x4244), r1, g1, b1);
    o.w = px_v4(__byte_perm(yw, 0, 0x4344), r1, g1, b1);
    return o;
}

template <int V>
__device__ __forceinline__ uint4 rgba(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    if (V == 4) return rgba_v4(yw, cb2, cr2);
    if (V == 0) return rgba_v0(yw, cb2, cr2);
    if (V == 1) return rgba_v1(yw, cb2, cr2);
    if (V == 2) return rgba_v2(yw, cb2, cr2);
    return rgba_v3(yw, cb2, cr2);
}

template <int V>
__global__ void __launch_bounds__(256) rgba_bench(uint32_t* out, int iters, uint32_t seed) {
    uint32_t s = seed + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        const uint32_t yw = s, c = s ^ (s >> 7);
#pragma unroll
        for (int k = 0; k < 4; k++) {  // 16 pixels per iteration, like one row pair of a unit
            const uint4 o = rgba<V>(yw + k * 0x01030507u, c + k * 0x11u, (c >> 16) + k * 0x21u);
            acc ^= o.x + 3 * o.y + 5 * o.z + 7 * o.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- MC, one 8-pixel row pair: inputs three words for each of two rows ----
__device__ __forceinline__ uint32_t mc_word_v0(uint32_t r0lo, uint32_t r0hi, uint32_t r1lo, uint32_t r1hi, int sh, int shb) {
    const uint32_t a = __funnelshift_r(r0lo, r0hi, sh), b = __funnelshift_rc(r0lo, r0hi, shb);
    const uint32_t c = __funnelshift_r(r1lo, r1hi, sh), d = __funnelshift_rc(r1lo, r1hi, shb);
    return avg4_u8x4(a, b, c, d);
}
// horizontal sums of one row in 16-bit lanes: hE = (p0+p1', p2+p3'), hO = (p1+p2', p3+p4') ...
struct HSum {
    uint32_t e0, o0, e1, o1;
};
__device__ __forceinline__ HSum hsum(uint32_t w0, uint32_t w1, uint32_t w2, int sh, int shb) {
    const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
    const uint32_t b0 = __funnelshift_rc(w0, w1, shb), b1 = __funnelshift_rc(w1, w2, shb);
    HSum h;
    h.e0 = __byte_perm(a0, 0, 0x4240) + __byte_perm(b0, 0, 0x4240);
    h.o0 = __byte_perm(a0, 0, 0x4341) + __byte_perm(b0, 0, 0x4341);
    h.e1 = __byte_perm(a1, 0, 0x4240) + __byte_perm(b1, 0, 0x4240);
    h.o1 = __byte_perm(a1, 0, 0x4341) + __byte_perm(b1, 0, 0x4341);
    return h;
}
__device__ __forceinline__ uint32_t vmix(uint32_t h0, uint32_t h1) { return ((h0 + h1 + 0x00020002u) >> 2) & 0x00FF00FFu; }

template <int V>
__global__ void __launch_bounds__(256) mc_bench(uint32_t* out, int iters, uint32_t seed) {
    uint32_t s = seed + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        const int sh = (s >> 28 & 3) * 8, ix = (s >> 27) & 1, iy = (s >> 26) & 1, shb = sh + 8 * ix;
        uint32_t w[5][3];
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) w[r][k] = s * (2 * r + 3) + k * 0x9E3779B9u;
        if (V == 0) {
#pragma unroll
            for (int r = 0; r < 4; r++) {  // 8 x 4 unit: 8 output words
                const uint32_t* a = w[r];
                const uint32_t* b = iy ? w[r + 1] : w[r];
                const uint32_t o0 = mc_word_v0(a[0], a[1], b[0], b[1], sh, shb);
                const uint32_t o1 = mc_word_v0(a[1], a[2], b[1], b[2], sh, shb);
                // residual add as in v1: split, add, merge
                const uint32_t p01 = __byte_perm(o0, 0u, 0x4140), p23 = __byte_perm(o0, 0u, 0x4342);
                const uint32_t q01 = __byte_perm(o1, 0u, 0x4140), q23 = __byte_perm(o1, 0u, 0x4342);
                const uint32_t s01 = __viaddmin_s16x2_relu(p01, s, 0x00FF00FFu), s23 = __viaddmin_s16x2_relu(p23, s >> 3, 0x00FF00FFu);
                const uint32_t t01 = __viaddmin_s16x2_relu(q01, s >> 5, 0x00FF00FFu), t23 = __viaddmin_s16x2_relu(q23, s >> 7, 0x00FF00FFu);
                acc ^= __byte_perm(s01, s23, 0x6420) + 3 * __byte_perm(t01, t23, 0x6420);
            }
        } else {
            HSum h[5];
#pragma unroll
            for (int r = 0; r < 5; r++) h[r] = hsum(w[r][0], w[r][1], w[r][2], sh, shb);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const HSum& a = h[r];
                const HSum b = iy ? h[r + 1] : h[r];
                const uint32_t e0 = __viaddmin_s16x2_relu(vmix(a.e0, b.e0), s, 0x00FF00FFu);
                const uint32_t o0 = __viaddmin_s16x2_relu(vmix(a.o0, b.o0), s >> 3, 0x00FF00FFu);
                const uint32_t e1 = __viaddmin_s16x2_relu(vmix(a.e1, b.e1), s >> 5, 0x00FF00FFu);
                const uint32_t o1 = __viaddmin_s16x2_relu(vmix(a.o1, b.o1), s >> 7, 0x00FF00FFu);
                acc ^= __byte_perm(e0, o0, 0x6240) + 3 * __byte_perm(e1, o1, 0x6240);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
__global__ void rgba_check(uint32_t* bad) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < 64; i++) {
        s = s * 1664525u + 1013904223u;
        const uint4 a = rgba<0>(s, s >> 5, s >> 13), b = rgba<V>(s, s >> 5, s >> 13);
        if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) atomicAdd(bad, 1u);
    }
}

template <typename F>
float time_it(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    uint32_t *out, *bad;
    const int grid = 148 * 8, iters = 2048;
    cudaMalloc(&out, grid * 256 * 4);
    cudaMalloc(&bad, 4);
    cudaMemset(bad, 0, 4);
    rgba_check<1><<<1024, 256>>>(bad);
    uint32_t h1, h2, h3;
    cudaMemcpy(&h1, bad, 4, cudaMemcpyDeviceToHost), cudaMemset(bad, 0, 4);
    rgba_check<2><<<1024, 256>>>(bad);
    cudaMemcpy(&h2, bad, 4, cudaMemcpyDeviceToHost), cudaMemset(bad, 0, 4);
    rgba_check<3><<<1024, 256>>>(bad);
    cudaMemcpy(&h3, bad, 4, cudaMemcpyDeviceToHost);
    printf("rgba mismatches vs v0: v1 %u v2 %u v3 %u\n", h1, h2, h3);
    cudaMemset(bad, 0, 4);
    rgba_check<4><<<1024, 256>>>(bad);
    cudaMemcpy(&h3, bad, 4, cudaMemcpyDeviceToHost);
    printf("v4 mismatches %u\n", h3);
    const double px = (double)grid * 256 * iters * 16;
    float ms;
    ms = time_it([&] { rgba_bench<0><<<grid, 256>>>(out, iters, 1); });
    printf("rgba v0 %.3f ms  %.2f Tpx/s\n", ms, px / ms / 1e9);
    ms = time_it([&] { rgba_bench<1><<<grid, 256>>>(out, iters, 1); });
    printf("rgba v1 %.3f ms  %.2f Tpx/s\n", ms, px / ms / 1e9);
    ms = time_it([&] { rgba_bench<2><<<grid, 256>>>(out, iters, 1); });
    printf("rgba v2 %.3f ms  %.2f Tpx/s\n", ms, px / ms / 1e9);
    ms = time_it([&] { rgba_bench<3><<<grid, 256>>>(out, iters, 1); });
    printf("rgba v3 %.3f ms  %.2f Tpx/s\n", ms, px / ms / 1e9);
    ms = time_it([&] { rgba_bench<4><<<grid, 256>>>(out, iters, 1); });
    printf("rgba v4 %.3f ms  %.2f Tpx/s\n", ms, px / ms / 1e9);
    const double mpx = (double)grid * 256 * iters * 32;
    ms = time_it([&] { mc_bench<0><<<grid, 256>>>(out, iters, 1); });
    printf("mc v0 %.3f ms  %.2f Tpx/s\n", ms, mpx / ms / 1e9);
    ms = time_it([&] { mc_bench<1><<<grid, 256>>>(out, iters, 1); });
    printf("mc v1 %.3f ms  %.2f Tpx/s\n", ms, mpx / ms / 1e9);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
