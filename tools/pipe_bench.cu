// pipe_bench.cu -- issue rates of the instruction classes the recon epilogue is built from, alone and mixed, to
// decide which pipe a piece of work should run on (round 2: the kernel is bound by the ALU pipe + issue slots).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/pipe_bench tools/pipe_bench.cu
// Each test: 1024 threads per SM (8 warps per scheduler), N independent chains per thread, reports warp instructions
// per cycle and scheduler (1.0 = the issue limit).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 512

enum Op { PRMT, LOP3, SHF, IADD3, VIADDMNMX, I2IP, IMAD, FFMA, FADD, F2I_U8, F2I_S32, I2F_U8, I2F_U16, I2FP, VIMNMX, IMADWIDE, IMADHI, DP4A, LEA, N_OPS };
static const char* op_names[] = {"PRMT", "LOP3", "SHF", "IADD3", "VIADDMNMX.RELU", "I2IP", "IMAD", "FFMA", "FADD", "F2I.U8.FLOOR", "F2I.S32.TRUNC",
                                 "I2F.U8.Bn", "I2F.U16", "I2FP.F32.U32", "VIMNMX", "IMAD.WIDE", "IMAD.HI", "IDP.4A", "LEA"};

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t k) {
    uint32_t d;
    if (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x2103;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %1, %2, 0x12345, 0x96;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %1, %2, 7;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == IADD3) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == VIADDMNMX) d = (uint32_t)__viaddmin_s32_relu((int)x, (int)k, 0xFFFFFF);
    if (OP == I2IP) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == FFMA) {
        float f;
        asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=f"(f) : "f"(__uint_as_float(x)), "f"(__uint_as_float(k)));
        d = __float_as_uint(f);
    }
    if (OP == FADD) {
        float f;
        asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(x)), "f"(__uint_as_float(k)));
        d = __float_as_uint(f);
    }
    if (OP == F2I_U8) asm volatile("cvt.rmi.u8.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(x)));
    if (OP == F2I_S32) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(x)));
    if (OP == I2F_U8) {
        float f;
        asm volatile("{ .reg .b8 a,b,c,e; mov.b32 {a,b,c,e}, %1; cvt.rn.f32.u8 %0, c; }" : "=f"(f) : "r"(x));
        d = __float_as_uint(f);
    }
    if (OP == I2F_U16) {
        float f;
        asm volatile("{ .reg .b16 a,b; mov.b32 {a,b}, %1; cvt.rn.f32.u16 %0, b; }" : "=f"(f) : "r"(x));
        d = __float_as_uint(f);
    }
    if (OP == I2FP) {
        float f;
        asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x));
        d = __float_as_uint(f);
    }
    if (OP == VIMNMX) asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == IMADWIDE) {
        unsigned long long w, c = ((unsigned long long)k << 32) | x;
        asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(w) : "r"(x), "r"(k), "l"(c));
        d = (uint32_t)(w >> 32);
    }
    if (OP == IMADHI) asm volatile("mad.hi.s32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(k));
    if (OP == LEA) asm volatile("{ .reg .u32 t; shl.b32 t, %1, 3; add.u32 %0, t, %2; }" : "=r"(d) : "r"(x), "r"(k));
    return d;
}

// A : B instructions interleaved na : nb per chain step
template <int A, int B, int NA, int NB>
__global__ void __launch_bounds__(1024, 1) mix_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
    uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = seed * (c + 1) + threadIdx.x, y[c] = seed ^ (c * 77 + threadIdx.x);
    const uint32_t k = seed | 1u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
#pragma unroll
            for (int a = 0; a < NA; a++) x[c] = step<A>(x[c], k);
#pragma unroll
            for (int b = 0; b < NB; b++) y[c] = step<B>(y[c], k);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= x[c] + y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B, int NA, int NB>
void run(uint32_t* out, long long* cyc, const char* label) {
    mix_kernel<A, B, NA, NB><<<148, 1024>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    // per scheduler: 8 warps x ITERS x CHAINS x (NA + NB) warp instructions
    const double inst = 8.0 * ITERS * CHAINS * (NA + NB);
    printf("%-34s  %.3f warp inst / cycle / scheduler  (%.2f cycles per instruction)\n", label, inst / avg, avg / inst);
}

#define SOLO(OP) run<OP, OP, 1, 0>(out, cyc, op_names[OP])
#define MIX(A, B, NA, NB)                                                        \
    {                                                                            \
        char l[96];                                                              \
        snprintf(l, sizeof l, "%d x %s + %d x %s", NA, op_names[A], NB, op_names[B]); \
        run<A, B, NA, NB>(out, cyc, l);                                          \
    }

int main() {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    SOLO(PRMT);
    SOLO(LOP3);
    SOLO(SHF);
    SOLO(IADD3);
    SOLO(VIADDMNMX);
    SOLO(VIMNMX);
    SOLO(I2IP);
    SOLO(IMAD);
    SOLO(FFMA);
    SOLO(FADD);
    SOLO(F2I_U8);
    SOLO(F2I_S32);
    SOLO(I2F_U8);
    SOLO(I2F_U16);
    SOLO(I2FP);
    SOLO(IMADWIDE);
    SOLO(IMADHI);
    SOLO(DP4A);
    SOLO(LEA);
    MIX(PRMT, IMADWIDE, 1, 1);
    MIX(PRMT, IMADWIDE, 2, 1);
    MIX(IMAD, IMADWIDE, 1, 1);
    MIX(FFMA, IMADWIDE, 1, 1);
    MIX(PRMT, IMADHI, 1, 1);
    MIX(PRMT, DP4A, 1, 1);
    MIX(IMAD, DP4A, 1, 1);
    MIX(PRMT, IMAD, 1, 1);
    MIX(PRMT, FFMA, 1, 1);
    MIX(PRMT, FFMA, 1, 2);
    MIX(IMAD, FFMA, 1, 1);
    MIX(PRMT, F2I_U8, 1, 1);
    MIX(FFMA, F2I_U8, 1, 1);
    MIX(FFMA, F2I_U8, 2, 1);
    MIX(IMAD, F2I_U8, 1, 1);
    MIX(PRMT, I2F_U8, 1, 1);
    MIX(F2I_U8, I2F_U8, 3, 1);
    MIX(PRMT, VIADDMNMX, 1, 1);
    MIX(PRMT, I2FP, 1, 1);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
