// tma_swizzle_probe.cu -- which 16-byte chunk of a shared-memory tile does a TMA tensor STORE put where?
// A 64-byte (then 128-byte) x 16-row box is stored from a 1024-byte aligned tile whose chunk k (16 bytes at offset
// 16k) is filled with the value k, once per swizzle mode; the printed table is the chunk id found at (row, 16-byte
// column) of the destination.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_swizzle_probe tools/tma_swizzle_probe.cu && /tmp/tma_swizzle_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__global__ void probe(const __grid_constant__ CUtensorMap tm) {
    __shared__ __align__(1024) uint8_t tile[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = (uint8_t)(i / 16);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(src), "r"(64), "r"(16)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    const uint64_t pitch = 1408, rows = 64;
    uint8_t* d;
    CK(cudaMalloc(&d, pitch * rows));
    const CUtensorMapSwizzle modes[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
    const char* names[4] = {"NONE", "32B", "64B", "128B"};
    for (int boxw = 64; boxw <= 128; boxw += 64)
        for (int m = 0; m < 4; m++) {
            CUtensorMap tm;
            const cuuint64_t dims[2] = {pitch, rows}, strides[1] = {pitch};
            const cuuint32_t box[2] = {(cuuint32_t)boxw, 16}, es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[m],
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                printf("box %d swizzle %s: encode failed (%d)\n", boxw, names[m], (int)r);
                continue;
            }
            CK(cudaMemset(d, 0xEE, pitch * rows));
            probe<<<1, 128>>>(tm);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("box %d swizzle %s: kernel failed: %s\n", boxw, names[m], cudaGetErrorString(e));
                return 1;
            }
            static uint8_t h[1408 * 64];
            CK(cudaMemcpy(h, d, pitch * rows, cudaMemcpyDeviceToHost));
            printf("box %d x 16, swizzle %s: chunk id at (row, 16-byte column)\n", boxw, names[m]);
            for (int row = 0; row < 16; row++) {
                printf("  row %2d:", row);
                for (int c = 0; c < boxw / 16; c++) printf(" %3d", h[(16 + row) * pitch + 64 + c * 16]);
                printf("\n");
            }
        }
    return 0;
}
