// kernels.cuh -- device-side data structures and launch wrappers shared by kernels.cu
// and context.cu.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the driver is reached through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/h263cu.h"

namespace h263dev {

// Device-side picture descriptor, built by the host for every picture of a step when the
// step is run (plane slots toggle per stream, so pointers are only known then).
// The two chroma planes of a picture are stored INTERLEAVED (CbCr pairs, like NV12): sample
// (x, y) of Cb sits at cur[1][y * pitch_c + 2 * x], of Cr one byte further (cur[2] == cur[1] + 1).
struct PicDev {
    uint8_t* cur[3];        // Y, Cb, Cr planes being reconstructed (Cb / Cr: element step CHROMA_STEP)
    const uint8_t* ref[3];  // reference planes (previous picture of the stream); may be null
    uint8_t* rgba;          // RGBA output (null when not requested)
    uint32_t first_event;   // offset of the picture's events in the step's event array
    uint32_t rgba_pitch;    // bytes
    uint16_t w, h;          // true luma dimensions
    uint16_t cw, ch;        // true chroma dimensions = ceil(w/2), ceil(h/2)
    uint16_t pitch_y, pitch_c;  // bytes per row of the luma plane / of the interleaved chroma plane
    uint8_t strength;       // QUANT_TO_STRENGTH[pquant]
    uint8_t flags;
    uint8_t pad[2];
    // the same planes as 32-bit offsets from the context's pools (tiled kernel): interior origin
    // of the Y plane in 4-byte units from y_pool, of the CbCr plane in 4-byte units from c_pool;
    // rgba_row0 = first row of the RGBA picture in the RGBA pool (rows of rgba_pitch bytes)
    uint32_t cur_y4, cur_c4, ref_y4, ref_c4, rgba_row0;
    uint32_t pad2;
};
constexpr int CHROMA_STEP = 2;  // bytes between horizontally adjacent samples of one chroma plane

// Pool base pointers of a context: kernel parameters of the tiled kernel (uniform registers),
// so that per-macroblock state in shared memory can be 32-bit offsets instead of pointers.
struct Pools {
    uint8_t* y;
    uint8_t* c;  // interleaved CbCr planes
    uint8_t* rgba;
    uint32_t pitch_y, pitch_c, rgba_pitch;  // row pitches shared by every plane of the context (bytes)
};

// Fused reconstruction of every macroblock of a step: inverse RLE + dequant + classify +
// IDCT + motion compensation + add/clamp -> planes, and BT.601 RGBA when `emit_rgba`.
// tiled != 0 selects recon_tile_kernel (padded reference planes), else the generic warp-per-macroblock
// recon_mb_kernel.
// tiled: 1 = every picture is a multiple of 16 in size, 2 = some are not (edge fix-up instantiation).
// wide_mv: some picture of the step may hold vectors beyond [-32, 31] half-pel units (no H263CU_PICFLAG_MV_IN_RANGE).
// rgba_map: TMA tensor map over the context's RGBA pool (2-D, rows of rgba_pitch bytes, box 64 bytes x 16 rows):
// the tiled kernel stores RGBA through it.
void launch_recon(const PicDev* pics, const h263cu_mb* mbs, const h263cu_event* events, uint32_t n_mbs,
                  int emit_rgba, int tiled, int wide_mv, const Pools& pools, const CUtensorMap* rgba_map, cudaStream_t stream);
// recon_tile.cu
void launch_recon_tile(const PicDev* pics, const h263cu_mb* mbs, const h263cu_event* events, uint32_t n_mbs,
                       int emit_rgba, int unaligned, int wide_mv, const Pools& pools, const CUtensorMap* rgba_map,
                       cudaStream_t stream);
// how the tiled kernel stores RGBA: 1 = TMA tensor stores from a shared-memory tile (the build default),
// 0 = direct global stores (build variant for A/B measurements)
int recon_tile_uses_tma();

// Plane padding (bytes / rows) reserved around every reconstruction plane: the tiled kernel
// replicates 16 luma pixels / 8 CbCr pairs (16 bytes either way) into it; the extra columns keep
// the interior origin 32 B aligned and absorb aligned-word over-reads.
constexpr int PAD_Y_COLS = 32, PAD_Y_ROWS = 16, PAD_C_COLS = 32, PAD_C_ROWS = 8;

// Deblocking post-filter (per plane, per picture) fused with the RGBA conversion.
// grid = (max tiles per picture, n_pics).
void launch_deblock_rgba(const PicDev* pics, uint32_t n_pics, uint32_t max_w, uint32_t max_h, cudaStream_t stream);
// deblock_tile.cu: the register-resident form for pictures whose sizes are multiples of 16
void launch_deblock_rgba_tile(const PicDev* pics, uint32_t n_pics, uint32_t max_w, uint32_t max_h, cudaStream_t stream);

// Stateless kernels behind h263cu_yuv420_to_rgba / h263cu_deblock (tight planes).
void launch_yuv420_to_rgba(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, uint32_t w, uint32_t h,
                           uint8_t* rgba, cudaStream_t stream);
void launch_deblock_plane(const uint8_t* in, uint8_t* out, uint32_t w, uint32_t h, int strength, cudaStream_t stream);

// Position-weighted checksums of the tight w x h window of a pitched plane:
// out[k] += sum_i (byte[i]+1) * ((uint32)(i*2654435761) | 1), i = y*row_bytes + x.
struct ChecksumJob {
    const uint8_t* base;
    uint32_t row_bytes;  // samples per row that count
    uint32_t rows;
    uint32_t pitch;
    uint32_t out_index;
    uint32_t step;       // bytes between samples (2 for a plane of the interleaved chroma pair)
    uint32_t pad;
};
void launch_checksums(const ChecksumJob* jobs, uint32_t n_jobs, unsigned long long* out, cudaStream_t stream);

}  // namespace h263dev
