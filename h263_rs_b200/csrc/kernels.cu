// kernels.cu -- hand-written sm_100a kernels of the reconstruction path.
//
//   recon_kernel          one warp per macroblock: inverse RLE + dequantisation + block
//                         classification (rle.rs:82-172), f32 IDCT in the reference's
//                         operation order (idct.rs:52-65, 82-201), full/half-pel motion
//                         compensation with edge clamping (gather.rs:16-126, 140-204),
//                         residual add + clamp, plane stores, fused BT.601 RGBA
//                         (bt601.rs:12-59) with 128-bit stores.
//   deblock_rgba_kernel   deblocking post-filter (deblock.rs:29-42, 99-127, 136-299) on
//                         32x32 tiles with a 2-pixel halo, fused with the RGBA conversion.
//   yuv420_to_rgba_kernel / deblock_plane_kernel   stateless sibling-crate drop-ins.
//   checksum_kernel       position-weighted checksums for full-size parity checks.
//
// Bound: HBM bandwidth / FP32+INT issue rate; no tensor cores (the 8x8 transform must keep
// the reference's summation order, and is not a dense contraction worth them).
// Compile with -fmad=false; the transform additionally uses __fmul_rn/__fadd_rn, which are
// never contracted.
#include "device_math.cuh"
#include "kernels.cuh"

namespace h263dev {

__constant__ float c_basis[8][8] = H263_BASIS_TABLE;
__constant__ uint8_t c_dezigzag[64] = H263_DEZIGZAG_LINEAR;

constexpr int WARPS_PER_CTA = 8;
constexpr int COEF_STRIDE = 68;  // floats per coefficient block: 64 + 4 pad (keeps 16 B alignment,
                                 // staggers the blocks over the banks)
constexpr unsigned FULL = 0xFFFFFFFFu;

enum { CLS_ZERO = 0, CLS_DC = 1, CLS_VERT = 3, CLS_FULL = 4 };  // Horiz is computed as Full (bit-identical)

// per-block info bits gathered while scattering events
constexpr uint32_t INFO_ROWS = 0xFFu;   // bit y: a coefficient event landed in row y
constexpr uint32_t INFO_COL = 0x100u;   // some event landed in a column x > 0
constexpr uint32_t INFO_DC = 0x200u;    // intra DC present
constexpr uint32_t INFO_OVF = 0x400u;   // zig-zag overflow: the block stays Zero (rle.rs:125-127)

struct __align__(16) WarpScratch {
    float coef[6 * COEF_STRIDE];  // dequantised coefficients, [block][y*8+x]
    float tbuf[64];               // row-pass output of the block in flight, [y*8+i]
    int16_t res[6][64];           // rounded residuals, [block][row*8+col]
    uint8_t rec[384];             // reconstructed MB: Y 16x16 | Cb 8x8 | Cr 8x8
};

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return (w >> (8 * k)) & 0xFFu; }

// clamp(pred + r, 0, 255) on four packed pixels
__device__ __forceinline__ uint32_t add_clamp4(uint32_t pred, int r0, int r1, int r2, int r3) {
    uint32_t o0 = (uint32_t)clamp_u8((int)byte_of(pred, 0) + r0);
    uint32_t o1 = (uint32_t)clamp_u8((int)byte_of(pred, 1) + r1);
    uint32_t o2 = (uint32_t)clamp_u8((int)byte_of(pred, 2) + r2);
    uint32_t o3 = (uint32_t)clamp_u8((int)byte_of(pred, 3) + r3);
    return o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
}

// Prediction for 8 horizontally adjacent pixels at (x0, y0) of a W x H plane, displaced by
// the half-pel vector (mvx, mvy).  Sample coordinates clamp to the plane (read_sample,
// gather.rs:16-31 = unrestricted-MV border extension); one direction interpolates with
// (a+b+1)>>1, both with (a+b+c+d+2)>>2 (gather.rs:34-40, 103-113).
__device__ __forceinline__ void mc_fetch8(const uint8_t* __restrict__ ref, int pitch, int W, int H, int x0, int y0,
                                          int mvx, int mvy, uint32_t& o0, uint32_t& o1) {
    const int dx = mvx >> 1, ix = mvx & 1, dy = mvy >> 1, iy = mvy & 1;  // floor / odd (types.rs:721-729)
    const int sx = x0 + dx, sy = y0 + dy;
    uint32_t a0, a1, b0 = 0, b1 = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    const bool inside = sx >= 0 && sy >= 0 && sx + 8 + ix <= W && sy + 1 + iy <= H;
    if (inside) {
        const int a = sx & 3;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(ref + (size_t)sy * pitch + (sx - a));
        const int sh = a * 8;
        uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
        a0 = __funnelshift_r(w0, w1, sh);
        a1 = __funnelshift_r(w1, w2, sh);
        if (ix) {
            b0 = __funnelshift_rc(w0, w1, sh + 8);
            b1 = __funnelshift_rc(w1, w2, sh + 8);
        }
        if (iy) {
            const uint32_t* wq = wp + (pitch >> 2);
            uint32_t v0 = __ldg(wq), v1 = __ldg(wq + 1), v2 = __ldg(wq + 2);
            c0 = __funnelshift_r(v0, v1, sh);
            c1 = __funnelshift_r(v1, v2, sh);
            if (ix) {
                d0 = __funnelshift_rc(v0, v1, sh + 8);
                d1 = __funnelshift_rc(v1, v2, sh + 8);
            }
        }
    } else {
        uint32_t px[2][9];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int cy = min(max(sy + r, 0), H - 1);
            const uint8_t* row = ref + (size_t)cy * pitch;
#pragma unroll
            for (int k = 0; k < 9; k++) px[r][k] = row[min(max(sx + k, 0), W - 1)];
        }
        a0 = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[0][3] << 24);
        a1 = px[0][4] | (px[0][5] << 8) | (px[0][6] << 16) | (px[0][7] << 24);
        b0 = px[0][1] | (px[0][2] << 8) | (px[0][3] << 16) | (px[0][4] << 24);
        b1 = px[0][5] | (px[0][6] << 8) | (px[0][7] << 16) | (px[0][8] << 24);
        c0 = px[1][0] | (px[1][1] << 8) | (px[1][2] << 16) | (px[1][3] << 24);
        c1 = px[1][4] | (px[1][5] << 8) | (px[1][6] << 16) | (px[1][7] << 24);
        d0 = px[1][1] | (px[1][2] << 8) | (px[1][3] << 16) | (px[1][4] << 24);
        d1 = px[1][5] | (px[1][6] << 8) | (px[1][7] << 16) | (px[1][8] << 24);
    }
    if (ix && iy) {
        o0 = avg4_u8x4(a0, b0, c0, d0);
        o1 = avg4_u8x4(a1, b1, c1, d1);
    } else if (ix) {
        o0 = avg2_u8x4(a0, b0);
        o1 = avg2_u8x4(a1, b1);
    } else if (iy) {
        o0 = avg2_u8x4(a0, c0);
        o1 = avg2_u8x4(a1, c1);
    } else {
        o0 = a0;
        o1 = a1;
    }
}

__device__ __forceinline__ void load_event(const h263cu_event* __restrict__ ev, uint32_t idx, bool wide, int& run,
                                           int& level) {
    if (wide) {
        run = __ldg(ev + 2 * idx) & 63;
        level = (int16_t)__ldg(ev + 2 * idx + 1);
    } else {
        uint32_t u = __ldg(ev + idx);
        run = (int)(u >> 10);
        level = ((int)(u << 22)) >> 22;  // sign-extend the 10-bit level
    }
}

// ---------------------------------------------------------------------------------------
// recon_mb_kernel (v0): one warp per macroblock, generic (any picture size, any vector,
// clamped sample fetches).  Kept as the fallback for pictures whose size is not a multiple
// of 16 and as an independent second implementation for on-GPU A/B checks of the tiled
// kernel below.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
    recon_mb_kernel(const PicDev* __restrict__ pics, const h263cu_mb* __restrict__ mbs,
                    const h263cu_event* __restrict__ events, uint32_t n_mbs, int emit_rgba) {
    __shared__ WarpScratch scratch[WARPS_PER_CTA];
    __shared__ uint8_t s_dezigzag[64];
    if (threadIdx.x < 64) s_dezigzag[threadIdx.x] = c_dezigzag[threadIdx.x];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mb_idx = blockIdx.x * WARPS_PER_CTA + warp;
    if (mb_idx >= n_mbs) return;  // warp-uniform
    WarpScratch& S = scratch[warp];

    // ---- macroblock record (24 bytes = 6 words) -------------------------------------------
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(mbs + mb_idx);
    const uint32_t mword = lane < 6 ? __ldg(mw + lane) : 0u;
    const uint32_t w0 = __shfl_sync(FULL, mword, 0), w1 = __shfl_sync(FULL, mword, 1);
    const uint32_t w2 = __shfl_sync(FULL, mword, 2), w3 = __shfl_sync(FULL, mword, 3);
    const uint32_t w4 = __shfl_sync(FULL, mword, 4), w5 = __shfl_sync(FULL, mword, 5);
    const PicDev& P = pics[w1 & 0xFFFFu];
    const int mbx = (w1 >> 16) & 0xFF, mby = w1 >> 24;
    const bool inter = (w2 & H263CU_MB_INTER) != 0;
    const bool wide = (w2 & H263CU_MB_WIDE) != 0;
    const int quant = (w2 >> 8) & 0xFF;
    const uint32_t c1 = (w2 >> 16) & 0xFF, c2 = c1 + (w2 >> 24), c3 = c2 + (w3 & 0xFF), c4 = c3 + ((w3 >> 8) & 0xFF),
                   c5 = c4 + ((w3 >> 16) & 0xFF), total = c5 + (w3 >> 24);
    const h263cu_event* ev = events + P.first_event + w0;

    uint32_t info[6] = {0, 0, 0, 0, 0, 0};
    const bool has_coefs = total > 0 || !inter;
    if (has_coefs) {
        // ---- zero the coefficient blocks, then scatter the dequantised events ----------------
        float4* cz = reinterpret_cast<float4*>(S.coef);
        for (int i = lane; i < 6 * COEF_STRIDE / 4; i += 32) cz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const int first_idx = inter ? 0 : 1;  // intra: DC occupies zig-zag index 0 (rle.rs:117-121)
        if (total <= 32) {
            // one event per lane; the zig-zag position is a segmented prefix sum of (run + 1)
            const bool active = (uint32_t)lane < total;
            const uint32_t e = (uint32_t)lane;
            const int b = (e >= c1) + (e >= c2) + (e >= c3) + (e >= c4) + (e >= c5);
            const int seg_start = b == 0 ? 0 : (b == 1 ? c1 : (b == 2 ? c2 : (b == 3 ? c3 : (b == 4 ? c4 : c5))));
            int run = 0, level = 1;
            if (active) load_event(ev, e, wide, run, level);
            int v = active ? run + 1 : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(FULL, v, d);
                if (lane - d >= seg_start) v += t;
            }
            const int pos = first_idx + v - 1;
            const bool ok = active && pos < 64;
            const uint32_t ovf_ballot = __ballot_sync(FULL, active && pos >= 64);
            uint32_t bits = 0;
            if (ok) {
                const int lin = s_dezigzag[pos];
                S.coef[b * COEF_STRIDE + lin] = (float)dequant(level, quant);
                bits = (1u << (lin >> 3)) | ((lin & 7) ? INFO_COL : 0u);
            }
            const uint32_t cs[7] = {0, c1, c2, c3, c4, c5, total};
#pragma unroll
            for (int bb = 0; bb < 6; bb++) {
                uint32_t m = __reduce_or_sync(FULL, (ok && b == bb) ? bits : 0u);
                const uint64_t seg = ((1ull << cs[bb + 1]) - 1ull) ^ ((1ull << cs[bb]) - 1ull);
                if (ovf_ballot & (uint32_t)seg) m |= INFO_OVF;
                info[bb] = m;
            }
        } else {
            // rare: more than 32 events in the macroblock -> lanes 0..5 walk one block each
            uint32_t mine = 0;
            if (lane < 6) {
                const uint32_t cs[7] = {0, c1, c2, c3, c4, c5, total};
                int idx = first_idx;
                for (uint32_t k = cs[lane]; k < cs[lane + 1]; k++) {
                    int run, level;
                    load_event(ev, k, wide, run, level);
                    idx += run;
                    if (idx >= 64) {
                        mine |= INFO_OVF;
                        break;
                    }
                    const int lin = s_dezigzag[idx];
                    S.coef[lane * COEF_STRIDE + lin] = (float)dequant(level, quant);
                    mine |= (1u << (lin >> 3)) | ((lin & 7) ? INFO_COL : 0u);
                    idx += 1;
                }
            }
#pragma unroll
            for (int bb = 0; bb < 6; bb++) info[bb] = __shfl_sync(FULL, mine, bb);
        }
        if (!inter) {
            // INTRADC codes: bytes 0..3 of w4, 0..1 of w5; 0 = dropped block
            uint32_t code = 0;
            if (lane < 6) code = lane < 4 ? byte_of(w4, lane) : byte_of(w5, lane - 4);
            uint32_t my_info = __shfl_sync(FULL, 0u, 0);
#pragma unroll
            for (int bb = 0; bb < 6; bb++)
                if (lane == bb) my_info = info[bb];
            const bool dc_ok = lane < 6 && code != 0 && !(my_info & INFO_OVF);
            if (dc_ok) S.coef[lane * COEF_STRIDE] = (float)intradc_level((int)code);
            const uint32_t dc_ballot = __ballot_sync(FULL, dc_ok);
#pragma unroll
            for (int bb = 0; bb < 6; bb++)
                if ((dc_ballot >> bb) & 1u) info[bb] |= INFO_DC;
        }
        __syncwarp();
    }

    // ---- classification (rle.rs:94-171) and block-serial IDCT ---------------------------------
    int cls[6];
    int dcres[6];
#pragma unroll
    for (int bb = 0; bb < 6; bb++) {
        const uint32_t m = info[bb];
        const uint32_t rows_ev = m & INFO_ROWS;
        int c;
        if ((m & INFO_OVF) || !(m & (INFO_ROWS | INFO_DC)))
            c = CLS_ZERO;
        else if (!(rows_ev & 0xFEu) && !(m & INFO_COL))
            c = CLS_DC;
        else if (!(m & INFO_COL))
            c = CLS_VERT;
        else
            c = CLS_FULL;
        cls[bb] = c;
        dcres[bb] = 0;
        if (c == CLS_DC) dcres[bb] = round_residual_dc(S.coef[bb * COEF_STRIDE]);
    }

    {
        const int i = lane & 7, jg = lane >> 3;
        float bi[8], bj0[8], bj1[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            bi[k] = c_basis[k][i];
            bj0[k] = c_basis[k][jg];
            bj1[k] = c_basis[k][jg + 4];
        }
#pragma unroll
        for (int bb = 0; bb < 6; bb++) {
            if (cls[bb] != CLS_FULL && cls[bb] != CLS_VERT) continue;  // warp-uniform
            const bool vert = cls[bb] == CLS_VERT;
            const uint32_t rows = (info[bb] & INFO_ROWS) | ((info[bb] & INFO_DC) ? 1u : 0u);
            const int n = __popc(rows);
            const float* cb = S.coef + bb * COEF_STRIDE;
            // row pass: t[y][i] = sum_x c[y][x] * B[x][i], ascending x (idct_1d, idct.rs:52-65);
            // only rows that hold a coefficient (all-zero rows give +0 and add nothing later)
            for (int base = 0; base < n; base += 4) {
                const int ridx = base + jg;
                if (ridx < n) {
                    const int y = __fns(rows, 0, ridx + 1);
                    float t;
                    if (vert) {
                        t = cb[y * 8];  // Vert: idct_1d runs over the first column directly (idct.rs:152-153)
                    } else {
                        const float4 ca = *reinterpret_cast<const float4*>(cb + y * 8);
                        const float4 cc = *reinterpret_cast<const float4*>(cb + y * 8 + 4);
                        t = 0.0f;
                        t = fadd(t, fmul(ca.x, bi[0]));
                        t = fadd(t, fmul(ca.y, bi[1]));
                        t = fadd(t, fmul(ca.z, bi[2]));
                        t = fadd(t, fmul(ca.w, bi[3]));
                        t = fadd(t, fmul(cc.x, bi[4]));
                        t = fadd(t, fmul(cc.y, bi[5]));
                        t = fadd(t, fmul(cc.z, bi[6]));
                        t = fadd(t, fmul(cc.w, bi[7]));
                    }
                    S.tbuf[y * 8 + i] = t;
                }
            }
            __syncwarp();
            // column pass: out[i][j] = sum_y t[y][i] * B[y][j], ascending y; pixel (x=i, y=j)
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int y = 0; y < 8; y++) {
                if ((rows >> y) & 1u) {
                    const float tv = S.tbuf[y * 8 + i];
                    a0 = fadd(a0, fmul(tv, bj0[y]));
                    a1 = fadd(a1, fmul(tv, bj1[y]));
                }
            }
            const int r0 = vert ? round_residual_scaled(a0) : round_residual(a0);
            const int r1 = vert ? round_residual_scaled(a1) : round_residual(a1);
            S.res[bb][jg * 8 + i] = (int16_t)r0;
            S.res[bb][(jg + 4) * 8 + i] = (int16_t)r1;
            __syncwarp();
        }
    }

    // ---- motion compensation + residual add + clamp -> planes ---------------------------------
    const int mv0x = (int8_t)byte_of(w4, 0), mv0y = (int8_t)byte_of(w4, 1);
    const int mv1x = (int8_t)byte_of(w4, 2), mv1y = (int8_t)byte_of(w4, 3);
    const int mv2x = (int8_t)byte_of(w5, 0), mv2y = (int8_t)byte_of(w5, 1);
    const int mv3x = (int8_t)byte_of(w5, 2), mv3y = (int8_t)byte_of(w5, 3);
    {
        // luma: lane -> pixel row (lane >> 1), 8-pixel half (lane & 1)
        const int rowpix = lane >> 1, half = lane & 1;
        const int b = ((rowpix >> 3) << 1) | half, j = rowpix & 7;
        const int c = half ? (rowpix >= 8 ? cls[3] : cls[1]) : (rowpix >= 8 ? cls[2] : cls[0]);
        const int dcv = half ? (rowpix >= 8 ? dcres[3] : dcres[1]) : (rowpix >= 8 ? dcres[2] : dcres[0]);
        uint32_t p0 = 0, p1 = 0;
        if (inter && P.ref[0]) {
            const int mvx = half ? (rowpix >= 8 ? mv3x : mv1x) : (rowpix >= 8 ? mv2x : mv0x);
            const int mvy = half ? (rowpix >= 8 ? mv3y : mv1y) : (rowpix >= 8 ? mv2y : mv0y);
            mc_fetch8(P.ref[0], P.pitch_y, P.w, P.h, mbx * 16 + half * 8, mby * 16 + rowpix, mvx, mvy, p0, p1);
        }
        if (c == CLS_DC) {
            p0 = add_clamp4(p0, dcv, dcv, dcv, dcv);
            p1 = add_clamp4(p1, dcv, dcv, dcv, dcv);
        } else if (c != CLS_ZERO) {
            const int4 rv = *reinterpret_cast<const int4*>(&S.res[b][j * 8]);
            p0 = add_clamp4(p0, (int16_t)(rv.x & 0xFFFF), rv.x >> 16, (int16_t)(rv.y & 0xFFFF), rv.y >> 16);
            p1 = add_clamp4(p1, (int16_t)(rv.z & 0xFFFF), rv.z >> 16, (int16_t)(rv.w & 0xFFFF), rv.w >> 16);
        }
        *reinterpret_cast<uint2*>(P.cur[0] + (size_t)(mby * 16 + rowpix) * P.pitch_y + mbx * 16 + half * 8) =
            make_uint2(p0, p1);
        *reinterpret_cast<uint2*>(&S.rec[rowpix * 16 + half * 8]) = make_uint2(p0, p1);
    }
    if (lane < 16) {
        // chroma: lane -> plane (lane >> 3), row (lane & 7); both planes share one vector
        const int plane = lane >> 3, j = lane & 7, b = 4 + plane;
        const int c = plane ? cls[5] : cls[4];
        const int dcv = plane ? dcres[5] : dcres[4];
        uint32_t p0 = 0, p1 = 0;
        if (inter && P.ref[1 + plane]) {
            const int cx = average_sum_of_mvs(mv0x + mv1x + mv2x + mv3x);
            const int cy = average_sum_of_mvs(mv0y + mv1y + mv2y + mv3y);
            mc_fetch8(P.ref[1 + plane], P.pitch_c, P.cw, P.ch, mbx * 8, mby * 8 + j, cx, cy, p0, p1);
        }
        if (c == CLS_DC) {
            p0 = add_clamp4(p0, dcv, dcv, dcv, dcv);
            p1 = add_clamp4(p1, dcv, dcv, dcv, dcv);
        } else if (c != CLS_ZERO) {
            const int4 rv = *reinterpret_cast<const int4*>(&S.res[b][j * 8]);
            p0 = add_clamp4(p0, (int16_t)(rv.x & 0xFFFF), rv.x >> 16, (int16_t)(rv.y & 0xFFFF), rv.y >> 16);
            p1 = add_clamp4(p1, (int16_t)(rv.z & 0xFFFF), rv.z >> 16, (int16_t)(rv.w & 0xFFFF), rv.w >> 16);
        }
        *reinterpret_cast<uint2*>(P.cur[1 + plane] + (size_t)(mby * 8 + j) * P.pitch_c + mbx * 8) = make_uint2(p0, p1);
        *reinterpret_cast<uint2*>(&S.rec[256 + plane * 64 + j * 8]) = make_uint2(p0, p1);
    }

    // ---- fused BT.601 YUV420 -> RGBA (bt601.rs:12-59), one 128-bit store per 4 pixels ---------
    if (emit_rgba && P.rgba) {
        __syncwarp();
#pragma unroll
        for (int g = lane; g < 64; g += 32) {
            const int row = g >> 2, xq = g & 3;
            const uint32_t yw = *reinterpret_cast<const uint32_t*>(&S.rec[row * 16 + xq * 4]);
            const uint32_t cbp = *reinterpret_cast<const uint16_t*>(&S.rec[256 + (row >> 1) * 8 + xq * 2]);
            const uint32_t crp = *reinterpret_cast<const uint16_t*>(&S.rec[320 + (row >> 1) * 8 + xq * 2]);
            const ChromaTerms t0 = chroma_terms((int)(cbp & 0xFF), (int)(crp & 0xFF));
            const ChromaTerms t1 = chroma_terms((int)(cbp >> 8), (int)(crp >> 8));
            uint4 o;
            o.x = yuv_pixel((int)byte_of(yw, 0), t0);
            o.y = yuv_pixel((int)byte_of(yw, 1), t0);
            o.z = yuv_pixel((int)byte_of(yw, 2), t1);
            o.w = yuv_pixel((int)byte_of(yw, 3), t1);
            *reinterpret_cast<uint4*>(P.rgba + (size_t)(mby * 16 + row) * P.rgba_pitch + (size_t)(mbx * 16 + xq * 4) * 4) = o;
        }
    }
}

// ---------------------------------------------------------------------------------------
// recon_tile_kernel (v1): one CTA per TILE_MBS consecutive macroblocks, three phases with the
// thread granularity each one wants:
//   phase 1  one thread per 8x8 block: walk the block's run/level events, dequantise, scatter
//            into a shared-memory coefficient slot, classify (rle.rs:82-172)
//   phase 2  8 lanes per coded block, 4 blocks per warp over the compacted slot list: row pass,
//            column pass, rounding (idct.rs:52-65, 170-198); residuals overwrite the slot
//   phase 3  one thread per 16x2 luma pixels + the 4+4 chroma pixels under them: branch-free
//            packed-byte motion compensation from PADDED reference planes, packed residual
//            add + clamp, plane stores, border replication, BT.601 RGBA with 128-bit stores;
//            everything stays in registers.
// Requires picture sizes that are multiples of 16 and reference planes whose 16 (luma) / 8
// (chroma) pixel border has been replicated by this kernel (the unrestricted-MV extension
// of gather.rs:16-31, materialised once per picture instead of clamping per sample).
// Vectors outside the baseline range [-32, 31] take the clamped path.
// ---------------------------------------------------------------------------------------
constexpr int TILE_MBS = 32;
constexpr int TILE_THREADS = 256;
constexpr int SLOT_CAP = 96;       // coefficient slots per pass (= 16 MBs x 6 blocks worst case)
constexpr int SLOT_FLOATS = 68;    // 64 + 4 pad: 16 B aligned, 4 consecutive slots hit distinct banks

// block meta word: cls[2:0] | rows[10:3] | slot[18:11] | dcres[28:19] (signed 10 bit)
__device__ __forceinline__ uint32_t pack_meta(int cls, uint32_t rows, uint32_t slot, int dcres) {
    return (uint32_t)cls | (rows << 3) | (slot << 11) | (((uint32_t)dcres & 0x3FFu) << 19);
}

// clamp(pred + residual, 0, 255) on four packed pixels; residuals as two s16x2 words
__device__ __forceinline__ uint32_t add_res4(uint32_t pred, uint32_t r01, uint32_t r23) {
    const uint32_t p01 = __byte_perm(pred, 0u, 0x4140);
    const uint32_t p23 = __byte_perm(pred, 0u, 0x4342);
    const uint32_t s01 = __viaddmin_s16x2_relu(p01, r01, 0x00FF00FFu);
    const uint32_t s23 = __viaddmin_s16x2_relu(p23, r23, 0x00FF00FFu);
    return __byte_perm(s01, s23, 0x6420);
}

// Four packed pixels of prediction from two rows of three aligned words each.
// sh = 8 * (address & 3); shb = sh + 8 when interpolating horizontally, else sh.
// With shb == sh and row1 == row0 the four-tap formula degenerates exactly:
// (4a+2)>>2 = a, (2a+2b+2)>>2 = (a+b+1)>>1 -- so one branch-free form covers all four modes.
__device__ __forceinline__ uint32_t mc_word(uint32_t r0lo, uint32_t r0hi, uint32_t r1lo, uint32_t r1hi, int sh, int shb) {
    const uint32_t a = __funnelshift_r(r0lo, r0hi, sh), b = __funnelshift_rc(r0lo, r0hi, shb);
    const uint32_t c = __funnelshift_r(r1lo, r1hi, sh), d = __funnelshift_rc(r1lo, r1hi, shb);
    return avg4_u8x4(a, b, c, d);
}

__device__ __forceinline__ uint4 rgba4(uint32_t yw, uint32_t cb2, uint32_t cr2) {
    // 4 luma pixels (bytes of yw) with two chroma samples (low two bytes of cb2 / cr2)
    const ChromaTerms t0 = chroma_terms((int)(cb2 & 0xFF), (int)(cr2 & 0xFF));
    const ChromaTerms t1 = chroma_terms((int)((cb2 >> 8) & 0xFF), (int)((cr2 >> 8) & 0xFF));
    uint4 o;
    o.x = yuv_pixel((int)byte_of(yw, 0), t0);
    o.y = yuv_pixel((int)byte_of(yw, 1), t0);
    o.z = yuv_pixel((int)byte_of(yw, 2), t1);
    o.w = yuv_pixel((int)byte_of(yw, 3), t1);
    return o;
}

struct TileSmem {
    float pool[SLOT_CAP * SLOT_FLOATS];  // coefficient slots -> row-pass output -> residuals (in place)
    uint32_t mbrec[TILE_MBS * 6];        // the tile's macroblock records
    uint32_t meta[TILE_MBS * 6];         // per block: class / rows / slot / DC residual
    uint32_t slotmeta[SLOT_CAP];         // per slot: rows | cls << 8
    float basis[64];
    uint32_t warp_count[8];
    uint8_t dezigzag[64];
};

__global__ void __launch_bounds__(TILE_THREADS, 4)
    recon_tile_kernel(const PicDev* __restrict__ pics, const h263cu_mb* __restrict__ mbs,
                      const h263cu_event* __restrict__ events, uint32_t n_mbs, int emit_rgba) {
    __shared__ __align__(16) TileSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mb_first = blockIdx.x * TILE_MBS;
    const int n_tile = (int)min((uint32_t)TILE_MBS, n_mbs - mb_first);

    // ---- stage the macroblock records and the constant tables ---------------------------------
    if (tid < n_tile * 6) S.mbrec[tid] = __ldg(reinterpret_cast<const uint32_t*>(mbs + mb_first) + tid);
    if (tid < 64) {
        S.basis[tid] = c_basis[tid >> 3][tid & 7];
        S.dezigzag[tid] = c_dezigzag[tid];
    }
    __syncthreads();

    // block handled by this thread in phase 1 (tid < 192): macroblock tid / 6, block tid % 6
    const int p1_mb = tid / 6, p1_b = tid - p1_mb * 6;
    const bool p1_valid = p1_mb < n_tile;
    uint32_t p1_nev = 0;
    if (p1_valid) {
        const uint32_t* r = &S.mbrec[p1_mb * 6];
        p1_nev = p1_b < 2 ? (r[2] >> (16 + 8 * p1_b)) & 0xFF : (r[3] >> (8 * (p1_b - 2))) & 0xFF;
    }
    // slots needed by the whole tile decide between one pass (32 MBs) and two passes (16 + 16)
    {
        const uint32_t bal = __ballot_sync(FULL, p1_nev > 0);
        if (lane == 0) S.warp_count[warp] = __popc(bal);
    }
    __syncthreads();
    int total_slots = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) total_slots += (int)S.warp_count[w];
    const int n_pass = total_slots <= SLOT_CAP ? 1 : 2;
    const int mb_per_pass = n_pass == 1 ? TILE_MBS : TILE_MBS / 2;
    __syncthreads();

    const float bt0 = S.basis[0 * 8 + (lane & 7)], bt1 = S.basis[1 * 8 + (lane & 7)], bt2 = S.basis[2 * 8 + (lane & 7)],
                bt3 = S.basis[3 * 8 + (lane & 7)], bt4 = S.basis[4 * 8 + (lane & 7)], bt5 = S.basis[5 * 8 + (lane & 7)],
                bt6 = S.basis[6 * 8 + (lane & 7)], bt7 = S.basis[7 * 8 + (lane & 7)];
    const float bt[8] = {bt0, bt1, bt2, bt3, bt4, bt5, bt6, bt7};

    for (int pass = 0; pass < n_pass; pass++) {
        const int m0 = pass * mb_per_pass, m1 = min(n_tile, m0 + mb_per_pass);
        if (m0 >= m1) break;  // uniform

        // ================= phase 1: slot assignment, zeroing, event walk ========================
        const bool mine = p1_valid && p1_mb >= m0 && p1_mb < m1;
        const bool want_slot = mine && p1_nev > 0;
        const uint32_t bal = __ballot_sync(FULL, want_slot);
        if (lane == 0) S.warp_count[warp] = __popc(bal);
        __syncthreads();
        int slot = __popc(bal & ((1u << lane) - 1u));
        int n_slots = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const int c = (int)S.warp_count[w];
            if (w < warp) slot += c;
            n_slots += c;
        }
        {
            float4* pz = reinterpret_cast<float4*>(S.pool);
            for (int i = tid; i < n_slots * (SLOT_FLOATS / 4); i += TILE_THREADS) pz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        if (mine) {
            const uint32_t* r = &S.mbrec[p1_mb * 6];
            const uint32_t w1 = r[1], w2 = r[2], w3 = r[3];
            const bool inter = (w2 & H263CU_MB_INTER) != 0, wide = (w2 & H263CU_MB_WIDE) != 0;
            const int quant = (w2 >> 8) & 0xFF;
            // events of the blocks before this one inside the macroblock
            const uint32_t n0 = (w2 >> 16) & 0xFF, n1 = w2 >> 24, n2 = w3 & 0xFF, n3 = (w3 >> 8) & 0xFF, n4 = (w3 >> 16) & 0xFF;
            uint32_t before = 0;
            before += p1_b > 0 ? n0 : 0;
            before += p1_b > 1 ? n1 : 0;
            before += p1_b > 2 ? n2 : 0;
            before += p1_b > 3 ? n3 : 0;
            before += p1_b > 4 ? n4 : 0;
            const h263cu_event* ev = events + pics[w1 & 0xFFFFu].first_event + r[0];
            float* cslot = S.pool + slot * SLOT_FLOATS;
            int idx = inter ? 0 : 1;
            uint32_t rows = 0, col = 0;
            bool ovf = false;
            int v00 = 0;
            for (uint32_t k = 0; k < p1_nev; k++) {
                int run, level;
                load_event(ev, before + k, wide, run, level);
                idx += run;
                if (idx >= 64) {
                    ovf = true;  // the whole block stays Zero, DC included (rle.rs:125-127)
                    break;
                }
                const int lin = S.dezigzag[idx];
                const int val = dequant(level, quant);
                cslot[lin] = (float)val;
                if (lin == 0) v00 = val;
                rows |= 1u << (lin >> 3);
                col |= (lin & 7) ? 1u : 0u;
                idx += 1;
            }
            bool has_dc = false;
            int dcv = 0;
            if (!inter) {
                const uint32_t code = p1_b < 4 ? byte_of(r[4], p1_b) : byte_of(r[5], p1_b - 4);
                has_dc = code != 0 && !ovf;
                dcv = intradc_level((int)code);
                if (has_dc && p1_nev > 0) cslot[0] = (float)dcv;
            }
            int cls, dcres = 0;
            uint32_t rmask = 0;
            if (ovf || (!rows && !has_dc)) {
                cls = CLS_ZERO;
            } else if (!(rows & 0xFEu) && !col) {
                cls = CLS_DC;
                dcres = round_residual_dc((float)(has_dc ? dcv : v00));
            } else {
                cls = col ? CLS_FULL : CLS_VERT;
                rmask = rows | (has_dc ? 1u : 0u);
            }
            S.meta[p1_mb * 6 + p1_b] = pack_meta(cls, rmask, (uint32_t)slot, dcres);
            if (p1_nev > 0) S.slotmeta[slot] = rmask | ((uint32_t)cls << 8);
        }
        __syncthreads();

        // ================= phase 2: IDCT, 4 slots per warp, 8 lanes per slot =====================
        for (int g = warp; g * 4 < n_slots; g += TILE_THREADS / 32) {
            const int s = g * 4 + (lane >> 3), t = lane & 7;
            uint32_t sm = s < n_slots ? S.slotmeta[s] : 0u;
            const int cls = (int)(sm >> 8);
            const bool need = cls == CLS_FULL || cls == CLS_VERT;
            const bool vert = cls == CLS_VERT;
            const uint32_t R = need ? (sm & 0xFFu) : 0u;
            uint32_t U = R;
            U |= __shfl_xor_sync(FULL, U, 8);
            U |= __shfl_xor_sync(FULL, U, 16);
            float* c = S.pool + min(s, SLOT_CAP - 1) * SLOT_FLOATS;
            // row pass: t[y][i] = sum_x c[y][x] * B[x][i] in ascending x (idct_1d); lane t = i
            float tv[8];
#pragma unroll
            for (int y = 0; y < 8; y++) {
                tv[y] = 0.0f;
                if ((U >> y) & 1u) {  // warp-uniform
                    if ((R >> y) & 1u) {
                        if (vert) {
                            tv[y] = c[y * 8];  // Vert: the first column feeds idct_1d directly (idct.rs:152-153)
                        } else {
                            const float4 ca = *reinterpret_cast<const float4*>(c + y * 8);
                            const float4 cc = *reinterpret_cast<const float4*>(c + y * 8 + 4);
                            float a = fadd(0.0f, fmul(ca.x, bt[0]));
                            a = fadd(a, fmul(ca.y, bt[1]));
                            a = fadd(a, fmul(ca.z, bt[2]));
                            a = fadd(a, fmul(ca.w, bt[3]));
                            a = fadd(a, fmul(cc.x, bt[4]));
                            a = fadd(a, fmul(cc.y, bt[5]));
                            a = fadd(a, fmul(cc.z, bt[6]));
                            a = fadd(a, fmul(cc.w, bt[7]));
                            tv[y] = a;
                        }
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int y = 0; y < 8; y++)
                if ((R >> y) & 1u) c[y * 8 + t] = tv[y];
            __syncwarp();
            // column pass: out[i][j] = sum_y t[y][i] * B[y][j] in ascending y; lane t = j (pixel row)
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int y = 0; y < 8; y++) {
                if ((U >> y) & 1u) {
                    if ((R >> y) & 1u) {
                        const float4 ta = *reinterpret_cast<const float4*>(c + y * 8);
                        const float4 tb = *reinterpret_cast<const float4*>(c + y * 8 + 4);
                        acc[0] = fadd(acc[0], fmul(ta.x, bt[y]));
                        acc[1] = fadd(acc[1], fmul(ta.y, bt[y]));
                        acc[2] = fadd(acc[2], fmul(ta.z, bt[y]));
                        acc[3] = fadd(acc[3], fmul(ta.w, bt[y]));
                        acc[4] = fadd(acc[4], fmul(tb.x, bt[y]));
                        acc[5] = fadd(acc[5], fmul(tb.y, bt[y]));
                        acc[6] = fadd(acc[6], fmul(tb.z, bt[y]));
                        acc[7] = fadd(acc[7], fmul(tb.w, bt[y]));
                    }
                }
            }
            int rr[8];
#pragma unroll
            for (int i = 0; i < 8; i++) rr[i] = vert ? round_residual_scaled(acc[i]) : round_residual(acc[i]);
            __syncwarp();
            if (need) {
                uint4 o;
                o.x = ((uint32_t)rr[0] & 0xFFFFu) | ((uint32_t)rr[1] << 16);
                o.y = ((uint32_t)rr[2] & 0xFFFFu) | ((uint32_t)rr[3] << 16);
                o.z = ((uint32_t)rr[4] & 0xFFFFu) | ((uint32_t)rr[5] << 16);
                o.w = ((uint32_t)rr[6] & 0xFFFFu) | ((uint32_t)rr[7] << 16);
                *reinterpret_cast<uint4*>(c + t * 8) = o;  // residual row t: first 16 bytes of slot row t
            }
        }
        __syncthreads();

        // ================= phase 3: MC + add + clamp + stores + RGBA, all in registers ===========
        const int n_units = (m1 - m0) * 16;  // unit = (macroblock, row pair q, half h): 16 luma x 2 rows
        for (int u = tid; u < n_units; u += TILE_THREADS) {
            // consecutive threads: h, then macroblock, then q -> neighbouring stores coalesce
            const int nmb = m1 - m0, nmb2 = nmb * 2;
            const int q = (nmb2 & (nmb2 - 1)) == 0 ? (u >> (__ffs(nmb2) - 1)) : (u / nmb2);  // uniform branch
            const int rem = u - q * nmb2;
            const int mbi = m0 + (rem >> 1), h = rem & 1;
            const uint32_t* r = &S.mbrec[mbi * 6];
            const uint32_t w1 = r[1], w2 = r[2], w4 = r[4], w5 = r[5];
            const PicDev& P = pics[w1 & 0xFFFFu];
            const int mbx = (w1 >> 16) & 0xFF, mby = w1 >> 24;
            const bool inter = (w2 & H263CU_MB_INTER) != 0;
            const int pitch_y = P.pitch_y, pitch_c = P.pitch_c;
            const int x0 = mbx * 16 + h * 8, y0 = mby * 16 + q * 2;
            const int cx0 = mbx * 8 + h * 4, cy0 = mby * 8 + q;
            const int lb = ((q >> 2) << 1) | h;  // luma block of this unit

            uint32_t y00 = 0, y01 = 0, y10 = 0, y11 = 0, cbw = 0, crw = 0;  // predictions
            if (inter) {
                const uint32_t mvw = lb < 2 ? (w4 >> (16 * lb)) : (w5 >> (16 * (lb - 2)));
                const int mvx = (int8_t)(mvw & 0xFF), mvy = (int8_t)((mvw >> 8) & 0xFF);
                const int m0x = (int8_t)byte_of(w4, 0), m0y = (int8_t)byte_of(w4, 1), m1x = (int8_t)byte_of(w4, 2),
                          m1y = (int8_t)byte_of(w4, 3), m2x = (int8_t)byte_of(w5, 0), m2y = (int8_t)byte_of(w5, 1),
                          m3x = (int8_t)byte_of(w5, 2), m3y = (int8_t)byte_of(w5, 3);
                const int cvx = average_sum_of_mvs(m0x + m1x + m2x + m3x), cvy = average_sum_of_mvs(m0y + m1y + m2y + m3y);
                const bool in_range = mvx >= -32 && mvx <= 31 && mvy >= -32 && mvy <= 31 && cvx >= -16 && cvx <= 15 &&
                                      cvy >= -16 && cvy <= 15;
                if (in_range) {
                    {
                        const int dx = mvx >> 1, ix = mvx & 1, dy = mvy >> 1, iy = mvy & 1;
                        const int sx = x0 + dx, sy = y0 + dy;
                        const int a = sx & 3, sh = a * 8, shb = sh + 8 * ix;
                        const uint8_t* base = P.ref[0] + (ptrdiff_t)sy * pitch_y + (sx - a);
                        const uint32_t* l0 = reinterpret_cast<const uint32_t*>(base);
                        const uint32_t* l1 = reinterpret_cast<const uint32_t*>(base + pitch_y);
                        const uint32_t* l2 = reinterpret_cast<const uint32_t*>(base + (1 + iy) * pitch_y);
                        const uint32_t a0 = __ldg(l0), a1 = __ldg(l0 + 1), a2 = __ldg(l0 + 2);
                        const uint32_t b0 = __ldg(l1), b1 = __ldg(l1 + 1), b2 = __ldg(l1 + 2);
                        const uint32_t c0 = __ldg(l2), c1 = __ldg(l2 + 1), c2 = __ldg(l2 + 2);
                        // second row of output row 0: row sy + iy
                        const uint32_t s0 = iy ? b0 : a0, s1 = iy ? b1 : a1, s2 = iy ? b2 : a2;
                        y00 = mc_word(a0, a1, s0, s1, sh, shb);
                        y01 = mc_word(a1, a2, s1, s2, sh, shb);
                        y10 = mc_word(b0, b1, c0, c1, sh, shb);
                        y11 = mc_word(b1, b2, c1, c2, sh, shb);
                    }
                    {
                        const int dx = cvx >> 1, ix = cvx & 1, dy = cvy >> 1, iy = cvy & 1;
                        const int sx = cx0 + dx, sy = cy0 + dy;
                        const int a = sx & 3, sh = a * 8, shb = sh + 8 * ix;
                        const ptrdiff_t off = (ptrdiff_t)sy * pitch_c + (sx - a);
                        const uint32_t* b0p = reinterpret_cast<const uint32_t*>(P.ref[1] + off);
                        const uint32_t* b1p = reinterpret_cast<const uint32_t*>(P.ref[1] + off + iy * pitch_c);
                        const uint32_t* r0p = reinterpret_cast<const uint32_t*>(P.ref[2] + off);
                        const uint32_t* r1p = reinterpret_cast<const uint32_t*>(P.ref[2] + off + iy * pitch_c);
                        cbw = mc_word(__ldg(b0p), __ldg(b0p + 1), __ldg(b1p), __ldg(b1p + 1), sh, shb);
                        crw = mc_word(__ldg(r0p), __ldg(r0p + 1), __ldg(r1p), __ldg(r1p + 1), sh, shb);
                    }
                } else {
                    // vectors beyond the baseline range: clamped per-sample fetch (generic path)
                    uint32_t t0, t1;
                    mc_fetch8(P.ref[0], pitch_y, P.w, P.h, x0, y0, mvx, mvy, y00, y01);
                    mc_fetch8(P.ref[0], pitch_y, P.w, P.h, x0, y0 + 1, mvx, mvy, y10, y11);
                    mc_fetch8(P.ref[1], pitch_c, P.cw, P.ch, mbx * 8, cy0, cvx, cvy, t0, t1);
                    cbw = h ? t1 : t0;
                    mc_fetch8(P.ref[2], pitch_c, P.cw, P.ch, mbx * 8, cy0, cvx, cvy, t0, t1);
                    crw = h ? t1 : t0;
                }
            }

            // ---- residuals ----
            {
                const uint32_t m = S.meta[mbi * 6 + lb];
                const int cls = (int)(m & 7u);
                if (cls == CLS_DC) {
                    const uint32_t d = (uint32_t)(((int)(m << 3)) >> 22) & 0xFFFFu;
                    const uint32_t dd = d | (d << 16);
                    y00 = add_res4(y00, dd, dd);
                    y01 = add_res4(y01, dd, dd);
                    y10 = add_res4(y10, dd, dd);
                    y11 = add_res4(y11, dd, dd);
                } else if (cls != CLS_ZERO) {
                    const float* c = S.pool + ((m >> 11) & 0xFFu) * SLOT_FLOATS + ((q * 2) & 7) * 8;
                    const uint4 ra = *reinterpret_cast<const uint4*>(c);
                    const uint4 rb = *reinterpret_cast<const uint4*>(c + 8);
                    y00 = add_res4(y00, ra.x, ra.y);
                    y01 = add_res4(y01, ra.z, ra.w);
                    y10 = add_res4(y10, rb.x, rb.y);
                    y11 = add_res4(y11, rb.z, rb.w);
                }
            }
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
                const uint32_t m = S.meta[mbi * 6 + 4 + pl];
                const int cls = (int)(m & 7u);
                uint32_t v = pl ? crw : cbw;
                if (cls == CLS_DC) {
                    const uint32_t d = (uint32_t)(((int)(m << 3)) >> 22) & 0xFFFFu;
                    const uint32_t dd = d | (d << 16);
                    v = add_res4(v, dd, dd);
                } else if (cls != CLS_ZERO) {
                    const float* c = S.pool + ((m >> 11) & 0xFFu) * SLOT_FLOATS + q * 8;
                    const uint2 rv = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(c) + h * 2);
                    v = add_res4(v, rv.x, rv.y);
                }
                if (pl) crw = v; else cbw = v;
            }

            // ---- plane stores (+ border replication for the next picture's prediction) ----
            uint8_t* py = P.cur[0] + (ptrdiff_t)y0 * pitch_y + x0;
            *reinterpret_cast<uint2*>(py) = make_uint2(y00, y01);
            *reinterpret_cast<uint2*>(py + pitch_y) = make_uint2(y10, y11);
            uint8_t* pcb = P.cur[1] + (ptrdiff_t)cy0 * pitch_c + cx0;
            uint8_t* pcr = P.cur[2] + (ptrdiff_t)cy0 * pitch_c + cx0;
            *reinterpret_cast<uint32_t*>(pcb) = cbw;
            *reinterpret_cast<uint32_t*>(pcr) = crw;
            const int mbw = P.w >> 4, mbh = P.h >> 4;
            const bool e_left = mbx == 0 && h == 0, e_right = mbx == mbw - 1 && h == 1;
            const bool e_top = mby == 0 && q == 0, e_bot = mby == mbh - 1 && q == 7;
            if (e_left | e_right | e_top | e_bot) {
                if (e_left | e_right) {
                    // 16 luma / 8 chroma pixels of horizontal extension for this unit's rows
                    const uint32_t l0 = e_left ? __byte_perm(y00, 0, 0x0000) : __byte_perm(y01, 0, 0x3333);
                    const uint32_t l1 = e_left ? __byte_perm(y10, 0, 0x0000) : __byte_perm(y11, 0, 0x3333);
                    uint8_t* q0 = e_left ? py - 16 : py + 8;
                    *reinterpret_cast<uint4*>(q0) = make_uint4(l0, l0, l0, l0);
                    *reinterpret_cast<uint4*>(q0 + pitch_y) = make_uint4(l1, l1, l1, l1);
                    const uint32_t cbe = e_left ? __byte_perm(cbw, 0, 0x0000) : __byte_perm(cbw, 0, 0x3333);
                    const uint32_t cre = e_left ? __byte_perm(crw, 0, 0x0000) : __byte_perm(crw, 0, 0x3333);
                    const int co = e_left ? -8 : 4;
                    *reinterpret_cast<uint2*>(pcb + co) = make_uint2(cbe, cbe);
                    *reinterpret_cast<uint2*>(pcr + co) = make_uint2(cre, cre);
                }
                if (e_top | e_bot) {
                    // vertical extension: 16 luma / 8 chroma rows above row 0 or below the last row,
                    // including the corner when the unit also sits on a vertical edge
                    const uint32_t v0 = e_top ? y00 : y10, v1 = e_top ? y01 : y11;
                    uint8_t* rowp = e_top ? py : py + pitch_y;
                    const int dir = e_top ? -pitch_y : pitch_y;
                    const uint32_t corner = e_left ? __byte_perm(v0, 0, 0x0000) : __byte_perm(v1, 0, 0x3333);
                    for (int k = 1; k <= 16; k++) {
                        uint8_t* d = rowp + (ptrdiff_t)k * dir;
                        *reinterpret_cast<uint2*>(d) = make_uint2(v0, v1);
                        if (e_left) *reinterpret_cast<uint4*>(d - 16) = make_uint4(corner, corner, corner, corner);
                        if (e_right) *reinterpret_cast<uint4*>(d + 8) = make_uint4(corner, corner, corner, corner);
                    }
                    const int cdir = e_top ? -pitch_c : pitch_c;
                    const uint32_t cbc = e_left ? __byte_perm(cbw, 0, 0x0000) : __byte_perm(cbw, 0, 0x3333);
                    const uint32_t crc = e_left ? __byte_perm(crw, 0, 0x0000) : __byte_perm(crw, 0, 0x3333);
                    const int co = e_left ? -8 : 4;
                    for (int k = 1; k <= 8; k++) {
                        uint8_t* db = pcb + (ptrdiff_t)k * cdir;
                        uint8_t* dr = pcr + (ptrdiff_t)k * cdir;
                        *reinterpret_cast<uint32_t*>(db) = cbw;
                        *reinterpret_cast<uint32_t*>(dr) = crw;
                        if (e_left | e_right) {
                            *reinterpret_cast<uint2*>(db + co) = make_uint2(cbc, cbc);
                            *reinterpret_cast<uint2*>(dr + co) = make_uint2(crc, crc);
                        }
                    }
                }
            }

            // ---- BT.601 RGBA (bt601.rs:12-59): 16 pixels, four 128-bit stores ----
            if (emit_rgba && P.rgba) {
                uint8_t* o = P.rgba + (size_t)y0 * P.rgba_pitch + (size_t)x0 * 4;
                *reinterpret_cast<uint4*>(o) = rgba4(y00, cbw, crw);
                *reinterpret_cast<uint4*>(o + 16) = rgba4(y01, cbw >> 16, crw >> 16);
                *reinterpret_cast<uint4*>(o + P.rgba_pitch) = rgba4(y10, cbw, crw);
                *reinterpret_cast<uint4*>(o + P.rgba_pitch + 16) = rgba4(y11, cbw >> 16, crw >> 16);
            }
        }
        __syncthreads();  // the pool is reused by the next pass
    }
}

void launch_recon(const PicDev* pics, const h263cu_mb* mbs, const h263cu_event* events, uint32_t n_mbs, int emit_rgba,
                  int tiled, cudaStream_t stream) {
    if (n_mbs == 0) return;
    if (tiled) {
        const uint32_t grid = (n_mbs + TILE_MBS - 1) / TILE_MBS;
        recon_tile_kernel<<<grid, TILE_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba);
    } else {
        const uint32_t grid = (n_mbs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        recon_mb_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba);
    }
}

// =======================================================================================
// Deblocking post-filter on a tile held in shared memory.
//   region = (TW + 4) x (TH + 4) pixels starting at plane position (ox - 2, oy - 2), stored
//   with row stride RS at column offset CO (so that the tile interior is 4-byte aligned).
//   Every output pixel depends only on input pixels at most 2 rows / columns away across an
//   8-aligned edge, so tiles whose origin is a multiple of 8 are independent given the halo.
//   Horizontal edges first over the whole plane, then vertical edges (deblock.rs:305-315).
// =======================================================================================
template <int TW, int TH, int RS, int CO>
__device__ __forceinline__ void deblock_tile(uint8_t* sm, const uint8_t* __restrict__ src, int pitch, int W, int H,
                                             int ox, int oy, int strength, int tid, int nthreads) {
    constexpr int RW = TW + 4, RH = TH + 4;
    for (int idx = tid; idx < RW * RH; idx += nthreads) {
        const int rx = idx % RW, ry = idx / RW;
        const int gx = min(max(ox - 2 + rx, 0), W - 1), gy = min(max(oy - 2 + ry, 0), H - 1);
        sm[ry * RS + CO + rx] = src[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    // horizontal edges: rows ey-2 .. ey+1 for ey = 8, 16, ... <= H - 2 (deblock.rs:136-181);
    // columns below 8*floor(W/8) use the SIMD (floor) arithmetic, the rest the scalar one.
    constexpr int NEH = TH / 8 + 1;
    const int simd_cols = (W >> 3) << 3;
    for (int item = tid; item < NEH * RW; item += nthreads) {
        const int e = item / RW, rx = item % RW;
        const int ey = oy + 8 * e, gx = ox - 2 + rx;
        if (ey < 8 || ey > H - 2 || gx < 0 || gx >= W) continue;
        uint8_t* p = sm + (8 * e) * RS + CO + rx;  // region row of sample A = 8e + 2 - 2
        int A = p[0], B = p[RS], C = p[2 * RS], D = p[3 * RS];
        deblock_process(A, B, C, D, strength, gx >= simd_cols);
        p[0] = (uint8_t)A, p[RS] = (uint8_t)B, p[2 * RS] = (uint8_t)C, p[3 * RS] = (uint8_t)D;
    }
    __syncthreads();
    // vertical edges: columns ex-2 .. ex+1 for ex = 8, 16, ... with ex + 2 <= W, only when
    // W >= 10 (deblock.rs:185-299); rows below 8*floor(H/8) use the SIMD arithmetic.
    constexpr int NEV = TW / 8 + 1;
    const int simd_rows = (H >> 3) << 3;
    for (int item = tid; item < NEV * TH; item += nthreads) {
        const int e = item / TH, ty = item % TH;
        const int ex = ox + 8 * e, gy = oy + ty;
        if (W < 10 || ex < 8 || ex + 2 > W || gy >= H) continue;
        uint8_t* p = sm + (ty + 2) * RS + CO + 8 * e;  // region column of sample A = 8e + 2 - 2
        int A = p[0], B = p[1], C = p[2], D = p[3];
        deblock_process(A, B, C, D, strength, gy >= simd_rows);
        p[0] = (uint8_t)A, p[1] = (uint8_t)B, p[2] = (uint8_t)C, p[3] = (uint8_t)D;
    }
    __syncthreads();
}

// Fused deblock (Y, Cb, Cr) + RGBA for one 32x32 luma tile of one picture.
__global__ void __launch_bounds__(256) deblock_rgba_kernel(const PicDev* __restrict__ pics) {
    __shared__ __align__(16) uint8_t sy[36 * 40];
    __shared__ __align__(16) uint8_t scb[20 * 24];
    __shared__ __align__(16) uint8_t scr[20 * 24];
    const PicDev& P = pics[blockIdx.y];
    const int W = P.w, H = P.h;
    const int tiles_x = (W + 31) >> 5, tiles_y = (H + 31) >> 5;
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int tid = threadIdx.x;
    const int strength = P.strength;
    deblock_tile<32, 32, 40, 2>(sy, P.cur[0], P.pitch_y, W, H, tx * 32, ty * 32, strength, tid, 256);
    deblock_tile<16, 16, 24, 2>(scb, P.cur[1], P.pitch_c, P.cw, P.ch, tx * 16, ty * 16, strength, tid, 256);
    deblock_tile<16, 16, 24, 2>(scr, P.cur[2], P.pitch_c, P.cw, P.ch, tx * 16, ty * 16, strength, tid, 256);
    if (!P.rgba) return;
    const int row = tid >> 3, xq = tid & 7;
    const int gx = tx * 32 + xq * 4, gy = ty * 32 + row;
    if (gx >= W || gy >= H) return;
    const uint32_t yw = *reinterpret_cast<const uint32_t*>(&sy[(row + 2) * 40 + 4 + xq * 4]);
    const uint32_t cbp = *reinterpret_cast<const uint16_t*>(&scb[((row >> 1) + 2) * 24 + 4 + xq * 2]);
    const uint32_t crp = *reinterpret_cast<const uint16_t*>(&scr[((row >> 1) + 2) * 24 + 4 + xq * 2]);
    const ChromaTerms t0 = chroma_terms((int)(cbp & 0xFF), (int)(crp & 0xFF));
    const ChromaTerms t1 = chroma_terms((int)(cbp >> 8), (int)(crp >> 8));
    uint4 o;
    o.x = yuv_pixel((int)byte_of(yw, 0), t0);
    o.y = yuv_pixel((int)byte_of(yw, 1), t0);
    o.z = yuv_pixel((int)byte_of(yw, 2), t1);
    o.w = yuv_pixel((int)byte_of(yw, 3), t1);
    *reinterpret_cast<uint4*>(P.rgba + (size_t)gy * P.rgba_pitch + (size_t)gx * 4) = o;
}

void launch_deblock_rgba(const PicDev* pics, uint32_t n_pics, uint32_t max_w, uint32_t max_h, cudaStream_t stream) {
    if (n_pics == 0) return;
    dim3 grid(((max_w + 31) / 32) * ((max_h + 31) / 32), n_pics);
    deblock_rgba_kernel<<<grid, 256, 0, stream>>>(pics);
}

// ---- stateless drop-ins ---------------------------------------------------------------
// yuv420_to_rgba (bt601.rs:105-196) on tight planes of any size: pixel (x, y) uses the
// chroma sample (x >> 1, y >> 1) of a ceil(w/2)-wide chroma plane (bt601.rs:115,133,182).
__global__ void __launch_bounds__(256)
    yuv420_to_rgba_kernel(const uint8_t* __restrict__ y, const uint8_t* __restrict__ cb, const uint8_t* __restrict__ cr,
                          uint32_t w, uint32_t h, uint32_t* __restrict__ rgba) {
    const uint32_t cw = (w + 1) >> 1;
    const size_t n = (size_t)w * h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t px = (uint32_t)(i % w), py = (uint32_t)(i / w);
        const size_t ci = (size_t)(py >> 1) * cw + (px >> 1);
        rgba[i] = yuv_pixel(y[i], chroma_terms(cb[ci], cr[ci]));
    }
}

void launch_yuv420_to_rgba(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, uint32_t w, uint32_t h, uint8_t* rgba,
                           cudaStream_t stream) {
    const size_t n = (size_t)w * h;
    if (n == 0) return;
    const uint32_t grid = (uint32_t)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
    yuv420_to_rgba_kernel<<<grid, 256, 0, stream>>>(y, cb, cr, w, h, reinterpret_cast<uint32_t*>(rgba));
}

__global__ void __launch_bounds__(256)
    deblock_plane_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int H, int strength) {
    __shared__ __align__(16) uint8_t sm[36 * 40];
    const int tiles_x = (W + 31) >> 5;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    deblock_tile<32, 32, 40, 2>(sm, in, W, W, H, tx * 32, ty * 32, strength, threadIdx.x, 256);
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 256) {
        const int rx = idx & 31, ry = idx >> 5;
        const int gx = tx * 32 + rx, gy = ty * 32 + ry;
        if (gx < W && gy < H) out[(size_t)gy * W + gx] = sm[(ry + 2) * 40 + 4 + rx];
    }
}

void launch_deblock_plane(const uint8_t* in, uint8_t* out, uint32_t w, uint32_t h, int strength, cudaStream_t stream) {
    if (w == 0 || h == 0) return;
    const uint32_t grid = ((w + 31) / 32) * ((h + 31) / 32);
    deblock_plane_kernel<<<grid, 256, 0, stream>>>(in, out, (int)w, (int)h, strength);
}

// ---- checksums --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) checksum_kernel(const ChecksumJob* __restrict__ jobs, unsigned long long* out) {
    const ChecksumJob J = jobs[blockIdx.y];
    const uint64_t n = (uint64_t)J.row_bytes * J.rows;
    unsigned long long s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % J.row_bytes), yy = (uint32_t)(i / J.row_bytes);
        const uint32_t v = J.base[(size_t)yy * J.pitch + x];
        s += (unsigned long long)(v + 1u) * (unsigned long long)(((uint32_t)i * 2654435761u) | 1u);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(FULL, s, d);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; k++) t += ws[k];
        atomicAdd(out + J.out_index, t);
    }
}

void launch_checksums(const ChecksumJob* jobs, uint32_t n_jobs, unsigned long long* out, cudaStream_t stream) {
    if (n_jobs == 0) return;
    dim3 grid(8, n_jobs);
    checksum_kernel<<<grid, 256, 0, stream>>>(jobs, out);
}

}  // namespace h263dev
