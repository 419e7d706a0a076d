// recon_tile.cu -- the hot kernel: fused reconstruction of a tile of 32 consecutive macroblocks
// per CTA (256 threads), for pictures whose size is a multiple of 16 and whose reference planes
// carry the replicated border.
//
//   phase 0  one thread per macroblock: motion vectors -> source offsets / alignment / half-pel
//            flags, destination offsets, edge flags (gather.rs:140-204, types.rs:721-729,759-768);
//            one thread per block: coded blocks are compacted into coefficient slots
//   phase 1  one thread per CODED block (slot): walk the run/level events, accumulate the zig-zag
//            index, dequantise, scatter into the slot, classify (rle.rs:82-172)
//   phase 2  8 lanes per slot, 4 slots per warp: row pass over the rows that hold a coefficient,
//            column pass, rounding (idct.rs:52-65,170-198); packed s16 residuals overwrite the slot
//   phase 3  one thread per 8x4 luma pixels + the 4x2 chroma pixels under them: half-pel
//            interpolation in 16-bit lanes (gather.rs:34-40,103-113), saturating residual add,
//            plane stores, border replication, BT.601 RGBA (bt601.rs:12-59) with 256-bit stores.
//
// Everything a thread needs per macroblock sits in shared memory as 32-bit offsets from the
// context's pool bases (kernel parameters), so the epilogue does no 64-bit pointer chasing.
// Bound: HBM bandwidth with the integer/FP32 issue rate as the secondary ceiling -- this file is
// written for instruction count: see DESIGN.md section 4 for the per-phase budget.
#include "recon_common.cuh"

namespace h263dev {

namespace {

constexpr int TILE_MBS = 32;
constexpr int TILE_THREADS = 256;
constexpr int SLOT_CAP = 96;     // coefficient slots per pass (= 16 MBs x 6 blocks worst case)
constexpr int SLOT_FLOATS = 68;  // 64 + 4 pad: 16 B aligned, consecutive slots start 4 banks apart

// MbDesc.flags
constexpr uint32_t MBF_INTER = 1u << 20;
constexpr uint32_t MBF_SLOW = 1u << 21;  // a vector leaves the replicated border: clamped per-sample path
constexpr uint32_t MBF_LEFT = 1u << 22, MBF_RIGHT = 1u << 23, MBF_TOP = 1u << 24, MBF_BOTTOM = 1u << 25;
constexpr uint32_t MBF_RGBA = 1u << 26;

struct __align__(16) MbDesc {  // 48 bytes
    uint32_t ysrc[4];   // per luma block: 4-byte offset from y_pool of the aligned word that holds the
                        // first source pixel of the block's row 0
    uint32_t csrc;      // same for the chroma block (offset from cb_pool / cr_pool)
    uint32_t ydst;      // 4-byte offset from y_pool of the macroblock's top-left luma pixel
    uint32_t cdst;      // 4-byte offset from cb_pool / cr_pool of its top-left chroma pixel
    uint32_t rgba;      // 16-byte offset from rgba_pool of its top-left RGBA pixel
    uint32_t flags;     // per luma block k bits [4k, 4k+4): align(2) | ix | iy ; chroma bits [16, 20); MBF_*
    uint32_t pitches;   // pitch_y | pitch_c << 16 (bytes)
    uint32_t rgba_pitch;
    uint32_t pic;
};

struct TileSmem {
    float pool[SLOT_CAP * SLOT_FLOATS];  // coefficient slots -> row-pass output -> residuals (in place)
    MbDesc mbd[TILE_MBS];
    uint32_t mbrec[TILE_MBS * 6];  // the tile's macroblock records
    uint32_t meta[TILE_MBS * 6];   // per block: cls[2:0] | slot[10:3] | dcres[31:16]
    uint2 slotdesc[SLOT_CAP];      // x = first event unit (absolute), y = nev | quant<<8 | wide<<13 | inter<<14 | block<<16 | dc<<24
    uint32_t slotmeta[SLOT_CAP];   // rows | cls << 8
    float basis[64];
    uint32_t warp_count[2][8];     // coded blocks per warp of block threads, per half of the tile
    uint8_t order[SLOT_CAP];       // slots sorted by the number of coefficient rows (per warp of slots)
    uint8_t dezigzag[64];
};

// H.263 dequantisation of a narrow (10-bit) level: |level| <= 512, so QP*(2|level|+1) stays below
// 2^15 and the wrapping-i16 arithmetic of rle.rs:130-133 cannot wrap. q2 = 2*QP, qc = QP - (QP even).
__device__ __forceinline__ int dequant_narrow(int level, int q2, int qc) {
    const int mag = q2 * abs(level) + qc;
    const int v = level < 0 ? -mag : mag;
    return max(min(v, 2047), -2048);
}

// rounding of the (already /4) transform output q: trunc(q * m + copysign(0.5, q)) with m = 1 for
// Full / Horiz blocks (x * 1 == x) and m = BASIS_TABLE[0][0] for Vert blocks, where
// (v * B00) / 4 == (v / 4) * B00 exactly (power-of-two scaling) -- idct.rs:143-145,161-163,189-190
__device__ __forceinline__ int round_q(float q, float m) { return __float2int_rz(fadd(fmul(q, m), copysign_half(q))); }
// (hi, lo) -> two s16 lanes, each `as i16` saturated and clamped to [-256, 255] (idct.rs:189-190)
__device__ __forceinline__ uint32_t pack_clamp_s16x2(int hi, int lo) {
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    asm("min.s16x2 %0, %0, %1;" : "+r"(d) : "r"(0x00FF00FFu));
    asm("max.s16x2 %0, %0, %1;" : "+r"(d) : "r"(0xFF00FF00u));
    return d;
}

// ---- half-pel interpolation in 16-bit lanes -------------------------------------------------
// One row of 8 pixels: a = bytes [s, s+8), b = bytes [s+ix, s+ix+8) of the 12 loaded bytes.
// Returns a + b per pixel as four words of two 16-bit lanes: e0 = (p0, p2), o0 = (p1, p3),
// e1 = (p4, p6), o1 = (p5, p7).  With ix = 0 this is 2a.
struct RowSum {
    uint32_t e0, o0, e1, o1;
};
__device__ __forceinline__ RowSum row_sum8(uint32_t w0, uint32_t w1, uint32_t w2, int sh, int shb) {
    const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
    const uint32_t b0 = __funnelshift_rc(w0, w1, shb), b1 = __funnelshift_rc(w1, w2, shb);
    RowSum h;
    h.e0 = __byte_perm(a0, 0, 0x4240) + __byte_perm(b0, 0, 0x4240);
    h.o0 = __byte_perm(a0, 0, 0x4341) + __byte_perm(b0, 0, 0x4341);
    h.e1 = __byte_perm(a1, 0, 0x4240) + __byte_perm(b1, 0, 0x4240);
    h.o1 = __byte_perm(a1, 0, 0x4341) + __byte_perm(b1, 0, 0x4341);
    return h;
}
// (top + bottom + 2) >> 2 per lane.  With bottom == top and ix == 0 this is exactly the pixel,
// with one of them it is (a + b + 1) >> 1: one form covers the four modes of gather.rs:34-40,103-113.
// wt = 2 - iy, wb = iy select "bottom == top" without a select, on the multiplier pipe.
__device__ __forceinline__ uint32_t vmix(uint32_t top, uint32_t bottom, uint32_t wt, uint32_t wb) {
    return ((top * wt + (bottom * wb + 0x00020002u)) >> 2) & 0x00FF00FFu;
}

// ---- BT.601 (bt601.rs:12-59), two instructions of clamp + pack per pixel ------------------
__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c) {
    uint32_t d;  // d = c[15:0] << 16 | sat_u8(a) << 8 | sat_u8(b)
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
struct CT {
    int r, g, b;
};
// chroma terms with the -16 luma offset and the rounding constant folded in:
//   R = (76309*y + r) >> 16 with r = 104597*(cr-128) + 32768 - 16*76309, etc.
__device__ __forceinline__ CT chroma_terms_folded(int cb, int cr) {
    CT t;
    t.r = cr * 104597 + (32768 - 128 * 104597 - 16 * 76309);
    t.g = cr * -53279 + (cb * -25675 + (32768 + 128 * 53279 + 128 * 25675 - 16 * 76309));
    t.b = cb * 132201 + (32768 - 128 * 132201 - 16 * 76309);
    return t;
}
__device__ __forceinline__ uint32_t rgba_px(int y, const CT& t) {
    const int r = (y * 76309 + t.r) >> 16, g = (y * 76309 + t.g) >> 16, b = (y * 76309 + t.b) >> 16;
    return pack_sat(g, r, pack_sat(255, b, 0));
}

__device__ __forceinline__ void st_global_v8(uint8_t* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ uint32_t splat_lo(uint32_t w) { return __byte_perm(w, 0, 0x0000); }
__device__ __forceinline__ uint32_t splat_hi(uint32_t w) { return __byte_perm(w, 0, 0x3333); }

}  // namespace

__global__ void __launch_bounds__(TILE_THREADS, 4)
    recon_tile_kernel(const PicDev* __restrict__ pics, const h263cu_mb* __restrict__ mbs,
                      const h263cu_event* __restrict__ events, uint32_t n_mbs, int emit_rgba, const Pools pools) {
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

    // ================= phase 0a: one thread per block (tid < 192): event counts ==================
    // block threads of macroblocks 0..15 sit in warps 0..2, of macroblocks 16..31 in warps 3..5
    const int p1_mb = tid / 6, p1_b = tid - p1_mb * 6;
    const bool p1_valid = p1_mb < n_tile;
    uint32_t p1_nev = 0;
    if (p1_valid) {
        const uint32_t* r = &S.mbrec[p1_mb * 6];
        p1_nev = p1_b < 2 ? (r[2] >> (16 + 8 * p1_b)) & 0xFF : (r[3] >> (8 * (p1_b - 2))) & 0xFF;
    }
    const uint32_t p1_bal = __ballot_sync(FULL, p1_nev > 0);
    if (lane == 0) S.warp_count[0][warp] = __popc(p1_bal);
    __syncthreads();
    int slot = __popc(p1_bal & ((1u << lane) - 1u));  // slot of this block thread when the tile runs in one pass
    int slots_half0 = 0, slots_half1 = 0;
#pragma unroll
    for (int w = 0; w < 6; w++) {
        const int c = (int)S.warp_count[0][w];
        if (w < warp) slot += c;
        if (w < 3) slots_half0 += c; else slots_half1 += c;
    }
    // one pass over 32 macroblocks when their coded blocks fit the slot pool, else two passes of 16
    const int n_pass = slots_half0 + slots_half1 <= SLOT_CAP ? 1 : 2;
    const int mb_per_pass = n_pass == 1 ? TILE_MBS : TILE_MBS / 2;
    const int unit_shift = n_pass == 1 ? 6 : 5;

    const float bt0 = S.basis[0 * 8 + (lane & 7)], bt1 = S.basis[1 * 8 + (lane & 7)], bt2 = S.basis[2 * 8 + (lane & 7)],
                bt3 = S.basis[3 * 8 + (lane & 7)], bt4 = S.basis[4 * 8 + (lane & 7)], bt5 = S.basis[5 * 8 + (lane & 7)],
                bt6 = S.basis[6 * 8 + (lane & 7)], bt7 = S.basis[7 * 8 + (lane & 7)];

    for (int pass = 0; pass < n_pass; pass++) {
        const int m0 = pass * mb_per_pass, m1 = min(n_tile, m0 + mb_per_pass);
        if (m0 >= m1) break;  // uniform
        const int n_slots = n_pass == 1 ? slots_half0 + slots_half1 : (pass == 0 ? slots_half0 : slots_half1);
        if (pass == 1) slot -= slots_half0;

        // ================= phase 0c: zero the slots, describe them ================================
        for (int i = tid; i < n_slots * 16; i += TILE_THREADS)
            *reinterpret_cast<float4*>(S.pool + (i >> 4) * SLOT_FLOATS + (i & 15) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p1_valid && p1_mb >= m0 && p1_mb < m1) {
            const uint32_t* r = &S.mbrec[p1_mb * 6];
            const uint32_t w1 = r[1], w2 = r[2], w3 = r[3];
            const bool inter = (w2 & H263CU_MB_INTER) != 0, wide = (w2 & H263CU_MB_WIDE) != 0;
            uint32_t code = 0;
            if (!inter) code = p1_b < 4 ? byte_of(r[4], p1_b) : byte_of(r[5], p1_b - 4);
            if (p1_nev > 0) {
                // events of the blocks before this one inside the macroblock
                const uint32_t n0 = (w2 >> 16) & 0xFF, n1 = w2 >> 24, n2 = w3 & 0xFF, n3 = (w3 >> 8) & 0xFF, n4 = (w3 >> 16) & 0xFF;
                uint32_t before = 0;
                before += p1_b > 0 ? n0 : 0;
                before += p1_b > 1 ? n1 : 0;
                before += p1_b > 2 ? n2 : 0;
                before += p1_b > 3 ? n3 : 0;
                before += p1_b > 4 ? n4 : 0;
                const uint32_t first = pics[w1 & 0xFFFFu].first_event + r[0] + (wide ? 2 * before : before);
                const uint32_t quant = (w2 >> 8) & 0xFF;
                S.slotdesc[slot] = make_uint2(first, p1_nev | ((quant & 31u) << 8) | (wide ? 1u << 13 : 0u) | (inter ? 1u << 14 : 0u) |
                                                         ((uint32_t)(p1_mb * 6 + p1_b) << 16) | (code << 24));
            } else {
                // no coefficients: Dc(level) when an intra DC is present, else Zero (rle.rs:94-104)
                const bool has_dc = code != 0;
                const int dcres = has_dc ? round_residual_dc((float)intradc_level((int)code)) : 0;
                S.meta[p1_mb * 6 + p1_b] = (uint32_t)(has_dc ? CLS_DC : CLS_ZERO) | ((uint32_t)dcres << 16);
            }
        }
        __syncthreads();

        // ================= phase 1: one thread per slot walks its events ==========================
        if (tid < ((n_slots + 31) & ~31)) {
            int key = 9;  // sort key: 8 - rows that need the transform; 9 = not a slot
            if (tid < n_slots) {
                const uint2 sd = S.slotdesc[tid];
                const int nev = (int)(sd.y & 0xFF), quant = (int)((sd.y >> 8) & 31);
                const bool wide = (sd.y >> 13) & 1u, inter = (sd.y >> 14) & 1u;
                const uint32_t blk = (sd.y >> 16) & 0xFFu, code = sd.y >> 24;
                const h263cu_event* ev = events + sd.x;
                float* cslot = S.pool + tid * SLOT_FLOATS;
                const int q2 = 2 * quant, qc = quant - 1 + (quant & 1);
                int idx = inter ? 0 : 1;  // intra: the DC occupies zig-zag index 0 (rle.rs:117-121)
                uint32_t rows = 0, colbits = 0;
                int v00 = 0;
                // No early exit: once the index passes 63 nothing is stored any more and the block is
                // classified Zero below (the whole block is dropped, DC included, rle.rs:125-127).
#pragma unroll 4
                for (int k = 0; k < nev; k++) {
                    int run, val;
                    if (!wide) {
                        const uint32_t u = __ldg(ev + k);
                        run = (int)(u >> 10);
                        val = dequant_narrow(((int)(u << 22)) >> 22, q2, qc);
                    } else {
                        run = __ldg(ev + 2 * k) & 63;
                        val = dequant((int16_t)__ldg(ev + 2 * k + 1), quant);
                    }
                    idx += run;
                    if (idx < 64) {
                        const int lin = S.dezigzag[idx];
                        cslot[lin] = (float)val;
                        if (lin == 0) v00 = val;
                        rows |= 1u << (lin >> 3);
                        colbits |= (uint32_t)lin;
                    }
                    idx += 1;
                }
                const bool ovf = idx > 64;  // some event landed on an index >= 64
                const bool col = (colbits & 7u) != 0;
                bool has_dc = false;
                int dcv = 0;
                if (!inter) {
                    has_dc = code != 0 && !ovf;
                    dcv = intradc_level((int)code);
                    if (has_dc) cslot[0] = (float)dcv;
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
                S.meta[blk] = (uint32_t)cls | ((uint32_t)tid << 3) | ((uint32_t)dcres << 16);
                S.slotmeta[tid] = rmask | ((uint32_t)cls << 8);
                key = 8 - __popc(rmask);
            }
            // counting sort of this warp's 32 slots by key (most rows first), so that the four slots a
            // warp transforms together in phase 2 have similar row counts; slots without a transform
            // (Dc / Zero) end up last
            int pos = 0;
#pragma unroll
            for (int j = 0; j < 9; j++) {
                const uint32_t b = __ballot_sync(FULL, key == j);
                if (j < key) pos += __popc(b);
                if (j == key) pos += __popc(b & ((1u << lane) - 1u));
            }
            if (key < 9) S.order[(tid & ~31) + pos] = (uint8_t)tid;
        } else if (warp == 7 && pass == 0) {
            // ============= phase 0b (in the shadow of phase 1): one thread per macroblock ===========
            const int m = lane;
            if (m < n_tile) {
                const uint32_t* r = &S.mbrec[m * 6];
                const uint32_t w1 = r[1], w2 = r[2], w4 = r[4], w5 = r[5];
                const uint32_t pic = w1 & 0xFFFFu;
                const PicDev& P = pics[pic];
                const int mbx = (w1 >> 16) & 0xFF, mby = w1 >> 24;
                const int pitch_y = P.pitch_y, pitch_c = P.pitch_c;
                const int mbw = P.w >> 4, mbh = P.h >> 4;
                MbDesc D;
                uint32_t flags = 0;
                if (w2 & H263CU_MB_INTER) {
                    flags |= MBF_INTER;
                    int sumx = 0, sumy = 0;
                    bool in_range = true;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t mvw = k < 2 ? (w4 >> (16 * k)) : (w5 >> (16 * (k - 2)));
                        const int mvx = (int8_t)(mvw & 0xFF), mvy = (int8_t)((mvw >> 8) & 0xFF);
                        sumx += mvx, sumy += mvy;
                        in_range &= mvx >= -32 && mvx <= 31 && mvy >= -32 && mvy <= 31;
                        const int sx = mbx * 16 + (k & 1) * 8 + (mvx >> 1), sy = mby * 16 + (k >> 1) * 8 + (mvy >> 1);
                        const int a = sx & 3;
                        D.ysrc[k] = P.ref_y4 + (uint32_t)((sy * pitch_y + (sx - a)) >> 2);
                        flags |= (uint32_t)(a | ((mvx & 1) << 2) | ((mvy & 1) << 3)) << (4 * k);
                    }
                    const int cvx = average_sum_of_mvs(sumx), cvy = average_sum_of_mvs(sumy);
                    in_range &= cvx >= -16 && cvx <= 15 && cvy >= -16 && cvy <= 15;
                    const int sx = mbx * 8 + (cvx >> 1), sy = mby * 8 + (cvy >> 1);
                    const int a = sx & 3;
                    D.csrc = P.ref_c4 + (uint32_t)((sy * pitch_c + (sx - a)) >> 2);
                    flags |= (uint32_t)(a | ((cvx & 1) << 2) | ((cvy & 1) << 3)) << 16;
                    if (!in_range) flags |= MBF_SLOW;
                } else {
                    D.ysrc[0] = D.ysrc[1] = D.ysrc[2] = D.ysrc[3] = D.csrc = 0;
                }
                D.ydst = P.cur_y4 + (uint32_t)((mby * 16 * pitch_y + mbx * 16) >> 2);
                D.cdst = P.cur_c4 + (uint32_t)((mby * 8 * pitch_c + mbx * 8) >> 2);
                D.rgba = P.rgba16 + (uint32_t)(mby * 16) * (P.rgba_pitch >> 4) + (uint32_t)(mbx * 4);
                if (mbx == 0) flags |= MBF_LEFT;
                if (mbx == mbw - 1) flags |= MBF_RIGHT;
                if (mby == 0) flags |= MBF_TOP;
                if (mby == mbh - 1) flags |= MBF_BOTTOM;
                if (emit_rgba && P.rgba) flags |= MBF_RGBA;
                D.flags = flags;
                D.pitches = (uint32_t)pitch_y | ((uint32_t)pitch_c << 16);
                D.rgba_pitch = P.rgba_pitch;
                D.pic = pic;
                S.mbd[m] = D;
            }
        }
        __syncthreads();

        // ================= phase 2: IDCT, 4 slots per warp, 8 lanes per slot =====================
        for (int g = warp; g * 4 < n_slots; g += TILE_THREADS / 32) {
            const int si = g * 4 + (lane >> 3), t = lane & 7;
            const int s = si < n_slots ? (int)S.order[si] : 0;
            const uint32_t sm = si < n_slots ? S.slotmeta[s] : 0u;
            const int cls = (int)(sm >> 8);
            const bool need = cls == CLS_FULL || cls == CLS_VERT;
            const bool vert = cls == CLS_VERT;
            const uint32_t R = need ? (sm & 0xFFu) : 0u;
            const int nmax = __reduce_max_sync(FULL, __popc(R));
            if (nmax == 0) continue;  // only Dc / Zero slots left (sorted last)
            float* c = S.pool + s * SLOT_FLOATS;
            // Per row y that holds a coefficient (ascending):
            //   row pass    t[y][i] = sum_x c[y][x] * B[x][i], ascending x (idct_1d, idct.rs:52-65), lane t = i;
            //               stored pre-divided by 4 (exact), which takes the /4 of idct.rs:189 out of the
            //               64-output rounding
            //   column pass out[i][j] += t[y][i] * B[y][j], lane t = j (pixel row)
            // All-zero rows contribute +-0 terms only and are skipped (exact).
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint32_t rem = R;
            for (int k = 0; k < nmax; k++) {
                const bool act = rem != 0;
                const int y = act ? __ffs(rem) - 1 : 0;
                rem &= rem - 1;
                const float4 ca = *reinterpret_cast<const float4*>(c + y * 8);
                const float4 cc = *reinterpret_cast<const float4*>(c + y * 8 + 4);
                const float bv = act ? S.basis[y * 8 + t] : 0.0f;  // exhausted lanes add +-0: no effect
                float a = fmul(ca.x, bt0);  // 0 + x == x
                a = fadd(a, fmul(ca.y, bt1));
                a = fadd(a, fmul(ca.z, bt2));
                a = fadd(a, fmul(ca.w, bt3));
                a = fadd(a, fmul(cc.x, bt4));
                a = fadd(a, fmul(cc.y, bt5));
                a = fadd(a, fmul(cc.z, bt6));
                a = fadd(a, fmul(cc.w, bt7));
                if (vert) a = ca.x;  // Vert: the first column feeds idct_1d directly (idct.rs:152-153)
                __syncwarp();        // every lane has read row y before it is overwritten
                if (act) c[y * 8 + t] = fmul(a, 0.25f);
                __syncwarp();
                const float4 ta = *reinterpret_cast<const float4*>(c + y * 8);
                const float4 tb = *reinterpret_cast<const float4*>(c + y * 8 + 4);
                acc[0] = fadd(acc[0], fmul(ta.x, bv));
                acc[1] = fadd(acc[1], fmul(ta.y, bv));
                acc[2] = fadd(acc[2], fmul(ta.z, bv));
                acc[3] = fadd(acc[3], fmul(ta.w, bv));
                acc[4] = fadd(acc[4], fmul(tb.x, bv));
                acc[5] = fadd(acc[5], fmul(tb.y, bv));
                acc[6] = fadd(acc[6], fmul(tb.z, bv));
                acc[7] = fadd(acc[7], fmul(tb.w, bv));
            }
            const float m = vert ? H263_B00 : 1.0f;
            int rr[8];
#pragma unroll
            for (int i = 0; i < 8; i++) rr[i] = round_q(acc[i], m);
            __syncwarp();
            if (need) {
                // residual row t in the lane order of phase 3: (r0,r2) (r1,r3) (r4,r6) (r5,r7)
                uint4 o;
                o.x = pack_clamp_s16x2(rr[2], rr[0]);
                o.y = pack_clamp_s16x2(rr[3], rr[1]);
                o.z = pack_clamp_s16x2(rr[6], rr[4]);
                o.w = pack_clamp_s16x2(rr[7], rr[5]);
                *reinterpret_cast<uint4*>(c + t * 4) = o;
            }
        }
        __syncthreads();

        // ================= phase 3: MC + add + clamp + stores + RGBA, all in registers ===========
        // unit = (row group rg of 4 luma rows, macroblock, h): 8 luma columns 8h..8h+7 of the 4 rows,
        // and the two chroma rows under them (8 columns) of ONE plane: Cb for h = 0, Cr for h = 1; the
        // chroma samples the RGBA conversion needs from the other plane come from the neighbour lane.
        // Consecutive threads: h, then macroblock -> neighbouring stores coalesce.
        {
            const int rg = tid >> unit_shift;
            const int mbi = m0 + ((tid >> 1) & (mb_per_pass - 1)), h = tid & 1;
            const bool unit_ok = rg < 4 && mbi < m1;
            const MbDesc& D = S.mbd[unit_ok ? mbi : m0];
            const uint32_t flags = unit_ok ? D.flags : 0u;
            const uint32_t pitch_y4 = (D.pitches & 0xFFFFu) >> 2, pitch_c4 = D.pitches >> 18;
            const int lb = ((rg >> 1) << 1) | h;  // luma block of this unit
            const int r0 = (rg & 1) * 4;          // first row of the unit inside its block

            RowSum ly[4];  // luma: 4 rows x 8 pixels in 16-bit lanes
            RowSum cy[2];  // chroma (own plane): 2 rows x 8 pixels
            if ((flags & (MBF_INTER | MBF_SLOW)) == MBF_INTER) {
                {
                    const uint32_t f = flags >> (4 * lb);
                    const int sh = (f & 3u) * 8, shb = sh + ((f & 4u) << 1);
                    const uint32_t wb = (f >> 3) & 1u, wt = 2u - wb;
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(pools.y) + D.ysrc[lb] + (uint32_t)r0 * pitch_y4;
                    RowSum hs[5];
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        // the fifth row is only needed for vertical interpolation; without it the load
                        // repeats row 3 (an L1 hit) and its weight is 0
                        const uint32_t* p = src + (r < 4 ? (uint32_t)r : 3u + wb) * pitch_y4;
                        hs[r] = row_sum8(__ldg(p), __ldg(p + 1), __ldg(p + 2), sh, shb);
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        ly[r].e0 = vmix(hs[r].e0, hs[r + 1].e0, wt, wb);
                        ly[r].o0 = vmix(hs[r].o0, hs[r + 1].o0, wt, wb);
                        ly[r].e1 = vmix(hs[r].e1, hs[r + 1].e1, wt, wb);
                        ly[r].o1 = vmix(hs[r].o1, hs[r + 1].o1, wt, wb);
                    }
                }
                {
                    const uint32_t f = flags >> 16;
                    const int sh = (f & 3u) * 8, shb = sh + ((f & 4u) << 1);
                    const uint32_t wb = (f >> 3) & 1u, wt = 2u - wb;
                    // chroma rows 2*rg, 2*rg+1 of the macroblock, all 8 columns, plane h
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(h ? pools.cr : pools.cb) + D.csrc + (uint32_t)(rg * 2) * pitch_c4;
                    RowSum hs[3];
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        const uint32_t* p = src + (r < 2 ? (uint32_t)r : 1u + wb) * pitch_c4;
                        hs[r] = row_sum8(__ldg(p), __ldg(p + 1), __ldg(p + 2), sh, shb);
                    }
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        cy[r].e0 = vmix(hs[r].e0, hs[r + 1].e0, wt, wb);
                        cy[r].o0 = vmix(hs[r].o0, hs[r + 1].o0, wt, wb);
                        cy[r].e1 = vmix(hs[r].e1, hs[r + 1].e1, wt, wb);
                        cy[r].o1 = vmix(hs[r].o1, hs[r + 1].o1, wt, wb);
                    }
                }
            } else if (flags & MBF_INTER) {
                // vectors beyond the replicated border: clamped per-sample fetch (generic path)
                const PicDev& P = pics[D.pic];
                const uint32_t* r = &S.mbrec[mbi * 6];
                const uint32_t w1 = r[1], w4 = r[4], w5 = r[5];
                const int mbx = (w1 >> 16) & 0xFF, mby = w1 >> 24;
                const uint32_t mvw = lb < 2 ? (w4 >> (16 * lb)) : (w5 >> (16 * (lb - 2)));
                const int mvx = (int8_t)(mvw & 0xFF), mvy = (int8_t)((mvw >> 8) & 0xFF);
                const int sumx = (int8_t)byte_of(w4, 0) + (int8_t)byte_of(w4, 2) + (int8_t)byte_of(w5, 0) + (int8_t)byte_of(w5, 2);
                const int sumy = (int8_t)byte_of(w4, 1) + (int8_t)byte_of(w4, 3) + (int8_t)byte_of(w5, 1) + (int8_t)byte_of(w5, 3);
                const int cvx = average_sum_of_mvs(sumx), cvy = average_sum_of_mvs(sumy);
#pragma unroll
                for (int rr = 0; rr < 4; rr++) {
                    uint32_t o0, o1;
                    mc_fetch8(P.ref[0], P.pitch_y, P.w, P.h, mbx * 16 + h * 8, mby * 16 + rg * 4 + rr, mvx, mvy, o0, o1);
                    ly[rr].e0 = __byte_perm(o0, 0, 0x4240), ly[rr].o0 = __byte_perm(o0, 0, 0x4341);
                    ly[rr].e1 = __byte_perm(o1, 0, 0x4240), ly[rr].o1 = __byte_perm(o1, 0, 0x4341);
                }
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    uint32_t o0, o1;
                    mc_fetch8(P.ref[1 + h], P.pitch_c, P.cw, P.ch, mbx * 8, mby * 8 + rg * 2 + rr, cvx, cvy, o0, o1);
                    cy[rr].e0 = __byte_perm(o0, 0, 0x4240), cy[rr].o0 = __byte_perm(o0, 0, 0x4341);
                    cy[rr].e1 = __byte_perm(o1, 0, 0x4240), cy[rr].o1 = __byte_perm(o1, 0, 0x4341);
                }
            } else {
                // intra: the prediction is the zero-initialised plane (picture.rs:42-48)
#pragma unroll
                for (int r = 0; r < 4; r++) ly[r] = RowSum{0, 0, 0, 0};
                cy[0] = cy[1] = RowSum{0, 0, 0, 0};
            }

            // ---- residuals: packed s16x2 add, saturate to [0, 255] (idct.rs:191-194) ----
            if (unit_ok) {
                const uint32_t m = S.meta[mbi * 6 + lb];
                const int cls = (int)(m & 7u);
                if (cls == CLS_DC) {
                    const uint32_t dd = __byte_perm(m, 0, 0x3232);  // (dcres, dcres)
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        ly[r].e0 = __viaddmin_s16x2_relu(ly[r].e0, dd, 0x00FF00FFu);
                        ly[r].o0 = __viaddmin_s16x2_relu(ly[r].o0, dd, 0x00FF00FFu);
                        ly[r].e1 = __viaddmin_s16x2_relu(ly[r].e1, dd, 0x00FF00FFu);
                        ly[r].o1 = __viaddmin_s16x2_relu(ly[r].o1, dd, 0x00FF00FFu);
                    }
                } else if (cls != CLS_ZERO) {
                    const float* c = S.pool + ((m >> 3) & 0xFFu) * SLOT_FLOATS + r0 * 4;
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(c + r * 4);
                        ly[r].e0 = __viaddmin_s16x2_relu(ly[r].e0, rv.x, 0x00FF00FFu);
                        ly[r].o0 = __viaddmin_s16x2_relu(ly[r].o0, rv.y, 0x00FF00FFu);
                        ly[r].e1 = __viaddmin_s16x2_relu(ly[r].e1, rv.z, 0x00FF00FFu);
                        ly[r].o1 = __viaddmin_s16x2_relu(ly[r].o1, rv.w, 0x00FF00FFu);
                    }
                }
                const uint32_t mc = S.meta[mbi * 6 + 4 + h];
                const int ccls = (int)(mc & 7u);
                if (ccls == CLS_DC) {
                    const uint32_t dd = __byte_perm(mc, 0, 0x3232);
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        cy[r].e0 = __viaddmin_s16x2_relu(cy[r].e0, dd, 0x00FF00FFu);
                        cy[r].o0 = __viaddmin_s16x2_relu(cy[r].o0, dd, 0x00FF00FFu);
                        cy[r].e1 = __viaddmin_s16x2_relu(cy[r].e1, dd, 0x00FF00FFu);
                        cy[r].o1 = __viaddmin_s16x2_relu(cy[r].o1, dd, 0x00FF00FFu);
                    }
                } else if (ccls != CLS_ZERO) {
                    const float* c = S.pool + ((mc >> 3) & 0xFFu) * SLOT_FLOATS + (rg * 2) * 4;
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(c + r * 4);
                        cy[r].e0 = __viaddmin_s16x2_relu(cy[r].e0, rv.x, 0x00FF00FFu);
                        cy[r].o0 = __viaddmin_s16x2_relu(cy[r].o0, rv.y, 0x00FF00FFu);
                        cy[r].e1 = __viaddmin_s16x2_relu(cy[r].e1, rv.z, 0x00FF00FFu);
                        cy[r].o1 = __viaddmin_s16x2_relu(cy[r].o1, rv.w, 0x00FF00FFu);
                    }
                }
            }

            // ---- plane stores (+ border replication for the next picture's prediction) ----
            const uint32_t pitch_y = pitch_y4 * 4, pitch_c = pitch_c4 * 4;
            uint32_t yw[4][2], cw[2][2];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                yw[r][0] = __byte_perm(ly[r].e0, ly[r].o0, 0x6240);
                yw[r][1] = __byte_perm(ly[r].e1, ly[r].o1, 0x6240);
            }
#pragma unroll
            for (int r = 0; r < 2; r++) {
                cw[r][0] = __byte_perm(cy[r].e0, cy[r].o0, 0x6240);
                cw[r][1] = __byte_perm(cy[r].e1, cy[r].o1, 0x6240);
            }
            if (unit_ok) {
                uint8_t* py = pools.y + (size_t)(D.ydst + (uint32_t)(rg * 4) * pitch_y4 + (uint32_t)(h * 2)) * 4;
                uint8_t* pc = (h ? pools.cr : pools.cb) + (size_t)(D.cdst + (uint32_t)(rg * 2) * pitch_c4) * 4;
#pragma unroll
                for (int r = 0; r < 4; r++) *reinterpret_cast<uint2*>(py + r * pitch_y) = make_uint2(yw[r][0], yw[r][1]);
#pragma unroll
                for (int r = 0; r < 2; r++) *reinterpret_cast<uint2*>(pc + r * pitch_c) = make_uint2(cw[r][0], cw[r][1]);
                // edges this unit owns: luma left for h = 0, luma right for h = 1, chroma both sides
                const uint32_t edge = flags & ((h ? MBF_RIGHT : MBF_LEFT) | (rg == 0 ? MBF_TOP : 0u) | (rg == 3 ? MBF_BOTTOM : 0u));
                const uint32_t cedge = flags & (MBF_LEFT | MBF_RIGHT);
                if (edge | cedge) {
                    const bool e_left = (edge & MBF_LEFT) != 0, e_right = (edge & MBF_RIGHT) != 0;
                    const bool e_top = (edge & MBF_TOP) != 0, e_bot = (edge & MBF_BOTTOM) != 0;
                    const bool c_left = (cedge & MBF_LEFT) != 0, c_right = (cedge & MBF_RIGHT) != 0;
                    if (e_left | e_right) {
                        // 16 luma pixels of horizontal extension for this unit's rows
                        const int lo = e_left ? -16 : 8;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint32_t v = e_left ? splat_lo(yw[r][0]) : splat_hi(yw[r][1]);
                            *reinterpret_cast<uint4*>(py + r * pitch_y + lo) = make_uint4(v, v, v, v);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 2; r++) {  // 8 chroma pixels on each side that is a picture edge
                        if (c_left) {
                            const uint32_t v = splat_lo(cw[r][0]);
                            *reinterpret_cast<uint2*>(pc + r * pitch_c - 8) = make_uint2(v, v);
                        }
                        if (c_right) {
                            const uint32_t v = splat_hi(cw[r][1]);
                            *reinterpret_cast<uint2*>(pc + r * pitch_c + 8) = make_uint2(v, v);
                        }
                    }
                    if (e_top | e_bot) {
                        // vertical extension: 16 luma / 8 chroma rows above row 0 or below the last row,
                        // including the corners when the unit also sits on a vertical edge
                        const uint32_t v0 = e_top ? yw[0][0] : yw[3][0], v1 = e_top ? yw[0][1] : yw[3][1];
                        uint8_t* rowp = e_top ? py : py + 3 * pitch_y;
                        const ptrdiff_t dir = e_top ? -(ptrdiff_t)pitch_y : (ptrdiff_t)pitch_y;
                        const uint32_t corner = e_left ? splat_lo(v0) : splat_hi(v1);
                        for (int k = 1; k <= 16; k++) {
                            uint8_t* d = rowp + k * dir;
                            *reinterpret_cast<uint2*>(d) = make_uint2(v0, v1);
                            if (e_left) *reinterpret_cast<uint4*>(d - 16) = make_uint4(corner, corner, corner, corner);
                            if (e_right) *reinterpret_cast<uint4*>(d + 8) = make_uint4(corner, corner, corner, corner);
                        }
                        const uint32_t c0 = e_top ? cw[0][0] : cw[1][0], c1 = e_top ? cw[0][1] : cw[1][1];
                        uint8_t* rc = e_top ? pc : pc + pitch_c;
                        const ptrdiff_t cdir = e_top ? -(ptrdiff_t)pitch_c : (ptrdiff_t)pitch_c;
                        const uint32_t cl = splat_lo(c0), cr = splat_hi(c1);
                        for (int k = 1; k <= 8; k++) {
                            uint8_t* d = rc + k * cdir;
                            *reinterpret_cast<uint2*>(d) = make_uint2(c0, c1);
                            if (c_left) *reinterpret_cast<uint2*>(d - 8) = make_uint2(cl, cl);
                            if (c_right) *reinterpret_cast<uint2*>(d + 8) = make_uint2(cr, cr);
                        }
                    }
                }
            }

            // ---- BT.601 RGBA (bt601.rs:12-59): 8 pixels per row, one 256-bit store per row ----
            // The unit needs chroma columns 4h..4h+3 of both planes: its own plane has them in
            // (e_h, o_h); the other plane's come from the neighbour lane (h ^ 1), which sends the half
            // of its row that it does not use itself.
            {
                uint32_t ce[2][2], co[2][2];  // [plane 0 = Cb, 1 = Cr][chroma row]: lanes (c0, c2) / (c1, c3)
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint32_t mine_e = h ? cy[r].e1 : cy[r].e0, mine_o = h ? cy[r].o1 : cy[r].o0;
                    const uint32_t send_e = h ? cy[r].e0 : cy[r].e1, send_o = h ? cy[r].o0 : cy[r].o1;
                    const uint32_t got_e = __shfl_xor_sync(FULL, send_e, 1), got_o = __shfl_xor_sync(FULL, send_o, 1);
                    ce[0][r] = h ? got_e : mine_e, co[0][r] = h ? got_o : mine_o;
                    ce[1][r] = h ? mine_e : got_e, co[1][r] = h ? mine_o : got_o;
                }
                if (flags & MBF_RGBA) {
                    uint8_t* o = pools.rgba + (size_t)D.rgba * 16 + (size_t)(rg * 4) * D.rgba_pitch + (size_t)(h * 32);
#pragma unroll
                    for (int cr2 = 0; cr2 < 2; cr2++) {
                        // chroma row cr2 serves luma rows 2*cr2, 2*cr2+1; sample j serves pixels 2j, 2j+1
                        const CT t0 = chroma_terms_folded((int)(ce[0][cr2] & 0xFFFFu), (int)(ce[1][cr2] & 0xFFFFu));
                        const CT t1 = chroma_terms_folded((int)(co[0][cr2] & 0xFFFFu), (int)(co[1][cr2] & 0xFFFFu));
                        const CT t2 = chroma_terms_folded((int)(ce[0][cr2] >> 16), (int)(ce[1][cr2] >> 16));
                        const CT t3 = chroma_terms_folded((int)(co[0][cr2] >> 16), (int)(co[1][cr2] >> 16));
#pragma unroll
                        for (int rr = 0; rr < 2; rr++) {
                            const RowSum& L = ly[cr2 * 2 + rr];
                            uint32_t px[8];
                            px[0] = rgba_px((int)(L.e0 & 0xFFFFu), t0);
                            px[1] = rgba_px((int)(L.o0 & 0xFFFFu), t0);
                            px[2] = rgba_px((int)(L.e0 >> 16), t1);
                            px[3] = rgba_px((int)(L.o0 >> 16), t1);
                            px[4] = rgba_px((int)(L.e1 & 0xFFFFu), t2);
                            px[5] = rgba_px((int)(L.o1 & 0xFFFFu), t2);
                            px[6] = rgba_px((int)(L.e1 >> 16), t3);
                            px[7] = rgba_px((int)(L.o1 >> 16), t3);
                            st_global_v8(o + (size_t)(cr2 * 2 + rr) * D.rgba_pitch, px);
                        }
                    }
                }
            }
        }
        if (pass + 1 < n_pass) __syncthreads();  // the pool is reused by the next pass
    }
}

void launch_recon_tile(const PicDev* pics, const h263cu_mb* mbs, const h263cu_event* events, uint32_t n_mbs, int emit_rgba,
                       const Pools& pools, cudaStream_t stream) {
    if (n_mbs == 0) return;
    const uint32_t grid = (n_mbs + TILE_MBS - 1) / TILE_MBS;
    recon_tile_kernel<<<grid, TILE_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
}

}  // namespace h263dev
