// recon_tile.cu -- the hot kernel: fused reconstruction, one WARP per 4 consecutive macroblocks,
// for pictures whose size is a multiple of 16 and whose reference planes carry the replicated
// border.  Warps are autonomous: after the constant tables are staged there
// is no CTA-wide barrier, only __syncwarp(), so a warp waiting on memory never holds up others.
//
//   phase 0  lanes 0..23 = the 24 blocks of the 4 macroblocks: event counts, motion vectors ->
//            source offsets / alignment / half-pel flags (gather.rs:140-204, types.rs:721-729,
//            759-768); coded blocks are compacted into slots, sorted by event count
//   phase 1  8 lanes per slot, 4 slots per pass: the lanes take one run/level event each, a
//            segmented prefix sum gives the zig-zag index, dequantise, scatter, classify
//            (rle.rs:82-172)
//   phase 2  same lanes: per row that holds a coefficient a row pass and a column-pass update,
//            then rounding (idct.rs:52-65,170-198); packed s16 residuals go to the slot's rows
//   phase 3  lane = 8x4 luma pixels + the 8x2 chroma pixels of one plane under them: half-pel
//            interpolation in 16-bit lanes (gather.rs:34-40,103-113), saturating residual add,
//            plane stores, border replication, BT.601 RGBA (bt601.rs:12-59) with 256-bit stores.
//
// Everything a lane needs per block sits in shared memory as 32-bit offsets from the context's
// pool bases (kernel parameters), so the epilogue does no 64-bit pointer chasing.
// Bound: HBM bandwidth with the integer/FP32 issue rate as the secondary ceiling -- this file is
// written for instruction count: see DESIGN.md section 4 for the per-phase budget.
#include "recon_common.cuh"

namespace h263dev {

namespace {

constexpr int WARP_MBS = 4;      // macroblocks per warp
#ifndef H263_CTA_WARPS
#define H263_CTA_WARPS 4  // without the staging barrier a CTA is just a group of warps: 4 beat 8 by 1-2 % (shorter tails)
#endif
#ifndef H263_PERSISTENT
#define H263_PERSISTENT 0
#endif
#ifndef H263_PREFETCH_L2
#define H263_PREFETCH_L2 0
#endif
#ifndef H263_LDG64
#define H263_LDG64 0
#endif
// Ablation builds for time attribution (results are wrong by design): 1 = no event walk / transform,
// 2 = no prediction loads (every macroblock treated as intra), 4 = no RGBA, 8 = no plane stores,
// 16 = RGBA computed but not stored, 32 = Cr predicted from the Cb plane (chroma load sectors halved),
// 64 = Cr stored onto the Cb plane (chroma store sectors halved), 128 = chroma loaded in the pattern of a lane that
// owns 4 columns of both planes.
#ifndef H263_ABLATE
#define H263_ABLATE 0
#endif
#ifndef H263_MIN_CTAS
#define H263_MIN_CTAS (32 / H263_CTA_WARPS)
#endif
// 1 = stage BASIS_TABLE / the de-zigzag map in shared memory behind a CTA barrier (versions up to v10);
// 0 = read them from global memory (L1-resident, 320 bytes): no per-CTA staging, no barrier at all, so
// a CTA's warps start on their record loads at once and small CTAs cost nothing extra.
#ifndef H263_SMEM_TABLES
#define H263_SMEM_TABLES 0
#endif
// L2 eviction priorities (createpolicy): bit 0 = RGBA stores evict_first (write-once output), bit 1 = plane
// stores evict_last (the next step's prediction source), bit 2 = prediction loads evict_first (dead after use)
#ifndef H263_L2_POLICY
#define H263_L2_POLICY 0
#endif
constexpr int CTA_WARPS = H263_CTA_WARPS;
// Persistent warps drawing tiles from a global counter were measured SLOWER than one CTA per 32
// consecutive macroblocks (325 vs 303 us per 1024-CIF step): neighbouring tiles then run on different
// SMs and lose the L1 sharing of overlapping prediction windows (L1 hit rate 48 % vs 58 %), and partial
// DRAM write atoms of chroma rows are no longer merged.  Kept as a build option for the record.
constexpr bool kPersistent = H263_PERSISTENT != 0;
// Tiles per warp (non-persistent build): warp w of CTA b takes tiles (b*T + it)*CTA_WARPS + w, it < T,
// so the CTA's warps stay on neighbouring macroblocks while the per-warp work averages out.
#ifndef H263_TILES_PER_WARP
#define H263_TILES_PER_WARP 1
#endif
constexpr int kTilesPerWarp = H263_TILES_PER_WARP;
constexpr int CTA_THREADS = CTA_WARPS * 32;
constexpr int WARP_BLOCKS = WARP_MBS * 6;
constexpr int RES_WORDS = 36;
constexpr int EV_CAP = 96;       // events walked at once (one slot has at most 64); more are walked in chunks
constexpr int SLOT_FLOATS = 68;  // 64 + 4 pad: 16 B aligned, the four slots of a pass start 4 banks apart

// per-macroblock flags (WarpSmem.mb[][3])
constexpr uint32_t MBF_INTER = 1u << 0;
constexpr uint32_t MBF_LEFT = 1u << 2, MBF_RIGHT = 1u << 3, MBF_TOP = 1u << 4, MBF_BOTTOM = 1u << 5;
constexpr uint32_t MBF_RGBA = 1u << 6;
// per-block flags (WarpSmem.bf[]): align(2) | ix | iy | slow
constexpr uint32_t BF_SLOW = 1u << 4;  // the vector leaves the replicated border: clamped per-sample path

struct __align__(16) WarpSmem {
    float coef[4][SLOT_FLOATS];        // coefficients of the four slots in flight -> row-pass output
    uint32_t res[WARP_BLOCKS][RES_WORDS];  // per slot: 8 residual rows of 16 bytes, lanes (r0,r2)(r1,r3)(r4,r6)(r5,r7);
                                       // 4 words of padding put the four slots of a pass on different banks
    uint32_t mb[WARP_MBS][8];          // ydst, cdst, rgba, flags, pitches, rgba_pitch, pic, -
    uint2 slotdesc[WARP_BLOCKS];       // x = first event unit (absolute), y = nev | quant<<8 | wide<<13 | inter<<14 | block<<16 | dc<<24
    uint32_t mbrec[WARP_BLOCKS];       // the four macroblock records
    uint32_t bd[WARP_BLOCKS];          // per block: 4-byte offset (from y_pool or cb/cr_pool) of the aligned word
                                       // that holds the first source pixel of the block's row 0
    uint32_t bf[WARP_BLOCKS];          // per block: BF_* flags
    uint32_t meta[WARP_BLOCKS];        // per block: cls[2:0] | slot[7:3] | dcres[31:16]
    uint32_t evbuf[EV_CAP];            // walked events of the slots in flight: lin[5:0] | dropped[15] | value[31:16]
    uint32_t sstart[WARP_BLOCKS];      // per slot: index of its first event among the warp's events
    uint32_t slotinfo[WARP_BLOCKS];    // per slot, gathered by the walk: rows[7:0] | column > 0 [8] | overflow [9]
    uint32_t slotcls[WARP_BLOCKS];     // per slot after classification: rows to transform[7:0] | cls[10:8] | has_dc[11]
    uint32_t order[WARP_BLOCKS];       // slots of the chunk sorted by rows to transform (most first)
};

struct TileSmem {
    WarpSmem w[CTA_WARPS];
#if H263_SMEM_TABLES
    float basis[64];
    uint8_t dezigzag[64];
#endif
};

#if !H263_SMEM_TABLES
__device__ const float g_basis[8][8] = H263_BASIS_TABLE;
__device__ const uint8_t g_dezigzag[64] = H263_DEZIGZAG_LINEAR;
#endif

// compile-time copy of BASIS_TABLE: with y and j unrolled these fold into immediates of the column pass
__device__ __forceinline__ constexpr float k_basis(int y, int j) {
    constexpr float T[8][8] = H263_BASIS_TABLE;
    return T[y][j];
}

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
// ---- half-pel interpolation in 16-bit lanes -------------------------------------------------
// The three aligned words that hold the 9 bytes a prediction row needs, p = word of the first pixel.
// With H263_LDG64 they come from two 8-byte loads of the enclosing 16-byte window (two requests
// instead of three on the L1 data pipe, the busiest unit of this kernel); odd = p is an odd word.
// `third` = the row needs its third word (it does unless the block starts word-aligned with a full-pel x vector);
// `row` = the row is needed at all (the extra row below a unit only when the vector is half-pel in y).  A load pass
// costs the L1 one look-up per distinct sector it touches, so lanes that do not need a word stay out of the pass.
__device__ __forceinline__ void load_row3(const uint32_t* p, bool odd, bool third, bool row, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
#if H263_LDG64
    const uint2* q = reinterpret_cast<const uint2*>(p - (odd ? 1 : 0));
    uint2 v0 = make_uint2(0u, 0u), v1 = make_uint2(0u, 0u);
    if (row) {
        v0 = __ldg(q);
        if (third || odd) v1 = __ldg(q + 1);
    }
    w0 = odd ? v0.y : v0.x, w1 = odd ? v1.x : v0.y, w2 = odd ? v1.y : v1.x;
#else
    w0 = w1 = w2 = 0u;
    if (row) {
        w0 = __ldg(p), w1 = __ldg(p + 1);
        if (third) w2 = __ldg(p + 2);
    }
#endif
}

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
// The multipliers are passed in registers (see opaque()): a product with an immediate multiplier
// AND an immediate addend does not encode, and the compiler would otherwise materialise the
// multiplier once per use.
__device__ __forceinline__ int opaque(int v) {
    asm("" : "+r"(v));
    return v;
}
__device__ __forceinline__ CT chroma_terms_folded(int cb, int cr, int kr, int kg, int kb) {
    CT t;
    t.r = cr * kr + (32768 - 128 * 104597 - 16 * 76309);
    t.g = cr * kg + (cb * -25675 + (32768 + 128 * 53279 + 128 * 25675 - 16 * 76309));
    t.b = cb * kb + (32768 - 128 * 132201 - 16 * 76309);
    return t;
}
__device__ __forceinline__ uint32_t rgba_px(int y, const CT& t) {
    const int r = (y * 76309 + t.r) >> 16, g = (y * 76309 + t.g) >> 16, b = (y * 76309 + t.b) >> 16;
    return pack_sat(g, r, pack_sat(255, b, 0));
}

__device__ __forceinline__ void st_global_v8(uint8_t* p, const uint32_t (&v)[8]) {
#if H263_L2_POLICY & 1
    asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
#else
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
#endif
}

__device__ __forceinline__ uint32_t splat_lo(uint32_t w) { return __byte_perm(w, 0, 0x0000); }
__device__ __forceinline__ uint32_t splat_hi(uint32_t w) { return __byte_perm(w, 0, 0x3333); }

}  // namespace


// PY / PC / PR: luma, chroma and RGBA row pitches in bytes when they are known at compile time (the
// standard picture formats; all planes of a context share them), 0 = read them from the descriptors.
// Constant pitches turn every row address of the epilogue into an immediate offset.
// EDGE: pictures whose size is not a multiple of 16.  The planes are macroblock-rounded; the part of the last
// macroblock column / row that lies outside the picture is overwritten with the picture's edge pixels before the
// stores (and the border replication continues from there), so that prediction reads beyond the true edge see
// read_sample's clamp (gather.rs:16-31).  Compiled out of the instantiations for aligned pictures.
// WIDE_MV: some vector of the step may leave the replicated border (no H263CU_PICFLAG_MV_IN_RANGE; unreachable from a
// parsed stream, mvd_pred.rs:70-117): the instantiation with the clamped per-sample path.  Without it vectors are
// clamped to the range in phase 0, and the 900 instructions of that path do not weigh on the register allocation.
template <int PY, int PC, int PR, bool EDGE, bool WIDE_MV>
__global__ void __launch_bounds__(CTA_THREADS, H263_MIN_CTAS)
    recon_tile_kernel(const PicDev* __restrict__ pics, const h263cu_mb* __restrict__ mbs,
                      const h263cu_event* __restrict__ events, uint32_t n_mbs, int emit_rgba, const Pools pools) {
    __shared__ __align__(16) TileSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if H263_SMEM_TABLES
    if (tid < 64) {
        S.basis[tid] = c_basis[tid >> 3][tid & 7];
        S.dezigzag[tid] = c_dezigzag[tid];
    }
    __syncthreads();  // the only CTA-wide barrier: from here on every warp runs on its own
#endif
    WarpSmem& W = S.w[warp];
    const int g = lane >> 3, t = lane & 7;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // Persistent warps with dynamic work distribution: every warp starts on the 4-macroblock tile of
    // its grid position and then draws tile numbers from a global counter, so that no warp slot idles
    // while a neighbour works on a heavier tile.  The next tile number is drawn at the top of a tile and
    // its record words are fetched before the epilogue, one tile ahead of their use.
    const uint32_t n_tiles = (n_mbs + WARP_MBS - 1) / WARP_MBS;
    uint32_t tile = (kPersistent ? blockIdx.x : blockIdx.x * kTilesPerWarp) * CTA_WARPS + warp;
    int tiles_left = kTilesPerWarp;
    const uint32_t* mbs32 = reinterpret_cast<const uint32_t*>(mbs);
    uint32_t rec = 0;
    if (tile < n_tiles && (uint32_t)lane < min((uint32_t)WARP_MBS, n_mbs - tile * WARP_MBS) * 6) rec = __ldg(mbs32 + (size_t)tile * WARP_BLOCKS + lane);
  while (tile < n_tiles) {
    const uint32_t mb0 = tile * WARP_MBS;
    const int n_w = (int)min((uint32_t)WARP_MBS, n_mbs - mb0);
    uint32_t next_tile = 0;
    if (!kPersistent) next_tile = --tiles_left > 0 ? tile + CTA_WARPS : n_tiles;
    else if (lane == 0) next_tile = gridDim.x * CTA_WARPS + atomicAdd(pools.work_counter, 1u);

    // ================= phase 0: lane = block (macroblock lane / 6, block lane % 6) ==============
    __syncwarp();  // the previous tile's epilogue has finished with the warp's shared memory
    if (lane < n_w * 6) W.mbrec[lane] = rec;
    __syncwarp();
    int n_slots;
    {
        const bool bvalid = lane < n_w * 6;
        const int bm = bvalid ? (lane * 43) >> 8 : 0, bb = bvalid ? lane - bm * 6 : 0;
        const uint32_t* r = &W.mbrec[bm * 6];
        const uint32_t w0 = r[0], w1 = r[1], w2 = r[2], w3 = r[3], w4 = r[4], w5 = r[5];
        const uint32_t pic = w1 & 0xFFFFu;
        const PicDev& P = pics[pic];
        const int mbx = (w1 >> 16) & 0xFF, mby = w1 >> 24;
        const bool inter = (w2 & H263CU_MB_INTER) != 0, wide = (w2 & H263CU_MB_WIDE) != 0;
        // nev[6] = bytes 2..7 of (w2, w3); intradc[6] = bytes 0..5 of (w4, w5)
        const uint32_t nev = bvalid ? __byte_perm(w2, w3, 0x4440u + (uint32_t)bb + 2u) & 0xFFu : 0u;
        const uint32_t code = inter ? 0u : __byte_perm(w4, w5, 0x4440u + (uint32_t)bb) & 0xFFu;

        // motion: source offset, alignment and half-pel flags of this block
        uint32_t boff = 0, bflags = 0;
        if (inter) {
            int mvx, mvy, sx, sy, pitch;
            uint32_t base;
            bool in_range;
            if (bb < 4) {
                // mv[bb] = bytes 2bb, 2bb+1 of (w4, w5), sign-extended
                mvx = (int)(int8_t)__byte_perm(w4, w5, 0x4440u + 2u * (uint32_t)bb), mvy = (int)(int8_t)__byte_perm(w4, w5, 0x4441u + 2u * (uint32_t)bb);
                if (!WIDE_MV) mvx = max(min(mvx, 31), -32), mvy = max(min(mvy, 31), -32);
                in_range = mvx >= -32 && mvx <= 31 && mvy >= -32 && mvy <= 31;
                sx = mbx * 16 + (bb & 1) * 8 + (mvx >> 1), sy = mby * 16 + (bb >> 1) * 8 + (mvy >> 1);
                pitch = PY ? PY : P.pitch_y, base = P.ref_y4;
            } else {
                // both chroma blocks use the average of the four luma vectors (gather.rs:182, types.rs:759-768)
                const int sumx = (int8_t)byte_of(w4, 0) + (int8_t)byte_of(w4, 2) + (int8_t)byte_of(w5, 0) + (int8_t)byte_of(w5, 2);
                const int sumy = (int8_t)byte_of(w4, 1) + (int8_t)byte_of(w4, 3) + (int8_t)byte_of(w5, 1) + (int8_t)byte_of(w5, 3);
                mvx = average_sum_of_mvs(sumx), mvy = average_sum_of_mvs(sumy);
                if (!WIDE_MV) mvx = max(min(mvx, 15), -16), mvy = max(min(mvy, 15), -16);
                in_range = mvx >= -16 && mvx <= 15 && mvy >= -16 && mvy <= 15;
                sx = mbx * 8 + (mvx >> 1), sy = mby * 8 + (mvy >> 1);
                pitch = PC ? PC : P.pitch_c, base = P.ref_c4;
            }
            const int a = sx & 3;
            boff = base + (uint32_t)((sy * pitch + (sx - a)) >> 2);
            bflags = (uint32_t)(a | ((mvx & 1) << 2) | ((mvy & 1) << 3)) | (!WIDE_MV || in_range ? 0u : BF_SLOW);
#if H263_PREFETCH_L2
            // the 9 source rows of this block (reference planes are DRAM-resident: written a step ago):
            // start the DRAM -> L2 transfer now, the epilogue loads them ~1000 instructions later
            if (bvalid && in_range) {
                const uint8_t* pbase = (bb < 4 ? pools.y : (bb == 4 ? pools.cb : pools.cr)) + (size_t)boff * 4;
#pragma unroll
                for (int rr = 0; rr < 9; rr++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pbase + (size_t)rr * pitch));
            }
#endif
        }
        if (bvalid) {
            W.bd[lane] = boff;
            W.bf[lane] = bflags;
            if (bb == 0) {
                const int pitch_y = PY ? PY : P.pitch_y, pitch_c = PC ? PC : P.pitch_c;
                const int mbw = (P.w + 15) >> 4, mbh = (P.h + 15) >> 4;
                // valid luma columns / rows of the last macroblock column / row (0 = all 16) and the same for
                // chroma (0 = all 8)
                const uint32_t ed = EDGE ? ((uint32_t)(P.w & 15) | ((uint32_t)(P.h & 15) << 4) | ((uint32_t)(P.cw & 7) << 8) |
                                            ((uint32_t)(P.ch & 7) << 12))
                                         : 0u;
                const uint32_t rgba_pitch = PR ? PR : P.rgba_pitch;
                uint32_t flags = inter ? MBF_INTER : 0u;
                if (mbx == 0) flags |= MBF_LEFT;
                if (mbx == mbw - 1) flags |= MBF_RIGHT;
                if (mby == 0) flags |= MBF_TOP;
                if (mby == mbh - 1) flags |= MBF_BOTTOM;
                if (emit_rgba && P.rgba) flags |= MBF_RGBA;
                *reinterpret_cast<uint4*>(&W.mb[bm][0]) =
                    make_uint4(P.cur_y4 + (uint32_t)((mby * 16 * pitch_y + mbx * 16) >> 2),
                               P.cur_c4 + (uint32_t)((mby * 8 * pitch_c + mbx * 8) >> 2),
                               P.rgba16 + (uint32_t)(mby * 16) * (rgba_pitch >> 4) + (uint32_t)(mbx * 4), flags);
                *reinterpret_cast<uint4*>(&W.mb[bm][4]) = make_uint4((uint32_t)pitch_y | ((uint32_t)pitch_c << 16), rgba_pitch, pic, ed);
            }
        }

        // coded blocks -> slots (block order); their events form one sequence per warp
        const bool coded = nev > 0;
        const uint32_t coded_mask = __ballot_sync(FULL, coded);
        const int pos = __popc(coded_mask & lt_mask);
        n_slots = __popc(coded_mask);
        uint32_t ev_incl = nev;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(FULL, ev_incl, d);
            if (lane >= d) ev_incl += x;
        }
        // events of the blocks before this one inside the macroblock
        const uint32_t before = (ev_incl - nev) - __shfl_sync(FULL, ev_incl - nev, bm * 6);
        if (coded) {
            W.sstart[pos] = ev_incl - nev;
            const uint32_t first = P.first_event + w0 + (wide ? 2 * before : before);
            const uint32_t quant = (w2 >> 8) & 31u;
            W.slotdesc[pos] = make_uint2(first, nev | (quant << 8) | (wide ? 1u << 13 : 0u) | (inter ? 1u << 14 : 0u) |
                                                    ((uint32_t)lane << 16) | (code << 24));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(events + first));
        } else if (bvalid) {
            // no coefficients: Dc(level) when an intra DC is present, else Zero (rle.rs:94-104)
            const bool has_dc = code != 0;
            const int dcres = has_dc ? round_residual_dc((float)intradc_level((int)code)) : 0;
            W.meta[lane] = (uint32_t)(has_dc ? CLS_DC : CLS_ZERO) | ((uint32_t)dcres << 16);
        }
    }
    __syncwarp();

    // ================= phases 1 + 2, per chunk of slots whose events fit the event buffer ==========
    // (one chunk unless the four macroblocks hold more than EV_CAP events)
    {
#if H263_SMEM_TABLES
        const float* const basis = S.basis;
#else
        const float* const basis = &g_basis[0][0];
#endif
        const float bt0 = basis[0 * 8 + t], bt1 = basis[1 * 8 + t], bt2 = basis[2 * 8 + t], bt3 = basis[3 * 8 + t],
                    bt4 = basis[4 * 8 + t], bt5 = basis[5 * 8 + t], bt6 = basis[6 * 8 + t], bt7 = basis[7 * 8 + t];
        float* c = W.coef[g];
        // lane = slot for the per-slot steps
        if (H263_ABLATE & 1) n_slots = 0;
        const bool is_slot = lane < n_slots;
        const uint2 my_sd = is_slot ? W.slotdesc[lane] : make_uint2(0u, 0u);
        const uint32_t my_start = is_slot ? W.sstart[lane] : 0xFFFFFFFFu;
        const uint32_t my_end = my_start + (my_sd.y & 0xFFu);
        int s_lo = 0;
        uint32_t e_lo = 0;
        while (s_lo < n_slots) {
            // slots of this chunk: the longest run starting at s_lo whose events fit the buffer
            // (a slot with more events than the buffer holds -- a malformed record, real blocks have at most
            // 64 -- is walked as far as the buffer goes: its index is past 63 by then and it is dropped)
            const int s_hi = max(s_lo + 1, s_lo + __popc(__ballot_sync(FULL, is_slot && lane >= s_lo && my_end - e_lo <= (uint32_t)EV_CAP)));
            const bool in_chunk = lane >= s_lo && lane < s_hi;
            const uint32_t e_end = __shfl_sync(FULL, my_end, s_hi - 1);
            const uint32_t e_hi = min(e_end, e_lo + (uint32_t)EV_CAP);
            if (in_chunk) W.slotinfo[lane] = 0u;
            __syncwarp();

            // ---- phase 1: lane = event.  The zig-zag index is a segmented prefix sum of run + 1 over the
            // events of a slot (rle.rs:117-135); dequantise and park (position, value) in the event buffer.
            uint32_t carry = 0;  // running sum of the slot that continues from the previous round
            for (uint32_t eb = e_lo; eb < e_hi; eb += 32) {
                const uint32_t e = eb + lane;
                const bool act = e < e_hi;
                // which slot: slots whose first event lies in this round mark their head
                const bool head = in_chunk && my_start >= eb && my_start < eb + 32;
                const uint32_t heads = __reduce_or_sync(FULL, head ? 1u << (my_start - eb) : 0u);
                const int n_before = __popc(__ballot_sync(FULL, in_chunk && my_start < eb));
                const uint32_t below = heads & (0xFFFFFFFFu >> (31 - lane));
                const int slot = s_lo + n_before - 1 + __popc(below);
                const int seg = below ? 31 - __clz(below) : -1;  // lane where my slot's events start in this round
                int v = 0, val = 0;
                bool inter = false;
                if (act) {
                    const uint2 sd = W.slotdesc[slot];
                    const uint32_t k = e - W.sstart[slot];
                    const int quant = (int)((sd.y >> 8) & 31u);
                    inter = (sd.y >> 14) & 1u;
                    int run;
                    if (!((sd.y >> 13) & 1u)) {
                        const uint32_t u = __ldg(events + sd.x + k);
                        run = (int)(u >> 10);
                        val = dequant_narrow(((int)(u << 22)) >> 22, 2 * quant, quant - 1 + (quant & 1));
                    } else {
                        run = __ldg(events + sd.x + 2 * k) & 63;
                        val = dequant((int16_t)__ldg(events + sd.x + 2 * k + 1), quant);
                    }
                    v = run + 1;
                }
                const int seg0 = max(seg, 0);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int x = __shfl_up_sync(FULL, v, d);
                    if (lane - d >= seg0) v += x;
                }
                if (seg < 0) v += (int)carry;
                carry = (uint32_t)__shfl_sync(FULL, v, 31);
                if (act) {
                    const int idx = (inter ? 0 : 1) + v - 1;  // intra: the DC occupies zig-zag index 0 (rle.rs:117-121)
                    uint32_t bits, ent;
                    if (idx < 64) {
#if H263_SMEM_TABLES
                        const int lin = S.dezigzag[idx];
#else
                        const int lin = g_dezigzag[idx];
#endif
                        ent = (uint32_t)lin | ((uint32_t)val << 16);
                        bits = (1u << (lin >> 3)) | ((lin & 7) ? 0x100u : 0u);
                    } else {
                        ent = 0x8000u;  // the whole block stays Zero, DC included (rle.rs:125-127)
                        bits = 0x200u;
                    }
                    W.evbuf[e - e_lo] = ent;
                    atomicOr(&W.slotinfo[slot], bits);
                }
            }
            __syncwarp();

            // ---- classification, lane = slot (rle.rs:94-171) ----
            int key = 4;
            if (in_chunk) {
                const uint32_t info = W.slotinfo[lane];
                const bool inter = (my_sd.y >> 14) & 1u;
                const uint32_t blk = (my_sd.y >> 16) & 0xFFu, code = my_sd.y >> 24;
                const bool ovf = (info & 0x200u) != 0, col = (info & 0x100u) != 0;
                const uint32_t rows = info & 0xFFu;
                const bool has_dc = !inter && code != 0 && !ovf;
                int cls, dcres = 0;
                uint32_t R = 0;
                if (ovf || (!rows && !has_dc)) {
                    cls = CLS_ZERO;
                } else if (!(rows & 0xFEu) && !col) {
                    cls = CLS_DC;
                    // Dc(level): the intra DC, or the one coefficient an inter block put on index 0
                    const int dc = has_dc ? intradc_level((int)code) : (int)W.evbuf[my_start - e_lo] >> 16;
                    dcres = round_residual_dc((float)dc);
                } else {
                    cls = col ? CLS_FULL : CLS_VERT;
                    R = rows | (has_dc ? 1u : 0u);
                }
                W.meta[blk] = (uint32_t)cls | ((uint32_t)lane << 3) | ((uint32_t)dcres << 16);
                W.slotcls[lane] = R | ((uint32_t)cls << 8) | (has_dc ? 0x800u : 0u);
                const int n = __popc(R);
                key = n >= 4 ? 0 : (n >= 2 ? 1 : (n == 1 ? 2 : 3));
            }
            // slots that need the transform first, those with many rows before those with few, so that the
            // four slots of a pass cost about the same
            int pos = s_lo, n_need = 0;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const uint32_t b = __ballot_sync(FULL, key == j);
                if (j < key) pos += __popc(b);
                if (j == key) pos += __popc(b & lt_mask);
                n_need += __popc(b);
            }
            if (key < 3) W.order[pos] = (uint32_t)lane;
            __syncwarp();

            // ---- phase 2: 4 slots per pass, 8 lanes per slot, lane t = column i of the block ---------
            for (int si0 = s_lo; si0 < s_lo + n_need; si0 += 4) {
                const int si = si0 + g;
                const bool valid = si < s_lo + n_need;
                const int sl = valid ? (int)W.order[si] : 0;
                const uint32_t sc = valid ? W.slotcls[sl] : 0u;
                const uint32_t R = sc & 0xFFu;
                const bool vert = ((sc >> 8) & 7u) == CLS_VERT;
                const uint2 sd = W.slotdesc[sl];
                const int nev = valid ? (int)(sd.y & 0xFFu) : 0;
                const uint32_t first = W.sstart[sl] - e_lo;
                // lane t clears row t of the slot, then the slot's events are scattered into it
                *reinterpret_cast<float4*>(c + t * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(c + t * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                const int nevmax = __reduce_max_sync(FULL, nev);
                __syncwarp();
                if (t < nev) {
                    const uint32_t ent = W.evbuf[first + t];
                    c[ent & 63u] = (float)((int)ent >> 16);
                }
                if (nevmax > 8) {  // rare: more than 8 events in a slot of this pass
                    for (int k = t + 8; k < nev; k += 8) {
                        const uint32_t ent = W.evbuf[first + k];
                        c[ent & 63u] = (float)((int)ent >> 16);
                    }
                }
                if (sc & 0x800u) {
                    if (t == 0) c[0] = (float)intradc_level((int)(sd.y >> 24));
                }
                // Per row y that holds a coefficient in ANY of the four slots (warp-uniform skip otherwise):
                //   row pass    t[y][i] = sum_x c[y][x] * B[x][i], ascending x (idct_1d, idct.rs:52-65), pre-divided
                //               by 4 (exact), which takes the /4 of idct.rs:189 out of the 64-output rounding
                //   column pass out[i][j] += t[y][i] * B[y][j] for the 8 pixel rows j, ascending y; B[y][j] are
                //               compile-time constants, t[y][i] never leaves the register
                // A slot that has no coefficient in row y reads zeros there (the slot was cleared), so its terms
                // are +-0 and change nothing -- skipping all-zero rows is exact, computing them is too.
                const uint32_t U = __reduce_or_sync(FULL, R);
                __syncwarp();
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    if (!((U >> y) & 1u)) continue;  // warp-uniform
                    const float4 ca = *reinterpret_cast<const float4*>(c + y * 8);
                    const float4 cc = *reinterpret_cast<const float4*>(c + y * 8 + 4);
                    float a = fmul(ca.x, bt0);  // 0 + x == x
                    a = fadd(a, fmul(ca.y, bt1));
                    a = fadd(a, fmul(ca.z, bt2));
                    a = fadd(a, fmul(ca.w, bt3));
                    a = fadd(a, fmul(cc.x, bt4));
                    a = fadd(a, fmul(cc.y, bt5));
                    a = fadd(a, fmul(cc.z, bt6));
                    a = fadd(a, fmul(cc.w, bt7));
                    if (vert) a = ca.x;  // Vert: the first column feeds idct_1d directly (idct.rs:152-153)
                    a = fmul(a, 0.25f);
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = fadd(acc[j], fmul(a, k_basis(y, j)));
                }
                if (R) {
                    // pixel (x = t, y = j): residual row j of the slot, 16-bit lane of column t in the lane
                    // order of phase 3, (r0,r2) (r1,r3) (r4,r6) (r5,r7)
                    const float m = vert ? H263_B00 : 1.0f;
                    uint16_t* rrow = reinterpret_cast<uint16_t*>(&W.res[sl][0]) + ((t & 1) + 2 * (t >> 2)) * 2 + ((t >> 1) & 1);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int r = round_q(acc[j], m);
                        rrow[j * 8] = (uint16_t)(int16_t)max(min(r, 255), -256);
                    }
                }
                __syncwarp();  // the coefficient slots are reused by the next pass
            }
            s_lo = s_hi;
            e_lo = e_end;
        }
    }

    // the next tile's record words: in flight during the epilogue
    if (kPersistent) next_tile = __shfl_sync(FULL, next_tile, 0);
    if ((kPersistent || kTilesPerWarp > 1) && next_tile < n_tiles && (uint32_t)lane < min((uint32_t)WARP_MBS, n_mbs - next_tile * WARP_MBS) * 6)
        rec = __ldg(mbs32 + (size_t)next_tile * WARP_BLOCKS + lane);

    // ================= phase 3: MC + add + clamp + stores + RGBA, all in registers ===============
    // lane = (macroblock, row group rg of 4 luma rows, h): luma columns 8h..8h+7 of the 4 rows and the
    // two chroma rows under them (8 columns) of ONE plane: Cb for h = 0, Cr for h = 1; the chroma
    // samples the RGBA conversion needs from the other plane come from the neighbour lane.
    {
        const int mbq = lane >> 3, rg = (lane >> 1) & 3, h = lane & 1;
        const bool unit_ok = mbq < n_w;
        const int mbi = unit_ok ? mbq : 0;
        const uint4 ma = *reinterpret_cast<const uint4*>(&W.mb[mbi][0]);
        const uint4 mv = *reinterpret_cast<const uint4*>(&W.mb[mbi][4]);
        const uint32_t flags = (unit_ok ? ma.w : 0u) & ~((H263_ABLATE & 2) ? MBF_INTER : 0u) & ~((H263_ABLATE & 4) ? MBF_RGBA : 0u);
        const uint32_t pitch_y4 = PY ? PY / 4 : (mv.x & 0xFFFFu) >> 2, pitch_c4 = PC ? PC / 4 : mv.x >> 18;
        const int lb = ((rg >> 1) << 1) | h;  // luma block of this unit
        const int r0 = (rg & 1) * 4;          // first row of the unit inside its block
        const int bl = mbi * 6 + lb, bc = mbi * 6 + 4 + h;

        RowSum ly[4];  // luma: 4 rows x 8 pixels in 16-bit lanes
        RowSum cy[2];  // chroma (own plane): 2 rows x 8 pixels
        if (flags & MBF_INTER) {
            const uint32_t fl = W.bf[bl], fc = W.bf[bc];
            if (!WIDE_MV || !(fl & BF_SLOW)) {
                const int sh = (fl & 3u) * 8, shb = sh + ((fl & 4u) << 1);
                const uint32_t wb = (fl >> 3) & 1u, wt = 2u - wb;
                const uint32_t so = W.bd[bl];
                const bool odd = (so & 1u) != 0;  // row pitches are multiples of 16 bytes: the same for every row
                const uint32_t* src = reinterpret_cast<const uint32_t*>(pools.y) + so + (uint32_t)r0 * pitch_y4;
                RowSum hs[5];
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    // the fifth row is only needed for vertical interpolation (its weight is 0 otherwise)
                    const uint32_t* p = src + (uint32_t)r * pitch_y4;
                    uint32_t w0, w1, w2;
                    load_row3(p, odd, (fl & 7u) != 0, r < 4 || wb != 0, w0, w1, w2);
                    hs[r] = row_sum8(w0, w1, w2, sh, shb);
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    ly[r].e0 = vmix(hs[r].e0, hs[r + 1].e0, wt, wb);
                    ly[r].o0 = vmix(hs[r].o0, hs[r + 1].o0, wt, wb);
                    ly[r].e1 = vmix(hs[r].e1, hs[r + 1].e1, wt, wb);
                    ly[r].o1 = vmix(hs[r].o1, hs[r + 1].o1, wt, wb);
                }
            } else {
                // vector beyond the replicated border: clamped per-sample fetch (generic path)
                const PicDev& P = pics[mv.z];
                const uint32_t* r = &W.mbrec[mbi * 6];
                const uint32_t w1 = r[1];
                const uint32_t mvw = lb < 2 ? (r[4] >> (16 * lb)) : (r[5] >> (16 * (lb - 2)));
#pragma unroll
                for (int rr = 0; rr < 4; rr++) {
                    uint32_t o0, o1;
                    mc_fetch8(P.ref[0], P.pitch_y, P.w, P.h, (int)((w1 >> 16) & 0xFF) * 16 + h * 8, (int)(w1 >> 24) * 16 + rg * 4 + rr,
                              (int8_t)(mvw & 0xFF), (int8_t)((mvw >> 8) & 0xFF), o0, o1);
                    ly[rr].e0 = __byte_perm(o0, 0, 0x4240), ly[rr].o0 = __byte_perm(o0, 0, 0x4341);
                    ly[rr].e1 = __byte_perm(o1, 0, 0x4240), ly[rr].o1 = __byte_perm(o1, 0, 0x4341);
                }
            }
            if (!WIDE_MV || !(fc & BF_SLOW)) {
                const int sh = (fc & 3u) * 8, shb = sh + ((fc & 4u) << 1);
                const uint32_t wb = (fc >> 3) & 1u, wt = 2u - wb;
                // chroma rows 2*rg, 2*rg+1 of the macroblock, all 8 columns, plane h
                const uint32_t so = W.bd[bc];
                const bool odd = (so & 1u) != 0;
                const uint32_t* src = reinterpret_cast<const uint32_t*>((H263_ABLATE & 32) ? pools.cb : (h ? pools.cr : pools.cb)) + so + (uint32_t)(rg * 2) * pitch_c4;
                RowSum hs[3];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t* p = src + (uint32_t)r * pitch_c4;
                    uint32_t w0, w1, w2;
                    if (H263_ABLATE & 128) {
                        // load pattern of a lane that owns 4 columns of BOTH planes (results wrong): 2 words per row and plane,
                        // the two h lanes side by side in one sector
                        const uint32_t* pb = reinterpret_cast<const uint32_t*>(pools.cb) + so + (uint32_t)(rg * 2 + r) * pitch_c4 + h;
                        const uint32_t* pr = reinterpret_cast<const uint32_t*>(pools.cr) + so + (uint32_t)(rg * 2 + r) * pitch_c4 + h;
                        w0 = w1 = w2 = 0u;
                        if (r < 2 || wb != 0) {
                            w0 = __ldg(pb), w2 = __ldg(pr);
                            if ((fc & 7u) != 0) w1 = __ldg(pb + 1), w2 ^= __ldg(pr + 1);
                        }
                    } else
                    load_row3(p, odd, (fc & 7u) != 0, r < 2 || wb != 0, w0, w1, w2);
                    hs[r] = row_sum8(w0, w1, w2, sh, shb);
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    cy[r].e0 = vmix(hs[r].e0, hs[r + 1].e0, wt, wb);
                    cy[r].o0 = vmix(hs[r].o0, hs[r + 1].o0, wt, wb);
                    cy[r].e1 = vmix(hs[r].e1, hs[r + 1].e1, wt, wb);
                    cy[r].o1 = vmix(hs[r].o1, hs[r + 1].o1, wt, wb);
                }
            } else {
                const PicDev& P = pics[mv.z];
                const uint32_t* r = &W.mbrec[mbi * 6];
                const uint32_t w1 = r[1], w4 = r[4], w5 = r[5];
                const int sumx = (int8_t)byte_of(w4, 0) + (int8_t)byte_of(w4, 2) + (int8_t)byte_of(w5, 0) + (int8_t)byte_of(w5, 2);
                const int sumy = (int8_t)byte_of(w4, 1) + (int8_t)byte_of(w4, 3) + (int8_t)byte_of(w5, 1) + (int8_t)byte_of(w5, 3);
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    uint32_t o0, o1;
                    mc_fetch8(P.ref[1 + h], P.pitch_c, P.cw, P.ch, (int)((w1 >> 16) & 0xFF) * 8, (int)(w1 >> 24) * 8 + rg * 2 + rr,
                              average_sum_of_mvs(sumx), average_sum_of_mvs(sumy), o0, o1);
                    cy[rr].e0 = __byte_perm(o0, 0, 0x4240), cy[rr].o0 = __byte_perm(o0, 0, 0x4341);
                    cy[rr].e1 = __byte_perm(o1, 0, 0x4240), cy[rr].o1 = __byte_perm(o1, 0, 0x4341);
                }
            }
        } else {
            // intra: the prediction is the zero-initialised plane (picture.rs:42-48)
#pragma unroll
            for (int r = 0; r < 4; r++) ly[r] = RowSum{0, 0, 0, 0};
            cy[0] = cy[1] = RowSum{0, 0, 0, 0};
        }

        // ---- residuals: packed s16x2 add, saturate to [0, 255] (idct.rs:191-194) ----
        if (unit_ok) {
            const uint32_t m = W.meta[bl];
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
                const uint32_t rs = (m >> 3) & 31u;
                const uint32_t* c = &W.res[rs][0];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(c + (r0 + r) * 4);
                    ly[r].e0 = __viaddmin_s16x2_relu(ly[r].e0, rv.x, 0x00FF00FFu);
                    ly[r].o0 = __viaddmin_s16x2_relu(ly[r].o0, rv.y, 0x00FF00FFu);
                    ly[r].e1 = __viaddmin_s16x2_relu(ly[r].e1, rv.z, 0x00FF00FFu);
                    ly[r].o1 = __viaddmin_s16x2_relu(ly[r].o1, rv.w, 0x00FF00FFu);
                }
            }
            const uint32_t mc = W.meta[bc];
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
                const uint32_t rs = (mc >> 3) & 31u;
                const uint32_t* c = &W.res[rs][0];
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(c + (rg * 2 + r) * 4);
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
        if constexpr (EDGE) {
            // keep `lo` bytes of a, take the rest from b (lo = 0..4)
            auto merge = [](uint32_t a, uint32_t b, int lo) {
                const uint32_t m = lo >= 4 ? 0xFFFFFFFFu : ((1u << (8 * lo)) - 1u);
                return (a & m) | (b & ~m);
            };
            const uint32_t ed = mv.w;
            const int vw = (int)(ed & 15u), vh = (int)((ed >> 4) & 15u), cvw = (int)((ed >> 8) & 7u), cvh = (int)((ed >> 12) & 7u);
            const bool fix_r = (flags & MBF_RIGHT) != 0, fix_b = (flags & MBF_BOTTOM) != 0;
            // right edge, luma: the pixel of column vw - 1 lives in the lane with h = (vw - 1) >> 3 of this row group
            {
                const int e = (vw - 1) & 15, keep = min(max(vw - 8 * h, 0), 8);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t mine = __byte_perm(yw[r][0], yw[r][1], (uint32_t)(e & 7)) & 0xFFu;
                    const uint32_t other = __shfl_xor_sync(FULL, mine, 1);
                    if (fix_r && vw) {
                        const uint32_t sp = splat_lo((e >> 3) == h ? mine : other);
                        yw[r][0] = merge(yw[r][0], sp, min(keep, 4));
                        yw[r][1] = merge(yw[r][1], sp, max(keep - 4, 0));
                    }
                }
            }
            // right edge, chroma: every lane holds all 8 columns of its plane
            if (fix_r && cvw) {
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint32_t sp = splat_lo(__byte_perm(cw[r][0], cw[r][1], (uint32_t)(cvw - 1)));
                    cw[r][0] = merge(cw[r][0], sp, min(cvw, 4));
                    cw[r][1] = merge(cw[r][1], sp, max(cvw - 4, 0));
                }
            }
            // bottom edge: rows below row vh - 1 (chroma cvh - 1) repeat it; its owner is the lane of the same
            // macroblock and h with rg = (vh - 1) >> 2 (chroma (cvh - 1) >> 1)
            {
                const int f = (vh - 1) & 15, rf = f & 3;
                const uint32_t own0 = rf == 0 ? yw[0][0] : (rf == 1 ? yw[1][0] : (rf == 2 ? yw[2][0] : yw[3][0]));
                const uint32_t own1 = rf == 0 ? yw[0][1] : (rf == 1 ? yw[1][1] : (rf == 2 ? yw[2][1] : yw[3][1]));
                const int src = (lane & ~6) | ((f >> 2) << 1);
                const uint32_t s0 = __shfl_sync(FULL, own0, src), s1 = __shfl_sync(FULL, own1, src);
                if (fix_b && vh) {
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (rg * 4 + r > f) yw[r][0] = s0, yw[r][1] = s1;
                }
                const int cf = (cvh - 1) & 7;
                const uint32_t c0 = (cf & 1) ? cw[1][0] : cw[0][0], c1 = (cf & 1) ? cw[1][1] : cw[0][1];
                const int csrc = (lane & ~6) | ((cf >> 1) << 1);
                const uint32_t t0 = __shfl_sync(FULL, c0, csrc), t1 = __shfl_sync(FULL, c1, csrc);
                if (fix_b && cvh) {
#pragma unroll
                    for (int r = 0; r < 2; r++)
                        if (rg * 2 + r > cf) cw[r][0] = t0, cw[r][1] = t1;
                }
            }
        }
        if (unit_ok && !(H263_ABLATE & 8)) {
            uint8_t* py = pools.y + (size_t)(ma.x + (uint32_t)(rg * 4) * pitch_y4 + (uint32_t)(h * 2)) * 4;
            uint8_t* pc = ((H263_ABLATE & 64) ? pools.cb : (h ? pools.cr : pools.cb)) + (size_t)(ma.y + (uint32_t)(rg * 2) * pitch_c4) * 4;
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
                const uint32_t rgba_pitch = PR ? PR : mv.y;
                uint8_t* o = pools.rgba + (size_t)ma.z * 16 + (size_t)(rg * 4) * rgba_pitch + (size_t)(h * 32);
                const int kr = opaque(104597), kg = opaque(-53279), kb = opaque(132201);
#pragma unroll
#pragma unroll
                for (int cr2 = 0; cr2 < 2; cr2++) {
                    // chroma row cr2 serves luma rows 2*cr2, 2*cr2+1; sample j serves pixels 2j, 2j+1
                    const CT t0 = chroma_terms_folded((int)(ce[0][cr2] & 0xFFFFu), (int)(ce[1][cr2] & 0xFFFFu), kr, kg, kb);
                    const CT t1 = chroma_terms_folded((int)(co[0][cr2] & 0xFFFFu), (int)(co[1][cr2] & 0xFFFFu), kr, kg, kb);
                    const CT t2 = chroma_terms_folded((int)(ce[0][cr2] >> 16), (int)(ce[1][cr2] >> 16), kr, kg, kb);
                    const CT t3 = chroma_terms_folded((int)(co[0][cr2] >> 16), (int)(co[1][cr2] >> 16), kr, kg, kb);
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
                        if (H263_ABLATE & 16) {  // RGBA computed, store suppressed (kept alive by an impossible condition)
                            if ((px[0] ^ px[1] ^ px[2] ^ px[3] ^ px[4] ^ px[5] ^ px[6] ^ px[7]) == 0x12345678u)
                                st_global_v8(o + (size_t)(cr2 * 2 + rr) * rgba_pitch, px);
                        } else
                        st_global_v8(o + (size_t)(cr2 * 2 + rr) * rgba_pitch, px);
                    }
                }
            }
        }
    }
    tile = next_tile;
  }  // tile loop
}

void launch_recon_tile(const PicDev* pics, const h263cu_mb* mbs, const h263cu_event* events, uint32_t n_mbs, int emit_rgba,
                       int unaligned, int wide_mv, const Pools& pools, cudaStream_t stream) {
    if (n_mbs == 0) return;
    const uint32_t per_cta = CTA_WARPS * WARP_MBS;
    uint32_t grid = (n_mbs + per_cta * kTilesPerWarp - 1) / (per_cta * kTilesPerWarp);
    if (kPersistent) {
        static int sm_count[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && sm_count[dev] == 0) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        const uint32_t resident = (uint32_t)(dev >= 0 && dev < 64 && sm_count[dev] > 0 ? sm_count[dev] : 148) * (32u / CTA_WARPS);
        grid = min((n_mbs + per_cta - 1) / per_cta, resident);
        cudaMemsetAsync(pools.work_counter, 0, sizeof(uint32_t), stream);  // tiles beyond the grid's first ones
    }
    // pitches of the standard formats (context.cu: pitch_y = 16*mbw + 64, pitch_c = round_up(8*mbw + 32, 16))
    const uint32_t py = pools.pitch_y, pc = pools.pitch_c, pr = pools.rgba_pitch;
    if (wide_mv) {  // hand-built side info with vectors beyond the range: clamped per-sample path, run-time pitches
        if (unaligned)
            recon_tile_kernel<0, 0, 0, true, true><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
        else
            recon_tile_kernel<0, 0, 0, false, true><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
    } else if (unaligned)  // some picture of the step is not a multiple of 16 in size: edge fix-up, run-time pitches
        recon_tile_kernel<0, 0, 0, true, false><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
    else if (py == 416 && pc == 208 && pr == 1408)  // CIF 352x288
        recon_tile_kernel<416, 208, 1408, false, false><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
    else if (py == 240 && pc == 128 && pr == 704)  // QCIF 176x144
        recon_tile_kernel<240, 128, 704, false, false><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
    else if (py == 768 && pc == 384 && pr == 2816)  // 4CIF 704x576
        recon_tile_kernel<768, 384, 2816, false, false><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
    else
        recon_tile_kernel<0, 0, 0, false, false><<<grid, CTA_THREADS, 0, stream>>>(pics, mbs, events, n_mbs, emit_rgba, pools);
}

}  // namespace h263dev
