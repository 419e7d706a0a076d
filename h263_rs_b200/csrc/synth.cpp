// synth.cpp -- synthetic Sorenson-flavour (and baseline H.263) bitstream generator.
//
// The reference has no encoder and ships no sample streams, so the
// benchmark and the parity tests need one.  It emits one byte-aligned, zero-padded packet
// per picture (the reference's Sorenson path ends a picture only at EOF, state.rs:193,411),
// I and P pictures with random but VALID syntax: MCBPC/CBPY/DQUANT/MVD/TCOEF codes from the
// same code tables the parser uses (vlc_codes.inc), escapes in all three widths, 4MV
// macroblocks, motion vectors that cross the picture borders, zig-zag overflows and
// truncated pictures.  It carries no picture content: coefficients are random draws whose
// statistics (events per block, level magnitudes, MB type mix) are parameters, reported
// with every benchmark number.  PRNG = splitmix64, so streams are reproducible anywhere.
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/h263synth.h"
#include "bitio.hpp"

// libh263synth.so stands alone: it carries its own copy of the code tables (the product keeps its copy in frontend.cpp)
namespace h263fe {
#include "vlc_codes.inc"
const VlcCode* vlc_codes(int id, int* count) {
    switch (id) {
        case T_MCBPC_I: *count = MCBPC_I_CODES_COUNT; return MCBPC_I_CODES;
        case T_MCBPC_P: *count = MCBPC_P_CODES_COUNT; return MCBPC_P_CODES;
        case T_CBPY: *count = CBPY_CODES_COUNT; return CBPY_CODES;
        case T_MVD: *count = MVD_CODES_COUNT; return MVD_CODES;
        default: *count = TCOEF_CODES_COUNT; return TCOEF_CODES;
    }
}
}  // namespace h263fe

using namespace h263fe;

namespace {
constexpr int H263CU_ERR_BAD_ARGUMENT = -100;  // the generator reports errors with the product library's codes

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return n ? (uint32_t)(next() % n) : 0; }
    bool pct(uint32_t p) { return below(100) < p; }
    bool permille(uint32_t p) { return below(1000) < p; }
    // geometric: number of failures before a success with success probability num/den
    uint32_t geometric(uint32_t num, uint32_t den, uint32_t cap) {
        uint32_t k = 0;
        while (k < cap && below(den) >= num) k++;
        return k;
    }
    int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }
};

struct EncTables {
    const char* mcbpc_i[6][2][2] = {};
    const char* mcbpc_p[6][2][2] = {};
    const char* mcbpc_i_stuffing = nullptr;
    const char* mcbpc_p_stuffing = nullptr;
    const char* cbpy[16] = {};
    const char* mvd[64] = {};          // index value + 32
    const char* tcoef[2][64][13] = {}; // [last][run][level]
    const char* tcoef_escape = nullptr;
    EncTables() {
        int n;
        const VlcCode* c = vlc_codes(T_MCBPC_I, &n);
        for (int i = 0; i < n; i++) {
            if (c[i].kind == 0) mcbpc_i[c[i].a][c[i].b][c[i].c] = c[i].bits;
            if (c[i].kind == 1) mcbpc_i_stuffing = c[i].bits;
        }
        c = vlc_codes(T_MCBPC_P, &n);
        for (int i = 0; i < n; i++) {
            if (c[i].kind == 0) mcbpc_p[c[i].a][c[i].b][c[i].c] = c[i].bits;
            if (c[i].kind == 1) mcbpc_p_stuffing = c[i].bits;
        }
        c = vlc_codes(T_CBPY, &n);
        for (int i = 0; i < n; i++)
            if (c[i].kind == 0) cbpy[c[i].a] = c[i].bits;
        c = vlc_codes(T_MVD, &n);
        for (int i = 0; i < n; i++)
            if (c[i].kind == 0) mvd[c[i].a + 32] = c[i].bits;
        c = vlc_codes(T_TCOEF, &n);
        for (int i = 0; i < n; i++) {
            if (c[i].kind == 0) tcoef[c[i].a][c[i].b][c[i].c] = c[i].bits;
            if (c[i].kind == 3) tcoef_escape = c[i].bits;
        }
    }
};
const EncTables& enc() {
    static const EncTables E;
    return E;
}

struct Mv {
    int x, y;
};
inline int median3(int a, int b, int c) {
    int lo = std::min(a, b), hi = std::max(a, b);
    return std::max(lo, std::min(hi, c));
}
// MVD that makes the decoder's wrap rule (mvd_pred.rs:77-116) reconstruct `target`
inline int mvd_for(int target, int pred) {
    int d = target - pred;
    if (d > 31) d -= 64;
    if (d < -32) d += 64;
    return d;
}

struct Gen {
    const h263cu_synth_params& P;
    Rng rng;
    uint32_t mb_w, mb_h;
    std::vector<Mv> mvs;
    explicit Gen(const h263cu_synth_params& p) : P(p), rng(p.seed * 0x9E3779B97F4A7C15ull + 0x1234567) {
        mb_w = (p.width + 15) / 16;
        mb_h = (p.height + 15) / 16;
    }

    void header(BitWriter& w, uint32_t tr, bool intra, uint32_t quant, bool disposable = false) {
        w.put(1, 17);  // PSC: sixteen zeros and a one
        if (P.flavour == 0) {
            w.put(P.version & 31, 5);
            w.put(tr & 255, 8);
            uint32_t W = P.width, H = P.height;
            if (W == 352 && H == 288)
                w.put(2, 3);
            else if (W == 176 && H == 144)
                w.put(3, 3);
            else if (W == 128 && H == 96)
                w.put(4, 3);
            else if (W == 320 && H == 240)
                w.put(5, 3);
            else if (W == 160 && H == 120)
                w.put(6, 3);
            else if (W < 256 && H < 256) {
                w.put(0, 3);
                w.put(W, 8);
                w.put(H, 8);
            } else {
                w.put(1, 3);
                w.put(W, 16);
                w.put(H, 16);
            }
            w.put(intra ? 0 : (disposable ? 2 : 1), 2);  // Sorenson picture type: 0 I, 1 P, 2 disposable P (picture.rs:319-325)
            w.put(P.deblock_flag ? 1 : 0, 1);
            w.put(quant, 5);
            w.put(0, 1);  // PEI
        } else {
            w.put(0, 5);  // GN = 0
            w.put(tr & 255, 8);
            uint32_t fmt = 3;
            if (P.width == 128) fmt = 1;
            if (P.width == 176) fmt = 2;
            if (P.width == 352) fmt = 3;
            if (P.width == 704) fmt = 4;
            if (P.width == 1408) fmt = 5;
            w.put(0x80 | fmt, 8);           // "10", no split screen / camera / freeze, source format
            w.put(intra ? 0x10 : 0x00, 5);  // this decoder reads the set bit as INTRA (picture.rs:57-61)
            w.put(quant, 5);
            w.put(0, 1);  // CPM
            w.put(0, 1);  // PEI
        }
    }

    // One coded block's TCOEF events.
    void events(BitWriter& w, bool intra, int quant) {
        const EncTables& E = enc();
        int idx = intra ? 1 : 0;
        bool overflow = rng.permille(P.permille_overflow);
        uint32_t mean10 = std::max<uint32_t>(P.mean_events_x10, 10);
        // 1 + geometric with mean (mean-1): success probability 10/mean10
        int n = 1 + (int)rng.geometric(10, mean10, 62);
        if (overflow && n < 2) n = 2;  // the first event of an inter block sits at index 0
        for (int k = 0; k < n; k++) {
            int remaining = 63 - idx;  // highest legal run from here
            bool last = k == n - 1;
            int run = (int)rng.geometric(2, 3, 6);
            if (rng.pct(10)) run += rng.range(0, 12);
            if (!overflow) {
                if (run > remaining) run = remaining;
                if (idx + run >= 63) last = true;  // no room for another event
            } else if (last) {
                run = 63;  // pushes the zig-zag index past 63: the whole block is dropped (rle.rs:125-127)
            } else if (run > remaining) {
                run = std::max(remaining, 0);
            }
            int mag = 1 + (int)rng.geometric(3, 5, 10);
            bool escape = rng.pct(P.pct_escape);
            if (escape && rng.pct(40)) mag += rng.range(10, 100);
            if (escape && rng.pct(10)) mag += rng.range(100, 900);
            // keep QP*(2|level|+1) inside i16
            int mag_cap = (32767 / std::max(quant, 1) - 1) / 2;
            mag = std::min(mag, mag_cap);
            int sign = (int)rng.below(2);
            const char* code = (!escape && run < 64 && mag <= 12) ? E.tcoef[last ? 1 : 0][run][mag] : nullptr;
            if (code) {
                w.put_code(code);
                w.put((uint32_t)sign, 1);
            } else {
                int level = sign ? -mag : mag;
                w.put_code(E.tcoef_escape);
                if (P.flavour == 0 && P.version == 1) {
                    if (mag > 63) {
                        level = std::max(std::min(level, 1023), -1023);
                        w.put(1, 1);
                        w.put(last ? 1 : 0, 1);
                        w.put((uint32_t)run, 6);
                        w.put((uint32_t)level & 0x7FF, 11);
                    } else {
                        w.put(0, 1);
                        w.put(last ? 1 : 0, 1);
                        w.put((uint32_t)run, 6);
                        w.put((uint32_t)level & 0x7F, 7);
                    }
                } else {
                    level = std::max(std::min(level, 127), -127);
                    w.put(last ? 1 : 0, 1);
                    w.put((uint32_t)run, 6);
                    w.put((uint32_t)level & 0xFF, 8);
                }
            }
            idx += run + 1;
            if (last) break;
        }
    }

    int pick_dc() {
        int v = 128 + rng.range(-40, 40) + rng.range(-40, 40);
        v = std::max(1, std::min(254, v));
        if (v == 128) v = 129;
        if (rng.permille(3)) v = 255;  // the 0xFF => 1024 special case (types.rs:955-961)
        return v;
    }

    Mv pick_mv(uint32_t col, uint32_t row) {
        Mv m;
        if (P.mv_mode == 1) {
            m.x = rng.range(-32, 31), m.y = rng.range(-32, 31);
        } else if (P.mv_mode == 2) {
            bool edge = col == 0 || row == 0 || col == mb_w - 1 || row == mb_h - 1;
            if (edge && rng.pct(70)) {
                m.x = col == 0 ? rng.range(-32, -1) : (col == mb_w - 1 ? rng.range(1, 31) : rng.range(-8, 8));
                m.y = row == 0 ? rng.range(-32, -1) : (row == mb_h - 1 ? rng.range(1, 31) : rng.range(-8, 8));
            } else {
                m.x = rng.range(-32, 31), m.y = rng.range(-32, 31);
            }
        } else {
            m.x = rng.range(-3, 3) + rng.range(-3, 3);
            m.y = rng.range(-3, 3) + rng.range(-3, 3);
            if (rng.pct(8)) m.x = rng.range(-32, 31), m.y = rng.range(-32, 31);
        }
        return m;
    }

    void picture(BitWriter& w, uint32_t index, bool intra_pic) {
        const EncTables& E = enc();
        uint32_t qlo = std::max(1u, std::min(31u, P.qp_min)), qhi = std::max(qlo, std::min(31u, P.qp_max));
        int quant = rng.range((int)qlo, (int)qhi);
        // drawn only when asked for, so that streams without disposable pictures stay what they were
        const bool disposable = !intra_pic && P.flavour == 0 && P.pct_disposable && rng.pct(P.pct_disposable);
        header(w, index, intra_pic, (uint32_t)quant, disposable);
        const uint32_t n_mb = mb_w * mb_h;
        mvs.assign((size_t)n_mb * 4, Mv{0, 0});
        uint32_t stop_at = n_mb;
        if (!intra_pic && rng.permille(P.truncate_permille)) stop_at = rng.below(n_mb);
        for (uint32_t n = 0; n < stop_at; n++) {
            const uint32_t col = n % mb_w, row = n / mb_w;
            if (rng.permille(2)) {  // MCBPC stuffing: takes no macroblock slot (state.rs:206)
                if (!intra_pic) w.put(0, 1);
                w.put_code(intra_pic ? E.mcbpc_i_stuffing : E.mcbpc_p_stuffing);
            }
            int type;  // 0 Inter 1 InterQ 2 Inter4V 3 Intra 4 IntraQ 5 Inter4Vq
            if (intra_pic) {
                type = 3;
            } else {
                uint32_t r = rng.below(100);
                if (r < P.pct_uncoded) {
                    w.put(1, 1);  // COD = 1
                    continue;
                }
                w.put(0, 1);
                r = rng.below(100);
                if (r < P.pct_intra)
                    type = 3;
                else if (r < P.pct_intra + P.pct_fourmv)
                    type = 2;
                else
                    type = 0;
            }
            int dq = 0;
            if (rng.pct(P.pct_dquant)) {
                static const int DQ[4] = {-1, -2, 1, 2};
                int pick = (int)rng.below(4);
                // keep the in-force quantiser inside [qlo, qhi] so the statistics stay put
                int nq = std::min(std::max(quant + DQ[pick], 1), 31);
                if (nq >= (int)qlo && nq <= (int)qhi) {
                    dq = DQ[pick];
                    type = type == 0 ? 1 : (type == 3 ? 4 : (type == 2 ? 5 : type));
                    // Inter4Vq only has a code in the long MCBPC tail; allowed, rare
                }
            }
            const bool intra = type == 3 || type == 4;
            bool cbp[6];
            for (int b = 0; b < 6; b++) cbp[b] = rng.pct(intra ? P.pct_cbp_intra : P.pct_cbp_inter);
            const char* mc = intra_pic ? E.mcbpc_i[type][cbp[4]][cbp[5]] : E.mcbpc_p[type][cbp[4]][cbp[5]];
            w.put_code(mc);
            int y = (cbp[0] << 3) | (cbp[1] << 2) | (cbp[2] << 1) | (int)cbp[3];
            w.put_code(E.cbpy[intra ? y : (~y & 15)]);
            if (type == 1 || type == 4 || type == 5) {
                static const int CODE[5] = {1, 0, -1, 2, 3};  // dq -2,-1,(0),+1,+2 -> 01,00,10,11
                w.put((uint32_t)CODE[dq + 2], 2);
                quant = std::min(std::max(quant + dq, 1), 31);
            }
            if (!intra) {
                const bool four = type == 2 || type == 5;
                Mv* cur = &mvs[(size_t)n * 4];
                Mv base = pick_mv(col, row);
                const Mv zero{0, 0};
                for (int k = 0; k < (four ? 4 : 1); k++) {
                    Mv c1, c2, c3;
                    if (k == 0 || k == 2)
                        c1 = col == 0 ? zero : mvs[(size_t)(n - 1) * 4 + k + 1];
                    else
                        c1 = cur[k - 1];
                    if (k < 2) {
                        c2 = row == 0 ? c1 : mvs[(size_t)(n - mb_w) * 4 + k + 2];
                        if (col == mb_w - 1)
                            c3 = zero;
                        else if (row == 0)
                            c3 = c1;
                        else
                            c3 = mvs[(size_t)(n - mb_w + 1) * 4 + 2];
                    } else {
                        c2 = cur[0], c3 = cur[1];
                    }
                    int px = median3(c1.x, c2.x, c3.x), py = median3(c1.y, c2.y, c3.y);
                    Mv t = base;
                    if (k > 0) {
                        t.x = std::min(31, std::max(-32, base.x + rng.range(-3, 3)));
                        t.y = std::min(31, std::max(-32, base.y + rng.range(-3, 3)));
                    }
                    w.put_code(E.mvd[mvd_for(t.x, px) + 32]);
                    w.put_code(E.mvd[mvd_for(t.y, py) + 32]);
                    cur[k] = t;
                }
                if (!four) cur[1] = cur[2] = cur[3] = cur[0];
            }
            for (int b = 0; b < 6; b++) {
                if (intra) w.put((uint32_t)pick_dc(), 8);
                if (cbp[b]) events(w, intra, quant);
            }
        }
        w.align_zero();
    }
};

}  // namespace

extern "C" {

void h263cu_synth_default_params(h263cu_synth_params* p, uint32_t width, uint32_t height, uint32_t n_pictures,
                                 uint64_t seed) {
    std::memset(p, 0, sizeof(*p));
    p->width = width, p->height = height, p->n_pictures = n_pictures, p->seed = seed;
    p->flavour = 0;
    p->version = 1;
    p->intra_period = 0;
    p->deblock_flag = 0;
    p->qp_min = 2, p->qp_max = 12;
    p->pct_uncoded = 25, p->pct_intra = 10, p->pct_fourmv = 5;
    p->pct_dquant = 5;
    p->pct_cbp_inter = 40, p->pct_cbp_intra = 70;
    p->mean_events_x10 = 35;
    p->pct_escape = 2;
    p->permille_overflow = 1;
    p->mv_mode = 0;
    p->truncate_permille = 0;
}

int64_t h263cu_synth_stream(const h263cu_synth_params* p, uint8_t* out, size_t cap, uint64_t* pkt_off,
                            uint32_t* pkt_len) {
    if (!p || p->width == 0 || p->height == 0 || p->width > 4080 || p->height > 4080) return H263CU_ERR_BAD_ARGUMENT;
    if (p->pct_uncoded > 100 || p->pct_intra + p->pct_fourmv > 100) return H263CU_ERR_BAD_ARGUMENT;
    if (p->flavour == 1) {
        bool ok = (p->width == 128 && p->height == 96) || (p->width == 176 && p->height == 144) ||
                  (p->width == 352 && p->height == 288) || (p->width == 704 && p->height == 576) ||
                  (p->width == 1408 && p->height == 1152);
        if (!ok) return H263CU_ERR_BAD_ARGUMENT;
    }
    Gen g(*p);
    uint64_t total = 0;
    for (uint32_t i = 0; i < p->n_pictures; i++) {
        BitWriter w;
        bool intra = i == 0 || (p->intra_period && i % p->intra_period == 0);
        g.picture(w, i, intra);
        if (out && total + w.bytes.size() <= cap) std::memcpy(out + total, w.bytes.data(), w.bytes.size());
        if (pkt_off) pkt_off[i] = total;
        if (pkt_len) pkt_len[i] = (uint32_t)w.bytes.size();
        total += w.bytes.size();
    }
    return (int64_t)total;
}

}  // extern "C"
