/*
 * h263_oracle.cpp -- CPU oracle: a literal restatement of ruffle-rs/h263-rs.
 *
 * TEST INFRASTRUCTURE ONLY (see h263_oracle.h).  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Build with
 *   g++ -O2 -ffp-contract=off   (never -ffast-math)
 * because the reference never fuses multiply-adds and the host CPU has FMA.
 * Integer arithmetic that the reference performs in i16 uses explicit int16_t
 * truncation = Rust release-mode wrapping (rle.rs:130-133).
 *
 * Supported syntax: Sorenson Spark pictures and baseline H.263 PTYPE pictures.
 * PLUSPTYPE headers return UnimplementedDecoding (the reference parses them, but
 * every option they can enable is unimplemented downstream); everything else follows
 * the reference including its error values.
 */
#include "h263_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct VlcCode {
    const char* bits;
    int len;
    int kind;
    int a;
    int b;
    int c;
};
#include "vlc_codes.inc"

typedef int Err;

/* ------------------------------------------------------------------------------------
 * VLC trees: built once from the flat code list; walked one bit at a time exactly like
 * H263Reader::read_vlc (reader.rs:272-290).
 * ---------------------------------------------------------------------------------- */
struct VlcNode {
    int child[2];  // -1 = absent
    int leaf;      // index into codes, -1 = interior
};
struct VlcTree {
    std::vector<VlcNode> nodes;
    const VlcCode* codes;
    void build(const VlcCode* c, int n) {
        codes = c;
        nodes.clear();
        nodes.push_back({{-1, -1}, -1});
        for (int i = 0; i < n; i++) {
            int cur = 0;
            for (int k = 0; k < c[i].len; k++) {
                int bit = c[i].bits[k] - '0';
                if (nodes[cur].child[bit] < 0) {
                    nodes[cur].child[bit] = (int)nodes.size();
                    nodes.push_back({{-1, -1}, -1});
                }
                cur = nodes[cur].child[bit];
            }
            nodes[cur].leaf = i;
        }
    }
};
VlcTree g_trees[5];
std::atomic<int> g_trees_ready{0};
void ensure_trees() {
    static std::atomic<int> lock{0};
    if (g_trees_ready.load(std::memory_order_acquire)) return;
    while (lock.exchange(1)) {
    }
    if (!g_trees_ready.load()) {
        g_trees[0].build(MCBPC_I_CODES, MCBPC_I_CODES_COUNT);
        g_trees[1].build(MCBPC_P_CODES, MCBPC_P_CODES_COUNT);
        g_trees[2].build(CBPY_CODES, CBPY_CODES_COUNT);
        g_trees[3].build(MVD_CODES, MVD_CODES_COUNT);
        g_trees[4].build(TCOEF_CODES, TCOEF_CODES_COUNT);
        g_trees_ready.store(1, std::memory_order_release);
    }
    lock.store(0);
}

/* ------------------------------------------------------------------------------------
 * H263Reader over an in-memory source (parser/reader.rs:15-441).  With a `&[u8]`
 * source the VecDeque buffering is unobservable: a read of n bits fails with
 * UnexpectedEof iff fewer than n bits remain (reader.rs:49-75).
 * ---------------------------------------------------------------------------------- */
struct Reader {
    const uint8_t* data;
    size_t len;
    size_t bits_read;

    // peek_bits (reader.rs:94-134)
    Err peek_bits(uint32_t n, uint64_t* out) const {
        if (n == 0) {
            *out = 0;
            return ORC_OK;
        }
        if (bits_read + n > len * 8) return ORC_ERR_UNHANDLED_IO_ERROR;
        uint64_t acc = 0;
        size_t p = bits_read;
        for (uint32_t i = 0; i < n; i++, p++) acc = (acc << 1) | ((data[p >> 3] >> (7 - (p & 7))) & 1);
        *out = acc;
        return ORC_OK;
    }
    // skip_bits (reader.rs:141-147)
    Err skip_bits(uint32_t n) {
        if (bits_read + n > len * 8) return ORC_ERR_UNHANDLED_IO_ERROR;
        bits_read += n;
        return ORC_OK;
    }
    // read_bits (reader.rs:159-164)
    Err read_bits(uint32_t n, uint64_t* out) {
        Err e = peek_bits(n, out);
        if (e) return e;
        return skip_bits(n);
    }
    // read_signed_bits (reader.rs:179-206), result sign-extended to 64 bits
    Err read_signed_bits(uint32_t n, int64_t* out) {
        uint64_t v;
        Err e = read_bits(n, &v);
        if (e) return e;
        if (n > 0 && n < 64 && ((v >> (n - 1)) & 1)) v |= ~uint64_t(0) << n;
        *out = (int64_t)v;
        return ORC_OK;
    }
    Err read_u8(uint8_t* out) {
        uint64_t v;
        Err e = read_bits(8, &v);
        if (e) return e;
        *out = (uint8_t)v;
        return ORC_OK;
    }
    // realignment_bits (reader.rs:215-217)
    uint32_t realignment_bits() const { return (8 - (uint32_t)(bits_read % 8)) % 8; }

    // recognize_start_code (reader.rs:240-258); *found = false for None.
    Err recognize_start_code(bool in_error, bool* found, uint32_t* skipped) {
        size_t checkpoint = bits_read;  // with_lookahead (reader.rs:429-440)
        uint32_t max_skip = realignment_bits();
        uint32_t skip = 0;
        uint64_t maybe;
        Err e = peek_bits(17, &maybe);
        *found = false;
        while (!e && maybe != 1) {
            if (!in_error && skip > max_skip) {
                bits_read = checkpoint;
                return ORC_OK;
            }
            e = skip_bits(1);
            if (e) break;
            skip += 1;
            e = peek_bits(17, &maybe);
        }
        bits_read = checkpoint;
        if (e) return e;
        *found = true;
        *skipped = skip;
        return ORC_OK;
    }

    // read_vlc (reader.rs:272-290): one read_bits(1) per tree level.
    Err read_vlc(const VlcTree& t, const VlcCode** out) {
        int idx = 0;
        for (;;) {
            const VlcNode& nd = t.nodes[idx];
            if (nd.leaf >= 0) {
                *out = &t.codes[nd.leaf];
                return ORC_OK;
            }
            uint64_t bit;
            Err e = read_bits(1, &bit);
            if (e) return e;
            idx = nd.child[bit];
            if (idx < 0) return ORC_ERR_INTERNAL_DECODER_ERROR;
        }
    }
};

/* ------------------------------------------------------------------------------------
 * Types (types.rs)
 * ---------------------------------------------------------------------------------- */
enum PicType { PT_I, PT_P, PT_PB, PT_IMPROVED_PB, PT_B, PT_EI, PT_EP, PT_RESERVED, PT_DISPOSABLE_P };

// PictureOption bits (types.rs:195-217)
enum : uint32_t {
    PO_SPLIT_SCREEN = 1u << 0,
    PO_DOCUMENT_CAMERA = 1u << 1,
    PO_RELEASE_FREEZE = 1u << 2,
    PO_UMV = 1u << 3,
    PO_SAC = 1u << 4,
    PO_AP = 1u << 5,
    PO_AIC = 1u << 6,
    PO_DEBLOCKING_FILTER = 1u << 7,
    PO_SLICE_STRUCTURED = 1u << 8,
    PO_RPS = 1u << 9,
    PO_ISD = 1u << 10,
    PO_ALT_INTER_VLC = 1u << 11,
    PO_MODIFIED_QUANT = 1u << 12,
    PO_RPR = 1u << 13,
    PO_RRU = 1u << 14,
    PO_ROUNDING_TYPE_ONE = 1u << 15,
    PO_USE_DEBLOCKER = 1u << 16,
};
// OPPTYPE_OPTIONS / MPPTYPE_OPTIONS (types.rs:220-241)
const uint32_t OPPTYPE_OPTIONS = PO_UMV | PO_SAC | PO_AP | PO_AIC | PO_DEBLOCKING_FILTER |
                                 PO_SLICE_STRUCTURED | PO_RPS | PO_ISD | PO_ALT_INTER_VLC |
                                 PO_MODIFIED_QUANT;
const uint32_t MPPTYPE_OPTIONS = PO_RPR | PO_RRU | PO_ROUNDING_TYPE_ONE;

// SourceFormat (types.rs:136-181): kind 0..5 standard, 6 = Extended(w,h); Reserved = 5.
struct SourceFormat {
    int kind;  // 0 SubQcif 1 QuarterCif 2 FullCif 3 FourCif 4 SixteenCif 5 Reserved 6 Extended
    uint16_t w, h;
    bool operator==(const SourceFormat& o) const {
        return kind == o.kind && (kind != 6 || (w == o.w && h == o.h));
    }
    // into_width_and_height (types.rs:168-180)
    bool dims(uint16_t* ow, uint16_t* oh) const {
        switch (kind) {
            case 0: *ow = 128, *oh = 96; return true;
            case 1: *ow = 176, *oh = 144; return true;
            case 2: *ow = 352, *oh = 288; return true;
            case 3: *ow = 704, *oh = 576; return true;
            case 4: *ow = 1408, *oh = 1152; return true;
            case 6: *ow = w, *oh = h; return true;
            default: return false;
        }
    }
};

// Picture (types.rs:13-122), the fields the decode path can observe.
struct Picture {
    int version = -1;  // None
    uint16_t temporal_reference = 0;
    bool has_format = false;
    SourceFormat format{5, 0, 0};
    uint32_t options = 0;
    bool has_plusptype = false;
    bool has_opptype = false;
    PicType picture_type = PT_I;
    uint8_t quantizer = 0;
};

enum MbType { MB_INTER = 0, MB_INTERQ = 1, MB_INTER4V = 2, MB_INTRA = 3, MB_INTRAQ = 4, MB_INTER4VQ = 5 };
inline bool mb_is_inter(int t) { return t == MB_INTER || t == MB_INTERQ || t == MB_INTER4V || t == MB_INTER4VQ; }
inline bool mb_is_intra(int t) { return t == MB_INTRA || t == MB_INTRAQ; }
inline bool mb_has_fourvec(int t) { return t == MB_INTER4V || t == MB_INTER4VQ; }
inline bool mb_has_quantizer(int t) { return t == MB_INTERQ || t == MB_INTRAQ || t == MB_INTER4VQ; }

struct MotionVector {
    int16_t x = 0, y = 0;
};

// HalfPel helpers (types.rs:721-798)
inline void into_lerp_parameters(int16_t v, int16_t* delta, bool* interp) {
    if (v % 2 == 0) {
        *delta = v / 2, *interp = false;
    } else if (v < 0) {
        *delta = (int16_t)(v / 2 - 1), *interp = true;
    } else {
        *delta = v / 2, *interp = true;
    }
}
inline int16_t hp_invert(int16_t v) { return v > 0 ? (int16_t)(v - 64) : (v < 0 ? (int16_t)(v + 64) : v); }
inline bool is_mv_within_range(int16_t v, int16_t range) { return -range <= v && v < range; }
inline int16_t average_sum_of_mvs(int16_t s) {
    int16_t whole = (int16_t)((s >> 4) << 1);
    int frac = s & 0x0F;
    if (frac <= 2) return whole;
    if (frac >= 14) return (int16_t)(whole + 2);
    return (int16_t)(whole + 1);
}
inline int16_t median_of(int16_t self, int16_t mhs, int16_t rhs) {
    if (self > mhs) {
        if (rhs > mhs) {
            return rhs > self ? self : rhs;
        }
        return mhs;
    } else if (mhs > rhs) {
        return rhs > self ? rhs : self;
    }
    return mhs;
}

struct TCoefficient {
    bool is_short;
    uint8_t run;
    int16_t level;
};
struct Block {
    int intradc = -1;  // raw 8-bit code, -1 = None
    std::vector<TCoefficient> tcoef;
};
// IntraDc::into_level (types.rs:955-961)
inline int16_t intradc_level(int code) { return code == 0xFF ? 1024 : (int16_t)((uint16_t)code << 3); }

enum DctClass { DCT_ZERO = 0, DCT_DC = 1, DCT_HORIZ = 2, DCT_VERT = 3, DCT_FULL = 4 };
// DecodedDctBlock (types.rs:902-916)
struct DecodedDctBlock {
    int cls = DCT_ZERO;
    float dc = 0.0f;
    float vec[8];
    float full[8][8];
};

/* ------------------------------------------------------------------------------------
 * DecodedPicture (decoder/picture.rs:8-143)
 * ---------------------------------------------------------------------------------- */
struct DecodedPicture {
    Picture header;
    SourceFormat format;
    int w = 0, h = 0;
    std::vector<uint8_t> luma, chroma_b, chroma_r;
    size_t chroma_samples_per_row = 0;

    // DecodedPicture::new (picture.rs:39-58)
    bool init(const Picture& hdr, const SourceFormat& fmt) {
        uint16_t fw, fh;
        if (!fmt.dims(&fw, &fh)) return false;
        header = hdr;
        format = fmt;
        w = fw, h = fh;
        luma.assign((size_t)w * h, 0);
        size_t cw = (size_t)std::ceil((float)w / 2.0f);
        size_t ch = (size_t)std::ceil((float)h / 2.0f);
        chroma_b.assign(cw * ch, 0);
        chroma_r.assign(cw * ch, 0);
        chroma_samples_per_row = cw;
        return true;
    }
};

/* ------------------------------------------------------------------------------------
 * Picture layer (parser/picture.rs)
 * ---------------------------------------------------------------------------------- */
// decode_sorenson_ptype (picture.rs:271-327)
Err decode_sorenson_ptype(Reader& r, SourceFormat* fmt, PicType* type, uint32_t* options) {
    uint64_t v;
    Err e = r.read_bits(3, &v);
    if (e) return e;
    bool have = true;
    uint32_t bit_count = 0;
    switch (v) {
        case 0: have = false, bit_count = 8; break;
        case 1: have = false, bit_count = 16; break;
        case 2: *fmt = {2, 0, 0}; break;
        case 3: *fmt = {1, 0, 0}; break;
        case 4: *fmt = {0, 0, 0}; break;
        case 5: *fmt = {6, 320, 240}; break;
        case 6: *fmt = {6, 160, 120}; break;
        default: *fmt = {5, 0, 0}; break;
    }
    if (!have) {
        uint64_t cw, chh;
        if ((e = r.read_bits(bit_count, &cw))) return e;
        if ((e = r.read_bits(bit_count, &chh))) return e;
        *fmt = {6, (uint16_t)cw, (uint16_t)chh};
    }
    if ((e = r.read_bits(2, &v))) return e;
    switch (v) {
        case 0: *type = PT_I; break;
        case 1: *type = PT_P; break;
        case 2: *type = PT_DISPOSABLE_P; break;
        default: *type = PT_RESERVED; break;
    }
    *options = 0;
    if ((e = r.read_bits(1, &v))) return e;
    if (v == 1) *options |= PO_USE_DEBLOCKER;
    return ORC_OK;
}

// decode_pei (picture.rs:577-595); the bytes are not observable downstream.
Err decode_pei(Reader& r) {
    for (;;) {
        uint64_t has_pei;
        Err e = r.read_bits(1, &has_pei);
        if (e) return e;
        if (has_pei == 1) {
            uint8_t b;
            if ((e = r.read_u8(&b))) return e;
        } else {
            return ORC_OK;
        }
    }
}

// decode_ptype (picture.rs:21-81). *plus = true when a PLUSPTYPE follows.
Err decode_ptype(Reader& r, uint32_t* options, bool* plus, SourceFormat* fmt, PicType* type) {
    *options = 0;
    *plus = false;
    uint8_t high;
    Err e = r.read_u8(&high);
    if (e) return e;
    if ((high & 0xC0) != 0x80) return ORC_ERR_INVALID_PTYPE;
    if (high & 0x20) *options |= PO_SPLIT_SCREEN;
    if (high & 0x10) *options |= PO_DOCUMENT_CAMERA;
    if (high & 0x08) *options |= PO_RELEASE_FREEZE;
    switch (high & 0x07) {
        case 0: return ORC_ERR_INVALID_PTYPE;
        case 1: *fmt = {0, 0, 0}; break;
        case 2: *fmt = {1, 0, 0}; break;
        case 3: *fmt = {2, 0, 0}; break;
        case 4: *fmt = {3, 0, 0}; break;
        case 5: *fmt = {4, 0, 0}; break;
        case 6: *fmt = {5, 0, 0}; break;
        default: *plus = true; return ORC_OK;
    }
    uint64_t low;
    if ((e = r.read_bits(5, &low))) return e;
    *type = (low & 0x10) ? PT_I : PT_P;  // sic: the reference maps the set bit to IFrame (picture.rs:57-61)
    if (low & 0x08) *options |= PO_UMV;
    if (low & 0x04) *options |= PO_SAC;
    if (low & 0x02) *options |= PO_AP;
    if (low & 0x01) *type = PT_PB;
    return ORC_OK;
}

// decode_picture (picture.rs:611-817). *none = true <=> Ok(None).
Err decode_picture(Reader& r, int decoder_options, const Picture* previous, Picture* out, bool* none) {
    size_t checkpoint = r.bits_read;  // with_transaction_union (reader.rs:404-420)
    *none = false;
    Err e;
    auto fail = [&](Err err) {
        r.bits_read = checkpoint;
        return err;
    };
    bool found;
    uint32_t skipped = 0;
    if ((e = r.recognize_start_code(false, &found, &skipped))) return fail(e);
    if (!found) return fail(ORC_ERR_MIDDLE_OF_BITSTREAM);
    if ((e = r.skip_bits(17 + skipped))) return fail(e);
    uint64_t gob_id;
    if ((e = r.read_bits(5, &gob_id))) return fail(e);

    Picture p;
    if (decoder_options & ORC_OPT_SORENSON_SPARK_BITSTREAM) {
        uint8_t tr;
        if ((e = r.read_u8(&tr))) return fail(e);
        SourceFormat fmt{5, 0, 0};
        PicType type;
        uint32_t options;
        {
            size_t cp2 = r.bits_read;
            if ((e = decode_sorenson_ptype(r, &fmt, &type, &options))) {
                r.bits_read = cp2;
                return fail(e);
            }
        }
        uint64_t q;
        if ((e = r.read_bits(5, &q))) return fail(e);
        if ((e = decode_pei(r))) return fail(e);
        p.version = (int)gob_id;  // "Sorenson abuses the GOB ID as a version field"
        p.temporal_reference = tr;
        p.has_format = true;
        p.format = fmt;
        p.options = options;
        p.picture_type = type;
        p.quantizer = (uint8_t)q;
        *out = p;
        return ORC_OK;
    } else if (gob_id != 0) {
        *none = true;
        r.bits_read = checkpoint;
        return ORC_OK;
    }

    uint8_t low_tr;
    if ((e = r.read_u8(&low_tr))) return fail(e);
    uint32_t options;
    bool plus;
    SourceFormat fmt{5, 0, 0};
    PicType type = PT_I;
    {
        size_t cp2 = r.bits_read;
        if ((e = decode_ptype(r, &options, &plus, &fmt, &type))) {
            r.bits_read = cp2;
            return fail(e);
        }
    }
    if (plus) {
        // decode_plusptype and its followers (picture.rs:136-267, 689-770): out of scope.
        return fail(ORC_ERR_UNIMPLEMENTED_DECODING);
    }
    // reference_picture_resampling (picture.rs:758-768): RPR is an MPPTYPE option (never set
    // without PLUSPTYPE); a format change against the previous picture reaches decode_rprp,
    // which is a stub returning UnimplementedDecoding (picture.rs:541-546).
    if (previous) {
        bool same = previous->has_format && previous->format == fmt;
        if (!same) return fail(ORC_ERR_UNIMPLEMENTED_DECODING);
    }
    uint64_t q;
    if ((e = r.read_bits(5, &q))) return fail(e);
    // decode_cpm_and_psbi (picture.rs:335-346)
    uint64_t cpm;
    if ((e = r.read_bits(1, &cpm))) return fail(e);
    if (cpm != 0) {
        uint64_t psbi;
        if ((e = r.read_bits(2, &psbi))) return fail(e);
    }
    if (type == PT_PB) {
        // decode_trb (3 bits without custom clock) + decode_dbquant (2 bits)
        uint64_t t;
        if ((e = r.read_bits(3, &t))) return fail(e);
        if ((e = r.read_bits(2, &t))) return fail(e);
    }
    if ((e = decode_pei(r))) return fail(e);
    p.version = -1;
    p.temporal_reference = low_tr;
    p.has_format = true;
    p.format = fmt;
    p.options = options;
    p.picture_type = type;
    p.quantizer = (uint8_t)q;
    *out = p;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * GOB layer (parser/gob.rs:50-71): a stub that only recognises picture start / EOS.
 * *none = true <=> Ok(None).
 * ---------------------------------------------------------------------------------- */
Err decode_gob(Reader& r, bool* none) {
    size_t checkpoint = r.bits_read;
    *none = false;
    bool found;
    uint32_t skipped = 0;
    Err e = r.recognize_start_code(false, &found, &skipped);
    if (!e && !found) e = ORC_ERR_INVALID_GOB_HEADER;
    if (!e) e = r.skip_bits(17 + skipped);
    uint64_t gob_id = 0;
    if (!e) e = r.read_bits(5, &gob_id);
    if (!e) {
        if (gob_id == 0 || gob_id == 15) {
            *none = true;
            r.bits_read = checkpoint;
            return ORC_OK;
        }
        e = ORC_ERR_UNIMPLEMENTED_DECODING;
    }
    r.bits_read = checkpoint;
    return e;
}

/* ------------------------------------------------------------------------------------
 * Macroblock layer (parser/macroblock.rs:414-549)
 * ---------------------------------------------------------------------------------- */
enum MbKind { MBK_UNCODED, MBK_STUFFING, MBK_CODED };
struct Macroblock {
    MbKind kind = MBK_UNCODED;
    int mb_type = MB_INTER;
    bool codes_luma[4] = {false, false, false, false};
    bool codes_chroma_b = false, codes_chroma_r = false;
    bool has_dquant = false;
    int8_t d_quantizer = 0;
    bool has_mv = false;
    MotionVector mv;
    bool has_addl = false;
    MotionVector addl[3];
};

// decode_motion_vector (macroblock.rs:414-438); UMV+PLUSPTYPE (read_umv) is out of scope.
Err decode_motion_vector(Reader& r, const Picture& pic, uint32_t running_options, MotionVector* mv) {
    size_t checkpoint = r.bits_read;
    if ((running_options & PO_UMV) && pic.has_plusptype) return ORC_ERR_UNIMPLEMENTED_DECODING;
    const VlcCode* c;
    Err e = r.read_vlc(g_trees[3], &c);
    if (!e && c->kind != 0) e = ORC_ERR_INVALID_MVD;
    int x = 0;
    if (!e) {
        x = c->a;
        e = r.read_vlc(g_trees[3], &c);
        if (!e && c->kind != 0) e = ORC_ERR_INVALID_MVD;
    }
    if (e) {
        r.bits_read = checkpoint;
        return e;
    }
    mv->x = (int16_t)x;
    mv->y = (int16_t)c->a;
    return ORC_OK;
}

// decode_macroblock (macroblock.rs:445-549)
// decode_disposable: EXTENSION beyond the reference (ORC_OPT_DECODE_DISPOSABLE): a Sorenson disposable P picture is
// parsed like a P picture; the reference itself fails it with UnimplementedDecoding (macroblock.rs:461-465).
Err decode_macroblock(Reader& r, const Picture& pic, uint32_t running_options, Macroblock* out, bool decode_disposable = false) {
    size_t checkpoint = r.bits_read;
    Err e = ORC_OK;
    Macroblock mb;
    auto fail = [&](Err err) {
        r.bits_read = checkpoint;
        return err;
    };
    uint64_t is_coded = 0;
    if (pic.picture_type != PT_I) {
        if ((e = r.read_bits(1, &is_coded))) return fail(e);
    }
    if (is_coded != 0) {
        mb.kind = MBK_UNCODED;
        *out = mb;
        return ORC_OK;
    }
    const VlcCode* c;
    if (pic.picture_type == PT_I) {
        if ((e = r.read_vlc(g_trees[0], &c))) return fail(e);
    } else if (pic.picture_type == PT_P || (decode_disposable && pic.picture_type == PT_DISPOSABLE_P)) {
        if ((e = r.read_vlc(g_trees[1], &c))) return fail(e);
    } else {
        return fail(ORC_ERR_UNIMPLEMENTED_DECODING);
    }
    if (c->kind == 1) {
        mb.kind = MBK_STUFFING;
        *out = mb;
        return ORC_OK;
    }
    if (c->kind == 2) return fail(ORC_ERR_INVALID_MACROBLOCK_HEADER);
    mb.kind = MBK_CODED;
    mb.mb_type = c->a;
    mb.codes_chroma_b = c->b != 0;
    mb.codes_chroma_r = c->c != 0;
    // MODB only exists for PbFrame pictures, which returned UnimplementedDecoding above.
    if ((e = r.read_vlc(g_trees[2], &c))) return fail(e);
    if (c->kind != 0) return fail(ORC_ERR_INVALID_MACROBLOCK_CODED_BITS);
    for (int i = 0; i < 4; i++) {
        bool v = ((c->a >> (3 - i)) & 1) != 0;
        mb.codes_luma[i] = mb_is_intra(mb.mb_type) ? v : !v;
    }
    if (running_options & PO_MODIFIED_QUANT) return fail(ORC_ERR_UNIMPLEMENTED_DECODING);
    if (mb_has_quantizer(mb.mb_type)) {
        // decode_dquant (macroblock.rs:257-270)
        uint64_t dq;
        if ((e = r.read_bits(2, &dq))) return fail(e);
        static const int8_t DQ[4] = {-1, -2, 1, 2};
        mb.has_dquant = true;
        mb.d_quantizer = DQ[dq];
    }
    if (mb_is_inter(mb.mb_type)) {  // is_any_pbframe() is impossible here
        if ((e = decode_motion_vector(r, pic, running_options, &mb.mv))) return fail(e);
        mb.has_mv = true;
    }
    if (mb_has_fourvec(mb.mb_type)) {
        for (int i = 0; i < 3; i++)
            if ((e = decode_motion_vector(r, pic, running_options, &mb.addl[i]))) return fail(e);
        mb.has_addl = true;
    }
    *out = mb;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * Block layer (parser/block.rs:670-755)
 * ---------------------------------------------------------------------------------- */
Err decode_block(Reader& r, int decoder_options, const Picture& pic, uint32_t running_options,
                 int mb_type, bool tcoef_present, Block* out) {
    size_t checkpoint = r.bits_read;
    Err e;
    Block blk;
    auto fail = [&](Err err) {
        r.bits_read = checkpoint;
        return err;
    };
    if (mb_is_intra(mb_type)) {
        uint8_t code;
        if ((e = r.read_u8(&code))) return fail(e);
        if (code == 0 || code == 128) return fail(ORC_ERR_INVALID_INTRA_DC);  // IntraDc::from_u8
        blk.intradc = code;
    }
    while (tcoef_present) {
        const VlcCode* c;
        if ((e = r.read_vlc(g_trees[4], &c))) return fail(e);
        if (c->kind == 2) return fail(ORC_ERR_INVALID_SHORT_COEFFICIENT);
        if (c->kind == 3) {
            uint32_t level_width = 8;
            if ((decoder_options & ORC_OPT_SORENSON_SPARK_BITSTREAM) && pic.version == 1) {
                uint64_t is11;
                if ((e = r.read_bits(1, &is11))) return fail(e);
                level_width = is11 == 1 ? 11 : 7;
            }
            uint64_t last, run;
            int64_t level;
            if ((e = r.read_bits(1, &last))) return fail(e);
            if ((e = r.read_bits(6, &run))) return fail(e);
            if ((e = r.read_signed_bits(level_width, &level))) return fail(e);
            if (level == 0) return fail(ORC_ERR_INVALID_LONG_COEFFICIENT);
            // `level == i16::MAX << level_width` (block.rs:716): the shift happens in i16, so
            // the constant is -256 / -128 / -2048, which no level_width-bit value can equal.
            int16_t forbidden = (int16_t)((uint16_t)0x7FFF << level_width);
            if ((int16_t)level == forbidden) {
                return fail((running_options & PO_MODIFIED_QUANT) ? ORC_ERR_UNIMPLEMENTED_DECODING
                                                                  : ORC_ERR_INVALID_LONG_COEFFICIENT);
            }
            blk.tcoef.push_back({false, (uint8_t)run, (int16_t)level});
            tcoef_present = last != 1;
        } else {
            uint64_t sign;
            if ((e = r.read_bits(1, &sign))) return fail(e);
            int16_t lv = (int16_t)c->c;
            blk.tcoef.push_back({true, (uint8_t)c->b, sign == 0 ? lv : (int16_t)-lv});
            tcoef_present = c->a == 0;
        }
    }
    *out = std::move(blk);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * Motion vector prediction (decoder/cpu/mvd_pred.rs)
 * ---------------------------------------------------------------------------------- */
typedef MotionVector Mv4[4];

// predict_candidate (mvd_pred.rs:27-67). `pv` = predictor_vectors[macroblocks_after_gob..].
MotionVector predict_candidate(const std::vector<std::array<MotionVector, 4>>& all, size_t after_gob,
                               const MotionVector cur[4], size_t mb_per_line, int index) {
    const std::array<MotionVector, 4>* pv = all.data() + after_gob;
    size_t pv_len = all.size() - after_gob;
    size_t current_mb = pv_len;
    size_t col_index = current_mb % mb_per_line;
    MotionVector zero;
    MotionVector mv1;
    if (index == 0 || index == 2) {
        mv1 = col_index == 0 ? zero : pv[current_mb - 1][index + 1];
    } else {
        mv1 = cur[index - 1];
    }
    size_t line_index = current_mb / mb_per_line;
    size_t last_line_mb = (line_index == 0 ? 0 : line_index - 1) * mb_per_line + col_index;
    MotionVector mv2;
    if (index == 0 || index == 1) {
        if (line_index == 0)
            mv2 = mv1;
        else
            mv2 = last_line_mb < pv_len ? pv[last_line_mb][index + 2] : mv1;
    } else {
        mv2 = cur[0];
    }
    bool is_end_of_line = col_index == (mb_per_line == 0 ? 0 : mb_per_line - 1);
    MotionVector mv3;
    if (index == 0 || index == 1) {
        if (is_end_of_line)
            mv3 = zero;
        else if (line_index == 0)
            mv3 = mv1;
        else
            mv3 = (last_line_mb + 1) < pv_len ? pv[last_line_mb + 1][2] : mv1;
    } else {
        mv3 = cur[1];
    }
    MotionVector r;
    r.x = median_of(mv1.x, mv2.x, mv3.x);
    r.y = median_of(mv1.y, mv2.y, mv3.y);
    return r;
}

// halfpel_decode (mvd_pred.rs:70-117).  For the supported pictures (no PLUSPTYPE) the
// running options never contain UNRESTRICTED_MOTION_VECTORS: it is an OPPTYPE option and
// state.rs:152-155 masks those out of non-PLUSPTYPE pictures (H263State.running_options is
// never updated and stays empty), so only the standard-range branch is reachable.
int16_t halfpel_decode(uint32_t running_options, const Picture& pic, int16_t predictor, int16_t mvd) {
    int16_t range = 32;
    int16_t out = (int16_t)(mvd + predictor);
    if ((running_options & PO_UMV) && !pic.has_plusptype) {
        if (is_mv_within_range(predictor, 32)) return out;
        range = 64;
    }
    if (!is_mv_within_range(out, range)) out = (int16_t)(hp_invert(mvd) + predictor);
    return out;
}
MotionVector mv_decode(uint32_t running_options, const Picture& pic, MotionVector pred, MotionVector mvd) {
    MotionVector r;
    r.x = halfpel_decode(running_options, pic, pred.x, mvd.x);
    r.y = halfpel_decode(running_options, pic, pred.y, mvd.y);
    return r;
}

/* ------------------------------------------------------------------------------------
 * inverse_rle (decoder/cpu/rle.rs:82-172)
 * ---------------------------------------------------------------------------------- */
// DEZIGZAG_MAPPING as (x, y) (rle.rs:6-71), generated from the zigzag scan rule and
// checked against the permutation property in tests/test_oracle_units.py.
struct XY {
    uint8_t x, y;
};
XY g_dezigzag[64];
struct DezigzagInit {
    DezigzagInit() {
        // Classic 8x8 zigzag: walk anti-diagonals, alternating direction.
        int idx = 0;
        for (int s = 0; s < 15; s++) {
            if (s % 2 == 0) {  // moving up-right: start at bottom-left of the diagonal
                for (int y = std::min(s, 7); y >= 0 && s - y <= 7; y--) g_dezigzag[idx++] = {(uint8_t)(s - y), (uint8_t)y};
            } else {
                for (int x = std::min(s, 7); x >= 0 && s - x <= 7; x--) g_dezigzag[idx++] = {(uint8_t)x, (uint8_t)(s - x)};
            }
        }
    }
} g_dezigzag_init;

void inverse_rle(const Block& enc, DecodedDctBlock* block, uint8_t quant) {
    if (enc.tcoef.empty()) {
        if (enc.intradc >= 0) {
            int16_t dc_level = intradc_level(enc.intradc);
            if (dc_level == 0) {
                block->cls = DCT_ZERO;
            } else {
                block->cls = DCT_DC;
                block->dc = (float)dc_level;
            }
        } else {
            block->cls = DCT_ZERO;
        }
        return;
    }
    float block_data[8][8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) block_data[y][x] = 0.0f;
    bool is_horiz = true, is_vert = true;
    size_t zigzag_index = 0;
    if (enc.intradc >= 0) {
        block_data[0][0] = (float)intradc_level(enc.intradc);
        zigzag_index += 1;
    }
    for (const TCoefficient& t : enc.tcoef) {
        zigzag_index += t.run;
        if (zigzag_index >= 64) return;  // the block keeps its previous value (Zero)
        XY p = g_dezigzag[zigzag_index];
        // i16 arithmetic, release-mode wrapping (rle.rs:130-133)
        int16_t absl = (int16_t)(t.level < 0 ? -t.level : t.level);
        int16_t dequantized = (int16_t)((int16_t)quant * (int16_t)((int16_t)(2 * absl) + 1));
        int16_t parity = (quant % 2 == 1) ? 0 : -1;
        int16_t sgn = t.level > 0 ? 1 : (t.level < 0 ? -1 : 0);
        int16_t value = (int16_t)(sgn * (int16_t)(dequantized + parity));
        value = std::min<int16_t>(std::max<int16_t>(value, -2048), 2047);
        float val = (float)value;
        block_data[p.y][p.x] = val;
        zigzag_index += 1;
        if (val != 0.0f) {
            if (p.y > 0) is_horiz = false;
            if (p.x > 0) is_vert = false;
        }
    }
    if (is_horiz && is_vert) {
        if (block_data[0][0] == 0.0f) {
            block->cls = DCT_ZERO;
        } else {
            block->cls = DCT_DC;
            block->dc = block_data[0][0];
        }
    } else if (is_horiz) {
        block->cls = DCT_HORIZ;
        for (int i = 0; i < 8; i++) block->vec[i] = block_data[0][i];
    } else if (is_vert) {
        block->cls = DCT_VERT;
        for (int i = 0; i < 8; i++) block->vec[i] = block_data[i][0];
    } else {
        block->cls = DCT_FULL;
        std::memcpy(block->full, block_data, sizeof(block_data));
    }
}

/* ------------------------------------------------------------------------------------
 * IDCT (decoder/cpu/idct.rs)
 * ---------------------------------------------------------------------------------- */
// BASIS_TABLE (idct.rs:39-48): f32 literals of cos(pi*(i+0.5)/8*freq), row 0 times 1/sqrt(2).
const float BASIS_TABLE[8][8] = {
    {0.70710677f, 0.70710677f, 0.70710677f, 0.70710677f, 0.70710677f, 0.70710677f, 0.70710677f, 0.70710677f},
    {0.98078525f, 0.8314696f, 0.5555702f, 0.19509023f, -0.19509032f, -0.55557036f, -0.83146966f, -0.9807853f},
    {0.9238795f, 0.38268343f, -0.38268352f, -0.9238796f, -0.9238795f, -0.38268313f, 0.3826836f, 0.92387956f},
    {0.8314696f, -0.19509032f, -0.9807853f, -0.55557f, 0.55557007f, 0.98078525f, 0.19509007f, -0.8314698f},
    {0.70710677f, -0.70710677f, -0.70710665f, 0.707107f, 0.70710677f, -0.70710725f, -0.70710653f, 0.7071068f},
    {0.5555702f, -0.9807853f, 0.19509041f, 0.83146936f, -0.8314698f, -0.19508928f, 0.9807853f, -0.55557007f},
    {0.38268343f, -0.9238795f, 0.92387974f, -0.3826839f, -0.38268384f, 0.9238793f, -0.92387974f, 0.3826839f},
    {0.19509023f, -0.55557f, 0.83146936f, -0.9807852f, 0.98078525f, -0.83147013f, 0.55557114f, -0.19508967f},
};

// idct_1d (idct.rs:52-65): out[i] = ((0 + in[0]*B[0][i]) + in[1]*B[1][i]) + ...
void idct_1d(const float in[8], float out[8]) {
    for (int i = 0; i < 8; i++) {
        float acc = 0.0f;
        for (int freq = 0; freq < 8; freq++) acc += in[freq] * BASIS_TABLE[freq][i];
        out[i] = acc;
    }
}
inline float signum(float v) { return std::signbit(v) ? -1.0f : 1.0f; }  // f32::signum (NaN unreachable)
// `as i16` saturating float->int cast
inline int16_t f32_as_i16(float v) {
    if (v != v) return 0;
    if (v <= -32768.0f) return -32768;
    if (v >= 32767.0f) return 32767;
    return (int16_t)v;  // truncates toward zero
}
inline int16_t clamp16(int16_t v, int16_t lo, int16_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// idct_channel (idct.rs:82-201)
void idct_channel(const std::vector<DecodedDctBlock>& levels, std::vector<uint8_t>& output,
                  size_t blk_per_line, size_t samples_per_line) {
    size_t output_height = output.size() / samples_per_line;
    size_t blk_height = levels.size() / blk_per_line;
    float inter[8][8], outp[8][8];
    for (size_t y_base = 0; y_base < blk_height; y_base++) {
        for (size_t x_base = 0; x_base < blk_per_line; x_base++) {
            size_t block_id = x_base + y_base * blk_per_line;
            if (block_id >= levels.size()) continue;
            long xs = std::min<long>(std::max<long>((long)samples_per_line - (long)x_base * 8, 0), 8);
            long ys = std::min<long>(std::max<long>((long)output_height - (long)y_base * 8, 0), 8);
            const DecodedDctBlock& b = levels[block_id];
            switch (b.cls) {
                case DCT_ZERO: break;
                case DCT_DC: {
                    float dc = b.dc;
                    int16_t clipped = clamp16(f32_as_i16(dc * 0.5f / 4.0f + signum(dc) * 0.5f), -256, 255);
                    for (long yo = 0; yo < ys; yo++)
                        for (long xo = 0; xo < xs; xo++) {
                            size_t idx = x_base * 8 + xo + (y_base * 8 + yo) * samples_per_line;
                            int16_t mocomp = output[idx];
                            output[idx] = (uint8_t)clamp16((int16_t)(clipped + mocomp), 0, 255);
                        }
                    break;
                }
                case DCT_HORIZ: {
                    idct_1d(b.vec, inter[0]);
                    for (long yo = 0; yo < ys; yo++)
                        for (long xo = 0; xo < xs; xo++) {
                            float idct = inter[0][xo];
                            size_t idx = x_base * 8 + xo + (y_base * 8 + yo) * samples_per_line;
                            int16_t clipped = clamp16(
                                f32_as_i16(idct * BASIS_TABLE[0][0] / 4.0f + signum(idct) * 0.5f), -256, 255);
                            int16_t mocomp = output[idx];
                            output[idx] = (uint8_t)clamp16((int16_t)(clipped + mocomp), 0, 255);
                        }
                    break;
                }
                case DCT_VERT: {
                    idct_1d(b.vec, inter[0]);
                    for (long yo = 0; yo < ys; yo++) {
                        float idct = inter[0][yo];
                        for (long xo = 0; xo < xs; xo++) {
                            size_t idx = x_base * 8 + xo + (y_base * 8 + yo) * samples_per_line;
                            int16_t clipped = clamp16(
                                f32_as_i16(idct * BASIS_TABLE[0][0] / 4.0f + signum(idct) * 0.5f), -256, 255);
                            int16_t mocomp = output[idx];
                            output[idx] = (uint8_t)clamp16((int16_t)(clipped + mocomp), 0, 255);
                        }
                    }
                    break;
                }
                case DCT_FULL: {
                    for (int row = 0; row < 8; row++) {
                        idct_1d(b.full[row], outp[row]);
                        for (int i = 0; i < 8; i++) inter[i][row] = outp[row][i];  // transposition
                    }
                    for (int row = 0; row < 8; row++) idct_1d(inter[row], outp[row]);
                    for (long xo = 0; xo < xs; xo++)
                        for (long yo = 0; yo < ys; yo++) {
                            float idct = outp[xo][yo];
                            size_t idx = x_base * 8 + xo + (y_base * 8 + yo) * samples_per_line;
                            int16_t clipped = clamp16(f32_as_i16(idct / 4.0f + signum(idct) * 0.5f), -256, 255);
                            int16_t mocomp = output[idx];
                            output[idx] = (uint8_t)clamp16((int16_t)(clipped + mocomp), 0, 255);
                        }
                    break;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Motion compensation (decoder/cpu/gather.rs)
 * ---------------------------------------------------------------------------------- */
// read_sample (gather.rs:16-31)
inline uint8_t read_sample(const uint8_t* px, size_t samples_per_row, size_t num_rows, long x, long y) {
    long mx = (long)(samples_per_row == 0 ? 0 : samples_per_row - 1);
    long my = (long)(num_rows == 0 ? 0 : num_rows - 1);
    x = std::min(std::max(x, 0L), mx);
    y = std::min(std::max(y, 0L), my);
    return px[(size_t)x + (size_t)y * samples_per_row];
}
// lerp (gather.rs:34-40)
inline uint8_t lerp(uint8_t a, uint8_t b, bool middle) {
    return middle ? (uint8_t)(((uint16_t)a + (uint16_t)b + 1) / 2) : a;
}
// gather_block (gather.rs:47-126)
void gather_block(const uint8_t* px, size_t px_len, size_t samples_per_row, size_t pos_x, size_t pos_y,
                  MotionVector mv, uint8_t* target) {
    int16_t x_delta, y_delta;
    bool x_interp, y_interp;
    into_lerp_parameters(mv.x, &x_delta, &x_interp);
    into_lerp_parameters(mv.y, &y_delta, &y_interp);
    long src_x = (long)pos_x + x_delta;
    long src_y = (long)pos_y + y_delta;
    size_t array_height = px_len / samples_per_row;
    long block_cols = std::min(std::max((long)samples_per_row - (long)pos_x, 0L), 8L);
    long block_rows = std::min(std::max((long)array_height - (long)pos_y, 0L), 8L);
    if (!x_interp && !y_interp) {
        if (block_cols == 8 && block_rows == 8 && src_x >= 0 && src_x <= (long)samples_per_row - 8 &&
            src_y >= 0 && src_y <= (long)array_height - 8) {
            for (long j = 0; j < 8; j++)
                std::memcpy(target + pos_x + (pos_y + j) * samples_per_row,
                            px + (size_t)src_x + (size_t)(src_y + j) * samples_per_row, 8);
        } else {
            for (long j = 0; j < block_rows; j++)
                for (long i = 0; i < block_cols; i++)
                    target[pos_x + i + (pos_y + j) * samples_per_row] =
                        read_sample(px, samples_per_row, array_height, src_x + i, src_y + j);
        }
    } else {
        for (long j = 0; j < block_rows; j++) {
            long v = src_y + j;
            for (long i = 0; i < block_cols; i++) {
                long u = src_x + i;
                uint8_t s00 = read_sample(px, samples_per_row, array_height, u, v);
                uint8_t s10 = read_sample(px, samples_per_row, array_height, u + 1, v);
                uint8_t s01 = read_sample(px, samples_per_row, array_height, u, v + 1);
                uint8_t s11 = read_sample(px, samples_per_row, array_height, u + 1, v + 1);
                uint8_t r;
                if (x_interp && y_interp) {
                    r = (uint8_t)(((uint16_t)s00 + s10 + s01 + s11 + 2) / 4);
                } else {
                    uint8_t m0 = lerp(s00, s10, x_interp);
                    uint8_t m1 = lerp(s01, s11, x_interp);
                    r = lerp(m0, m1, y_interp);
                }
                target[pos_x + i + (pos_y + j) * samples_per_row] = r;
            }
        }
    }
}

// gather (gather.rs:140-204)
Err gather(const std::vector<int8_t>& mb_types, const DecodedPicture* reference,
           const std::vector<std::array<MotionVector, 4>>& mvs, size_t mb_per_line, DecodedPicture& np) {
    size_t n = std::min(mb_types.size(), mvs.size());
    for (size_t i = 0; i < n; i++) {
        if (!mb_is_inter(mb_types[i])) continue;
        if (!reference) return ORC_ERR_UNCODED_IFRAME_BLOCKS;
        // The reference indexes the new picture with the reference picture's stride; with
        // differing dimensions that is an out-of-bounds write or garbage.
        if (reference->w != np.w || reference->h != np.h) return ORC_ERR_REFERENCE_WOULD_ABORT;
        size_t lspr = (size_t)reference->w;
        size_t px = (i % mb_per_line) * 16, py = (i / mb_per_line) * 16;
        const std::array<MotionVector, 4>& mv = mvs[i];
        gather_block(reference->luma.data(), reference->luma.size(), lspr, px, py, mv[0], np.luma.data());
        gather_block(reference->luma.data(), reference->luma.size(), lspr, px + 8, py, mv[1], np.luma.data());
        gather_block(reference->luma.data(), reference->luma.size(), lspr, px, py + 8, mv[2], np.luma.data());
        gather_block(reference->luma.data(), reference->luma.size(), lspr, px + 8, py + 8, mv[3], np.luma.data());
        MotionVector c;
        c.x = average_sum_of_mvs((int16_t)(mv[0].x + mv[1].x + mv[2].x + mv[3].x));
        c.y = average_sum_of_mvs((int16_t)(mv[0].y + mv[1].y + mv[2].y + mv[3].y));
        size_t cspr = reference->chroma_samples_per_row;
        size_t cx = (i % mb_per_line) * 8, cy = (i / mb_per_line) * 8;
        gather_block(reference->chroma_b.data(), reference->chroma_b.size(), cspr, cx, cy, c, np.chroma_b.data());
        gather_block(reference->chroma_r.data(), reference->chroma_r.size(), cspr, cx, cy, c, np.chroma_r.data());
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * yuv::bt601 (yuv/src/bt601.rs)
 * ---------------------------------------------------------------------------------- */
// yuv_to_rgba_4x (bt601.rs:12-59), i32 lanes (wide::i32x4: wrapping ops, arithmetic >>)
inline void yuv_to_rgba_4x(const uint8_t y4[4], const uint8_t cb2[2], const uint8_t cr2[2], uint8_t rgba[16]) {
    for (int k = 0; k < 4; k++) {
        int32_t y = (int32_t)y4[k] - 16;
        int32_t cb = (int32_t)cb2[k / 2] - 128;
        int32_t cr = (int32_t)cr2[k / 2] - 128;
        int32_t gray = y * 76309;
        int32_t cr2r = cr * 104597;
        int32_t cr2g = cr * -53279;
        int32_t cb2g = cb * -25675;
        int32_t cb2b = cb * 132201;
        int32_t half = 32768;
        int32_t r = (gray + cr2r + half) >> 16;
        int32_t g = (gray + cr2g + cb2g + half) >> 16;
        int32_t b = (gray + cb2b + half) >> 16;
        r = std::min(std::max(r, 0), 255);
        g = std::min(std::max(g, 0), 255);
        b = std::min(std::max(b, 0), 255);
        rgba[4 * k + 0] = (uint8_t)r;
        rgba[4 * k + 1] = (uint8_t)g;
        rgba[4 * k + 2] = (uint8_t)b;
        rgba[4 * k + 3] = 255;
    }
}

// yuv420_to_rgba (bt601.rs:105-196)
void yuv420_to_rgba(const uint8_t* y, const uint8_t* chroma_b, const uint8_t* chroma_r, size_t y_len,
                    size_t y_width, uint8_t* rgba) {
    if (y_len == 0) return;
    size_t br_width = (y_width + 1) / 2;
    size_t y_height = y_len / y_width;
    size_t rgba_stride = y_width * 4;
    for (size_t luma_row = 0; luma_row < y_height; luma_row++) {
        size_t chroma_row = luma_row / 2;
        size_t y_rem = y_width % 4;
        size_t rgba_rem = y_rem * 4;
        const uint8_t* y_row = y + luma_row * y_width;
        const uint8_t* cb_row = chroma_b + chroma_row * br_width;
        const uint8_t* cr_row = chroma_r + chroma_row * br_width;
        uint8_t* rgba_row = rgba + luma_row * rgba_stride;
        size_t chunks = (y_width - y_rem) / 4;  // zip() stops at the shortest iterator = luma chunks
        for (size_t c = 0; c < chunks; c++) yuv_to_rgba_4x(y_row + 4 * c, cb_row + 2 * c, cr_row + 2 * c, rgba_row + 16 * c);
        if (y_rem != 0) {
            uint8_t yy[4] = {0, 0, 0, 0}, cb[2] = {0, 0}, cr[2] = {0, 0};
            for (size_t x = y_width - y_rem; x < y_width; x++) {
                yy[x % 4] = y_row[x];
                cb[(x % 4) / 2] = cb_row[x / 2];
                cr[(x % 4) / 2] = cr_row[x / 2];
            }
            uint8_t tmp[16];
            yuv_to_rgba_4x(yy, cb, cr, tmp);
            for (size_t i = rgba_stride - rgba_rem; i < rgba_stride; i++) rgba_row[i] = tmp[i % 16];
        }
    }
}

/* ------------------------------------------------------------------------------------
 * deblock (deblock/src/deblock.rs)
 * ---------------------------------------------------------------------------------- */
const uint8_t QUANT_TO_STRENGTH[32] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 7,
                                       7, 8, 8, 8, 9, 9, 9, 10, 10, 10, 11, 11, 11, 12, 12, 12};

inline int16_t i16signum(int16_t x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }
inline int16_t i16abs(int16_t x) { return (int16_t)(x < 0 ? -x : x); }
// up_down_ramp (deblock.rs:13-15 / 66-69)
inline int16_t up_down_ramp(int16_t x, int16_t strength) {
    int16_t ax = i16abs(x);
    int16_t inner = std::max<int16_t>((int16_t)(2 * (int16_t)(ax - strength)), 0);
    return (int16_t)(i16signum(x) * std::max<int16_t>((int16_t)(ax - inner), 0));
}
// scalar process (deblock.rs:29-42): truncating `/`
inline void process_scalar(uint8_t* A, uint8_t* B, uint8_t* C, uint8_t* D, int strength) {
    int16_t a = *A, b = *B, c = *C, d = *D;
    int16_t dd = (int16_t)((a - 4 * b + 4 * c - d) / 8);
    int16_t d1 = up_down_ramp(dd, (int16_t)strength);
    int16_t lim = i16abs((int16_t)(d1 / 2));
    int16_t d2 = clamp16((int16_t)((a - d) / 4), (int16_t)-lim, lim);
    *A = (uint8_t)(int16_t)(a - d2);
    *B = (uint8_t)clamp16((int16_t)(b + d1), 0, 255);
    *C = (uint8_t)clamp16((int16_t)(c - d1), 0, 255);
    *D = (uint8_t)(int16_t)(d + d2);
}
// one lane of process_simd (deblock.rs:99-127): arithmetic shift right (floor)
inline void process_simd_lane(uint8_t* A, uint8_t* B, uint8_t* C, uint8_t* D, int strength) {
    int16_t a = *A, b = *B, c = *C, d = *D;
    int16_t dd = (int16_t)((int16_t)(a - 4 * b + 4 * c - d) >> 3);
    int16_t d1 = up_down_ramp(dd, (int16_t)strength);
    int16_t lim = i16abs((int16_t)(d1 >> 1));
    int16_t d2 = clamp16((int16_t)((int16_t)(a - d) >> 2), (int16_t)-lim, lim);
    *A = (uint8_t)(int16_t)(a - d2);
    *B = (uint8_t)clamp16((int16_t)(b + d1), 0, 255);
    *C = (uint8_t)clamp16((int16_t)(c - d1), 0, 255);
    *D = (uint8_t)(int16_t)(d + d2);
}
// deblock_horiz (deblock.rs:136-181)
void deblock_horiz(uint8_t* res, size_t len, size_t width, int strength) {
    size_t height = len / width;
    if (height < 2) return;  // the reference underflows `height - 2` and aborts here
    size_t simd_cols = (width / 8) * 8;
    for (size_t edge_y = 8; edge_y <= height - 2; edge_y += 8) {
        uint8_t* ra = res + (edge_y - 2) * width;
        uint8_t* rb = ra + width;
        uint8_t* rc = rb + width;
        uint8_t* rd = rc + width;
        for (size_t x = 0; x < simd_cols; x++) process_simd_lane(ra + x, rb + x, rc + x, rd + x, strength);
        for (size_t x = simd_cols; x < width; x++) process_scalar(ra + x, rb + x, rc + x, rd + x, strength);
    }
}
// deblock_vert (deblock.rs:185-299)
void deblock_vert(uint8_t* res, size_t len, size_t width, int strength) {
    if (width < 10) return;
    size_t height = len / width;
    size_t simd_rows = (height / 8) * 8;
    size_t chunks = (width - 2) / 8;  // row[2..].chunks_exact(8)
    for (size_t y = 0; y < height; y++) {
        uint8_t* row = res + y * width;
        for (size_t k = 0; k < chunks; k++) {
            uint8_t* chunk = row + 2 + 8 * k;
            if (y < simd_rows)
                process_simd_lane(chunk + 4, chunk + 5, chunk + 6, chunk + 7, strength);
            else
                process_scalar(chunk + 4, chunk + 5, chunk + 6, chunk + 7, strength);
        }
    }
}
// deblock (deblock.rs:305-315)
void deblock(const uint8_t* data, size_t len, size_t width, int strength, uint8_t* result) {
    std::memcpy(result, data, len);
    if (len == 0 || width == 0) return;
    deblock_horiz(result, len, width, strength);
    deblock_vert(result, len, width, strength);
}

}  // namespace

/* ------------------------------------------------------------------------------------
 * H263State (decoder/state.rs)
 * ---------------------------------------------------------------------------------- */
struct orc_state {
    int decoder_options = 0;
    // last_picture / reference_picture (state.rs:23-31).  get_reference_picture()
    // (state.rs:72-78) returns the entry keyed by *last_picture* whenever
    // reference_picture is Some, so the map reduces to "the last decoded picture" plus a
    // flag saying whether a non-disposable picture has been decoded since the last reset.
    bool has_last = false;
    bool has_reference = false;
    DecodedPicture last;
    // EXTENSION (ORC_OPT_DECODE_DISPOSABLE): the last NON-disposable picture, which is what a disposable-aware decoder
    // predicts from.  Unused (and never filled) without the option: the reference predicts from the last picture.
    DecodedPicture reference;
    uint32_t running_options = 0;  // never updated by the reference (stays empty)

    bool trace = false;
    std::vector<int8_t> t_mb_type, t_coded;
    std::vector<uint8_t> t_quant, t_nev, t_run;
    std::vector<int16_t> t_mv, t_intradc, t_level;
};

namespace {

// decode_next_picture (state.rs:138-489)
Err decode_next_picture(orc_state* st, Reader& reader) {
    ensure_trees();
    size_t checkpoint = reader.bits_read;  // with_transaction (reader.rs:376-389)
    auto fail = [&](Err e) {
        reader.bits_read = checkpoint;
        return e;
    };
    Picture next;
    bool none;
    Err e = decode_picture(reader, st->decoder_options, st->has_last ? &st->last.header : nullptr, &next, &none);
    if (e) return fail(e);
    if (none) return fail(ORC_ERR_MIDDLE_OF_BITSTREAM);

    uint32_t next_running_options;
    if (next.has_plusptype && next.has_opptype)
        next_running_options = next.options;
    else if (next.has_plusptype)
        next_running_options = (next.options & ~OPPTYPE_OPTIONS) | (st->running_options & OPPTYPE_OPTIONS);
    else
        next_running_options = (next.options & ~OPPTYPE_OPTIONS & ~MPPTYPE_OPTIONS) |
                               (st->running_options & (OPPTYPE_OPTIONS | MPPTYPE_OPTIONS));

    SourceFormat format;
    if (next.has_format)
        format = next.format;
    else if (next.picture_type == PT_I)
        return fail(ORC_ERR_PICTURE_FORMAT_MISSING);
    else if (st->has_last)
        format = st->last.format;
    else
        return fail(ORC_ERR_PICTURE_FORMAT_MISSING);

    const bool ext_disposable = (st->decoder_options & ORC_OPT_DECODE_DISPOSABLE) != 0;
    const DecodedPicture* reference_picture = (st->has_reference && st->has_last) ? (ext_disposable ? &st->reference : &st->last) : nullptr;

    uint16_t ow, oh;
    if (!format.dims(&ow, &oh)) return fail(ORC_ERR_PICTURE_FORMAT_INVALID);
    size_t mb_per_line = (size_t)std::ceil((double)ow / 16.0);
    size_t mb_height = (size_t)std::ceil((double)oh / 16.0);
    // mb_per_line == 0 makes `len % mb_per_line` (state.rs:200) divide by zero.
    if (mb_per_line == 0) return fail(ORC_ERR_REFERENCE_WOULD_ABORT);
    size_t level_w = mb_per_line * 16, level_h = mb_height * 16;
    size_t capacity = mb_per_line * mb_height;

    uint8_t in_force_quantizer = next.quantizer;
    std::vector<std::array<MotionVector, 4>> predictor_vectors;
    std::vector<int8_t> macroblock_types;
    predictor_vectors.reserve(capacity);
    macroblock_types.reserve(capacity);
    size_t macroblocks_after_gob = 0;

    DecodedPicture np;
    if (!np.init(next, format)) return fail(ORC_ERR_PICTURE_FORMAT_INVALID);

    std::vector<DecodedDctBlock> luma_levels(level_w * level_h / 64);
    std::vector<DecodedDctBlock> chroma_b_levels(level_w * level_h / 4 / 64);
    std::vector<DecodedDctBlock> chroma_r_levels(level_w * level_h / 4 / 64);

    const bool is_sorenson = (st->decoder_options & ORC_OPT_SORENSON_SPARK_BITSTREAM) != 0;
    const bool tr = st->trace;
    std::vector<int8_t> t_mb_type, t_coded;
    std::vector<uint8_t> t_quant, t_nev, t_run;
    std::vector<int16_t> t_mv, t_intradc, t_level;

    for (;;) {
        Macroblock mb;
        Err me = decode_macroblock(reader, np.header, next_running_options, &mb, ext_disposable);
        size_t pos_x = (macroblock_types.size() % mb_per_line) * 16;
        size_t pos_y = (macroblock_types.size() / mb_per_line) * 16;
        std::array<MotionVector, 4> motion_vectors{};
        int8_t mb_type;
        int16_t tr_dc[6] = {-1, -1, -1, -1, -1, -1};
        uint8_t tr_nev[6] = {0, 0, 0, 0, 0, 0};
        bool coded = false;

        if (!me && mb.kind == MBK_STUFFING) {
            continue;
        } else if (!me && mb.kind == MBK_UNCODED) {
            if (np.header.picture_type == PT_I) return fail(ORC_ERR_UNCODED_IFRAME_BLOCKS);
            mb_type = MB_INTER;
        } else if (!me) {
            coded = true;
            int8_t quantizer = (int8_t)((int8_t)in_force_quantizer + (mb.has_dquant ? mb.d_quantizer : 0));
            in_force_quantizer = (uint8_t)std::min<int8_t>(std::max<int8_t>(quantizer, 1), 31);
            if (mb_is_inter(mb.mb_type)) {
                MotionVector mv1 = mb.has_mv ? mb.mv : MotionVector();
                MotionVector pred = predict_candidate(predictor_vectors, macroblocks_after_gob,
                                                      motion_vectors.data(), mb_per_line, 0);
                motion_vectors[0] = mv_decode(next_running_options, np.header, pred, mv1);
                if (mb.has_addl) {
                    for (int k = 1; k < 4; k++) {
                        pred = predict_candidate(predictor_vectors, macroblocks_after_gob,
                                                 motion_vectors.data(), mb_per_line, k);
                        motion_vectors[k] = mv_decode(next_running_options, np.header, pred, mb.addl[k - 1]);
                    }
                } else {
                    motion_vectors[1] = motion_vectors[2] = motion_vectors[3] = motion_vectors[0];
                }
            }
            // Six blocks: Y0 Y1 Y2 Y3 Cb Cr (state.rs:287-381)
            for (int bi = 0; bi < 6; bi++) {
                bool present = bi < 4 ? mb.codes_luma[bi] : (bi == 4 ? mb.codes_chroma_b : mb.codes_chroma_r);
                Block blk;
                Err be = decode_block(reader, st->decoder_options, np.header, next_running_options,
                                      mb.mb_type, present, &blk);
                if (be) return fail(be);  // `?`: block errors (incl. EOF) abort the picture
                std::vector<DecodedDctBlock>* levels;
                size_t bx, by, bpl;
                if (bi < 4) {
                    levels = &luma_levels;
                    bx = pos_x + (bi & 1) * 8, by = pos_y + (bi >> 1) * 8, bpl = level_w / 8;
                } else {
                    levels = bi == 4 ? &chroma_b_levels : &chroma_r_levels;
                    bx = pos_x / 2, by = pos_y / 2, bpl = mb_per_line;
                }
                size_t block_id = bx / 8 + (by / 8) * bpl;
                // rle.rs:89-90: out-of-range index would panic (abort).
                if (block_id >= levels->size()) return fail(ORC_ERR_REFERENCE_WOULD_ABORT);
                inverse_rle(blk, &(*levels)[block_id], in_force_quantizer);
                if (tr) {
                    tr_dc[bi] = (int16_t)blk.intradc;
                    tr_nev[bi] = (uint8_t)blk.tcoef.size();
                    for (const TCoefficient& t : blk.tcoef) {
                        t_run.push_back(t.run);
                        t_level.push_back(t.level);
                    }
                }
            }
            mb_type = (int8_t)mb.mb_type;
        } else if ((me == ORC_ERR_INVALID_MACROBLOCK_HEADER || me == ORC_ERR_INVALID_MACROBLOCK_CODED_BITS) &&
                   !is_sorenson) {
            // Attempt to recover from macroblock errors (state.rs:387-408)
            bool gnone;
            Err ge = decode_gob(reader, &gnone);
            if (!ge && gnone) break;
            // Ok(Some(gob)) is unreachable: decode_gob never returns Some.
            if (ge == ORC_ERR_UNHANDLED_IO_ERROR || ge == ORC_ERR_INVALID_GOB_HEADER) break;
            return fail(ge);
        } else if (me == ORC_ERR_UNHANDLED_IO_ERROR) {
            break;  // EOF ends the picture (state.rs:411)
        } else {
            return fail(me);
        }

        predictor_vectors.push_back(motion_vectors);
        macroblock_types.push_back(mb_type);
        if (tr) {
            t_mb_type.push_back(mb_type);
            t_coded.push_back(coded ? 1 : 0);
            t_quant.push_back(in_force_quantizer);
            for (int k = 0; k < 4; k++) {
                t_mv.push_back(motion_vectors[k].x);
                t_mv.push_back(motion_vectors[k].y);
            }
            for (int k = 0; k < 6; k++) {
                t_intradc.push_back(tr_dc[k]);
                t_nev.push_back(tr_nev[k]);
            }
        }
    }

    // Pad a picture that ended early (state.rs:419-427)
    if (predictor_vectors.size() < capacity) predictor_vectors.resize(capacity);
    if (macroblock_types.size() < capacity) {
        if (tr) {
            for (size_t i = macroblock_types.size(); i < capacity; i++) {
                t_mb_type.push_back(MB_INTER);
                t_coded.push_back(0);
                t_quant.push_back(in_force_quantizer);
                for (int k = 0; k < 8; k++) t_mv.push_back(0);
                for (int k = 0; k < 6; k++) {
                    t_intradc.push_back(-1);
                    t_nev.push_back(0);
                }
            }
        }
        macroblock_types.resize(capacity, MB_INTER);
    }

    e = gather(macroblock_types, reference_picture, predictor_vectors, mb_per_line, np);
    if (e) return fail(e);
    idct_channel(luma_levels, np.luma, mb_per_line * 2, (size_t)ow);
    size_t cspr = np.chroma_samples_per_row;
    idct_channel(chroma_b_levels, np.chroma_b, mb_per_line, cspr);
    idct_channel(chroma_r_levels, np.chroma_r, mb_per_line, cspr);

    // Reference bookkeeping (state.rs:464-483)
    if (np.header.picture_type == PT_I) st->has_reference = false;
    st->has_last = true;
    if (np.header.picture_type != PT_DISPOSABLE_P) {
        st->has_reference = true;
        if (ext_disposable) st->reference = np;  // a copy: `last` and `reference` part ways at the next disposable picture
    }
    st->last = std::move(np);
    if (tr) {
        st->t_mb_type.swap(t_mb_type);
        st->t_coded.swap(t_coded);
        st->t_quant.swap(t_quant);
        st->t_nev.swap(t_nev);
        st->t_run.swap(t_run);
        st->t_mv.swap(t_mv);
        st->t_intradc.swap(t_intradc);
        st->t_level.swap(t_level);
    }
    return ORC_OK;
}

inline uint64_t weighted_sum(const uint8_t* p, size_t n) {
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) s += (uint64_t)(p[i] + 1u) * (uint64_t)(((uint32_t)i * 2654435761u) | 1u);
    return s;
}

}  // namespace

/* ------------------------------------------------------------------------------------
 * C ABI
 * ---------------------------------------------------------------------------------- */
extern "C" {

orc_state* orc_state_new(int decoder_options) {
    ensure_trees();
    orc_state* s = new orc_state();
    s->decoder_options = decoder_options;
    return s;
}
void orc_state_free(orc_state* s) { delete s; }

int orc_decode_next_picture(orc_state* s, const uint8_t* data, size_t len) {
    Reader r{data, len, 0};
    return decode_next_picture(s, r);
}

int orc_last_picture_info(orc_state* s, int* width, int* height, int* temporal_reference, int* picture_type,
                          int* quantizer, int* deblock_flag, int* version) {
    if (!s->has_last) return -1;
    const DecodedPicture& p = s->last;
    if (width) *width = p.w;
    if (height) *height = p.h;
    if (temporal_reference) *temporal_reference = p.header.temporal_reference;
    if (picture_type) {
        switch (p.header.picture_type) {
            case PT_I: *picture_type = ORC_PIC_I; break;
            case PT_P: *picture_type = ORC_PIC_P; break;
            case PT_DISPOSABLE_P: *picture_type = ORC_PIC_DISPOSABLE_P; break;
            default: *picture_type = ORC_PIC_OTHER; break;
        }
    }
    if (quantizer) *quantizer = p.header.quantizer;
    if (deblock_flag) *deblock_flag = (p.header.options & PO_USE_DEBLOCKER) ? 1 : 0;
    if (version) *version = p.header.version;
    return 0;
}

int orc_last_picture_yuv(orc_state* s, uint8_t* y, uint8_t* cb, uint8_t* cr) {
    if (!s->has_last) return -1;
    const DecodedPicture& p = s->last;
    std::memcpy(y, p.luma.data(), p.luma.size());
    std::memcpy(cb, p.chroma_b.data(), p.chroma_b.size());
    std::memcpy(cr, p.chroma_r.data(), p.chroma_r.size());
    return 0;
}

void orc_state_set_trace(orc_state* s, int enable) { s->trace = enable != 0; }
int orc_trace_counts(orc_state* s, int* n_mbs, int* n_events) {
    *n_mbs = (int)s->t_mb_type.size();
    *n_events = (int)s->t_run.size();
    return 0;
}
int orc_trace_copy(orc_state* s, int8_t* mb_type, int8_t* coded, uint8_t* quant, int16_t* mv, int16_t* intradc,
                   uint8_t* nev, uint8_t* run, int16_t* level) {
    size_t n = s->t_mb_type.size(), ne = s->t_run.size();
    std::memcpy(mb_type, s->t_mb_type.data(), n);
    std::memcpy(coded, s->t_coded.data(), n);
    std::memcpy(quant, s->t_quant.data(), n);
    std::memcpy(mv, s->t_mv.data(), n * 8 * sizeof(int16_t));
    std::memcpy(intradc, s->t_intradc.data(), n * 6 * sizeof(int16_t));
    std::memcpy(nev, s->t_nev.data(), n * 6);
    std::memcpy(run, s->t_run.data(), ne);
    std::memcpy(level, s->t_level.data(), ne * sizeof(int16_t));
    return 0;
}

void orc_yuv420_to_rgba(const uint8_t* y, const uint8_t* cb, const uint8_t* cr, size_t y_len, size_t y_width,
                        uint8_t* out) {
    yuv420_to_rgba(y, cb, cr, y_len, y_width, out);
}
void orc_deblock(const uint8_t* in, size_t len, size_t width, int strength, uint8_t* out) {
    deblock(in, len, width, strength, out);
}
void orc_deblock_process(uint8_t* abcd, int strength, int simd) {
    if (simd)
        process_simd_lane(abcd, abcd + 1, abcd + 2, abcd + 3, strength);
    else
        process_scalar(abcd, abcd + 1, abcd + 2, abcd + 3, strength);
}
int orc_quant_to_strength(int quant) { return (quant >= 0 && quant < 32) ? QUANT_TO_STRENGTH[quant] : -1; }

int orc_inverse_rle(int intradc_code, int n_events, const uint8_t* run, const int16_t* level, int quant,
                    float* coefs) {
    Block b;
    b.intradc = intradc_code;
    for (int i = 0; i < n_events; i++) b.tcoef.push_back({true, run[i], level[i]});
    DecodedDctBlock d;
    inverse_rle(b, &d, (uint8_t)quant);
    for (int i = 0; i < 64; i++) coefs[i] = 0.0f;
    switch (d.cls) {
        case DCT_DC: coefs[0] = d.dc; break;
        case DCT_HORIZ:
            for (int i = 0; i < 8; i++) coefs[i] = d.vec[i];
            break;
        case DCT_VERT:
            for (int i = 0; i < 8; i++) coefs[8 * i] = d.vec[i];
            break;
        case DCT_FULL: std::memcpy(coefs, d.full, sizeof(d.full)); break;
        default: break;
    }
    return d.cls;
}

void orc_idct_block(int cls, const float* coefs, uint8_t* pixels) {
    std::vector<DecodedDctBlock> levels(1);
    DecodedDctBlock& d = levels[0];
    d.cls = cls;
    d.dc = coefs[0];
    if (cls == DCT_HORIZ)
        for (int i = 0; i < 8; i++) d.vec[i] = coefs[i];
    if (cls == DCT_VERT)
        for (int i = 0; i < 8; i++) d.vec[i] = coefs[8 * i];
    std::memcpy(d.full, coefs, sizeof(d.full));
    std::vector<uint8_t> out(pixels, pixels + 64);
    idct_channel(levels, out, 1, 8);
    std::memcpy(pixels, out.data(), 64);
}
void orc_idct_1d(const float* in, float* out) { idct_1d(in, out); }

void orc_gather_block(const uint8_t* src, int width, int height, int pos_x, int pos_y, int mv_x, int mv_y,
                      uint8_t* dst) {
    MotionVector mv;
    mv.x = (int16_t)mv_x, mv.y = (int16_t)mv_y;
    gather_block(src, (size_t)width * height, (size_t)width, (size_t)pos_x, (size_t)pos_y, mv, dst);
}
int orc_average_sum_of_mvs(int sum) { return average_sum_of_mvs((int16_t)sum); }
int orc_halfpel_decode(int predictor, int mvd) {
    Picture p;
    return halfpel_decode(0, p, (int16_t)predictor, (int16_t)mvd);
}
int orc_median_of(int a, int b, int c) { return median_of((int16_t)a, (int16_t)b, (int16_t)c); }

int orc_read_vlc(int table, const uint8_t* data, size_t len, size_t* bitpos, int* out4) {
    ensure_trees();
    if (table < 0 || table > 4) return ORC_ERR_INTERNAL_DECODER_ERROR;
    Reader r{data, len, *bitpos};
    const VlcCode* c;
    Err e = r.read_vlc(g_trees[table], &c);
    *bitpos = r.bits_read;
    if (e) return e;
    out4[0] = c->kind, out4[1] = c->a, out4[2] = c->b, out4[3] = c->c;
    return ORC_OK;
}

int orc_read_bits(const uint8_t* data, size_t len, size_t* bitpos, int nbits, int is_signed, int peek,
                  int64_t* value) {
    Reader r{data, len, *bitpos};
    Err e;
    if (is_signed) {
        e = r.read_signed_bits((uint32_t)nbits, value);
    } else {
        uint64_t v = 0;
        e = r.read_bits((uint32_t)nbits, &v);
        *value = (int64_t)v;
    }
    if (!e && !peek) *bitpos = r.bits_read;
    return e;
}

int orc_recognize_start_code(const uint8_t* data, size_t len, size_t bitpos, int in_error, int* skipped) {
    Reader r{data, len, bitpos};
    bool found;
    uint32_t sk = 0;
    Err e = r.recognize_start_code(in_error != 0, &found, &sk);
    if (e) return e;
    *skipped = found ? (int)sk : -1;
    return ORC_OK;
}

int orc_decode_block(const uint8_t* data, size_t len, size_t* bitpos, int decoder_options, int version,
                     int is_intra, int tcoef_present, int* intradc_code, int* n_events, uint8_t* run,
                     int16_t* level, uint8_t* is_short, int cap) {
    ensure_trees();
    Reader r{data, len, *bitpos};
    Picture pic;
    pic.version = version;
    Block b;
    Err e = decode_block(r, decoder_options, pic, 0, is_intra ? MB_INTRA : MB_INTER, tcoef_present != 0, &b);
    *bitpos = r.bits_read;
    if (e) return e;
    *intradc_code = b.intradc;
    *n_events = (int)b.tcoef.size();
    for (int i = 0; i < (int)b.tcoef.size() && i < cap; i++) {
        run[i] = b.tcoef[i].run;
        level[i] = b.tcoef[i].level;
        is_short[i] = b.tcoef[i].is_short;
    }
    return ORC_OK;
}

double orc_bench_decode(const uint8_t* blob, const uint64_t* pkt_off, const uint32_t* pkt_len,
                        const uint32_t* pic_first, int n_streams, int decoder_options, int do_deblock,
                        int threads, uint64_t* pixels, uint64_t* checksum) {
    ensure_trees();
    if (threads < 1) threads = 1;
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    std::atomic<uint64_t> px_total{0}, ck_total{0};
    auto worker = [&]() {
        uint64_t px = 0, ck = 0;
        std::vector<uint8_t> rgba, dy, dcb, dcr;
        for (;;) {
            int s = next.fetch_add(1);
            if (s >= n_streams) break;
            orc_state st;
            st.decoder_options = decoder_options;
            for (uint32_t i = pic_first[s]; i < pic_first[s + 1]; i++) {
                Reader r{blob + pkt_off[i], pkt_len[i], 0};
                Err e = decode_next_picture(&st, r);
                if (e) {
                    failed.store(e);
                    return;
                }
                const DecodedPicture& p = st.last;
                rgba.resize(p.luma.size() * 4);
                if (do_deblock) {
                    int strength = QUANT_TO_STRENGTH[p.header.quantizer & 31];
                    dy.resize(p.luma.size());
                    dcb.resize(p.chroma_b.size());
                    dcr.resize(p.chroma_r.size());
                    deblock(p.luma.data(), p.luma.size(), (size_t)p.w, strength, dy.data());
                    deblock(p.chroma_b.data(), p.chroma_b.size(), p.chroma_samples_per_row, strength, dcb.data());
                    deblock(p.chroma_r.data(), p.chroma_r.size(), p.chroma_samples_per_row, strength, dcr.data());
                    yuv420_to_rgba(dy.data(), dcb.data(), dcr.data(), dy.size(), (size_t)p.w, rgba.data());
                } else {
                    yuv420_to_rgba(p.luma.data(), p.chroma_b.data(), p.chroma_r.data(), p.luma.size(), (size_t)p.w,
                                   rgba.data());
                }
                px += p.luma.size();
                uint64_t pid = ((uint64_t)s << 20) + (i - pic_first[s]);
                ck += weighted_sum(rgba.data(), rgba.size()) * ((0x9E3779B97F4A7C15ull * (pid + 1)) | 1ull);
            }
        }
        px_total.fetch_add(px);
        ck_total.fetch_add(ck);
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (failed.load()) return -(double)failed.load();
    *pixels = px_total.load();
    *checksum = ck_total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Step-wise baseline driver.  Worker threads live as long as the batch (they are not created per step), every stream
// keeps its RGBA output buffer (what a consumer of the reference would hold), and the optional checksum is computed
// AFTER the timed region: the seconds returned cover decode_next_picture + [deblock] + yuv420_to_rgba only.
struct orc_batch {
    std::vector<orc_state> states;
    std::vector<std::vector<uint8_t>> rgba;  // per stream
    uint64_t step_index = 0;
    // persistent pool
    std::vector<std::thread> pool;
    std::mutex m;
    std::condition_variable cv, done;
    uint64_t generation = 0;
    int running = 0;
    bool quit = false;
    // job of the current step
    const uint8_t* blob = nullptr;
    const uint64_t* pkt_off = nullptr;
    const uint32_t* pkt_len = nullptr;
    int do_deblock = 0;
    std::atomic<int> next{0}, failed{0};
    std::atomic<uint64_t> px_total{0};

    void work() {
        uint64_t px = 0;
        std::vector<uint8_t> dy, dcb, dcr;
        const int n = (int)states.size();
        for (;;) {
            int s = next.fetch_add(1);
            if (s >= n) break;
            orc_state& st = states[(size_t)s];
            Reader r{blob + pkt_off[s], pkt_len[s], 0};
            Err e = decode_next_picture(&st, r);
            if (e) {
                failed.store(e);
                continue;
            }
            const DecodedPicture& p = st.last;
            std::vector<uint8_t>& out = rgba[(size_t)s];
            out.resize(p.luma.size() * 4);
            if (do_deblock) {
                int strength = QUANT_TO_STRENGTH[p.header.quantizer & 31];
                dy.resize(p.luma.size());
                dcb.resize(p.chroma_b.size());
                dcr.resize(p.chroma_r.size());
                deblock(p.luma.data(), p.luma.size(), (size_t)p.w, strength, dy.data());
                deblock(p.chroma_b.data(), p.chroma_b.size(), p.chroma_samples_per_row, strength, dcb.data());
                deblock(p.chroma_r.data(), p.chroma_r.size(), p.chroma_samples_per_row, strength, dcr.data());
                yuv420_to_rgba(dy.data(), dcb.data(), dcr.data(), dy.size(), (size_t)p.w, out.data());
            } else {
                yuv420_to_rgba(p.luma.data(), p.chroma_b.data(), p.chroma_r.data(), p.luma.size(), (size_t)p.w, out.data());
            }
            px += p.luma.size();
        }
        px_total.fetch_add(px);
    }
    void thread_main() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
            }
            work();
            std::lock_guard<std::mutex> lk(m);
            if (--running == 0) done.notify_all();
        }
    }
    ~orc_batch() {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv.notify_all();
        for (auto& t : pool) t.join();
    }
};

orc_batch* orc_batch_new(int n_streams, int decoder_options) {
    ensure_trees();
    orc_batch* b = new orc_batch();
    b->states.resize((size_t)n_streams);
    b->rgba.resize((size_t)n_streams);
    for (auto& s : b->states) s.decoder_options = decoder_options;
    return b;
}
void orc_batch_free(orc_batch* b) { delete b; }

double orc_batch_step(orc_batch* b, const uint8_t* blob, const uint64_t* pkt_off, const uint32_t* pkt_len,
                      int do_deblock, int threads, uint64_t* pixels, uint64_t* checksum) {
    const int n = (int)b->states.size();
    if (threads < 1) threads = 1;
    if (threads > n) threads = n;
    // the pool grows to the largest thread count asked for (outside the timed region)
    while ((int)b->pool.size() < threads - 1) b->pool.emplace_back(&orc_batch::thread_main, b);
    const int helpers = (int)b->pool.size();  // every pool thread takes part
    const uint64_t step_index = b->step_index++;
    b->blob = blob, b->pkt_off = pkt_off, b->pkt_len = pkt_len, b->do_deblock = do_deblock;
    b->next.store(0), b->failed.store(0), b->px_total.store(0);
    auto t0 = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lk(b->m);
        b->running = helpers;
        b->generation++;
    }
    b->cv.notify_all();
    b->work();
    {
        std::unique_lock<std::mutex> lk(b->m);
        b->done.wait(lk, [&] { return b->running == 0; });
    }
    auto t1 = std::chrono::steady_clock::now();
    if (b->failed.load()) return -(double)b->failed.load();
    if (pixels) *pixels += b->px_total.load();
    if (checksum) {  // untimed: position-weighted checksum of every stream's RGBA
        uint64_t ck = 0;
        for (int s = 0; s < n; s++) {
            uint64_t pid = ((uint64_t)s << 20) + step_index;
            ck += weighted_sum(b->rgba[(size_t)s].data(), b->rgba[(size_t)s].size()) * ((0x9E3779B97F4A7C15ull * (pid + 1)) | 1ull);
        }
        *checksum += ck;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_batch_stream_yuv(orc_batch* b, int s, uint8_t* y, uint8_t* cb, uint8_t* cr) {
    if (s < 0 || s >= (int)b->states.size()) return -1;
    return orc_last_picture_yuv(&b->states[(size_t)s], y, cb, cr);
}

}  // extern "C"
