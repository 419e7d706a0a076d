// context.cu -- device context, step objects and the device half of the C ABI
// (include/h263cu.h).  The context owns all device memory: per stream two reconstruction
// slots (current / reference, ping-pong) for Y and the interleaved CbCr plane and a two-deep RGBA ring; side
// info arrives through pinned cudaMemcpyAsync.  Everything here is plumbing around the
// kernels in kernels.cu -- there is no CPU fallback: without a usable device the entry
// points return H263CU_ERR_NO_DEVICE / H263CU_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "device_math.cuh"
#include "kernels.cuh"

using namespace h263dev;

// frontend.cpp: the threaded parse with the parsers' state change held back until parse_step_finish(accept)
namespace h263fe {
int parse_step_deferred(h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                        const uint32_t* stream_ids, uint32_t n, int threads, h263cu_pic* pics, h263cu_mb* mbs, uint32_t mb_cap,
                        h263cu_event* events, uint32_t ev_cap, uint32_t* n_pics_out, uint32_t* n_mbs_out, uint32_t* n_units_out,
                        int* per_pic_err, int32_t* pic_of_input, uint32_t max_w, uint32_t max_h);
void parse_step_finish(h263cu_parser* const* parsers, uint32_t n, bool accept);
}  // namespace h263fe

extern "C" const uint8_t h263cu_quant_to_strength[32] = H263_QUANT_TO_STRENGTH;

namespace {

#define CU_TRY(expr)                      \
    do {                                  \
        cudaError_t _e = (expr);          \
        if (_e != cudaSuccess) return map_cuda_error(_e); \
    } while (0)

int map_cuda_error(cudaError_t e) {
    switch (e) {
        case cudaErrorNoDevice:
        case cudaErrorInsufficientDriver:
        case cudaErrorInvalidDevice: return H263CU_ERR_NO_DEVICE;
        case cudaErrorMemoryAllocation: return H263CU_ERR_OUT_OF_MEMORY;
        default: return H263CU_ERR_CUDA;
    }
}

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct StreamState {
    bool has_pic = false;
    uint8_t cur_slot = 0;  // slot that holds the last decoded picture
    int8_t rgba_slot = -1; // ring slot that holds the last picture's RGBA (-1 = none)
    uint16_t w = 0, h = 0;
    uint8_t pic_type = 0, pquant = 0;
    uint16_t tr = 0;
    uint32_t stamp = 0;
    uint32_t seen = 0;    // epoch of the last h263cu_decode_step that named this stream (duplicate check)
    // the prediction source of the next picture: the last NON-disposable picture.  Identical to the last picture unless
    // the stream carries disposable P pictures (H263CU_PICFLAG_DISPOSABLE), which are shown but never predicted from.
    bool has_ref = false;
    uint8_t ref_slot = 0;
    uint16_t ref_w = 0, ref_h = 0;
    bool padded = false;  // the reference picture's planes carry the replicated border (tiled kernel)
};

// where the next picture of a stream is reconstructed: never over its prediction source
inline int next_slot(const StreamState& st) { return st.has_ref ? (st.ref_slot ^ 1) : (st.has_pic ? (st.cur_slot ^ 1) : 0); }

// state.rs:464-483 on the device side: the picture becomes the stream's last picture and, unless it is disposable,
// the reference of the next one
inline void advance_stream(StreamState& st, const h263cu_pic& p, bool want_rgba, int rgba_ring, bool tiled) {
    const int slot = next_slot(st);
    st.cur_slot = (uint8_t)slot;
    st.has_pic = true;
    st.w = p.width, st.h = p.height;
    st.pic_type = p.pic_type, st.pquant = p.pquant, st.tr = p.temporal_reference;
    st.rgba_slot = want_rgba ? (int8_t)rgba_ring : (int8_t)-1;
    if (!(p.flags & H263CU_PICFLAG_DISPOSABLE)) {
        st.has_ref = true;
        st.ref_slot = (uint8_t)slot;
        st.ref_w = p.width, st.ref_h = p.height;
        st.padded = tiled;
    }
}

}  // namespace

struct h263cu_step {
    std::vector<h263cu_pic> pics;
    h263cu_mb* d_mbs = nullptr;
    h263cu_event* d_events = nullptr;
    uint32_t n_mbs = 0, n_units = 0;
    size_t mb_cap = 0, ev_cap = 0;  // capacities in elements (ring reuse)
    uint32_t max_w = 0, max_h = 0;
};

struct h263cu_ctx {
    int device = 0;
    uint32_t max_streams = 0, max_w = 0, max_h = 0;
    uint32_t mbw = 0, mbh = 0;
    uint32_t pitch_y = 0, pitch_c = 0, rgba_pitch = 0;
    size_t y_slot = 0, c_slot = 0, rgba_slot = 0;
    uint8_t *y_pool = nullptr, *c_pool = nullptr, *rgba_pool = nullptr;  // c_pool: interleaved CbCr planes
    CUtensorMap rgba_map;  // TMA view of rgba_pool: rows of rgba_pitch bytes, box 64 bytes x 16 rows (one macroblock)
    std::vector<StreamState> streams;
    uint32_t stamp = 0, decode_epoch = 0;
    uint32_t rgba_parity = 0;

    cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_pics = nullptr;  // descriptor uploads: overlap the previous step's kernels
    // PicDev staging ring (pinned host + device)
    static constexpr int PIC_RING = 4;
    PicDev* h_pics[PIC_RING] = {};
    PicDev* d_pics[PIC_RING] = {};
    cudaEvent_t pics_done[PIC_RING] = {};  // the kernels that read d_pics[i] have finished
    cudaEvent_t pics_up[PIC_RING] = {};    // d_pics[i] is uploaded
    size_t pics_cap = 0;
    int pic_ring_pos = 0;
    // side-info ring used by h263cu_submit_step*
    h263cu_step ring[2];
    cudaEvent_t ring_h2d_done[2] = {}, ring_run_done[2] = {};
    int ring_pos = 0;
    // pinned staging of h263cu_decode_step, one set per side-info ring slot (grow only)
    struct Staging {
        h263cu_pic* pics = nullptr;
        h263cu_mb* mbs = nullptr;
        h263cu_event* events = nullptr;
        size_t pic_cap = 0, mb_cap = 0, ev_cap = 0;
    } staging[2];
    std::vector<int32_t> pic_of_input;
    std::vector<uint64_t> rgba_offsets;
    // RGBA ring read-back tracking
    cudaEvent_t rgba_written[2] = {}, rgba_read[2] = {};
    // timing
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint64_t launches = 0, tiled_launches = 0;
    // optional per-kernel timing (h263cu_profile_*): event pairs, kind 0 = recon, 1 = deblock
    bool profiling = false;
    std::vector<cudaEvent_t> prof_free;
    struct ProfSpan {
        cudaEvent_t a, b;
        int kind;
    };
    std::vector<ProfSpan> prof_spans;
    // checksum scratch
    ChecksumJob* d_jobs = nullptr;
    unsigned long long* d_sums = nullptr;
    size_t jobs_cap = 0;

    int force_kernel = 0;  // H263CU_KERNEL=mb|tile overrides the per-step choice (A/B checks)
    // host time spent inside h263cu_decode_step, split into the bitstream parse and everything else (staging, driver
    // calls): the north-star asks for the host parse time beside the device time
    double host_parse_s = 0.0, host_other_s = 0.0;
    uint64_t host_calls = 0;
    // interior origin (pixel 0,0) of a plane; the padding lies at negative offsets
    uint8_t* plane(int p, uint32_t stream, int slot) const {
        const size_t idx = (size_t)stream * 2 + (size_t)slot;
        if (p == 0) return y_pool + idx * y_slot + (size_t)PAD_Y_ROWS * pitch_y + PAD_Y_COLS;
        // Cb and Cr share one interleaved plane: Cr starts one byte after Cb, samples are CHROMA_STEP bytes apart
        return c_pool + idx * c_slot + (size_t)PAD_C_ROWS * pitch_c + PAD_C_COLS + (p == 2 ? 1 : 0);
    }
    uint8_t* rgba(uint32_t stream, int slot) const { return rgba_pool + ((size_t)slot * max_streams + stream) * rgba_slot; }
};

namespace {

int ensure_pic_ring(h263cu_ctx* c, size_t n) {
    if (n <= c->pics_cap) return 0;
    CU_TRY(cudaStreamSynchronize(c->s_main));
    size_t cap = std::max<size_t>(n, 256);
    for (int i = 0; i < h263cu_ctx::PIC_RING; i++) {
        if (c->h_pics[i]) cudaFreeHost(c->h_pics[i]);
        if (c->d_pics[i]) cudaFree(c->d_pics[i]);
        c->h_pics[i] = nullptr, c->d_pics[i] = nullptr;
        CU_TRY(cudaHostAlloc((void**)&c->h_pics[i], cap * sizeof(PicDev), cudaHostAllocDefault));
        CU_TRY(cudaMalloc((void**)&c->d_pics[i], cap * sizeof(PicDev)));
    }
    c->pics_cap = cap;
    return 0;
}

int step_reserve(h263cu_step* s, size_t n_mbs, size_t n_units) {
    if (n_mbs > s->mb_cap) {
        if (s->d_mbs) cudaFree(s->d_mbs);
        s->d_mbs = nullptr;
        size_t cap = n_mbs + n_mbs / 8 + 64;
        CU_TRY(cudaMalloc((void**)&s->d_mbs, cap * sizeof(h263cu_mb)));
        s->mb_cap = cap;
    }
    if (n_units + 8 > s->ev_cap) {
        if (s->d_events) cudaFree(s->d_events);
        s->d_events = nullptr;
        size_t cap = n_units + n_units / 8 + 64;
        CU_TRY(cudaMalloc((void**)&s->d_events, cap * sizeof(h263cu_event)));
        s->ev_cap = cap;
    }
    return 0;
}

int prof_take(h263cu_ctx* c, cudaEvent_t* a, cudaEvent_t* b) {
    for (cudaEvent_t* e : {a, b}) {
        if (!c->prof_free.empty()) {
            *e = c->prof_free.back();
            c->prof_free.pop_back();
        } else {
            CU_TRY(cudaEventCreate(e));
        }
    }
    return 0;
}
void prof_end(h263cu_ctx* c, cudaEvent_t a, cudaEvent_t b, int kind) {
    cudaEventRecord(b, c->s_main);
    c->prof_spans.push_back({a, b, kind});
}

// The device-side descriptor of one picture, given the state of its stream BEFORE the picture (run_step and the graph
// builder, which walks a simulated copy of the state, share it).
PicDev make_picdev(const h263cu_ctx* c, const h263cu_pic& p, const StreamState& st, bool want_rgba, int rgba_ring) {
    const int ref_slot = st.ref_slot, new_slot = next_slot(st);
    PicDev d;
    std::memset(&d, 0, sizeof(d));
    for (int k = 0; k < 3; k++) {
        d.cur[k] = c->plane(k, p.stream, new_slot);
        d.ref[k] = st.has_ref ? c->plane(k, p.stream, ref_slot) : nullptr;
    }
    d.rgba = want_rgba ? c->rgba(p.stream, rgba_ring) : nullptr;
    d.cur_y4 = (uint32_t)((d.cur[0] - c->y_pool) >> 2);
    d.cur_c4 = (uint32_t)((d.cur[1] - c->c_pool) >> 2);
    d.ref_y4 = st.has_ref ? (uint32_t)((d.ref[0] - c->y_pool) >> 2) : 0u;
    d.ref_c4 = st.has_ref ? (uint32_t)((d.ref[1] - c->c_pool) >> 2) : 0u;
    d.rgba_row0 = want_rgba ? (uint32_t)((size_t)(d.rgba - c->rgba_pool) / c->rgba_pitch) : 0u;
    d.first_event = p.first_event;
    d.rgba_pitch = c->rgba_pitch;
    d.w = p.width, d.h = p.height;
    d.cw = (uint16_t)((p.width + 1) / 2), d.ch = (uint16_t)((p.height + 1) / 2);
    d.pitch_y = (uint16_t)c->pitch_y, d.pitch_c = (uint16_t)c->pitch_c;
    d.strength = h263cu_quant_to_strength[p.pquant & 31];
    d.flags = p.flags;
    return d;
}

// Validates a step against the context and the per-stream state, builds the PicDev array,
// enqueues the kernels on s_main and advances the per-stream reference bookkeeping
// (state.rs:464-483: the picture just decoded becomes the reference of the next one).
// Nothing of the per-stream state changes unless the kernels have been enqueued: a step that is
// refused, or that fails on the way to the launch, leaves every stream as it was.
int run_step(h263cu_ctx* c, h263cu_step* s, uint32_t out_flags, bool lean = false) {
    const uint32_t n = (uint32_t)s->pics.size();
    if (n == 0) return 0;
    if (n > 65535) return H263CU_ERR_CAPACITY;  // h263cu_mb.pic is 16 bits wide
    const bool want_rgba = (out_flags & H263CU_OUT_RGBA) != 0;
    const bool want_deblock = want_rgba && (out_flags & H263CU_OUT_DEBLOCK) != 0;
    const uint32_t stamp = ++c->stamp;  // marks the streams named by this step (duplicate check only)
    uint32_t max_w = 0, max_h = 0;
    bool tiled = true, aligned16 = true, wide_mv = false;
    for (uint32_t i = 0; i < n; i++) {
        const h263cu_pic& p = s->pics[i];
        if (p.stream >= c->max_streams) return H263CU_ERR_CAPACITY;
        if (p.width == 0 || p.height == 0 || p.width > c->max_w || p.height > c->max_h) return H263CU_ERR_CAPACITY;
        if ((uint32_t)p.mb_w * 16 < p.width || (uint32_t)p.mb_h * 16 < p.height ||
            (uint32_t)p.mb_w > c->mbw || (uint32_t)p.mb_h > c->mbh)
            return H263CU_ERR_BAD_ARGUMENT;
        if ((uint64_t)p.first_mb + p.n_mbs > s->n_mbs || (uint64_t)p.first_event + p.n_event_units > s->n_units ||
            p.n_mbs != (uint32_t)p.mb_w * p.mb_h)
            return H263CU_ERR_BAD_ARGUMENT;
        StreamState& st = c->streams[p.stream];
        if (st.stamp == stamp) return H263CU_ERR_BAD_ARGUMENT;  // a stream appears once per step
        st.stamp = stamp;
        if (p.flags & H263CU_PICFLAG_HAS_INTER) {
            if (!st.has_ref) return H263CU_ERR_UNCODED_IFRAME_BLOCKS;             // gather.rs:149
            if (st.ref_w != p.width || st.ref_h != p.height) return H263CU_ERR_REFERENCE_WOULD_ABORT;
        }
        max_w = std::max<uint32_t>(max_w, p.width);
        max_h = std::max<uint32_t>(max_h, p.height);
        // the tiled kernel needs references with a replicated border; sizes that are not multiples of 16 take its
        // edge fix-up instantiation, and the register-resident deblock kernel needs aligned pictures
        if ((p.width | p.height) & 15) aligned16 = false;
        if ((p.flags & H263CU_PICFLAG_HAS_INTER) && !(p.flags & H263CU_PICFLAG_MV_IN_RANGE)) wide_mv = true;
        if ((p.flags & H263CU_PICFLAG_HAS_INTER) && !st.padded) tiled = false;
    }
    if (c->force_kernel == 1) tiled = false;
    int e = ensure_pic_ring(c, n);
    if (e) return e;
    const int slot = c->pic_ring_pos;
    CU_TRY(cudaEventSynchronize(c->pics_done[slot]));
    const int rgba_ring = (int)(c->rgba_parity & 1u);
    PicDev* hp = c->h_pics[slot];
    for (uint32_t i = 0; i < n; i++) {
        const h263cu_pic& p = s->pics[i];
        const PicDev d = make_picdev(c, p, c->streams[p.stream], want_rgba, rgba_ring);
        hp[i] = d;
    }
    cudaEvent_t pa = nullptr, pb = nullptr, qa = nullptr, qb = nullptr;
    if (c->profiling) {
        // the event pairs are taken before anything is enqueued: a failure here leaves no trace
        if ((e = prof_take(c, &pa, &pb))) return e;
        if (want_deblock && (e = prof_take(c, &qa, &qb))) {
            c->prof_free.push_back(pa), c->prof_free.push_back(pb);
            return e;
        }
    }
    // the descriptors travel on their own stream, so the copy overlaps the previous step's kernels
    // (pics_done[slot] was waited for above: the kernels that read this ring slot four steps ago are done).
    // Lean form (one small step, e.g. a single stream's picture): everything goes to s_main in order -- there is
    // nothing to overlap, and every driver call saved is a microsecond or two of the caller's time per picture.
    if (lean) {
        CU_TRY(cudaMemcpyAsync(c->d_pics[slot], hp, n * sizeof(PicDev), cudaMemcpyHostToDevice, c->s_main));
    } else {
        CU_TRY(cudaMemcpyAsync(c->d_pics[slot], hp, n * sizeof(PicDev), cudaMemcpyHostToDevice, c->s_pics));
        CU_TRY(cudaEventRecord(c->pics_up[slot], c->s_pics));
        CU_TRY(cudaStreamWaitEvent(c->s_main, c->pics_up[slot], 0));
    }
    if (want_rgba) {
        // do not overwrite an RGBA ring slot that is still being read back
        CU_TRY(cudaStreamWaitEvent(c->s_main, c->rgba_read[rgba_ring], 0));
    }
    if (pa) cudaEventRecord(pa, c->s_main);
    const Pools pools{c->y_pool, c->c_pool, c->rgba_pool, c->pitch_y, c->pitch_c, c->rgba_pitch};
    launch_recon(c->d_pics[slot], s->d_mbs, s->d_events, s->n_mbs, want_rgba && !want_deblock, tiled ? (aligned16 ? 1 : 2) : 0, wide_mv, pools,
                 &c->rgba_map, c->s_main);
    if (pa) prof_end(c, pa, pb, 0);
    if (want_deblock) {
        if (qa) cudaEventRecord(qa, c->s_main);
        if (aligned16 && c->force_kernel != 1)
            launch_deblock_rgba_tile(c->d_pics[slot], n, max_w, max_h, c->s_main);
        else
            launch_deblock_rgba(c->d_pics[slot], n, max_w, max_h, c->s_main);
        if (qa) prof_end(c, qa, qb, 1);
    }
    CU_TRY(cudaGetLastError());
    // ---- the step is on the device: commit the bookkeeping ----
    c->pic_ring_pos = (c->pic_ring_pos + 1) % h263cu_ctx::PIC_RING;
    c->launches += want_deblock ? 2 : 1;
    if (tiled) c->tiled_launches++;
    for (uint32_t i = 0; i < n; i++) advance_stream(c->streams[s->pics[i].stream], s->pics[i], want_rgba, rgba_ring, tiled);
    CU_TRY(cudaEventRecord(c->pics_done[slot], c->s_main));
    if (want_rgba) {
        if (!lean) CU_TRY(cudaEventRecord(c->rgba_written[rgba_ring], c->s_main));  // lean: the read-back follows on s_main
        c->rgba_parity++;
    }
    return 0;
}

// Checks caller-built side info before it goes to the device, where the kernels index with it unchecked: every
// record belongs to the picture whose range it lies in and sits at its raster position, its events stay inside the
// picture's share of the event array.  HAS_INTER / MV_IN_RANGE are derived from the records, not taken on trust: a
// picture with an inter macroblock gets HAS_INTER, one with a vector beyond [-32, 31] loses MV_IN_RANGE (and takes
// the clamped-fetch instantiation).  Side info produced by the library's own parser skips this pass.
int validate_side_info(std::vector<h263cu_pic>& pics, const h263cu_mb* mbs, uint32_t n_mbs, uint32_t n_units) {
    for (size_t i = 0; i < pics.size(); i++) {
        h263cu_pic& p = pics[i];
        if ((uint64_t)p.first_mb + p.n_mbs > n_mbs || (uint64_t)p.first_event + p.n_event_units > n_units ||
            p.n_mbs != (uint32_t)p.mb_w * p.mb_h || p.mb_w == 0)
            return H263CU_ERR_BAD_ARGUMENT;
        bool any_inter = false, in_range = true;
        const h263cu_mb* m = mbs + p.first_mb;
        for (uint32_t k = 0; k < p.n_mbs; k++, m++) {
            if (m->pic != i || (uint32_t)m->mby * p.mb_w + m->mbx != k || m->mbx >= p.mb_w) return H263CU_ERR_BAD_ARGUMENT;
            uint32_t ev = 0;
            for (int b = 0; b < 6; b++) ev += m->nev[b];
            if (m->flags & H263CU_MB_WIDE) ev *= 2;
            if ((uint64_t)m->ev_off + ev > p.n_event_units) return H263CU_ERR_BAD_ARGUMENT;
            if (m->flags & H263CU_MB_INTER) {
                any_inter = true;
                for (int b = 0; b < 4; b++)
                    in_range &= m->u.mv[b][0] >= -32 && m->u.mv[b][0] <= 31 && m->u.mv[b][1] >= -32 && m->u.mv[b][1] <= 31;
            }
        }
        if (any_inter) p.flags |= H263CU_PICFLAG_HAS_INTER;
        if (!in_range) p.flags &= (uint8_t)~H263CU_PICFLAG_MV_IN_RANGE;
    }
    return 0;
}

int upload_into(h263cu_ctx* c, h263cu_step* s, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs,
                uint32_t n_mbs, const h263cu_event* events, uint32_t n_units, cudaStream_t stream, bool trusted) {
    int e = step_reserve(s, n_mbs, n_units);
    if (e) return e;
    try {
        s->pics.assign(pics, pics + n_pics);
    } catch (const std::bad_alloc&) {
        return H263CU_ERR_OUT_OF_MEMORY;
    }
    s->n_mbs = n_mbs, s->n_units = n_units;
    if (!trusted && (e = validate_side_info(s->pics, mbs, n_mbs, n_units))) {
        s->pics.clear();
        return e;
    }
    if (n_mbs) CU_TRY(cudaMemcpyAsync(s->d_mbs, mbs, (size_t)n_mbs * sizeof(h263cu_mb), cudaMemcpyHostToDevice, stream));
    if (n_units)
        CU_TRY(cudaMemcpyAsync(s->d_events, events, (size_t)n_units * sizeof(h263cu_event), cudaMemcpyHostToDevice, stream));
    (void)c;
    return 0;
}

template <typename T>
int grow_pinned(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr, *cap = 0;
    const size_t n = need + need / 4 + 64;
    CU_TRY(cudaHostAlloc((void**)p, n * sizeof(T), cudaHostAllocDefault));
    *cap = n;
    return 0;
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
int make_rgba_map(CUtensorMap* map, void* base, uint64_t pitch, uint64_t rows) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    std::memset(map, 0, sizeof(*map));
    if (!recon_tile_uses_tma()) return 0;
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) {
        cudaGetLastError();
        return H263CU_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {pitch, rows}, strides[1] = {pitch};
    const cuuint32_t box[2] = {64, 16}, es[2] = {1, 1};
    if (enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return H263CU_ERR_CUDA;
    return 0;
}
void free_step_buffers(h263cu_step* s) {
    if (s->d_mbs) cudaFree(s->d_mbs);
    if (s->d_events) cudaFree(s->d_events);
    s->d_mbs = nullptr, s->d_events = nullptr;
    s->mb_cap = s->ev_cap = 0;
}

}  // namespace

extern "C" {

int h263cu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

h263cu_ctx* h263cu_create(int device, uint32_t max_streams, uint32_t max_width, uint32_t max_height, uint32_t flags,
                          int* err) {
    (void)flags;
    int dummy;
    if (!err) err = &dummy;
    *err = 0;
    if (max_streams == 0 || max_width == 0 || max_height == 0 || max_width > 4080 || max_height > 4080) {
        *err = H263CU_ERR_BAD_ARGUMENT;
        return nullptr;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        *err = H263CU_ERR_NO_DEVICE;
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        *err = H263CU_ERR_NO_DEVICE;
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        *err = H263CU_ERR_CUDA;
        return nullptr;
    }
    h263cu_ctx* c = new (std::nothrow) h263cu_ctx();
    if (!c) {
        *err = H263CU_ERR_OUT_OF_MEMORY;
        return nullptr;
    }
    c->device = device;
    c->max_streams = max_streams, c->max_w = max_width, c->max_h = max_height;
    c->mbw = (max_width + 15) / 16, c->mbh = (max_height + 15) / 16;
    // planes are MB-rounded so that whole-macroblock stores never need predication; the
    // padding is never read as picture content (sample coordinates clamp to the true size)
    // plus a border (PAD_*) into which the tiled kernel replicates the edge pixels, so that
    // motion compensation needs no per-sample clamping (unrestricted-MV extension)
    c->pitch_y = c->mbw * 16 + 2 * PAD_Y_COLS;
    c->pitch_c = c->mbw * 8 * CHROMA_STEP + 2 * PAD_C_COLS;  // interleaved CbCr rows: the same pitch as luma
    c->rgba_pitch = c->mbw * 16 * 4;
    c->y_slot = (size_t)c->pitch_y * (c->mbh * 16 + 2 * PAD_Y_ROWS);
    c->c_slot = (size_t)c->pitch_c * (c->mbh * 8 + 2 * PAD_C_ROWS);
    if (const char* k = getenv("H263CU_KERNEL")) c->force_kernel = !strcmp(k, "mb") ? 1 : (!strcmp(k, "tile") ? 2 : 0);
    c->rgba_slot = (size_t)c->rgba_pitch * c->mbh * 16;
    c->streams.resize(max_streams);
    auto fail = [&](int code) {
        *err = code;
        h263cu_destroy(c);
        return (h263cu_ctx*)nullptr;
    };
    const size_t pad = 256;  // aligned-word prediction loads may run a few bytes past a row
    // the tiled kernel addresses the pools through 32-bit offsets: 4-byte units for the planes,
    // 16-byte units for RGBA
    if (c->y_slot * 2 * max_streams + pad >= (16ull << 30) || c->rgba_slot * 2 * max_streams + pad >= (64ull << 30)) return fail(H263CU_ERR_CAPACITY);
    if (cudaMalloc((void**)&c->y_pool, c->y_slot * 2 * max_streams + pad) != cudaSuccess) return fail(H263CU_ERR_OUT_OF_MEMORY);
    if (cudaMalloc((void**)&c->c_pool, c->c_slot * 2 * max_streams + pad) != cudaSuccess) return fail(H263CU_ERR_OUT_OF_MEMORY);
    if (cudaMalloc((void**)&c->rgba_pool, c->rgba_slot * 2 * max_streams + pad) != cudaSuccess) return fail(H263CU_ERR_OUT_OF_MEMORY);
    {
        // TMA tensor map over the RGBA pool: the reconstruction kernel stores one 64-byte x 16-row box per macroblock
        // (cp.async.bulk.tensor.2d) from a shared-memory tile
        const int e = make_rgba_map(&c->rgba_map, c->rgba_pool, c->rgba_pitch, (uint64_t)2 * max_streams * c->mbh * 16);
        if (e) return fail(e);
    }
    if (cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&c->s_pics, cudaStreamNonBlocking) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    for (int i = 0; i < h263cu_ctx::PIC_RING; i++)
        if (cudaEventCreateWithFlags(&c->pics_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->pics_up[i], cudaEventDisableTiming) != cudaSuccess)
            return fail(H263CU_ERR_CUDA);
    for (int i = 0; i < 2; i++) {
        if (cudaEventCreateWithFlags(&c->ring_h2d_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ring_run_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->rgba_written[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->rgba_read[i], cudaEventDisableTiming) != cudaSuccess)
            return fail(H263CU_ERR_CUDA);
    }
    if (cudaEventCreate(&c->t0) != cudaSuccess || cudaEventCreate(&c->t1) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    // planes start out zeroed (DecodedPicture::new, picture.rs:39-58); every macroblock of a
    // picture is rewritten by the kernel, so this only matters for defensive reads
    cudaMemsetAsync(c->y_pool, 0, c->y_slot * 2 * max_streams + pad, c->s_main);
    cudaMemsetAsync(c->c_pool, 0, c->c_slot * 2 * max_streams + pad, c->s_main);
    if (cudaStreamSynchronize(c->s_main) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    return c;
}

void h263cu_destroy(h263cu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->s_main) cudaStreamSynchronize(c->s_main);
    if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
    if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
    if (c->s_pics) cudaStreamSynchronize(c->s_pics);
    for (int i = 0; i < 2; i++) free_step_buffers(&c->ring[i]);
    for (auto& st : c->staging) {
        if (st.pics) cudaFreeHost(st.pics);
        if (st.mbs) cudaFreeHost(st.mbs);
        if (st.events) cudaFreeHost(st.events);
    }
    for (int i = 0; i < h263cu_ctx::PIC_RING; i++) {
        if (c->h_pics[i]) cudaFreeHost(c->h_pics[i]);
        if (c->d_pics[i]) cudaFree(c->d_pics[i]);
        if (c->pics_done[i]) cudaEventDestroy(c->pics_done[i]);
        if (c->pics_up[i]) cudaEventDestroy(c->pics_up[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (c->ring_h2d_done[i]) cudaEventDestroy(c->ring_h2d_done[i]);
        if (c->ring_run_done[i]) cudaEventDestroy(c->ring_run_done[i]);
        if (c->rgba_written[i]) cudaEventDestroy(c->rgba_written[i]);
        if (c->rgba_read[i]) cudaEventDestroy(c->rgba_read[i]);
    }
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    for (auto& sp : c->prof_spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (auto e : c->prof_free) cudaEventDestroy(e);
    if (c->d_jobs) cudaFree(c->d_jobs);
    if (c->d_sums) cudaFree(c->d_sums);
    if (c->y_pool) cudaFree(c->y_pool);
    if (c->c_pool) cudaFree(c->c_pool);
    if (c->rgba_pool) cudaFree(c->rgba_pool);
    if (c->s_main) cudaStreamDestroy(c->s_main);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    if (c->s_pics) cudaStreamDestroy(c->s_pics);
    delete c;
}

int h263cu_device_of(h263cu_ctx* c) { return c ? c->device : -1; }

void* h263cu_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void h263cu_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

h263cu_step* h263cu_step_upload(h263cu_ctx* c, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs,
                                uint32_t n_mbs, const h263cu_event* events, uint32_t n_units, int* err) {
    int dummy;
    if (!err) err = &dummy;
    *err = 0;
    if (!c || !pics || (!mbs && n_mbs) || (!events && n_units)) {
        *err = H263CU_ERR_BAD_ARGUMENT;
        return nullptr;
    }
    cudaSetDevice(c->device);
    h263cu_step* s = new (std::nothrow) h263cu_step();
    if (!s) {
        *err = H263CU_ERR_OUT_OF_MEMORY;
        return nullptr;
    }
    int e = upload_into(c, s, pics, n_pics, mbs, n_mbs, events, n_units, c->s_main, false);
    if (e) {
        free_step_buffers(s);
        delete s;
        *err = e;
        return nullptr;
    }
    return s;
}

void h263cu_step_free(h263cu_ctx* c, h263cu_step* s) {
    if (!s) return;
    if (c) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->s_main);
    }
    free_step_buffers(s);
    delete s;
}

int h263cu_step_run(h263cu_ctx* c, h263cu_step* s, uint32_t out_flags) {
    if (!c || !s) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    return run_step(c, s, out_flags);
}

static int submit_common(h263cu_ctx* c, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs, uint32_t n_mbs,
                         const h263cu_event* events, uint32_t n_units, uint32_t out_flags, bool trusted = false, bool lean = false) {
    if (!c || !pics || (!mbs && n_mbs) || (!events && n_units)) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    const int slot = c->ring_pos;
    h263cu_step* s = &c->ring[slot];
    if (n_mbs > s->mb_cap || (size_t)n_units + 8 > s->ev_cap) CU_TRY(cudaEventSynchronize(c->ring_run_done[slot]));
    int e;
    if (lean) {
        // in-order on s_main: the kernels that read this ring slot two submits ago precede these copies on the stream
        if ((e = upload_into(c, s, pics, n_pics, mbs, n_mbs, events, n_units, c->s_main, trusted))) return e;
        CU_TRY(cudaEventRecord(c->ring_h2d_done[slot], c->s_main));
    } else {
        // the copy engine may not overwrite side info that a kernel is still reading
        CU_TRY(cudaStreamWaitEvent(c->s_h2d, c->ring_run_done[slot], 0));
        if ((e = upload_into(c, s, pics, n_pics, mbs, n_mbs, events, n_units, c->s_h2d, trusted))) return e;
        CU_TRY(cudaEventRecord(c->ring_h2d_done[slot], c->s_h2d));
        CU_TRY(cudaStreamWaitEvent(c->s_main, c->ring_h2d_done[slot], 0));
    }
    e = run_step(c, s, out_flags, lean);
    if (e) return e;
    c->ring_pos ^= 1;  // the ring slot is taken only by a step that runs
    CU_TRY(cudaEventRecord(c->ring_run_done[slot], c->s_main));
    return 0;
}

int h263cu_submit_step(h263cu_ctx* c, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs, uint32_t n_mbs,
                       const h263cu_event* events, uint32_t n_units, uint32_t out_flags) {
    return submit_common(c, pics, n_pics, mbs, n_mbs, events, n_units, out_flags);
}

static int submit_readback(h263cu_ctx* c, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs, uint32_t n_mbs,
                           const h263cu_event* events, uint32_t n_units, uint32_t out_flags, uint8_t* host_rgba,
                           const uint64_t* rgba_offsets, bool trusted, bool lean = false) {
    if (!host_rgba) return H263CU_ERR_BAD_ARGUMENT;
    out_flags |= H263CU_OUT_RGBA;
    int e = submit_common(c, pics, n_pics, mbs, n_mbs, events, n_units, out_flags, trusted, lean);
    if (e) return e;
    const int ring = (int)((c->rgba_parity - 1) & 1u);  // the slot run_step just wrote
    const cudaStream_t s_d2h = lean ? c->s_main : c->s_d2h;
    if (!lean) CU_TRY(cudaStreamWaitEvent(c->s_d2h, c->rgba_written[ring], 0));
    uint64_t off = 0;
    uint32_t i = 0;
    while (i < n_pics) {
        const h263cu_pic& p = pics[i];
        const uint64_t dst = rgba_offsets ? rgba_offsets[i] : off;
        const size_t tight = (size_t)p.width * 4;
        const bool dense = tight == c->rgba_pitch && (size_t)p.height * c->rgba_pitch == c->rgba_slot;
        if (dense) {
            // coalesce runs of consecutive streams with contiguous destinations into one copy
            uint32_t j = i + 1;
            while (j < n_pics && pics[j].stream == pics[j - 1].stream + 1 && pics[j].width == p.width &&
                   pics[j].height == p.height &&
                   (!rgba_offsets || rgba_offsets[j] == rgba_offsets[j - 1] + c->rgba_slot))
                j++;
            CU_TRY(cudaMemcpyAsync(host_rgba + dst, c->rgba(p.stream, ring), (size_t)(j - i) * c->rgba_slot,
                                   cudaMemcpyDeviceToHost, s_d2h));
            off = dst + (uint64_t)(j - i) * c->rgba_slot;
            i = j;
        } else {
            CU_TRY(cudaMemcpy2DAsync(host_rgba + dst, tight, c->rgba(p.stream, ring), c->rgba_pitch, tight, p.height,
                                     cudaMemcpyDeviceToHost, s_d2h));
            off = dst + (uint64_t)tight * p.height;
            i++;
        }
    }
    CU_TRY(cudaEventRecord(c->rgba_read[ring], s_d2h));
    return 0;
}

int h263cu_submit_step_readback(h263cu_ctx* c, const h263cu_pic* pics, uint32_t n_pics, const h263cu_mb* mbs,
                                uint32_t n_mbs, const h263cu_event* events, uint32_t n_units, uint32_t out_flags,
                                uint8_t* host_rgba, const uint64_t* rgba_offsets) {
    return submit_readback(c, pics, n_pics, mbs, n_mbs, events, n_units, out_flags, host_rgba, rgba_offsets, false);
}

// The body of h263cu_decode_step; `positions` (may be NULL: input i sits at position i) says where input i's RGBA goes in
// host_rgba, in units of rgba_stride (the group form scatters the pictures of one device over a buffer shared by all).
static int decode_step_scattered(h263cu_ctx* c, h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                                 const uint32_t* stream_ids, uint32_t n, int threads, uint32_t out_flags, uint8_t* host_rgba,
                                 uint64_t rgba_stride, const uint32_t* positions, int* per_pic_err, uint32_t* n_decoded) {
    if (!c || !parsers || !packets || !lens) return H263CU_ERR_BAD_ARGUMENT;
    if (n_decoded) *n_decoded = 0;
    if (n == 0) return 0;
    const auto t_enter = std::chrono::steady_clock::now();
    cudaSetDevice(c->device);
    // capacities: a picture holds at most mbw * mbh macroblocks of this context; every event costs at least
    // 3 bits of bitstream and takes at most 2 units
    size_t bytes = 0;
    for (uint32_t i = 0; i < n; i++) bytes += lens[i];
    const size_t mb_need = (size_t)n * c->mbw * c->mbh, ev_need = bytes * 16 / 3 + 16 * (size_t)n;
    if (mb_need > 0xFFFFFFFFull || ev_need > 0xFFFFFFFFull) return H263CU_ERR_CAPACITY;
    // Stream ids are checked before anything is parsed; a picture larger than the context fails as that picture's
    // error inside the parse.  The parsers advance only when the device stage has accepted the step, so a failing
    // call leaves every stream as it was (decode_next_picture is a transaction, state.rs:120-137).
    c->decode_epoch++;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t sid = stream_ids ? stream_ids[i] : i;
        if (sid >= c->max_streams) return H263CU_ERR_CAPACITY;
        if (c->streams[sid].seen == c->decode_epoch) return H263CU_ERR_BAD_ARGUMENT;
        c->streams[sid].seen = c->decode_epoch;
    }
    const int slot = c->ring_pos;  // the ring slot submit_common is about to use
    h263cu_ctx::Staging& st = c->staging[slot];
    // the copy engine has finished with this staging set (it was read two submits ago)
    CU_TRY(cudaEventSynchronize(c->ring_h2d_done[slot]));
    int e;
    if ((e = grow_pinned(&st.pics, &st.pic_cap, n)) || (e = grow_pinned(&st.mbs, &st.mb_cap, mb_need)) ||
        (e = grow_pinned(&st.events, &st.ev_cap, ev_need)))
        return e;
    try {
        c->pic_of_input.resize(n);
    } catch (const std::bad_alloc&) {
        return H263CU_ERR_OUT_OF_MEMORY;
    }
    uint32_t np = 0, nm = 0, nu = 0;
    const auto t_parse0 = std::chrono::steady_clock::now();
    e = h263fe::parse_step_deferred(parsers, packets, lens, stream_ids, n, threads, st.pics, st.mbs, (uint32_t)st.mb_cap, st.events,
                                    (uint32_t)st.ev_cap, &np, &nm, &nu, per_pic_err, c->pic_of_input.data(), c->max_w, c->max_h);
    const auto t_parse1 = std::chrono::steady_clock::now();
    c->host_parse_s += std::chrono::duration<double>(t_parse1 - t_parse0).count();
    c->host_other_s += std::chrono::duration<double>(t_parse0 - t_enter).count();
    c->host_calls++;
    struct Tail {  // whatever follows the parse on any return path counts as "other"
        h263cu_ctx* c;
        std::chrono::steady_clock::time_point t;
        ~Tail() { c->host_other_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); }
    } tail{c, t_parse1};
    if (e) {
        h263fe::parse_step_finish(parsers, n, false);
        return e;
    }
    if (n_decoded) *n_decoded = np;
    if (np == 0) return 0;
    // one picture of at most a 4CIF's macroblocks: the single-stream case (H263State), latency bound on driver calls
    const bool lean = np == 1 && nm <= 1584;
    if (!host_rgba) {
        e = submit_common(c, st.pics, np, st.mbs, nm, st.events, nu, out_flags, true, lean);
    } else {
        try {
            c->rgba_offsets.resize(np);
        } catch (const std::bad_alloc&) {
            h263fe::parse_step_finish(parsers, n, false);
            return H263CU_ERR_OUT_OF_MEMORY;
        }
        for (uint32_t i = 0; i < n; i++)
            if (c->pic_of_input[i] >= 0) c->rgba_offsets[(size_t)c->pic_of_input[i]] = (uint64_t)(positions ? positions[i] : i) * rgba_stride;
        e = submit_readback(c, st.pics, np, st.mbs, nm, st.events, nu, out_flags, host_rgba, c->rgba_offsets.data(), true, lean);
    }
    // the device stage has the step (or refused it): only now do the parsers move on
    h263fe::parse_step_finish(parsers, n, e == 0);
    return e;
}

int h263cu_decode_step(h263cu_ctx* c, h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                       const uint32_t* stream_ids, uint32_t n, int threads, uint32_t out_flags, uint8_t* host_rgba,
                       uint64_t rgba_stride, int* per_pic_err, uint32_t* n_decoded) {
    return decode_step_scattered(c, parsers, packets, lens, stream_ids, n, threads, out_flags, host_rgba, rgba_stride, nullptr,
                                 per_pic_err, n_decoded);
}

int h263cu_sync(h263cu_ctx* c) {
    if (!c) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    CU_TRY(cudaStreamSynchronize(c->s_h2d));
    CU_TRY(cudaStreamSynchronize(c->s_main));
    CU_TRY(cudaStreamSynchronize(c->s_d2h));
    return 0;
}

int h263cu_stream_info(h263cu_ctx* c, uint32_t stream, uint32_t* width, uint32_t* height, uint32_t* pic_type,
                       uint32_t* pquant, uint32_t* temporal_reference) {
    if (!c || stream >= c->max_streams) return H263CU_ERR_BAD_ARGUMENT;
    const StreamState& st = c->streams[stream];
    if (!st.has_pic) return H263CU_ERR_NO_PICTURE;
    if (width) *width = st.w;
    if (height) *height = st.h;
    if (pic_type) *pic_type = st.pic_type;
    if (pquant) *pquant = st.pquant;
    if (temporal_reference) *temporal_reference = st.tr;
    return 0;
}

uint32_t h263cu_stream_dims(h263cu_ctx* c, uint32_t stream) {
    if (!c || stream >= c->max_streams || !c->streams[stream].has_pic) return 0u;
    return ((uint32_t)c->streams[stream].w << 16) | c->streams[stream].h;
}

int h263cu_read_yuv(h263cu_ctx* c, uint32_t stream, uint8_t* y, uint8_t* cb, uint8_t* cr) {
    if (!c || stream >= c->max_streams || !y || !cb || !cr) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    const StreamState& st = c->streams[stream];
    if (!st.has_pic) return H263CU_ERR_NO_PICTURE;
    CU_TRY(cudaStreamSynchronize(c->s_main));
    const size_t cw = (st.w + 1) / 2, ch = (st.h + 1) / 2;
    CU_TRY(cudaMemcpy2D(y, st.w, c->plane(0, stream, st.cur_slot), c->pitch_y, st.w, st.h, cudaMemcpyDeviceToHost));
    // the chroma planes live interleaved on the device: copy the CbCr rows and split them here
    std::vector<uint8_t> pairs;
    try {
        pairs.resize(cw * ch * CHROMA_STEP);
    } catch (const std::bad_alloc&) {
        return H263CU_ERR_OUT_OF_MEMORY;
    }
    CU_TRY(cudaMemcpy2D(pairs.data(), cw * CHROMA_STEP, c->plane(1, stream, st.cur_slot), c->pitch_c, cw * CHROMA_STEP, ch,
                        cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < cw * ch; i++) cb[i] = pairs[2 * i], cr[i] = pairs[2 * i + 1];
    return 0;
}

int h263cu_read_rgba(h263cu_ctx* c, uint32_t stream, uint8_t* rgba) {
    if (!c || stream >= c->max_streams || !rgba) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    const StreamState& st = c->streams[stream];
    if (!st.has_pic || st.rgba_slot < 0) return H263CU_ERR_NO_PICTURE;
    CU_TRY(cudaStreamSynchronize(c->s_main));
    CU_TRY(cudaMemcpy2D(rgba, (size_t)st.w * 4, c->rgba(stream, st.rgba_slot), c->rgba_pitch, (size_t)st.w * 4, st.h,
                        cudaMemcpyDeviceToHost));
    return 0;
}

int h263cu_checksums(h263cu_ctx* c, const uint32_t* streams, uint32_t n, uint64_t* out4) {
    if (!c || !streams || !out4) return H263CU_ERR_BAD_ARGUMENT;
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    std::vector<ChecksumJob> jobs;
    jobs.reserve((size_t)n * 4);
    for (uint32_t i = 0; i < n; i++) {
        if (streams[i] >= c->max_streams) return H263CU_ERR_BAD_ARGUMENT;
        const StreamState& st = c->streams[streams[i]];
        if (!st.has_pic) return H263CU_ERR_NO_PICTURE;
        const uint32_t cw = (st.w + 1u) / 2u, ch = (st.h + 1u) / 2u;
        jobs.push_back({c->plane(0, streams[i], st.cur_slot), st.w, st.h, c->pitch_y, i * 4 + 0, 1u, 0u});
        jobs.push_back({c->plane(1, streams[i], st.cur_slot), cw, ch, c->pitch_c, i * 4 + 1, (uint32_t)CHROMA_STEP, 0u});
        jobs.push_back({c->plane(2, streams[i], st.cur_slot), cw, ch, c->pitch_c, i * 4 + 2, (uint32_t)CHROMA_STEP, 0u});
        if (st.rgba_slot >= 0)
            jobs.push_back({c->rgba(streams[i], st.rgba_slot), (uint32_t)st.w * 4u, st.h, c->rgba_pitch, i * 4 + 3, 1u, 0u});
    }
    if (jobs.size() > c->jobs_cap) {
        if (c->d_jobs) cudaFree(c->d_jobs);
        if (c->d_sums) cudaFree(c->d_sums);
        c->d_jobs = nullptr, c->d_sums = nullptr;
        c->jobs_cap = 0;
        const size_t cap = (size_t)n * 4 + 64;
        CU_TRY(cudaMalloc((void**)&c->d_jobs, cap * sizeof(ChecksumJob)));
        CU_TRY(cudaMalloc((void**)&c->d_sums, cap * sizeof(unsigned long long)));
        c->jobs_cap = cap;
    }
    CU_TRY(cudaMemcpyAsync(c->d_jobs, jobs.data(), jobs.size() * sizeof(ChecksumJob), cudaMemcpyHostToDevice, c->s_main));
    CU_TRY(cudaMemsetAsync(c->d_sums, 0, (size_t)n * 4 * sizeof(unsigned long long), c->s_main));
    // grid.y is limited to 65535 jobs per launch
    for (size_t first = 0; first < jobs.size(); first += 60000) {
        const uint32_t cnt = (uint32_t)std::min<size_t>(60000, jobs.size() - first);
        launch_checksums(c->d_jobs + first, cnt, c->d_sums, c->s_main);
        c->launches++;
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out4, c->d_sums, (size_t)n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->s_main));
    CU_TRY(cudaStreamSynchronize(c->s_main));
    return 0;
}

int h263cu_timer_start(h263cu_ctx* c) {
    if (!c) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    CU_TRY(cudaEventRecord(c->t0, c->s_main));
    return 0;
}
int h263cu_timer_stop(h263cu_ctx* c, float* ms) {
    if (!c || !ms) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    CU_TRY(cudaEventRecord(c->t1, c->s_main));
    CU_TRY(cudaEventSynchronize(c->t1));
    CU_TRY(cudaEventElapsedTime(ms, c->t0, c->t1));
    return 0;
}
uint64_t h263cu_launch_count(h263cu_ctx* c) { return c ? c->launches : 0; }
uint64_t h263cu_tiled_launch_count(h263cu_ctx* c) { return c ? c->tiled_launches : 0; }

int h263cu_host_times(h263cu_ctx* c, double* parse_seconds, double* other_seconds, uint64_t* calls, int reset) {
    if (!c) return H263CU_ERR_BAD_ARGUMENT;
    if (parse_seconds) *parse_seconds = c->host_parse_s;
    if (other_seconds) *other_seconds = c->host_other_s;
    if (calls) *calls = c->host_calls;
    if (reset) c->host_parse_s = c->host_other_s = 0.0, c->host_calls = 0;
    return 0;
}

int h263cu_profile_enable(h263cu_ctx* c, int enable) {
    if (!c) return H263CU_ERR_BAD_ARGUMENT;
    c->profiling = enable != 0;
    return 0;
}
int h263cu_profile_read(h263cu_ctx* c, double* ms2, uint64_t* launches2) {
    if (!c || !ms2 || !launches2) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    CU_TRY(cudaStreamSynchronize(c->s_main));
    ms2[0] = ms2[1] = 0.0;
    launches2[0] = launches2[1] = 0;
    for (const auto& sp : c->prof_spans) {
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, sp.a, sp.b));
        ms2[sp.kind] += ms;
        launches2[sp.kind]++;
        c->prof_free.push_back(sp.a);
        c->prof_free.push_back(sp.b);
    }
    c->prof_spans.clear();
    return 0;
}

int h263cu_readback_wait(h263cu_ctx* c, uint32_t age) {
    if (!c || age > 1) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    if (c->rgba_parity <= age) return H263CU_ERR_NO_PICTURE;  // no such step yet
    const int ring = (int)((c->rgba_parity - 1u - age) & 1u);
    CU_TRY(cudaEventSynchronize(c->rgba_read[ring]));
    return 0;
}

// ---- resident steps as ONE CUDA graph ---------------------------------------------------------------------------
// A single stream's pictures are dependent launches of ~10 us each: issuing them one by one is bound by the launch
// path.  h263cu_graph_build captures the launches of n resident steps into one graph; the descriptors of all their
// pictures are computed once from a simulated walk of the per-stream state and live in their own device array.
struct h263cu_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    PicDev* d_pics = nullptr;
    uint32_t out_flags = 0;
    uint64_t launches = 0, tiled_launches = 0;
    uint32_t rgba_steps = 0;
    uint32_t rgba_parity0 = 0;  // parity of the RGBA ring the capture assumed
    std::vector<uint32_t> streams;        // streams the graph touches
    std::vector<StreamState> before, after;  // their state the capture assumed / leaves behind
};

h263cu_graph* h263cu_graph_build(h263cu_ctx* c, h263cu_step* const* steps, uint32_t n_steps, uint32_t out_flags, int* err) {
    int dummy;
    if (!err) err = &dummy;
    *err = 0;
    if (!c || !steps || n_steps == 0) {
        *err = H263CU_ERR_BAD_ARGUMENT;
        return nullptr;
    }
    cudaSetDevice(c->device);
    const bool want_rgba = (out_flags & H263CU_OUT_RGBA) != 0;
    const bool want_deblock = want_rgba && (out_flags & H263CU_OUT_DEBLOCK) != 0;
    h263cu_graph* g = new (std::nothrow) h263cu_graph();
    if (!g) {
        *err = H263CU_ERR_OUT_OF_MEMORY;
        return nullptr;
    }
    auto fail = [&](int code) {
        *err = code;
        h263cu_graph_free(c, g);
        return (h263cu_graph*)nullptr;
    };
    g->out_flags = out_flags;
    g->rgba_parity0 = c->rgba_parity & 1u;
    // ---- walk the steps over a copy of the stream state: the checks of run_step, the descriptors, the end state ----
    std::vector<StreamState> sim;
    std::vector<PicDev> picdev;
    struct StepPlan {
        size_t first_pic;
        uint32_t max_w, max_h;
        bool tiled, aligned16, wide_mv;
    };
    std::vector<StepPlan> plan;
    try {
        sim = c->streams;
        std::vector<uint8_t> touched(c->max_streams, 0);
        uint32_t parity = c->rgba_parity;
        for (uint32_t k = 0; k < n_steps; k++) {
            const h263cu_step* s = steps[k];
            if (!s || s->pics.empty()) return fail(H263CU_ERR_BAD_ARGUMENT);
            StepPlan sp{picdev.size(), 0, 0, true, true, false};
            const uint32_t stamp = ++c->stamp;
            for (const h263cu_pic& p : s->pics) {
                if (p.stream >= c->max_streams || p.width == 0 || p.height == 0 || p.width > c->max_w || p.height > c->max_h)
                    return fail(H263CU_ERR_CAPACITY);
                if ((uint32_t)p.mb_w * 16 < p.width || (uint32_t)p.mb_h * 16 < p.height || (uint32_t)p.mb_w > c->mbw ||
                    (uint32_t)p.mb_h > c->mbh || (uint64_t)p.first_mb + p.n_mbs > s->n_mbs ||
                    (uint64_t)p.first_event + p.n_event_units > s->n_units || p.n_mbs != (uint32_t)p.mb_w * p.mb_h)
                    return fail(H263CU_ERR_BAD_ARGUMENT);
                StreamState& st = sim[p.stream];
                if (st.stamp == stamp) return fail(H263CU_ERR_BAD_ARGUMENT);
                st.stamp = stamp;
                if (p.flags & H263CU_PICFLAG_HAS_INTER) {
                    if (!st.has_ref) return fail(H263CU_ERR_UNCODED_IFRAME_BLOCKS);
                    if (st.ref_w != p.width || st.ref_h != p.height) return fail(H263CU_ERR_REFERENCE_WOULD_ABORT);
                    if (!(p.flags & H263CU_PICFLAG_MV_IN_RANGE)) sp.wide_mv = true;
                    if (!st.padded) sp.tiled = false;
                }
                if ((p.width | p.height) & 15) sp.aligned16 = false;
                sp.max_w = std::max<uint32_t>(sp.max_w, p.width), sp.max_h = std::max<uint32_t>(sp.max_h, p.height);
                if (!touched[p.stream]) touched[p.stream] = 1, g->streams.push_back(p.stream);
            }
            if (c->force_kernel == 1) sp.tiled = false;
            const int ring = (int)(parity & 1u);
            for (const h263cu_pic& p : s->pics) {
                StreamState& st = sim[p.stream];
                picdev.push_back(make_picdev(c, p, st, want_rgba, ring));
                advance_stream(st, p, want_rgba, ring, sp.tiled);
            }
            if (want_rgba) parity++, g->rgba_steps++;
            g->launches += want_deblock ? 2 : 1;
            if (sp.tiled) g->tiled_launches++;
            plan.push_back(sp);
        }
        for (uint32_t sid : g->streams) g->before.push_back(c->streams[sid]), g->after.push_back(sim[sid]);
    } catch (const std::bad_alloc&) {
        return fail(H263CU_ERR_OUT_OF_MEMORY);
    }
    if (cudaMalloc((void**)&g->d_pics, picdev.size() * sizeof(PicDev)) != cudaSuccess) return fail(H263CU_ERR_OUT_OF_MEMORY);
    if (cudaMemcpy(g->d_pics, picdev.data(), picdev.size() * sizeof(PicDev), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(H263CU_ERR_CUDA);
    // ---- capture ----
    if (cudaStreamSynchronize(c->s_main) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    if (cudaStreamBeginCapture(c->s_main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail(H263CU_ERR_CUDA);
    const Pools pools{c->y_pool, c->c_pool, c->rgba_pool, c->pitch_y, c->pitch_c, c->rgba_pitch};
    for (uint32_t k = 0; k < n_steps; k++) {
        const h263cu_step* s = steps[k];
        const StepPlan& sp = plan[k];
        const PicDev* dp = g->d_pics + sp.first_pic;
        launch_recon(dp, s->d_mbs, s->d_events, s->n_mbs, want_rgba && !want_deblock, sp.tiled ? (sp.aligned16 ? 1 : 2) : 0, sp.wide_mv, pools,
                     &c->rgba_map, c->s_main);
        if (want_deblock) {
            if (sp.aligned16 && c->force_kernel != 1)
                launch_deblock_rgba_tile(dp, (uint32_t)s->pics.size(), sp.max_w, sp.max_h, c->s_main);
            else
                launch_deblock_rgba(dp, (uint32_t)s->pics.size(), sp.max_w, sp.max_h, c->s_main);
        }
    }
    if (cudaStreamEndCapture(c->s_main, &g->graph) != cudaSuccess || !g->graph) {
        cudaGetLastError();
        return fail(H263CU_ERR_CUDA);
    }
    if (cudaGraphInstantiate(&g->exec, g->graph, 0) != cudaSuccess) {
        cudaGetLastError();
        return fail(H263CU_ERR_CUDA);
    }
    return g;
}

int h263cu_graph_launch(h263cu_ctx* c, h263cu_graph* g) {
    if (!c || !g || !g->exec) return H263CU_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    // the graph holds plane addresses: it runs only from the state it was captured for
    if (g->rgba_steps && (c->rgba_parity & 1u) != g->rgba_parity0) return H263CU_ERR_BAD_ARGUMENT;
    for (size_t i = 0; i < g->streams.size(); i++) {
        const StreamState &a = c->streams[g->streams[i]], &b = g->before[i];
        if (a.has_pic != b.has_pic || (a.has_pic && a.cur_slot != b.cur_slot) || a.has_ref != b.has_ref ||
            (a.has_ref && (a.ref_slot != b.ref_slot || a.ref_w != b.ref_w || a.ref_h != b.ref_h || a.padded != b.padded)))
            return H263CU_ERR_BAD_ARGUMENT;
    }
    if (g->rgba_steps) {  // read-backs of earlier steps may still hold the RGBA ring
        CU_TRY(cudaStreamWaitEvent(c->s_main, c->rgba_read[0], 0));
        CU_TRY(cudaStreamWaitEvent(c->s_main, c->rgba_read[1], 0));
    }
    CU_TRY(cudaGraphLaunch(g->exec, c->s_main));
    for (size_t i = 0; i < g->streams.size(); i++) {
        StreamState& st = c->streams[g->streams[i]];
        const uint32_t seen = st.seen, stamp = st.stamp;
        st = g->after[i];
        st.seen = seen, st.stamp = stamp;
    }
    c->rgba_parity += g->rgba_steps;
    c->launches += g->launches;
    c->tiled_launches += g->tiled_launches;
    if (g->rgba_steps) {
        CU_TRY(cudaEventRecord(c->rgba_written[0], c->s_main));
        CU_TRY(cudaEventRecord(c->rgba_written[1], c->s_main));
    }
    return 0;
}

void h263cu_graph_free(h263cu_ctx* c, h263cu_graph* g) {
    if (!g) return;
    if (c) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->s_main);
    }
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    if (g->d_pics) cudaFree(g->d_pics);
    delete g;
}

// ---- one process, several GPUs: a group of contexts fed by ONE shared pool of parser threads ----
struct h263cu_group {
    std::vector<h263cu_ctx*> ctx;
    uint32_t streams_per_device = 0;
    int threads = 0;
    // per-device argument slices of the current step (reused)
    struct Slice {
        std::vector<h263cu_parser*> parsers;
        std::vector<const uint8_t*> packets;
        std::vector<size_t> lens;
        std::vector<uint32_t> ids, input, pos;
        std::vector<int> errs;
    };
    std::vector<Slice> slice;
};

h263cu_group* h263cu_group_create(const int* devices, uint32_t n_devices, uint32_t streams_per_device, uint32_t max_width,
                                  uint32_t max_height, int threads, int* err) {
    int dummy;
    if (!err) err = &dummy;
    *err = 0;
    if (!devices || n_devices == 0 || n_devices > 64 || streams_per_device == 0) {
        *err = H263CU_ERR_BAD_ARGUMENT;
        return nullptr;
    }
    h263cu_group* g = new (std::nothrow) h263cu_group();
    if (!g) {
        *err = H263CU_ERR_OUT_OF_MEMORY;
        return nullptr;
    }
    g->streams_per_device = streams_per_device, g->threads = threads;
    try {
        g->slice.resize(n_devices);
    } catch (const std::bad_alloc&) {
        delete g;
        *err = H263CU_ERR_OUT_OF_MEMORY;
        return nullptr;
    }
    for (uint32_t d = 0; d < n_devices; d++) {
        h263cu_ctx* c = h263cu_create(devices[d], streams_per_device, max_width, max_height, 0, err);
        if (!c) {
            h263cu_group_destroy(g);
            return nullptr;
        }
        g->ctx.push_back(c);
    }
    return g;
}

void h263cu_group_destroy(h263cu_group* g) {
    if (!g) return;
    for (h263cu_ctx* c : g->ctx) h263cu_destroy(c);
    delete g;
}

uint32_t h263cu_group_size(const h263cu_group* g) { return g ? (uint32_t)g->ctx.size() : 0u; }
h263cu_ctx* h263cu_group_ctx(h263cu_group* g, uint32_t index) { return g && index < g->ctx.size() ? g->ctx[index] : nullptr; }

int h263cu_group_decode_step(h263cu_group* g, h263cu_parser* const* parsers, const uint8_t* const* packets, const size_t* lens,
                             const uint32_t* stream_ids, uint32_t n, uint32_t out_flags, uint8_t* host_rgba, uint64_t rgba_stride,
                             int* per_pic_err, uint32_t* n_decoded) {
    if (!g || !parsers || !packets || !lens) return H263CU_ERR_BAD_ARGUMENT;
    if (n_decoded) *n_decoded = 0;
    const uint32_t nd = (uint32_t)g->ctx.size();
    try {
        for (auto& sl : g->slice) sl.parsers.clear(), sl.packets.clear(), sl.lens.clear(), sl.ids.clear(), sl.input.clear(), sl.pos.clear();
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t sid = stream_ids ? stream_ids[i] : i;
            const uint32_t d = sid % nd, local = sid / nd;
            if (local >= g->streams_per_device) return H263CU_ERR_CAPACITY;
            h263cu_group::Slice& sl = g->slice[d];
            sl.parsers.push_back(parsers[i]), sl.packets.push_back(packets[i]), sl.lens.push_back(lens[i]);
            sl.ids.push_back(local), sl.input.push_back(i), sl.pos.push_back(d * g->streams_per_device + local);
        }
        for (auto& sl : g->slice) sl.errs.assign(sl.ids.size(), 0);
    } catch (const std::bad_alloc&) {
        return H263CU_ERR_OUT_OF_MEMORY;
    }
    // Device by device: the pictures of device d are parsed on ALL of the group's parser threads and queued on d (upload,
    // kernels and read-back are asynchronous), then the threads move on to device d + 1 while d works and copies back.
    // No core is tied to a GPU: a device with fewer or cheaper pictures just takes less of the pool's time.
    int first_err = 0;
    uint32_t total = 0;
    for (uint32_t d = 0; d < nd; d++) {
        h263cu_group::Slice& sl = g->slice[d];
        const uint32_t m = (uint32_t)sl.ids.size();
        if (m == 0) continue;
        uint32_t got = 0;
        // RGBA of global stream s goes to host_rgba + ((s % n_devices) * streams_per_device + s / n_devices) * rgba_stride:
        // device-major, so that the pictures of one device are contiguous and leave in few large copies
        int e;
        if (host_rgba) {
            e = decode_step_scattered(g->ctx[d], sl.parsers.data(), sl.packets.data(), sl.lens.data(), sl.ids.data(), m, g->threads,
                                      out_flags, host_rgba, rgba_stride, sl.pos.data(), sl.errs.data(), &got);
        } else {
            e = h263cu_decode_step(g->ctx[d], sl.parsers.data(), sl.packets.data(), sl.lens.data(), sl.ids.data(), m, g->threads, out_flags,
                                   nullptr, 0, sl.errs.data(), &got);
        }
        if (e && !first_err) first_err = e;
        total += got;
        if (per_pic_err)
            for (uint32_t k = 0; k < m; k++) per_pic_err[sl.input[k]] = e ? e : sl.errs[k];
    }
    if (n_decoded) *n_decoded = total;
    return first_err;
}

int h263cu_group_sync(h263cu_group* g) {
    if (!g) return H263CU_ERR_BAD_ARGUMENT;
    int first = 0;
    for (h263cu_ctx* c : g->ctx) {
        const int e = h263cu_sync(c);
        if (e && !first) first = e;
    }
    return first;
}

// ---- stateless drop-ins: host buffers in, host buffers out ---------------------------------
// One scratch set per device (device buffer, pinned staging, stream), used under one lock: the calls run on the
// device that is current on the calling thread, like any CUDA runtime call.
namespace {
struct DropInScratch {
    uint8_t* dev = nullptr;
    size_t dev_cap = 0;
    uint8_t* pinned = nullptr;
    size_t pinned_cap = 0;
    cudaStream_t stream = nullptr;
};
std::mutex g_scratch_mutex;
DropInScratch g_scratch[64];

int scratch_get(size_t dev_bytes, size_t pinned_bytes, DropInScratch** out) {
    int device = 0;
    CU_TRY(cudaGetDevice(&device));
    if (device < 0 || device >= 64) return H263CU_ERR_NO_DEVICE;
    DropInScratch& sc = g_scratch[device];
    if (!sc.stream) CU_TRY(cudaStreamCreateWithFlags(&sc.stream, cudaStreamNonBlocking));
    if (dev_bytes > sc.dev_cap) {
        if (sc.dev) cudaFree(sc.dev);
        sc.dev = nullptr, sc.dev_cap = 0;
        CU_TRY(cudaMalloc((void**)&sc.dev, dev_bytes + dev_bytes / 4));
        sc.dev_cap = dev_bytes + dev_bytes / 4;
    }
    if (pinned_bytes > sc.pinned_cap) {
        if (sc.pinned) cudaFreeHost(sc.pinned);
        sc.pinned = nullptr, sc.pinned_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&sc.pinned, pinned_bytes + pinned_bytes / 4, cudaHostAllocDefault));
        sc.pinned_cap = pinned_bytes + pinned_bytes / 4;
    }
    *out = &sc;
    return 0;
}
}  // namespace

int h263cu_yuv420_to_rgba(const uint8_t* y, const uint8_t* chroma_b, const uint8_t* chroma_r, size_t y_len,
                          size_t y_width, uint8_t* rgba_out) {
    if (y_len == 0) return 0;  // bt601.rs:106-112
    if (!y || !chroma_b || !chroma_r || !rgba_out || y_width == 0 || y_len % y_width != 0 || y_width > 65535)
        return H263CU_ERR_BAD_ARGUMENT;
    if (h263cu_device_count() == 0) return H263CU_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    const size_t h = y_len / y_width, cw = (y_width + 1) / 2, ch = (h + 1) / 2, clen = cw * ch;
    const size_t o_cb = round_up(y_len, 256), o_cr = o_cb + round_up(clen, 256), o_out = o_cr + round_up(clen, 256);
    DropInScratch* sc;
    int e = scratch_get(o_out + y_len * 4, o_out + y_len * 4, &sc);
    if (e) return e;
    // pinned staging in both directions: one asynchronous copy each way on the drop-ins' own stream
    std::memcpy(sc->pinned, y, y_len);
    std::memcpy(sc->pinned + o_cb, chroma_b, clen);
    std::memcpy(sc->pinned + o_cr, chroma_r, clen);
    CU_TRY(cudaMemcpyAsync(sc->dev, sc->pinned, o_cr + clen, cudaMemcpyHostToDevice, sc->stream));
    launch_yuv420_to_rgba(sc->dev, sc->dev + o_cb, sc->dev + o_cr, (uint32_t)y_width, (uint32_t)h, sc->dev + o_out, sc->stream);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(sc->pinned + o_out, sc->dev + o_out, y_len * 4, cudaMemcpyDeviceToHost, sc->stream));
    CU_TRY(cudaStreamSynchronize(sc->stream));
    std::memcpy(rgba_out, sc->pinned + o_out, y_len * 4);
    return 0;
}

int h263cu_deblock(const uint8_t* data, size_t len, size_t width, uint8_t strength, uint8_t* out) {
    if (len == 0) return 0;
    if (!data || !out || width == 0 || len % width != 0 || width > 65535) return H263CU_ERR_BAD_ARGUMENT;
    if (h263cu_device_count() == 0) return H263CU_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    const size_t h = len / width;
    const size_t o_out = round_up(len, 256);
    DropInScratch* sc;
    int e = scratch_get(o_out + len, o_out + len, &sc);
    if (e) return e;
    std::memcpy(sc->pinned, data, len);
    CU_TRY(cudaMemcpyAsync(sc->dev, sc->pinned, len, cudaMemcpyHostToDevice, sc->stream));
    launch_deblock_plane(sc->dev, sc->dev + o_out, (uint32_t)width, (uint32_t)h, strength, sc->stream);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(sc->pinned + o_out, sc->dev + o_out, len, cudaMemcpyDeviceToHost, sc->stream));
    CU_TRY(cudaStreamSynchronize(sc->stream));
    std::memcpy(out, sc->pinned + o_out, len);
    return 0;
}

}  // extern "C"
