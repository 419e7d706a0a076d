// flv.cpp -- FLV container feed for the host front end.
//
// The reference decodes what its caller hands it: one H263Reader::from_source(&packet[..]) per FLV
// video tag (Sorenson Spark travels as FLV video codec 2, one picture per tag).  The demuxer is the
// caller's business (Ruffle), not part of h263-rs; this file is that piece for the batched decoder:
// a zero-copy scan of an FLV byte stream that lists the H.263 picture packets (offset, size,
// timestamp, frame type) so that they can be fed to h263cu_parse_step / h263cu_decode_step, and the
// matching muxer for the synthetic generator (the repo has no encoder, so test files are built here).
//
// FLV layout (Adobe "Video File Format Specification v10", section "The FLV File Format"):
//   header  'F' 'L' 'V' version flags(audio=4|video=1) data_offset(be32, 9)
//   body    PreviousTagSize0(be32) { tag PreviousTagSize(be32) }*
//   tag     type(u8: 8 audio, 9 video, 18 script; bit 5 = encrypted/filtered) data_size(be24)
//           timestamp(be24) timestamp_ext(u8, bits 31..24) stream_id(be24, 0) data[data_size]
//   video   data[0] = frame_type(4 bits: 1 key, 2 inter, 3 disposable inter, 5 info) | codec_id(4 bits: 2 = H.263)
#include <cstdint>
#include <cstring>

#include "../../include/h263cu.h"

namespace {
inline uint32_t be24(const uint8_t* p) { return ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2]; }
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put24(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 16), p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)v; }
inline void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24), p[1] = (uint8_t)(v >> 16), p[2] = (uint8_t)(v >> 8), p[3] = (uint8_t)v; }
}  // namespace

extern "C" {

int64_t h263cu_flv_scan(const uint8_t* data, size_t len, h263cu_flv_packet* out, size_t cap, uint32_t* n_other_tags) {
    if (n_other_tags) *n_other_tags = 0;
    if (!data || (!out && cap)) return H263CU_ERR_BAD_ARGUMENT;
    if (len < 9 || data[0] != 'F' || data[1] != 'L' || data[2] != 'V') return H263CU_ERR_INVALID_BITSTREAM;
    const uint32_t body = be32(data + 5);
    if (body < 9 || body > len) return H263CU_ERR_INVALID_BITSTREAM;
    size_t pos = body;
    int64_t n = 0;
    uint32_t other = 0;
    for (;;) {
        // PreviousTagSize + 11-byte tag header; a file cut short ends the scan without an error (streaming input)
        if (pos + 4 + 11 > len) break;
        const uint8_t* t = data + pos + 4;
        const uint32_t type = t[0] & 0x1Fu, filtered = t[0] & 0x20u;
        const uint32_t size = be24(t + 1);
        const uint32_t ts = be24(t + 4) | ((uint32_t)t[7] << 24);
        const size_t payload = pos + 4 + 11;
        if (payload + size > len) break;
        if (type == 9 && !filtered && size >= 1 && (data[payload] & 0x0F) == 2 && (data[payload] >> 4) != 5) {
            if ((size_t)n < cap) {
                h263cu_flv_packet& p = out[n];
                p.offset = payload + 1;
                p.size = size - 1;
                p.timestamp_ms = ts;
                p.frame_type = (uint8_t)(data[payload] >> 4);
                p.codec_id = 2;
                p.reserved = 0, p.reserved2 = 0;
            }
            n++;
        } else {
            other++;
        }
        pos = payload + size;
    }
    if (n_other_tags) *n_other_tags = other;
    return n;
}

int64_t h263cu_flv_mux(const uint8_t* packets, const uint64_t* pkt_off, const uint32_t* pkt_len, const uint8_t* frame_types,
                       uint32_t n, uint32_t ms_per_picture, uint32_t filler_every, uint8_t* out, size_t cap) {
    if ((!packets || !pkt_off || !pkt_len) && n) return H263CU_ERR_BAD_ARGUMENT;
    // size pass
    uint64_t need = 9 + 4;
    for (uint32_t i = 0; i < n; i++) {
        need += 11 + 1 + (uint64_t)pkt_len[i] + 4;
        if (filler_every && i % filler_every == 0) need += 11 + 2 + 4 + 11 + 3 + 4;  // one audio tag + one script tag
    }
    if (!out || cap < need) return (int64_t)need;
    uint8_t* p = out;
    std::memcpy(p, "FLV\x01", 4);
    p[4] = filler_every ? 5 : 1;
    put32(p + 5, 9);
    put32(p + 9, 0);
    p += 13;
    auto tag = [&](uint8_t type, uint32_t ts, const uint8_t* head, uint32_t nhead, const uint8_t* body, uint32_t nbody) {
        p[0] = type;
        put24(p + 1, nhead + nbody);
        put24(p + 4, ts & 0xFFFFFFu);
        p[7] = (uint8_t)(ts >> 24);
        put24(p + 8, 0);
        std::memcpy(p + 11, head, nhead);
        if (nbody) std::memcpy(p + 11 + nhead, body, nbody);
        put32(p + 11 + nhead + nbody, 11 + nhead + nbody);
        p += 11 + nhead + nbody + 4;
    };
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t ts = i * ms_per_picture;
        if (filler_every && i % filler_every == 0) {
            const uint8_t audio[2] = {0x2F, 0x00}, script[3] = {0x02, 0x00, 0x00};
            tag(8, ts, audio, 2, nullptr, 0);
            tag(18, ts, script, 3, nullptr, 0);
        }
        const uint8_t ft = frame_types ? frame_types[i] : (uint8_t)(i == 0 ? 1 : 2);
        const uint8_t head = (uint8_t)((ft << 4) | 2);
        tag(9, ts, &head, 1, packets + pkt_off[i], pkt_len[i]);
    }
    return (int64_t)(p - out);
}

}  // extern "C"
