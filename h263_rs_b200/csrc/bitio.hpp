// bitio.hpp -- MSB-first bit reader / writer and table-driven VLC decode for the host
// front end.  Replaces the reference's byte-at-a-time VecDeque reader and one-bit-per-step
// tree walk (h263/src/parser/reader.rs:49-58, 272-290) with a 64-bit window and one
// table lookup per code; observable behaviour (values, EOF conditions) is identical.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace h263fe {

struct VlcCode {
    const char* bits;
    int len;
    int kind;
    int a;
    int b;
    int c;
};

struct VlcEntry {  // 4 bytes: one aligned load per symbol
    uint8_t len_kind;  // bits 4..0 = code length in bits, bits 6..5 = kind (0 valid, 1 stuffing, 2 invalid, 3 escape)
    int8_t a;
    uint8_t b;
    uint8_t c;
    unsigned len() const { return len_kind & 31u; }
    unsigned kind() const { return len_kind >> 5; }
};

struct VlcTable {
    int max_len = 0;
    std::vector<VlcEntry> lut;  // 2^max_len entries
    void build(const VlcCode* codes, int n);
};

enum { T_MCBPC_I = 0, T_MCBPC_P = 1, T_CBPY = 2, T_MVD = 3, T_TCOEF = 4 };
const VlcTable& vlc_table(int id);
const VlcCode* vlc_codes(int id, int* count);

// Reads MSB-first from an in-memory packet.  A read of n bits fails (returns false)
// iff fewer than n bits remain -- the reference's UnexpectedEof (reader.rs:49-75).
// The next bits sit MSB-aligned in a 64-bit window that is refilled from memory only when fewer than 32
// of them are left, so the dependent chain per symbol is shift -> table load -> shift instead of a fresh
// 8-byte load + byte swap per read.
struct BitReader {
    const uint8_t* data;
    size_t total_bits;

    BitReader(const uint8_t* d, size_t len) : data(d), total_bits(len * 8), pos_(0) { refill(); }
    size_t pos() const { return pos_; }
    size_t avail() const { return total_bits - pos_; }
    void seek(size_t p) {
        pos_ = p;
        refill();
    }

    // Next n (<= 32) bits, zero padded past the end of the packet.
    inline uint32_t peek_padded(unsigned n) const { return n == 0 ? 0u : (uint32_t)(win_ >> (64 - n)); }
    inline void consume(unsigned n) {
        pos_ += n;
        if (n >= have_) {
            refill();
        } else {
            win_ <<= n;
            have_ -= n;
            if (have_ < 32) refill();
        }
    }
    inline bool read(unsigned n, uint32_t* out) {
        if (n > avail()) return false;
        *out = peek_padded(n);
        consume(n);
        return true;
    }
    inline bool read_signed(unsigned n, int32_t* out) {
        uint32_t v;
        if (!read(n, &v)) return false;
        *out = (int32_t)(v << (32 - n)) >> (32 - n);
        return true;
    }
    inline bool skip(unsigned n) {
        if (n > avail()) return false;
        consume(n);
        return true;
    }
    // One VLC symbol.  Returns false on EOF (the serial walk would run out of bits before
    // reaching a leaf: codes are prefix free, so that happens iff the zero-padded lookup
    // lands on a code longer than what is left).
    inline bool read_vlc(const VlcTable& t, const VlcEntry** out) {
        const VlcEntry& e = t.lut[peek_padded((unsigned)t.max_len)];
        if (e.len() > avail()) return false;
        consume(e.len());
        *out = &e;
        return true;
    }
    // One VLC symbol followed by `extra` (<= 8) plain bits, fetched from the same window (max_len + extra <= 32).
    // EOF semantics as two separate reads: the code must fit, then the extra bits must fit.
    inline bool read_vlc_bits(const VlcTable& t, unsigned extra, const VlcEntry** out, uint32_t* bits, bool* eof_in_extra) {
        const uint64_t w = win_;
        const VlcEntry& e = t.lut[(uint32_t)(w >> (64 - t.max_len))];
        const unsigned len = e.len();
        *eof_in_extra = false;
        if (len > avail()) return false;
        *out = &e;
        if (e.kind() != 0) {  // stuffing / invalid / escape: the caller decides what follows
            consume(len);
            *bits = 0;
            return true;
        }
        if (len + extra > avail()) {
            consume(len);
            *eof_in_extra = true;
            return false;
        }
        *bits = (uint32_t)((w << len) >> (64 - extra));
        consume(len + extra);
        return true;
    }

  private:
    size_t pos_;
    uint64_t win_;   // bits pos_.. MSB-aligned, zero padded past the end of the packet
    unsigned have_;  // how many leading bits of win_ are backed by the packet or its zero padding (>= 32 after a refill)

    inline void refill() {
        const size_t byte = pos_ >> 3, nbytes = total_bits >> 3;
        uint64_t w = 0;
        if (byte + 8 <= nbytes) {
            uint64_t raw;
            std::memcpy(&raw, data + byte, 8);
            w = __builtin_bswap64(raw);
        } else {
            for (size_t i = 0; i < 8; i++) w = (w << 8) | (byte + i < nbytes ? data[byte + i] : 0);
        }
        win_ = w << (pos_ & 7);
        have_ = 64 - (unsigned)(pos_ & 7);  // >= 57
    }
};

struct BitWriter {
    std::vector<uint8_t> bytes;
    uint64_t acc = 0;
    int nacc = 0;
    void put(uint32_t value, int nbits) {
        for (int i = nbits - 1; i >= 0; i--) {
            acc = (acc << 1) | ((value >> i) & 1u);
            if (++nacc == 8) {
                bytes.push_back((uint8_t)acc);
                acc = 0;
                nacc = 0;
            }
        }
    }
    void put_code(const char* bits) {
        for (const char* p = bits; *p; p++) put((uint32_t)(*p - '0'), 1);
    }
    void align_zero() {
        while (nacc != 0) put(0, 1);
    }
};

}  // namespace h263fe
