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

#define H263_AI __attribute__((always_inline)) inline

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
struct BitReader {
    const uint8_t* data;
    size_t total_bits;

    BitReader(const uint8_t* d, size_t len) : data(d), total_bits(len * 8), ptr_(d), end_(d + len), buf_(0), cnt_(0), left_(len * 8) {
        refill();
    }
    H263_AI size_t pos() const { return total_bits - left_; }
    H263_AI size_t avail() const { return left_; }  // kept as a counter: one subtraction per consume
    H263_AI void seek(size_t p) {
        ptr_ = data + (p >> 3);
        buf_ = 0, cnt_ = 0;
        refill();
        const unsigned r = (unsigned)(p & 7);  // p <= total_bits: the byte that holds bit p exists whenever r != 0
        buf_ <<= r;
        cnt_ -= r;
        refill();
        left_ = total_bits - p;
    }

    // Next n (<= 32) bits, zero padded past the end of the packet.
    H263_AI uint32_t peek_padded(unsigned n) const { return n == 0 ? 0u : (uint32_t)(buf_ >> (64 - n)); }
    // The buffer is topped up after every consume without a data-dependent branch (the only branch is the
    // end-of-packet test, taken in the last 8 bytes).  The load address of a refill depends on the PREVIOUS symbol's
    // length only, so the dependent chain per symbol is shift -> table load -> shift -> or: the 8-byte load and the
    // byte swap run beside it.  At least 56 bits are valid after every call (fewer only in the packet's last bytes,
    // where the missing ones read as zeros), so any field or code + sign (<= 32 bits) can be taken from it.
    H263_AI void consume(unsigned n) {
        buf_ <<= n;
        cnt_ -= n;
        left_ -= n;
        refill();
    }
    H263_AI bool read(unsigned n, uint32_t* out) {
        if (n > avail()) return false;
        *out = peek_padded(n);
        consume(n);
        return true;
    }
    H263_AI bool read_signed(unsigned n, int32_t* out) {
        uint32_t v;
        if (!read(n, &v)) return false;
        *out = (int32_t)(v << (32 - n)) >> (32 - n);
        return true;
    }
    H263_AI bool skip(unsigned n) {
        if (n > avail()) return false;
        consume(n);
        return true;
    }
    // One VLC symbol.  Returns false on EOF (the serial walk would run out of bits before
    // reaching a leaf: codes are prefix free, so that happens iff the zero-padded lookup
    // lands on a code longer than what is left).
    H263_AI bool read_vlc(const VlcTable& t, const VlcEntry** out) {
        const VlcEntry& e = t.lut[peek_padded((unsigned)t.max_len)];
        if (e.len() > avail()) return false;
        consume(e.len());
        *out = &e;
        return true;
    }
    // The same with the table held by the caller as a raw pointer + width (locals of the caller: no reload of the
    // vector's fields after every store the compiler cannot disambiguate)
    H263_AI bool read_vlc(const VlcEntry* __restrict lut, unsigned max_len, const VlcEntry** out) {
        const VlcEntry& e = lut[(uint32_t)(buf_ >> (64 - max_len))];
        if (e.len() > avail()) return false;
        consume(e.len());
        *out = &e;
        return true;
    }
    // One VLC symbol followed by `extra` (<= 8) plain bits, fetched from the same window (max_len + extra <= 32).
    // EOF semantics as two separate reads: the code must fit, then the extra bits must fit.
    H263_AI bool read_vlc_bits(const VlcTable& t, unsigned extra, const VlcEntry** out, uint32_t* bits, bool* eof_in_extra) {
        const uint64_t w = buf_;
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
    // the window itself, for decoders that look at more than one field per fetch (frontend.cpp's TCOEF loop)
    H263_AI uint64_t window() const { return buf_; }

  private:
    const uint8_t* ptr_;  // next byte to fetch
    const uint8_t* end_;
    uint64_t buf_;   // the next bits, MSB-aligned; bits below the top cnt_ are either upcoming data or zero
    unsigned cnt_;   // how many of the top bits are accounted for (fetched and not yet consumed)
    size_t left_;    // bits of the packet not yet consumed

    H263_AI void refill() {
        if (__builtin_expect(ptr_ + 8 <= end_, 1)) {
            uint64_t raw;
            std::memcpy(&raw, ptr_, 8);
            buf_ |= __builtin_bswap64(raw) >> cnt_;
            ptr_ += (63 - cnt_) >> 3;
            cnt_ |= 56;
        } else {
            // the last bytes of the packet: byte by byte, then zeros
            buf_ &= cnt_ ? ~0ull << (64 - cnt_) : 0ull;  // drop the look-ahead bits a fast refill may have left below cnt_
            while (cnt_ <= 56 && ptr_ < end_) {
                buf_ |= (uint64_t)*ptr_++ << (56 - cnt_);
                cnt_ += 8;
            }
        }
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
