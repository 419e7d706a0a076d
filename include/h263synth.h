/*
 * h263synth.h -- C ABI of libh263synth.so: the synthetic Sorenson-flavour bitstream generator.
 *
 * The reference repository has no encoder and ships no sample streams (BASELINE.json north_star: "a small synthetic
 * Sorenson-flavour bitstream generator is written first").  The generator is test and benchmark tooling: it is built
 * into its own shared library so that neither the product library (libh263cu.so) carries it nor the CPU reference arm
 * of bench.py has to map any product code to get its input streams.  It speaks the bitstream syntax the reference
 * parses (h263/src/parser/picture.rs:271-327,611-660, macroblock.rs:23-549,
 * block.rs:39-755) with the same code tables the product's parser uses (csrc/vlc_codes.inc).
 */
#ifndef H263SYNTH_H
#define H263SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h263cu_synth_params {
    uint32_t width, height;
    uint32_t n_pictures;
    uint64_t seed;
    uint32_t flavour;      /* 0 = Sorenson Spark, 1 = baseline H.263 (standard sizes only) */
    uint32_t version;      /* Sorenson version field: 0 or 1 (1 = 7/11-bit escapes) */
    uint32_t intra_period; /* an I picture every N pictures; 0 = only the first */
    uint32_t deblock_flag; /* value of the Sorenson DeblockingFlag */
    uint32_t qp_min, qp_max;
    uint32_t pct_uncoded;  /* P pictures: % of MBs with COD=1 */
    uint32_t pct_intra;    /* P pictures: % of MBs coded INTRA */
    uint32_t pct_fourmv;   /* P pictures: % of MBs coded INTER4V */
    uint32_t pct_dquant;   /* % of coded MBs carrying DQUANT */
    uint32_t pct_cbp_inter; /* % of blocks of an inter MB that carry coefficients */
    uint32_t pct_cbp_intra; /* % of blocks of an intra MB that carry AC coefficients */
    uint32_t mean_events_x10; /* mean TCOEF events per coded block, times 10 */
    uint32_t pct_escape;   /* % of events forced through the ESCAPE path */
    uint32_t permille_overflow; /* per-mille of coded blocks whose runs overflow the zig-zag */
    uint32_t mv_mode;      /* 0 small vectors, 1 full range uniform, 2 biased across borders */
    uint32_t truncate_permille; /* per-mille of P pictures that end early (padding path) */
    uint32_t pct_disposable;  /* Sorenson only: % of P pictures sent as disposable P pictures (type code 2) */
    uint32_t reserved[3];
} h263cu_synth_params;

void h263cu_synth_default_params(h263cu_synth_params* p, uint32_t width, uint32_t height, uint32_t n_pictures,
                                 uint64_t seed);
/* Writes the stream's packets back to back into out (one byte-aligned, zero-padded packet
 * per picture) and their offsets/lengths; returns the number of bytes needed (call with
 * cap = 0 to size the buffer) or a negative error. */
int64_t h263cu_synth_stream(const h263cu_synth_params* p, uint8_t* out, size_t cap, uint64_t* pkt_off,
                            uint32_t* pkt_len);

#ifdef __cplusplus
}
#endif
#endif /* H263SYNTH_H */
