/*
 * huff_oracle.h — CPU restatement of yechuan51/huffman's GPU compressor path and
 * its decompressor.  TEST INFRASTRUCTURE ONLY: nothing under huffman_b200/ may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity pin status: the reference ships no golden vectors (its only test is a
 * romeo.txt round trip, Makefile:17-29).  This oracle is pinned by
 *   (i)  the unmodified reference `extract` (Decompressor.cu) built into
 *        oracle/_ref/ decoding the oracle's files bit-exactly,
 *   (ii) the unmodified reference GPU `archive` (Compressor.cu +
 *        gpuHuffmanConstruction.h, built for sm_100a into oracle/_ref/) run on a
 *        B200; its outputs are committed as sha256 fixtures in tests/golden/.
 *
 * All file:line citations are into /root/reference.
 *   C: Compressor.cu   D: Decompressor.cu   h: gpuHuffmanConstruction.h
 */
#ifndef HUFF_ORACLE_H
#define HUFF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HO_NSYM 65536

typedef struct {
    uint32_t U;                 /* non-zero bins (C:378-385) */
    uint16_t order[HO_NSYM];    /* rank -> symbol, ascending (freq, symbol) (C:387-393, C:419-425) */
    uint8_t  len[HO_NSYM];      /* by symbol; 0 when absent */
    uint64_t code[HO_NSYM];     /* by symbol; root->leaf path, right aligned */
    uint32_t maxlen;
    uint64_t table_bits;        /* sum over ranks of 16+8+len (C:454-483) */
    uint64_t payload_bits;      /* sum hist[s]*len[s] (C:541-558) */
} ho_codebook_t;

enum {
    HO_OK = 0,
    HO_ERR_CODE_TOO_LONG = 1,   /* a path longer than 64 bits */
    HO_ERR_CAPACITY = 2,
    HO_ERR_FORMAT = 3
};

/* C:38-48: bins of little-endian byte pairs; trailing odd byte is not counted. */
void ho_histogram(const uint8_t *in, uint64_t n, uint64_t *hist);

/* C:378-425 + h:353-494 + h:551-579 through the two-queue equivalence. */
int ho_codebook(const uint64_t *hist, ho_codebook_t *cb);

/* Independent restatement of the reference's ROUND mechanism (h:137-151,
 * h:163-209, h:353-466): pivot search with its clamp, pair (2i,2i+1), merge with
 * carried-over nodes first on ties.  Produces the same fields as ho_codebook;
 * tests assert both agree. */
int ho_codebook_rounds(const uint64_t *hist, ho_codebook_t *cb);

/* exact size of the .compressed file for this input */
uint64_t ho_compressed_size(const ho_codebook_t *cb, uint64_t n);

/* C:427-487, C:541-601, C:637-669: whole file image ("ideal stream"). */
int ho_compress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n);

/* D:68-108, D:129-182, D:243-291 */
int ho_decompressed_size(const uint8_t *in, uint64_t n, uint64_t *out_n);
int ho_decompress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n);

/* SURVEY 2.3 clean-domain predicate: 1 when the reference GPU binary's output is
 * the ideal stream for this input (defects R1/R2 do not trigger). */
int ho_reference_clean(const uint8_t *in, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
