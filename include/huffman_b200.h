/*
 * huffman_b200.h — C ABI of libhuffb200.so, the B200 (sm_100a) Huffman codec that
 * is a drop-in for yechuan51/huffman's compress / decompress path.
 *
 * The reference exposes no library API: its boundary is two programs
 * (`archive <file>` -> `<file>.compressed`, Compressor.cu:315-321, :427-429;
 * `extract <file>` -> `DECOMPRESSED_FILE`, Decompressor.cu:47-63, :104-105) and the
 * on-disk format.  Each entry point below names the reference
 * code whose RESULT it reproduces.  Citations: C: Compressor.cu,
 * h: gpuHuffmanConstruction.h, D: Decompressor.cu (all under /root/reference).
 *
 * Conventions
 *   - every function returns 0 (HF_OK) or an HF_ERR_* code; nothing aborts or
 *     throws across the ABI (the reference abort()s, h:24-32);
 *   - `d_` pointers are device memory on the context's GPU, `h_` pointers host;
 *   - stage functions only enqueue work on the context's stream and return;
 *     functions documented "synchronises" wait for the stream;
 *   - one context per host thread and per GPU; a context owns its workspace;
 *   - there is NO CPU fallback: every call fails with HF_ERR_CUDA when no GPU
 *     is usable.
 */
#ifndef HUFFMAN_B200_H
#define HUFFMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF_NSYM 65536u            /* 16-bit little-endian byte-pair alphabet (C:38-48, C:323) */
#define HF_MAX_CODE_BITS 64u

enum {
    HF_OK = 0,
    HF_ERR_CUDA = 1,              /* CUDA runtime error; see hf_last_error */
    HF_ERR_ARG = 2,
    HF_ERR_CAPACITY = 3,          /* output buffer too small */
    HF_ERR_FORMAT = 4,            /* malformed .compressed image */
    HF_ERR_CODE_TOO_LONG = 5,     /* a tree path longer than 64 bits */
    HF_ERR_INTERNAL = 6,          /* device-side watchdog / consistency failure */
    HF_ERR_IO = 7
};

typedef struct hf_ctx hf_ctx;

/* what hf_codebook_info returns; the same numbers the reference prints or derives */
typedef struct {
    uint32_t n_unique;            /* "Unique symbols count" (C:378-385) */
    uint32_t max_code_bits;       /* maxCL (h:462-464) */
    uint64_t table_bits;          /* header table: sum of 16+8+len (C:454-483) */
    uint64_t payload_bits;        /* sum hist[s]*len[s] (C:548-558 minus bitCounter) */
    uint32_t status;              /* HF_OK or HF_ERR_CODE_TOO_LONG */
    uint32_t reserved;
} hf_cb_info_t;

typedef struct {
    uint32_t n_unique;            /* D:69-71 */
    uint32_t is_odd;              /* D:76 */
    uint32_t last_byte;           /* D:77-80 */
    uint32_t max_code_bits;
    uint64_t original_bytes;      /* D:243-255 */
    uint64_t payload_start_bit;   /* bit offset of the first payload bit from byte 0 of the file */
    uint32_t status;              /* HF_OK or HF_ERR_FORMAT */
    uint32_t reserved;
} hf_header_info_t;

/* ---- lifecycle -------------------------------------------------------- */

/* device: CUDA ordinal.  stream: the cudaStream_t to enqueue on (e.g. torch's current
 * stream); NULL is the legacy default stream, which is what the reference runs on. */
int hf_ctx_create(hf_ctx **ctx, int device, void *stream);
int hf_ctx_destroy(hf_ctx *ctx);
int hf_ctx_set_stream(hf_ctx *ctx, void *stream);
int hf_sync(hf_ctx *ctx);                         /* synchronises */
const char *hf_last_error(hf_ctx *ctx);           /* static storage owned by ctx */
const char *hf_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t hf_launch_count(hf_ctx *ctx);

/* Per-kernel device times, for bench.py's roofline: while enabled, every kernel launch of this
 * context is bracketed by CUDA events on the context's stream (the reference prints wall
 * timers instead, C:356-399, C:492-593, h:697-698).  hf_profile_read synchronises, sums the
 * recorded launches by kernel name into out[0..*n_out) and clears the record. */
typedef struct {
    char name[48];
    uint32_t launches;
    float total_ms;
} hf_kernel_time_t;
int hf_profile_enable(hf_ctx *ctx, int on);
int hf_profile_read(hf_ctx *ctx, hf_kernel_time_t *out, uint32_t cap, uint32_t *n_out);

/* development aid (phase-timing builds): synchronises and copies a piece of the context's
 * device workspace to the host */
int hf_debug_read_ws(hf_ctx *ctx, uint64_t off, void *h_dst, uint64_t bytes);

/* pinned host buffers for the host-facing calls (C:343 uses cudaHostAlloc) */
int hf_host_alloc(void **h_ptr, size_t bytes);
int hf_host_free(void *h_ptr);

/* ---- sizes ------------------------------------------------------------ */

size_t hf_codebook_bytes(void);                   /* device bytes of one codebook object */
size_t hf_decode_table_bytes(void);               /* device bytes of one decode-table object */
uint64_t hf_compress_bound(uint64_t n_bytes);     /* safe capacity for a .compressed image */

/* ---- compress stages (device pointers, asynchronous) ------------------ */

/* C:38-48 calculateFrequency.  Adds the counts of the little-endian byte pairs of
 * d_in[0 .. n_bytes & ~1) into d_hist[65536] (u64, caller zeroes; sums of several
 * shards may accumulate in place).  d_in must be 2-byte aligned. */
int hf_histogram(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, uint64_t *d_hist);

/* C:378-425 (count, stable sort, non-zero suffix) + h:353-494, h:551-579 (tree, code
 * lengths, code words, bit polarity).  Deterministic in d_hist alone, so every rank
 * of a sharded job builds the identical codebook from the all-reduced histogram. */
int hf_build_codebook(hf_ctx *ctx, const uint64_t *d_hist, void *d_codebook);

/* synchronises; copies the codebook summary to the host */
int hf_codebook_info(hf_ctx *ctx, const void *d_codebook, hf_cb_info_t *h_info);

/* test / inspection hook: copies order[65536] (u16, rank -> symbol), len[65536] (u8, by
 * symbol) and code[65536] (u64, by symbol, right aligned) to host arrays; synchronises */
int hf_codebook_export(hf_ctx *ctx, const void *d_codebook, uint16_t *h_order,
                       uint8_t *h_len, uint64_t *h_code);

/* dot(hist, len) for a shard's own histogram: the shard's payload bit count that the
 * ranks all-gather to obtain their global start bits.  *d_bits device u64. */
int hf_shard_payload_bits(hf_ctx *ctx, const uint64_t *d_shard_hist, const void *d_codebook,
                          uint64_t *d_bits);

/* C:427-487, C:637-669: bytes 0..2(3) and the MSB-first header bit stream (U entries of
 * symbol16,len8,code, then the 64-bit size) into d_file.  The last, partial header byte
 * is written zero-padded; hf_encode merges the first payload bits into it. */
int hf_header_pack(hf_ctx *ctx, const void *d_codebook, uint64_t n_bytes, uint32_t last_byte,
                   uint8_t *d_file, uint64_t capacity);

/* C:50-74 + C:541-588 + C:597-601 (populateCWLength, scan, encodeFromCW, tail flush):
 * bits per 512-symbol unit, a two-level scan, one warp-independent packing kernel.  Writes the code words of the byte pairs of
 * d_in[0 .. n_bytes & ~1) as one MSB-first bit stream starting `start_bit` bits after
 * d_stream.  Bits of the first byte before the start phase are preserved (they belong to
 * the header or to the previous shard); the final partial byte is zero padded. */
int hf_encode(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, const void *d_codebook,
              uint8_t *d_stream, uint64_t start_bit);

/* whole `archive` data path on device buffers: histogram, codebook, header, encode.
 * Synchronises once (to learn the size); *h_file_bytes receives the image size. */
int hf_compress(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, uint8_t *d_file,
                uint64_t capacity, uint64_t *h_file_bytes);

/* ---- decompress stages ------------------------------------------------ */

/* D:68-103, D:129-182, D:243-255: parses the header of a .compressed image in device
 * memory and builds the decode tables.  Synchronises; fills *h_info. */
int hf_parse_header(hf_ctx *ctx, const uint8_t *d_file, uint64_t file_bytes,
                    void *d_decode_table, hf_header_info_t *h_info);

/* decode tables straight from a codebook (for shards / tests that skip the file image) */
int hf_decode_table_from_codebook(hf_ctx *ctx, const void *d_codebook, void *d_decode_table);

/* D:259-284 translateFile: decodes n_symbols code words starting `start_bit` bits after
 * d_stream (the stream holds stream_bytes bytes) into d_out as little-endian pairs.
 * Self-synchronising: needs no side index. */
int hf_decode(hf_ctx *ctx, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit,
              uint64_t n_symbols, const void *d_decode_table, uint8_t *d_out);

/* One rank's BYTE RANGE of a stream sharded over several GPUs (the reference is
 * single-GPU and its format has no offset index, D:259-284).  d_range points at the first byte of
 * the range; halo_bytes (>= 16) after it are readable: the following bytes of the same stream,
 * zeros past its end.
 *   hf_range_overflow: where does the first code word AFTER the range start?  Found speculatively
 *     by self-synchronising over the last 256 KiB of the range from a guessed start; d_result[1] =
 *     bits past the range end.  The ranks all-gather these values: each is the next rank's first bit.
 *   hf_decode_range: decodes the code words that start inside the range, the first one first_bit
 *     bits into it, into d_out (at most out_symbols).  d_result (device u64[4]): [1] overflow of
 *     the last code word past the range end (must equal what hf_range_overflow predicted),
 *     [2] symbols decoded, [3] flags: 4 invalid code, 8 out_symbols too small.  Any flag means d_out must not
 *     be used.  Both calls are asynchronous. */
int hf_range_overflow(hf_ctx *ctx, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes,
                      const void *d_decode_table, uint64_t *d_result);
int hf_decode_range(hf_ctx *ctx, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes,
                    uint64_t first_bit, const void *d_decode_table, uint8_t *d_out, uint64_t out_symbols,
                    uint64_t *d_result);

/* Kept for ABI compatibility: the decoder has ONE mode, exact self-synchronising decode (an earlier revision also had a
 * speculative single-pass decoder that this switch turned off).  The argument is stored and ignored. */
int hf_set_decode_mode(hf_ctx *ctx, int exact_only);

/* whole `extract` data path on device buffers.  Synchronises. */
int hf_decompress(hf_ctx *ctx, const uint8_t *d_file, uint64_t file_bytes, uint8_t *d_out,
                  uint64_t capacity, uint64_t *h_out_bytes);

/* ---- optional side index -------------------------- */

/* The reference format has no offset array (Decompressor.cu:259-284 decodes serially), so hf_decompress finds the code
 * word boundaries by self-synchronisation.  For an image THIS library writes, the compressor can also produce a side
 * index — one 16-bit record per 256 bits of payload (first code word boundary, code words starting there), 1/16 of the
 * image — with which the decoder skips the synchronisation pass.  The image itself is unchanged (byte-identical to
 * the reference's); the index is never required: hf_decompress_indexed checks it against the image and against
 * every walk it drives, and decodes without it when it is absent, stale, made for another alignment of the image
 * (address modulo 16) or wrong.  d_index must be 16-byte aligned. */
uint64_t hf_index_bound(uint64_t n_bytes);        /* safe capacity for the index of an n_bytes input */
int hf_compress_indexed(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, uint8_t *d_file, uint64_t capacity,
                        uint64_t *h_file_bytes, uint8_t *d_index, uint64_t index_capacity, uint64_t *h_index_bytes);
int hf_decompress_indexed(hf_ctx *ctx, const uint8_t *d_file, uint64_t file_bytes, const uint8_t *d_index,
                          uint64_t index_bytes, uint8_t *d_out, uint64_t capacity, uint64_t *h_out_bytes);

/* ---- host-buffer calls (what the CLIs and the end-to-end benchmark use) - */

/* h_in / h_file should be pinned (hf_host_alloc) for full PCIe speed; pageable works.
 * Copies are chunked and overlapped with the histogram / encode kernels. */
int hf_compress_host(hf_ctx *ctx, const uint8_t *h_in, uint64_t n_bytes, uint8_t *h_file,
                     uint64_t capacity, uint64_t *h_file_bytes);
int hf_decompressed_size_host(const uint8_t *h_file, uint64_t file_bytes, uint64_t *h_out_bytes);
int hf_decompress_host(hf_ctx *ctx, const uint8_t *h_file, uint64_t file_bytes, uint8_t *h_out,
                       uint64_t capacity, uint64_t *h_out_bytes);

/* ---- program-level entry points (the reference's two mains) ----------- */

/* C:315-632: reads `path`, writes `path`.compressed, prints the reference's progress
 * lines to stdout.  Returns HF_OK also when the file does not exist (C:325-330). */
int hf_archive_file(hf_ctx *ctx, const char *path);
/* D:47-114: reads `path`, writes ./DECOMPRESSED_FILE (D:104-105 collision naming). */
int hf_extract_file(hf_ctx *ctx, const char *path);

#ifdef __cplusplus
}
#endif
#endif
