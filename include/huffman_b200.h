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
 * recorded launches by kernel name into out[0..*n_out) and clears the record.  on = 1: every kernel; on = 2: only the
 * kernels that move the data (histogram, unit bits, pack, synchronisation, write, codebook sort + tree). */
typedef struct {
    char name[48];
    uint32_t launches;
    float total_ms;
} hf_kernel_time_t;
int hf_profile_enable(hf_ctx *ctx, int on);
int hf_profile_read(hf_ctx *ctx, hf_kernel_time_t *out, uint32_t cap, uint32_t *n_out);

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
 * bits per 512-symbol unit with their scan in the same pass (decoupled look-back), one warp-independent packing kernel.
 * Writes the code words of the byte pairs of
 * d_in[0 .. n_bytes & ~1) as one MSB-first bit stream starting `start_bit` bits after
 * d_stream.  Bits of the first byte before the start phase are preserved (they belong to
 * the header or to the previous shard); the final partial byte is zero padded. */
int hf_encode(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, const void *d_codebook,
              uint8_t *d_stream, uint64_t start_bit);

/* whole `archive` data path on device buffers: histogram, codebook, header, encode.  The sizes and the start bit of the
 * payload stay on the device (the capacity is checked there): nothing waits for the host until everything is enqueued.
 * Synchronises once, at the end; *h_file_bytes receives the image size (also when it exceeds the capacity). */
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

/* whole `extract` data path on device buffers: the header is parsed and the tables are built on the device, the decode
 * starts where the device says the header ends, the capacity is checked there.  Synchronises once, at the end. */
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

/* ---- sharded job: ONE stream over several GPUs ------------ */

/* The reference is single-GPU (h:678 queries device 0; no cudaSetDevice, no NCCL anywhere): this layer is new, the
 * format is not — the slices of the ranks, laid end to end in rank order, are the byte-identical file.  One process
 * (or host thread) and one context per GPU.  The input is cut at multiples of 16 bytes (hf_shard_bounds); every rank
 * histograms its chunk, the histogram is all-reduced, every rank builds the identical codebook (C:378-425, h:353-494),
 * an all-gather of one u64 (the shard's payload bits) gives every rank its global start bit ON THE DEVICE, and the
 * rank packs its slice of the single stream at that bit phase (C:541-588); seam bytes are OR-merged from an all-gather
 * of 48 bytes.  Decompression (new: D:259-284 decodes serially on the host): the header is broadcast, every rank
 * synchronises its byte range speculatively, an all-gather of the overflows hands every rank its first bit on the
 * device, and an all-gather of the counts its output offset.  A step has three collectives each way plus the header
 * broadcast, all on the context's stream (NCCL, loaded at run time), and ONE host synchronisation, at its end. */
#define HF_SHARD_HALO 32u         /* bytes of the following slice kept behind each slice (read-ahead of the decoder) */
#define HF_SHARD_REC_BYTES 48u    /* seam record of a rank: its first 32 bytes, its last byte, padding */
#define HF_SHARD_MAX_RANKS 64u
#define HF_HEADER_MAX 720928u     /* >= any header: 4 + 65536 * (16 + 8 + 64) / 8 + 8, rounded up to 16 */

typedef struct { char internal[128]; } hf_unique_id_t;      /* an ncclUniqueId */

typedef struct {
    uint64_t first_byte;          /* index of slice[0] in the whole image */
    uint64_t range_bytes;         /* bytes this rank owns: the ranges of consecutive ranks tile the image */
    uint64_t start_bit, end_bit;  /* global bits of this rank's payload */
    uint64_t image_bytes;         /* size of the whole image */
    uint64_t n_total;             /* original byte count of the whole input */
    uint64_t needed_capacity;     /* bytes this rank's slice buffer must hold (HF_ERR_CAPACITY: retry with at least this) */
} hf_slice_info_t;

typedef struct {
    uint64_t n_total;             /* original byte count (D:243-255) */
    uint64_t out_offset;          /* byte offset of this rank's output in the original */
    uint64_t out_bytes;           /* bytes this rank decoded into d_out */
    uint64_t payload_start_bit;
    uint64_t needed_symbols;      /* symbols this rank's range holds (status 2: what the buffer must take) */
    uint32_t max_code_bits, is_odd, last_byte;
    uint32_t status;              /* 0 done; 1 the stream does not re-synchronise where the ranks speculated: gather it on
                                     one rank (hf_gather_image) and hf_decompress it there; 2 an output buffer was too small.
                                     The same on every rank. */
} hf_shard_out_t;

/* rank 0 creates the id, the caller distributes it (MPI, a file, torch.distributed, ...), every rank calls hf_comm_init */
int hf_comm_unique_id(hf_unique_id_t *id);
int hf_comm_init(hf_ctx *ctx, const hf_unique_id_t *id, int rank, int nranks);
int hf_comm_destroy(hf_ctx *ctx);
uint64_t hf_collective_count(hf_ctx *ctx);        /* NCCL calls this context has issued (bench.py) */

/* d_chunk: this rank's bytes of the input (even offset, 16-byte aligned cut); d_slice: 16-byte aligned, capacity >=
 * hf_compress_bound(chunk_bytes) + HF_SHARD_HALO + 64 is enough unless the chunk codes worse than 16 bits per byte pair
 * under the codebook of the WHOLE input; when any rank's slice does not fit, EVERY rank returns HF_ERR_CAPACITY (the
 * capacities travel with the bit counts) with h_info->needed_capacity set, and the job is retried as a whole.  On
 * return d_slice[0 .. range_bytes) is this rank's part of the image and HF_SHARD_HALO bytes of the next rank's part
 * follow it.  Synchronises once, at the end. */
int hf_compress_sharded(hf_ctx *ctx, const uint8_t *d_chunk, uint64_t chunk_bytes, uint64_t n_total, uint32_t last_byte,
                        uint8_t *d_slice, uint64_t capacity, hf_slice_info_t *h_info);
/* d_slice: range_bytes of the image followed by halo_bytes (>= 16) of read-ahead (the next bytes of the image, zeros
 * past its end); only rank 0's slice must start at image byte 0.  d_out takes this rank's part of the output
 * (out_capacity bytes; n_total / nranks + 64 KiB is enough unless the ranges are very uneven).  Synchronises once. */
int hf_decompress_sharded(hf_ctx *ctx, const uint8_t *d_slice, uint64_t range_bytes, uint64_t halo_bytes, uint64_t image_bytes,
                          uint8_t *d_out, uint64_t out_capacity, hf_shard_out_t *h_out);
/* every rank's range to rank 0's d_image (the one file; also the fall-back of status 1) */
int hf_gather_image(hf_ctx *ctx, const uint8_t *d_slice, uint64_t first_byte, uint64_t range_bytes, uint8_t *d_image,
                    uint64_t image_capacity);

/* The phases hf_compress_sharded / hf_decompress_sharded run between their collectives, for callers with another
 * transport (and for tests: eight contexts on one GPU emulate eight ranks).  All asynchronous except the last of each.
 *   compress:   local (histogram of my chunk)                       -> sum d_hist_local over the ranks into d_hist_total
 *               bits (codebook from the total; my (payload bits, slice capacity) into d_allbits[2 * rank])
 *                                                                    -> all-gather d_allbits
 *               pack (plan, header on rank 0, my slice, my seam record into d_recs[rank * HF_SHARD_REC_BYTES])
 *                                                                    -> all-gather d_recs
 *               seams (OR-merge, read-ahead; synchronises, fills h_info)
 *   decompress: header (rank 0 copies the head of its slice into d_hdr[HF_HEADER_MAX])  -> broadcast d_hdr
 *               sync (tables; speculative synchronisation of my range; (overflow, bits) into d_probe[2 * rank])
 *                                                                    -> all-gather d_probe
 *               write (my first bit from the chain; head repair; symbols out; result into d_res[4 * rank])
 *                                                                    -> all-gather d_res
 *               finish (checks the speculation, output offsets; synchronises, fills h_out) */
int hf_shard_compress_local(hf_ctx *ctx, const uint8_t *d_chunk, uint64_t chunk_bytes, uint64_t *d_hist_local);
int hf_shard_compress_bits(hf_ctx *ctx, const uint64_t *d_hist_total, const uint64_t *d_hist_local, int rank,
                           uint64_t slice_capacity, uint64_t *d_allbits);
int hf_shard_compress_pack(hf_ctx *ctx, const uint8_t *d_chunk, uint64_t chunk_bytes, uint64_t n_total, uint32_t last_byte,
                           int rank, int nranks, const uint64_t *d_allbits, uint8_t *d_slice, uint64_t capacity,
                           uint8_t *d_recs);
int hf_shard_compress_seams(hf_ctx *ctx, uint64_t n_total, int rank, int nranks, const uint64_t *d_allbits,
                            const uint8_t *d_recs, uint8_t *d_slice, hf_slice_info_t *h_info);
int hf_shard_decompress_header(hf_ctx *ctx, int rank, const uint8_t *d_slice, uint64_t avail_bytes, uint8_t *d_hdr);
int hf_shard_decompress_sync(hf_ctx *ctx, int rank, const uint8_t *d_hdr, uint64_t image_bytes, const uint8_t *d_slice,
                             uint64_t range_bytes, uint64_t halo_bytes, uint64_t *d_probe);
int hf_shard_decompress_write(hf_ctx *ctx, int rank, int nranks, const uint64_t *d_probe, const uint8_t *d_slice,
                              uint64_t range_bytes, uint64_t halo_bytes, uint8_t *d_out, uint64_t out_capacity,
                              uint64_t *d_res);
int hf_shard_decompress_finish(hf_ctx *ctx, int rank, int nranks, const uint64_t *d_probe, const uint64_t *d_res,
                               hf_shard_out_t *h_out);

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
