// common.cuh — shared declarations of libhuffb200 (sm_100a only).
// Layouts here are private to the library; include/huffman_b200.h is the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stddef.h>

#include "../../include/huffman_b200.h"

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "libhuffb200 is written for sm_100a (B200) only"
#endif

namespace hf {

constexpr uint32_t NSYM = HF_NSYM;

// ---- device-resident codebook (hf_codebook_bytes) ---------------------------------
struct Codebook {
    uint32_t U;
    uint32_t maxlen;
    unsigned long long table_bits;
    unsigned long long payload_bits;
    uint32_t status;
    uint32_t pad0;
    unsigned long long pad1;
    uint16_t order[NSYM];      // rank -> symbol, ascending (count, symbol)
    uint8_t len[NSYM];         // by symbol, 0 when absent
    unsigned long long code[NSYM];   // by symbol, right aligned, root->leaf
    // the encoder's shared-memory table: (1 << len) | code for 1 <= len <= 23 (0 = longer or absent),
    // low 16 bits in p16, high 8 bits in p8, both indexed by sym ^ (sym >> 8); contiguous and 16-byte
    // aligned (copied to shared memory with 128-bit loads)
    alignas(16) uint16_t p16[NSYM];
    uint8_t p8[NSYM];
    uint8_t lenf[NSYM];        // code length by folded symbol (enc_count_kernel's shared-memory plane)
};
static_assert(offsetof(Codebook, p16) % 16 == 0, "encoder planes must be 16-byte aligned");
static_assert(offsetof(Codebook, p8) == offsetof(Codebook, p16) + NSYM * 2, "encoder planes must be contiguous");
static_assert(offsetof(Codebook, lenf) % 16 == 0, "bulk copies need 16-byte alignment");

// ---- device-resident decode tables (hf_decode_table_bytes) ------------------------
// Level 1: K1-bit direct table (copied to shared memory by the decode kernels).
//   leaf    (sym << 8) | len              1 <= len <= K1, bit 7 clear
//   sub     (off << 8) | 0x80 | sub_bits  -> 2^sub_bits entries at t2[off]
//   0       hole (incomplete code)
// Level 2: per-prefix direct tables, sub_bits <= k2cap (12, or 8 when 12 would overflow t2).
//   leaf    (sym << 8) | total_len
//   list    0x80000080 | (i << 8): codes longer than K1 + sub_bits that share this slot's prefix form a
//           linked list through `longs`, head longs[i] (a handful of entries: prefix codes)
constexpr uint32_t K1 = 12;
constexpr uint32_t K2MAX = 12;
constexpr uint32_t T2_CAP = 1u << 20;          // entries; 2^K1 prefixes x 2^8 always fits
#ifndef HF_FLAT_MAX
#define HF_FLAT_MAX 22
#endif
constexpr uint32_t FLAT_MAX = HF_FLAT_MAX;     // index bits of the flat second-level planes
constexpr uint32_t MICRO_K = 14;               // index bits of the shared-memory plane; micro trees add up to 4
struct LongCode {
    unsigned long long code_left;               // left aligned in 64 bits
    uint32_t leaf;                              // (sym << 8) | len
    uint32_t next;                              // next list entry (same encoding as the slot), 0 at the end
};
struct DecodeTable {
    uint32_t U;
    uint32_t maxlen;
    uint32_t minlen;
    uint32_t len_gcd;          // gcd of all code lengths: code word boundaries are start + k*gcd
    uint32_t n_long;
    uint32_t t2_used;
    uint32_t status;
    uint32_t single_sym;       // 0x10000 | sym when U == 1 with a zero-length code
    uint32_t k2cap;
    uint32_t pad[3];
    uint32_t t1[1u << K1];
    uint32_t sub_depth[1u << K1];
    uint32_t t2[T2_CAP];
    LongCode longs[NSYM];
    // planes of the word-walk kernels (decode2.cu), derived from t1 / t2 by the dt_planes / dt_micro kernels
    //   t14     by the next 14 bits, in shared memory:
    //             (sym << 16) | (len << 1)   a code of at most 14 bits (bit 15 clear, len >= 1)
    //             (base << 16) | 0x8000 | starts   every code below this prefix has 15..18 bits (a complete "micro
    //                                        tree" of depth <= 4): starts bit j - 1 = a leaf starts at slot j of the next
    //                                        4 bits (slot 0 always starts one; bit 15 marks the entry), its symbols are
    //                                        leaves[base ...] in code order
    //             0                          deeper, incomplete or absent: take the flat planes
    //   leaves  symbols of the micro trees, in shared memory of the write kernel
    //   lenflat / flat2   length / (sym << 8) | len by the next k2 bits (<= k2 <= 22), L2 resident
    uint32_t k2;
    uint32_t pad2[3];
    alignas(16) uint32_t t14[1u << MICRO_K];
    alignas(16) uint16_t leaves[NSYM];
    // d14: what the synchronisation walks need of t14 (lengths only).  A code of at most 14 bits:
    // (len << 28) | (n << 8) | (bits << 4) | 0xC, n = code words these 14 bits hold completely and their total bits
    // (0x10C: not here, take the flat planes).  Else the micro tree as 16 two-bit fields, field j = (length at slot j)
    // - 15; a micro tree never has the low nibble 0xC: slot 0 at depth 1 covers slot 1 as well.
    alignas(16) uint32_t d14[1u << MICRO_K];
    alignas(16) uint16_t micro_sym[16u << MICRO_K];     // build scratch: symbol at every slot of every prefix
    alignas(16) uint8_t lenflat[1u << FLAT_MAX];
    alignas(16) uint32_t flat2[1u << FLAT_MAX];
};

static_assert(offsetof(DecodeTable, leaves) == offsetof(DecodeTable, t14) + (4u << MICRO_K), "t14 and leaves are loaded by one bulk copy");
static_assert(offsetof(DecodeTable, t14) % 16 == 0 && offsetof(DecodeTable, d14) % 16 == 0, "bulk copies need 16-byte alignment");

// ---- where a slice of the image lies ---------------------------------------------------------------
// Computed on the device (sharded.cu shard_plan_kernel) from the codebook's table size and the payload bit counts of
// the ranks (one rank: the codebook's own payload_bits), so that neither the sizes nor the start bit visit the host
// between the histogram and the last packed byte.  The stages that write the image return at once when status is set.
struct ShardPlan {
    unsigned long long start_bit, end_bit;      // global bits (from image byte 0) of this rank's payload slice
    unsigned long long first_byte;              // global byte of the rank's buffer[0]
    unsigned long long range_bytes;             // bytes the rank owns: the ranges of consecutive ranks tile the image
    unsigned long long own_bytes;               // bytes the rank's encoder touches (range_bytes, or one more: the seam byte)
    unsigned long long image_bytes;             // size of the whole image
    unsigned long long local_start_bit;         // start_bit - 8 * first_byte: where the encoder starts in the rank's buffer
    unsigned long long need_bytes;              // capacity this rank's buffer must have
    unsigned long long status;                  // HF_OK, HF_ERR_CAPACITY, HF_ERR_CODE_TOO_LONG
};

// ---- context ----------------------------------------------------------------------
struct Ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int sm_count;
    uint64_t launches;
    char err[512];
    // workspace (device), grown on demand
    void *ws;
    size_t ws_bytes;
    // small pinned host scratch for summaries
    void *h_scratch;
    // second stream + events for the host-buffer pipelines
    cudaStream_t copy_stream;
    cudaStream_t d2h_stream;        // results leave on their own stream (hf_decompress_host)
    unsigned long long *h_pipe;     // pinned: running symbol counts of the slices (PIPE_SLOTS)
    cudaEvent_t ev[8];
    // cached device buffers for the host-facing calls
    void *d_in; size_t d_in_bytes;
    void *d_out; size_t d_out_bytes;
    void *d_cb;
    void *d_tab;
    void *d_hist;
    void *d_scan;                   // block totals of the decoder's chunk scan (SCAN_BLOCKS_MAX u64)
    // kernels whose dynamic shared-memory limit has been raised on THIS context's device (a process may hold
    // contexts on several GPUs; the attribute is per device)
    bool smem_attr[8];
    int write_kernel;               // development aid (environment HF_WRITE_KERNEL=3 / 4): one write kernel for every chunk
    uint32_t write_split;           // symbols per chunk from which a chunk is dec_write4_kernel's (HF_WRITE_SPLIT)
    // sharded job (sharded.cu): NCCL communicator, device state, summed histogram, header staging
    void *comm;
    int rank, nranks;
    uint64_t collectives;
    void *d_shard, *d_hist2, *d_hdr;
    // pinned ring of the file programs (programs.cu): ring_n slots of ring_slot bytes, an event each
    uint8_t *ring;
    size_t ring_slot;
    int ring_n;
    cudaEvent_t ring_ev[4];
    // optional per-kernel timing (hf_profile_*): event pairs around every launch
    bool prof_on;
    bool prof_major_only;           // hf_profile_enable(ctx, 2): only the kernels that move the data
    bool prof_open;                 // a begin event is pending
    uint32_t prof_n;                // recorded pairs
    cudaEvent_t *prof_ev;           // 2 * PROF_CAP events, created on first enable
    const char **prof_name;         // PROF_CAP static strings
};
enum { ATTR_HIST = 0, ATTR_ENCODE, ATTR_PARSE, ATTR_SYNC, ATTR_WRITE, ATTR_INDEX, ATTR_WRITE4 };
// chunks (16 KiB of payload) with this many code words or more (under ~11 bits each) are written by dec_write4_kernel.
// (Measured with its 32 warps: the 1 GiB Zipf(1.2) stream, 10.6 bits per code word under its own codebook, 3.22 -> 2.27 ms
// with it; the text class of the mixed stream, 11.3 bits, is faster in dec_write3_kernel: 1.37 against ~1.7 ms per GiB.)
constexpr uint32_t WRITE_SPLIT_DEFAULT = 131072 / 11;
constexpr uint32_t PROF_CAP = 8192;
constexpr uint32_t PIPE_SLOTS = 4096;
constexpr uint32_t SCAN_BLOCKS_MAX = 1u << 16;  // x 4096 chunks x 16 KiB = 4 TiB of payload
void prof_begin(Ctx *c, const char *name);
void prof_end(Ctx *c);

int set_err(Ctx *c, int code, const char *fmt, ...);
int ensure_ws(Ctx *c, size_t bytes);

// The context's workspace (ctx->ws) is one allocation with two regions:
//   [0, WS_SCRATCH_BYTES)   short-lived scratch of one stage at a time: the histogram's per-CTA partial bins
//                           (hist.cu, sm_count x 128 KiB), the codebook builder's arrays (codebook.cu CbWork), the
//                           table builder's source arrays (decode.cu TabSrc); each stage is done with it when the next
//                           one starts (one stream);
//   [WS_STAGE_OFFSET, ...)  what lives across the kernels of an encode (unit / group bit counts, encode2.cu Enc2Work)
//                           or of a decode (DecWork + DecLayout, decode_common.cuh), sized by the input.
constexpr size_t WS_SCRATCH_BYTES = 24u << 20;
constexpr size_t WS_STAGE_OFFSET = WS_SCRATCH_BYTES;

#define HF_CUDA(ctx, call)                                                              \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess)                                                          \
            return hf::set_err((ctx), HF_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                               cudaGetErrorString(_e));                                 \
    } while (0)

// HF_PROF(ctx, "kernel") goes directly before a launch, HF_LAUNCH_CHECK(ctx) directly after
#define HF_PROF(ctx, name)                                                              \
    do {                                                                                \
        if ((ctx)->prof_on) hf::prof_begin((ctx), (name));                              \
    } while (0)

#define HF_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        if ((ctx)->prof_open) hf::prof_end((ctx));                                      \
        HF_CUDA((ctx), cudaGetLastError());                                             \
    } while (0)

// ---- stage launchers (each in its own .cu) ------------------------------------------
int launch_histogram(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, unsigned long long *d_hist);
int launch_codebook(Ctx *c, const unsigned long long *d_hist, Codebook *d_cb);
int launch_shard_bits(Ctx *c, const unsigned long long *d_hist, const Codebook *d_cb,
                      unsigned long long *d_bits);
// d_last: device address of the odd last input byte (its value is read on the device), or NULL: `last_byte` is it.
// plan: NULL, or the device plan whose status gates the stage and whose local_start_bit replaces start_bit.
int launch_header_pack(Ctx *c, const Codebook *d_cb, uint64_t n_bytes, uint32_t last_byte, const uint8_t *d_last,
                       uint8_t *d_file, uint64_t capacity, const ShardPlan *plan);
int launch_encode(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb,
                  uint8_t *d_stream, uint64_t start_bit, const ShardPlan *plan);
int launch_encode_index(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb, const uint8_t *d_stream,
                        uint64_t start_bit, uint16_t *d_rec, uint64_t n_subs);
int launch_parse_header(Ctx *c, const uint8_t *d_file, uint64_t file_bytes, DecodeTable *d_tab,
                        hf_header_info_t *d_info);
int launch_table_from_codebook(Ctx *c, const Codebook *d_cb, DecodeTable *d_tab);
int launch_decode(Ctx *c, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit,
                  uint64_t n_symbols, const DecodeTable *d_tab, uint8_t *d_out);
// a decode in slices (decode.cu; hf_decompress_host pipelines host copies around them)
struct DecodeJob {
    const uint8_t *frame;               // 16-byte aligned base the chunks are counted from
    unsigned long long frame_bytes, F0, nch, n_symbols;
    const DecodeTable *tab;
    uint16_t *out16;
    void *work;
    unsigned long long *total;          // device: symbols decoded by the slices so far
};
int decode_begin(Ctx *c, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit, uint64_t n_symbols,
                 const DecodeTable *d_tab, uint8_t *d_out, DecodeJob *job);
int decode_slice(Ctx *c, const DecodeJob &job, unsigned long long c0, unsigned long long c1);
int launch_decode_indexed(Ctx *c, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit, uint64_t n_symbols,
                          const DecodeTable *d_tab, uint8_t *d_out, const uint16_t *d_rec, uint64_t n_subs);
uint64_t index_subs(const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit);
int launch_decode_range(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes, uint64_t first_bit,
                        bool tail_only, const DecodeTable *d_tab, uint8_t *d_out, uint64_t out_symbols,
                        unsigned long long *d_result);
int launch_range_sync(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes, const DecodeTable *d_tab,
                      unsigned long long *d_probe);
int launch_range_write(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes,
                       const unsigned long long *d_first_bit, const DecodeTable *d_tab, uint8_t *d_out, uint64_t out_symbols,
                       unsigned long long *d_result);
// api.cu helpers shared with programs.cu
int ensure_buf(Ctx *c, void **p, size_t *have, size_t bytes);
double now_ms();
int compress_async(Ctx *c, const uint8_t *d_in, uint64_t n, const uint8_t *d_last, uint32_t last_byte, uint8_t *d_file,
                   uint64_t capacity, ShardPlan **d_plan);
int compress_result(Ctx *c, const ShardPlan *d_plan, uint64_t capacity, uint64_t *h_file_bytes);
int check_decode_flags(Ctx *c);
int launch_plan_single(Ctx *c, const Codebook *d_cb, uint64_t n_total, uint64_t capacity, ShardPlan **d_plan);
void shard_release(Ctx *c);
void ring_release(Ctx *c);
int launch_decompress_image(Ctx *c, const uint8_t *d_file, uint64_t file_bytes, const hf_header_info_t *d_info,
                            const DecodeTable *d_tab, uint8_t *d_out, uint64_t capacity);

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint4 ld_stream_v4(const void *p)
{   // streaming 128-bit load: read once, keep out of L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// The 32 bytes at p (a whole sector) as two 128-bit halves.  wide: p is 32-byte aligned — ONE 256-bit load (sm_100:
// LDG.E.256), so a warp whose lanes read 32 bytes each touches every sector once instead of once per half.
__device__ __forceinline__ void ld_stream_2v4(const void *p, bool wide, uint4 &a, uint4 &b)
{
    if (wide)
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
    else { a = ld_stream_v4(p); b = ld_stream_v4(reinterpret_cast<const uint8_t *>(p) + 16); }
}
__device__ __forceinline__ void st_stream_v4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t bswap32(uint32_t w) { return __byte_perm(w, 0, 0x0123); }

// ---- bulk copies (TMA, non-tensor form) of the table planes into shared memory ---------------------
// One thread arms an mbarrier with the byte count and issues cp.async.bulk global -> shared copies; everybody waits
// on the barrier's phase.  Replaces per-thread LDG.128 + STS.128 staging loops (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
constexpr uint32_t BULK_PIECE = 32768;                  // bytes per cp.async.bulk
// thread 0 of the CTA: src (16-byte aligned, `bytes` a multiple of 16) -> shared address dst, completion on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    for (uint32_t o = 0; o < bytes; o += BULK_PIECE)
        bulk_g2s(dst + o, reinterpret_cast<const uint8_t *>(src) + o, min(BULK_PIECE, bytes - o), bar);
}

// All threads of the CTA: the planes [src, src + bytes) land at shared address dst; `bar` = shared address of 8 bytes
// (8-byte aligned) for the mbarrier.  Contains a CTA barrier; one use per kernel (phase 0 of the barrier).
__device__ __forceinline__ void cta_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes);
        bulk_load(dst, src, bytes, bar);
    }
    __syncthreads();                                    // the barrier is initialised
    mbar_wait(bar, 0);
}
#endif

}  // namespace hf
