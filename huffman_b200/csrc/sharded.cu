// sharded.cu — one stream over several GPUs behind the C ABI (include/huffman_b200.h "sharded job").
//
// The reference is single-GPU (gpuHuffmanConstruction.h:678 queries device 0; no cudaSetDevice, no NCCL anywhere), so
// everything here is new; the format it produces is still the reference's: the slices of the ranks, laid
// end to end in rank order, are the byte-identical file.
//
// One process (or host thread) per GPU, one context each.  Every stage is enqueued on the context's stream and so are
// the collectives (NCCL, loaded with dlopen so that the library has no link-time dependency and shares the copy a host
// framework has already loaded); what the stages need from a collective — bit counts, hand-over bits — they read from
// DEVICE memory, so a step has no host synchronisation between its first kernel and its last, and exactly three
// collectives each way:
//   compress    all-reduce   65,536 x u64   the histogram                     (512 KiB)
//               all-gather   1 x u64        payload bits of every shard       -> shard_plan_kernel: global start bits
//               all-gather   48 B           every slice's first 32 bytes and last byte: seam bytes, read-ahead
//   decompress  broadcast    HF_HEADER_MAX  the image's header (rank 0 holds it)
//               all-gather   2 x u64        overflow of every range's last code word (speculative), range bits
//               all-gather   4 x u64        true overflow, symbol count, flags of every range
// The phase functions (hf_shard_*) are what hf_compress_sharded / hf_decompress_sharded run between their collectives;
// they are part of the ABI so that the ranks of a job can be driven by another transport (tests emulate eight ranks on
// one GPU with eight contexts; torch.distributed over gloo drives the CPU stand-in).
#include <dlfcn.h>

#include "common.cuh"
#include "decode_common.cuh"

namespace hf {

constexpr uint32_t HALO = HF_SHARD_HALO;                // bytes of the next slice kept behind each slice
constexpr uint32_t REC_BYTES = HF_SHARD_REC_BYTES;      // seam record: HALO head bytes + 1 tail byte, padded
constexpr uint32_t MAX_RANKS = HF_SHARD_MAX_RANKS;

// ---- NCCL, loaded at run time ------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclUint8 = 1, ncclUint64 = 5 };
enum { ncclSum = 0 };
struct Nccl {
    void *lib;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
};
static Nccl g_nccl;

static const char *nccl_load()
{
    if (g_nccl.lib) return nullptr;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the process already holds (e.g. torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return "libnccl.so.2 not found";
#define SYM(field, name)                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                               \
    if (!g_nccl.field) return "libnccl: missing symbol " name;
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Broadcast, "ncclBroadcast")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return nullptr;
}

#define HF_NCCL(ctx, call)                                                                          \
    do {                                                                                            \
        int _e = (call);                                                                            \
        if (_e != 0)                                                                                \
            return hf::set_err((ctx), HF_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,      \
                               g_nccl.GetErrorString(_e));                                          \
    } while (0)

// ---- device state of a sharded job (ctx->d_shard) ----------------------------------------------------
struct ShardState {
    ShardPlan plan;
    unsigned long long allbits[2 * MAX_RANKS];          // (payload bits, slice capacity) of every shard
    unsigned long long probe[2 * MAX_RANKS];            // decompress: (speculative overflow, range bits) of every range
    unsigned long long res[4 * MAX_RANKS];              // decompress: (-, overflow, symbols, flags) of every range
    unsigned long long first_bit;                       // decompress: my first code word, bits into my range
    unsigned long long out_offset;                      // decompress: symbols of the ranks before me
    unsigned long long spec_ok;                         // decompress: every true overflow confirmed the speculated one
    unsigned long long cap_short;                       // decompress: some rank's output buffer was too small
    alignas(16) uint8_t recs[MAX_RANKS * REC_BYTES];    // compress: seam records of every rank
};

static int ensure_shard(Ctx *c)
{
    if (c->d_shard) return HF_OK;
    HF_CUDA(c, cudaMalloc(&c->d_shard, sizeof(ShardState)));
    HF_CUDA(c, cudaMalloc(&c->d_hist2, NSYM * 8));
    HF_CUDA(c, cudaMalloc(&c->d_hdr, HF_HEADER_MAX));
    return HF_OK;
}

// ---- compress ----------------------------------------------------------------------------------------
struct SliceGeom {                                      // the plan arithmetic, shared by the kernels below
    unsigned long long first, end, own, next, start, bits;
};
// allbits: one (payload bits, slice capacity) pair per rank, `stride` words apart (1: bit counts only)
__device__ __forceinline__ SliceGeom slice_geom(const unsigned long long *allbits, uint32_t stride, uint32_t nranks, uint32_t r,
                                                unsigned long long start0, unsigned long long image_bytes)
{
    unsigned long long s = start0;
    for (uint32_t k = 0; k < r; k++) s += allbits[stride * k];
    SliceGeom g;
    g.start = s;
    g.bits = allbits[stride * r];
    g.first = r == 0 ? 0ull : s / 8;
    g.end = r + 1 == nranks ? image_bytes : (s + g.bits + 7) / 8;       // one past the last byte rank r wrote
    g.next = r + 1 == nranks ? image_bytes : (s + g.bits) / 8;          // first byte of rank r + 1
    g.own = g.end - g.first;
    return g;
}
__device__ __forceinline__ unsigned long long image_size(const unsigned long long *allbits, uint32_t stride, uint32_t nranks,
                                                         unsigned long long start0)
{
    unsigned long long s = start0;
    for (uint32_t k = 0; k < nranks; k++) s += allbits[stride * k];
    return (s + 7) / 8;
}
__device__ __forceinline__ unsigned long long slice_need(const SliceGeom &g, uint32_t with_halo)
{
    return (g.own > g.next - g.first ? g.own : g.next - g.first) + (with_halo ? HALO + 4 : 0);
}

// one thread: my slice of the image from the bit counts of all shards (one shard: the codebook's own payload_bits)
// sharded (stride 2): the capacities of all ranks travel with the bit counts, so every rank reaches the SAME verdict on
// whether the job fits — a rank that could not write its slice would leave its neighbours without seam bytes
__global__ void shard_plan_kernel(const Codebook *__restrict__ cb, const unsigned long long *__restrict__ allbits,
                                  uint32_t stride, uint32_t rank, uint32_t nranks, unsigned long long n_total,
                                  unsigned long long capacity, uint32_t with_halo, ShardPlan *__restrict__ plan)
{
    const unsigned long long start0 = (3 + (n_total & 1)) * 8ull + cb->table_bits + 64;
    const unsigned long long image_bytes = image_size(allbits, stride, nranks, start0);
    const SliceGeom g = slice_geom(allbits, stride, nranks, rank, start0, image_bytes);
    plan->start_bit = g.start;
    plan->end_bit = g.start + g.bits;
    plan->first_byte = g.first;
    plan->range_bytes = g.next - g.first;
    plan->own_bytes = g.own;
    plan->image_bytes = image_bytes;
    plan->local_start_bit = g.start - 8 * g.first;
    plan->need_bytes = slice_need(g, with_halo);
    bool fits = plan->need_bytes <= capacity;
    if (stride == 2)
        for (uint32_t r = 0; r < nranks; r++)
            if (slice_need(slice_geom(allbits, stride, nranks, r, start0, image_bytes), with_halo) > allbits[2 * r + 1]) fits = false;
    plan->status = cb->status ? cb->status : (fits ? 0ull : (unsigned long long)HF_ERR_CAPACITY);
}

// after the encoder: the bytes behind my last one are zero up to the end of the read-ahead, and my seam record —
// my first HALO bytes and my last byte — is ready for the all-gather
__global__ void shard_tail_kernel(const ShardPlan *__restrict__ plan, uint8_t *slice, uint8_t *rec)
{
    if (plan->status) return;
    const unsigned long long own = plan->own_bytes, range = plan->range_bytes;
    const uint32_t t = threadIdx.x;
    for (unsigned long long i = own + t; i < range + HALO + 1; i += blockDim.x) slice[i] = 0;
    __syncthreads();
    if (t < REC_BYTES) {
        uint8_t v = 0;
        const unsigned long long m = own < HALO ? own : HALO;
        if (t < HALO) v = t < m ? slice[t] : 0;
        else if (t == HALO) v = own ? slice[own - 1] : 0;
        rec[t] = v;
    }
}

// seam bytes and read-ahead: every other slice's first HALO bytes and last byte, OR-ed in where they fall into my
// window [first_byte, first_byte + range_bytes + HALO) (one thread per byte of every record)
__global__ void shard_seams_kernel(const Codebook *__restrict__ cb, const unsigned long long *__restrict__ allbits,
                                   uint32_t rank, uint32_t nranks, unsigned long long n_total,
                                   const uint8_t *__restrict__ recs, const ShardPlan *__restrict__ plan, uint8_t *slice)
{
    if (plan->status) return;
    const uint32_t r = blockIdx.x, j = threadIdx.x;
    if (r == rank || j > HALO) return;
    const unsigned long long start0 = (3 + (n_total & 1)) * 8ull + cb->table_bits + 64;
    const SliceGeom g = slice_geom(allbits, 2, nranks, r, start0, plan->image_bytes);
    if (g.own == 0) return;
    unsigned long long at;                              // global byte this record byte belongs to
    if (j < HALO) {
        if (j >= g.own) return;
        at = g.first + j;
    } else {
        if (g.own <= HALO) return;                      // the tail byte is already part of the head
        at = g.end - 1;
    }
    const unsigned long long lo = plan->first_byte, hi = lo + plan->range_bytes + HALO;
    if (at < lo || at >= hi) return;
    const uint8_t v = recs[r * REC_BYTES + j];
    if (v == 0) return;
    // several records can meet in one byte (a seam byte holds the last bits of one slice and the first of the next)
    uint8_t *p = slice + (at - lo);
    uint32_t *w = reinterpret_cast<uint32_t *>((uintptr_t)p & ~(uintptr_t)3);
    atomicOr(w, (uint32_t)v << (8 * ((uintptr_t)p & 3)));
}

// ---- decompress --------------------------------------------------------------------------------------
// the hand-over chain (one thread): first[0] is where the header ends; a range no code word starts in passes the
// bit on.  Leaves my first bit for the write phase.
__global__ void shard_chain_kernel(const hf_header_info_t *__restrict__ info, const unsigned long long *__restrict__ probe,
                                   uint32_t rank, uint32_t nranks, ShardState *st)
{
    unsigned long long first = info->payload_start_bit;
    for (uint32_t r = 0; r < rank; r++) {
        const unsigned long long rb = probe[2 * r + 1];
        first = first >= rb ? first - rb : probe[2 * r];
    }
    st->first_bit = first;
}

// One distinct symbol with a zero-length code: the payload is empty and there is nothing to shard;
// rank 0 writes the original_bytes / 2 copies of the symbol and reports them as its range's count.
__global__ void shard_single_fill_kernel(const DecodeTable *__restrict__ tab, const hf_header_info_t *__restrict__ info,
                                         uint32_t rank, uint16_t *__restrict__ out, unsigned long long out_symbols,
                                         unsigned long long *__restrict__ res)
{
    if (!(tab->single_sym & 0x10000u)) return;
    const unsigned long long n = rank == 0 ? info->original_bytes / 2 : 0;
    const bool fits = n <= out_symbols;
    const uint16_t s = (uint16_t)tab->single_sym;
    if (fits)
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (unsigned long long)gridDim.x * blockDim.x) out[i] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) { res[1] = 0; res[2] = n; res[3] = fits ? 0ull : 8ull; }
}

// after the second all-gather: did every range end where the speculation said, and where does my output start?
__global__ void shard_finish_kernel(const hf_header_info_t *__restrict__ info, const DecodeTable *__restrict__ tab,
                                    uint32_t rank, uint32_t nranks, ShardState *st)
{
    unsigned long long first = info->payload_start_bit, off = 0, ok = 1, cap = 0;
    const bool no_payload = tab->single_sym != 0;       // nothing to hand over
    for (uint32_t r = 0; r < nranks; r++) {
        const unsigned long long rb = st->probe[2 * r + 1];
        const unsigned long long over = first >= rb ? first - rb : st->probe[2 * r];
        if ((st->res[4 * r + 3] & 4) || (!no_payload && st->res[4 * r + 1] != over)) ok = 0;
        if (st->res[4 * r + 3] & 8) cap = 1;
        if (r < rank) off += st->res[4 * r + 2];
        first = over;
    }
    st->out_offset = off;
    st->spec_ok = ok;
    st->cap_short = cap;
}

// the plan of an unsharded compress (hf_compress, hf_compress_host): one rank whose bit count is the codebook's own
int launch_plan_single(Ctx *c, const Codebook *d_cb, uint64_t n_total, uint64_t capacity, ShardPlan **d_plan)
{
    int rc = ensure_shard(c);
    if (rc) return rc;
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    shard_plan_kernel<<<1, 1, 0, c->stream>>>(d_cb, &d_cb->payload_bits, 1u, 0u, 1u, n_total, capacity, 0u, &st->plan);
    HF_LAUNCH_CHECK(c);
    *d_plan = &st->plan;
    return HF_OK;
}

void shard_release(Ctx *c)
{
    if (c->comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    if (c->d_shard) cudaFree(c->d_shard);
    if (c->d_hist2) cudaFree(c->d_hist2);
    if (c->d_hdr) cudaFree(c->d_hdr);
    c->d_shard = c->d_hist2 = c->d_hdr = nullptr;
}

}  // namespace hf

using namespace hf;

#define CTX(c) reinterpret_cast<Ctx *>(c)
#define NEED_CTX(c)                         \
    if (!(c)) return HF_ERR_ARG;            \
    do {                                    \
        cudaError_t _e = cudaSetDevice(CTX(c)->device); \
        if (_e != cudaSuccess) return set_err(CTX(c), HF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(_e)); \
    } while (0)

extern "C" {

int hf_comm_unique_id(hf_unique_id_t *id)
{
    if (!id) return HF_ERR_ARG;
    if (nccl_load()) return HF_ERR_CUDA;
    static_assert(sizeof(hf_unique_id_t) == sizeof(ncclUniqueId), "hf_unique_id_t is an ncclUniqueId");
    return g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id)) == 0 ? HF_OK : HF_ERR_CUDA;
}

int hf_comm_init(hf_ctx *ctx, const hf_unique_id_t *id, int rank, int nranks)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!id || rank < 0 || nranks < 1 || rank >= nranks || nranks > (int)MAX_RANKS)
        return set_err(c, HF_ERR_ARG, "hf_comm_init: rank %d of %d", rank, nranks);
    if (const char *why = nccl_load()) return set_err(c, HF_ERR_CUDA, "hf_comm_init: %s", why);
    if (c->comm) { g_nccl.CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    HF_NCCL(c, g_nccl.CommInitRank(&comm, nranks, uid, rank));
    c->comm = comm;
    c->rank = rank;
    c->nranks = nranks;
    return ensure_shard(c);
}

int hf_comm_destroy(hf_ctx *ctx)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (c->comm) {
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
    }
    c->rank = 0;
    c->nranks = 1;
    return HF_OK;
}

// ---- compress phases ---------------------------------------------------------------------------------
int hf_shard_compress_local(hf_ctx *ctx, const uint8_t *d_chunk, uint64_t chunk_bytes, uint64_t *d_hist_local)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_hist_local || (!d_chunk && chunk_bytes)) return set_err(c, HF_ERR_ARG, "hf_shard_compress_local: null pointer");
    HF_CUDA(c, cudaMemsetAsync(d_hist_local, 0, NSYM * 8, c->stream));
    return launch_histogram(c, d_chunk, chunk_bytes, reinterpret_cast<unsigned long long *>(d_hist_local));
}

__global__ void shard_capacity_kernel(unsigned long long *slot, unsigned long long capacity) { *slot = capacity; }

int hf_shard_compress_bits(hf_ctx *ctx, const uint64_t *d_hist_total, const uint64_t *d_hist_local, int rank,
                           uint64_t slice_capacity, uint64_t *d_allbits)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_hist_total || !d_hist_local || !d_allbits || rank < 0 || rank >= (int)MAX_RANKS)
        return set_err(c, HF_ERR_ARG, "hf_shard_compress_bits: bad argument");
    Codebook *cb = reinterpret_cast<Codebook *>(c->d_cb);
    int rc = launch_codebook(c, reinterpret_cast<const unsigned long long *>(d_hist_total), cb);
    if (rc) return rc;
    rc = launch_shard_bits(c, reinterpret_cast<const unsigned long long *>(d_hist_local), cb,
                           reinterpret_cast<unsigned long long *>(d_allbits) + 2 * rank);
    if (rc) return rc;
    shard_capacity_kernel<<<1, 1, 0, c->stream>>>(reinterpret_cast<unsigned long long *>(d_allbits) + 2 * rank + 1, slice_capacity);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int hf_shard_compress_pack(hf_ctx *ctx, const uint8_t *d_chunk, uint64_t chunk_bytes, uint64_t n_total, uint32_t last_byte,
                           int rank, int nranks, const uint64_t *d_allbits, uint8_t *d_slice, uint64_t capacity, uint8_t *d_recs)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_allbits || !d_slice || !d_recs || rank < 0 || nranks < 1 || rank >= nranks || nranks > (int)MAX_RANKS)
        return set_err(c, HF_ERR_ARG, "hf_shard_compress_pack: bad argument");
    if ((uintptr_t)d_slice & 15) return set_err(c, HF_ERR_ARG, "hf_shard_compress_pack: the slice buffer must be 16-byte aligned");
    if (capacity < HALO + 64) return set_err(c, HF_ERR_CAPACITY, "hf_shard_compress_pack: capacity %llu too small", (unsigned long long)capacity);
    int rc = ensure_shard(c);
    if (rc) return rc;
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    Codebook *cb = reinterpret_cast<Codebook *>(c->d_cb);
    shard_plan_kernel<<<1, 1, 0, c->stream>>>(cb, reinterpret_cast<const unsigned long long *>(d_allbits), 2u, (uint32_t)rank,
                                              (uint32_t)nranks, n_total, capacity, 1u, &st->plan);
    HF_LAUNCH_CHECK(c);
    if (rank == 0) {
        rc = launch_header_pack(c, cb, n_total, last_byte, nullptr, d_slice, capacity, &st->plan);      // zeroes the header region first
        if (rc) return rc;
    } else {
        HF_CUDA(c, cudaMemsetAsync(d_slice, 0, 16, c->stream));        // the bits before my start phase are the neighbour's
    }
    rc = launch_encode(c, d_chunk, chunk_bytes, cb, d_slice, 0, &st->plan);
    if (rc) return rc;
    shard_tail_kernel<<<1, 256, 0, c->stream>>>(&st->plan, d_slice, d_recs + (size_t)rank * REC_BYTES);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int hf_shard_compress_seams(hf_ctx *ctx, uint64_t n_total, int rank, int nranks, const uint64_t *d_allbits,
                            const uint8_t *d_recs, uint8_t *d_slice, hf_slice_info_t *h_info)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_allbits || !d_recs || !d_slice || !h_info || !c->d_shard) return set_err(c, HF_ERR_ARG, "hf_shard_compress_seams: bad argument");
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    Codebook *cb = reinterpret_cast<Codebook *>(c->d_cb);
    if (nranks > 1) {
        shard_seams_kernel<<<nranks, 64, 0, c->stream>>>(cb, reinterpret_cast<const unsigned long long *>(d_allbits), (uint32_t)rank,
                                                         (uint32_t)nranks, n_total, d_recs, &st->plan, d_slice);
        HF_LAUNCH_CHECK(c);
    }
    ShardPlan *h = reinterpret_cast<ShardPlan *>((uint8_t *)c->h_scratch + 2560);
    HF_CUDA(c, cudaMemcpyAsync(h, &st->plan, sizeof(ShardPlan), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));       // the one synchronisation of a compress step
    h_info->first_byte = h->first_byte;
    h_info->range_bytes = h->range_bytes;
    h_info->start_bit = h->start_bit;
    h_info->end_bit = h->end_bit;
    h_info->image_bytes = h->image_bytes;
    h_info->n_total = n_total;
    h_info->needed_capacity = h->need_bytes;
    if (h->status == HF_ERR_CAPACITY)        // on EVERY rank when any rank's slice does not fit: the job is retried as a whole
        return set_err(c, HF_ERR_CAPACITY, "sharded compress: a slice does not fit (this rank's needs %llu bytes)",
                       (unsigned long long)h->need_bytes);
    if (h->status) return set_err(c, (int)h->status, "sharded compress: a code word is longer than 64 bits");
    return HF_OK;
}

int hf_compress_sharded(hf_ctx *ctx, const uint8_t *d_chunk, uint64_t chunk_bytes, uint64_t n_total, uint32_t last_byte,
                        uint8_t *d_slice, uint64_t capacity, hf_slice_info_t *h_info)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!h_info) return set_err(c, HF_ERR_ARG, "hf_compress_sharded: null pointer");
    int rc = ensure_shard(c);
    if (rc) return rc;
    const int rank = c->rank, nranks = c->nranks < 1 ? 1 : c->nranks;
    if (nranks > 1 && !c->comm) return set_err(c, HF_ERR_ARG, "hf_compress_sharded: hf_comm_init first");
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    uint64_t *local = reinterpret_cast<uint64_t *>(c->d_hist), *total = reinterpret_cast<uint64_t *>(c->d_hist2);
    ncclComm_t comm = (ncclComm_t)c->comm;
    rc = hf_shard_compress_local(ctx, d_chunk, chunk_bytes, local);
    if (rc) return rc;
    if (nranks > 1) { HF_NCCL(c, g_nccl.AllReduce(local, total, NSYM, ncclUint64, ncclSum, comm, c->stream)); c->collectives++; }
    else total = local;
    rc = hf_shard_compress_bits(ctx, total, local, rank, capacity, (uint64_t *)st->allbits);
    if (rc) return rc;
    if (nranks > 1) { HF_NCCL(c, g_nccl.AllGather(st->allbits + 2 * rank, st->allbits, 2, ncclUint64, comm, c->stream)); c->collectives++; }
    rc = hf_shard_compress_pack(ctx, d_chunk, chunk_bytes, n_total, last_byte, rank, nranks, (const uint64_t *)st->allbits, d_slice,
                                capacity, st->recs);
    if (rc) return rc;
    if (nranks > 1) {
        HF_NCCL(c, g_nccl.AllGather(st->recs + (size_t)rank * REC_BYTES, st->recs, REC_BYTES, ncclUint8, comm, c->stream));
        c->collectives++;
    }
    return hf_shard_compress_seams(ctx, n_total, rank, nranks, (const uint64_t *)st->allbits, st->recs, d_slice, h_info);
}

// ---- decompress phases -------------------------------------------------------------------------------
int hf_shard_decompress_header(hf_ctx *ctx, int rank, const uint8_t *d_slice, uint64_t avail_bytes, uint8_t *d_hdr)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_hdr) return set_err(c, HF_ERR_ARG, "hf_shard_decompress_header: null pointer");
    if (rank != 0) return HF_OK;                        // the broadcast fills the others' copies
    if (!d_slice) return set_err(c, HF_ERR_ARG, "hf_shard_decompress_header: null pointer");
    const uint64_t m = avail_bytes < HF_HEADER_MAX ? avail_bytes : HF_HEADER_MAX;
    HF_CUDA(c, cudaMemsetAsync(d_hdr, 0, HF_HEADER_MAX, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(d_hdr, d_slice, m, cudaMemcpyDeviceToDevice, c->stream));
    return HF_OK;
}

int hf_shard_decompress_sync(hf_ctx *ctx, int rank, const uint8_t *d_hdr, uint64_t image_bytes, const uint8_t *d_slice,
                             uint64_t range_bytes, uint64_t halo_bytes, uint64_t *d_probe)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_hdr || !d_probe || (!d_slice && range_bytes) || rank < 0 || rank >= (int)MAX_RANKS)
        return set_err(c, HF_ERR_ARG, "hf_shard_decompress_sync: bad argument");
    if (halo_bytes < 16) return set_err(c, HF_ERR_ARG, "hf_shard_decompress_sync: at least 16 bytes of read-ahead");
    // every rank parses the (same) header and builds the (same) tables; the header's own byte count bounds what it reads
    hf_header_info_t *d_info = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->d_hist + NSYM * 8);
    DecodeTable *tab = reinterpret_cast<DecodeTable *>(c->d_tab);
    const uint64_t hdr_bytes = image_bytes < HF_HEADER_MAX ? image_bytes : HF_HEADER_MAX;
    int rc = launch_parse_header(c, d_hdr, hdr_bytes, tab, d_info);
    if (rc) return rc;
    return launch_range_sync(c, d_slice, range_bytes, halo_bytes, tab, reinterpret_cast<unsigned long long *>(d_probe) + 2 * rank);
}

int hf_shard_decompress_write(hf_ctx *ctx, int rank, int nranks, const uint64_t *d_probe, const uint8_t *d_slice,
                              uint64_t range_bytes, uint64_t halo_bytes, uint8_t *d_out, uint64_t out_capacity, uint64_t *d_res)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_probe || !d_res || !d_out || rank < 0 || nranks < 1 || rank >= nranks || nranks > (int)MAX_RANKS)
        return set_err(c, HF_ERR_ARG, "hf_shard_decompress_write: bad argument");
    int rc = ensure_shard(c);
    if (rc) return rc;
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    hf_header_info_t *d_info = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->d_hist + NSYM * 8);
    shard_chain_kernel<<<1, 1, 0, c->stream>>>(d_info, reinterpret_cast<const unsigned long long *>(d_probe), (uint32_t)rank,
                                               (uint32_t)nranks, st);
    HF_LAUNCH_CHECK(c);
    rc = launch_range_write(c, d_slice, range_bytes, halo_bytes, &st->first_bit, reinterpret_cast<DecodeTable *>(c->d_tab), d_out,
                            out_capacity / 2, reinterpret_cast<unsigned long long *>(d_res) + 4 * rank);
    if (rc) return rc;
    shard_single_fill_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(reinterpret_cast<DecodeTable *>(c->d_tab), d_info, (uint32_t)rank,
                                                                    reinterpret_cast<uint16_t *>(d_out), out_capacity / 2,
                                                                    reinterpret_cast<unsigned long long *>(d_res) + 4 * rank);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int hf_shard_decompress_finish(hf_ctx *ctx, int rank, int nranks, const uint64_t *d_probe, const uint64_t *d_res,
                               hf_shard_out_t *h_out)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_probe || !d_res || !h_out || !c->d_shard) return set_err(c, HF_ERR_ARG, "hf_shard_decompress_finish: bad argument");
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    hf_header_info_t *d_info = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->d_hist + NSYM * 8);
    if ((const void *)d_probe != (const void *)st->probe)
        HF_CUDA(c, cudaMemcpyAsync(st->probe, d_probe, 16 * nranks, cudaMemcpyDeviceToDevice, c->stream));
    if ((const void *)d_res != (const void *)st->res)
        HF_CUDA(c, cudaMemcpyAsync(st->res, d_res, 32 * nranks, cudaMemcpyDeviceToDevice, c->stream));
    shard_finish_kernel<<<1, 1, 0, c->stream>>>(d_info, reinterpret_cast<DecodeTable *>(c->d_tab), (uint32_t)rank, (uint32_t)nranks, st);
    HF_LAUNCH_CHECK(c);
    uint8_t *h = (uint8_t *)c->h_scratch + 2560;
    hf_header_info_t *hi = reinterpret_cast<hf_header_info_t *>(h);
    unsigned long long *hs = reinterpret_cast<unsigned long long *>(h + 64);      // out_offset, spec_ok, cap_short
    unsigned long long *hr = reinterpret_cast<unsigned long long *>(h + 96);      // my result
    uint32_t *ht = reinterpret_cast<uint32_t *>(h + 160);                          // table summary
    HF_CUDA(c, cudaMemcpyAsync(hi, d_info, sizeof(hf_header_info_t), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(hs, &st->out_offset, 24, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(hr, st->res + 4 * rank, 32, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(ht, c->d_tab, 48, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));       // the one synchronisation of a decompress step
    if (hi->status || ht[6]) return set_err(c, HF_ERR_FORMAT, "sharded decompress: malformed header");
    const unsigned long long nsym_total = hi->original_bytes / 2;
    unsigned long long off = hs[0] < nsym_total ? hs[0] : nsym_total, mine = hr[2];
    if (off + mine > nsym_total) mine = nsym_total - off;       // padding bits past the last code word decode to nothing of ours
    h_out->n_total = hi->original_bytes;
    h_out->out_offset = 2 * off;
    h_out->out_bytes = 2 * mine;
    h_out->max_code_bits = ht[1];
    h_out->is_odd = hi->is_odd;
    h_out->last_byte = hi->last_byte;
    h_out->payload_start_bit = hi->payload_start_bit;
    h_out->needed_symbols = hr[2];
    // 0: done.  1: the stream did not re-synchronise where the ranks speculated (or holds invalid bits): the caller
    // gathers the image on one rank and decodes it there.  2: a rank's output buffer was too small (needed_symbols).
    h_out->status = !hs[1] ? 1u : (hs[2] ? 2u : 0u);      // the same on every rank: the fall-back is a collective
    return HF_OK;
}

int hf_decompress_sharded(hf_ctx *ctx, const uint8_t *d_slice, uint64_t range_bytes, uint64_t halo_bytes, uint64_t image_bytes,
                          uint8_t *d_out, uint64_t out_capacity, hf_shard_out_t *h_out)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!h_out) return set_err(c, HF_ERR_ARG, "hf_decompress_sharded: null pointer");
    int rc = ensure_shard(c);
    if (rc) return rc;
    const int rank = c->rank, nranks = c->nranks < 1 ? 1 : c->nranks;
    if (nranks > 1 && !c->comm) return set_err(c, HF_ERR_ARG, "hf_decompress_sharded: hf_comm_init first");
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    ncclComm_t comm = (ncclComm_t)c->comm;
    uint8_t *d_hdr = reinterpret_cast<uint8_t *>(c->d_hdr);
    rc = hf_shard_decompress_header(ctx, rank, d_slice, range_bytes + halo_bytes, d_hdr);
    if (rc) return rc;
    if (nranks > 1) { HF_NCCL(c, g_nccl.Broadcast(d_hdr, d_hdr, HF_HEADER_MAX, ncclUint8, 0, comm, c->stream)); c->collectives++; }
    rc = hf_shard_decompress_sync(ctx, rank, d_hdr, image_bytes, d_slice, range_bytes, halo_bytes, (uint64_t *)st->probe);
    if (rc) return rc;
    if (nranks > 1) { HF_NCCL(c, g_nccl.AllGather(st->probe + 2 * rank, st->probe, 2, ncclUint64, comm, c->stream)); c->collectives++; }
    rc = hf_shard_decompress_write(ctx, rank, nranks, (const uint64_t *)st->probe, d_slice, range_bytes, halo_bytes, d_out,
                                   out_capacity, (uint64_t *)st->res);
    if (rc) return rc;
    if (nranks > 1) { HF_NCCL(c, g_nccl.AllGather(st->res + 4 * rank, st->res, 4, ncclUint64, comm, c->stream)); c->collectives++; }
    return hf_shard_decompress_finish(ctx, rank, nranks, (const uint64_t *)st->probe, (const uint64_t *)st->res, h_out);
}

// the whole image on rank 0 (the fall-back of hf_decompress_sharded's status 1, and what a caller who wants one file
// does after hf_compress_sharded): d_image on rank 0 receives every rank's range at its first_byte
int hf_gather_image(hf_ctx *ctx, const uint8_t *d_slice, uint64_t first_byte, uint64_t range_bytes, uint8_t *d_image,
                    uint64_t image_capacity)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    int rc = ensure_shard(c);
    if (rc) return rc;
    const int rank = c->rank, nranks = c->nranks < 1 ? 1 : c->nranks;
    if (nranks == 1) {
        if (range_bytes > image_capacity) return set_err(c, HF_ERR_CAPACITY, "hf_gather_image: capacity");
        HF_CUDA(c, cudaMemcpyAsync(d_image, d_slice, range_bytes, cudaMemcpyDeviceToDevice, c->stream));
        return HF_OK;
    }
    if (!c->comm) return set_err(c, HF_ERR_ARG, "hf_gather_image: hf_comm_init first");
    ShardState *st = reinterpret_cast<ShardState *>(c->d_shard);
    ncclComm_t comm = (ncclComm_t)c->comm;
    unsigned long long *h = reinterpret_cast<unsigned long long *>((uint8_t *)c->h_scratch + 2560);
    h[2 * rank] = first_byte;
    h[2 * rank + 1] = range_bytes;
    HF_CUDA(c, cudaMemcpyAsync(st->probe + 2 * rank, h + 2 * rank, 16, cudaMemcpyHostToDevice, c->stream));
    HF_NCCL(c, g_nccl.AllGather(st->probe + 2 * rank, st->probe, 2, ncclUint64, comm, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(h, st->probe, 16 * nranks, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (rank == 0) {
        for (int r = 0; r < nranks; r++)
            if (h[2 * r] + h[2 * r + 1] > image_capacity) return set_err(c, HF_ERR_CAPACITY, "hf_gather_image: capacity");
    }
    HF_NCCL(c, g_nccl.GroupStart());
    if (rank == 0) {
        for (int r = 1; r < nranks; r++)
            if (h[2 * r + 1]) HF_NCCL(c, g_nccl.Recv(d_image + h[2 * r], h[2 * r + 1], ncclUint8, r, comm, c->stream));
    } else if (range_bytes) {
        HF_NCCL(c, g_nccl.Send(d_slice, range_bytes, ncclUint8, 0, comm, c->stream));
    }
    HF_NCCL(c, g_nccl.GroupEnd());
    if (rank == 0 && range_bytes) HF_CUDA(c, cudaMemcpyAsync(d_image + first_byte, d_slice, range_bytes, cudaMemcpyDeviceToDevice, c->stream));
    return HF_OK;
}

uint64_t hf_collective_count(hf_ctx *ctx) { return ctx ? CTX(ctx)->collectives : 0; }

}  // extern "C"
