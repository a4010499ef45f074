// api.cu — the C ABI of libhuffb200 (include/huffman_b200.h): context, stage wrappers,
// whole-file device calls, host-buffer pipelines and the two program-level entry points
// that mirror the reference's mains (/root/reference/Compressor.cu:315-632,
// /root/reference/Decompressor.cu:47-114).  No CPU fallback anywhere: without a usable
// GPU every call returns HF_ERR_CUDA.
#include <stdarg.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <sys/time.h>

#include "common.cuh"

namespace hf {

size_t codebook_alloc_bytes();          // codebook.cu

int set_err(Ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int ensure_ws(Ctx *c, size_t bytes)
{
    if (bytes <= c->ws_bytes) return HF_OK;
    size_t want = bytes + (bytes >> 2);
    if (want < (32u << 20)) want = 32u << 20;
    if (c->ws) HF_CUDA(c, cudaFree(c->ws));     // synchronises: no kernel still uses the old block
    c->ws = nullptr; c->ws_bytes = 0;
    HF_CUDA(c, cudaMalloc(&c->ws, want));
    c->ws_bytes = want;
    return HF_OK;
}

// the kernels that move the data (a second profiling level times only these: an event pair costs ~5 us of a
// launch-bound sequence of small kernels, 0.2 ms of a sharded step)
static bool prof_major(const char *name)
{
    static const char *const major[] = {"hist_smem", "enc_bits", "encode2", "dec_sync4", "dec_write", "cb_sort_tree"};
    for (const char *m : major) if (strncmp(name, m, strlen(m)) == 0) return true;
    return false;
}

void prof_begin(Ctx *c, const char *name)
{
    if (c->prof_n >= PROF_CAP) return;
    if (c->prof_major_only && !prof_major(name)) return;
    c->prof_name[c->prof_n] = name;
    if (cudaEventRecord(c->prof_ev[2 * c->prof_n], c->stream) == cudaSuccess) c->prof_open = true;
}

void prof_end(Ctx *c)
{
    cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], c->stream);
    c->prof_n++;
    c->prof_open = false;
}

int ensure_buf(Ctx *c, void **p, size_t *have, size_t bytes)
{
    if (bytes <= *have) return HF_OK;
    if (*p) HF_CUDA(c, cudaFree(*p));
    *p = nullptr; *have = 0;
    HF_CUDA(c, cudaMalloc(p, bytes + 256));
    *have = bytes;
    return HF_OK;
}

double now_ms()
{
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

static uint32_t preamble_bytes(uint64_t n) { return 3 + (uint32_t)(n & 1); }

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

const char *hf_version(void) { return "huffman_b200 0.1 (sm_100a)"; }

int hf_ctx_create(hf_ctx **out, int device, void *stream)
{
    if (!out) return HF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return HF_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return HF_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HF_ERR_CUDA;
    if (prop.major < 10) return HF_ERR_CUDA;            // sm_100a code only
    Ctx *c = (Ctx *)calloc(1, sizeof(Ctx));
    if (!c) return HF_ERR_ARG;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->stream = (cudaStream_t)stream;                   // NULL = the legacy default stream, as the reference uses
    c->own_stream = false;
    c->nranks = 1;
    {   // development aids for A/B timing: HF_WRITE_KERNEL=3 / 4 gives every chunk to one write kernel, HF_WRITE_SPLIT moves the split
        const char *e = getenv("HF_WRITE_KERNEL");
        c->write_kernel = e ? atoi(e) : 0;
        e = getenv("HF_WRITE_SPLIT");
        c->write_split = e ? (uint32_t)atoi(e) : WRITE_SPLIT_DEFAULT;
    }
    bool ok = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&c->h_pipe, PIPE_SLOTS * 8) == cudaSuccess;
    for (int i = 0; ok && i < 8; i++) ok = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_scratch, 4096) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_cb, codebook_alloc_bytes()) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_tab, sizeof(DecodeTable)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_hist, NSYM * 8 + 256) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_scan, SCAN_BLOCKS_MAX * 8) == cudaSuccess;
    if (!ok) { hf_ctx_destroy(reinterpret_cast<hf_ctx *>(c)); return HF_ERR_CUDA; }
    *out = reinterpret_cast<hf_ctx *>(c);
    return HF_OK;
}

int hf_ctx_destroy(hf_ctx *ctx)
{
    if (!ctx) return HF_OK;
    Ctx *c = CTX(ctx);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->d2h_stream) { cudaStreamSynchronize(c->d2h_stream); cudaStreamDestroy(c->d2h_stream); }
    if (c->h_pipe) cudaFreeHost(c->h_pipe);
    for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    shard_release(c);
    ring_release(c);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->ws) cudaFree(c->ws);
    if (c->d_in) cudaFree(c->d_in);
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_cb) cudaFree(c->d_cb);
    if (c->d_tab) cudaFree(c->d_tab);
    if (c->d_hist) cudaFree(c->d_hist);
    if (c->d_scan) cudaFree(c->d_scan);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    if (c->prof_ev) {
        for (uint32_t i = 0; i < 2 * PROF_CAP; i++) if (c->prof_ev[i]) cudaEventDestroy(c->prof_ev[i]);
        free(c->prof_ev);
        free(c->prof_name);
    }
    free(c);
    return HF_OK;
}

int hf_ctx_set_stream(hf_ctx *ctx, void *stream)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)stream;
    return HF_OK;
}

int hf_sync(hf_ctx *ctx)
{
    NEED_CTX(ctx);
    HF_CUDA(CTX(ctx), cudaStreamSynchronize(CTX(ctx)->stream));
    return HF_OK;
}

const char *hf_last_error(hf_ctx *ctx) { return ctx ? CTX(ctx)->err : "no context"; }
uint64_t hf_launch_count(hf_ctx *ctx) { return ctx ? CTX(ctx)->launches : 0; }

int hf_profile_enable(hf_ctx *ctx, int on)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (on && !c->prof_ev) {
        c->prof_ev = (cudaEvent_t *)calloc(2 * PROF_CAP, sizeof(cudaEvent_t));
        c->prof_name = (const char **)calloc(PROF_CAP, sizeof(char *));
        if (!c->prof_ev || !c->prof_name) return set_err(c, HF_ERR_ARG, "hf_profile_enable: out of memory");
        for (uint32_t i = 0; i < 2 * PROF_CAP; i++) HF_CUDA(c, cudaEventCreate(&c->prof_ev[i]));
    }
    c->prof_on = on != 0;
    c->prof_major_only = on == 2;
    c->prof_open = false;
    c->prof_n = 0;
    return HF_OK;
}

int hf_profile_read(hf_ctx *ctx, hf_kernel_time_t *out, uint32_t cap, uint32_t *n_out)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!n_out || (cap && !out)) return set_err(c, HF_ERR_ARG, "hf_profile_read: null pointer");
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t n = 0;
    for (uint32_t i = 0; i < c->prof_n; i++) {
        float ms = 0.f;
        HF_CUDA(c, cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        uint32_t k = 0;
        while (k < n && strncmp(out[k].name, c->prof_name[i], sizeof(out[k].name) - 1) != 0) k++;
        if (k == n) {
            if (n >= cap) continue;
            memset(&out[k], 0, sizeof(out[k]));
            strncpy(out[k].name, c->prof_name[i], sizeof(out[k].name) - 1);
            n++;
        }
        out[k].launches++;
        out[k].total_ms += ms;
    }
    *n_out = n;
    c->prof_n = 0;
    return HF_OK;
}

int hf_host_alloc(void **h_ptr, size_t bytes)
{
    if (!h_ptr) return HF_ERR_ARG;
    return cudaMallocHost(h_ptr, bytes ? bytes : 1) == cudaSuccess ? HF_OK : HF_ERR_CUDA;
}
int hf_host_free(void *h_ptr) { return cudaFreeHost(h_ptr) == cudaSuccess ? HF_OK : HF_ERR_CUDA; }

size_t hf_codebook_bytes(void) { return codebook_alloc_bytes(); }
size_t hf_decode_table_bytes(void) { return sizeof(DecodeTable); }

uint64_t hf_compress_bound(uint64_t n)
{   // header: <= min(65536, n/2) entries of <= 11 bytes; payload: an optimal prefix code never
    // exceeds the 16-bit fixed-length code, i.e. n bytes
    uint64_t u = n / 2 < NSYM ? n / 2 : NSYM;
    return 4 + 11 * u + 8 + n + 32;
}

// ---- compress stages ---------------------------------------------------------------
int hf_histogram(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, uint64_t *d_hist)
{
    NEED_CTX(ctx);
    if (!d_hist || (!d_in && n_bytes)) return set_err(CTX(ctx), HF_ERR_ARG, "hf_histogram: null pointer");
    return launch_histogram(CTX(ctx), d_in, n_bytes, reinterpret_cast<unsigned long long *>(d_hist));
}

int hf_build_codebook(hf_ctx *ctx, const uint64_t *d_hist, void *d_codebook)
{
    NEED_CTX(ctx);
    if (!d_hist || !d_codebook) return set_err(CTX(ctx), HF_ERR_ARG, "hf_build_codebook: null pointer");
    return launch_codebook(CTX(ctx), reinterpret_cast<const unsigned long long *>(d_hist),
                           reinterpret_cast<Codebook *>(d_codebook));
}

int hf_codebook_info(hf_ctx *ctx, const void *d_codebook, hf_cb_info_t *h_info)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_codebook || !h_info) return set_err(c, HF_ERR_ARG, "hf_codebook_info: null pointer");
    Codebook *h = reinterpret_cast<Codebook *>(c->h_scratch);       // only the 48-byte summary is copied
    HF_CUDA(c, cudaMemcpyAsync(h, d_codebook, offsetof(Codebook, order), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    h_info->n_unique = h->U;
    h_info->max_code_bits = h->maxlen;
    h_info->table_bits = h->table_bits;
    h_info->payload_bits = h->payload_bits;
    h_info->status = h->status;
    h_info->reserved = 0;
    return HF_OK;
}

int hf_codebook_export(hf_ctx *ctx, const void *d_codebook, uint16_t *h_order, uint8_t *h_len, uint64_t *h_code)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    const Codebook *cb = reinterpret_cast<const Codebook *>(d_codebook);
    if (!cb) return set_err(c, HF_ERR_ARG, "hf_codebook_export: null pointer");
    if (h_order) HF_CUDA(c, cudaMemcpyAsync(h_order, cb->order, sizeof(cb->order), cudaMemcpyDeviceToHost, c->stream));
    if (h_len) HF_CUDA(c, cudaMemcpyAsync(h_len, cb->len, sizeof(cb->len), cudaMemcpyDeviceToHost, c->stream));
    if (h_code) HF_CUDA(c, cudaMemcpyAsync(h_code, cb->code, sizeof(cb->code), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HF_OK;
}

int hf_shard_payload_bits(hf_ctx *ctx, const uint64_t *d_shard_hist, const void *d_codebook, uint64_t *d_bits)
{
    NEED_CTX(ctx);
    if (!d_shard_hist || !d_codebook || !d_bits) return set_err(CTX(ctx), HF_ERR_ARG, "hf_shard_payload_bits: null pointer");
    return launch_shard_bits(CTX(ctx), reinterpret_cast<const unsigned long long *>(d_shard_hist),
                             reinterpret_cast<const Codebook *>(d_codebook),
                             reinterpret_cast<unsigned long long *>(d_bits));
}

int hf_header_pack(hf_ctx *ctx, const void *d_codebook, uint64_t n_bytes, uint32_t last_byte, uint8_t *d_file,
                   uint64_t capacity)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_codebook || !d_file) return set_err(c, HF_ERR_ARG, "hf_header_pack: null pointer");
    uint64_t u = n_bytes / 2 < NSYM ? n_bytes / 2 : NSYM;
    if (capacity < 4 + 11 * u + 8 + 4)
        return set_err(c, HF_ERR_CAPACITY, "hf_header_pack: capacity %llu below the header bound %llu",
                       (unsigned long long)capacity, (unsigned long long)(4 + 11 * u + 12));
    return launch_header_pack(c, reinterpret_cast<const Codebook *>(d_codebook), n_bytes, last_byte, nullptr, d_file, capacity, nullptr);
}

int hf_encode(hf_ctx *ctx, const uint8_t *d_in, uint64_t n_bytes, const void *d_codebook, uint8_t *d_stream,
              uint64_t start_bit)
{
    NEED_CTX(ctx);
    if ((!d_in && n_bytes) || !d_codebook || !d_stream) return set_err(CTX(ctx), HF_ERR_ARG, "hf_encode: null pointer");
    return launch_encode(CTX(ctx), d_in, n_bytes, reinterpret_cast<const Codebook *>(d_codebook), d_stream, start_bit, nullptr);
}

// the side index object: this header, then n_subs u16 records (encode2.cu enc_index_kernel)
struct IndexHeader {
    uint32_t magic, version;            // "HFIX", 1
    uint32_t align16;                   // address of the image modulo 16 the records were made for
    uint32_t reserved;
    unsigned long long payload_start_bit, n_subs, file_bytes, original_bytes;
    unsigned long long pad[2];
};
static_assert(sizeof(IndexHeader) == 64, "index header is 64 bytes");
static const uint32_t INDEX_MAGIC = 0x58494648u;

// hf_compress_indexed: codebook -> sizes -> header -> payload -> side index.  The index needs the stream's geometry on the
// host, so this variant learns the sizes before it packs (hf_compress does not: compress_async below).
// *h_last is read after the synchronisation inside hf_codebook_info.
static int compress_after_hist(Ctx *c, const uint8_t *d_in, uint64_t n, const uint8_t *h_last, uint8_t *d_file,
                               uint64_t capacity, uint64_t *h_file_bytes, uint8_t *d_index = nullptr,
                               uint64_t index_capacity = 0, uint64_t *h_index_bytes = nullptr)
{
    hf_ctx *ctx = reinterpret_cast<hf_ctx *>(c);
    Codebook *cb = reinterpret_cast<Codebook *>(c->d_cb);
    int rc = launch_codebook(c, reinterpret_cast<unsigned long long *>(c->d_hist), cb);
    if (rc) return rc;
    hf_cb_info_t info;
    rc = hf_codebook_info(ctx, cb, &info);
    if (rc) return rc;
    if (info.status) return set_err(c, (int)info.status, "codebook: a code word is longer than 64 bits");
    const uint64_t bits = info.table_bits + 64 + info.payload_bits;
    const uint64_t total = preamble_bytes(n) + (bits + 7) / 8;
    if (h_file_bytes) *h_file_bytes = total;
    if (total > capacity)
        return set_err(c, HF_ERR_CAPACITY, "hf_compress: need %llu bytes, capacity %llu", (unsigned long long)total,
                       (unsigned long long)capacity);
    // the image fits (total <= capacity): the worst-case header bound of the standalone stage call does not apply
    rc = launch_header_pack(c, cb, n, (n & 1) ? *h_last : 0, nullptr, d_file, capacity, nullptr);
    if (rc) return rc;
    rc = launch_encode(c, d_in, n, cb, d_file + preamble_bytes(n), info.table_bits + 64, nullptr);
    if (rc || !d_index) return rc;
    // the side index: records for exactly the stream just packed, at this image's alignment
    if (h_index_bytes) *h_index_bytes = 0;
    if (info.max_code_bits == 0 || n < 2) return HF_OK;             // no payload: nothing to index
    const uint8_t *d_stream = d_file + preamble_bytes(n);
    const uint64_t stream_bytes = total - preamble_bytes(n), start_bit = info.table_bits + 64;
    const uint64_t n_subs = index_subs(d_stream, stream_bytes, start_bit);
    const uint64_t need = sizeof(IndexHeader) + 2 * n_subs;
    if (need > index_capacity)
        return set_err(c, HF_ERR_CAPACITY, "hf_compress_indexed: index needs %llu bytes, capacity %llu",
                       (unsigned long long)need, (unsigned long long)index_capacity);
    if ((uintptr_t)d_index & 15) return set_err(c, HF_ERR_ARG, "hf_compress_indexed: index buffer must be 16-byte aligned");
    rc = launch_encode_index(c, d_in, n, cb, d_stream, start_bit, reinterpret_cast<uint16_t *>(d_index + sizeof(IndexHeader)), n_subs);
    if (rc) return rc;
    IndexHeader *h = reinterpret_cast<IndexHeader *>((uint8_t *)c->h_scratch + 3584);
    memset(h, 0, sizeof(*h));
    h->magic = INDEX_MAGIC; h->version = 1; h->align16 = (uint32_t)((uintptr_t)d_file & 15);
    h->payload_start_bit = preamble_bytes(n) * 8ull + start_bit; h->n_subs = n_subs; h->file_bytes = total; h->original_bytes = n;
    HF_CUDA(c, cudaMemcpyAsync(d_index, h, sizeof(*h), cudaMemcpyHostToDevice, c->stream));
    if (h_index_bytes) *h_index_bytes = need;
    return HF_OK;
}

// hf_compress / hf_compress_host after the histogram: codebook -> plan (sizes, start bit, capacity check: all on the
// device) -> header -> payload; nothing waits for the host.  d_last / last_byte: see launch_header_pack.
extern "C++" int hf::compress_async(Ctx *c, const uint8_t *d_in, uint64_t n, const uint8_t *d_last, uint32_t last_byte, uint8_t *d_file,
                          uint64_t capacity, ShardPlan **d_plan)
{
    Codebook *cb = reinterpret_cast<Codebook *>(c->d_cb);
    int rc = launch_codebook(c, reinterpret_cast<unsigned long long *>(c->d_hist), cb);
    if (rc) return rc;
    rc = launch_plan_single(c, cb, n, capacity, d_plan);
    if (rc) return rc;
    rc = launch_header_pack(c, cb, n, last_byte, (n & 1) ? d_last : nullptr, d_file, capacity, *d_plan);
    if (rc) return rc;
    return launch_encode(c, d_in, n, cb, d_file, 0, *d_plan);
}

// the plan after the stream has drained: the image size, or why there is no image
extern "C++" int hf::compress_result(Ctx *c, const ShardPlan *d_plan, uint64_t capacity, uint64_t *h_file_bytes)
{
    ShardPlan *h = reinterpret_cast<ShardPlan *>((uint8_t *)c->h_scratch + 2560);
    HF_CUDA(c, cudaMemcpyAsync(h, d_plan, sizeof(ShardPlan), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h_file_bytes) *h_file_bytes = h->image_bytes;
    if (h->status == HF_ERR_CAPACITY)
        return set_err(c, HF_ERR_CAPACITY, "hf_compress: need %llu bytes, capacity %llu", (unsigned long long)h->image_bytes,
                       (unsigned long long)capacity);
    if (h->status) return set_err(c, (int)h->status, "codebook: a code word is longer than 64 bits");
    return HF_OK;
}

int hf_compress(hf_ctx *ctx, const uint8_t *d_in, uint64_t n, uint8_t *d_file, uint64_t capacity,
                uint64_t *h_file_bytes)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if ((!d_in && n) || !d_file) return set_err(c, HF_ERR_ARG, "hf_compress: null pointer");
    if (capacity < 12) return set_err(c, HF_ERR_CAPACITY, "hf_compress: capacity %llu is below any image", (unsigned long long)capacity);
    HF_CUDA(c, cudaMemsetAsync(c->d_hist, 0, NSYM * 8, c->stream));
    int rc = launch_histogram(c, d_in, n, reinterpret_cast<unsigned long long *>(c->d_hist));
    if (rc) return rc;
    ShardPlan *d_plan = nullptr;
    rc = compress_async(c, d_in, n, n ? d_in + n - 1 : nullptr, 0, d_file, capacity, &d_plan);
    if (rc) return rc;
    return compress_result(c, d_plan, capacity, h_file_bytes);
}

uint64_t hf_index_bound(uint64_t n)
{   // records for the largest image n bytes can give, at any alignment
    const uint64_t frame = hf_compress_bound(n) + 16;
    const uint64_t nch = (frame * 8 + 131071) / 131072 + 1;
    return sizeof(IndexHeader) + 2 * nch * 512;
}

int hf_compress_indexed(hf_ctx *ctx, const uint8_t *d_in, uint64_t n, uint8_t *d_file, uint64_t capacity,
                        uint64_t *h_file_bytes, uint8_t *d_index, uint64_t index_capacity, uint64_t *h_index_bytes)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if ((!d_in && n) || !d_file || !d_index) return set_err(c, HF_ERR_ARG, "hf_compress_indexed: null pointer");
    HF_CUDA(c, cudaMemsetAsync(c->d_hist, 0, NSYM * 8, c->stream));
    int rc = launch_histogram(c, d_in, n, reinterpret_cast<unsigned long long *>(c->d_hist));
    if (rc) return rc;
    uint8_t *h_last = reinterpret_cast<uint8_t *>(c->h_scratch) + 2048;
    *h_last = 0;
    if (n & 1) HF_CUDA(c, cudaMemcpyAsync(h_last, d_in + n - 1, 1, cudaMemcpyDeviceToHost, c->stream));
    rc = compress_after_hist(c, d_in, n, h_last, d_file, capacity, h_file_bytes, d_index, index_capacity, h_index_bytes);
    if (rc) return rc;
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HF_OK;
}

// ---- decompress stages -------------------------------------------------------------
int hf_parse_header(hf_ctx *ctx, const uint8_t *d_file, uint64_t file_bytes, void *d_decode_table,
                    hf_header_info_t *h_info)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_file || !d_decode_table || !h_info) return set_err(c, HF_ERR_ARG, "hf_parse_header: null pointer");
    if (file_bytes < 11) return set_err(c, HF_ERR_FORMAT, "hf_parse_header: %llu bytes is shorter than any image",
                                        (unsigned long long)file_bytes);
    // device-side info block lives at the end of the hist buffer's allocation
    hf_header_info_t *d_info = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->d_hist + NSYM * 8);
    DecodeTable *tab = reinterpret_cast<DecodeTable *>(d_decode_table);
    int rc = launch_parse_header(c, d_file, file_bytes, tab, d_info);
    if (rc) return rc;
    hf_header_info_t *h = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->h_scratch + 1024);
    uint32_t *h_tab = reinterpret_cast<uint32_t *>((uint8_t *)c->h_scratch + 1536);
    HF_CUDA(c, cudaMemcpyAsync(h, d_info, sizeof(hf_header_info_t), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(h_tab, tab, 48, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    *h_info = *h;
    h_info->max_code_bits = h_tab[1];
    if (h_tab[6] != 0 && h_info->status == HF_OK) h_info->status = h_tab[6];
    if (h_info->status) return set_err(c, HF_ERR_FORMAT, "hf_parse_header: malformed header");
    return HF_OK;
}

int hf_decode_table_from_codebook(hf_ctx *ctx, const void *d_codebook, void *d_decode_table)
{
    NEED_CTX(ctx);
    if (!d_codebook || !d_decode_table) return set_err(CTX(ctx), HF_ERR_ARG, "hf_decode_table_from_codebook: null pointer");
    return launch_table_from_codebook(CTX(ctx), reinterpret_cast<const Codebook *>(d_codebook),
                                      reinterpret_cast<DecodeTable *>(d_decode_table));
}

int hf_decode(hf_ctx *ctx, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit, uint64_t n_symbols,
              const void *d_decode_table, uint8_t *d_out)
{
    NEED_CTX(ctx);
    if (!d_decode_table || (n_symbols && (!d_stream || !d_out))) return set_err(CTX(ctx), HF_ERR_ARG, "hf_decode: null pointer");
    return launch_decode(CTX(ctx), d_stream, stream_bytes, start_bit, n_symbols,
                         reinterpret_cast<const DecodeTable *>(d_decode_table), d_out);
}

int hf_decode_range(hf_ctx *ctx, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes, uint64_t first_bit,
                    const void *d_decode_table, uint8_t *d_out, uint64_t out_symbols, uint64_t *d_result)
{
    NEED_CTX(ctx);
    if (!d_range || !d_decode_table || !d_out || !d_result) return set_err(CTX(ctx), HF_ERR_ARG, "hf_decode_range: null pointer");
    return launch_decode_range(CTX(ctx), d_range, range_bytes, halo_bytes, first_bit, false,
                               reinterpret_cast<const DecodeTable *>(d_decode_table), d_out, out_symbols,
                               reinterpret_cast<unsigned long long *>(d_result));
}

int hf_range_overflow(hf_ctx *ctx, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes,
                      const void *d_decode_table, uint64_t *d_result)
{
    NEED_CTX(ctx);
    if (!d_range || !d_decode_table || !d_result) return set_err(CTX(ctx), HF_ERR_ARG, "hf_range_overflow: null pointer");
    return launch_decode_range(CTX(ctx), d_range, range_bytes, halo_bytes, 0, true,
                               reinterpret_cast<const DecodeTable *>(d_decode_table), nullptr, 0,
                               reinterpret_cast<unsigned long long *>(d_result));
}

// what the decode kernels report (DecWork, decode_common.cuh) sits at the start of the workspace's stage region
extern "C++" int hf::check_decode_flags(Ctx *c)
{
    unsigned long long *h = reinterpret_cast<unsigned long long *>((uint8_t *)c->h_scratch + 3072);
    HF_CUDA(c, cudaMemcpyAsync(h, (uint8_t *)c->ws + WS_STAGE_OFFSET, 64, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h[4 + 1]) return set_err(c, HF_ERR_FORMAT, "hf_decode: the payload holds bits that are no code word");
    return HF_OK;
}

int hf_decompress(hf_ctx *ctx, const uint8_t *d_file, uint64_t file_bytes, uint8_t *d_out, uint64_t capacity,
                  uint64_t *h_out_bytes)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_file) return set_err(c, HF_ERR_ARG, "hf_decompress: null pointer");
    if (file_bytes < 11) return set_err(c, HF_ERR_FORMAT, "hf_decompress: %llu bytes is shorter than any image",
                                        (unsigned long long)file_bytes);
    if (!d_out) capacity = 0;
    hf_header_info_t *d_info = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->d_hist + NSYM * 8);
    DecodeTable *tab = reinterpret_cast<DecodeTable *>(c->d_tab);
    int rc = launch_parse_header(c, d_file, file_bytes, tab, d_info);
    if (rc) return rc;
    rc = launch_decompress_image(c, d_file, file_bytes, d_info, tab, d_out, capacity);
    if (rc) return rc;
    // the one synchronisation: what the header said, what the kernels flagged
    hf_header_info_t *h = reinterpret_cast<hf_header_info_t *>((uint8_t *)c->h_scratch + 1024);
    uint32_t *h_tab = reinterpret_cast<uint32_t *>((uint8_t *)c->h_scratch + 1536);
    unsigned long long *h_work = reinterpret_cast<unsigned long long *>((uint8_t *)c->h_scratch + 3072);
    HF_CUDA(c, cudaMemcpyAsync(h, d_info, sizeof(hf_header_info_t), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(h_tab, tab, 48, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaMemcpyAsync(h_work, (uint8_t *)c->ws + WS_STAGE_OFFSET, 64, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h->status || h_tab[6]) return set_err(c, HF_ERR_FORMAT, "hf_decompress: malformed header");
    if (h_out_bytes) *h_out_bytes = h->original_bytes;
    if (h_work[4 + 2])
        return set_err(c, d_out ? HF_ERR_CAPACITY : HF_ERR_ARG, "hf_decompress: need %llu bytes, capacity %llu",
                       (unsigned long long)h->original_bytes, (unsigned long long)capacity);
    if (h_work[4 + 1]) return set_err(c, HF_ERR_FORMAT, "hf_decode: the payload holds bits that are no code word");
    return HF_OK;
}

int hf_decompress_indexed(hf_ctx *ctx, const uint8_t *d_file, uint64_t file_bytes, const uint8_t *d_index,
                          uint64_t index_bytes, uint8_t *d_out, uint64_t capacity, uint64_t *h_out_bytes)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!d_index || index_bytes < sizeof(IndexHeader)) return hf_decompress(ctx, d_file, file_bytes, d_out, capacity, h_out_bytes);
    hf_header_info_t info;
    int rc = hf_parse_header(ctx, d_file, file_bytes, c->d_tab, &info);
    if (rc) return rc;
    if (h_out_bytes) *h_out_bytes = info.original_bytes;
    if (info.original_bytes > capacity)
        return set_err(c, HF_ERR_CAPACITY, "hf_decompress_indexed: need %llu bytes, capacity %llu",
                       (unsigned long long)info.original_bytes, (unsigned long long)capacity);
    if (info.original_bytes && !d_out) return set_err(c, HF_ERR_ARG, "hf_decompress_indexed: null output");
    const uint64_t nsym = info.original_bytes / 2;
    IndexHeader *h = reinterpret_cast<IndexHeader *>((uint8_t *)c->h_scratch + 3584);
    HF_CUDA(c, cudaMemcpyAsync(h, d_index, sizeof(*h), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    DecodeTable *tab = reinterpret_cast<DecodeTable *>(c->d_tab);
    // the index must be the one of THIS image at THIS alignment; anything else decodes without it
    bool usable = nsym != 0 && info.max_code_bits != 0 && h->magic == INDEX_MAGIC && h->version == 1 &&
                  h->align16 == (uint32_t)((uintptr_t)d_file & 15) && h->payload_start_bit == info.payload_start_bit &&
                  h->file_bytes == file_bytes && h->original_bytes == info.original_bytes &&
                  ((uintptr_t)d_index & 15) == 0 && index_bytes >= sizeof(IndexHeader) + 2 * h->n_subs &&
                  h->n_subs == index_subs(d_file, file_bytes, info.payload_start_bit);
    if (usable) {
        rc = launch_decode_indexed(c, d_file, file_bytes, info.payload_start_bit, nsym, tab, d_out,
                                   reinterpret_cast<const uint16_t *>(d_index + sizeof(IndexHeader)), h->n_subs);
        if (rc == HF_OK) rc = check_decode_flags(c);
        usable = rc == HF_OK;                           // a record that does not match the walk: decode again without
    }
    if (!usable && nsym) {
        rc = launch_decode(c, d_file, file_bytes, info.payload_start_bit, nsym, tab, d_out);
        if (rc) return rc;
        rc = check_decode_flags(c);
        if (rc) return rc;
    }
    if (info.is_odd) HF_CUDA(c, cudaMemsetAsync(d_out + info.original_bytes - 1, (int)info.last_byte, 1, c->stream));   // D:286-289
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HF_OK;
}

// ---- host-buffer calls ---------------------------------------------------------------
static const uint64_t H2D_CHUNK = 32ull << 20;

int hf_compress_host(hf_ctx *ctx, const uint8_t *h_in, uint64_t n, uint8_t *h_file, uint64_t capacity,
                     uint64_t *h_file_bytes)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if ((!h_in && n) || !h_file) return set_err(c, HF_ERR_ARG, "hf_compress_host: null pointer");
    const uint64_t bound = hf_compress_bound(n);
    int rc = ensure_buf(c, &c->d_in, &c->d_in_bytes, n + 16);
    if (rc) return rc;
    rc = ensure_buf(c, &c->d_out, &c->d_out_bytes, bound + 16);
    if (rc) return rc;
    uint8_t *d_in = reinterpret_cast<uint8_t *>(c->d_in);
    // the image starts 16 - preamble bytes into the buffer so that the bit stream is 16-byte aligned
    uint8_t *d_file = reinterpret_cast<uint8_t *>(c->d_out) + (16 - preamble_bytes(n));
    unsigned long long *d_hist = reinterpret_cast<unsigned long long *>(c->d_hist);

    HF_CUDA(c, cudaMemsetAsync(d_hist, 0, NSYM * 8, c->stream));
    // copy chunk k on the copy stream while the histogram of chunk k-1 runs
    HF_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    HF_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev[0], 0));
    int e = 0;
    for (uint64_t o = 0; o < n; o += H2D_CHUNK) {
        uint64_t len = n - o < H2D_CHUNK ? n - o : H2D_CHUNK;
        HF_CUDA(c, cudaMemcpyAsync(d_in + o, h_in + o, len, cudaMemcpyHostToDevice, c->copy_stream));
        cudaEvent_t ev = c->ev[1 + (e++ % 6)];
        HF_CUDA(c, cudaEventRecord(ev, c->copy_stream));
        HF_CUDA(c, cudaStreamWaitEvent(c->stream, ev, 0));
        rc = launch_histogram(c, d_in + o, len & ~1ull, d_hist);     // chunks are even-sized; the odd last byte is not a symbol
        if (rc) return rc;
    }
    uint64_t total = 0;
    ShardPlan *d_plan = nullptr;
    rc = compress_async(c, d_in, n, nullptr, (n & 1) ? h_in[n - 1] : 0u, d_file, bound, &d_plan);
    if (rc) return rc;
    rc = compress_result(c, d_plan, bound, &total);
    if (h_file_bytes) *h_file_bytes = total;
    if (rc) return rc;
    if (total > capacity)
        return set_err(c, HF_ERR_CAPACITY, "hf_compress_host: need %llu bytes, capacity %llu", (unsigned long long)total,
                       (unsigned long long)capacity);
    HF_CUDA(c, cudaMemcpyAsync(h_file, d_file, total, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HF_OK;
}

int hf_decompressed_size_host(const uint8_t *h_file, uint64_t file_bytes, uint64_t *h_out_bytes)
{
    // walks the header on the host just far enough to find the 64-bit size (D:68-103, D:243-255);
    // used by callers to size the output buffer before the GPU call
    if (!h_file || !h_out_bytes || file_bytes < 11) return HF_ERR_FORMAT;
    uint32_t U = (uint32_t)h_file[0] | ((uint32_t)h_file[1] << 8);
    uint32_t pre = 3 + (h_file[2] != 0);
    if (U == 0) U = (file_bytes - pre == 8) ? 0 : 65536;
    const uint8_t *s = h_file + pre;
    uint64_t nb = file_bytes - pre, pos = 0;
    auto byte_at = [&](uint64_t bit) -> uint32_t {
        uint64_t i = bit >> 3;
        uint32_t sh = (uint32_t)(bit & 7);
        uint32_t a = i < nb ? s[i] : 0, b = i + 1 < nb ? s[i + 1] : 0;
        return (((a << 8) | b) >> (8 - sh)) & 0xFF;
    };
    for (uint32_t k = 0; k < U; k++) {
        if ((pos >> 3) + 3 > nb) return HF_ERR_FORMAT;
        uint32_t len = byte_at(pos + 16);
        if (len == 0 && U > 1) return HF_ERR_FORMAT;
        pos += 24 + len;
    }
    if ((pos + 64 + 7) / 8 > nb) return HF_ERR_FORMAT;
    uint64_t n = 0;
    for (int i = 0; i < 8; i++) n |= (uint64_t)byte_at(pos + 8 * i) << (8 * i);
    *h_out_bytes = n;
    return HF_OK;
}

// Host buffers in, host buffers out.  Large images are pipelined: the header is parsed from the first MiB while the
// payload is still crossing PCIe in slices; each slice is decoded as it lands and its symbols leave on a third stream,
// so the H2D copy of the image, the kernels and the D2H copy of the output overlap (the reference reads the file with
// fread per byte and decodes on the host, D:259-291).
static const uint64_t PIPE_MIN_BYTES = 64ull << 20;
static const uint64_t PIPE_SLICE_CHUNKS = 16384;        // 256 MiB of payload per slice

int hf_decompress_host(hf_ctx *ctx, const uint8_t *h_file, uint64_t file_bytes, uint8_t *h_out, uint64_t capacity,
                       uint64_t *h_out_bytes)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    if (!h_file) return set_err(c, HF_ERR_ARG, "hf_decompress_host: null pointer");
    if (file_bytes < 11) return set_err(c, HF_ERR_FORMAT, "hf_decompress_host: image too short");
    int rc = ensure_buf(c, &c->d_out, &c->d_out_bytes, file_bytes + 128);
    if (rc) return rc;
    uint8_t *d_file = reinterpret_cast<uint8_t *>(c->d_out);
    // the header (at most ~720 KiB) first
    const bool piped = file_bytes >= PIPE_MIN_BYTES;
    const uint64_t first = piped ? (1ull << 20) : file_bytes;
    HF_CUDA(c, cudaMemcpyAsync(d_file, h_file, first, cudaMemcpyHostToDevice, c->stream));
    hf_header_info_t info;
    rc = hf_parse_header(ctx, d_file, file_bytes, c->d_tab, &info);     // synchronises
    if (rc) return rc;
    if (h_out_bytes) *h_out_bytes = info.original_bytes;
    if (info.original_bytes > capacity)
        return set_err(c, HF_ERR_CAPACITY, "hf_decompress_host: need %llu bytes, capacity %llu",
                       (unsigned long long)info.original_bytes, (unsigned long long)capacity);
    if (info.original_bytes && !h_out) return set_err(c, HF_ERR_ARG, "hf_decompress_host: null output");
    rc = ensure_buf(c, &c->d_in, &c->d_in_bytes, info.original_bytes + 16);
    if (rc) return rc;
    uint8_t *d_out = reinterpret_cast<uint8_t *>(c->d_in);
    const uint64_t nsym = info.original_bytes / 2;
    DecodeTable *tab = reinterpret_cast<DecodeTable *>(c->d_tab);

    if (!piped || nsym == 0 || info.n_unique <= 1) {
        if (first < file_bytes)
            HF_CUDA(c, cudaMemcpyAsync(d_file + first, h_file + first, file_bytes - first, cudaMemcpyHostToDevice, c->stream));
        if (nsym) {
            rc = launch_decode(c, d_file, file_bytes, info.payload_start_bit, nsym, tab, d_out);
            if (rc) return rc;
            HF_CUDA(c, cudaMemcpyAsync(h_out, d_out, nsym * 2, cudaMemcpyDeviceToHost, c->stream));
            rc = check_decode_flags(c);
            if (rc) return rc;
        } else {
            HF_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        if (info.is_odd) h_out[info.original_bytes - 1] = (uint8_t)info.last_byte;      // D:286-289
        return HF_OK;
    }

    DecodeJob job{};
    rc = decode_begin(c, d_file, file_bytes, info.payload_start_bit, nsym, tab, d_out, &job);
    if (rc) return rc;
    const uint64_t frame_off = (uint64_t)(job.frame - d_file);          // file byte of frame byte 0
    const uint64_t slice_bytes = PIPE_SLICE_CHUNKS * 16384ull;
    const uint64_t nsl = (job.nch + PIPE_SLICE_CHUNKS - 1) / PIPE_SLICE_CHUNKS;
    if (nsl > PIPE_SLOTS) return set_err(c, HF_ERR_ARG, "hf_decompress_host: image too large");
    cudaEvent_t *ev = (cudaEvent_t *)calloc(2 * nsl, sizeof(cudaEvent_t));
    if (!ev) return set_err(c, HF_ERR_ARG, "hf_decompress_host: out of memory");
    auto cleanup = [&](int code) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->d2h_stream);
        for (uint64_t i = 0; i < 2 * nsl; i++) if (ev[i]) cudaEventDestroy(ev[i]);
        free(ev);
        return code;
    };
#define PIPE_CUDA(call)                                                                                     \
    do {                                                                                                    \
        cudaError_t _e = (call);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return cleanup(set_err(c, HF_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e))); \
    } while (0)
    for (uint64_t i = 0; i < 2 * nsl; i++) PIPE_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    // the payload crosses PCIe slice by slice (each with 64 bytes of the next: the kernels look a few bytes ahead)
    for (uint64_t k = 0; k < nsl; k++) {
        uint64_t lo = frame_off + k * slice_bytes, hi = frame_off + (k + 1) * slice_bytes + 64;
        if (lo < first) lo = first;
        if (hi > file_bytes) hi = file_bytes;
        if (hi > lo) PIPE_CUDA(cudaMemcpyAsync(d_file + lo, h_file + lo, hi - lo, cudaMemcpyHostToDevice, c->copy_stream));
        PIPE_CUDA(cudaEventRecord(ev[2 * k], c->copy_stream));
    }
    uint64_t sent = 0;                                  // symbols already on their way to the host
    auto drain = [&](uint64_t k) -> int {               // the symbols slice k completed leave on the third stream
        cudaError_t e = cudaEventSynchronize(ev[2 * k + 1]);
        if (e != cudaSuccess) return set_err(c, HF_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
        uint64_t tot = c->h_pipe[k] < nsym ? c->h_pipe[k] : nsym;
        if (k + 1 == nsl) tot = nsym;
        if (tot > sent) {
            e = cudaStreamWaitEvent(c->d2h_stream, ev[2 * k + 1], 0);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(h_out + 2 * sent, d_out + 2 * sent, 2 * (tot - sent), cudaMemcpyDeviceToHost, c->d2h_stream);
            if (e != cudaSuccess) return set_err(c, HF_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
            sent = tot;
        }
        return HF_OK;
    };
    for (uint64_t k = 0; k < nsl; k++) {
        PIPE_CUDA(cudaStreamWaitEvent(c->stream, ev[2 * k], 0));
        const uint64_t c0 = k * PIPE_SLICE_CHUNKS, c1 = (k + 1) * PIPE_SLICE_CHUNKS < job.nch ? (k + 1) * PIPE_SLICE_CHUNKS : job.nch;
        rc = decode_slice(c, job, c0, c1);
        if (rc) return cleanup(rc);
        PIPE_CUDA(cudaMemcpyAsync(&c->h_pipe[k], job.total, 8, cudaMemcpyDeviceToHost, c->stream));
        PIPE_CUDA(cudaEventRecord(ev[2 * k + 1], c->stream));
        if (k) { rc = drain(k - 1); if (rc) return cleanup(rc); }
    }
    rc = drain(nsl - 1);
    if (rc) return cleanup(rc);
#undef PIPE_CUDA
    rc = cleanup(HF_OK);
    if (rc) return rc;
    rc = check_decode_flags(c);
    if (rc) return rc;
    if (info.is_odd) h_out[info.original_bytes - 1] = (uint8_t)info.last_byte;          // D:286-289
    return HF_OK;
}

}  // extern "C"
