// programs.cu — the two program-level entry points that mirror the reference's mains
// (/root/reference/Compressor.cu:315-632 `archive`, /root/reference/Decompressor.cu:47-114 `extract`): same
// command line, same output file names, the reference's stdout lines in the reference's order.
//
// The reference reads the whole file into one pinned buffer, copies it, computes, copies everything back and
// writes it (C:342-346, C:365-367, C:585-588).  Here a file streams through a small RING of pinned buffers
// (up to 4 x 32 MiB, whatever the file size): the fread of chunk k + 1 runs while chunk k crosses PCIe and is
// histogrammed, and on the way out the fwrite of chunk k runs while chunk k + 1 crosses PCIe.  Pinned memory is
// bounded and allocated for the file at hand (a 160 KB text file gets one 1 MiB slot).
//
// Side index: with HF_SIDE_INDEX=1 in the environment `archive f` also writes
// f.compressed.idx (the records of hf_compress_indexed); `extract f.compressed` uses f.compressed.idx when it is
// there and describes this image, and decodes without it otherwise.  The .compressed file is the same bytes either way.
#include <dirent.h>
#include <stdlib.h>

#include <string>

#include "common.cuh"

namespace hf {

static const size_t RING_SLOT_MAX = 32u << 20;
static const int RING_SLOTS = 4;

static int ensure_ring(Ctx *c, uint64_t file_bytes)
{
    size_t slot = RING_SLOT_MAX;
    while (slot > (1u << 20) && slot / 2 >= file_bytes) slot /= 2;
    int n = (int)((file_bytes + slot - 1) / slot);
    n = n < 1 ? 1 : (n > RING_SLOTS ? RING_SLOTS : n);
    if (c->ring && c->ring_slot >= slot && c->ring_n >= n) return HF_OK;
    if (c->ring) { cudaFreeHost(c->ring); c->ring = nullptr; }
    HF_CUDA(c, cudaMallocHost((void **)&c->ring, slot * n));
    c->ring_slot = slot;
    c->ring_n = n;
    for (int i = 0; i < RING_SLOTS; i++)
        if (!c->ring_ev[i]) HF_CUDA(c, cudaEventCreateWithFlags(&c->ring_ev[i], cudaEventDisableTiming));
    return HF_OK;
}

void ring_release(Ctx *c)
{
    if (c->ring) cudaFreeHost(c->ring);
    c->ring = nullptr;
    for (int i = 0; i < RING_SLOTS; i++) if (c->ring_ev[i]) { cudaEventDestroy(c->ring_ev[i]); c->ring_ev[i] = nullptr; }
}

// file -> device through the ring; after(off, len) is called when a chunk's copy has been enqueued on the copy
// stream and ring_ev of its slot recorded (the caller makes its compute stream wait on that event)
template <typename F>
static int ring_upload(Ctx *c, FILE *f, uint64_t n, uint8_t *d_dst, uint8_t *last_byte, F after)
{
    for (uint64_t off = 0, k = 0; off < n; off += c->ring_slot, k++) {
        const int s = (int)(k % c->ring_n);
        const uint64_t len = n - off < c->ring_slot ? n - off : c->ring_slot;
        if (k >= (uint64_t)c->ring_n) HF_CUDA(c, cudaEventSynchronize(c->ring_ev[s]));    // the slot's last copy is done
        uint8_t *h = c->ring + (size_t)s * c->ring_slot;
        if (fread(h, 1, len, f) != len) return set_err(c, HF_ERR_IO, "short read");
        if (last_byte) *last_byte = h[len - 1];
        HF_CUDA(c, cudaMemcpyAsync(d_dst + off, h, len, cudaMemcpyHostToDevice, c->copy_stream));
        HF_CUDA(c, cudaEventRecord(c->ring_ev[s], c->copy_stream));
        int rc = after(off, len, c->ring_ev[s]);
        if (rc) return rc;
    }
    return HF_OK;
}

// device -> file through the ring: the fwrite of a chunk runs while the next one crosses PCIe
static int ring_download(Ctx *c, const uint8_t *d_src, uint64_t n, FILE *o)
{
    uint64_t pending_off = 0, pending_len = 0;
    int pending_slot = -1;
    for (uint64_t off = 0, k = 0; off < n || pending_slot >= 0; off += c->ring_slot, k++) {
        int s = -1;
        uint64_t len = 0;
        if (off < n) {
            s = (int)(k % c->ring_n);
            len = n - off < c->ring_slot ? n - off : c->ring_slot;
            if (c->ring_n == 1 && pending_slot >= 0) {                  // one slot: write it out before it is refilled
                HF_CUDA(c, cudaEventSynchronize(c->ring_ev[pending_slot]));
                if (fwrite(c->ring + (size_t)pending_slot * c->ring_slot, 1, pending_len, o) != pending_len) return set_err(c, HF_ERR_IO, "short write");
                pending_slot = -1;
            }
            HF_CUDA(c, cudaMemcpyAsync(c->ring + (size_t)s * c->ring_slot, d_src + off, len, cudaMemcpyDeviceToHost, c->d2h_stream));
            HF_CUDA(c, cudaEventRecord(c->ring_ev[s], c->d2h_stream));
        }
        if (pending_slot >= 0) {
            HF_CUDA(c, cudaEventSynchronize(c->ring_ev[pending_slot]));
            if (fwrite(c->ring + (size_t)pending_slot * c->ring_slot, 1, pending_len, o) != pending_len) return set_err(c, HF_ERR_IO, "short write");
        }
        pending_slot = s; pending_off = off; pending_len = len;
        (void)pending_off;
        // with more than two slots a copy may run ahead of the write by one chunk only: the slot it overwrites
        // (k + 1) % ring_n has been written unless ring_n == 2, where the write above already freed it
    }
    return HF_OK;
}

static bool file_exists(const std::string &name)
{   // D:222-240: a file or a directory of that name
    FILE *fp = fopen(name.c_str(), "rb");
    if (fp) { fclose(fp); return true; }
    DIR *d = opendir(name.c_str());
    if (d) { closedir(d); return true; }
    return false;
}

static bool env_on(const char *name)
{
    const char *e = getenv(name);
    return e && e[0] && e[0] != '0';
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

int hf_archive_file(hf_ctx *ctx, const char *path)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    FILE *f = fopen(path, "rb");
    if (!f) {                                                   // C:325-330
        printf("%s file does not exist\nProcess has been terminated\n", path);
        return HF_OK;
    }
    struct Closer { FILE *f; ~Closer() { if (f) fclose(f); } } close_in{f};
    fseeko(f, 0, SEEK_END);
    const uint64_t n = (uint64_t)ftello(f);
    fseeko(f, 0, SEEK_SET);
    printf("The size of the sum of ORIGINAL files is: %llu bytes\n", (unsigned long long)n);   // C:335
    const bool timing = env_on("HF_TIMING"), want_index = env_on("HF_SIDE_INDEX");
    const double w0 = now_ms();
    const uint64_t bound = hf_compress_bound(n);
    int rc = ensure_ring(c, n > bound ? n : bound);
    if (rc) return rc;
    rc = ensure_buf(c, &c->d_in, &c->d_in_bytes, n + 16);
    if (rc) return rc;
    rc = ensure_buf(c, &c->d_out, &c->d_out_bytes, bound + 16);
    if (rc) return rc;
    uint8_t *d_in = reinterpret_cast<uint8_t *>(c->d_in);
    const uint32_t pre = 3 + (uint32_t)(n & 1);
    uint8_t *d_file = reinterpret_cast<uint8_t *>(c->d_out) + (16 - pre);       // the bit stream starts 16-byte aligned
    unsigned long long *d_hist = reinterpret_cast<unsigned long long *>(c->d_hist);
    cudaEvent_t t[4] = {nullptr, nullptr, nullptr, nullptr};
    struct Ev { cudaEvent_t *t; ~Ev() { for (int i = 0; i < 4; i++) if (t[i]) cudaEventDestroy(t[i]); } } free_ev{t};
    for (int i = 0; i < 4; i++) HF_CUDA(c, cudaEventCreate(&t[i]));
    const double w1 = now_ms();

    // ---- "Histograming" (C:356-399): disk -> ring -> device, the histogram of a chunk under the next chunk's read ----
    HF_CUDA(c, cudaEventRecord(t[0], c->stream));
    HF_CUDA(c, cudaMemsetAsync(d_hist, 0, NSYM * 8, c->stream));
    HF_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    HF_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev[0], 0));
    uint8_t last = 0;
    rc = ring_upload(c, f, n, d_in, &last, [&](uint64_t off, uint64_t len, cudaEvent_t ev) -> int {
        HF_CUDA(c, cudaStreamWaitEvent(c->stream, ev, 0));
        return launch_histogram(c, d_in + off, len & ~1ull, d_hist);    // chunks are even-sized; the odd last byte is no symbol
    });
    if (rc) return rc;
    HF_CUDA(c, cudaEventRecord(t[1], c->stream));
    const double w2 = now_ms();

    // ---- construction (h:695-784), header (C:427-487), "Encoding" (C:492-593): sizes and start bit stay on the device ----
    uint64_t total = 0, index_bytes = 0;
    uint8_t *d_index = nullptr;
    if (want_index) {
        // the side index needs the stream's geometry on the host: this variant learns the sizes before it packs
        const uint64_t icap = hf_index_bound(n);
        HF_CUDA(c, cudaMalloc((void **)&d_index, icap));
        HF_CUDA(c, cudaEventRecord(t[2], c->stream));
        rc = hf_compress_indexed(ctx, d_in, n, d_file, bound, &total, d_index, icap, &index_bytes);
        if (rc) { cudaFree(d_index); return rc; }
    } else {
        ShardPlan *d_plan = nullptr;
        Codebook *cb = reinterpret_cast<Codebook *>(c->d_cb);
        rc = launch_codebook(c, d_hist, cb);
        if (rc) return rc;
        HF_CUDA(c, cudaEventRecord(t[2], c->stream));
        rc = launch_plan_single(c, cb, n, bound, &d_plan);
        if (rc) return rc;
        rc = launch_header_pack(c, cb, n, last, nullptr, d_file, bound, d_plan);
        if (rc) return rc;
        rc = launch_encode(c, d_in, n, cb, d_file, 0, d_plan);
        if (rc) return rc;
        HF_CUDA(c, cudaEventRecord(t[3], c->stream));
        rc = compress_result(c, d_plan, bound, &total);
        if (rc) return rc;
    }
    if (want_index) { HF_CUDA(c, cudaEventRecord(t[3], c->stream)); HF_CUDA(c, cudaStreamSynchronize(c->stream)); }
    const double w3 = now_ms();
    hf_cb_info_t info;
    rc = hf_codebook_info(ctx, c->d_cb, &info);
    if (rc) return rc;
    float ms_hist = 0, ms_cons = 0, ms_enc = 0;
    cudaEventElapsedTime(&ms_hist, t[0], t[1]);
    cudaEventElapsedTime(&ms_cons, t[1], t[2]);
    cudaEventElapsedTime(&ms_enc, t[2], t[3]);
    if (want_index) ms_cons = 0;                                // construction runs inside the indexed call: counted with the encoding

    // ---- image -> file (C:585-601), the write of a chunk under the next chunk's copy ----
    const std::string outp = std::string(path) + ".compressed";                             // C:427-429
    FILE *o = fopen(outp.c_str(), "wb");
    if (!o) { if (d_index) cudaFree(d_index); return set_err(c, HF_ERR_IO, "hf_archive_file: cannot write %s", outp.c_str()); }
    rc = ring_download(c, d_file, total, o);
    if (fclose(o) != 0 && !rc) rc = set_err(c, HF_ERR_IO, "hf_archive_file: cannot write %s", outp.c_str());
    if (!rc && d_index && index_bytes) {
        const std::string idxp = outp + ".idx";
        FILE *x = fopen(idxp.c_str(), "wb");
        if (!x) rc = set_err(c, HF_ERR_IO, "hf_archive_file: cannot write %s", idxp.c_str());
        else { rc = ring_download(c, d_index, index_bytes, x); if (fclose(x) != 0 && !rc) rc = set_err(c, HF_ERR_IO, "hf_archive_file: cannot write %s", idxp.c_str()); }
    }
    if (d_index) cudaFree(d_index);
    if (rc) return rc;
    const double w4 = now_ms();

    // the reference's lines, in its order (C:385, C:399, h:704-705, h:780-782, C:490, C:545, C:571, C:593, C:611-631)
    const unsigned long long pending = info.table_bits & 7;     // bits of the header's last byte (the reference's bitCounter)
    const unsigned long long alloc_bits = ((pending + info.payload_bits + 7) / 8) * 8;
    printf("Unique symbols count: %u\n", info.n_unique);
    printf("Histograming took %g ms\n", (w2 - w1) > ms_hist ? (w2 - w1) : ms_hist);
    printf("threadsPerBlock: %d\nnumBlocks: %d\n", 1024, 1);
    printf("construction time: %.3f ms, symbols/s: %.3f\n", ms_cons, ms_cons > 0 ? (float)info.n_unique / (ms_cons * 1e-3f) : 0.f);
    printf("Start Encoding\nstart init thrust\n");
    printf("Number of bytes allocated for h_encode_buffer: %llu\n", alloc_bits);
    printf("Encoding took %g ms\n", (double)ms_enc + (w4 - w3));
    printf("The size of the COMPRESSED file is: %llu bytes\n", (unsigned long long)total);
    printf("Compressed file's size is [%g%%] of the original files.\n", 100.0f * (float)total / (float)n);
    if (total > n) printf("\nWARNING: The compressed file's size is larger than the sum of the originals.\n\n");
    printf("\nCreated compressed file: %s\nCompression is complete\n", outp.c_str());
    if (timing)
        fprintf(stderr, "[hf timing] setup %.1f ms, read+H2D+histogram %.1f ms (device %.2f), codebook %.2f ms, header+encode %.2f ms, "
                        "wait %.1f ms, D2H+write %.1f ms\n", w1 - w0, w2 - w1, ms_hist, ms_cons, ms_enc, w3 - w2, w4 - w3);
    return HF_OK;
}

int hf_extract_file(hf_ctx *ctx, const char *path)
{
    NEED_CTX(ctx);
    Ctx *c = CTX(ctx);
    FILE *f = fopen(path, "rb");
    if (!f) { printf("%s does not exist\n", path); return HF_OK; }      // D:59-63
    struct Closer { FILE *f; ~Closer() { if (f) fclose(f); } } close_in{f};
    const bool timing = env_on("HF_TIMING");
    const double w0 = now_ms();
    fseeko(f, 0, SEEK_END);
    const uint64_t nb = (uint64_t)ftello(f);
    fseeko(f, 0, SEEK_SET);
    if (nb < 11) return set_err(c, HF_ERR_FORMAT, "hf_extract_file: %llu bytes is shorter than any image", (unsigned long long)nb);
    // a side index next to the image (written by `HF_SIDE_INDEX=1 archive`): optional, verified, never required
    const std::string idxp = std::string(path) + ".idx";
    FILE *x = getenv("HF_SIDE_INDEX") && !env_on("HF_SIDE_INDEX") ? nullptr : fopen(idxp.c_str(), "rb");
    struct Closer close_idx{x};
    uint64_t xb = 0;
    if (x) { fseeko(x, 0, SEEK_END); xb = (uint64_t)ftello(x); fseeko(x, 0, SEEK_SET); }
    int rc = ensure_ring(c, nb);
    if (rc) return rc;
    rc = ensure_buf(c, &c->d_out, &c->d_out_bytes, nb + 256);
    if (rc) return rc;
    // the index is made for one alignment of the image (its address modulo 16): archive packs at 16 - preamble
    uint8_t first4[4] = {0, 0, 0, 0};
    if (fread(first4, 1, 3, f) != 3) return set_err(c, HF_ERR_IO, "hf_extract_file: short read");
    fseeko(f, 0, SEEK_SET);
    const uint32_t pre = 3 + (first4[2] != 0);
    uint8_t *d_file = reinterpret_cast<uint8_t *>(c->d_out) + (16 - pre);
    HF_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    HF_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev[0], 0));
    rc = ring_upload(c, f, nb, d_file, nullptr, [&](uint64_t, uint64_t, cudaEvent_t ev) -> int {
        HF_CUDA(c, cudaStreamWaitEvent(c->stream, ev, 0));
        return HF_OK;
    });
    if (rc) return rc;
    uint8_t *d_index = nullptr;
    if (x && xb >= 64) {
        HF_CUDA(c, cudaMalloc((void **)&d_index, xb + 16));
        rc = ring_upload(c, x, xb, d_index, nullptr, [&](uint64_t, uint64_t, cudaEvent_t ev) -> int {
            HF_CUDA(c, cudaStreamWaitEvent(c->stream, ev, 0));
            return HF_OK;
        });
        if (rc) { cudaFree(d_index); return rc; }
    }
    const double w1 = now_ms();
    hf_header_info_t info;
    rc = hf_parse_header(ctx, d_file, nb, c->d_tab, &info);             // synchronises: the output size
    if (rc) { if (d_index) cudaFree(d_index); return rc; }
    uint64_t n = info.original_bytes;
    rc = ensure_buf(c, &c->d_in, &c->d_in_bytes, n + 16);
    if (rc) { if (d_index) cudaFree(d_index); return rc; }
    uint8_t *d_outbuf = reinterpret_cast<uint8_t *>(c->d_in);
    if (d_index) rc = hf_decompress_indexed(ctx, d_file, nb, d_index, xb, d_outbuf, n + 16, &n);
    else rc = hf_decompress(ctx, d_file, nb, d_outbuf, n + 16, &n);
    if (d_index) cudaFree(d_index);
    if (rc) return rc;
    const double w2 = now_ms();
    std::string name = "DECOMPRESSED_FILE";                                             // D:104
    if (file_exists(name)) {                                                            // D:185-219
        for (int k = 1; k < 10; k++) {
            name = "DECOMPRESSED_FILE(" + std::to_string(k) + ")";
            if (!file_exists(name)) break;
        }
    }
    FILE *o = fopen(name.c_str(), "wb");
    if (!o) return set_err(c, HF_ERR_IO, "hf_extract_file: cannot write %s", name.c_str());
    rc = ring_download(c, d_outbuf, n, o);
    if (fclose(o) != 0 && !rc) rc = set_err(c, HF_ERR_IO, "hf_extract_file: cannot write %s", name.c_str());
    if (rc) return rc;
    printf("Decompression is complete\n");                                              // D:113
    if (timing)
        fprintf(stderr, "[hf timing] read+H2D %.1f ms, parse+decode %.1f ms, D2H+write %.1f ms%s\n", w1 - w0, w2 - w1, now_ms() - w2,
                d_index ? " (side index)" : "");
    return HF_OK;
}

}  // extern "C"
