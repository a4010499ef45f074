// decode.cu — GPU decompressor: header parse, decode-table build, self-synchronising
// subsequence decode.  The reference has NO GPU decoder: its Decompressor.cu is a host
// program that fread()s one byte at a time and walks a pointer tree one bit at a time
// (/root/reference/Decompressor.cu:68-103, :129-182, :243-291).  Its format carries no
// gap / offset array, so parallel decode must find code word boundaries
// by itself.
//
// Header (dec_parse_kernel): entry k+1 starts 24+len_k bits after entry k — a serial chain.
//   One CTA chases it SPECULATIVELY from shared-memory staging: thread i assumes the next
//   entries keep the current stride and reads the length of entry k+i; the first mismatch
//   confirms up to 1024 entries per step (the table is sorted by count, so lengths come in
//   long runs).  Any header parses correctly, sorted ones ~1000x faster than entry by entry.
//
// Tables (dt_*): a general two-level table for any code length up to 64 bits (12-bit primary, per-prefix
//   secondary tables of up to +12 bits, a linear list for still longer codes); decode2.cu derives the
//   shared-memory planes of the hot kernels from it.
//
// Payload: chunks of 512 subsequences x 256 bits; the stages below run on a range of chunks (the whole stream,
// or one slice of a pipelined host-buffer decode):
//   dec_sync3_kernel (decode2.cu)   code word boundaries and counts per subsequence, exact
//   dec_fix2_kernel  (decode2.cu)   repair of the first spans of a group from the previous group's overflow
//   dec_scan1/2/3_kernel            exclusive scan of the chunk symbol counts
//   dec_write3_kernel (decode2.cu)  symbols out
//
// Algorithmic bytes: C read + N written (the payload is read twice: traffic ~ 2C + N + C/16).
#include "common.cuh"
#include "decode_common.cuh"

namespace hf {

// -----------------------------------------------------------------------------------
// decode-table build from (sym, len, code)[U]
struct TabSrc {                         // workspace arrays filled by the header parser or from a Codebook
    uint32_t sym[NSYM];
    uint32_t len[NSYM];
    unsigned long long code[NSYM];
    unsigned long long entry_pos[NSYM];
    unsigned long long len_mask[2];     // bit (len-1) set for every length in use
    uint32_t U;
    uint32_t pad;
};
static_assert(sizeof(TabSrc) <= WS_SCRATCH_BYTES, "TabSrc lives in the workspace's scratch region");

__global__ void dt_from_codebook_kernel(const Codebook *__restrict__ cb, TabSrc *__restrict__ src)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) src->U = cb->U;
    if (k < cb->U) {
        uint32_t s = cb->order[k];
        src->sym[k] = s; src->len[k] = cb->len[s]; src->code[k] = cb->code[s];
    }
}

__global__ void dt_depth_kernel(const TabSrc *__restrict__ src, DecodeTable *__restrict__ tab,
                                unsigned long long *len_mask)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t U = src->U;
    if (k >= U) return;
    uint32_t len = src->len[k];
    if (len == 0 || len > 64) { if (!(U == 1 && len == 0)) atomicExch(&tab->status, (uint32_t)HF_ERR_FORMAT); return; }
    atomicOr(len_mask, 1ull << (len - 1));
    if (len > K1) atomicMax(&tab->sub_depth[(uint32_t)(src->code[k] >> (len - K1))], len - K1);
}

__global__ void __launch_bounds__(1024, 1)
dt_offsets_kernel(const TabSrc *__restrict__ src, DecodeTable *__restrict__ tab, const unsigned long long *len_mask)
{
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_cap;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t per = (1u << K1) / 1024;     // 4 prefixes per thread
    uint32_t cap = K2MAX;
    for (int attempt = 0; attempt < 2; attempt++) {
        uint32_t sz[per], sum = 0;
        for (uint32_t j = 0; j < per; j++) {
            uint32_t d = tab->sub_depth[tid * per + j];
            sz[j] = d ? (1u << min(d, cap)) : 0u;
            sum += sz[j];
        }
        uint32_t x = sum;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = s_w[lane], t = s;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
            s_w[lane] = t - s;
            if (lane == 31) s_w[32] = t;
        }
        __syncthreads();
        uint32_t total = s_w[32];
        if (total > T2_CAP && attempt == 0) { cap = 8; __syncthreads(); continue; }   // 2^12 * 2^8 always fits
        uint32_t run = x - sum + s_w[wid];
        for (uint32_t j = 0; j < per; j++) {
            uint32_t p = tid * per + j;
            uint32_t d = tab->sub_depth[p];
            if (d) { tab->t1[p] = (run << 8) | E_SUB | min(d, cap); run += sz[j]; }
        }
        if (tid == 0) { tab->t2_used = total; s_cap = cap; }
        break;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = len_mask[0];
        uint32_t U = src->U;
        tab->U = U;
        uint32_t g = 0, mn = 0, mx = 0;
        for (uint32_t l = 1; l <= 64; l++)
            if ((m >> (l - 1)) & 1) {
                if (!mn) mn = l;
                mx = l;
                uint32_t a = g, b = l;                  // gcd(g, l); gcd(0, l) = l
                while (b) { uint32_t t = a % b; a = b; b = t; }
                g = a;
            }
        tab->maxlen = mx; tab->minlen = mn; tab->len_gcd = g ? g : 1;
        tab->n_long = 0;
        tab->single_sym = (U == 1 && src->len[0] == 0) ? (0x10000u | src->sym[0]) : 0u;
        tab->k2cap = s_cap;
    }
}

__global__ void dt_fill_kernel(const TabSrc *__restrict__ src, DecodeTable *__restrict__ tab)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= src->U) return;
    uint32_t len = src->len[k], sym = src->sym[k];
    unsigned long long code = src->code[k];
    if (len == 0 || len > 64) return;
    const uint32_t leaf = (sym << 8) | len;
    if (len <= K1) {
        uint32_t base = (uint32_t)code << (K1 - len), n = 1u << (K1 - len);
        for (uint32_t i = 0; i < n; i++) tab->t1[base + i] = leaf;
        return;
    }
    uint32_t p = (uint32_t)(code >> (len - K1));
    uint32_t e1 = tab->t1[p];
    uint32_t sb = e1 & 31u, off = e1 >> 8;
    uint32_t rem = len - K1;
    if (rem <= sb) {
        uint32_t low = (uint32_t)(code & ((1ull << rem) - 1));
        uint32_t base = low << (sb - rem), n = 1u << (sb - rem);
        for (uint32_t i = 0; i < n; i++) tab->t2[off + base + i] = leaf;
    } else {
        uint32_t idx = (uint32_t)((code >> (rem - sb)) & ((1u << sb) - 1));
        const uint32_t slot = atomicAdd(&tab->n_long, 1u);
        LongCode lc;
        lc.code_left = code << (64 - len);
        lc.leaf = leaf;
        lc.next = atomicExch(&tab->t2[off + idx], E_LIST | E_SUB | (slot << 8));   // push on the slot's list
        tab->longs[slot] = lc;
    }
}

// -----------------------------------------------------------------------------------
// header parse
__device__ __forceinline__ uint32_t hdr_byte_at_bit(const uint8_t *sm, uint32_t bit)
{   // 8 bits starting at bit offset `bit` of the staged bytes
    uint32_t i = bit >> 3, sh = bit & 7;
    return ((((uint32_t)sm[i] << 8) | sm[i + 1]) >> (8 - sh)) & 0xFFu;
}

constexpr uint32_t HDR_STAGE = 64 * 1024;               // staged header bytes

__global__ void __launch_bounds__(1024, 1)
dec_parse_kernel(const uint8_t *__restrict__ file, unsigned long long file_bytes, TabSrc *__restrict__ src,
                 DecodeTable *__restrict__ tab, hf_header_info_t *__restrict__ info)
{
    extern __shared__ uint8_t sm[];
    __shared__ unsigned long long s_pos;                // bit position (from the stream start) of the next entry
    __shared__ uint32_t s_k, s_done, s_err;
    const uint32_t tid = threadIdx.x;

    uint32_t U = 0, is_odd = 0, last = 0, pre = 3;
    bool ok = file_bytes >= 3;
    if (ok) {
        U = (uint32_t)file[0] | ((uint32_t)file[1] << 8);           // D:69
        is_odd = file[2] != 0;                                      // D:76
        pre = 3 + is_odd;
        ok = file_bytes >= pre;
        if (ok && is_odd) last = file[3];                           // D:77-80
        if (ok && U == 0) U = (file_bytes - pre == 8) ? 0u : 65536u;   // D:70-71; U = 0 is our N < 2 case
    }
    const unsigned long long stream_bytes = ok ? file_bytes - pre : 0;
    const uint8_t *stream = file + pre;
    if (tid == 0) { s_pos = 0; s_k = 0; s_done = (U == 0 || !ok); s_err = !ok; }
    __syncthreads();

    __shared__ uint32_t s_first;                        // first entry of a step that is not where the stride predicts
    while (!s_done) {
        // stage HDR_STAGE bytes starting at the byte holding s_pos
        const unsigned long long b0 = s_pos >> 3;
        for (uint32_t i = tid; i < HDR_STAGE + 16; i += 1024)
            sm[i] = (b0 + i < stream_bytes) ? stream[b0 + i] : 0;
        __syncthreads();
        // Steps of up to 1024 entries: thread i assumes the entries keep the stride of the first one and reads the
        // length of entry k + i at its predicted place; everything before the first mismatch is confirmed at once.
        for (;;) {
            const unsigned long long pos = s_pos;
            const uint32_t k = s_k;
            if (k >= U) { if (tid == 0) s_done = 1; break; }
            const uint32_t rel = (uint32_t)(pos - b0 * 8);
            if ((rel >> 3) + 64 > HDR_STAGE) break;                     // restage
            if ((pos >> 3) + 3 > stream_bytes) { if (tid == 0) { s_err = 1; s_done = 1; } break; }
            const uint32_t stride = 24 + hdr_byte_at_bit(sm, rel + 16);
            uint32_t n_can = (HDR_STAGE * 8 - 64 - rel) / stride;      // entries whose length byte is staged
            if (n_can < 32 && (rel >> 3) > 1024) break;                // a short step near the end of the stage: restage
            n_can = min(min(n_can, 1024u), U - k);
            if (tid == 0) s_first = n_can;
            __syncthreads();
            uint32_t L = 0;
            if (tid < n_can) {
                const uint32_t myrel = rel + tid * stride;
                const bool inb = ((pos + (unsigned long long)tid * stride) >> 3) + 3 <= stream_bytes;
                L = inb ? hdr_byte_at_bit(sm, myrel + 16) : 0xFFFFFFFFu;    // out of the stream: never matches
                if (24 + L != stride) atomicMin(&s_first, tid);
            }
            __syncthreads();
            const uint32_t f = s_first;                                 // entries 0 .. f sit at their predicted places
            const uint32_t n_ok = min(f + 1, n_can);
            if (tid < n_ok && L != 0xFFFFFFFFu) {
                src->entry_pos[k + tid] = pos + (unsigned long long)tid * stride;
                src->len[k + tid] = L;
            }
            if (tid == n_ok - 1) {
                if (L == 0xFFFFFFFFu) { s_err = 1; s_done = 1; }        // an entry that is needed lies past the stream
                s_pos = pos + (unsigned long long)tid * stride + 24 + L;
                s_k = k + n_ok;
            }
            __syncthreads();
            if (s_done) break;
        }
        __syncthreads();
    }
    if (tid == 0) {
        unsigned long long pos = s_pos;
        unsigned long long n = 0;
        bool err = s_err;
        if (!err && (pos + 64 + 7) / 8 > stream_bytes) err = true;
        if (!err) {
            for (int i = 0; i < 8; i++) {                              // D:243-255
                unsigned long long bit = pos + 8 * i;
                uint32_t bi = (uint32_t)(bit & 7);
                uint32_t v = ((((uint32_t)stream[bit >> 3] << 8) |
                               ((bit >> 3) + 1 < stream_bytes ? stream[(bit >> 3) + 1] : 0u)) >> (8 - bi)) & 0xFFu;
                n |= (unsigned long long)v << (8 * i);
            }
            if ((n & 1) != is_odd) err = true;
        }
        src->U = err ? 0 : U;
        info->n_unique = U; info->is_odd = is_odd; info->last_byte = last;
        info->original_bytes = n;
        info->payload_start_bit = pre * 8ull + pos + 64;
        info->status = err ? HF_ERR_FORMAT : HF_OK;
        info->max_code_bits = 0;
        if (err) tab->status = HF_ERR_FORMAT;
    }
}

// up to 64 bits at an arbitrary bit position of a byte buffer, zero past the end
__device__ __forceinline__ unsigned long long bits_at(const uint8_t *p, unsigned long long nbytes,
                                                      unsigned long long bitpos, uint32_t nbits)
{
    unsigned long long v = 0;
    unsigned long long b = bitpos >> 3;
    uint32_t sh = (uint32_t)(bitpos & 7), got = 0;
    // first (partial) byte, then whole bytes
    while (got < nbits) {
        uint32_t byte = b < nbytes ? p[b] : 0u;
        uint32_t avail = 8 - sh;
        uint32_t take = min(avail, nbits - got);
        uint32_t chunk = (byte >> (avail - take)) & ((1u << take) - 1);
        v = (v << take) | chunk;
        got += take; sh = 0; b++;
    }
    return v;
}

// pulls (sym, code) of every entry out of the file image once the entry positions are known
__global__ void dec_entries_kernel(const uint8_t *__restrict__ file, unsigned long long file_bytes,
                                   TabSrc *__restrict__ src)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= src->U) return;
    const uint32_t pre = 3 + (file[2] != 0);
    const uint8_t *stream = file + pre;
    const unsigned long long nb = file_bytes - pre;
    const unsigned long long pos = src->entry_pos[k];
    uint32_t sym = (uint32_t)bits_at(stream, nb, pos, 16);           // D:178-182
    uint32_t len = (uint32_t)bits_at(stream, nb, pos + 16, 8);       // D:93
    unsigned long long code = 0;
    if (len && len <= 64) code = bits_at(stream, nb, pos + 24, len); // D:129-163
    src->sym[k] = sym;
    src->len[k] = len;
    src->code[k] = code;
}

// Exclusive scan of the chunk symbol counts -> chunkBase, in three small steps: every block of 4096 chunks scans
// itself (dec_scan1_kernel), one CTA scans the block totals (dec_scan2_kernel), dec_scan3_kernel adds them back.
constexpr uint32_t DSCAN_PER_BLOCK = 4096;

__global__ void __launch_bounds__(1024)
dec_scan1_kernel(DecWork *work, unsigned long long nch, unsigned long long c0, unsigned long long c1,
                 unsigned long long *block_tot)
{
    __shared__ unsigned long long s_w[33];
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr uint32_t PER = DSCAN_PER_BLOCK / 1024;
    const unsigned long long i0 = c0 + (unsigned long long)blockIdx.x * DSCAN_PER_BLOCK + tid * PER;
    uint32_t v[PER];
    unsigned long long sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < PER; j++) { v[j] = i0 + j < c1 ? L.chunkCnt[i0 + j] : 0u; sum += v[j]; }
    unsigned long long x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned long long s = s_w[lane], t = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
        s_w[lane] = t - s;
        if (lane == 31) s_w[32] = t;
    }
    __syncthreads();
    unsigned long long run = x - sum + s_w[wid];
#pragma unroll
    for (uint32_t j = 0; j < PER; j++) { if (i0 + j < c1) L.chunkBase[i0 + j] = run; run += v[j]; }
    if (tid == 0) block_tot[blockIdx.x] = s_w[32];
}

__global__ void __launch_bounds__(1024)
dec_scan2_kernel(DecWork *work, unsigned long long *block_tot, uint32_t nblocks)
{
    __shared__ unsigned long long s_w[33];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t per = (nblocks + 1023) / 1024;
    const uint32_t lo = min(nblocks, tid * per), hi = min(nblocks, (tid + 1) * per);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += block_tot[i];
    unsigned long long x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned long long s = s_w[lane], t = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
        s_w[lane] = t - s;
        if (lane == 31) s_w[32] = t;
    }
    __syncthreads();
    // work->result[2]: symbols of the chunks before this slice (0 for the first or only slice), then of all so far
    const unsigned long long carry = work->result[2];
    unsigned long long run = carry + x - sum + s_w[wid];
    for (uint32_t i = lo; i < hi; i++) { const unsigned long long v = block_tot[i]; block_tot[i] = run; run += v; }
    __syncthreads();
    if (tid == 0) work->result[2] = carry + s_w[32];
}

__global__ void __launch_bounds__(1024)
dec_scan3_kernel(DecWork *work, unsigned long long nch, unsigned long long c0, unsigned long long c1,
                 const unsigned long long *block_tot)
{
    DecLayout L(work, nch);
    const unsigned long long add = block_tot[blockIdx.x];
    if (add == 0) return;
    for (uint32_t j = threadIdx.x; j < DSCAN_PER_BLOCK; j += 1024) {
        const unsigned long long i = c0 + (unsigned long long)blockIdx.x * DSCAN_PER_BLOCK + j;
        if (i < c1) L.chunkBase[i] += add;
    }
}

// U == 1 with a zero-length code: the payload is empty, every symbol is the same
__global__ void dec_fill_kernel(const DecodeTable *__restrict__ tab, const DecWork *__restrict__ work,
                                uint16_t *__restrict__ out)
{
    if (!(tab->single_sym & 0x10000u)) return;
    const unsigned long long n_symbols = work->start[1];
    const uint16_t s = (uint16_t)tab->single_sym;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n_symbols; i += stride) out[i] = s;
}

// -----------------------------------------------------------------------------------
int launch_table_planes(Ctx *c, DecodeTable *d_tab);             // decode2.cu
int launch_sync2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                 unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                 unsigned long long c0, unsigned long long c1, bool tail_only, bool speculative);
int launch_write2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                  const DecodeTable *d_tab, DecWork *work, unsigned long long nch, unsigned long long c0,
                  unsigned long long c1, uint16_t *out, bool check);
int launch_idx_chunks(Ctx *c, DecWork *work, unsigned long long nch);
int launch_idx_total(Ctx *c, DecWork *work);
int launch_fix2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                unsigned long long c0, unsigned long long c1, bool speculative);
int launch_fix_head(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long range_end_bit,
                    const DecodeTable *d_tab, DecWork *work, unsigned long long nch);

// where a decode starts and how much it may write (DecWork::start), from host values ...
__global__ void dec_start_kernel(DecWork *work, unsigned long long F0, unsigned long long n_symbols)
{
    work->start[0] = F0;
    work->start[1] = n_symbols;
}

// ... from a header parsed on the device (hf_decompress: nothing visits the host before the decode is enqueued).  A
// malformed header or an output that does not fit decodes nothing: the first code word is put behind everything.
__global__ void dec_start_info_kernel(DecWork *work, const hf_header_info_t *__restrict__ info,
                                      const DecodeTable *__restrict__ tab, unsigned long long lead_bits,
                                      unsigned long long capacity, uint8_t *d_out)
{
    const unsigned long long n = info->original_bytes;
    bool ok = info->status == 0 && tab->status == 0;
    if (ok && n > capacity) { work->flags[2] = 1; ok = false; }
    work->start[0] = ok ? lead_bits + info->payload_start_bit : NO_START - 1;
    work->start[1] = ok ? n / 2 : 0;
    if (ok && info->is_odd) d_out[n - 1] = (uint8_t)info->last_byte;      // D:286-289
}

// ... from the hand-over bit of a sharded stream: *first_bit = bits from the range's first byte to its first code word
// (the predecessor's overflow); a range no code word starts in keeps F0 at or behind its end
__global__ void dec_start_range_kernel(DecWork *work, const unsigned long long *__restrict__ first_bit,
                                       unsigned long long lead_bits, unsigned long long out_symbols)
{
    work->start[0] = lead_bits + *first_bit;
    work->start[1] = out_symbols;
}

static int build_tables(Ctx *c, TabSrc *src, DecodeTable *d_tab)
{
    // zero everything but the (large) long list; t2 zero = invalid
    HF_CUDA(c, cudaMemsetAsync(d_tab, 0, offsetof(DecodeTable, longs), c->stream));
    HF_CUDA(c, cudaMemsetAsync(src->len_mask, 0, sizeof(src->len_mask), c->stream));
    HF_PROF(c, "dt_depth_kernel"); dt_depth_kernel<<<NSYM / 256, 256, 0, c->stream>>>(src, d_tab, src->len_mask);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dt_offsets_kernel"); dt_offsets_kernel<<<1, 1024, 0, c->stream>>>(src, d_tab, src->len_mask);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dt_fill_kernel"); dt_fill_kernel<<<NSYM / 256, 256, 0, c->stream>>>(src, d_tab);
    HF_LAUNCH_CHECK(c);
    return launch_table_planes(c, d_tab);
}

int launch_table_from_codebook(Ctx *c, const Codebook *d_cb, DecodeTable *d_tab)
{
    int rc = ensure_ws(c, sizeof(TabSrc));
    if (rc) return rc;
    TabSrc *src = reinterpret_cast<TabSrc *>(c->ws);
    HF_PROF(c, "dt_from_codebook_kernel"); dt_from_codebook_kernel<<<NSYM / 256, 256, 0, c->stream>>>(d_cb, src);
    HF_LAUNCH_CHECK(c);
    return build_tables(c, src, d_tab);
}

int launch_parse_header(Ctx *c, const uint8_t *d_file, uint64_t file_bytes, DecodeTable *d_tab,
                        hf_header_info_t *d_info)
{
    int rc = ensure_ws(c, sizeof(TabSrc));
    if (rc) return rc;
    TabSrc *src = reinterpret_cast<TabSrc *>(c->ws);
    if (!c->smem_attr[ATTR_PARSE]) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HDR_STAGE + 32));
        c->smem_attr[ATTR_PARSE] = true;
    }
    // status of an earlier use must not leak into this parse
    HF_CUDA(c, cudaMemsetAsync(&d_tab->status, 0, sizeof(uint32_t), c->stream));
    HF_PROF(c, "dec_parse_kernel"); dec_parse_kernel<<<1, 1024, HDR_STAGE + 32, c->stream>>>(d_file, file_bytes, src, d_tab, d_info);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_entries_kernel"); dec_entries_kernel<<<NSYM / 256, 256, 0, c->stream>>>(d_file, file_bytes, src);
    HF_LAUNCH_CHECK(c);
    return build_tables(c, src, d_tab);
}

// the scan of the chunk counts of [c0, c1); work->result[2] carries the symbols before c0 in and those before c1 out
static int launch_scan(Ctx *c, DecWork *work, unsigned long long nch, unsigned long long c0, unsigned long long c1)
{
    if (c1 <= c0) return HF_OK;
    const uint32_t nblocks = (uint32_t)((c1 - c0 + DSCAN_PER_BLOCK - 1) / DSCAN_PER_BLOCK);
    unsigned long long *block_tot = reinterpret_cast<unsigned long long *>(c->d_scan);
    if (nblocks > SCAN_BLOCKS_MAX) return set_err(c, HF_ERR_ARG, "hf_decode: stream too large");
    HF_PROF(c, "dec_scan1_kernel"); dec_scan1_kernel<<<nblocks, 1024, 0, c->stream>>>(work, nch, c0, c1, block_tot);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_scan2_kernel"); dec_scan2_kernel<<<1, 1024, 0, c->stream>>>(work, block_tot, nblocks);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_scan3_kernel"); dec_scan3_kernel<<<nblocks, 1024, 0, c->stream>>>(work, nch, c0, c1, block_tot);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// The decode kernels over the chunks [c0, c1) of the frame: the whole stream in one go, or one slice of a pipelined
// host-buffer decode (slices in ascending order; c0 a multiple of GROUP_CHUNKS).
static int launch_decode_exact(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                               unsigned long long range_end_bit, const DecodeTable *d_tab,
                               uint16_t *out16, DecWork *work, unsigned long long nch, unsigned long long c0,
                               unsigned long long c1)
{
    int rc = launch_sync2(c, frame, frame_bytes, range_end_bit, d_tab, work, nch, c0, c1, false, false);
    if (rc) return rc;
    rc = launch_fix2(c, frame, frame_bytes, range_end_bit, d_tab, work, nch, c0, c1, false);
    if (rc) return rc;
    rc = launch_scan(c, work, nch, c0, c1);
    if (rc) return rc;
    return launch_write2(c, frame, frame_bytes, d_tab, work, nch, c0, c1, out16, false);
}

// A decode of one stream, all at once (launch_decode) or in slices of chunks as its bytes arrive from the host
// (hf_decompress_host): decode_begin lays out the frame and the work area, decode_slice runs the exact kernels on
// the chunks [c0, c1) — slices in ascending order, c0 a multiple of GROUP_CHUNKS.  job.total is the device address of the
// running symbol count (symbols decoded by the slices so far).
int decode_begin(Ctx *c, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit, uint64_t n_symbols,
                 const DecodeTable *d_tab, uint8_t *d_out, DecodeJob *job)
{
    if ((uintptr_t)d_out & 1) return set_err(c, HF_ERR_ARG, "hf_decode: output must be 2-byte aligned");
    if ((start_bit >> 3) > stream_bytes) return set_err(c, HF_ERR_ARG, "hf_decode: start bit past the stream");
    d_stream += start_bit >> 3;
    stream_bytes -= start_bit >> 3;
    start_bit &= 7;
    job->frame = reinterpret_cast<const uint8_t *>((uintptr_t)d_stream & ~(uintptr_t)15);
    const unsigned long long lead = (uintptr_t)d_stream & 15;
    job->frame_bytes = lead + stream_bytes;
    job->F0 = lead * 8 + start_bit;
    job->nch = (job->frame_bytes * 8 + CHUNK_BITS - 1) / CHUNK_BITS;
    if (job->nch == 0) job->nch = 1;                    // zero-length codes: nothing to read
    if (job->nch > 0x7FFFFFFFull) return set_err(c, HF_ERR_ARG, "hf_decode: stream too large");
    job->n_symbols = n_symbols;
    job->tab = d_tab;
    job->out16 = reinterpret_cast<uint16_t *>(d_out);
    const size_t off = WS_STAGE_OFFSET;
    int rc = ensure_ws(c, off + DecLayout::bytes(job->nch));
    if (rc) return rc;
    job->work = (uint8_t *)c->ws + off;
    job->total = reinterpret_cast<unsigned long long *>(job->work) + 2;     // DecWork::result[2]
    HF_CUDA(c, cudaMemsetAsync(job->work, 0, sizeof(DecWork), c->stream));
    dec_start_kernel<<<1, 1, 0, c->stream>>>(reinterpret_cast<DecWork *>(job->work), job->F0, n_symbols);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_fill_kernel"); dec_fill_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d_tab, reinterpret_cast<DecWork *>(job->work), job->out16);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int decode_slice(Ctx *c, const DecodeJob &job, unsigned long long c0, unsigned long long c1)
{
    return launch_decode_exact(c, job.frame, job.frame_bytes, job.frame_bytes * 8, job.tab, job.out16,
                               reinterpret_cast<DecWork *>(job.work), job.nch, c0, c1);
}

// hf_decompress: the whole image, its header parsed on the device (d_info): the frame is the image itself, the first
// code word sits wherever the header ends, and the output capacity is checked on the device.  Nothing here waits for
// the host; the caller reads d_info and the flags after its one synchronisation.
int launch_decompress_image(Ctx *c, const uint8_t *d_file, uint64_t file_bytes, const hf_header_info_t *d_info,
                            const DecodeTable *d_tab, uint8_t *d_out, uint64_t capacity)
{
    if ((uintptr_t)d_out & 1) return set_err(c, HF_ERR_ARG, "hf_decompress: output must be 2-byte aligned");
    const uint8_t *frame = reinterpret_cast<const uint8_t *>((uintptr_t)d_file & ~(uintptr_t)15);
    const unsigned long long lead = (uintptr_t)d_file & 15;
    const unsigned long long frame_bytes = lead + file_bytes;
    unsigned long long nch = (frame_bytes * 8 + CHUNK_BITS - 1) / CHUNK_BITS;
    if (nch == 0) nch = 1;
    if (nch > 0x7FFFFFFFull) return set_err(c, HF_ERR_ARG, "hf_decompress: image too large");
    const size_t off = WS_STAGE_OFFSET;
    int rc = ensure_ws(c, off + DecLayout::bytes(nch));
    if (rc) return rc;
    DecWork *work = reinterpret_cast<DecWork *>((uint8_t *)c->ws + off);
    HF_CUDA(c, cudaMemsetAsync(work, 0, sizeof(DecWork), c->stream));
    dec_start_info_kernel<<<1, 1, 0, c->stream>>>(work, d_info, d_tab, lead * 8, capacity, d_out);
    HF_LAUNCH_CHECK(c);
    uint16_t *out16 = reinterpret_cast<uint16_t *>(d_out);
    HF_PROF(c, "dec_fill_kernel"); dec_fill_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d_tab, work, out16);
    HF_LAUNCH_CHECK(c);
    return launch_decode_exact(c, frame, frame_bytes, frame_bytes * 8, d_tab, out16, work, nch, 0, nch);
}

// Decode with the records of a side index (n_subs u16, as the compressor's enc_index_kernel wrote them for this
// stream at this alignment) in place of the synchronisation kernels.  The write kernel checks every record against
// the walk it drives (flags[1] on a mismatch: the caller decodes again without the index).
int launch_decode_indexed(Ctx *c, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit, uint64_t n_symbols,
                          const DecodeTable *d_tab, uint8_t *d_out, const uint16_t *d_rec, uint64_t n_subs)
{
    if (n_symbols == 0) return HF_OK;
    DecodeJob job{};
    int rc = decode_begin(c, d_stream, stream_bytes, start_bit, n_symbols, d_tab, d_out, &job);
    if (rc) return rc;
    if (n_subs != job.nch * DEC_THREADS) return set_err(c, HF_ERR_FORMAT, "index does not fit the stream");
    DecWork *work = reinterpret_cast<DecWork *>(job.work);
    DecLayout L(work, job.nch);
    HF_CUDA(c, cudaMemcpyAsync(L.info, d_rec, n_subs * 2, cudaMemcpyDeviceToDevice, c->stream));
    rc = launch_idx_chunks(c, work, job.nch);
    if (rc) return rc;
    rc = launch_scan(c, work, job.nch, 0, job.nch);
    if (rc) return rc;
    rc = launch_idx_total(c, work);
    if (rc) return rc;
    return launch_write2(c, job.frame, job.frame_bytes, d_tab, work, job.nch, 0, job.nch, job.out16, true);
}

// number of index records of a stream of stream_bytes bytes whose first code word sits start_bit bits after d_stream
uint64_t index_subs(const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit)
{
    d_stream += start_bit >> 3;
    stream_bytes -= start_bit >> 3;
    const unsigned long long frame_bytes = ((uintptr_t)d_stream & 15) + stream_bytes;
    unsigned long long nch = (frame_bytes * 8 + CHUNK_BITS - 1) / CHUNK_BITS;
    if (nch == 0) nch = 1;
    return nch * DEC_THREADS;
}

int launch_decode(Ctx *c, const uint8_t *d_stream, uint64_t stream_bytes, uint64_t start_bit,
                  uint64_t n_symbols, const DecodeTable *d_tab, uint8_t *d_out)
{
    if (n_symbols == 0) return HF_OK;
    DecodeJob job{};
    int rc = decode_begin(c, d_stream, stream_bytes, start_bit, n_symbols, d_tab, d_out, &job);
    if (rc) return rc;
    return decode_slice(c, job, 0, job.nch);
}

// result[4] of a range call from the exact kernels' work area: -, overflow, symbols, flags
__global__ void dec_result_kernel(const DecWork *__restrict__ work, unsigned long long *__restrict__ result)
{
    result[0] = 0;
    result[1] = work->result[1];
    result[2] = work->result[2];
    result[3] = (work->flags[1] ? 4ull : 0ull) | (work->result[2] > work->start[1] ? 8ull : 0ull);
}

// one rank's byte range of a sharded stream.  tail_only: just the overflow of the
// range's last code word past its end, found speculatively from the last chunk (result[1]).
int launch_decode_range(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes, uint64_t first_bit,
                        bool tail_only, const DecodeTable *d_tab, uint8_t *d_out, uint64_t out_symbols,
                        unsigned long long *d_result)
{
    if (!tail_only && ((uintptr_t)d_out & 1)) return set_err(c, HF_ERR_ARG, "hf_decode_range: output must be 2-byte aligned");
    if ((first_bit >> 3) > range_bytes) return set_err(c, HF_ERR_ARG, "hf_decode_range: first bit past the range");
    const uint8_t *p = d_range + (first_bit >> 3);
    const uint8_t *frame = reinterpret_cast<const uint8_t *>((uintptr_t)p & ~(uintptr_t)15);
    const unsigned long long lead = (unsigned long long)((uintptr_t)p & 15);
    const unsigned long long F0 = lead * 8 + (first_bit & 7);
    const unsigned long long own = range_bytes - (first_bit >> 3);
    const unsigned long long frame_bytes = lead + own + halo_bytes;        // readable
    const unsigned long long end_bit = (lead + own) * 8;                   // code words starting at or after it are not ours
    unsigned long long nch = (end_bit + CHUNK_BITS - 1) / CHUNK_BITS;
    if (nch == 0) nch = 1;
    if (nch > 0x7FFFFFFFull) return set_err(c, HF_ERR_ARG, "hf_decode_range: range too large");
    const size_t off = WS_STAGE_OFFSET;
    int rc = ensure_ws(c, off + DecLayout::bytes(nch));
    if (rc) return rc;
    DecWork *work = reinterpret_cast<DecWork *>((uint8_t *)c->ws + off);
    HF_CUDA(c, cudaMemsetAsync(work, 0, sizeof(DecWork), c->stream));
    dec_start_kernel<<<1, 1, 0, c->stream>>>(work, F0, tail_only ? ~0ull >> 8 : out_symbols);
    HF_LAUNCH_CHECK(c);
    if (tail_only) {
        rc = launch_sync2(c, frame, frame_bytes, end_bit, d_tab, work, nch, 0, nch, true, true);
        if (rc) return rc;
        // the groups of the tail converge one by one on guessed starts; the repair carries the chain from the
        // first of them (the lead-in, 240 KiB or more when the range is that long) to the range end
        rc = launch_fix2(c, frame, frame_bytes, end_bit, d_tab, work, nch, tail_first_chunk(nch) + GROUP_CHUNKS, nch, true);
        if (rc) return rc;
    } else {
        rc = launch_decode_exact(c, frame, frame_bytes, end_bit, d_tab, reinterpret_cast<uint16_t *>(d_out), work, nch, 0, nch);
        if (rc) return rc;
    }
    HF_PROF(c, "dec_result_kernel"); dec_result_kernel<<<1, 1, 0, c->stream>>>(work, d_result);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// ---- a rank's byte range in two phases around the hand-over collective (sharded.cu) ------------------------------
// Phase A needs nothing from the other ranks: the whole range is synchronised SPECULATIVELY (every span from a guess,
// the chunk chain repaired on the assumption that the stream re-synchronises), which already yields what the next
// rank waits for, the overflow of the range's last code word past its end (d_probe[0]; d_probe[1] = the range's bits).
// Phase B runs once the predecessor's overflow is known ON THE DEVICE (*d_first_bit, bits from the range's first byte):
// the head of the range is repaired from that bit, the chunk counts are scanned and the symbols written.  Round 1 ran a
// separate speculative pass over the range's tail for the hand-over and then the exact pass; here the one pass serves both.
struct RangeGeom {
    const uint8_t *frame;
    unsigned long long lead, frame_bytes, end_bit, nch;
};
static int range_geom(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes, RangeGeom *g)
{
    g->frame = reinterpret_cast<const uint8_t *>((uintptr_t)d_range & ~(uintptr_t)15);
    g->lead = (unsigned long long)((uintptr_t)d_range & 15);
    g->frame_bytes = g->lead + range_bytes + halo_bytes;
    g->end_bit = (g->lead + range_bytes) * 8;
    g->nch = (g->end_bit + CHUNK_BITS - 1) / CHUNK_BITS;
    if (g->nch == 0) g->nch = 1;
    if (g->nch > 0x7FFFFFFFull) return set_err(c, HF_ERR_ARG, "sharded decode: range too large");
    return ensure_ws(c, WS_STAGE_OFFSET + DecLayout::bytes(g->nch));
}

__global__ void range_probe_kernel(const DecWork *__restrict__ work, unsigned long long range_bits,
                                   unsigned long long *__restrict__ probe)
{
    probe[0] = work->result[1];
    probe[1] = range_bits;
}

int launch_range_sync(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes, const DecodeTable *d_tab,
                      unsigned long long *d_probe)
{
    RangeGeom g;
    int rc = range_geom(c, d_range, range_bytes, halo_bytes, &g);
    if (rc) return rc;
    DecWork *work = reinterpret_cast<DecWork *>((uint8_t *)c->ws + WS_STAGE_OFFSET);
    HF_CUDA(c, cudaMemsetAsync(work, 0, sizeof(DecWork), c->stream));
    if (range_bytes) {
        rc = launch_sync2(c, g.frame, g.frame_bytes, g.end_bit, d_tab, work, g.nch, 0, g.nch, false, true);
        if (rc) return rc;
        rc = launch_fix2(c, g.frame, g.frame_bytes, g.end_bit, d_tab, work, g.nch, 0, g.nch, true);
        if (rc) return rc;
    }
    range_probe_kernel<<<1, 1, 0, c->stream>>>(work, range_bytes * 8ull, d_probe);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_range_write(Ctx *c, const uint8_t *d_range, uint64_t range_bytes, uint64_t halo_bytes,
                       const unsigned long long *d_first_bit, const DecodeTable *d_tab, uint8_t *d_out, uint64_t out_symbols,
                       unsigned long long *d_result)
{
    if ((uintptr_t)d_out & 1) return set_err(c, HF_ERR_ARG, "sharded decode: output must be 2-byte aligned");
    RangeGeom g;
    int rc = range_geom(c, d_range, range_bytes, halo_bytes, &g);       // the same geometry (and workspace) as phase A
    if (rc) return rc;
    DecWork *work = reinterpret_cast<DecWork *>((uint8_t *)c->ws + WS_STAGE_OFFSET);
    dec_start_range_kernel<<<1, 1, 0, c->stream>>>(work, d_first_bit, g.lead * 8, out_symbols);
    HF_LAUNCH_CHECK(c);
    if (range_bytes) {
        rc = launch_fix_head(c, g.frame, g.frame_bytes, g.end_bit, d_tab, work, g.nch);
        if (rc) return rc;
        rc = launch_scan(c, work, g.nch, 0, g.nch);
        if (rc) return rc;
        rc = launch_write2(c, g.frame, g.frame_bytes, d_tab, work, g.nch, 0, g.nch, reinterpret_cast<uint16_t *>(d_out), false);
        if (rc) return rc;
    }
    HF_PROF(c, "dec_result_kernel"); dec_result_kernel<<<1, 1, 0, c->stream>>>(work, d_result);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
