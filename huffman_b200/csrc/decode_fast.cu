// decode_fast.cu — single-pass self-synchronising decoder.
//
// The reference decodes on the host, one bit and one fread() per step
// (/root/reference/Decompressor.cu:259-291); its format has no offset index, so a
// parallel decoder has to find the code word boundaries itself.  decode.cu does that exactly
// (decode to count, repair, scan, decode again to write).  This kernel does it in ONE pass over
// the payload and is the path normally taken; decode.cu stays as the exact fallback.
//
// Chunk = 512 subsequences of 256 bits (16 KiB), one per thread of a persistent 512-thread CTA
// (2 CTAs per SM); chunks are claimed in order from an atomic counter.
//   1. every thread decodes its subsequence from a guessed start INTO shared-memory slots
//      (symbols kept, not just counted) and records the boundaries it met (256-bit mask);
//   2. fix-point inside the CTA: a thread whose true start (the predecessor's overflow) differs
//      from its guess decodes from there only until it lands on a recorded boundary — a few code
//      words — and keeps the rest of its slot.  When the chunk has converged on its guessed start
//      it PUBLISHES the overflow of its last subsequence (511 subsequences of self-synchronisation
//      lie behind that value, so it does not depend on the guess) and takes the previous chunk's
//      published overflow as thread 0's true start: the inter-chunk repair is just more rounds of
//      the same fix-point, and no chunk waits for its predecessor to FINISH;
//   3. CTA scan of the symbol counts, decoupled look-back over the chunks for the output offset.
//      If a repair changes an overflow that was already published (a stream that does not
//      synchronise within a whole chunk) a flag is raised and the exact path of decode.cu redoes
//      the job;
//   4. slots are compacted through a shared staging window and leave with aligned 128-bit stores;
//      a thread with more symbols than its slot holds decodes a second time straight into the window.
// The payload is read once and every symbol is decoded once (plus the few re-synchronisation code
// words): algorithmic bytes C + N, traffic ~ C + N.
#include "common.cuh"
#include "decode_common.cuh"

namespace hf {

constexpr int DF_THREADS = 512;
constexpr uint32_t DF_S = 256;                                  // bits per subsequence
constexpr uint32_t DF_SW = DF_S / 32;
constexpr uint32_t DF_CHUNK_BITS = DF_THREADS * DF_S;           // 131,072 bits = 16 KiB
constexpr uint32_t DF_CAP = 40;                                 // symbols a slot holds
constexpr uint32_t DF_MARGIN = 8;                               // re-synchronisation prefix a slot holds
constexpr uint32_t DF_SLOT = DF_MARGIN + DF_CAP + 1;            // half-words; odd: lanes spread over the banks
constexpr uint32_t DF_PAD_WORDS = 8;
constexpr uint32_t DF_BITS_WORDS = DF_THREADS * DF_SW + DF_PAD_WORDS;   // 4104 words
constexpr uint32_t DF_WIN = 8192 - 8;                           // staging window (symbols)
constexpr uint32_t DF_SPIN_LIMIT = 1u << 26;

// descriptor of a chunk: state(2) | overflow E (7) | value (55)
constexpr uint32_t DF_ST_INVALID = 0, DF_ST_E = 1, DF_ST_AGG = 2, DF_ST_INCL = 3;
constexpr unsigned long long DF_VAL_MASK = (1ull << 55) - 1;

constexpr uint32_t DF_SW_PADDED = (smem_words_padded(DF_BITS_WORDS) + 3) & ~3u;
constexpr size_t DF_SMEM = (1u << K1) * 4 + DF_SW_PADDED * 4 + DF_THREADS * DF_SLOT * 2 + 8192 * 2;

// flags in result[3]
constexpr unsigned long long DF_F_MISMATCH = 1, DF_F_BAD = 4, DF_F_CAPACITY = 8, DF_F_SPIN = 16;

struct DfWork {
    unsigned long long counter;
    unsigned long long pad;
    unsigned long long result[4];           // unused, E_last, n_symbols, flags
    unsigned long long phase_cycles[8];     // HF_DF_TIMING builds: cycles thread 0 of every CTA spent per phase
    // followed by desc[nch]
};

#ifdef HF_DF_TIMING
#define DF_TICK(k) do { if (tid == 0) { long long _n = clock64(); atomicAdd(&P.work->phase_cycles[k], (unsigned long long)(_n - t_last)); t_last = _n; } } while (0)
#else
#define DF_TICK(k) do { } while (0)
#endif

struct DfParams {
    const uint8_t *frame;                   // 16-byte aligned
    long long hi_valid;                     // readable bytes from frame: [0, hi_valid)
    long long range_end_bit;                // code words starting at or after this frame bit are not ours
    uint32_t F0;                            // frame bit of the first code word (< 256)
    uint32_t tail_only;                     // 1: only the last chunk, speculatively, for result[1]; no output
    unsigned long long nch;
    unsigned long long out_limit;           // symbols the output may hold
    const DecodeTable *tab;
    DfWork *work;
    uint16_t *out;
};

__device__ __forceinline__ bool mask_test(const unsigned long long (&m)[4], uint32_t pos)
{
    unsigned long long w = pos < 128 ? (pos < 64 ? m[0] : m[1]) : (pos < 192 ? m[2] : m[3]);
    return (w >> (pos & 63)) & 1ull;
}

__device__ __forceinline__ uint32_t mask_count_below(const unsigned long long (&m)[4], uint32_t pos)
{
    uint32_t n = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; s++) {
        if (pos >= (s + 1) * 64) n += __popcll(m[s]);
        else if (pos > s * 64) n += __popcll(m[s] & ((1ull << (pos - s * 64)) - 1ull));
    }
    return n;
}

// decodes [p, lim) of the subsequence at sw bit `sub0` into slot[0..), recording the boundaries
__device__ __forceinline__ void df_decode_sub(const TabView &T, const uint32_t *sw, uint32_t sub0, uint32_t p,
                                              uint32_t lim, uint16_t *slot, uint32_t &cnt, uint32_t &end,
                                              unsigned long long (&mask)[4], uint32_t &bad)
{
    SmemFetch f{sw};
    BitReader<SmemFetch> r{f};
    r.init(sub0 + p);
    uint32_t pos = p, n = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; s++) {
        unsigned long long m = 0;
        const uint32_t seg_end = min(lim, (s + 1) * 64);
        while (pos < seg_end) {
            const uint32_t e = decode_one(T, r, sub0 + pos, bad);
            const uint32_t len = e & 0x7Fu;
            m |= 1ull << (pos & 63);
            if (n < DF_CAP) slot[n] = (uint16_t)(e >> 8);
            n++;
            pos += len;
            r.skip(len);
        }
        mask[s] = m;
    }
    cnt = n;
    end = pos - lim;
}

__global__ void __launch_bounds__(DF_THREADS, 2)
dec_fast_kernel(const DfParams P)
{
    extern __shared__ __align__(16) uint32_t df_smem[];
    uint32_t *st1 = df_smem;                                    // 2^K1
    uint32_t *sw = st1 + (1u << K1);                            // DF_BITS_WORDS
    uint16_t *slots = reinterpret_cast<uint16_t *>(sw + DF_SW_PADDED);
    uint16_t *sout = slots + DF_THREADS * DF_SLOT;              // 8192 staging symbols (16-byte aligned: see DF_SLOT)
    __shared__ uint32_t s_end[DF_THREADS];
    __shared__ uint32_t s_w[34];
    __shared__ unsigned long long s_bc[2];
    __shared__ unsigned long long s_chunk;
    __shared__ uint32_t s_prev, s_repair, s_espec;

    const DecodeTable *tab = P.tab;
    if (tab->single_sym) return;                                // zero-length code: dec_fill_kernel writes the output
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long *desc = reinterpret_cast<unsigned long long *>(P.work + 1);
    for (uint32_t i = tid; i < (1u << K1); i += DF_THREADS) st1[i] = tab->t1[i];
    const TabView T{st1, tab->t2, tab->longs, tab->n_long};
    const uint32_t g = tab->len_gcd;
    uint16_t *slot = slots + tid * DF_SLOT;                     // [0, MARGIN): re-sync prefix, [MARGIN, MARGIN + CAP): symbols
    unsigned long long flags = 0;
#ifdef HF_DF_TIMING
    long long t_last = clock64();
#endif

    for (;;) {
        __syncthreads();                                        // sw / sout / s_chunk reuse
        DF_TICK(7);
        if (tid == 0) s_chunk = P.tail_only ? (P.work->counter++ ? P.nch : P.nch - 1) : atomicAdd(&P.work->counter, 1ull);
        __syncthreads();
        const unsigned long long c = s_chunk;
        if (c >= P.nch) break;

        // ---- stage the chunk's bits (+ look-ahead) as big-endian words ----
        const long long chunk_bit0 = (long long)(c * DF_CHUNK_BITS);
        const long long byte0 = chunk_bit0 >> 3;                // multiple of 16
        for (uint32_t i = tid; i < DF_BITS_WORDS / 4; i += DF_THREADS) {
            const long long b = byte0 + 16ll * i;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (b < P.hi_valid) v = ld_stream_v4(P.frame + b);
            uint32_t *dst = sw + smem_word_index(4 * i);
            dst[0] = bswap32(v.x); dst[1] = bswap32(v.y); dst[2] = bswap32(v.z); dst[3] = bswap32(v.w);
        }

        // ---- 1. speculative decode of my subsequence ----
        const uint32_t sub0 = tid * DF_S;
        const long long subpos = chunk_bit0 + sub0;             // frame bit of my subsequence
        const bool fixed = c == 0 && tid == 0 && !P.tail_only;  // holds the first code word: exact start
        const long long room = P.range_end_bit - subpos;
        const uint32_t lim = room <= 0 ? 0u : (room >= (long long)DF_S ? DF_S : (uint32_t)room);
        const bool active = lim > 0;
        uint32_t p = 0, cnt = 0, end = 0, bad = 0, pre_k = 0, skip_j = 0;
        bool patched = false;
        unsigned long long mask[4] = {0, 0, 0, 0};
        if (active) p = fixed ? P.F0 : (subpos >= (long long)P.F0 ? spec_start((unsigned long long)subpos, P.F0, g) : 0u);
        if (tid == 0) { s_prev = p; s_repair = 0; }
        __syncthreads();
        DF_TICK(0);                                             // claim + stage
        if (active) {
            if (p < lim) df_decode_sub(T, sw, sub0, p, lim, slot + DF_MARGIN, cnt, end, mask, bad);
            else end = p - lim;                                 // a code word longer than what is left of the subsequence
        }

        // ---- 2. fix-point: my true start is my predecessor's overflow; thread 0's predecessor is the
        //         previous chunk, whose overflow arrives through its descriptor ----
        bool have_prev = c == 0 || P.tail_only;
        for (uint32_t it = 0; it < 4 * DF_THREADS; it++) {
            s_end[tid] = end;
            __syncthreads();
            if (it == 0) DF_TICK(1);                            // first decode (until the slowest warp is done)
#ifdef HF_DF_TIMING
            if (tid == 0) atomicAdd(&P.work->phase_cycles[6], 1ull);     // fix-point rounds
#endif
            int changed = 0;
            if (active && !fixed) {
                const uint32_t q = tid ? s_end[tid - 1] : s_prev;
                if (q != p) {
                    const uint32_t old_end = end;
                    bool done = false;
                    if (!patched && q < lim) {
                        // decode from q until a boundary of the recorded walk (or the end of the subsequence)
                        SmemFetch f{sw};
                        BitReader<SmemFetch> r{f};
                        r.init(sub0 + q);
                        uint32_t pos = q, k = 0;
                        bool hit = false;
                        while (pos < lim) {
                            if (pos >= p && mask_test(mask, pos)) { hit = true; break; }
                            if (k == DF_MARGIN) break;
                            const uint32_t e = decode_one(T, r, sub0 + pos, bad);
                            slot[k++] = (uint16_t)(e >> 8);
                            pos += e & 0x7Fu;
                            r.skip(e & 0x7Fu);
                        }
                        if (hit) {
                            pre_k = k; skip_j = mask_count_below(mask, pos); patched = true; done = true;
                        } else if (pos >= lim) {                // ran off the subsequence: the prefix is everything
                            pre_k = k; skip_j = cnt; end = pos - lim; patched = true; done = true;
                        }
                    }
                    if (!done) {                                // start over from q
                        pre_k = 0; skip_j = 0; patched = false; cnt = 0;
                        mask[0] = mask[1] = mask[2] = mask[3] = 0;
                        if (q < lim) df_decode_sub(T, sw, sub0, q, lim, slot + DF_MARGIN, cnt, end, mask, bad);
                        else end = q - lim;
                    }
                    p = q;
                    changed = end != old_end;
                }
            }
            if (__syncthreads_or(changed)) continue;
            if (have_prev) break;
            // converged on the guessed chunk start: publish my overflow (robust: 512 subsequences of
            // self-synchronisation lie behind it), then take the predecessor's as thread 0's true start
            if (tid == 0) {
                const uint32_t E = s_end[DF_THREADS - 1] & 0x7Fu;
                s_espec = E;
                st_release_u64(&desc[c], ((unsigned long long)DF_ST_E << 62) | ((unsigned long long)E << 55));
                unsigned long long d;
                uint32_t spins = 0;
                while (((d = ld_acquire_u64(&desc[c - 1])) >> 62) == DF_ST_INVALID)
                    if (++spins > DF_SPIN_LIMIT) { flags |= DF_F_SPIN; break; }
                s_prev = (uint32_t)((d >> 55) & 0x7Fu);
                s_repair = s_prev != p;
            }
            have_prev = true;
            __syncthreads();
            if (!s_repair) break;
        }
        if (c == 0 && !P.tail_only && tid == 0) {
            const uint32_t E = s_end[DF_THREADS - 1] & 0x7Fu;
            s_espec = E;
            st_release_u64(&desc[0], ((unsigned long long)DF_ST_E << 62) | ((unsigned long long)E << 55));
        }
        DF_TICK(2);                                             // fix-point
        // the thread whose subsequence holds the end of the range reports the overflow past it
        if (active && room <= (long long)DF_S) P.work->result[1] = end;
        if (P.tail_only) break;
        if (active && bad) flags |= DF_F_BAD;

        // ---- 3. counts -> CTA scan -> look-back ----
        const uint32_t n_mine = active ? pre_k + cnt - skip_j : 0u;
        uint32_t x = n_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = lane < DF_THREADS / 32 ? s_w[lane] : 0u, t = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
            if (lane < DF_THREADS / 32) s_w[lane] = t - s;
            if (lane == 31) s_w[32] = t;
        }
        __syncthreads();
        const uint32_t off = x - n_mine + s_w[wid];             // chunk-relative index of my first symbol
        const uint32_t total = s_w[32];
        if (wid == 0) {
            const unsigned long long E = s_end[DF_THREADS - 1] & 0x7Fu;   // final overflow of the chunk
            // a successor may already have used the overflow published before the repair
            if (lane == 0 && c + 1 < P.nch && E != s_espec) flags |= DF_F_MISMATCH;
            unsigned long long prefix = 0;
            if (c == 0) {
                if (lane == 0) st_release_u64(&desc[0], ((unsigned long long)DF_ST_INCL << 62) | (E << 55) | total);
            } else {
                if (lane == 0) st_release_u64(&desc[c], ((unsigned long long)DF_ST_AGG << 62) | (E << 55) | total);
                long long look = (long long)c - 1;
                uint32_t spins = 0;
                for (;;) {
                    const long long idx = look - (long long)lane;
                    unsigned long long d = idx >= 0 ? ld_acquire_u64(&desc[idx]) : ((unsigned long long)DF_ST_INCL << 62);
                    const uint32_t st = (uint32_t)(d >> 62);
                    const uint32_t m_incl = __ballot_sync(0xFFFFFFFFu, st == DF_ST_INCL);
                    const uint32_t m_inv = __ballot_sync(0xFFFFFFFFu, st < DF_ST_AGG);
                    const uint32_t first = m_incl ? (uint32_t)__ffs(m_incl) - 1 : 32u;
                    const uint32_t need = first < 32 ? ((2u << first) - 1u) : 0xFFFFFFFFu;
                    if (m_inv & need) {
                        if (++spins > DF_SPIN_LIMIT) { flags |= DF_F_SPIN; break; }
                        continue;
                    }
                    unsigned long long v = (lane <= first) ? (d & DF_VAL_MASK) : 0ull;
#pragma unroll
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                    prefix += v;
                    if (first < 32) break;
                    look -= 32;
                }
                if (lane == 0)
                    st_release_u64(&desc[c], ((unsigned long long)DF_ST_INCL << 62) | (E << 55) | (prefix + total));
            }
            if (lane == 0) {
                s_bc[0] = prefix;
                if (c + 1 == P.nch) P.work->result[2] = prefix + total;
            }
        }
        __syncthreads();
        DF_TICK(3);                                             // scan + look-back
        const unsigned long long base = s_bc[0];
        unsigned long long tot = total;
        if (base >= P.out_limit) tot = 0;
        else if (base + tot > P.out_limit) tot = P.out_limit - base;      // padding garbage past the last symbol

        // ---- 4. compact the slots through the staging window, aligned 128-bit stores ----
        const uint32_t mis = (uint32_t)(base & 7);              // staging slot j <-> output symbol base - mis + j
        const bool over = cnt > DF_CAP;                         // slot too small: decode again, straight into the window
        SmemFetch rf{sw};
        BitReader<SmemFetch> rr{rf};
        uint32_t rpos = p, rnext = 0;
        if (over && p < lim) rr.init(sub0 + p);
        for (unsigned long long w0 = 0; w0 < tot; w0 += DF_WIN) {
            const uint32_t wn = (uint32_t)min((unsigned long long)DF_WIN, tot - w0);
            // my symbols [i0, i1) fall into this window
            const long long lo = (long long)w0 - (long long)off, hi = lo + wn;
            const uint32_t i0 = lo < 0 ? 0u : (uint32_t)min((long long)n_mine, lo);
            const uint32_t i1 = hi < 0 ? 0u : (uint32_t)min((long long)n_mine, hi);
            const int dbase = (int)mis + (int)off - (int)w0;    // my symbol i lands at sout[dbase + i]
            if (!over) {
                for (uint32_t i = i0; i < i1; i++)
                    sout[dbase + (int)i] = i < pre_k ? slot[i] : slot[DF_MARGIN + skip_j + (i - pre_k)];
            } else {
                while (rnext < i1) {
                    const uint32_t e = decode_one(T, rr, sub0 + rpos, bad);
                    if (rnext >= i0) sout[dbase + (int)rnext] = (uint16_t)(e >> 8);
                    rpos += e & 0x7Fu;
                    rr.skip(e & 0x7Fu);
                    rnext++;
                }
            }
            __syncthreads();
            DF_TICK(4);                                         // compaction
            uint16_t *dst = P.out + base + w0 - mis;            // 16-byte aligned when out is
            const uint32_t nvec = (mis + wn + 7) / 8;
            for (uint32_t q = tid; q < nvec; q += DF_THREADS) {
                const uint32_t j0 = q * 8;
                if (j0 >= mis && j0 + 8 <= mis + wn && (((uintptr_t)(dst + j0) & 15) == 0)) {
                    st_stream_v4(dst + j0, reinterpret_cast<const uint4 *>(sout)[q]);
                } else {
                    for (uint32_t j = j0; j < j0 + 8; j++)
                        if (j >= mis && j < mis + wn) dst[j] = sout[j];
                }
            }
            __syncthreads();
            DF_TICK(5);                                         // flush
        }
        if (base + total > P.out_limit && c + 1 < P.nch) flags |= DF_F_CAPACITY;
    }
    if (flags) atomicOr(&P.work->result[3], flags);
}

size_t df_work_bytes(unsigned long long nch) { return sizeof(DfWork) + nch * 8; }

unsigned long long df_chunks(unsigned long long range_end_bit)
{
    unsigned long long n = (range_end_bit + DF_CHUNK_BITS - 1) / DF_CHUNK_BITS;
    return n ? n : 1;
}

// enqueues the single-pass decode of frame bits [F0, range_end_bit); *work is zeroed here
int launch_decode_fast(Ctx *c, const uint8_t *frame, long long hi_valid, long long range_end_bit, uint32_t F0,
                       unsigned long long out_limit, const DecodeTable *d_tab, uint16_t *out, void *work_mem,
                       unsigned long long nch, bool tail_only)
{
    DfWork *work = reinterpret_cast<DfWork *>(work_mem);
    HF_CUDA(c, cudaMemsetAsync(work, 0, tail_only ? sizeof(DfWork) : df_work_bytes(nch), c->stream));
    static bool attr = false;
    if (!attr) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM));
        attr = true;
    }
    DfParams P;
    P.frame = frame; P.hi_valid = hi_valid; P.range_end_bit = range_end_bit;
    P.F0 = F0; P.tail_only = tail_only ? 1u : 0u; P.nch = nch; P.out_limit = out_limit; P.tab = d_tab; P.work = work; P.out = out;
    unsigned long long grid = nch < (unsigned long long)(2 * c->sm_count) ? nch : (unsigned long long)(2 * c->sm_count);
    if (tail_only) grid = 1;
    HF_PROF(c, tail_only ? "decode_tail_kernel" : "decode_kernel"); dec_fast_kernel<<<(unsigned)grid, DF_THREADS, DF_SMEM, c->stream>>>(P);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
