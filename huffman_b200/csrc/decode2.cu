// decode2.cu — the word-walk kernels of the exact decoder: dec_sync2_kernel (code word boundaries and
// symbol counts per 256-bit subsequence) and dec_write2_kernel (symbols out).  They replace the two
// hot kernels of decode.cu (dec_sync_kernel, dec_write_kernel) and keep their global interface
// (DecLayout: info / chunkCnt / chunkE), so the inter-chunk repair and the scan of decode.cu stay.
//
// The reference decodes on the host, one bit and one fread() per step, by chasing tree pointers
// (/root/reference/Decompressor.cu:259-291); its format has no offset index.
//
// What is different from decode.cu's first kernels (ncu: 76 + 65 thread instructions per symbol, 39 %
// of the lanes active in the synchronisation kernel, 55 % of the symbols of the 16 GiB bench stream
// looked up in global memory):
//   * a thread keeps its 256-bit subsequence (+ one look-ahead word) in REGISTERS, loaded straight
//     from global memory; the walk is unrolled over the 8 words, so the 32-bit window of a code word
//     is ONE funnel shift of two registers — no shared-memory staging of the payload, no bit reader;
//   * the tables are direct and much deeper: a 64 KiB length plane indexed by 16 bits (the
//     synchronisation walks need no symbols) and a 128 KiB (symbol, length) plane indexed by 15 bits,
//     each in the shared memory of a persistent CTA; longer codes take ONE gather from a flat
//     second-level plane (<= 22 bits, L2 resident); anything beyond goes through decode.cu's tables;
//   * re-synchronisation is detected at four checkpoints per subsequence (the first code word boundary
//     at or after bits 0/64/128/192, with the symbol count of every 64-bit segment), which costs two
//     instructions per checkpoint instead of a 256-bit boundary mask maintained per symbol.
//
// Algorithmic bytes: dec_sync2 reads C; dec_write2 reads C and writes N.
#include "common.cuh"
#include "decode_common.cuh"

namespace hf {

constexpr uint32_t SYNC_K = 16;                                 // index bits of the length plane
constexpr uint32_t WRITE_K = 15;                                // index bits of the (symbol, length) plane
constexpr int W2_THREADS = 2 * DEC_THREADS;                     // a pair of chunks per step
constexpr uint32_t W2_WIN = 40960;                              // output staging window (symbols)
constexpr size_t S2_SMEM = 1u << SYNC_K;
constexpr size_t W2_SMEM = (4u << WRITE_K) + (W2_WIN + 8) * 2;

// ---- planes ------------------------------------------------------------------------------------
// (sym << 8) | len of the code word that is a prefix of the left-aligned window, from t1 / t2; 0 when
// the window holds no complete code word of at most K1 + sub bits (hole, or a long-list code)
__device__ __forceinline__ uint32_t lookup_win32(const DecodeTable *tab, uint32_t win)
{
    uint32_t e = tab->t1[win >> (32 - K1)];
    if (e & E_SUB) {
        const uint32_t sb = e & 31u;
        const uint32_t idx2 = sb ? (win << K1) >> (32 - sb) : 0u;
        e = tab->t2[(e >> 8) + idx2];
        if (e & (E_LIST | E_SUB)) e = 0;
    }
    return e;
}

__global__ void dt_planes_kernel(DecodeTable *__restrict__ tab)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t k2 = tab->maxlen;
    k2 = k2 < SYNC_K ? SYNC_K : (k2 > FLAT_MAX ? FLAT_MAX : k2);
    if (i == 0) tab->k2 = k2;
    if (i < (1u << SYNC_K)) {
        const uint32_t e = lookup_win32(tab, i << (32 - SYNC_K));
        tab->len16[i] = (e & 0x7Fu) <= SYNC_K ? (uint8_t)(e & 0x7Fu) : (uint8_t)0;
    }
    if (i < (1u << WRITE_K)) {
        const uint32_t e = lookup_win32(tab, i << (32 - WRITE_K));
        tab->lut15[i] = (e & 0x7Fu) <= WRITE_K ? e : 0u;
    }
    if (i < (1u << k2)) {
        const uint32_t e = lookup_win32(tab, i << (32 - k2));
        const bool ok = e && (e & 0x7Fu) <= k2;
        tab->flat2[i] = ok ? e : 0u;
        tab->lenflat[i] = ok ? (uint8_t)(e & 0x7Fu) : (uint8_t)0;
    }
}

int launch_table_planes(Ctx *c, DecodeTable *d_tab)
{
    HF_PROF(c, "dt_planes_kernel"); dt_planes_kernel<<<(1u << FLAT_MAX) / 256, 256, 0, c->stream>>>(d_tab);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// ---- the rare way: any code length, bits straight from global memory ----------------------------
// returns (sym << 8) | len, len >= 1; bit 31 set when the bits are no code word (a hole in the code)
constexpr uint32_t SLOW_BAD = 0x80000000u;
__device__ __noinline__ uint32_t slow_decode(const DecodeTable *tab, const uint8_t *frame, unsigned long long frame_bytes,
                                             unsigned long long bit)
{
    GlobalFetch f{frame, frame_bytes, bit >> 5};
    const unsigned long long w64 = peek64(f, (uint32_t)(bit & 31));
    const uint32_t win = (uint32_t)(w64 >> 32);
    uint32_t e = tab->t1[win >> (32 - K1)];
    if (e & E_SUB) {
        const uint32_t sb = e & 31u;
        const uint32_t idx2 = sb ? (win << K1) >> (32 - sb) : 0u;
        e = tab->t2[(e >> 8) + idx2];
        if (e & E_LIST) {
            uint32_t cur = e;
            e = 0;
            while (cur & E_LIST) {
                const LongCode lc = tab->longs[(cur >> 8) & 0xFFFFu];
                if (((w64 ^ lc.code_left) >> (64 - (lc.leaf & 0x7Fu))) == 0) { e = lc.leaf; break; }
                cur = lc.next;
            }
        }
    }
    if (e == 0) e = SLOW_BAD | 1u;      // hole in the code: flag it, step one bit so the walk ends
    return e;
}

// ---- a thread's subsequence in registers --------------------------------------------------------
// r[0..7]: the 8 big-endian words of subsequence (c, t); r[8]: the first word of the next one
__device__ __forceinline__ void load_sub(uint32_t (&r)[9], const uint8_t *frame, unsigned long long frame_bytes,
                                         unsigned long long c, uint32_t t, uint32_t lane)
{
    const unsigned long long b = c * (CHUNK_BITS / 8) + (unsigned long long)t * (SUB_BITS / 8);
    uint4 a = make_uint4(0, 0, 0, 0), d = make_uint4(0, 0, 0, 0);
    if (b < frame_bytes) a = __ldg(reinterpret_cast<const uint4 *>(frame + b));          // the frame is 16-byte aligned
    if (b + 16 < frame_bytes) d = __ldg(reinterpret_cast<const uint4 *>(frame + b + 16));
    r[0] = bswap32(a.x); r[1] = bswap32(a.y); r[2] = bswap32(a.z); r[3] = bswap32(a.w);
    r[4] = bswap32(d.x); r[5] = bswap32(d.y); r[6] = bswap32(d.z); r[7] = bswap32(d.w);
    uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, r[0], 1);
    if (lane == 31) {
        nx = 0;
        if (b + 32 < frame_bytes) nx = bswap32(__ldg(reinterpret_cast<const uint32_t *>(frame + b + 32)));
    }
    r[8] = nx;
}

// Walks the code words of a subsequence from bit `start` to `lim`.  chkpos: byte k = (first code word
// boundary at or after bit 64k) - 64k; chkcnt: byte k = code words starting in [64k, 64k + 64).
// RESYNC: chkpos / chkcnt describe an earlier walk; stop at the first checkpoint both walks share and
// keep the earlier walk's record from there on (`end` is then the earlier walk's).  CHK_NONE marks "no
// earlier walk": its bytes (0xFF) equal no checkpoint offset (those are below 192).
constexpr uint32_t CHK_NONE = 0xFFFFFFFFu;
template <bool RESYNC>
__device__ __forceinline__ void walk_len(const uint32_t (&r)[9], const uint8_t *s_len, const DecodeTable *tab,
                                         const uint8_t *frame, unsigned long long frame_bytes, unsigned long long sub_bit0,
                                         uint32_t k2shift, uint32_t start, uint32_t lim, uint32_t &chkpos,
                                         uint32_t &chkcnt, uint32_t &end, uint32_t &bad)
{
    uint32_t pos = start, n = 0, npos = 0, ncnt = 0, wl = lim, keep = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if ((w & 1) == 0) {
            const int k = w >> 1;
            if (k > 0) { ncnt |= n << (8 * (k - 1)); n = 0; }
            const uint32_t rel = (pos - 32u * w) & 0xFFu;
            if (RESYNC && k > 0) {       // still walking, and on the earlier walk's boundary: the walks have met
                if (pos < wl && rel == ((chkpos >> (8 * k)) & 0xFFu)) { wl = 0; keep = 0xFFFFFFFFu << (8 * k); }
            }
            npos |= rel << (8 * k);
        }
        const uint32_t lw = min(wl, 32u * (w + 1));
        while (pos < lw) {
            const uint32_t win = __funnelshift_l(r[w + 1], r[w], pos);
            uint32_t len = s_len[win >> (32 - SYNC_K)];
            if (len == 0) {
                len = __ldg(tab->lenflat + (win >> k2shift));
                if (len == 0) {
                    const uint32_t e = slow_decode(tab, frame, frame_bytes, sub_bit0 + pos);
                    bad |= e >> 31;
                    len = e & 0x7Fu;
                }
            }
            pos += len;
            n++;
        }
    }
    ncnt |= n << 24;
    if (RESYNC && keep) {
        chkpos = (npos & ~keep) | (chkpos & keep);
        chkcnt = (ncnt & ~keep) | (chkcnt & keep);
    } else {
        chkpos = npos;
        chkcnt = ncnt;
        end = pos - lim;
    }
}

// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_THREADS, 3)
dec_sync2_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes, unsigned long long F0,
                 unsigned long long range_end_bit, const DecodeTable *__restrict__ tab, DecWork *work,
                 unsigned long long nch, unsigned long long c_first, unsigned long long c_last, uint32_t speculative,
                 const unsigned long long *gate)
{
    if (gate && !(*gate & DF_GATE_MASK)) return;        // the single-pass decoder succeeded
    extern __shared__ __align__(16) uint8_t s_len[];    // 2^SYNC_K
    __shared__ uint32_t s_end[DEC_THREADS];
    __shared__ uint32_t s_red[DEC_THREADS / 32];
    if (tab->single_sym) return;                        // empty payload, see dec_fill_kernel
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab->len16);
        uint4 *dst = reinterpret_cast<uint4 *>(s_len);
        for (uint32_t i = tid; i < S2_SMEM / 16; i += DEC_THREADS) dst[i] = __ldg(src + i);
    }
    const uint32_t g = speculative ? 1u : tab->len_gcd;
    const uint32_t k2shift = 32u - tab->k2;
    const uint32_t sub0 = tid * SUB_BITS;
    uint32_t bad = 0;

    for (unsigned long long c = c_first + blockIdx.x; c < c_last; c += gridDim.x) {
        __syncthreads();                                // planes loaded / s_end, s_red reuse
        uint32_t r[9];
        load_sub(r, frame, frame_bytes, c, tid, lane);
        const unsigned long long X = c * CHUNK_BITS + sub0;
        const uint32_t lim = sub_limit(c, tid, range_end_bit);      // code words starting at or after the range end are not ours
        const bool fixed = (c == 0 && tid == 0 && !speculative);    // holds the first payload bit: exact start
        uint32_t p = fixed ? (uint32_t)F0 : (X >= F0 ? spec_start(X, F0, g) : 0u);
        uint32_t end = 0, chkpos = CHK_NONE, chkcnt = 0;
        if (lim) {
            if (p < lim) walk_len<false>(r, s_len, tab, frame, frame_bytes, X, k2shift, p, lim, chkpos, chkcnt, end, bad);
            else end = p - lim;
        }
        // fix-point: my true start is my predecessor's overflow
        for (uint32_t it = 0; it < DEC_THREADS + 1; it++) {
            s_end[tid] = end;
            __syncthreads();
            int changed = 0;
            if (tid > 0 && !fixed && lim) {
                const uint32_t q = s_end[tid - 1];
                if (q != p) {
                    const uint32_t old_end = end;
                    if (q < lim) walk_len<true>(r, s_len, tab, frame, frame_bytes, X, k2shift, q, lim, chkpos, chkcnt, end, bad);
                    else { chkpos = CHK_NONE; chkcnt = 0; end = q - lim; }
                    changed = end != old_end;
                    p = q;
                }
            }
            if (!__syncthreads_or(changed)) break;
        }
        const uint32_t cnt = __dp4a(chkcnt, 0x01010101u, 0u);
        L.info[c * DEC_THREADS + tid] = (uint16_t)((p & 63u) | (cnt << 6));
        uint32_t v = cnt;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) s_red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int i = 0; i < DEC_THREADS / 32; i++) tot += s_red[i];
            L.chunkCnt[c] = tot;
            L.chunkE2[c] = 0xFFFFFFFFu;
        }
        if (tid == DEC_THREADS - 1) L.chunkE[c] = end;
        // the thread whose subsequence holds the end of the range reports the overflow past it
        if (lim && sub_limit(c, tid + 1, range_end_bit) == 0) work->result[1] = end;
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(W2_THREADS, 1)
dec_write2_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes, unsigned long long F0,
                  const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch,
                  unsigned long long n_symbols, uint16_t *__restrict__ out, const unsigned long long *gate)
{
    if (gate && !(*gate & DF_GATE_MASK)) return;
    extern __shared__ __align__(16) uint32_t w2_smem[];
    uint32_t *s_lut = w2_smem;                                              // 2^WRITE_K
    uint16_t *sout = reinterpret_cast<uint16_t *>(s_lut + (1u << WRITE_K)); // W2_WIN + 8
    __shared__ uint32_t s_w[33];
    if (tab->single_sym) return;
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab->lut15);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t i = tid; i < (4u << WRITE_K) / 16; i += W2_THREADS) dst[i] = __ldg(src + i);
    }
    const uint32_t k2shift = 32u - tab->k2;
    const uint32_t half = tid / DEC_THREADS, t = tid % DEC_THREADS;
    const uint32_t sub0 = t * SUB_BITS;
    uint32_t bad = 0;
    const unsigned long long npair = (nch + 1) / 2;

    for (unsigned long long cp = blockIdx.x; cp < npair; cp += gridDim.x) {
        __syncthreads();                                // planes loaded / sout, s_w reuse
        const unsigned long long c = 2 * cp + half;
        const bool have = c < nch;
        const unsigned long long base = L.chunkBase[2 * cp];
        if (base >= n_symbols) continue;                // uniform over the CTA
        const uint32_t inf = have ? L.info[c * DEC_THREADS + t] : 0u;
        const uint32_t cnt = inf >> 6;
        uint32_t pos = (c == 0 && t == 0) ? (uint32_t)F0 : (inf & 63u);     // the stream head may sit past bit 63
        uint32_t r[9];
        load_sub(r, frame, frame_bytes, have ? c : 0ull, t, lane);

        // exclusive scan of the counts over the pair
        uint32_t x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            const uint32_t s = s_w[lane];
            uint32_t v = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += y; }
            s_w[lane] = v - s;
            if (lane == 31) s_w[32] = v;
        }
        __syncthreads();
        const uint32_t off = x - cnt + s_w[wid];        // pair-relative index of my first symbol
        unsigned long long total = s_w[32];
        if (base + total > n_symbols) total = n_symbols - base;   // garbage past the payload end is dropped
        const uint32_t my_end = (uint32_t)min((unsigned long long)(off + cnt), total);
        uint32_t o = off;                               // pair-relative index of my next symbol

        const uint32_t mis = (uint32_t)(base & 7);      // staging slot j <-> output symbol base - mis + j
        for (unsigned long long w0 = 0; w0 < total; w0 += W2_WIN) {
            const uint32_t wend = (uint32_t)min(total, w0 + W2_WIN);
            const uint32_t o_end = min(my_end, wend);
            if (o < o_end) {
                uint16_t *sp = sout + (o - (uint32_t)w0 + mis);
                uint16_t *const sp_end = sout + (o_end - (uint32_t)w0 + mis);
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint32_t lw = 32u * (w + 1);
                    while (pos < lw && sp < sp_end) {
                        const uint32_t win = __funnelshift_l(r[w + 1], r[w], pos);
                        uint32_t e = s_lut[win >> (32 - WRITE_K)];
                        if (e == 0) {
                            e = __ldg(tab->flat2 + (win >> k2shift));
                            if (e == 0) {
                                e = slow_decode(tab, frame, frame_bytes, c * CHUNK_BITS + sub0 + pos);
                                bad |= e >> 31;
                            }
                        }
                        *sp++ = (uint16_t)(e >> 8);
                        pos += e & 0x7Fu;
                    }
                }
                o = o_end;
            }
            __syncthreads();
            // flush [w0, wend): staging slots [mis, mis + n)
            const uint32_t n = wend - (uint32_t)w0;
            uint16_t *dst = out + base + w0 - mis;      // 16-byte aligned when out is
            const uint32_t nvec = (mis + n + 7) / 8;
            for (uint32_t q = tid; q < nvec; q += W2_THREADS) {
                const uint32_t j0 = q * 8;
                if (j0 >= mis && j0 + 8 <= mis + n && (((uintptr_t)(dst + j0) & 15) == 0)) {
                    st_stream_v4(dst + j0, reinterpret_cast<const uint4 *>(sout)[q]);
                } else {
                    for (uint32_t j = j0; j < j0 + 8; j++)
                        if (j >= mis && j < mis + n) dst[j] = sout[j];
                }
            }
            __syncthreads();
        }
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// -------------------------------------------------------------------------------------------------
int launch_sync2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long F0,
                 unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                 unsigned long long c_first, unsigned long long c_last, uint32_t speculative, const unsigned long long *gate)
{
    static bool attr = false;
    if (!attr) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_sync2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_SMEM));
        HF_CUDA(c, cudaFuncSetAttribute(dec_write2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W2_SMEM));
        attr = true;
    }
    unsigned long long grid = c_last - c_first;
    if (grid > (unsigned long long)(3 * c->sm_count)) grid = 3 * c->sm_count;
    HF_PROF(c, "dec_sync2_kernel");
    dec_sync2_kernel<<<(unsigned)grid, DEC_THREADS, S2_SMEM, c->stream>>>(frame, frame_bytes, F0, range_end_bit, d_tab, work, nch,
                                                                        c_first, c_last, speculative, gate);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_write2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long F0,
                  const DecodeTable *d_tab, DecWork *work, unsigned long long nch, unsigned long long n_symbols,
                  uint16_t *out, const unsigned long long *gate)
{
    static bool attr = false;
    if (!attr) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_write2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W2_SMEM));
        attr = true;
    }
    unsigned long long grid = (nch + 1) / 2;
    if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "dec_write2_kernel");
    dec_write2_kernel<<<(unsigned)grid, W2_THREADS, W2_SMEM, c->stream>>>(frame, frame_bytes, F0, d_tab, work, nch, n_symbols, out, gate);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
