// encode.cu — Huffman encoder: bit offsets by a counting pre-pass over SEGMENTS, then one packing
// pass in which every CTA works on its own segments with no inter-CTA dependency at all.
//
// Replaces populateCWLength + thrust::transform_inclusive_scan + encodeFromCW + the host
// tail flush (/root/reference/Compressor.cu:50-74, :152-313, :541-601, :673-684): the
// reference materialises 12 bytes of scratch per symbol and then binary-searches the
// offsets once per OUTPUT byte.
//
// enc_count_kernel   the input is cut into ~8 segments per SM (whole tiles).  Each CTA sums the code
//     lengths of its segments with the 64 KiB length plane in shared memory; the CTA that finishes
//     last turns the per-segment bit counts into start offsets (exclusive scan).  Reads N.
// encode_kernel      one persistent 768-thread CTA per SM claims segments from an atomic counter and
//     walks the tiles (12,288 symbols) of a segment in order with a running bit offset:
//       1. 2 x 128-bit loads per thread (8 symbols each; the next tile is prefetched while this one
//          is packed), code lookup, length sums, CTA-wide exclusive scan;
//       2. every thread streams its codes through a 64-bit accumulator (branch-free flush) into the
//          shared staging window, 8192 words at the tile's 128-bit phase; a tile whose bits exceed
//          the window (more than ~21 bits per symbol) is packed window by window;
//       3. a 32-bit word belongs to the tile that holds its first bit: the owner completes its last,
//          partial word by encoding the symbols that FOLLOW the tile until the word is full, so tiles
//          and segments exchange nothing; windows leave with aligned 128-bit stores.
//     Reads N, writes C.
// An earlier version resolved the offsets in the packing kernel itself (decoupled look-back); with
// 4096-symbol tiles in flight on every SM the look-back walks cost more than the extra read.
//
// The code table of the 65,536-symbol alphabet lives in SHARED memory: 24-bit entries
// (1 << len) | code  for len <= 23 — the leading one bit carries the length — stored as two
// planes (u16 low part, u8 high part: 192 KiB of the SM's 227 KiB) and indexed by the symbol
// with its low byte XOR-folded with its high byte, so that skewed first bytes (text, Zipf) still
// spread over the 32 banks.  Gathers from shared memory cost bank conflicts only; the same gathers
// from global memory cost one L1 wavefront per distinct line.  Entry 0 means "longer than 23 bits"
// (or absent): (len, code) then come from the global codebook.
// Bits before the start phase in the first byte are preserved; the last byte is zero padded.
//
// Algorithmic bytes: N read + C written (traffic 2N + C).  Roofline: HBM.
#include "common.cuh"

namespace hf {

constexpr int ENC_TEAM = 768;                       // threads per team
constexpr int ENC_NT = 1;                           // teams per CTA: consecutive tiles of a segment, in flight together
constexpr int ENC_THREADS = ENC_TEAM * ENC_NT;
constexpr int ENC_WARPS = ENC_TEAM / 32;            // warps per team
constexpr int ENC_GROUPS = 2;                       // 8-symbol groups per thread
constexpr uint32_t ENC_TILE_SYMS = ENC_TEAM * 8 * ENC_GROUPS;       // 4096
constexpr uint32_t ENC_WIN = 8192;                  // staging words per team (multiple of 4): ~22 bits per symbol
constexpr uint32_t ENC_PLANE_BYTES = NSYM * 3;      // p16 + p8
constexpr size_t ENC_SMEM = ENC_PLANE_BYTES + (size_t)ENC_NT * ENC_WIN * 4;
constexpr uint32_t ENC_SPIN_LIMIT = 1u << 28;
constexpr uint32_t ENC_MAX_SEGS = 2048;
constexpr int CNT_THREADS = 512;
constexpr uint32_t V_FALLBACK = 0x80000000u;        // v = V_FALLBACK | symbol: take (len, code) from global memory
constexpr unsigned long long NOT_FINAL = ~0ull;

struct EncWork {                                    // lives in ctx->ws, zeroed per launch
    unsigned long long counter;                     // next segment (encode_kernel)
    unsigned long long done;                        // CTAs of enc_count_kernel that have finished
    unsigned long long seg_bits[ENC_MAX_SEGS];
    unsigned long long seg_start[ENC_MAX_SEGS];     // exclusive scan of seg_bits
    unsigned long long phase_cycles[8];             // HF_ENC_TIMING builds
};

#ifdef HF_ENC_TIMING
#define ENC_TICK(k) do { if (tid == 0) { long long _n = clock64(); atomicAdd(&work->phase_cycles[k], (unsigned long long)(_n - t_last)); t_last = _n; } } while (0)
#else
#define ENC_TICK(k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t fold16(uint32_t sym) { return sym ^ (sym >> 8); }      // involution on 16 bits

__device__ __forceinline__ void team_sync(uint32_t team)
{
    asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(ENC_TEAM) : "memory");
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CNT_THREADS)
enc_count_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb, EncWork *work,
                 uint32_t nseg, uint64_t seg_syms)
{
    extern __shared__ __align__(16) uint8_t cnt_smem[];        // lenf plane, 64 KiB
    __shared__ unsigned long long s_part[CNT_THREADS / 32];
    __shared__ bool s_last;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(cb->lenf);
        uint4 *dst = reinterpret_cast<uint4 *>(cnt_smem);
        for (uint32_t i = tid; i < NSYM / 16; i += CNT_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint16_t *in16 = reinterpret_cast<const uint16_t *>(in_bytes);
    const bool aligned = ((uintptr_t)in_bytes & 15) == 0;
    for (uint32_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const uint64_t s0 = (uint64_t)seg * seg_syms;
        const uint64_t s1 = min(n_sym, s0 + seg_syms);
        unsigned long long sum = 0;
        uint64_t done = s0;
        if (aligned) {                                          // seg_syms is a multiple of 8: s0 is 16-byte aligned
            const uint64_t nvec = (s1 - s0) / 8;
            const uint4 *src = reinterpret_cast<const uint4 *>(in16 + s0);
            uint32_t acc = 0;
            for (uint64_t i = tid; i < nvec; i += CNT_THREADS) {
                const uint4 x = ld_stream_v4(src + i);
                acc += cnt_smem[fold16(x.x & 0xFFFFu)] + cnt_smem[fold16(x.x >> 16)] +
                       cnt_smem[fold16(x.y & 0xFFFFu)] + cnt_smem[fold16(x.y >> 16)] +
                       cnt_smem[fold16(x.z & 0xFFFFu)] + cnt_smem[fold16(x.z >> 16)] +
                       cnt_smem[fold16(x.w & 0xFFFFu)] + cnt_smem[fold16(x.w >> 16)];
                if (acc > 0xF0000000u) { sum += acc; acc = 0; }
            }
            sum += acc;
            done = s0 + nvec * 8;
        }
        for (uint64_t s = done + tid; s < s1; s += CNT_THREADS) sum += cnt_smem[fold16(in16[s])];
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if (lane == 0) s_part[wid] = sum;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
            for (int i = 0; i < CNT_THREADS / 32; i++) t += s_part[i];
            work->seg_bits[seg] = t;
        }
        __syncthreads();
    }
    // the CTA that finishes last scans the segment totals
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&work->done, 1ull) + 1 == gridDim.x;
    }
    __syncthreads();
    if (s_last && wid == 0) {
        __threadfence();
        unsigned long long run = 0;
        for (uint32_t b = 0; b < nseg; b += 32) {
            const uint32_t i = b + lane;
            const unsigned long long v = i < nseg ? *((volatile unsigned long long *)&work->seg_bits[i]) : 0ull;
            unsigned long long x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += y;
            }
            if (i < nseg) work->seg_start[i] = run + x - v;
            run += __shfl_sync(0xFFFFFFFFu, x, 31);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// predicated shared-memory OR without a branch (the compiler turns `if (p) atomicOr(...)` into a
// divergent branch with a generic-address conversion on each side)
__device__ __forceinline__ void red_or_shared_if(bool p, uint32_t saddr, uint32_t v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.shared.or.b32 [%1], %2;\n\t}"
                 :: "r"((uint32_t)p), "r"(saddr), "r"(v) : "memory");
}

// append `len` (<= 32) bits to the right-aligned accumulator; a full word is OR-ed into the (zeroed)
// staging window.  nb = bits held (< 32 on entry), w = tile-relative word the next flush writes,
// wbase = first word of the window, sbase = shared-window byte address of the window.
// Branch-free: the flush is predicated, so lanes with and without a full word do not diverge.
__device__ __forceinline__ void put_bits(uint32_t sbase, uint32_t wbase, unsigned long long &acc, uint32_t &nb,
                                         uint32_t &w, uint32_t code, uint32_t len)
{
    acc = (acc << len) | code;                      // len <= 32: the 64-bit shift is well defined
    nb += len;
    const bool fl = nb >= 32;
    const uint32_t nb2 = fl ? nb - 32 : nb;
    const uint32_t word = (uint32_t)(acc >> nb2);   // only used when fl
    const uint32_t i = w - wbase;
    red_or_shared_if(fl && i < ENC_WIN, sbase + 4u * i, word);
    w += fl ? 1u : 0u;
    nb = nb2;
}

// (len, code) of one symbol from the shared planes, or from the global codebook when it is not there
__device__ __forceinline__ void lookup_slow(const uint16_t *p16, const uint8_t *p8, const Codebook *cb, uint32_t sym,
                                            uint32_t &len, unsigned long long &code)
{
    const uint32_t f = fold16(sym);
    const uint32_t x = (uint32_t)p16[f] | ((uint32_t)p8[f] << 16);
    if (x) { len = 31 - __clz(x); code = x ^ (1u << len); }
    else { len = cb->len[sym]; code = cb->code[sym]; }
}

// one group of up to 8 symbols packed the slow, general way: any code length, any window position
__device__ __noinline__ void pack_group_slow(const uint16_t *in16, uint64_t s0, uint32_t count, const uint16_t *p16,
                                             const uint8_t *p8, const Codebook *cb, uint32_t sbase, uint32_t wbase,
                                             uint32_t pos)
{
    uint32_t w = pos >> 5, nb = pos & 31;
    unsigned long long acc = 0;
#pragma unroll 1
    for (uint32_t j = 0; j < count; j++) {
        uint32_t len;
        unsigned long long code;
        lookup_slow(p16, p8, cb, in16[s0 + j], len, code);
        if (len > 32) { put_bits(sbase, wbase, acc, nb, w, (uint32_t)(code >> 32), len - 32); len = 32; }
        put_bits(sbase, wbase, acc, nb, w, (uint32_t)code, len);
    }
    // tail word, shared with the next thread
    red_or_shared_if(nb != 0 && w - wbase < ENC_WIN, sbase + 4u * (w - wbase), (uint32_t)(acc << (32 - nb)));
}

__global__ void __launch_bounds__(ENC_THREADS, 1)
encode_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb,
              uint8_t *stream, uint64_t start_bit, EncWork *work, uint32_t ntiles, uint32_t nseg, uint32_t seg_tiles)
{
    extern __shared__ __align__(16) uint8_t enc_smem[];
    const uint16_t *p16 = reinterpret_cast<const uint16_t *>(enc_smem);
    const uint8_t *p8 = enc_smem + NSYM * 2;
    __shared__ unsigned long long s_scan[ENC_NT][ENC_WARPS + 1];
    __shared__ unsigned long long s_final[ENC_NT];
    __shared__ unsigned long long s_prefix[ENC_NT];
    __shared__ unsigned long long s_chain_prefix;               // bits before tile s_chain_tile (the teams' hand-over)
    __shared__ uint32_t s_chain_tile;
    __shared__ uint32_t s_seg;

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t team = tid / ENC_TEAM, ttid = tid % ENC_TEAM, wid = ttid >> 5;
    uint32_t *stage = reinterpret_cast<uint32_t *>(enc_smem + ENC_PLANE_BYTES) + team * ENC_WIN;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
    const uint16_t *in16 = reinterpret_cast<const uint16_t *>(in_bytes);

    // the table planes: one coalesced copy per CTA, L2 resident after the first
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(cb->p16);
        uint4 *dst = reinterpret_cast<uint4 *>(enc_smem);
        for (uint32_t i = tid; i < ENC_PLANE_BYTES / 16; i += ENC_THREADS) dst[i] = __ldg(src + i);
    }

    // aligned frame: bit 0 of the frame is the 16-byte boundary at or below `stream`
    uint8_t *frame = reinterpret_cast<uint8_t *>((uintptr_t)stream & ~(uintptr_t)15);
    const unsigned long long bit0 = ((uintptr_t)stream & 15) * 8ull + start_bit;   // first payload bit, frame coordinates
    const bool in_aligned = ((uintptr_t)in_bytes & 15) == 0;

    // raw input of a tile: full tiles of an aligned input come as 2 x 128 bit per thread
    uint4 raw[ENC_GROUPS];
    auto is_full = [&](uint32_t t) { return in_aligned && t < ntiles && (uint64_t)(t + 1) * ENC_TILE_SYMS <= n_sym; };
    auto load_raw = [&](uint32_t t) {
        if (is_full(t)) {
#pragma unroll
            for (int g = 0; g < ENC_GROUPS; g++)
                raw[g] = ld_stream_v4(in_bytes + ((uint64_t)t * ENC_TILE_SYMS + (g * ENC_TEAM + ttid) * 8) * 2);
        }
    };
    // the next tile is pulled into L2 while this one is packed (no registers held across the tile)
    auto prefetch_tile = [&](uint32_t t) {
        if (is_full(t) && ttid < ENC_TILE_SYMS * 2 / 128)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(in_bytes + (uint64_t)t * ENC_TILE_SYMS * 2 + ttid * 128));
    };

#ifdef HF_ENC_TIMING
    long long t_last = clock64();
#endif
    for (;;) {
        __syncthreads();                                        // planes loaded / s_seg, staging reuse
        ENC_TICK(7);
        if (tid == 0) {
            const uint32_t sg = (uint32_t)atomicAdd(&work->counter, 1ull);
            s_seg = sg;
            if (sg < nseg) { s_chain_tile = sg * seg_tiles; s_chain_prefix = work->seg_start[sg]; }   // bits before this segment
        }
        __syncthreads();
        const uint32_t seg = s_seg;
        if (seg >= nseg) break;
        const uint32_t t_lo = seg * seg_tiles, t_hi = min(ntiles, t_lo + seg_tiles);

        // the teams take the tiles of the segment in turn; the running bit offset passes from tile to tile
        // through shared memory as soon as a tile's bit count is known (before it is packed)
        for (uint32_t tile = t_lo + team; tile < t_hi; tile += ENC_NT) {
            const uint64_t sym0 = (uint64_t)tile * ENC_TILE_SYMS;
            const uint32_t nsym = (uint32_t)min((uint64_t)ENC_TILE_SYMS, n_sym - sym0);
            const bool full = is_full(tile);
            load_raw(tile);
            if (tile + ENC_NT < t_hi) prefetch_tile(tile + ENC_NT);

            // ---- 1. symbols -> table entries, length sums ----
            uint32_t v[ENC_GROUPS][8];              // (1 << len) | code, or V_FALLBACK | symbol
            uint32_t glen[ENC_GROUPS];
#pragma unroll
            for (int g = 0; g < ENC_GROUPS; g++) {
                uint32_t sym[8];
                if (full) {
                    const uint4 x = raw[g];
                    sym[0] = x.x & 0xFFFFu; sym[1] = x.x >> 16; sym[2] = x.y & 0xFFFFu; sym[3] = x.y >> 16;
                    sym[4] = x.z & 0xFFFFu; sym[5] = x.z >> 16; sym[6] = x.w & 0xFFFFu; sym[7] = x.w >> 16;
                } else {
                    const uint32_t s_base = (g * ENC_TEAM + ttid) * 8;
#pragma unroll
                    for (int j = 0; j < 8; j++) sym[j] = (s_base + j < nsym) ? (uint32_t)in16[sym0 + s_base + j] : 0x10000u;
                }
                uint32_t L = 0, zero = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t f = fold16(sym[j] & 0xFFFFu);
                    uint32_t x = (uint32_t)p16[f] | ((uint32_t)p8[f] << 16);
                    if (!full && sym[j] > 0xFFFFu) x = 1u;         // no symbol: zero bits
                    v[g][j] = x;
                    zero = min(zero, x);
                    L += 31 - __clz(x | 1u);
                }
                if (zero == 0) {                                    // some code is longer than 23 bits: rare
                    L = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (v[g][j] == 0) { v[g][j] = V_FALLBACK | sym[j]; L += cb->len[sym[j]]; }
                        else L += 31 - __clz(v[g][j]);
                    }
                }
                glen[g] = L;
            }

            ENC_TICK(0);                                        // load + lookup (thread 0's view)
            // ---- CTA exclusive scan of (glen[0], glen[1]) packed in one 64-bit value ----
            const unsigned long long pk = (unsigned long long)glen[0] | ((unsigned long long)glen[1] << 32);
            unsigned long long x = pk;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_scan[team][wid] = x;
            team_sync(team);
            if (wid == 0) {
                unsigned long long s = lane < ENC_WARPS ? s_scan[team][lane] : 0ull;
                unsigned long long t = s;
#pragma unroll
                for (int o = 1; o < ENC_WARPS; o <<= 1) {
                    unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o);
                    if (lane >= o) t += y;
                }
                if (lane < ENC_WARPS) s_scan[team][lane] = t - s;
                if (lane == ENC_WARPS - 1) {
                    s_scan[team][ENC_WARPS] = t;                // totals of both groups
                    // hand-over: take the bits before my tile, pass on the bits before the next one
                    const uint32_t bits = (uint32_t)t + (uint32_t)(t >> 32);
                    uint32_t spins = 0;
                    while (*((volatile uint32_t *)&s_chain_tile) != tile)
                        if (++spins > ENC_SPIN_LIMIT) break;
                    const unsigned long long pfx = *((volatile unsigned long long *)&s_chain_prefix);
                    s_prefix[team] = pfx;
                    *((volatile unsigned long long *)&s_chain_prefix) = pfx + bits;
                    __threadfence_block();
                    *((volatile uint32_t *)&s_chain_tile) = tile + 1;
                }
            }
            // the window is zeroed while warp 0 scans
            for (uint32_t i = ttid; i < ENC_WIN / 4; i += ENC_TEAM)
                reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
            team_sync(team);
            ENC_TICK(1);                                        // scan + zeroing + hand-over
            const unsigned long long excl = x - pk + s_scan[team][wid];
            const unsigned long long tot = s_scan[team][ENC_WARPS];
            const uint32_t totA = (uint32_t)tot, totB = (uint32_t)(tot >> 32);
            const uint32_t tile_bits = totA + totB;
            const uint32_t off[ENC_GROUPS] = {(uint32_t)excl, totA + (uint32_t)(excl >> 32)};

            const unsigned long long gbit = bit0 + s_prefix[team];  // tile's first bit, frame coordinates
            const unsigned long long gend = gbit + tile_bits;
            const uint32_t phase = (uint32_t)(gbit & 127);
            const unsigned long long G0 = (gbit - phase) >> 5;      // frame word of tile-relative word 0 (multiple of 4)
            const bool last_tile = tile + 1 == ntiles;
            // words I own: those whose first bit is mine (tile 0 also owns the word the stream starts in)
            const unsigned long long own_lo = tile == 0 ? (gbit >> 5) : ((gbit + 31) >> 5);
            const unsigned long long own_hi = tile_bits ? ((gend - 1) >> 5) : 0;       // valid when tile_bits > 0
            const bool owns = tile_bits > 0 && own_hi >= own_lo;
            const uint32_t hi_rel = (uint32_t)(own_hi - G0);
            const uint32_t npass = owns ? hi_rel / ENC_WIN + 1 : 0;

            for (uint32_t pass = 0; pass < npass; pass++) {
                const uint32_t wbase = pass * ENC_WIN;
                if (pass) {
                    team_sync(team);                                // the previous window has left
                    for (uint32_t i = ttid; i < ENC_WIN / 4; i += ENC_TEAM)
                        reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
                    team_sync(team);
                }

                // ---- 2. pack this thread's codes into the window ----
#pragma unroll
                for (int g = 0; g < ENC_GROUPS; g++) {
                    uint32_t pos = phase + off[g];
                    const uint32_t w_lo = pos >> 5, w_hi = (pos + glen[g]) >> 5;
                    if (w_hi < wbase || w_lo >= wbase + ENC_WIN) continue;      // nothing of mine in this window
                    const bool inside = w_lo >= wbase && w_hi + 1 < wbase + ENC_WIN;   // no word of mine leaves it
                    const uint32_t fb = v[g][0] | v[g][1] | v[g][2] | v[g][3] | v[g][4] | v[g][5] | v[g][6] | v[g][7];
                    if (inside && !(fb & V_FALLBACK)) {
                        // pairs of codes (<= 46 bits) are merged in registers and OR-ed in with up to three
                        // predicated shared-memory reductions: no carried accumulator, no divergence
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            const uint32_t x0 = v[g][j], x1 = v[g][j + 1];
                            const uint32_t l0 = 31 - __clz(x0), l1 = 31 - __clz(x1);
                            const uint32_t a32 = __funnelshift_lc(0u, x0, 32 - l0);   // left aligned, the leading one falls off
                            const uint32_t b32 = __funnelshift_lc(0u, x1, 32 - l1);
                            const uint32_t hi = a32 | (b32 >> l0);                    // l0 <= 23
                            const uint32_t lo = __funnelshift_r(0u, b32, l0);
                            const uint32_t L = l0 + l1, sh = pos & 31;
                            const uint32_t sa = sbase + 4u * ((pos >> 5) - wbase);
                            red_or_shared_if(L != 0, sa, hi >> sh);
                            red_or_shared_if(sh + L > 32, sa + 4, __funnelshift_r(lo, hi, sh));
                            red_or_shared_if(sh + L > 64, sa + 8, __funnelshift_r(0u, lo, sh));
                            pos += L;
                        }
                    } else {
                        // window edge or a code longer than 23 bits (rare): symbol by symbol from the input again
                        const uint32_t s_base = (g * ENC_TEAM + ttid) * 8;
                        pack_group_slow(in16, sym0 + s_base, s_base < nsym ? min(8u, nsym - s_base) : 0u, p16, p8, cb, sbase,
                                        wbase, pos);
                    }
                }
                team_sync(team);
                ENC_TICK(2);                                    // pack

                // ---- 3. the stream head (tile 0) and my last, partial word ----
                const bool last_pass = pass + 1 == npass;
                if (ttid == 0) {
                    if (pass == 0 && tile == 0) {
                        // preserve the bits of the first byte that precede the start phase
                        const uint32_t b = frame[gbit >> 3];
                        const uint32_t keep = b & ~(0xFFu >> (gbit & 7));
                        stage[(gbit >> 5) - G0] |= keep << (24 - 8 * (uint32_t)((gbit >> 3) & 3));
                    }
                    if (last_pass) {
                        unsigned long long fin = last_tile ? gend : NOT_FINAL;
                        uint32_t have = (uint32_t)(gend & 31);      // bits of my last word that are mine
                        if (have && !last_tile) {
                            // complete the word with the codes of the symbols that follow the tile
                            uint32_t word = 0;
                            uint64_t s = sym0 + nsym;
                            unsigned long long end = gend;
                            while (have < 32 && s < n_sym) {
                                uint32_t len;
                                unsigned long long code;
                                lookup_slow(p16, p8, cb, in16[s], len, code);
                                if (len) {
                                    const unsigned long long left = code << (64 - len);    // left aligned
                                    word |= (uint32_t)(left >> 32) >> have;
                                    have += len;
                                    end += len;
                                }
                                s++;
                            }
                            stage[hi_rel - wbase] |= word;
                            if (have < 32) fin = end;               // the input ended inside my word: it is the last one
                        }
                        s_final[team] = fin;
                    }
                }
                team_sync(team);
                ENC_TICK(3);                                    // head / tail completion
                const unsigned long long fin = last_pass ? s_final[team] : NOT_FINAL;

                // my words of this window leave: 128-bit stores where a whole group is mine, else words / bytes
                const unsigned long long W0 = G0 + wbase;
                uint32_t n_groups = ENC_WIN / 4;
                if (own_hi < W0 + ENC_WIN) n_groups = (uint32_t)((own_hi - W0 + 4) >> 2);
                uint32_t *gw = reinterpret_cast<uint32_t *>(frame);
                for (uint32_t q = ttid; q < n_groups; q += ENC_TEAM) {
                    const unsigned long long w0 = W0 + 4ull * q;
                    uint4 o = reinterpret_cast<const uint4 *>(stage)[q];
                    o.x = bswap32(o.x); o.y = bswap32(o.y); o.z = bswap32(o.z); o.w = bswap32(o.w);
                    const bool whole = w0 >= own_lo && w0 + 3 <= own_hi && !(tile == 0 && w0 <= (bit0 >> 5)) &&
                                       !(fin != NOT_FINAL && w0 + 3 == own_hi);
                    if (whole) {
                        st_stream_v4(gw + w0, o);
                    } else {
                        const uint32_t vv[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const unsigned long long wk = w0 + k;
                            if (wk < own_lo || wk > own_hi) continue;
                            uint32_t b_lo = 0, b_hi = 4;            // byte range [b_lo, b_hi) of this word to store
                            if (wk == (bit0 >> 5)) b_lo = (uint32_t)((bit0 >> 3) & 3);   // bytes before the stream are not ours
                            if (fin != NOT_FINAL && wk == own_hi) b_hi = (uint32_t)(((fin - 1) >> 3) & 3) + 1;   // bytes holding bits
                            if (b_lo == 0 && b_hi == 4) gw[wk] = vv[k];
                            else
                                for (uint32_t b = b_lo; b < b_hi; b++)
                                    frame[wk * 4 + b] = (uint8_t)(vv[k] >> (8 * b));   // little-endian view of the swapped word
                        }
                    }
                }
            }
            team_sync(team);                                        // staging, s_scan, s_final reuse
            ENC_TICK(4);                                        // copy-out
        }
    }
}

int launch_encode_old(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb, uint8_t *d_stream,
                      uint64_t start_bit, uint32_t maxlen_hint)
{
    (void)maxlen_hint;
    const uint64_t n_sym = n_bytes / 2;
    if (n_sym == 0) return HF_OK;
    if ((uintptr_t)d_in & 1) return set_err(c, HF_ERR_ARG, "hf_encode: input must be 2-byte aligned");
    if ((uintptr_t)d_cb & 15) return set_err(c, HF_ERR_ARG, "hf_encode: codebook must be 16-byte aligned");
    const uint64_t nt64 = (n_sym + ENC_TILE_SYMS - 1) / ENC_TILE_SYMS;
    if (nt64 > 0x7FFFFFFFull) return set_err(c, HF_ERR_ARG, "hf_encode: input too large");
    const uint32_t ntiles = (uint32_t)nt64;
    // ~8 segments per SM (dynamic claims balance segments of different entropy), whole tiles each
    uint32_t seg_tiles = (ntiles + 8 * c->sm_count - 1) / (8 * c->sm_count);
    if (seg_tiles < 1) seg_tiles = 1;
    uint32_t nseg = (ntiles + seg_tiles - 1) / seg_tiles;
    while (nseg > ENC_MAX_SEGS) { seg_tiles++; nseg = (ntiles + seg_tiles - 1) / seg_tiles; }
    // the encode workspace sits behind the codebook workspace so the two never alias a live buffer
    const size_t off = 8u << 20;
    int rc = ensure_ws(c, off + sizeof(EncWork));
    if (rc) return rc;
    EncWork *work = reinterpret_cast<EncWork *>((uint8_t *)c->ws + off);
    HF_CUDA(c, cudaMemsetAsync(work, 0, sizeof(EncWork), c->stream));

    static bool attr_set = false;
    if (!attr_set) {
        HF_CUDA(c, cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM));
        HF_CUDA(c, cudaFuncSetAttribute(enc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NSYM));
        attr_set = true;
    }
    // start_bit may exceed 8: fold whole bytes into the pointer
    d_stream += start_bit >> 3;
    start_bit &= 7;
    const uint32_t cgrid = nseg < (uint32_t)(3 * c->sm_count) ? nseg : (uint32_t)(3 * c->sm_count);
    HF_PROF(c, "enc_count_kernel"); enc_count_kernel<<<cgrid, CNT_THREADS, NSYM, c->stream>>>(d_in, n_sym, d_cb, work, nseg, (uint64_t)seg_tiles * ENC_TILE_SYMS);
    HF_LAUNCH_CHECK(c);
    const uint32_t grid = nseg < (uint32_t)c->sm_count ? nseg : (uint32_t)c->sm_count;
    HF_PROF(c, "encode_kernel"); encode_kernel<<<grid, ENC_THREADS, ENC_SMEM, c->stream>>>(d_in, n_sym, d_cb, d_stream, start_bit, work, ntiles, nseg, seg_tiles);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
