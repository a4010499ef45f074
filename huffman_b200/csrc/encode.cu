// encode.cu — single-pass Huffman encoder: code lookup, decoupled-lookback exclusive scan of
// the bit offsets and bit packing in ONE kernel.
//
// Replaces populateCWLength + thrust::transform_inclusive_scan + encodeFromCW + the host
// tail flush (/root/reference/Compressor.cu:50-74, :152-313, :541-601, :673-684): the
// reference materialises 12 bytes of scratch per symbol and then binary-searches the
// offsets once per OUTPUT byte.  Here the input is read once and the output written once.
//
// The code table of the 65,536-symbol alphabet lives in SHARED memory: 24-bit entries
// (1 << len) | code  for len <= 23 — the leading one bit carries the length — stored as two
// planes (u16 low part, u8 high part: 192 KiB of the SM's 227 KiB) and indexed by the symbol
// with its low byte XOR-folded with its high byte, so that skewed first bytes (text, Zipf) still
// spread over the 32 banks.  Gathers from shared memory cost bank conflicts only; the same gathers
// from global memory cost one L1 wavefront per distinct line and bounded the first version of this
// kernel.  Entry 0 means "longer than 23 bits" (or absent): (len, code) come from global memory.
//
// One persistent 1024-thread CTA per SM = 4 independent teams of 256 threads (named barriers),
// each looping over tiles of 4096 symbols claimed from an atomic counter; the next tile's index
// and input are fetched while the current tile is packed:
//   1. 2 x 128-bit loads per thread (8 symbols each), table lookup, length sums,
//      team-wide exclusive scan of both groups at once (packed 2 x 32 bit);
//   2. warp 0 of the team publishes the tile's bit count and resolves its global bit offset by
//      decoupled look-back over the predecessors' descriptors (acquire/release);
//   3. every symbol's code is OR-ed (shared-memory atomics, no carried state, no divergent flush)
//      into the team's staging window (2048 words at the tile's 128-bit phase); a tile whose bits
//      exceed the window (more than ~16 bits per symbol) is packed window by window;
//   4. a 32-bit word belongs to the tile that holds its first bit: the owner completes its last,
//      partial word by encoding the symbols that FOLLOW the tile until the word is full, so tiles
//      exchange nothing but the look-back prefix; windows leave with aligned 128-bit stores.
// Bits before the start phase in the first byte are preserved; the last byte is zero padded.
//
// Algorithmic bytes: N read + C written.  Roofline: HBM.
#include "common.cuh"

namespace hf {

constexpr int ENC_TEAM = 256;                       // threads per team
constexpr int ENC_NT = 3;                           // teams per CTA
constexpr int ENC_THREADS = ENC_TEAM * ENC_NT;
constexpr int ENC_GROUPS = 2;                       // 8-symbol groups per thread
constexpr uint32_t ENC_TILE_SYMS = ENC_TEAM * 8 * ENC_GROUPS;       // 4096
constexpr uint32_t ENC_WIN = 2048;                  // staging words per team (multiple of 4)
constexpr uint32_t ENC_PLANE_BYTES = NSYM * 3;      // p16 + p8
constexpr size_t ENC_SMEM = ENC_PLANE_BYTES + (size_t)ENC_NT * ENC_WIN * 4;
constexpr uint32_t ST_INVALID = 0, ST_AGG = 1, ST_INCL = 2;
constexpr uint32_t SPIN_LIMIT = 1u << 26;
constexpr uint32_t V_FALLBACK = 0x80000000u;        // v = V_FALLBACK | symbol: take (len, code) from global memory
constexpr unsigned long long NOT_FINAL = ~0ull;

struct EncWork {                                    // lives in ctx->ws, zeroed per launch
    unsigned long long counter;                     // next tile
    unsigned long long error;
    // followed by desc[ntiles]
};

__device__ __forceinline__ void team_sync(uint32_t team)
{
    asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(ENC_TEAM) : "memory");
}

__device__ __forceinline__ uint32_t fold16(uint32_t sym) { return sym ^ (sym >> 8); }      // involution on 16 bits

// OR `len` (<= 32) bits, given left aligned in c32, into the window at tile-relative bit `pos`
__device__ __forceinline__ void or_code(uint32_t *stage, uint32_t wbase, uint32_t pos, uint32_t c32, uint32_t len)
{
    const uint32_t sh = pos & 31, w = (pos >> 5) - wbase;
    if (w < ENC_WIN) atomicOr(&stage[w], c32 >> sh);
    if (sh + len > 32 && w + 1 < ENC_WIN) atomicOr(&stage[w + 1], __funnelshift_r(0u, c32, sh));
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

__global__ void __launch_bounds__(ENC_THREADS, 1)
encode_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb,
              uint8_t *stream, uint64_t start_bit, EncWork *work, uint32_t ntiles)
{
    extern __shared__ __align__(16) uint8_t enc_smem[];
    const uint16_t *p16 = reinterpret_cast<const uint16_t *>(enc_smem);
    const uint8_t *p8 = enc_smem + NSYM * 2;
    __shared__ unsigned long long s_scan[ENC_NT][12];
    __shared__ unsigned long long s_bcast[ENC_NT];
    __shared__ unsigned long long s_final[ENC_NT];
    __shared__ uint32_t s_tile[ENC_NT];

    const uint32_t tid = threadIdx.x, team = tid / ENC_TEAM, ttid = tid % ENC_TEAM;
    const uint32_t lane = tid & 31, twid = ttid >> 5;
    uint32_t *stage = reinterpret_cast<uint32_t *>(enc_smem + ENC_PLANE_BYTES) + team * ENC_WIN;
    const uint16_t *in16 = reinterpret_cast<const uint16_t *>(in_bytes);

    // the table planes: one coalesced copy per CTA, L2 resident after the first
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(cb->p16);
        uint4 *dst = reinterpret_cast<uint4 *>(enc_smem);
        for (uint32_t i = tid; i < ENC_PLANE_BYTES / 16; i += ENC_THREADS) dst[i] = __ldg(src + i);
    }
    if (ttid == 0) s_tile[team] = (uint32_t)atomicAdd(&work->counter, 1ull);
    __syncthreads();

    unsigned long long *desc = reinterpret_cast<unsigned long long *>(work + 1);
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
    uint32_t tile = s_tile[team];
    load_raw(tile);

    while (tile < ntiles) {
        const uint64_t sym0 = (uint64_t)tile * ENC_TILE_SYMS;
        const uint32_t nsym = (uint32_t)min((uint64_t)ENC_TILE_SYMS, n_sym - sym0);
        const bool full = is_full(tile);

        // ---- 1. symbols -> table entries, length sums ----
        uint32_t v[ENC_GROUPS][8];                  // (1 << len) | code, or V_FALLBACK | symbol
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

        // ---- team exclusive scan of (glen[0], glen[1]) packed in one 64-bit value;
        //      the claim of the next tile rides on the same barriers ----
        if (ttid == 0) s_tile[team] = (uint32_t)atomicAdd(&work->counter, 1ull);
        const unsigned long long pk = (unsigned long long)glen[0] | ((unsigned long long)glen[1] << 32);
        unsigned long long x = pk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[team][twid] = x;
        team_sync(team);
        if (twid == 0) {
            unsigned long long s = lane < ENC_TEAM / 32 ? s_scan[team][lane] : 0ull;
            unsigned long long t = s;
#pragma unroll
            for (int o = 1; o < ENC_TEAM / 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o);
                if (lane >= o) t += y;
            }
            if (lane < ENC_TEAM / 32) s_scan[team][lane] = t - s;
            if (lane == ENC_TEAM / 32 - 1) s_scan[team][8] = t;   // totals of both groups
        }
        team_sync(team);
        const uint32_t next_tile = s_tile[team];
        load_raw(next_tile);                                    // in flight while this tile is packed
        const unsigned long long excl = x - pk + s_scan[team][twid];
        const unsigned long long tot = s_scan[team][8];
        const uint32_t totA = (uint32_t)tot, totB = (uint32_t)(tot >> 32);
        const uint32_t tile_bits = totA + totB;
        const uint32_t off[ENC_GROUPS] = {(uint32_t)excl, totA + (uint32_t)(excl >> 32)};

        // ---- 2. decoupled look-back for the tile's exclusive bit prefix ----
        if (twid == 0) {
            unsigned long long prefix = 0;
            if (tile == 0) {
                if (lane == 0) st_release_u64(&desc[0], ((unsigned long long)ST_INCL << 62) | tile_bits);
            } else {
                if (lane == 0) st_release_u64(&desc[tile], ((unsigned long long)ST_AGG << 62) | tile_bits);
                int look = (int)tile - 1;
                uint32_t spins = 0;
                for (;;) {
                    int idx = look - (int)lane;
                    unsigned long long d = idx >= 0 ? ld_acquire_u64(&desc[idx]) : ((unsigned long long)ST_INCL << 62);
                    uint32_t st = (uint32_t)(d >> 62);
                    uint32_t m_incl = __ballot_sync(0xFFFFFFFFu, st == ST_INCL);
                    uint32_t m_inv = __ballot_sync(0xFFFFFFFFu, st == ST_INVALID);
                    uint32_t first = m_incl ? (uint32_t)__ffs(m_incl) - 1 : 32u;
                    uint32_t need = first < 32 ? ((2u << first) - 1u) : 0xFFFFFFFFu;
                    if (m_inv & need) {             // a needed predecessor has not published yet
                        if (++spins > SPIN_LIMIT) { if (lane == 0) atomicExch(&work->error, 1ull); break; }
                        continue;
                    }
                    unsigned long long val = (lane <= first) ? (d & 0x3FFFFFFFFFFFFFFFull) : 0ull;
#pragma unroll
                    for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, o);
                    prefix += val;
                    if (first < 32) break;
                    look -= 32;
                }
                if (lane == 0) st_release_u64(&desc[tile], ((unsigned long long)ST_INCL << 62) | (prefix + tile_bits));
            }
            if (lane == 0) s_bcast[team] = prefix;
        }
        // the first window can be zeroed while warp 0 looks back
        for (uint32_t i = ttid; i < ENC_WIN / 4; i += ENC_TEAM)
            reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
        team_sync(team);
        const unsigned long long gbit = bit0 + s_bcast[team];       // tile's first bit, frame coordinates
        const unsigned long long gend = gbit + tile_bits;
        const uint32_t phase = (uint32_t)(gbit & 127);
        const unsigned long long G0 = (gbit - phase) >> 5;          // frame word of tile-relative word 0 (multiple of 4)
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
                team_sync(team);                                    // the previous window has left
                for (uint32_t i = ttid; i < ENC_WIN / 4; i += ENC_TEAM)
                    reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
                team_sync(team);
            }

            // ---- 3. OR this thread's codes into the window ----
#pragma unroll
            for (int g = 0; g < ENC_GROUPS; g++) {
                uint32_t pos = phase + off[g];
                if (((pos + glen[g]) >> 5) < wbase || (pos >> 5) >= wbase + ENC_WIN) continue;   // nothing of mine in this window
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t e = v[g][j];
                    if (e & V_FALLBACK) {
                        const uint32_t s = e & 0xFFFFu;
                        const uint32_t len = cb->len[s];
                        const unsigned long long code = cb->code[s];
                        if (len > 32) {
                            or_code(stage, wbase, pos, (uint32_t)((code << (64 - len)) >> 32), 32);
                            or_code(stage, wbase, pos + 32, (uint32_t)(code << (64 - len)), len - 32);
                        } else if (len) {
                            or_code(stage, wbase, pos, (uint32_t)code << (32 - len), len);
                        }
                        pos += len;
                    } else {
                        const uint32_t len = 31 - __clz(e);
                        or_code(stage, wbase, pos, __funnelshift_lc(0u, e, 32 - len), len);   // the leading one falls off
                        pos += len;
                    }
                }
            }
            team_sync(team);

            // ---- 4. the stream head (tile 0) and my last, partial word ----
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
                    uint32_t have = (uint32_t)(gend & 31);          // bits of my last word that are mine
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
                                const unsigned long long left = code << (64 - len);        // left aligned
                                word |= (uint32_t)(left >> 32) >> have;
                                have += len;
                                end += len;
                            }
                            s++;
                        }
                        stage[hi_rel - wbase] |= word;
                        if (have < 32) fin = end;                   // the input ended inside my word: it is the last one
                    }
                    s_final[team] = fin;
                }
            }
            team_sync(team);
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
                        uint32_t b_lo = 0, b_hi = 4;                // byte range [b_lo, b_hi) of this word to store
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
        team_sync(team);                                            // staging, s_tile, s_final reuse
        tile = next_tile;
    }
}

static size_t enc_work_bytes(uint32_t ntiles) { return sizeof(EncWork) + (size_t)ntiles * 8; }

int launch_encode(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb, uint8_t *d_stream,
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
    // the encode workspace sits behind the codebook workspace so the two never alias a live buffer
    const size_t off = 8u << 20;
    int rc = ensure_ws(c, off + enc_work_bytes(ntiles));
    if (rc) return rc;
    EncWork *work = reinterpret_cast<EncWork *>((uint8_t *)c->ws + off);
    HF_CUDA(c, cudaMemsetAsync(work, 0, enc_work_bytes(ntiles), c->stream));

    static bool attr_set = false;
    if (!attr_set) {
        HF_CUDA(c, cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM));
        attr_set = true;
    }
    // start_bit may exceed 8: fold whole bytes into the pointer
    d_stream += start_bit >> 3;
    start_bit &= 7;
    const uint32_t ctas = (ntiles + ENC_NT - 1) / ENC_NT;
    const uint32_t grid = ctas < (uint32_t)c->sm_count ? ctas : (uint32_t)c->sm_count;
    HF_PROF(c, "encode_kernel"); encode_kernel<<<grid, ENC_THREADS, ENC_SMEM, c->stream>>>(d_in, n_sym, d_cb, d_stream, start_bit, work, ntiles);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
