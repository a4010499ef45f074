// encode.cu — single-pass Huffman encoder: code-length lookup, decoupled-lookback
// exclusive scan of bit offsets and bit packing in ONE kernel.
//
// Replaces populateCWLength + thrust::transform_inclusive_scan + encodeFromCW + the host
// tail flush (/root/reference/Compressor.cu:50-74, :152-313, :541-601, :673-684): the
// reference materialises 12 bytes of scratch per symbol and then binary-searches the
// offsets once per OUTPUT byte.  Here the input is read once and the output written once.
//
// Per tile of 8192 symbols (512 threads x 2 groups x 8 symbols, one 128-bit load each):
//   1. gather (len, code) per symbol from the 256 KiB enc32 table (L1/L2 resident),
//      sum the lengths, CTA-wide exclusive scan of both groups at once (packed 2 x 32 bit);
//   2. warp 0 publishes the tile's bit count and resolves its global bit offset by
//      decoupled look-back over the predecessors' descriptors (acquire/release);
//   3. every thread streams its codes through a 64-bit funnel accumulator into a shared
//      staging buffer laid out at the tile's 128-bit phase, so that
//   4. the staging buffer is copied out with aligned 128-bit stores.  The word shared with
//      the previous tile is completed by THIS tile from the predecessor's published tail
//      word, so every output word is written exactly once: no memset, no global atomics.
// Bits before the start phase in the first byte are preserved; the last byte is zero padded.
//
// Algorithmic bytes: N read + C written.  Roofline: HBM.
#include "common.cuh"

namespace hf {

constexpr int ENC_THREADS = 512;
constexpr int ENC_GROUPS = 2;                       // 8-symbol groups per thread
constexpr uint32_t ENC_TILE_SYMS = ENC_THREADS * 8 * ENC_GROUPS;    // 8192
constexpr uint32_t ST_INVALID = 0, ST_AGG = 1, ST_INCL = 2;
constexpr uint32_t SPIN_LIMIT = 1u << 26;

struct EncWork {                                    // lives in ctx->ws, zeroed per launch
    unsigned long long counter;                     // next tile
    unsigned long long error;
    // followed by desc[ntiles], tail[ntiles]
};

template <bool LONG>
struct EncCfg {
    static constexpr uint32_t MAXLEN = LONG ? 64 : ENC32_MAX_LEN;
    static constexpr uint32_t STAGE_WORDS = ENC_TILE_SYMS * MAXLEN / 32 + 8;   // + 128-bit phase + slack
};

// append `len` (<= 32) bits of `code` to the thread's funnel; flush full words to staging
__device__ __forceinline__ void put_bits(uint32_t *stage, unsigned long long &acc, uint32_t &nb, uint32_t &w,
                                         bool &shared_first, uint32_t code, uint32_t len)
{
    acc |= (unsigned long long)code << (64 - nb - len);       // nb < 32, len <= 32; len == 0 adds nothing
    nb += len;
    if (nb >= 32) {
        uint32_t word = (uint32_t)(acc >> 32);
        if (shared_first) { atomicOr(&stage[w], word); shared_first = false; }
        else stage[w] = word;
        w++;
        acc <<= 32;
        nb -= 32;
    }
}

template <bool LONG>
__global__ void __launch_bounds__(ENC_THREADS)
encode_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb,
              uint8_t *stream, uint64_t start_bit, EncWork *work, uint32_t ntiles)
{
    // variant selection on the device keeps hf_encode asynchronous (see launch_encode)
    if ((cb->maxlen > ENC32_MAX_LEN) != LONG) return;

    extern __shared__ __align__(16) uint32_t stage[];
    __shared__ unsigned long long s_scan[20];
    __shared__ unsigned long long s_bcast[2];
    __shared__ uint32_t s_tile;

    unsigned long long *desc = reinterpret_cast<unsigned long long *>(work + 1);
    unsigned long long *tails = desc + ntiles;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // aligned frame: bit 0 of the frame is the 16-byte boundary at or below `stream`
    uint8_t *frame = reinterpret_cast<uint8_t *>((uintptr_t)stream & ~(uintptr_t)15);
    const unsigned long long bit0 = ((uintptr_t)stream & 15) * 8ull + start_bit;   // first payload bit, frame coordinates
    const uint32_t *enc32 = cb->enc32;

    for (;;) {
        __syncthreads();                            // staging / s_tile reuse
        if (tid == 0) s_tile = (uint32_t)atomicAdd(&work->counter, 1ull);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint64_t sym0 = (uint64_t)tile * ENC_TILE_SYMS;
        const uint32_t nsym = (uint32_t)min((uint64_t)ENC_TILE_SYMS, n_sym - sym0);

        // ---- 1. load symbols, gather codes, sum lengths ----
        uint32_t e[ENC_GROUPS][8];                  // fast path: enc32 entries; long path: symbols
        uint32_t glen[ENC_GROUPS];
#pragma unroll
        for (int g = 0; g < ENC_GROUPS; g++) {
            const uint32_t s_base = (g * ENC_THREADS + tid) * 8;      // first symbol of this group in the tile
            uint32_t sym[8];
            if (s_base + 8 <= nsym) {
                uint4 v = ld_stream_v4(in_bytes + (sym0 + s_base) * 2);
                sym[0] = v.x & 0xFFFFu; sym[1] = v.x >> 16; sym[2] = v.y & 0xFFFFu; sym[3] = v.y >> 16;
                sym[4] = v.z & 0xFFFFu; sym[5] = v.z >> 16; sym[6] = v.w & 0xFFFFu; sym[7] = v.w >> 16;
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint64_t s = sym0 + s_base + j;
                    sym[j] = (s_base + j < nsym)
                                 ? (uint32_t)in_bytes[2 * s] | ((uint32_t)in_bytes[2 * s + 1] << 8)
                                 : 0x10000u;        // marker: no symbol
                }
            }
            uint32_t L = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (LONG) {
                    e[g][j] = sym[j];
                    L += sym[j] < NSYM ? cb->len[sym[j]] : 0;
                } else {
                    uint32_t x = sym[j] < NSYM ? __ldg(&enc32[sym[j]]) : 0u;
                    e[g][j] = x;
                    L += x >> 27;
                }
            }
            glen[g] = L;
        }

        // ---- CTA exclusive scan of (glen[0], glen[1]) packed in one 64-bit value ----
        unsigned long long pk = (unsigned long long)glen[0] | ((unsigned long long)glen[1] << 32);
        unsigned long long x = pk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long s = lane < ENC_THREADS / 32 ? s_scan[lane] : 0ull;
            unsigned long long t = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o);
                if (lane >= o) t += y;
            }
            if (lane < ENC_THREADS / 32) s_scan[lane] = t - s;
            if (lane == 31) s_scan[16] = t;         // totals of both groups
        }
        __syncthreads();
        const unsigned long long excl = x - pk + s_scan[wid];
        const unsigned long long tot = s_scan[16];
        const uint32_t totA = (uint32_t)tot, totB = (uint32_t)(tot >> 32);
        const uint32_t tile_bits = totA + totB;
        uint32_t off[ENC_GROUPS] = {(uint32_t)excl, totA + (uint32_t)(excl >> 32)};

        // ---- 2. decoupled look-back for the tile's exclusive bit prefix ----
        if (wid == 0) {
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
                    unsigned long long v = (lane <= first) ? (d & 0x3FFFFFFFFFFFFFFFull) : 0ull;
#pragma unroll
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                    prefix += v;
                    if (first < 32) break;
                    look -= 32;
                }
                if (lane == 0) st_release_u64(&desc[tile], ((unsigned long long)ST_INCL << 62) | (prefix + tile_bits));
            }
            if (lane == 0) s_bcast[0] = prefix;
        }
        // zero the staging buffer while warp 0 looks back
        for (uint32_t i = tid; i < EncCfg<LONG>::STAGE_WORDS / 4; i += ENC_THREADS)
            reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const unsigned long long gbit = bit0 + s_bcast[0];          // tile's first bit, frame coordinates
        const unsigned long long gend = gbit + tile_bits;
        const uint32_t phase = (uint32_t)(gbit & 127);
        const unsigned long long G0 = (gbit - phase) >> 5;          // frame word of staging word 0 (multiple of 4)

        // ---- 3. pack this thread's codes into staging ----
#pragma unroll
        for (int g = 0; g < ENC_GROUPS; g++) {
            uint32_t pos = phase + off[g];
            uint32_t w = pos >> 5;
            uint32_t nb = pos & 31;
            bool shared_first = nb != 0;
            unsigned long long acc = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (LONG) {
                    uint32_t s = e[g][j];
                    if (s < NSYM) {
                        uint32_t len = cb->len[s];
                        unsigned long long code = cb->code[s];
                        if (len > 32) { put_bits(stage, acc, nb, w, shared_first, (uint32_t)(code >> 32), len - 32); len = 32; }
                        put_bits(stage, acc, nb, w, shared_first, (uint32_t)code, len);
                    }
                } else {
                    put_bits(stage, acc, nb, w, shared_first, e[g][j] & 0x07FFFFFFu, e[g][j] >> 27);
                }
            }
            if (nb) atomicOr(&stage[w], (uint32_t)(acc >> 32));     // tail word shared with the next thread
        }
        __syncthreads();

        // ---- 4. seam handling, then copy out ----
        const unsigned long long first_w = gbit >> 5;               // frame word holding the tile's first bit
        const unsigned long long end_w = gend >> 5;                 // frame word holding the bit after the tile
        const bool last_tile = tile + 1 == ntiles;
        if (tid == 0) {
            const bool tail_is_seam = end_w == first_w;             // tiny tile: tail word needs the predecessor's bits
            if (!tail_is_seam && !last_tile)
                st_release_u64(&tails[tile], (1ull << 63) | stage[end_w - G0]);
            uint32_t carry = 0;
            if (tile == 0) {
                // preserve the bits of the first byte that precede the start phase
                uint32_t b = frame[gbit >> 3];
                uint32_t keep = b & ~(0xFFu >> (gbit & 7));
                carry = keep << (24 - 8 * (uint32_t)((gbit >> 3) & 3));
            } else {
                unsigned long long t;
                uint32_t spins = 0;
                while (((t = ld_acquire_u64(&tails[tile - 1])) >> 63) == 0)
                    if (++spins > SPIN_LIMIT) { atomicExch(&work->error, 2ull); break; }
                carry = (uint32_t)t;
            }
            stage[first_w - G0] |= carry;
            if (tail_is_seam && !last_tile)
                st_release_u64(&tails[tile], (1ull << 63) | stage[end_w - G0]);
        }
        __syncthreads();

        // words [first_w, end_w) are complete.  The very first word of the stream is stored
        // bytewise from the start byte on; the last tile also stores the final partial bytes.
        const uint32_t n_groups = (uint32_t)((end_w - G0 + 4) >> 2);
        uint32_t *gw = reinterpret_cast<uint32_t *>(frame);
        for (uint32_t q = tid; q < n_groups; q += ENC_THREADS) {
            const unsigned long long w0 = G0 + 4ull * q;
            uint4 v = reinterpret_cast<const uint4 *>(stage)[q];
            v.x = bswap32(v.x); v.y = bswap32(v.y); v.z = bswap32(v.z); v.w = bswap32(v.w);
            const bool has_stream_head = (tile == 0) && (w0 <= first_w);
            if (w0 >= first_w && w0 + 4 <= end_w && !has_stream_head) {
                st_stream_v4(gw + w0, v);
            } else {
                uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned long long wk = w0 + k;
                    if (wk < first_w || wk > end_w) continue;
                    uint32_t b_lo = 0, b_hi = 4;                    // byte range [b_lo, b_hi) of this word to store
                    if (tile == 0 && wk == first_w) b_lo = (uint32_t)((gbit >> 3) & 3);
                    if (wk == end_w) {
                        if (!last_tile) continue;                   // completed by the next tile
                        b_hi = (uint32_t)(((gend & 31) + 7) >> 3);  // bytes holding payload bits
                    }
                    if (b_lo == 0 && b_hi == 4) gw[wk] = vv[k];
                    else
                        for (uint32_t b = b_lo; b < b_hi; b++)
                            frame[wk * 4 + b] = (uint8_t)(vv[k] >> (8 * b));   // little-endian view of the swapped word
                }
            }
        }
    }
}

static size_t enc_work_bytes(uint32_t ntiles) { return sizeof(EncWork) + (size_t)ntiles * 16; }

int launch_encode(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb, uint8_t *d_stream,
                  uint64_t start_bit, uint32_t maxlen_hint)
{
    const uint64_t n_sym = n_bytes / 2;
    if (n_sym == 0) return HF_OK;
    if ((uintptr_t)d_in & 1) return set_err(c, HF_ERR_ARG, "hf_encode: input must be 2-byte aligned");
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
    const size_t smem_fast = EncCfg<false>::STAGE_WORDS * 4, smem_long = EncCfg<true>::STAGE_WORDS * 4;
    if (!attr_set) {
        HF_CUDA(c, cudaFuncSetAttribute(encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
        HF_CUDA(c, cudaFuncSetAttribute(encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_long));
        attr_set = true;
    }
    // start_bit may exceed 8: fold whole bytes into the pointer
    d_stream += start_bit >> 3;
    start_bit &= 7;
    int occ_fast = 0, occ_long = 0;
    HF_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_fast, encode_kernel<false>, ENC_THREADS, smem_fast));
    HF_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_long, encode_kernel<true>, ENC_THREADS, smem_long));
    if (maxlen_hint == 0 || maxlen_hint <= ENC32_MAX_LEN) {
        uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)c->sm_count * (occ_fast > 0 ? occ_fast : 1));
        HF_PROF(c, "encode_kernel<false>"); encode_kernel<false><<<grid, ENC_THREADS, smem_fast, c->stream>>>(d_in, n_sym, d_cb, d_stream, start_bit, work, ntiles);
        HF_LAUNCH_CHECK(c);
    }
    if (maxlen_hint == 0 || maxlen_hint > ENC32_MAX_LEN) {
        uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)c->sm_count * (occ_long > 0 ? occ_long : 1));
        HF_PROF(c, "encode_kernel<true>"); encode_kernel<true><<<grid, ENC_THREADS, smem_long, c->stream>>>(d_in, n_sym, d_cb, d_stream, start_bit, work, ntiles);
        HF_LAUNCH_CHECK(c);
    }
    return HF_OK;
}

}  // namespace hf
