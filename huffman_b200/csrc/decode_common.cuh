// decode_common.cuh — bit readers and the table walk shared by the exact decoder (decode.cu) and the
// single-pass decoder (decode_fast.cu).  Private to libhuffb200.
#pragma once
#include "common.cuh"

namespace hf {

constexpr uint32_t E_SUB = 0x80u;                               // entry flag: sub-table / escape
constexpr uint32_t E_LIST = 0x80000000u;                           // level-2 slot: head of a long-code list

// -----------------------------------------------------------------------------------
// bit access.  Fetch functors return big-endian 32-bit word i of some bit string.
// The chunk's words live in shared memory with one pad word after every 8 (a subsequence is 8 words):
// lane t reading word j of ITS subsequence hits bank (9t + j) % 32, so the lanes of a warp do not collide.
__device__ __forceinline__ uint32_t smem_word_index(uint32_t i) { return i + (i >> 3); }
constexpr uint32_t smem_words_padded(uint32_t n) { return n + (n >> 3) + 1; }
struct SmemFetch {
    const uint32_t *w;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return w[smem_word_index(i)]; }
};
struct GlobalFetch {                    // frame words straight from global memory, zero past the end
    const uint8_t *frame;
    unsigned long long frame_bytes;
    unsigned long long word0;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const
    {
        unsigned long long b = (word0 + i) * 4ull;
        if (b + 4 <= frame_bytes) return bswap32(*reinterpret_cast<const uint32_t *>(frame + b));
        uint32_t v = 0;
        for (int k = 0; k < 4; k++)
            if (b + k < frame_bytes) v |= (uint32_t)frame[b + k] << (24 - 8 * k);
        return v;
    }
};

template <typename F>
struct BitReader {
    F f;
    uint32_t wi;                        // next word to pull
    unsigned long long win;             // upcoming bits, left aligned
    uint32_t avail;                     // valid bits in win (kept > 32)
    __device__ __forceinline__ void init(uint32_t bitpos)
    {
        wi = bitpos >> 5;
        uint32_t sh = bitpos & 31;
        win = (((unsigned long long)f(wi) << 32) | f(wi + 1)) << sh;
        avail = 64 - sh;
        wi += 2;
        if (avail <= 32) { win |= (unsigned long long)f(wi++) << (32 - avail); avail += 32; }
    }
    __device__ __forceinline__ void consume(uint32_t n)     // n <= 32 per call
    {
        win <<= n;
        avail -= n;
        if (avail <= 32) { win |= (unsigned long long)f(wi++) << (32 - avail); avail += 32; }
    }
    __device__ __forceinline__ void skip(uint32_t n)
    {
        while (n > 32) { consume(32); n -= 32; }
        consume(n);
    }
};

template <typename F>
__device__ __forceinline__ unsigned long long peek64(const F &f, uint32_t bitpos)
{
    uint32_t i = bitpos >> 5, sh = bitpos & 31;
    unsigned long long hi = ((unsigned long long)f(i) << 32) | f(i + 1);
    if (sh == 0) return hi;
    return (hi << sh) | ((unsigned long long)f(i + 2) >> (32 - sh));
}

// 32 bits of the staged chunk starting at bit `b` (stateless: two loads and a funnel shift, no refill branch)
__device__ __forceinline__ uint32_t smem_window32(const uint32_t *sw, uint32_t b)
{
    const uint32_t i = b >> 5;
    return __funnelshift_l(sw[smem_word_index(i + 1)], sw[smem_word_index(i)], b & 31);
}

struct TabView {
    const uint32_t *t1;                 // shared (kernels A, C) or global (kernel B)
    const uint32_t *t2;
    const LongCode *longs;
    uint32_t n_long;
};

// one code word at bit `b` of the staged chunk; returns (sym << 8) | len, len >= 1.  K1 + sub bits <= 24 <= 32.
__device__ __forceinline__ uint32_t decode_at(const TabView &T, const uint32_t *sw, uint32_t b, uint32_t &bad)
{
    const uint32_t win = smem_window32(sw, b);
    uint32_t e = T.t1[win >> (32 - K1)];
    if (e & E_SUB) {
        const uint32_t sb = e & 31u;
        const uint32_t idx2 = (win << K1) >> (32 - sb);
        e = __ldg(&T.t2[(e >> 8) + idx2]);
        if (e & E_LIST) {               // longer than K1 + sub bits: walk the slot's list of long codes
            const unsigned long long w64 = peek64(SmemFetch{sw}, b);
            uint32_t cur = e;
            e = 0;
            while (cur & E_LIST) {
                const LongCode lc = T.longs[(cur >> 8) & 0xFFFFu];
                if (((w64 ^ lc.code_left) >> (64 - (lc.leaf & 0x7Fu))) == 0) { e = lc.leaf; break; }
                cur = lc.next;
            }
        }
    }
    if (e == 0) { bad = 1; e = 1; }     // hole in the code: flag it, step one bit so the walk ends
    return e;
}

// one code word at the reader's position; returns (sym << 8) | len, len >= 1
template <typename F>
__device__ __forceinline__ uint32_t decode_one(const TabView &T, const BitReader<F> &r, uint32_t bitpos,
                                               uint32_t &bad)
{
    uint32_t e = T.t1[(uint32_t)(r.win >> (64 - K1))];
    if (e & E_SUB) {
        const uint32_t sb = e & 31u;
        const uint32_t idx2 = (uint32_t)((r.win << K1) >> (64 - sb));
        e = __ldg(&T.t2[(e >> 8) + idx2]);
        if (e & E_LIST) {               // longer than K1 + sub bits: walk the slot's list of long codes
            const unsigned long long w64 = peek64(r.f, bitpos);
            uint32_t cur = e;
            e = 0;
            while (cur & E_LIST) {
                const LongCode lc = T.longs[(cur >> 8) & 0xFFFFu];
                if (((w64 ^ lc.code_left) >> (64 - (lc.leaf & 0x7Fu))) == 0) { e = lc.leaf; break; }
                cur = lc.next;
            }
        }
    }
    if (e == 0) { bad = 1; e = 1; }     // hole in the code: flag it, step one bit so the walk ends
    return e;
}

// ---- geometry and global work area of the exact decoder (decode.cu, decode2.cu) ----
constexpr int DEC_THREADS = 512;
constexpr uint32_t SUB_BITS = 256;                              // bits per subsequence (thread)
constexpr uint32_t CHUNK_BITS = DEC_THREADS * SUB_BITS;         // 131072 bits = 16 KiB
constexpr uint32_t CHUNK_WORDS = CHUNK_BITS / 32;               // 4096
constexpr uint32_t CHUNK_PAD_WORDS = 8;                         // look-ahead past the chunk
constexpr uint32_t WIN_SYMS = 16384;                            // output staging window (symbols)
constexpr uint32_t SW_PADDED = (smem_words_padded(CHUNK_WORDS + CHUNK_PAD_WORDS) + 3) & ~3u;   // staged chunk, padded layout

// result flags of the single-pass decoder (decode_fast.cu) that send the job to the exact kernels below
constexpr unsigned long long DF_GATE_MASK = 1 | 2 | 4 | 16;

struct DecWork {
    unsigned long long result[4];       // [1] overflow of the last code word past the range end, [2] symbols in the range
    unsigned long long flags[4];        // [0] any chunk failed to sync, [1] invalid code met, [2] table error
    // followed by: chunkBase[nch] u64, chunkCnt[nch] u32, chunkE[nch] u32, chunkE2[nch] u32 (0xFFFFFFFF = unchanged),
    //              info[nch * DEC_THREADS] u16
};

struct DecLayout {
    unsigned long long *chunkBase;
    uint32_t *chunkCnt, *chunkE, *chunkE2;
    uint16_t *info;
    static size_t bytes(uint64_t nch) { return sizeof(DecWork) + nch * (8 + 4 + 4 + 4 + 2 * (size_t)DEC_THREADS) + 16; }
    __host__ __device__ DecLayout(DecWork *w, uint64_t nch)
    {
        uint8_t *p = reinterpret_cast<uint8_t *>(w + 1);
        chunkBase = reinterpret_cast<unsigned long long *>(p); p += nch * 8;
        chunkCnt = reinterpret_cast<uint32_t *>(p); p += nch * 4;
        chunkE = reinterpret_cast<uint32_t *>(p); p += nch * 4;
        chunkE2 = reinterpret_cast<uint32_t *>(p); p += nch * 4;
        info = reinterpret_cast<uint16_t *>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);   // 64-bit stores of 4 records
    }
};


// bits of subsequence t of chunk c that lie before the end of the range (0 .. SUB_BITS)
__device__ __forceinline__ uint32_t sub_limit(unsigned long long c, uint32_t t, unsigned long long range_end_bit)
{
    const unsigned long long x = c * CHUNK_BITS + (unsigned long long)t * SUB_BITS;
    if (x >= range_end_bit) return 0u;
    const unsigned long long room = range_end_bit - x;
    return room >= SUB_BITS ? SUB_BITS : (uint32_t)room;
}

__device__ __forceinline__ uint32_t spec_start(unsigned long long X, unsigned long long F0, uint32_t g)
{   // first offset >= 0 from frame bit X at which a code word can start (boundaries are F0 + k*g)
    if (g <= 1) return 0;
    uint32_t r = (uint32_t)((X - F0) % g);
    return r ? g - r : 0;
}


}  // namespace hf
