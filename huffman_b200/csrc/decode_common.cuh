// decode_common.cuh — bit readers and the table walk shared by the exact decoder (decode.cu) and the
// word-walk kernels (decode2.cu).  Private to libhuffb200.
#pragma once
#include "common.cuh"

namespace hf {

constexpr uint32_t E_SUB = 0x80u;                               // entry flag: sub-table / escape
constexpr uint32_t E_LIST = 0x80000000u;                           // level-2 slot: head of a long-code list

// -----------------------------------------------------------------------------------
// bit access.  Fetch functors return big-endian 32-bit word i of some bit string.
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
__device__ __forceinline__ unsigned long long peek64(const F &f, uint32_t bitpos)
{
    uint32_t i = bitpos >> 5, sh = bitpos & 31;
    unsigned long long hi = ((unsigned long long)f(i) << 32) | f(i + 1);
    if (sh == 0) return hi;
    return (hi << sh) | ((unsigned long long)f(i + 2) >> (32 - sh));
}

// ---- geometry and global work area of the exact decoder (decode.cu, decode2.cu) ----
constexpr int DEC_THREADS = 512;
constexpr uint32_t SUB_BITS = 256;                              // bits per subsequence (thread)
constexpr uint32_t CHUNK_BITS = DEC_THREADS * SUB_BITS;         // 131072 bits = 16 KiB
// the synchronisation kernel's unit of convergence (a "group"): one chunk, converged on by one warp (decode2.cu)
constexpr uint32_t GROUP_CHUNKS = 1;
// a speculative range call (hf_range_overflow) works on the last TAIL_CHUNKS chunks of the range, from a group boundary
constexpr unsigned long long TAIL_CHUNKS = 16;
__host__ __device__ inline unsigned long long tail_first_chunk(unsigned long long nch)
{
    const unsigned long long ngroups = (nch + GROUP_CHUNKS - 1) / GROUP_CHUNKS, tg = TAIL_CHUNKS / GROUP_CHUNKS;
    return (ngroups > tg ? ngroups - tg : 0) * GROUP_CHUNKS;
}


// NO_START: no exact start is known (a speculative pass: every span starts from a guess)
constexpr unsigned long long NO_START = ~0ull;
struct DecWork {
    unsigned long long result[4];       // [1] overflow of the last code word past the range end, [2] symbols in the range
    unsigned long long flags[4];        // [0] a group does not start where the one before ends (serial kernel needed),
                                        // [1] invalid code met, [2] output capacity too small, [3] a group was left to dec_regroup4_kernel
    // where the decode starts and how much it may write, DEVICE resident so that a header parsed on the device or a
    // hand-over bit that arrives by a collective never has to visit the host: [0] frame bit of the first code word
    // (F0; chunks and lanes before it hold no code word), [1] symbols to write at most, [2] spare, [3] spare
    unsigned long long start[4];
    // followed by: chunkBase[nch] u64, chunkCnt[nch] u32, chunkE[nch] u32 (overflow past the chunk's end),
    //              chunkE2[nch] u32 (0xFFFFFFFF, or CHUNK_DIRTY on the first chunk of a group that must be redone),
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
