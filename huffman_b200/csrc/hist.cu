// hist.cu — 65,536-bin histogram of little-endian byte pairs (replaces
// calculateFrequency, /root/reference/Compressor.cu:38-48, one global atomic per symbol).
//
// Design (B200): one persistent CTA per SM holds ALL 65,536 bins in shared memory as
// packed 16-bit counters (2 per 32-bit word, 128 KiB of the SM's 227 KiB), so every
// count is one shared-memory atomic and the input is read exactly once with 128-bit
// streaming loads.  A 16-bit field that reaches SPILL (bit 14) is spilled (SPILL counts moved
// to the 64-bit global bin) by the one thread that saw the crossing, so fields never wrap.
// The bin index is XOR-folded (low byte ^= high byte, an involution) so that text-like
// inputs, whose first bytes cluster, still spread over the 32 banks.
// CTA partials go to global memory with coalesced stores and a second small kernel
// folds them into the caller's u64 histogram (no 65,536 x 148 contended global atomics).
//
// Algorithmic bytes: N read.  Roofline: HBM.
#include "common.cuh"

namespace hf {

constexpr int HIST_THREADS = 1024;
constexpr int HIST_WORDS = 32768;           // packed u16 pairs
constexpr int HIST_UNROLL = 2;              // 2 x 16 B counted while the next 2 x 16 B are in flight, per thread
constexpr uint64_t HIST_SMALL_BYTES = 1u << 18;

__device__ __forceinline__ uint32_t fold(uint32_t sym) { return sym ^ (sym >> 8); }   // involution on 16 bits

// The spill is two steps: the add that sets the field's bit 14, then the subtraction of SPILL.  Other threads keep adding
// to the field in between; it stays inside its 16 bits as long as fewer than 0x10000 - 2 * SPILL + 1 = 32,769 counts land
// in that window — an add is at most 256 (a warp's hot symbol over one step) and the window is a few instructions of the
// thread that saw the crossing, so the 1024 threads of the CTA would have to deliver over a hundred such adds each within it.
// (Round 1 spilled at bit 15, which left a window of 0x7F00 counts: the advisor's finding.)
constexpr uint32_t SPILL = 0x4000u;

// adds n (<= 256) to the bin of `sym`
__device__ __forceinline__ void count_sym(uint32_t *sh, unsigned long long *ghist, uint32_t sym, uint32_t n = 1)
{
    uint32_t ix = fold(sym);
    uint32_t w = ix & 0x7FFFu;
    uint32_t sa = (ix >> 11) & 16u;                 // field select: 0 or 16
    uint32_t inc = n << sa;
    uint32_t old = atomicAdd(&sh[w], inc);
    if (((old + inc) & ~old) & (SPILL << sa)) {     // my add set bit 14 of the field
        atomicSub(&sh[w], SPILL << sa);
        atomicAdd(&ghist[sym], (unsigned long long)SPILL);
    }
}

// Warp aggregation of the hot bin: atomics of lanes that hit the SAME shared-memory word serialise, and skewed
// inputs (text, Zipf, two-valued data) send most lanes to one bin.  The first symbol of lane 0 is taken as the
// warp's candidate: its occurrences in the 256 symbols of this step are counted in registers and added with ONE
// atomic; everything else goes bin by bin.  A candidate that pays (>= ~10 % of the symbols) is kept for the next
// steps; otherwise the warp counts bin by bin for 15 steps and then tries the symbol it meets first.
// Returns how often the candidate occurred (warp-wide), so the caller can stop trying on flat inputs.
__device__ __forceinline__ uint32_t count_vec(uint32_t *sh, unsigned long long *ghist, const uint4 &v, uint32_t hot)
{
    const uint32_t s[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                           v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
    uint32_t nh = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) nh += (s[j] == hot);
    const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, nh);
    if ((threadIdx.x & 31) == 0) count_sym(sh, ghist, hot, tot);
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (s[j] != hot) count_sym(sh, ghist, s[j]);
    return tot;
}

// the same bin by bin, for the lanes of a partly filled warp
__device__ __forceinline__ void count_vec_partial(uint32_t *sh, unsigned long long *ghist, const uint4 &v)
{
    count_sym(sh, ghist, v.x & 0xFFFFu); count_sym(sh, ghist, v.x >> 16);
    count_sym(sh, ghist, v.y & 0xFFFFu); count_sym(sh, ghist, v.y >> 16);
    count_sym(sh, ghist, v.z & 0xFFFFu); count_sym(sh, ghist, v.z >> 16);
    count_sym(sh, ghist, v.w & 0xFFFFu); count_sym(sh, ghist, v.w >> 16);
}

__global__ void __launch_bounds__(HIST_THREADS, 1)
hist_smem_kernel(const uint4 *__restrict__ in, uint64_t n_vec, uint32_t *__restrict__ partials,
                 unsigned long long *__restrict__ ghist)
{
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < HIST_WORDS / 4; i += HIST_THREADS)
        reinterpret_cast<uint4 *>(sh)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const uint64_t step = (uint64_t)gridDim.x * HIST_THREADS * HIST_UNROLL;
    uint64_t i = (uint64_t)blockIdx.x * HIST_THREADS * HIST_UNROLL + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;         // whole warps only: count_vec uses warp-wide operations
    uint32_t skip = 0;                              // steps for which this warp does not try to aggregate
    uint32_t hot = 0x10000u;                        // the warp's candidate symbol (none yet)
    // software pipeline: HIST_UNROLL vectors are counted while the next HIST_UNROLL are in flight
    auto whole = [&](uint64_t at) { return (at - lane) + 31 + (HIST_UNROLL - 1) * HIST_THREADS < n_vec; };
    uint4 nv[HIST_UNROLL];
    bool have = whole(i);
    if (have) {
#pragma unroll
        for (int j = 0; j < HIST_UNROLL; j++) nv[j] = ld_stream_v4(in + i + j * HIST_THREADS);
    }
    while (have) {
        uint4 v[HIST_UNROLL];
#pragma unroll
        for (int j = 0; j < HIST_UNROLL; j++) v[j] = nv[j];
        const bool more = whole(i + step);
        if (more) {
#pragma unroll
            for (int j = 0; j < HIST_UNROLL; j++) nv[j] = ld_stream_v4(in + i + step + j * HIST_THREADS);
        }
        if (skip == 0) {
            if (hot > 0xFFFFu) hot = __shfl_sync(0xFFFFFFFFu, v[0].x & 0xFFFFu, 0);
            uint32_t hits = 0;
#pragma unroll
            for (int j = 0; j < HIST_UNROLL; j++) hits += count_vec(sh, ghist, v[j], hot);
            if (hits < HIST_UNROLL * 24) { skip = 15 * (4 / HIST_UNROLL); hot = 0x10000u; }   // under ~10 % of the symbols: not worth it
        } else {
            skip--;
#pragma unroll
            for (int j = 0; j < HIST_UNROLL; j++) count_vec_partial(sh, ghist, v[j]);
        }
        i += step;
        have = more;
    }
    for (; i < n_vec; i += HIST_THREADS)            // ragged end of this CTA's last strip, bin by bin
        count_vec_partial(sh, ghist, ld_stream_v4(in + i));
    __syncthreads();

    uint4 *dst = reinterpret_cast<uint4 *>(partials + (size_t)blockIdx.x * HIST_WORDS);
    for (int k = threadIdx.x; k < HIST_WORDS / 4; k += HIST_THREADS)
        dst[k] = reinterpret_cast<uint4 *>(sh)[k];
}

// folds the CTA partials into the caller's histogram; one thread per packed word
__global__ void hist_reduce_kernel(const uint32_t *__restrict__ partials, int n_parts,
                                   unsigned long long *__restrict__ ghist)
{
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= HIST_WORDS) return;
    unsigned long long a0 = 0, a1 = 0;
#pragma unroll 4
    for (int b = 0; b < n_parts; b++) {
        uint32_t v = partials[(size_t)b * HIST_WORDS + w];
        a0 += v & 0xFFFFu;
        a1 += v >> 16;
    }
    if (a0) ghist[fold(w)] += a0;
    if (a1) ghist[fold(w | 0x8000u)] += a1;
}

// heads, tails and tiny inputs: straight 64-bit global atomics
__global__ void hist_scalar_kernel(const uint16_t *__restrict__ in, uint64_t n_sym,
                                   unsigned long long *__restrict__ ghist)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_sym; i += stride) atomicAdd(&ghist[in[i]], 1ull);
}

int launch_histogram(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, unsigned long long *d_hist)
{
    uint64_t n_sym = n_bytes / 2;
    if (n_sym == 0) return HF_OK;
    if ((uintptr_t)d_in & 1) return set_err(c, HF_ERR_ARG, "hf_histogram: input must be 2-byte aligned");

    const uint16_t *p = reinterpret_cast<const uint16_t *>(d_in);
    if (n_bytes < HIST_SMALL_BYTES) {
        int blocks = (int)((n_sym + 255) / 256);
        HF_PROF(c, "hist_scalar_kernel"); hist_scalar_kernel<<<blocks, 256, 0, c->stream>>>(p, n_sym, d_hist);
        HF_LAUNCH_CHECK(c);
        return HF_OK;
    }
    // symbols before the first 16-byte boundary and after the last one: scalar path
    uint64_t head = ((16 - ((uintptr_t)d_in & 15)) & 15) / 2;
    if (head > n_sym) head = n_sym;
    uint64_t n_vec = (n_sym - head) / 8;
    uint64_t tail = n_sym - head - n_vec * 8;
    if (head) { HF_PROF(c, "hist_scalar_kernel"); hist_scalar_kernel<<<1, 32, 0, c->stream>>>(p, head, d_hist); HF_LAUNCH_CHECK(c); }
    if (tail) {
        HF_PROF(c, "hist_scalar_kernel"); hist_scalar_kernel<<<1, 32, 0, c->stream>>>(p + head + n_vec * 8, tail, d_hist);
        HF_LAUNCH_CHECK(c);
    }
    if (n_vec) {
        uint64_t per_cta = (uint64_t)HIST_THREADS * HIST_UNROLL;          // vectors per CTA strip
        int grid = (int)((n_vec + per_cta - 1) / per_cta);
        if (grid > c->sm_count) grid = c->sm_count;
        if ((size_t)grid * HIST_WORDS * sizeof(uint32_t) > WS_SCRATCH_BYTES) grid = (int)(WS_SCRATCH_BYTES / (HIST_WORDS * sizeof(uint32_t)));
        size_t part_bytes = (size_t)grid * HIST_WORDS * sizeof(uint32_t);
        int rc = ensure_ws(c, part_bytes);
        if (rc) return rc;
        if (!c->smem_attr[ATTR_HIST]) {
            HF_CUDA(c, cudaFuncSetAttribute(hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            HIST_WORDS * 4));
            c->smem_attr[ATTR_HIST] = true;
        }
        HF_PROF(c, "hist_smem_kernel"); hist_smem_kernel<<<grid, HIST_THREADS, HIST_WORDS * 4, c->stream>>>(
            reinterpret_cast<const uint4 *>(p + head), n_vec, (uint32_t *)c->ws, d_hist);
        HF_LAUNCH_CHECK(c);
        HF_PROF(c, "hist_reduce_kernel"); hist_reduce_kernel<<<HIST_WORDS / 256, 256, 0, c->stream>>>((const uint32_t *)c->ws, grid, d_hist);
        HF_LAUNCH_CHECK(c);
    }
    return HF_OK;
}

}  // namespace hf
