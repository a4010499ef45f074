// decode2.cu — the hot kernels of the decoder: dec_sync4_kernel (code word boundaries and symbol counts per
// 256-bit subsequence), dec_fix2_kernel (inter-group repair) and dec_write3_kernel (symbols out), plus the
// shared-memory table planes they read.  decode.cu holds the header parser, the general tables, the scan and
// the orchestration; the kernels meet in the work area of decode_common.cuh (DecLayout: info / chunkCnt / chunkE).
//
// The reference decodes on the host, one bit and one fread() per step, by chasing tree pointers
// (/root/reference/Decompressor.cu:259-291); its format has no offset index.
//
// Design notes (the first kernels of this round staged the payload per chunk, kept a 64-bit bit reader and a
// 12-bit first-level table: ncu showed 76 + 65 thread instructions per symbol, 39 % of the lanes active and
// 55 % of the symbols of the 16 GiB bench stream looked up in global memory):
//   * the write kernel keeps a thread's 256-bit subsequence (+ one look-ahead word) in REGISTERS, loaded
//     straight from global memory; the walk is unrolled over the words, so the 32-bit window of a code word
//     is a funnel shift of two registers — no shared-memory staging of the payload, no bit reader;
//   * the table of the hot loop is ONE 64 KiB shared-memory plane indexed by 14 bits whose entries are either a
//     code word of at most 14 bits or a "micro tree": the shape (leaf-start bits) of the complete
//     depth-4 subtree below that prefix, from which the length of a 15..18-bit code is a few bit operations and
//     its symbol one more shared-memory load (leaves[base + rank], 128 KiB, write kernel only).  The 65,536
//     byte pairs of a flat or mixed stream get 17/18-bit codes, so a plain direct table of that depth would
//     need 1 MiB; here every code of up to 18 bits is resolved without leaving the SM.  Deeper or incomplete
//     prefixes take ONE gather from a flat second-level plane (<= 22 bits, L2 resident); anything beyond goes
//     through decode.cu's tables;
//   * the synchronisation kernel gives every warp a whole chunk and every lane a 512-byte span of it (round 1 gave a
//     thread 128 bytes and converged in teams of four warps behind named barriers: 40 % of its walk iterations were
//     repeated walks, 17 % of its stalls barrier waits); the per-subsequence records double as the checkpoints at which
//     a repeated walk recognises the earlier one;
//   * what the per-source-line and per-SASS-instruction views of ncu found later is listed in profiles/README.md
//     (shared-memory base addresses rebuilt inside loops, loop state in local memory around a call, loads that
//     were waited for where they were issued, single-lane tails).
//
// Algorithmic bytes: dec_sync4 reads C; dec_write3 reads C and writes N.
#include "common.cuh"
#include "decode_common.cuh"

namespace hf {

constexpr uint32_t MICRO_D = 4;                                 // depth of a micro tree
constexpr uint32_t MICRO_MAX = MICRO_K + MICRO_D;               // longest code resolved in shared memory (18)
// t14 entry of a micro tree: base << 16 | MICRO_FLAG | starts of slots 1..15 (bit j - 1 = a leaf starts at slot j;
// slot 0 always starts one).  A short code's entry, sym << 16 | len << 1, never has bit 15 set.
constexpr uint32_t MICRO_FLAG = 0x8000u;
__host__ __device__ __forceinline__ uint32_t micro_starts(uint32_t e) { return ((e & 0x7FFFu) << 1) | 1u; }
__device__ __forceinline__ uint32_t micro_leaves(uint32_t e) { return __popc(e & 0x7FFFu) + 1u; }
#ifndef W3_WARPS
#define W3_WARPS 24                                             // warps of the write kernel's CTA
#endif
constexpr int W3_THREADS = W3_WARPS * 32;
#ifndef W3_PHASE
#define W3_PHASE 2                                              // words per walk phase of the write kernel (1, 2, 4)
#endif
// output staging window of one warp (symbols, multiple of 8): what is left of the SM's 227 KiB beside the planes
constexpr uint32_t W3_WIN = (((232448u - (4u << MICRO_K) - NSYM * 2 - 64u) / W3_WARPS) / 2 - 8) & ~7u;
constexpr uint32_t W3_BAR = (4u << MICRO_K) + NSYM * 2 + W3_WARPS * (W3_WIN + 8) * 2;    // the mbarrier of the plane load
constexpr size_t W3_SMEM = W3_BAR + 16;
static_assert(W3_SMEM <= 232448, "dec_write3: shared memory");

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
    k2 = k2 < MICRO_K ? MICRO_K : (k2 > FLAT_MAX ? FLAT_MAX : k2);
    if (i == 0) tab->k2 = k2;
    if (i < (1u << MICRO_K)) {
        // entry of the 14-bit prefix i: a short code, or the shape of the micro tree below it
        const uint32_t w0 = i << (32 - MICRO_K);
        const uint32_t e0 = lookup_win32(tab, w0);
        uint32_t entry = 0, dw = 0x10Cu;                  // not here: no bits, counted as one code word by the flat path
        if (e0 && (e0 & 0x7Fu) <= MICRO_K) {
            entry = ((e0 >> 8) << 16) | ((e0 & 0x7Fu) << 1);
            // the lengths-only twin also says how many code words these 14 bits hold completely, and their bits
            uint32_t tot = e0 & 0x7Fu, cnt = 1;
            for (;;) {
                const uint32_t e = lookup_win32(tab, w0 << tot);
                const uint32_t len = e & 0x7Fu;
                if (e == 0 || len == 0 || tot + len > MICRO_K) break;
                tot += len;
                cnt++;
            }
            dw = ((e0 & 0x7Fu) << 28) | (cnt << 8) | (tot << 4) | 0xCu;
        } else {
            uint32_t mask = 0, depths = 0;
            bool ok = true;
            for (uint32_t j = 0; j < (1u << MICRO_D); j++) {
                const uint32_t e = lookup_win32(tab, w0 | (j << (32 - MICRO_MAX)));
                const uint32_t len = e & 0x7Fu;
                if (e == 0 || len <= MICRO_K || len > MICRO_MAX) { ok = false; break; }
                tab->micro_sym[i * (1u << MICRO_D) + j] = (uint16_t)(e >> 8);
                if ((j & ((1u << (MICRO_MAX - len)) - 1u)) == 0) mask |= 1u << j;       // first slot of its leaf
                depths |= (len - MICRO_K - 1u) << (2 * j);
            }
            // the two planes must tell the same lengths: every leaf an aligned block of slots of one depth.  A prefix
            // code always is; a damaged table that is not goes to the flat planes, which both kernels read alike.
            for (uint32_t j = 0; ok && j < (1u << MICRO_D); j++) {
                const uint32_t d = (depths >> (2 * j)) & 3u, first = j & ~((8u >> d) - 1u);
                if (((depths >> (2 * first)) & 3u) != d) ok = false;
            }
            if (ok) { entry = (mask >> 1) | MICRO_FLAG; dw = depths; }      // the base comes from dt_micro_kernel
        }
        tab->t14[i] = entry;
        tab->d14[i] = dw;

    }
    if (i < (1u << k2)) {
        const uint32_t e = lookup_win32(tab, i << (32 - k2));
        const bool ok = e && (e & 0x7Fu) <= k2;
        tab->flat2[i] = ok ? e : 0u;
        tab->lenflat[i] = ok ? (uint8_t)(e & 0x7Fu) : (uint8_t)0;
    }
}

// bases of the micro trees: exclusive scan of their leaf counts in prefix order (one CTA, 16 prefixes per thread)
__global__ void __launch_bounds__(1024, 1)
dt_micro_kernel(DecodeTable *__restrict__ tab)
{
    __shared__ uint32_t s_w[33];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr uint32_t PER = (1u << MICRO_K) / 1024;    // 16 prefixes per thread
    uint4 v[PER / 4];
    uint4 *ent = reinterpret_cast<uint4 *>(tab->t14 + tid * PER);
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < PER / 4; j++) {
        v[j] = ent[j];
        const uint32_t e4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int k = 0; k < 4; k++) if (e4[k] & MICRO_FLAG) sum += micro_leaves(e4[k]);
    }
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
        const uint32_t t0 = s_w[lane];
        uint32_t t = t0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
        s_w[lane] = t - t0;
    }
    __syncthreads();
    uint32_t base = x - sum + s_w[wid];
#pragma unroll
    for (uint32_t j = 0; j < PER / 4; j++) {
        uint32_t e4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!(e4[k] & MICRO_FLAG)) continue;
            const uint32_t shape = e4[k] & 0xFFFFu, n = micro_leaves(shape);
            e4[k] = base + n > NSYM ? 0u : ((base << 16) | shape);      // cannot overflow for a prefix code
            base += n;
        }
        ent[j] = make_uint4(e4[0], e4[1], e4[2], e4[3]);
    }
}

// the leaf symbols of every micro tree, in slot order
__global__ void dt_leaves_kernel(DecodeTable *__restrict__ tab)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (1u << MICRO_K)) return;
    const uint32_t e = tab->t14[p];
    if (!(e & MICRO_FLAG)) return;
    uint32_t base = e >> 16;
    for (uint32_t m = micro_starts(e); m; m &= m - 1)
        tab->leaves[base++] = tab->micro_sym[p * (1u << MICRO_D) + (__ffs(m) - 1)];
}

int launch_table_planes(Ctx *c, DecodeTable *d_tab)
{
    HF_PROF(c, "dt_planes_kernel"); dt_planes_kernel<<<(1u << FLAT_MAX) / 256, 256, 0, c->stream>>>(d_tab);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dt_micro_kernel"); dt_micro_kernel<<<1, 1024, 0, c->stream>>>(d_tab);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dt_leaves_kernel"); dt_leaves_kernel<<<(1u << MICRO_K) / 256, 256, 0, c->stream>>>(d_tab);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

__device__ __forceinline__ uint32_t msb(uint32_t x)              // position of the highest set bit (x != 0)
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

// Code word at the head of `win` from a micro-tree entry e: its length and the index of its leaf.  b = the 4 bits
// after the 14-bit prefix.  Starts at or below b, moved to the top of a word: the distance down to the leaf's start is
// a count of leading zeros, its rank a population count; the starts above b, bit-reversed and moved to the top, give
// the distance up to the next leaf the same way (the flag bit is the start of "slot 16").
__device__ __forceinline__ void micro_decode(uint32_t e, uint32_t win, uint32_t &len, uint32_t &leaf)
{
    const uint32_t b = (win >> (32 - MICRO_MAX)) & ((1u << MICRO_D) - 1u);
    const uint32_t below = ((e << 1) | 1u) << (31u - b);        // slot b at bit 31, slot 0 (always a start) at 31 - b
    const uint32_t above = __brev(e) << b;                      // slot b + 1 at bit 31
    const uint32_t slots = 63u - msb(below) - msb(above);       // clz + clz + 1: slots of the leaf, 1, 2, 4 or 8
    len = MICRO_MAX - msb(slots);
    leaf = (e >> 16) + __popc(below) - 1u;
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

// ---- a thread's subsequence in registers (write kernel) ------------------------------------------
// r[0..7]: the 8 big-endian words of subsequence (c, t); r[8]: the first word of the next one.  Loaded in two halves, for a
// prefetch: the loads now (nothing waits on them), the byte swaps and the neighbour's
// first word when the subsequence is walked.  raw[0..7]: the 8 words as loaded; raw[8]: lane 31's look-ahead word
__device__ __forceinline__ void load_sub_raw(uint32_t (&raw)[9], const uint8_t *frame, unsigned long long frame_bytes,
                                             unsigned long long c, uint32_t t, uint32_t lane)
{
    const unsigned long long b = c * (CHUNK_BITS / 8) + (unsigned long long)t * (SUB_BITS / 8);
    uint4 a = make_uint4(0, 0, 0, 0), d = make_uint4(0, 0, 0, 0);
    // (two halves: one 256-bit load here makes the kernel SLOWER, 15.0 -> 16.0 ms on the 16 GiB stream)
    if (b < frame_bytes) a = __ldg(reinterpret_cast<const uint4 *>(frame + b));          // the frame is 16-byte aligned
    if (b + 16 < frame_bytes) d = __ldg(reinterpret_cast<const uint4 *>(frame + b + 16));
    raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = d.x; raw[5] = d.y; raw[6] = d.z; raw[7] = d.w;
    raw[8] = 0;
    if (lane == 31 && b + 32 < frame_bytes) raw[8] = __ldg(reinterpret_cast<const uint32_t *>(frame + b + 32));
}
__device__ __forceinline__ void finish_sub(uint32_t (&r)[9], const uint32_t (&raw)[9], uint32_t lane)
{
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = bswap32(raw[i]);
    const uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, r[0], 1);
    r[8] = lane == 31 ? bswap32(raw[8]) : nx;
}

// ---- shared-memory access by 32-bit address -------------------------------------------------------
// the 32-bit shared address of p, passed through a shuffle so that the compiler keeps it in a register instead of
// re-deriving it (S2UR SR_CgaCtaId + three uniform ops) at every use; call with the whole warp
__device__ __forceinline__ uint32_t opaque_shared_addr(const void *p)
{
    return __shfl_sync(0xFFFFFFFFu, (uint32_t)__cvta_generic_to_shared(p), 0);
}

__device__ __forceinline__ uint32_t lds32(uint32_t shared_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// Groups the cheap repair (dec_fix2_kernel) gave up on are marked in chunkE2 of their first chunk
constexpr uint32_t CHUNK_DIRTY = 0xFFFFFFFEu;           // in chunkE2 of a group's first chunk

// -------------------------------------------------------------------------------------------------
// dec_sync4_kernel: every WARP converges on one chunk by itself (no teams, no CTA or named barriers after the
// table load).  A lane owns LANE_SUBS = 16 consecutive subsequences (512 bytes of payload): four times the span
// of round 1's team kernel, so a guessed start costs a quarter as many repeated walks per payload bit (a wrong
// walk re-joins the true chain after tens of bits on skewed codes, after ~2,000 bits on the nearly fixed-length
// codes of flat data).  The lane streams its span through a private shared-memory row, two subsequences (64
// bytes, two whole sectors) at a time; the next pair is in flight in registers while this one is walked.  The
// warp walks in lock step per subsequence (lanes wait for each other every 256 bits: ~21 code words), and the
// per-subsequence records — which are also the output — are the checkpoints at which a repeated walk from a
// corrected start recognises the earlier one.
constexpr int S4_THREADS = 1024;
constexpr int S4_WARPS = S4_THREADS / 32;
constexpr uint32_t LANE_SUBS = DEC_THREADS / 32;                // 16
constexpr uint32_t LANE_BITS = LANE_SUBS * SUB_BITS;            // 4096
// a warp's rows are stored column-major — word w of lane l at (w * 32 + l) — so a lane reads and writes its own bank
// whatever word it is at (with one padded row per lane the lanes collide as soon as they drift apart inside a
// subsequence: the walk is bound by shared-memory wavefronts and ALU issue about equally); the u16 records likewise
constexpr uint32_t ROW4_WORDS = SUB_BITS / 32 + 1;             // a subsequence + the word that follows it
constexpr uint32_t REC4_WORDS = LANE_SUBS / 2;                  // 16 u16 records
constexpr uint32_t ROW4_STRIDE = 32 * 4;                        // bytes between consecutive words of a lane
constexpr uint32_t REC4_STRIDE = 32 * 2;                        // bytes between consecutive records of a lane
constexpr uint32_t S4_TAB_BYTES = 4u << MICRO_K;                // the d14 plane
// dec_sync4's rows are RINGS of two subsequences (words 0-7: the even ones, 8-15: the odd ones, 16: a copy of word 0):
// the word that follows a subsequence is the first word of the next one, which the lane has already stored, so a
// subsequence is ONE 32-byte load — a whole sector, read once — instead of two halves and a look-ahead word
constexpr uint32_t S4_ROW_WORDS = 2 * (SUB_BITS / 32) + 1;
constexpr uint32_t S4_BAR = S4_TAB_BYTES + S4_THREADS * (S4_ROW_WORDS + REC4_WORDS) * 4;    // the mbarrier of the plane load
constexpr size_t S4_SMEM = S4_BAR + 16;
static_assert(GROUP_CHUNKS == 1, "dec_sync4: a warp converges on one chunk");

__device__ __forceinline__ void sts16(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

struct Sync4Ctx {
    const uint8_t *frame;
    unsigned long long frame_bytes, F0, range_end_bit, nch;
    const DecodeTable *tab;
    DecWork *work;
    uint32_t d14_a;                     // shared address of the d14 plane
    uint32_t g, k2shift;
};

__device__ __forceinline__ uint32_t lane_limit(unsigned long long X, unsigned long long range_end_bit)
{   // bits of the lane's span at frame bit X that lie before the end of the range (0 .. LANE_BITS)
    if (X >= range_end_bit) return 0u;
    const unsigned long long room = range_end_bit - X;
    return room >= LANE_BITS ? LANE_BITS : (uint32_t)room;
}

// the 32 bytes of the frame at byte b (one sector) and the word after them, as loaded; zero past the end
__device__ __forceinline__ void load_sub_raw4(uint4 &a, uint4 &d, uint32_t &next, const uint8_t *frame,
                                              unsigned long long frame_bytes, unsigned long long b)
{
    a = make_uint4(0, 0, 0, 0); d = a; next = 0;
    if (b + 32 <= frame_bytes) ld_stream_2v4(frame + b, ((uintptr_t)frame & 31) == 0, a, d);
    else {
        if (b < frame_bytes) a = ld_stream_v4(frame + b);
        if (b + 16 < frame_bytes) d = ld_stream_v4(frame + b + 16);
    }
    if (b + 32 < frame_bytes) next = __ldg(reinterpret_cast<const uint32_t *>(frame + b + 32));
}

// the 32 bytes of the frame at byte b (one sector), as loaded; zero past the end.  a32: the frame is 32-byte aligned
// (one 256-bit load; otherwise two halves)
__device__ __forceinline__ void load_sub8(uint32_t (&r)[8], const uint8_t *frame, unsigned long long frame_bytes,
                                          unsigned long long b, bool a32)
{
    if (a32 && b + 32 <= frame_bytes) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "l"(frame + b));
    } else {
        uint4 a = make_uint4(0, 0, 0, 0), d = a;
        if (b < frame_bytes) a = ld_stream_v4(frame + b);
        if (b + 16 < frame_bytes) d = ld_stream_v4(frame + b + 16);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = d.w;
    }
}
// registers -> one half of the lane's ring (big-endian words); the even half leaves a copy of its first word behind word 15
__device__ __forceinline__ void store_half(uint32_t row_a, uint32_t odd, const uint32_t (&r)[8])
{
    const uint32_t ha = row_a + odd * (8u * ROW4_STRIDE);
    const uint32_t w0 = bswap32(r[0]);
    sts32(ha, w0);
#pragma unroll
    for (int i = 1; i < 8; i++) sts32(ha + i * ROW4_STRIDE, bswap32(r[i]));
    if (!odd) sts32(row_a + 16u * ROW4_STRIDE, w0);
}

// The lanes with `live` set walk their spans from bit `wstart` (relative to the span) up to `lim`, writing one
// record per subsequence into their record rows.  have_rec: the row describes an earlier walk of this span; the
// new walk stops at the first subsequence it enters where the earlier one did and keeps the earlier records (and
// the caller's `end`) from there on.  A walk that runs to `lim` sets end = overflow past it.
template <bool MULTI>
__device__ __forceinline__ void walk_lane(const Sync4Ctx &S, uint32_t row_a, uint32_t rec_a, unsigned long long span_bit0,
                                          bool live, bool have_rec, uint32_t wstart, uint32_t lim, uint32_t &end, uint32_t &bad)
{
    const uint32_t d14_a = S.d14_a, k2shift = S.k2shift;
    const unsigned long long span_byte0 = span_bit0 >> 3;
    uint32_t pos = wstart;
    const bool a32 = ((uintptr_t)S.frame & 31) == 0;
    uint32_t r[8];                                      // the subsequence in flight
    if (live) {
        load_sub8(r, S.frame, S.frame_bytes, span_byte0, a32);
        store_half(row_a, 0u, r);
        load_sub8(r, S.frame, S.frame_bytes, span_byte0 + 32u, a32);
    }
#pragma unroll 1
    for (uint32_t k = 0; k < LANE_SUBS; k++) {
        if (!__any_sync(0xFFFFFFFFu, live)) break;
        if (live) {
            // subsequence k + 1 has arrived (after the last one: the first word of the next span): into the other
            // half of the ring, behind subsequence k; k + 2 goes in flight
            store_half(row_a, (k + 1) & 1u, r);
            if (k + 2 < LANE_SUBS) load_sub8(r, S.frame, S.frame_bytes, span_byte0 + 32ull * (k + 2), a32);
            else if (k + 2 == LANE_SUBS) {
                const unsigned long long b = span_byte0 + 32ull * LANE_SUBS;
                r[0] = b < S.frame_bytes ? __ldg(reinterpret_cast<const uint32_t *>(S.frame + b)) : 0u;
            }
            const uint32_t sub0 = SUB_BITS * k;
            const uint32_t rel = pos - sub0;            // where this walk enters the subsequence (< 64 for codes <= 64 bits)
            if (have_rec && pos < lim) {                // entering where the earlier walk did: the walks have met
                const uint32_t old = lds16(rec_a + REC4_STRIDE * k);
                if (old >= 64u && rel == (old & 63u)) live = false;
            }
            if (live) {
                const uint32_t lw = min(lim, sub0 + SUB_BITS);
                uint32_t n = 0;
                while (pos < lw) {
                    const uint32_t wa = row_a + ((pos << 2) & (15u * ROW4_STRIDE));     // word (pos / 32) mod 16 of my ring
                    const uint32_t win = __funnelshift_l(lds32(wa + ROW4_STRIDE), lds32(wa), pos);
                    const uint32_t e14 = lds32(d14_a + ((win >> (30 - MICRO_K)) & ((4u << MICRO_K) - 4u)));
                    const uint32_t deep = (MICRO_K + 1) + ((e14 >> ((win >> (32 - MICRO_MAX - 1)) & 30u)) & 3u);  // micro tree: 2 bits per slot
                    const bool micro = (e14 & 0xFu) != 0xCu;
                    uint32_t len = micro ? deep : (e14 >> 28), cnt = 1;
                    if (MULTI) {                        // all the code words the 14 bits hold, when they end inside the subsequence
                        const uint32_t tot = (e14 >> 4) & 0xFu;
                        if (!micro && pos + tot <= lw) { len = tot; cnt = (e14 >> 8) & 0xFu; }
                    }
                    if (len == 0) {
                        len = __ldg(S.tab->lenflat + (win >> k2shift));
                        if (len == 0) {
                            const uint32_t e = slow_decode(S.tab, S.frame, S.frame_bytes, span_bit0 + pos);
                            bad |= e >> 31;
                            len = e & 0x7Fu;
                        }
                    }
                    pos += len;
                    n += cnt;
                }
                sts16(rec_a + REC4_STRIDE * k, n ? ((rel & 63u) | (n << 6)) : 0u);
            }
        }
    }
    if (live) end = pos - lim;          // ran to the limit (pos >= lim)
}

// One warp converges on chunk c.  xstart: frame bit at which a code word is KNOWN to start (the stream head, or the
// true overflow of the chunk before when a chunk is redone), NO_START when there is none.  Spans that end before it
// hold no code word; the lane whose span holds it starts there; every other lane 0 starts from a guess like the lanes
// behind it, and dec_fix2_kernel repairs the chunk's first subsequences afterwards.  dense: code words this warp
// counted in its last chunk (see DENSE4_MIN).  mark: reset the chunk's repair mark (the regroup kernel must not).
// Short code words come several to a 14-bit look-up; when the warp's last chunk held more than DENSE4_MIN of them (under
// ~10 bits each) the walks take all the code words an entry holds in one step (MULTI), which costs every step a few
// instructions more.
constexpr uint32_t DENSE4_MIN = CHUNK_BITS / 10u;
template <bool MULTI>
__device__ __forceinline__ void sync_chunk(const Sync4Ctx &S, uint32_t row_a, uint32_t rec_a, unsigned long long c,
                                           unsigned long long xstart, bool mark, uint32_t &bad, uint32_t &dense)
{
    const uint32_t lane = threadIdx.x & 31;
    DecLayout L(S.work, S.nch);
    const unsigned long long X = c * CHUNK_BITS + (unsigned long long)lane * LANE_BITS;
    const bool have_x = xstart != NO_START;
    const bool before = have_x && X + LANE_BITS <= xstart;          // my span ends before the first code word
    const bool holds = have_x && xstart >= X && xstart < X + LANE_BITS;
    const uint32_t lim = before ? 0u : lane_limit(X, S.range_end_bit);
    const bool fixed = lane == 0 || holds;              // no predecessor in the warp (guess), or the known start
    uint32_t p = holds ? (uint32_t)(xstart - X) : (X >= S.F0 ? spec_start(X, S.F0, S.g) : 0u);
    // `end`: overflow of the walk from p.  The record row describes the walk from rec_p, which ended at rec_end.
    // memo: up to four (start + 1, end) pairs of walks of this span.  Data that does not re-synchronise (a long run
    // of one code word is periodic) makes the fix-point hand a lane the same few starts again and again; a remembered
    // start costs no walk, and the records of the final start are rebuilt once at the end.
    uint32_t end = 0, rec_p = p, rec_end = 0, mslot = 0, nwalk = 0;
    unsigned long long memo = 0;
    bool have_rec = false;
    const bool movable = !fixed && lim != 0;
    // first pass: everybody walks
    bool want = true, final = false;
    uint32_t wstart = p;
    for (;;) {
        bool live = want && wstart < lim;
        if (want && !live) {                            // starts at or past my limit: no code word of mine
#pragma unroll
            for (uint32_t k = 0; k < LANE_SUBS; k++) sts16(rec_a + REC4_STRIDE * k, 0u);
            end = lim ? wstart - lim : 0u;
            rec_p = wstart; rec_end = end; have_rec = false;
        }
        if (live) end = rec_end;                        // a walk that merges keeps the recorded walk's end
        walk_lane<MULTI>(S, row_a, rec_a, X, live, have_rec, wstart, lim, end, bad);
        if (live) {
            rec_p = wstart; rec_end = end; have_rec = true;
            memo = (memo & ~(0xFFFFull << (16 * mslot))) | ((unsigned long long)(((wstart + 1) << 8) | end) << (16 * mslot));
            mslot = (mslot + 1) & 3;
            nwalk++;
        }
        if (final) break;
        // fix-point: my true start is my predecessor's overflow
        for (;;) {
            uint32_t q = __shfl_up_sync(0xFFFFFFFFu, end, 1);
            want = movable && q != p;
            uint32_t hit = 0;                           // ((q + 1) << 8) | end of a walk from q done before
            if (want && nwalk >= 2) {                   // the first correction of a guess cannot be a repeat
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t en = (uint32_t)(memo >> (16 * i)) & 0xFFFFu;
                    if ((en >> 8) == q + 1) hit = en;
                }
            }
            if (want) { p = q; wstart = q; }
            if (hit) { end = hit & 0xFFu; want = false; }       // the record row stays with the walk it describes
            const bool any_want = __any_sync(0xFFFFFFFFu, want);
            const bool any_hit = __any_sync(0xFFFFFFFFu, hit != 0);
            if (any_want) break;                        // somebody walks
            if (!any_hit) { final = true; break; }      // nothing moved: converged
        }
        if (final) {                                    // the final start was a remembered one: rebuild its records
            want = lim != 0 && rec_p != p;
            wstart = p;
            if (!__any_sync(0xFFFFFFFFu, want)) break;
        }
    }

    // ---- records out: 16 per lane, 32 contiguous bytes; chunk total and overflow ----
    uint32_t r8[8], total = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r8[i] = lds16(rec_a + REC4_STRIDE * (2 * i)) | (lds16(rec_a + REC4_STRIDE * (2 * i + 1)) << 16);
        total += ((r8[i] & 0xFFFFu) >> 6) + (r8[i] >> 22);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(L.info + c * DEC_THREADS + lane * LANE_SUBS);
    dst[0] = make_uint4(r8[0], r8[1], r8[2], r8[3]);
    dst[1] = make_uint4(r8[4], r8[5], r8[6], r8[7]);
    total = __reduce_add_sync(0xFFFFFFFFu, total);
    dense = total;
    if (lane == 0) {
        L.chunkCnt[c] = total;
        if (mark) L.chunkE2[c] = 0xFFFFFFFFu;
    }
    if (lane == 31) L.chunkE[c] = end;
    // the lane whose span holds the end of the range reports the overflow past it
    if (lim && lane_limit(X + LANE_BITS, S.range_end_bit) == 0) S.work->result[1] = end;
}

__device__ __forceinline__ void sync4_setup(uint32_t *smem, const DecodeTable *tab, uint32_t &d14_a, uint32_t &row_a, uint32_t &rec_a)
{
    const uint32_t tid = threadIdx.x;
    d14_a = opaque_shared_addr(smem);
    cta_bulk_load(d14_a, tab->d14, S4_TAB_BYTES, d14_a + S4_BAR);          // lengths only; one bulk copy, no per-thread staging
    const uint32_t wid = tid >> 5, lane = tid & 31;
    row_a = d14_a + S4_TAB_BYTES + wid * (S4_ROW_WORDS * ROW4_STRIDE) + lane * 4u;
    rec_a = d14_a + S4_TAB_BYTES + S4_WARPS * (S4_ROW_WORDS * ROW4_STRIDE) + wid * (LANE_SUBS * REC4_STRIDE) + lane * 2u;
}

__global__ void __launch_bounds__(S4_THREADS, 1)
dec_sync4_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes,
                 unsigned long long range_end_bit, const DecodeTable *__restrict__ tab, DecWork *work,
                 unsigned long long nch, unsigned long long c_first, unsigned long long c_last, uint32_t speculative)
{
    extern __shared__ __align__(16) uint32_t s4_smem[];
    if (tab->single_sym) return;                        // empty payload, see dec_fill_kernel
    uint32_t d14_a, row_a, rec_a;
    sync4_setup(s4_smem, tab, d14_a, row_a, rec_a);
    // speculative: the start of the range is not known (yet): every span starts from a guess, code word boundaries
    // are not assumed to lie on a lattice
    const unsigned long long F0 = speculative ? 0ull : work->start[0];
    const Sync4Ctx S{frame, frame_bytes, F0, range_end_bit, nch, tab, work, d14_a, speculative ? 1u : tab->len_gcd, 32u - tab->k2};
    __syncthreads();                                    // plane loaded; the warps are on their own from here
    const uint32_t wid = threadIdx.x >> 5;
    uint32_t bad = 0, dense = 0;
    for (unsigned long long c = c_first + (unsigned long long)blockIdx.x * S4_WARPS + wid; c < c_last;
         c += (unsigned long long)gridDim.x * S4_WARPS) {
        const unsigned long long xs = speculative ? NO_START : F0;
        if (dense > DENSE4_MIN) sync_chunk<true>(S, row_a, rec_a, c, xs, true, bad, dense);
        else sync_chunk<false>(S, row_a, rec_a, c, xs, true, bad, dense);
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// Chunks dec_fix2_kernel gave up on (marked CHUNK_DIRTY): the chain of the chunk before never meets the chain the
// chunk recorded from its guessed start (data that does not re-synchronise, e.g. a long run of one code word).  The
// marks are written before this kernel starts and never changed by it, so runs of marked chunks are disjoint and each
// belongs to exactly one warp: the warp that finds a run's head redoes its chunks one after the other from their TRUE
// starts, and walks on through unmarked chunks for as long as the overflow it hands over is not the start they
// recorded — up to, not into, the next run's head.  All decisions are taken by lane 0 and broadcast.  What this leaves
// open (a chain that reaches another run) is found by dec_verify_kernel and settled by the serial kernel.
__global__ void __launch_bounds__(S4_THREADS, 1)
dec_regroup4_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes,
                    unsigned long long range_end_bit, const DecodeTable *__restrict__ tab, DecWork *work,
                    unsigned long long nch, unsigned long long c_first, unsigned long long c_last, uint32_t speculative)
{
    if (work->flags[3] == 0) return;                    // no chunk was given up on
    extern __shared__ __align__(16) uint32_t s4_smem[];
    if (tab->single_sym) return;
    uint32_t d14_a, row_a, rec_a;
    sync4_setup(s4_smem, tab, d14_a, row_a, rec_a);
    const unsigned long long F0 = speculative ? 0ull : work->start[0];
    const Sync4Ctx S{frame, frame_bytes, F0, range_end_bit, nch, tab, work, d14_a, speculative ? 1u : tab->len_gcd, 32u - tab->k2};
    __syncthreads();
    DecLayout L(work, nch);
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long c_lo = c_first ? c_first : 1;      // the first chunk of a stream has an exact start
    uint32_t bad = 0, dense = 0;
    for (unsigned long long c = c_lo + (unsigned long long)blockIdx.x * S4_WARPS + wid; c < c_last;
         c += (unsigned long long)gridDim.x * S4_WARPS) {
        uint32_t head = 0, s = 0;
        if (lane == 0) {
            head = L.chunkE2[c] == CHUNK_DIRTY && !(c > c_lo && L.chunkE2[c - 1] == CHUNK_DIRTY);
            s = L.chunkE[c - 1];
        }
        head = __shfl_sync(0xFFFFFFFFu, head, 0);
        if (!head) continue;
        s = __shfl_sync(0xFFFFFFFFu, s, 0);
        bool in_run = true;
        for (unsigned long long cur = c;;) {
            sync_chunk<false>(S, row_a, rec_a, cur, cur * CHUNK_BITS + s, false, bad, dense);
            const unsigned long long next = cur + 1;
            if (next >= c_last) break;
            uint32_t dirty = 0, first = 0, e = 0;
            __syncwarp();
            if (lane == 0) {
                dirty = L.chunkE2[next] == CHUNK_DIRTY;
                first = L.info[next * DEC_THREADS];
                e = L.chunkE[cur];                      // written by lane 31 before the __syncwarp
            }
            dirty = __shfl_sync(0xFFFFFFFFu, dirty, 0);
            first = __shfl_sync(0xFFFFFFFFu, first, 0);
            e = __shfl_sync(0xFFFFFFFFu, e, 0);
            if (dirty) {
                if (!in_run) break;                     // another run's head: that run's warp takes it from there
            } else {
                in_run = false;
                if (e == (first & 63u)) break;          // the chain meets what the next chunk recorded
            }
            s = e;
            cur = next;
        }
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// the chunk that holds the first code word: boundaries up to and including its own have nothing to repair
__device__ __forceinline__ unsigned long long first_chunk(const DecWork *work, uint32_t speculative)
{
    return speculative ? 0ull : work->start[0] / CHUNK_BITS;
}

// every group must start where the group before it ends; what the parallel repairs left open goes to the serial kernel
__global__ void dec_verify_kernel(const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch,
                                  unsigned long long g_first, unsigned long long g_last, uint32_t speculative)
{
    const unsigned long long g = (g_first ? g_first : 1) + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_last || tab->single_sym || g <= first_chunk(work, speculative)) return;
    DecLayout L(work, nch);
    const unsigned long long c = g * GROUP_CHUNKS;
    if (c >= nch) return;
    if (L.chunkE[c - 1] != (uint32_t)(L.info[c * DEC_THREADS] & 63u)) atomicExch(&work->flags[0], 1ull);
}



// -------------------------------------------------------------------------------------------------
// Every WARP works on its own: a unit of 32 consecutive subsequences (1 KiB of payload), whose output offset
// it derives itself from the chunk's records (no CTA-wide scan, no CTA barrier after the planes are loaded).
// The symbols of a unit are compacted in the warp's staging window and leave with aligned 128-bit stores.
__global__ void __launch_bounds__(W3_THREADS, 1)
dec_write3_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes,
                  const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch,
                  unsigned long long c0, unsigned long long c1,
                  uint16_t *__restrict__ out, uint32_t check, uint32_t upw, uint32_t dense_min)
{
    extern __shared__ __align__(16) uint32_t w3_smem[];
    constexpr uint32_t WIN = W3_WIN;
    uint32_t *s_t14 = w3_smem;                                                  // 2^MICRO_K
    uint16_t *s_leaves = reinterpret_cast<uint16_t *>(s_t14 + (1u << MICRO_K)); // NSYM
    if (tab->single_sym) return;
    const unsigned long long F0 = work->start[0], n_symbols = work->start[1];
    const unsigned long long head_sub = F0 / SUB_BITS;         // the subsequence that holds the first code word
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint16_t *sout = s_leaves + NSYM + wid * (WIN + 8);                         // this warp's window
    // the planes and the window by 32-bit shared address: a compare, an add and the address of a look-up are one
    // instruction each instead of the two or three of a generic 64-bit pointer
    const uint32_t t14_a = opaque_shared_addr(w3_smem), lv_a = t14_a + (4u << MICRO_K);
    const uint32_t sout_a = lv_a + NSYM * 2u + wid * ((WIN + 8) * 2u);
    if (dense_min != 0xFFFFFFFFu) {     // every chunk of this CTA is dec_write4_kernel's: leave before the planes load
        const unsigned long long upc = (DEC_THREADS / 32) / upw;                // runs per chunk
        bool any = false;
        for (unsigned long long run = (unsigned long long)blockIdx.x * W3_WARPS + wid + (unsigned long long)lane * gridDim.x * W3_WARPS;
             run < (c1 - c0) * upc && !any; run += 32ull * gridDim.x * W3_WARPS)
            any = L.chunkCnt[c0 + run / upc] < dense_min;
        if (!__syncthreads_or(any)) return;
    }
    {   // t14 and leaves are neighbours in the table and in shared memory: one bulk copy (192 KiB)
        const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(w3_smem);
        cta_bulk_load(a0, tab->t14, (4u << MICRO_K) + NSYM * 2, a0 + W3_BAR);
    }
    const uint32_t k2shift = 32u - tab->k2;
    uint32_t bad = 0;
    constexpr uint32_t UPC = DEC_THREADS / 32;          // units per chunk
    // A warp takes a RUN of upw consecutive units of one chunk (upw = 16, a whole chunk, on large streams; fewer on
    // small ones so that every SM has work): the output offset is derived once per run and carried from unit to
    // unit, and the next unit's record and payload words are loaded while the current one is walked.
    const unsigned long long nruns = (c1 - c0) * (UPC / upw);
    for (unsigned long long run = (unsigned long long)blockIdx.x * W3_WARPS + wid; run < nruns;
         run += (unsigned long long)gridDim.x * W3_WARPS) {
        const unsigned long long ug0 = c0 * UPC + run * upw;
        const unsigned long long c = ug0 / UPC;
        const uint32_t u0 = (uint32_t)(ug0 % UPC);
        const unsigned long long cbase = L.chunkBase[c];
        if (cbase >= n_symbols || (c + 1) * DEC_THREADS <= head_sub) continue;     // nothing left to write / before the first code word
        if (L.chunkCnt[c] >= dense_min) continue;       // a chunk of short code words: dec_write4_kernel's
        uint32_t ninf = L.info[c * DEC_THREADS + 32 * u0 + lane];
        uint32_t nr[9];
        load_sub_raw(nr, frame, frame_bytes, c, 32 * u0 + lane, lane);
        // symbols of the chunk before my run: lane l sums the 16 records [16 l, 16 l + 16) of the chunk
        unsigned long long base = cbase;
        if (u0) {
            const uint4 *ip = reinterpret_cast<const uint4 *>(L.info + c * DEC_THREADS);
            const uint4 a = ip[2 * lane], d = ip[2 * lane + 1];
            const uint32_t w8[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) sum += ((w8[i] & 0xFFFFu) >> 6) + (w8[i] >> 22);
            uint32_t x = (lane < 2 * u0) ? sum : 0u;
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
            base += x;
        }
        for (uint32_t u = u0; u < u0 + upw; u++) {
        const uint32_t t = 32 * u + lane;               // my subsequence of the chunk
        const uint32_t inf = ninf;
        uint32_t r[9];
        finish_sub(r, nr, lane);
        if (u + 1 < u0 + upw) {                         // the next unit of the run: in flight during this one's walk
            ninf = L.info[c * DEC_THREADS + t + 32];
            load_sub_raw(nr, frame, frame_bytes, c, t + 32, lane);
        }
        const uint32_t cnt = inf >> 6;
        uint32_t pos = (c * DEC_THREADS + t == head_sub) ? (uint32_t)(F0 % SUB_BITS) : (inf & 63u);   // the head may sit past bit 63
        uint32_t x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        const uint32_t off = x - cnt;                   // unit-relative index of my first symbol
        const uint32_t unit_total = __shfl_sync(0xFFFFFFFFu, x, 31);
        unsigned long long total = unit_total;
        if (base >= n_symbols) break;
        if (base + total > n_symbols) total = n_symbols - base;   // garbage past the payload end is dropped
        const uint32_t my_end = (uint32_t)min((unsigned long long)(off + cnt), total);
        uint32_t o = off;                               // unit-relative index of my next symbol

        const uint32_t mis = (uint32_t)(base & 7);      // staging slot j <-> output symbol base - mis + j
        for (uint32_t w0 = 0; w0 < (uint32_t)total; w0 += WIN) {
            const uint32_t wend = min((uint32_t)total, w0 + WIN);
            const uint32_t o_end = min(my_end, wend);
            if (o < o_end) {
                uint32_t sp = sout_a + 2u * (o - w0 + mis);
                const uint32_t sp_end = sout_a + 2u * (o_end - w0 + mis);
                // phases of W3_PHASE words: the window's two words are picked by selects inside a phase; longer
                // phases keep more lanes busy (a lane leaves a phase when its position passes the phase's end)
#pragma unroll
                for (int w = 0; w < 8; w += W3_PHASE) {
                    const uint32_t lw = 32u * (w + W3_PHASE);
                    while (pos < lw && sp < sp_end) {
                        uint32_t hi = r[w], lo = r[w + 1];
#pragma unroll
                        for (int k = 1; k < W3_PHASE; k++)
                            if (pos >= 32u * (w + k)) { hi = r[w + k]; lo = r[w + k + 1]; }
                        const uint32_t win = __funnelshift_l(lo, hi, pos);
                        const uint32_t e14 = lds32(t14_a + ((win >> (30 - MICRO_K)) & ((4u << MICRO_K) - 4u)));
                        uint32_t len, sym;
                        if (e14 & MICRO_FLAG) {
                            uint32_t leaf;
                            micro_decode(e14, win, len, leaf);
                            sym = lds16(lv_a + 2u * leaf);
                        } else {
                            len = (e14 >> 1) & 0x7Fu;
                            sym = e14 >> 16;
                            if (len == 0) {
                                uint32_t e = __ldg(tab->flat2 + (win >> k2shift));
                                if (e == 0) {
                                    e = slow_decode(tab, frame, frame_bytes, c * CHUNK_BITS + t * SUB_BITS + pos);
                                    bad |= e >> 31;
                                }
                                len = e & 0x7Fu;
                                sym = (e >> 8) & 0xFFFFu;
                            }
                        }
                        sts16(sp, sym);
                        sp += 2;
                        pos += len;
                    }
                }
                o = o_end;
            }
            __syncwarp();
            // flush [w0, wend): staging slots [mis, mis + n)
            const uint32_t n = wend - w0;
            uint16_t *dst = out + base + w0 - mis;      // 16-byte aligned when out is
            const uint32_t endslot = mis + n, nvec = (endslot + 7) / 8;
            if (((uintptr_t)dst & 15) == 0) {
                for (uint32_t q = lane; q < nvec; q += 32) {
                    const uint32_t j0 = q * 8;
                    if (j0 >= mis && j0 + 8 <= endslot) st_stream_v4(dst + j0, reinterpret_cast<const uint4 *>(sout)[q]);
                }
                // the partial vectors at the two ends, a symbol per lane: lanes 0-7 the first vector when it is
                // partial, lanes 8-15 the last one when it is partial and not the first
                const uint32_t last0 = endslot & ~7u;
                const uint32_t j = lane < 8 ? lane : last0 + (lane - 8);
                const bool part = lane < 8 ? (mis != 0 || endslot < 8) : (lane < 16 && last0 != 0 && (endslot & 7u) != 0);
                if (part && j >= mis && j < endslot) dst[j] = sout[j];
            } else {
                for (uint32_t q = lane; q < nvec; q += 32)
                    for (uint32_t j = q * 8; j < q * 8 + 8; j++)
                        if (j >= mis && j < endslot) dst[j] = sout[j];
            }
            __syncwarp();
        }
        if (check) {
            // records that did not come from the synchronisation kernels (a side index): a walk of `cnt` code words
            // must end exactly where the next subsequence says its first code word starts
            uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, inf, 1);
            if (lane == 31) {
                const unsigned long long ns = c * DEC_THREADS + t + 1;
                nxt = ns < nch * DEC_THREADS ? (uint32_t)L.info[ns] : 0u;
            }
            const bool last = base + off + cnt == n_symbols;        // my last code word is the stream's last: no successor
            if (cnt && my_end == off + cnt && !last && ((nxt >> 6) == 0 || pos < SUB_BITS || pos - SUB_BITS != (nxt & 63u))) bad = 1;
        }
        base += unit_total;
        }
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// -------------------------------------------------------------------------------------------------
// dec_write4_kernel: every WARP decodes one chunk, every LANE a contiguous span of 16 subsequences (512 bytes of
// payload) into a contiguous run of output symbols.  dec_write3_kernel gives a lane ONE subsequence of a unit and
// walks it in phases of 64 bits (the subsequence sits in registers, which need static indices): lanes wait for each
// other at every phase end, and a unit of short code words does not fit the warp's staging window, so only a third of
// the lanes work at a time.  Here
//   * the lane streams its span through a private shared-memory row (9 words, column-major: a lane stays in its own
//     bank), so the window of a code word is two loads by a dynamic index and the lanes wait for each other once per
//     subsequence (256 bits), not four times;
//   * the symbols go to a private 16-entry ring (column-major u16) and leave, eight at a time, as ONE aligned 16-byte
//     global store per lane: no staging window, no window passes, no per-unit scan.  The lanes run in lock step, so
//     after a block of eight steps every active lane has eight symbols pending and the stores of a warp issue together;
//     the two partial vectors at the ends of a lane's run (once per 4096 bits) leave a symbol at a time;
//   * the planes arrive by bulk copies (cp.async.bulk + mbarrier), the payload by 32-byte loads one subsequence ahead.
// Nothing is shared between lanes after the offset scan: no barrier of any kind inside the chunk loop.
// Shared memory: t14 64 KiB | rings 1 KiB per warp | rows 1152 B per warp | the mbarrier.  The leaves of the micro trees
// (code words of 15-18 bits, rare in the chunks this kernel takes) are read through L1 instead: without their 128 KiB
// there is room for 32 warps, and this kernel is bound by latency (16 warps beside the leaves: 2.08 ms on 4 GiB of the
// mixed stream; leaves through L1: 16 warps 2.34, 24 warps 1.94, 32 warps 1.82).  A second plane with the SECOND code
// word a window holds (two symbols per look-up) was built and measured: 1.82 -> 1.82 / 1.85 (the lanes of a warp leave a
// subsequence together, and the extra look-up and the shorter blocks between flush tests cost what the pairs save).
#ifndef W4_WARPS
#define W4_WARPS 32
#endif
constexpr int W4_THREADS = W4_WARPS * 32;
constexpr uint32_t W4_RINGS = 4u << MICRO_K;                    // behind the t14 plane; offsets inside the dynamic shared memory
constexpr uint32_t W4_ROWS = W4_RINGS + W4_WARPS * 1024u;
constexpr uint32_t W4_BAR = W4_ROWS + W4_WARPS * (ROW4_WORDS * ROW4_STRIDE);
constexpr size_t W4_SMEM = W4_BAR + 16;
static_assert(W4_RINGS % 1024 == 0, "the ring address is formed by OR");
static_assert(W4_SMEM <= 232448, "dec_write4: shared memory");
constexpr uint32_t RING_STRIDE = 64;                            // bytes between consecutive symbols of a lane's ring
constexpr uint32_t RING_MASK = 15u * RING_STRIDE;

__global__ void __launch_bounds__(W4_THREADS, 1)
dec_write4_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes,
                  const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch,
                  unsigned long long c0, unsigned long long c1, uint16_t *__restrict__ out, uint32_t dense_min)
{
    extern __shared__ __align__(1024) uint8_t w4_smem[];
    if (tab->single_sym) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t a0 = opaque_shared_addr(w4_smem);
    if (a0 & 1023u) __trap();                           // the ring address below is base | offset
    {   // nothing of mine among this CTA's chunks (the usual case on data of long code words): leave before the planes load
        DecLayout L0(work, nch);
        bool any = false;
        for (unsigned long long c = c0 + (unsigned long long)blockIdx.x * W4_WARPS + wid + (unsigned long long)lane * gridDim.x * W4_WARPS;
             c < c1 && !any; c += 32ull * gridDim.x * W4_WARPS)
            any = L0.chunkCnt[c] >= dense_min;
        if (!__syncthreads_or(any)) return;
    }
    const unsigned long long F0 = work->start[0], n_symbols = work->start[1];
    const unsigned long long head_sub = F0 / SUB_BITS;  // the subsequence that holds the first code word
    const uint32_t head_pos = (uint32_t)(F0 % SUB_BITS);
    DecLayout L(work, nch);
    const uint32_t t14_a = a0;
    const uint32_t ring_a = a0 + W4_RINGS + wid * 1024u + lane * 2u;
    const uint32_t row_a = a0 + W4_ROWS + wid * (ROW4_WORDS * ROW4_STRIDE) + lane * 4u;
    const uint32_t k2shift = 32u - tab->k2;
    cta_bulk_load(a0, tab->t14, 4u << MICRO_K, a0 + W4_BAR);       // t14 only (see above); the warps are on their own from here
    uint32_t bad = 0;
    for (unsigned long long c = c0 + (unsigned long long)blockIdx.x * W4_WARPS + wid; c < c1;
         c += (unsigned long long)gridDim.x * W4_WARPS) {
        const unsigned long long cbase = L.chunkBase[c];
        if (cbase >= n_symbols || (c + 1) * DEC_THREADS <= head_sub) continue;     // nothing left to write / before the first code word
        if (L.chunkCnt[c] < dense_min) continue;        // a chunk of long code words: dec_write3_kernel's
        // my 16 records, my symbols, where they go
        uint32_t r8[8];
        {
            const uint4 *ip = reinterpret_cast<const uint4 *>(L.info + c * DEC_THREADS);
            const uint4 a = ip[2 * lane], d = ip[2 * lane + 1];
            r8[0] = a.x; r8[1] = a.y; r8[2] = a.z; r8[3] = a.w; r8[4] = d.x; r8[5] = d.y; r8[6] = d.z; r8[7] = d.w;
        }
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) tot += ((r8[i] & 0xFFFFu) >> 6) + (r8[i] >> 22);
        uint32_t x = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        const unsigned long long o0 = cbase + (x - tot);            // index of my first symbol
        const uint32_t mine = o0 >= n_symbols ? 0u : (uint32_t)min((unsigned long long)tot, n_symbols - o0);   // garbage past the end is dropped
        uint16_t *gp = out + o0;
        const uint32_t i0 = (uint32_t)((uintptr_t)gp >> 1) & 7u;    // ring slot s <-> gp[s], gp 16-byte aligned
        gp -= i0;
        // j: ring position of my next symbol, in units of RING_STRIDE (slot = j / 64 mod 16); fl: first slot not stored
        uint32_t j = i0 * RING_STRIDE, fl = 0;
        const uint32_t jtot = (i0 + mine) * RING_STRIDE;
        const unsigned long long sub0 = c * DEC_THREADS + (unsigned long long)lane * LANE_SUBS;
        const uint32_t head_k = (head_sub >= sub0 && head_sub < sub0 + LANE_SUBS) ? (uint32_t)(head_sub - sub0) : 0xFFFFu;
        const unsigned long long span_byte0 = sub0 * (SUB_BITS / 8);
        const bool live = mine != 0;
        uint4 na, nd;
        uint32_t nnext = 0;
        if (live) load_sub_raw4(na, nd, nnext, frame, frame_bytes, span_byte0);
#pragma unroll 1
        for (uint32_t k = 0; k < LANE_SUBS; k++) {
            const uint32_t rec = r8[0] & 0xFFFFu;
#pragma unroll
            for (int i = 0; i < 7; i++) r8[i] = __funnelshift_r(r8[i], r8[i + 1], 16);
            r8[7] >>= 16;
            if (!__any_sync(0xFFFFFFFFu, j < jtot)) break;
            if (live) {
                sts32(row_a + 0 * ROW4_STRIDE, bswap32(na.x)); sts32(row_a + 1 * ROW4_STRIDE, bswap32(na.y));
                sts32(row_a + 2 * ROW4_STRIDE, bswap32(na.z)); sts32(row_a + 3 * ROW4_STRIDE, bswap32(na.w));
                sts32(row_a + 4 * ROW4_STRIDE, bswap32(nd.x)); sts32(row_a + 5 * ROW4_STRIDE, bswap32(nd.y));
                sts32(row_a + 6 * ROW4_STRIDE, bswap32(nd.z)); sts32(row_a + 7 * ROW4_STRIDE, bswap32(nd.w));
                sts32(row_a + 8 * ROW4_STRIDE, bswap32(nnext));
                if (k + 1 < LANE_SUBS) load_sub_raw4(na, nd, nnext, frame, frame_bytes, span_byte0 + 32ull * (k + 1));
            }
            const uint32_t cnt = rec >> 6;
            uint32_t pos = k == head_k ? head_pos : (rec & 63u);
            const uint32_t jend = min(j + cnt * RING_STRIDE, jtot);
            while (j < jend) {
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    if (j < jend) {
                        const uint32_t wa = row_a + ((pos << 2) & (7u * ROW4_STRIDE));     // word (pos / 32) mod 8 of my row
                        const uint32_t win = __funnelshift_l(lds32(wa + ROW4_STRIDE), lds32(wa), pos);
                        const uint32_t e14 = lds32(t14_a + ((win >> (30 - MICRO_K)) & ((4u << MICRO_K) - 4u)));
                        uint32_t len, sym;
                        if (e14 & MICRO_FLAG) {
                            uint32_t leaf;
                            micro_decode(e14, win, len, leaf);
                            sym = __ldg(tab->leaves + leaf);       // rare in these chunks: through L1, not in shared memory
                        } else {
                            len = (e14 >> 1) & 0x7Fu;
                            sym = e14 >> 16;
                            if (len == 0) {
                                uint32_t e = __ldg(tab->flat2 + (win >> k2shift));
                                if (e == 0) {
                                    e = slow_decode(tab, frame, frame_bytes, (sub0 + k) * SUB_BITS + pos);
                                    bad |= e >> 31;
                                }
                                len = e & 0x7Fu;
                                sym = (e >> 8) & 0xFFFFu;
                            }
                        }
                        sts16(ring_a | (j & RING_MASK), sym);
                        j += RING_STRIDE;
                        pos += len;
                    }
                }
                if (j - fl >= 8 * RING_STRIDE) {
                    // eight symbols from ring slots (fl / 64 mod 16) ..: one aligned vector (the first one of my run may
                    // begin before my first symbol: a symbol at a time then)
                    const uint32_t ra = ring_a | (fl & (8u * RING_STRIDE));
                    uint32_t s8[8];
#pragma unroll
                    for (int t = 0; t < 8; t++) s8[t] = lds16(ra + t * RING_STRIDE);
                    uint16_t *dst = gp + (fl / RING_STRIDE);
                    if (fl == 0 && i0 != 0) {
#pragma unroll
                        for (int t = 1; t < 8; t++) if ((uint32_t)t >= i0) dst[t] = (uint16_t)s8[t];
                    } else {
#ifdef W4_EXP_NOSTORE
                        if (s8[0] == 0x12345 && s8[7] == 0x54321)
#endif
                        st_stream_v4(dst, make_uint4(s8[0] | (s8[1] << 16), s8[2] | (s8[3] << 16), s8[4] | (s8[5] << 16), s8[6] | (s8[7] << 16)));
                    }
                    fl += 8 * RING_STRIDE;
                }
            }
        }
        // what is left of my run: fewer than eight symbols, one at a time
        {
            const uint32_t ra = ring_a | (fl & (8u * RING_STRIDE));
            const uint32_t s_lo = fl / RING_STRIDE, s_hi = j / RING_STRIDE;
#pragma unroll
            for (uint32_t t = 0; t < 8; t++) {
                const uint32_t s = s_lo + t;
                if (s >= i0 && s < s_hi) gp[s] = (uint16_t)lds16(ra + t * RING_STRIDE);
            }
        }
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// chunk totals of records that came from a side index; a count no subsequence can have marks the index invalid
__global__ void idx_chunks_kernel(DecWork *work, unsigned long long nch)
{
    const unsigned long long c = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= nch) return;
    const uint32_t lane = threadIdx.x & 31;
    DecLayout L(work, nch);
    const uint4 *ip = reinterpret_cast<const uint4 *>(L.info + c * DEC_THREADS);
    const uint4 a = ip[2 * lane], d = ip[2 * lane + 1];
    const uint32_t w8[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
    uint32_t sum = 0, worst = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t lo = (w8[i] & 0xFFFFu) >> 6, hi = w8[i] >> 22;
        sum += lo + hi;
        worst = max(worst, max(lo, hi));
    }
    sum = __reduce_add_sync(0xFFFFFFFFu, sum);
    worst = __reduce_max_sync(0xFFFFFFFFu, worst);
    if (lane == 0) {
        L.chunkCnt[c] = sum;
        L.chunkE[c] = 0;
        L.chunkE2[c] = 0xFFFFFFFFu;
        if (worst > SUB_BITS) atomicExch(&work->flags[1], 1ull);
    }
}

// the records of a side index must account for every symbol of the stream, no more, no fewer
__global__ void idx_total_kernel(DecWork *work)
{
    if (work->result[2] != work->start[1]) atomicExch(&work->flags[1], 1ull);
}

int launch_idx_total(Ctx *c, DecWork *work)
{
    HF_PROF(c, "idx_total_kernel"); idx_total_kernel<<<1, 1, 0, c->stream>>>(work);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_idx_chunks(Ctx *c, DecWork *work, unsigned long long nch)
{
    HF_PROF(c, "idx_chunks_kernel"); idx_chunks_kernel<<<(unsigned)((nch * 32 + 255) / 256), 256, 0, c->stream>>>(work, nch);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// -------------------------------------------------------------------------------------------------
// Inter-group repair.  A group converges on a GUESSED start of its first span; its true start is the overflow
// of the group before.  One thread per chunk boundary walks from the true start, subsequence by subsequence
// (registers + the t14 plane read through L1), until it lands on a start the synchronisation kernel recorded.
__device__ __forceinline__ void load_sub_single(uint32_t (&r)[9], const uint8_t *frame, unsigned long long frame_bytes,
                                                unsigned long long c, uint32_t t)
{
    const unsigned long long b = c * (CHUNK_BITS / 8) + (unsigned long long)t * (SUB_BITS / 8);
    uint4 a = make_uint4(0, 0, 0, 0), d = make_uint4(0, 0, 0, 0);
    if (b < frame_bytes) a = __ldg(reinterpret_cast<const uint4 *>(frame + b));
    if (b + 16 < frame_bytes) d = __ldg(reinterpret_cast<const uint4 *>(frame + b + 16));
    r[0] = bswap32(a.x); r[1] = bswap32(a.y); r[2] = bswap32(a.z); r[3] = bswap32(a.w);
    r[4] = bswap32(d.x); r[5] = bswap32(d.y); r[6] = bswap32(d.z); r[7] = bswap32(d.w);
    r[8] = 0;
    if (b + 32 < frame_bytes) r[8] = bswap32(__ldg(reinterpret_cast<const uint32_t *>(frame + b + 32)));
}

// code words of subsequence (c, t) starting in [q, lim): their number and the overflow of the last one
__device__ __forceinline__ void sub_count(const DecodeTable *tab, const uint8_t *frame, unsigned long long frame_bytes,
                                          unsigned long long c, uint32_t t, uint32_t q, uint32_t lim, uint32_t k2shift,
                                          uint32_t &end, uint32_t &cnt, uint32_t &bad)
{
    cnt = 0;
    if (q >= lim) { end = q - lim; return; }
    uint32_t r[9];
    load_sub_single(r, frame, frame_bytes, c, t);
    uint32_t pos = q, n = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t lw = min(lim, 32u * (w + 1));
        while (pos < lw) {
            const uint32_t win = __funnelshift_l(r[w + 1], r[w], pos);
            const uint32_t e14 = __ldg(tab->t14 + (win >> (32 - MICRO_K)));
            uint32_t len;
            if (e14 & MICRO_FLAG) {
                uint32_t leaf;
                micro_decode(e14, win, len, leaf);
            } else {
                len = (e14 >> 1) & 0x7Fu;
                if (len == 0) {
                    len = __ldg(tab->lenflat + (win >> k2shift));
                    if (len == 0) {
                        const uint32_t e = slow_decode(tab, frame, frame_bytes, c * CHUNK_BITS + t * SUB_BITS + pos);
                        bad |= e >> 31;
                        len = e & 0x7Fu;
                    }
                }
            }
            pos += len;
            n++;
        }
    }
    cnt = n;
    end = pos - lim;
}

// repairs chunk c from the true start `s` (offset inside the chunk's subsequence 0): walks subsequence by
// subsequence, rewriting the records, until the walk lands on a start the synchronisation kernel recorded.
// Returns true when that happened before the chunk ended; otherwise chunkE[c] is the chunk's new overflow.
__device__ bool fix_chunk2(const DecodeTable *tab, const uint8_t *frame, unsigned long long frame_bytes,
                           unsigned long long range_end_bit, DecWork *work, DecLayout &L, unsigned long long c, uint32_t s,
                           uint32_t &bad, uint32_t t0 = 0)
{
    uint16_t *info = L.info + c * DEC_THREADS;
    const uint32_t k2shift = 32u - tab->k2;
    uint32_t q = s;
    long long delta = 0;
    for (uint32_t t = t0; t < DEC_THREADS; t++) {
        const uint32_t lim = sub_limit(c, t, range_end_bit);
        if (lim == 0) break;
        uint32_t end, cnt;
        sub_count(tab, frame, frame_bytes, c, t, q, lim, k2shift, end, cnt, bad);
        const uint32_t old = info[t];
        delta += (long long)cnt - (long long)(old >> 6);
        info[t] = (uint16_t)((q & 63u) | (cnt << 6));
        if (sub_limit(c, t + 1, range_end_bit) == 0) {          // the range ends in this subsequence
            work->result[1] = end;
            break;
        }
        if (t + 1 == DEC_THREADS) {
            L.chunkCnt[c] = (uint32_t)((long long)L.chunkCnt[c] + delta);
            if (end != L.chunkE[c]) { L.chunkE[c] = end; return false; }
            return true;
        }
        if ((uint32_t)(info[t + 1] & 63u) == end) break;
        q = end;
    }
    L.chunkCnt[c] = (uint32_t)((long long)L.chunkCnt[c] + delta);
    return true;
}

// does a walk of chunk c from `s` land on a recorded start within FIX_PROBE subsequences?  (nothing is written)
constexpr uint32_t FIX_PROBE = 16;
__device__ bool probe_chunk(const DecodeTable *tab, const uint8_t *frame, unsigned long long frame_bytes,
                            unsigned long long range_end_bit, DecLayout &L, unsigned long long c, uint32_t s, uint32_t &bad)
{
    const uint16_t *info = L.info + c * DEC_THREADS;
    const uint32_t k2shift = 32u - tab->k2;
    uint32_t q = s;
    for (uint32_t t = 0; t < FIX_PROBE; t++) {
        const uint32_t lim = sub_limit(c, t, range_end_bit);
        if (lim == 0) return true;
        uint32_t end, cnt;
        sub_count(tab, frame, frame_bytes, c, t, q, lim, k2shift, end, cnt, bad);
        if (sub_limit(c, t + 1, range_end_bit) == 0) return true;   // the range ends here
        if ((uint32_t)(info[t + 1] & 63u) == end) return true;
        q = end;
    }
    return false;
}

// One thread per group boundary (the chunks inside a group are consistent by construction).  Data that
// re-synchronises meets the recorded chain after a subsequence or two; a group that does not within FIX_PROBE
// subsequences is left to dec_regroup_kernel.
__global__ void dec_fix2_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes,
                                unsigned long long range_end_bit, const DecodeTable *__restrict__ tab, DecWork *work,
                                unsigned long long nch, unsigned long long g_first, unsigned long long g_last,
                                uint32_t speculative)
{
    const unsigned long long g = (g_first ? g_first : 1) + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_last || tab->single_sym || g <= first_chunk(work, speculative)) return;
    const unsigned long long c = g * GROUP_CHUNKS;
    if (c >= nch) return;
    DecLayout L(work, nch);
    const uint32_t s = L.chunkE[c - 1];
    if (s == (uint32_t)(L.info[c * DEC_THREADS] & 63u)) return;
    uint32_t bad = 0;
    if (probe_chunk(tab, frame, frame_bytes, range_end_bit, L, c, s, bad)) {
        fix_chunk2(tab, frame, frame_bytes, range_end_bit, work, L, c, s, bad);
        if (bad) atomicExch(&work->flags[1], 1ull);
    } else {
        L.chunkE2[c] = CHUNK_DIRTY;
        atomicExch(&work->flags[3], 1ull);
    }
}

// The safety net (dec_verify_kernel found a group that does not start where the one before it ends): one thread
// carries the true start forward chunk by chunk.  Correct for any stream; slow; not seen on real data.
__global__ void dec_fix2_serial_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes,
                                       unsigned long long range_end_bit, const DecodeTable *__restrict__ tab,
                                       DecWork *work, unsigned long long nch, unsigned long long c0,
                                       unsigned long long c1, uint32_t speculative)
{
    if (work->flags[0] == 0 || tab->single_sym) return;
    DecLayout L(work, nch);
    uint32_t bad = 0;
    const unsigned long long cf = first_chunk(work, speculative);
    for (unsigned long long c = max(c0 ? c0 : 1ull, cf + 1); c < c1; c++) {
        const uint32_t s = L.chunkE[c - 1];
        if (s == (uint32_t)(L.info[c * DEC_THREADS] & 63u)) continue;
        fix_chunk2(tab, frame, frame_bytes, range_end_bit, work, L, c, s, bad);
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
    work->flags[0] = 0;                                 // settled (the next slice starts clean)
}


// A range whose chunks were synchronised speculatively (every span from a guess, dec_fix2 between the chunks) learns
// where its first code word starts (work->start[0], e.g. the predecessor rank's overflow, which arrives by a
// collective): the chunks before it count nothing, the subsequences before it in its chunk neither, and one thread
// walks from it until it lands on a start the synchronisation recorded — after a subsequence or two on data that
// re-synchronises; a chain that does not is carried on chunk by chunk, and the overflow past the range end follows it.
__global__ void __launch_bounds__(256)
dec_fix_head_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes, unsigned long long range_end_bit,
                    const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch)
{
    if (tab->single_sym) return;
    DecLayout L(work, nch);
    const unsigned long long F0 = work->start[0];
    const bool none = F0 >= range_end_bit;              // no code word starts in this range: it only passes the bit on
    const unsigned long long cf = none ? nch : F0 / CHUNK_BITS;
    for (unsigned long long c = threadIdx.x; c < cf; c += blockDim.x) L.chunkCnt[c] = 0;
    if (none) {
        if (threadIdx.x == 0) { work->result[1] = F0 - range_end_bit; work->start[1] = 0; }
        return;
    }
    if (threadIdx.x != 0) return;
    const uint32_t t0 = (uint32_t)((F0 % CHUNK_BITS) / SUB_BITS);
    uint16_t *info = L.info + cf * DEC_THREADS;
    uint32_t gone = 0;
    for (uint32_t t = 0; t < t0; t++) { gone += info[t] >> 6; info[t] = 0; }
    L.chunkCnt[cf] -= gone;
    uint32_t bad = 0;
    if ((uint32_t)(info[t0] & 63u) != (uint32_t)(F0 % SUB_BITS) || (info[t0] >> 6) == 0 || F0 % SUB_BITS >= 64) {
        bool met = fix_chunk2(tab, frame, frame_bytes, range_end_bit, work, L, cf, (uint32_t)(F0 % SUB_BITS), bad, t0);
        for (unsigned long long c = cf + 1; !met && c < nch; c++)
            met = fix_chunk2(tab, frame, frame_bytes, range_end_bit, work, L, c, L.chunkE[c - 1], bad);
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

int launch_fix_head(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long range_end_bit,
                    const DecodeTable *d_tab, DecWork *work, unsigned long long nch)
{
    HF_PROF(c, "dec_fix_head_kernel"); dec_fix_head_kernel<<<1, 256, 0, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// chunks [c0, c1) (a slice of the stream, or all of it); everything before c0 is final
int launch_fix2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                unsigned long long c0, unsigned long long c1, bool speculative)
{
    if (c1 <= c0) return HF_OK;
    const unsigned long long g_first = c0 / GROUP_CHUNKS, g_last = (c1 + GROUP_CHUNKS - 1) / GROUP_CHUNKS;
    if (g_last <= (g_first ? g_first : 1)) return HF_OK;
    const unsigned long long ng = g_last - (g_first ? g_first : 1);
    HF_CUDA(c, cudaMemsetAsync(&work->flags[3], 0, 8, c->stream));
    HF_PROF(c, "dec_fix2_kernel"); dec_fix2_kernel<<<(unsigned)((ng + 127) / 128), 128, 0, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch, g_first, g_last, speculative ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    unsigned long long grid = (ng + S4_WARPS - 1) / S4_WARPS;
    if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "dec_regroup4_kernel");
    dec_regroup4_kernel<<<(unsigned)grid, S4_THREADS, S4_SMEM, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch, g_first, g_last, speculative ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_verify_kernel"); dec_verify_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, c->stream>>>(d_tab, work, nch, g_first, g_last, speculative ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_fix2_serial_kernel"); dec_fix2_serial_kernel<<<1, 1, 0, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch, c0, c1, speculative ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// -------------------------------------------------------------------------------------------------
// -------------------------------------------------------------------------------------------------
// chunks [c0, c1), c0 a multiple of GROUP_CHUNKS; tail_only ignores the range
int launch_sync2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                 unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                 unsigned long long c0, unsigned long long c1, bool tail_only, bool speculative)
{
    if (!c->smem_attr[ATTR_SYNC]) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_sync4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S4_SMEM));
        HF_CUDA(c, cudaFuncSetAttribute(dec_regroup4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S4_SMEM));
        c->smem_attr[ATTR_SYNC] = true;
    }
    unsigned long long ngroups = (nch + GROUP_CHUNKS - 1) / GROUP_CHUNKS;
    // tail_only: the overflow past the range end, speculatively from a guessed start TAIL_CHUNKS chunks
    // (240 .. 256 KiB of self-synchronisation) before it
    unsigned long long g_first = tail_only ? tail_first_chunk(nch) / GROUP_CHUNKS : 0;
    if (!tail_only) {
        if (c0 % GROUP_CHUNKS) return set_err(c, HF_ERR_INTERNAL, "decode slice does not start at a group");
        g_first = c0 / GROUP_CHUNKS;
        ngroups = (c1 + GROUP_CHUNKS - 1) / GROUP_CHUNKS;
        if (ngroups <= g_first) return HF_OK;
    }
    unsigned long long grid4 = (ngroups - g_first + S4_WARPS - 1) / S4_WARPS;
    if (grid4 > (unsigned long long)c->sm_count) grid4 = c->sm_count;
    HF_PROF(c, "dec_sync4_kernel");
    dec_sync4_kernel<<<(unsigned)grid4, S4_THREADS, S4_SMEM, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch,
                                                                        g_first, ngroups, (tail_only || speculative) ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_write2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes,
                  const DecodeTable *d_tab, DecWork *work, unsigned long long nch, unsigned long long c0,
                  unsigned long long c1, uint16_t *out, bool check)
{
    if (c1 <= c0) return HF_OK;
    if (!c->smem_attr[ATTR_WRITE]) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_write3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
        c->smem_attr[ATTR_WRITE] = true;
    }
    // Chunks of short code words (dense_min symbols or more in their 16 KiB) go to dec_write4_kernel, the others to
    // dec_write3_kernel: both kernels run over the range and each skips the other's chunks.  Records that came from a
    // side index are checked by dec_write3_kernel only.
    uint32_t dense_min = check ? 0xFFFFFFFFu : c->write_split;
    if (c->write_kernel == 3) dense_min = 0xFFFFFFFFu;
    if (c->write_kernel == 4 && !check) dense_min = 0;
    if (dense_min != 0xFFFFFFFFu) {
        if (!c->smem_attr[ATTR_WRITE4]) {
            HF_CUDA(c, cudaFuncSetAttribute(dec_write4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W4_SMEM));
            c->smem_attr[ATTR_WRITE4] = true;
        }
        unsigned long long grid = (c1 - c0 + W4_WARPS - 1) / W4_WARPS;
        if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
        HF_PROF(c, "dec_write4_kernel");
        dec_write4_kernel<<<(unsigned)grid, W4_THREADS, W4_SMEM, c->stream>>>(frame, frame_bytes, d_tab, work, nch, c0, c1, out, dense_min);
        HF_LAUNCH_CHECK(c);
        if (dense_min == 0) return HF_OK;
    }
    // units per run: a whole chunk per warp when that still gives every warp of the machine 16 runs or more (fewer
    // runs per warp leave the warps that got one less idle at the end)
    uint32_t upw = DEC_THREADS / 32;
    while (upw > 1 && (c1 - c0) * ((DEC_THREADS / 32) / upw) < 16ull * c->sm_count * W3_WARPS) upw >>= 1;
    const unsigned long long nruns = (c1 - c0) * ((DEC_THREADS / 32) / upw);
    unsigned long long grid = (nruns + W3_WARPS - 1) / W3_WARPS;
    if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "dec_write3_kernel");
    dec_write3_kernel<<<(unsigned)grid, W3_THREADS, W3_SMEM, c->stream>>>(frame, frame_bytes, d_tab, work, nch, c0, c1, out, check ? 1u : 0u, upw, dense_min);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
