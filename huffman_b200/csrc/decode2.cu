// decode2.cu — the hot kernels of the decoder: dec_sync3_kernel (code word boundaries and symbol counts per
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
//   * re-synchronisation is detected at checkpoints every 128 bits (the first code word boundary at or after
//     the checkpoint, with the symbol count of every segment), which costs two instructions per checkpoint
//     instead of a 256-bit boundary mask maintained per symbol;
//   * what the per-source-line and per-SASS-instruction views of ncu found later is listed in profiles/README.md
//     (shared-memory base addresses rebuilt inside loops, loop state in local memory around a call, loads that
//     were waited for where they were issued, single-lane tails).
//
// Algorithmic bytes: dec_sync3 reads C; dec_write3 reads C and writes N.
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
constexpr size_t W3_SMEM = (4u << MICRO_K) + NSYM * 2 + (size_t)W3_WARPS * (W3_WIN + 8) * 2;

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

// ---- synchronisation kernel ---------------------------------------------------------------------
// A thread walks a SPAN of 4 subsequences (1024 bits).  Longer spans mean fewer repeated walks: a walk
// from a wrong start re-joins the true one after tens of bits on skewed codes but only after ~2,000 bits
// on the nearly fixed-length codes of flat data, and every subsequence a wrong walk crosses costs the
// fix-point one more round.  A CTA of 1024 threads holds 8 TEAMS of 4 warps; a team stages its own GROUP (one
// chunk, 16 KiB) in shared memory, one padded row of 33 words per thread (bank = (thread + word) mod 32; the pad
// word repeats the next row's first word so a 32-bit window never leaves the row), and converges on it behind its
// own named barrier.  Small teams wait less on their slowest warp (teams of 8 warps: +7 % kernel time, of 16: +15 %);
// the price is one more group boundary per chunk for the repair pass, which is cheap.
constexpr int S3_THREADS = 1024;
constexpr int TEAM_THREADS = DEC_TEAM_THREADS;
constexpr int S3_TEAMS = S3_THREADS / TEAM_THREADS;
constexpr uint32_t SPAN_SUBS = DEC_SPAN_SUBS;
constexpr uint32_t SPAN_BITS = SPAN_SUBS * SUB_BITS;            // 1024
constexpr uint32_t SPAN_WORDS = SPAN_BITS / 32;                 // 32
constexpr uint32_t ROW_WORDS = SPAN_WORDS + 1;
constexpr unsigned long long GROUP_BITS = (unsigned long long)TEAM_THREADS * SPAN_BITS;
constexpr uint32_t SEG_BITS = 128;                              // checkpoint spacing
constexpr uint32_t NSEG = SPAN_BITS / SEG_BITS;                 // 8
constexpr size_t S3_SMEM = (4u << MICRO_K) + (size_t)S3_THREADS * ROW_WORDS * 4;

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

// Record of a walk: chkpos byte k = (first code word boundary at or after bit 128k) - 128k; chkcnt byte k = code
// words starting in [128k, 128k + 128).  CHK_NONE (bytes 0xFF, no checkpoint offset is that large) = no walk yet.
struct Chk { uint32_t pos[2], cnt[2]; };
constexpr uint32_t CHK_NONE = 0xFFFFFFFFu;

// Walks the code words of a span from bit `start` to `lim`.  RESYNC: `rec` describes an earlier walk; stop at
// the first checkpoint both walks share and keep the earlier record from there on (`end` stays the earlier one's).
template <bool RESYNC, bool MULTI>
__device__ __forceinline__ void walk_span(const uint32_t *row, uint32_t t14_a, const DecodeTable *tab,
                                          const uint8_t *frame, unsigned long long frame_bytes,
                                          unsigned long long span_bit0, uint32_t k2shift, uint32_t start, uint32_t lim,
                                          Chk &rec, uint32_t &end, uint32_t &bad)
{
    uint32_t pos = start, n = 0, wl = lim;
    uint32_t npos[2] = {0, 0}, ncnt[2] = {0, 0};
    // t14_a: the table's 32-bit shared address as a plain register value (opaque_shared_addr); left to itself the
    // compiler rebuilds the shared window base inside every walk loop, four uniform instructions per code word
    const uint32_t row_a = (uint32_t)__cvta_generic_to_shared(row);
    int kept = NSEG;                                    // checkpoints >= kept keep the earlier record
#pragma unroll
    for (int k = 0; k < (int)NSEG; k++) {
        if (k > 0) { ncnt[(k - 1) >> 2] |= n << (8 * ((k - 1) & 3)); n = 0; }
        const uint32_t rel = (pos - SEG_BITS * k) & 0xFFu;
        if (RESYNC && k > 0) {          // still walking, and on the earlier walk's boundary: the walks have met
            if (pos < wl && rel == ((rec.pos[k >> 2] >> (8 * (k & 3))) & 0xFFu)) { wl = 0; kept = k; }
        }
        npos[k >> 2] |= rel << (8 * (k & 3));
        const uint32_t lw = min(wl, SEG_BITS * (k + 1));
        while (pos < lw) {
            const uint32_t wa = row_a + ((pos >> 5) << 2);
            const uint32_t win = __funnelshift_l(lds32(wa + 4), lds32(wa), pos);
            const uint32_t e14 = lds32(t14_a + ((win >> (30 - MICRO_K)) & ((4u << MICRO_K) - 4u)));    // the d14 plane
            const uint32_t deep = (MICRO_K + 1) + ((e14 >> ((win >> (32 - MICRO_MAX - 1)) & 30u)) & 3u);  // micro tree: 2 bits per slot
            const bool micro = (e14 & 0xFu) != 0xCu;
            uint32_t len = micro ? deep : (e14 >> 28), cnt = 1;
            if (MULTI) {                            // all the code words the 14 bits hold, when they end inside the segment
                const uint32_t tot = (e14 >> 4) & 0xFu;
                if (!micro && pos + tot <= lw) { len = tot; cnt = (e14 >> 8) & 0xFu; }
            }
            if (len == 0) {
                len = __ldg(tab->lenflat + (win >> k2shift));
                if (len == 0) {
                    const uint32_t e = slow_decode(tab, frame, frame_bytes, span_bit0 + pos);
                    bad |= e >> 31;
                    len = e & 0x7Fu;
                }
            }
            pos += len;
            n += cnt;
        }
    }
    ncnt[1] |= n << 24;
    if (RESYNC && kept < (int)NSEG) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // bytes of half h that belong to checkpoints >= kept
            const int first = kept - 4 * h;             // first kept byte of this half (may be <= 0 or >= 4)
            const uint32_t keep = first <= 0 ? 0xFFFFFFFFu : (first >= 4 ? 0u : 0xFFFFFFFFu << (8 * first));
            rec.pos[h] = (npos[h] & ~keep) | (rec.pos[h] & keep);
            rec.cnt[h] = (ncnt[h] & ~keep) | (rec.cnt[h] & keep);
        }
    } else {
        rec.pos[0] = npos[0]; rec.pos[1] = npos[1];
        rec.cnt[0] = ncnt[0]; rec.cnt[1] = ncnt[1];
        end = pos - lim;
    }
}

__device__ __forceinline__ uint32_t span_limit(unsigned long long X, unsigned long long range_end_bit)
{   // bits of the span at frame bit X that lie before the end of the range (0 .. SPAN_BITS)
    if (X >= range_end_bit) return 0u;
    const unsigned long long room = range_end_bit - X;
    return room >= SPAN_BITS ? SPAN_BITS : (uint32_t)room;
}

__device__ __forceinline__ void team_sync(uint32_t team)
{
    asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(TEAM_THREADS) : "memory");
}
__device__ __forceinline__ bool team_or(uint32_t team, bool pred)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred), "r"(team + 1), "r"(TEAM_THREADS) : "memory");
    return r != 0;
}

// what the threads of a synchronisation CTA share
struct SyncCtx {
    const uint8_t *frame;
    unsigned long long frame_bytes, F0, range_end_bit, nch;
    const DecodeTable *tab;
    DecWork *work;
    uint32_t t14_a;                     // shared address of the d14 plane (opaque_shared_addr)
    uint32_t *s_bits, *s_wend, *s_red;
    uint32_t g, k2shift;                // gcd of the code lengths (1 when speculating), 32 - k2
};

// One team converges on one group (GROUP_CHUNKS chunks): per-subsequence records, chunk totals and overflows.
// exact: the group's first code word starts `start` bits into it (the stream head, or the true overflow of the group
// before when a group is redone); otherwise the first span starts from a guess like every other one.
// dense: code words my warp counted in the group before (0 at first).  Short code words come several to a 14-bit
// look-up; when the warp's last 32 Kbit held more than DENSE_MIN of them (under ~10 bits each) the walks take all the
// code words an entry holds in one step (walk_span<.., true>), which costs every step a few instructions more.
constexpr uint32_t DENSE_MIN = 32u * SPAN_BITS / 10u;
__device__ __forceinline__ void sync_group(const SyncCtx &S, unsigned long long grp, bool exact, uint32_t start, uint32_t &bad,
                                           uint32_t &dense)
{
    const bool multi = dense > DENSE_MIN;
    const uint8_t *frame = S.frame;
    const unsigned long long frame_bytes = S.frame_bytes, F0 = S.F0, range_end_bit = S.range_end_bit, nch = S.nch;
    const DecodeTable *tab = S.tab;
    DecWork *work = S.work;
    const uint32_t t14_a = S.t14_a;
    uint32_t *s_bits = S.s_bits, *s_wend = S.s_wend, *s_red = S.s_red;
    const uint32_t g = S.g, k2shift = S.k2shift;
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t team = tid / TEAM_THREADS, tt = tid % TEAM_THREADS;      // my team, my index in it
    const uint32_t *row = s_bits + tid * ROW_WORDS;
    uint32_t *team_rows = s_bits + team * TEAM_THREADS * ROW_WORDS;

    team_sync(team);                                // my team's rows, s_wend, s_red are free again
    // ---- stage the group: 8 coalesced 128-bit loads per thread ----
    const unsigned long long gbyte0 = grp * (GROUP_BITS / 8);
#pragma unroll
    for (uint32_t k = 0; k < SPAN_WORDS / 4; k++) {
        const uint32_t v = tt + k * TEAM_THREADS;   // 16-byte vector of the group
        const unsigned long long b = gbyte0 + 16ull * v;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (b < frame_bytes) x = ld_stream_v4(frame + b);          // the frame is 16-byte aligned
        uint32_t *dst = team_rows + (v >> 3) * ROW_WORDS + 4 * (v & 7);
        dst[0] = bswap32(x.x); dst[1] = bswap32(x.y); dst[2] = bswap32(x.z); dst[3] = bswap32(x.w);
    }
    {   // the pad word of my row: the first word of the next span
        const unsigned long long b = gbyte0 + (unsigned long long)(tt + 1) * (SPAN_BITS / 8);
        uint32_t x = 0;
        if (b < frame_bytes) x = bswap32(__ldg(reinterpret_cast<const uint32_t *>(frame + b)));
        s_bits[tid * ROW_WORDS + SPAN_WORDS] = x;
    }
    team_sync(team);

    const unsigned long long X = grp * GROUP_BITS + (unsigned long long)tt * SPAN_BITS;
    const uint32_t lim = span_limit(X, range_end_bit);     // code words starting at or after the range end are not ours
    const bool fixed = exact && tt == 0;                    // exact: the group's first code word starts at `start`
    uint32_t p = fixed ? start : (X >= F0 ? spec_start(X, F0, g) : 0u);
    uint32_t end = 0;
    Chk rec{{CHK_NONE, CHK_NONE}, {0, 0}};
    // `rec` describes the walk from rec_p, which ended at rec_end.  memo: up to four (start + 1, end) pairs of
    // walks this thread has done on this span.  Data that does not re-synchronise (a long run of one code word
    // is periodic: a walk that enters it out of phase leaves it out of phase) makes the fix-point hand a lane
    // the same few starts again and again; a remembered start costs no walk, and the record of the final start
    // is rebuilt once at the end.
    uint32_t rec_p = p, rec_end = 0, mslot = 0, nwalk = 0;
    unsigned long long memo = 0;
    auto memo_add = [&](uint32_t st, uint32_t en) {
        memo = (memo & ~(0xFFFFull << (16 * mslot))) | ((unsigned long long)(((st + 1) << 8) | en) << (16 * mslot));
        mslot = (mslot + 1) & 3;
        nwalk++;
    };
    if (lim) {
        if (p < lim) {
            if (multi) walk_span<false, true>(row, t14_a, tab, frame, frame_bytes, X, k2shift, p, lim, rec, end, bad);
            else walk_span<false, false>(row, t14_a, tab, frame, frame_bytes, X, k2shift, p, lim, rec, end, bad);
            memo_add(p, end);
        }
        else end = p - lim;
        rec_end = end;
    }
    // Fix-point: my true start is my predecessor's overflow.  Inside a warp the overflows travel by
    // shuffle and the warps iterate on their own; the warps then exchange their last overflow through
    // shared memory, which usually moves only lane 0 of each warp.
    const bool movable = !fixed && lim != 0;
    uint32_t q0 = p;                                // lane 0's start: the guess, then the previous warp's overflow
    for (uint32_t round = 0; round < TEAM_THREADS / 32 + 2; round++) {
        for (;;) {
            uint32_t q = __shfl_up_sync(0xFFFFFFFFu, end, 1);
            if (lane == 0) q = q0;
            const bool need = movable && q != p;
            if (!__any_sync(0xFFFFFFFFu, need)) break;
            if (need) {
                uint32_t hit = 0;                   // ((q + 1) << 8) | end of a walk from q done before
                if (nwalk >= 2) {                   // the first correction of a guess cannot be a repeat
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t en = (uint32_t)(memo >> (16 * j)) & 0xFFFFu;
                        if ((en >> 8) == q + 1) hit = en;
                    }
                }
                if (hit) {
                    end = hit & 0xFFu;              // rec stays with the walk it describes
                } else if (q < lim) {
                    end = rec_end;                  // a merge keeps the recorded walk's end
                    if (multi) walk_span<true, true>(row, t14_a, tab, frame, frame_bytes, X, k2shift, q, lim, rec, end, bad);
                    else walk_span<true, false>(row, t14_a, tab, frame, frame_bytes, X, k2shift, q, lim, rec, end, bad);
                    rec_p = q; rec_end = end;
                    memo_add(q, end);
                } else {
                    rec.pos[0] = rec.pos[1] = CHK_NONE; rec.cnt[0] = rec.cnt[1] = 0; end = q - lim;
                    rec_p = q; rec_end = end;
                }
                p = q;
            }
        }
        if (lane == 31) s_wend[wid] = end;
        team_sync(team);
        q0 = (tt >> 5) ? s_wend[wid - 1] : p;
        if (!team_or(team, lane == 0 && movable && q0 != p)) break;
    }
    if (lim && rec_p != p) {                        // the final start was a remembered one: rebuild its record
        if (p < lim) {
            end = rec_end;
            if (multi) walk_span<true, true>(row, t14_a, tab, frame, frame_bytes, X, k2shift, p, lim, rec, end, bad);
            else walk_span<true, false>(row, t14_a, tab, frame, frame_bytes, X, k2shift, p, lim, rec, end, bad);
        } else {
            rec.pos[0] = rec.pos[1] = CHK_NONE; rec.cnt[0] = rec.cnt[1] = 0; end = p - lim;
        }
    }

    // ---- per-subsequence records: start offset (6 bits) | code words (10 bits), 4 per thread ----
    uint32_t cnt4[SPAN_SUBS];
    uint32_t total = 0;
    unsigned long long packed = 0;
#pragma unroll
    for (uint32_t j = 0; j < SPAN_SUBS; j++) {
        const uint32_t pj = (rec.pos[j >> 1] >> (16 * (j & 1))) & 0xFFu;            // checkpoint 2j
        const uint32_t cj = (rec.cnt[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;          // segments 2j, 2j + 1
        cnt4[j] = (cj & 0xFFu) + (cj >> 8);
        total += cnt4[j];
        packed |= (unsigned long long)((pj & 63u) | (cnt4[j] << 6)) << (16 * j);
    }
    const unsigned long long sub_index = grp * (GROUP_BITS / SUB_BITS) + (unsigned long long)tt * SPAN_SUBS;
    if (sub_index < nch * DEC_THREADS)              // info holds whole chunks: 4 records never straddle its end
        *reinterpret_cast<unsigned long long *>(L.info + sub_index) = packed;
    // chunk totals: a chunk is 128 consecutive threads (4 warps)
    uint32_t v = total;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) s_red[wid] = v;
    dense = v;
    team_sync(team);
    constexpr uint32_t TPC = DEC_THREADS / SPAN_SUBS;       // threads per chunk
    const unsigned long long c = grp * GROUP_CHUNKS + tt / TPC;
    if (c < nch) {
        if (tt % TPC == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (uint32_t i = 0; i < TPC / 32; i++) tot += s_red[wid + i];
            L.chunkCnt[c] = tot;
            L.chunkE2[c] = 0xFFFFFFFFu;
        }
        if (tt % TPC == TPC - 1) L.chunkE[c] = end;
    }
    // the thread whose span holds the end of the range reports the overflow past it
    if (lim && span_limit(X + SPAN_BITS, range_end_bit) == 0) work->result[1] = end;
}

__global__ void __launch_bounds__(S3_THREADS, 1)
dec_sync3_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes, unsigned long long F0,
                 unsigned long long range_end_bit, const DecodeTable *__restrict__ tab, DecWork *work,
                 unsigned long long nch, unsigned long long g_first, unsigned long long g_last, uint32_t speculative)
{
    extern __shared__ __align__(16) uint32_t s3_smem[];
    uint32_t *s_t14 = s3_smem;                          // 2^MICRO_K
    uint32_t *s_bits = s3_smem + (1u << MICRO_K);       // S3_THREADS rows of ROW_WORDS
    __shared__ uint32_t s_wend[S3_THREADS / 32];
    __shared__ uint32_t s_red[S3_THREADS / 32];
    if (tab->single_sym) return;                        // empty payload, see dec_fill_kernel
    const uint32_t tid = threadIdx.x, team = tid / TEAM_THREADS;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab->d14);       // lengths only
        uint4 *dst = reinterpret_cast<uint4 *>(s_t14);
        for (uint32_t i = tid; i < (4u << MICRO_K) / 16; i += S3_THREADS) dst[i] = __ldg(src + i);
    }
    const SyncCtx S{frame, frame_bytes, F0, range_end_bit, nch, tab, work, opaque_shared_addr(s_t14), s_bits, s_wend, s_red,
                    speculative ? 1u : tab->len_gcd, 32u - tab->k2};
    uint32_t bad = 0;
    __syncthreads();                                    // planes loaded

    uint32_t dense = 0;
    for (unsigned long long grp = g_first + (unsigned long long)blockIdx.x * S3_TEAMS + team; grp < g_last;
         grp += (unsigned long long)gridDim.x * S3_TEAMS)
        sync_group(S, grp, grp == 0 && !speculative, (uint32_t)F0, bad, dense);
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// Groups the cheap repair (dec_fix2_kernel) gave up on: the chain of the group before never meets the chain the
// group recorded from its guessed start (data that does not re-synchronise, e.g. a long run of one code word).
// A team takes the head of every run of such groups and redoes the groups one after the other from their TRUE
// starts, walking on into the following groups for as long as the overflow it hands over is not the start they
// recorded.  Runs are independent of each other; a chain that reaches another run's groups is caught by
// dec_verify_kernel and settled by the serial kernel.
constexpr uint32_t CHUNK_DIRTY = 0xFFFFFFFEu;           // in chunkE2 of a group's first chunk

__global__ void __launch_bounds__(S3_THREADS, 1)
dec_regroup_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes, unsigned long long F0,
                   unsigned long long range_end_bit, const DecodeTable *__restrict__ tab, DecWork *work,
                   unsigned long long nch, unsigned long long g_first, unsigned long long g_last, uint32_t speculative)
{
    if (work->flags[3] == 0) return;                    // no group was given up on
    extern __shared__ __align__(16) uint32_t s3_smem[];
    uint32_t *s_t14 = s3_smem;
    uint32_t *s_bits = s3_smem + (1u << MICRO_K);
    __shared__ uint32_t s_wend[S3_THREADS / 32];
    __shared__ uint32_t s_red[S3_THREADS / 32];
    if (tab->single_sym) return;
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, team = tid / TEAM_THREADS;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab->d14);
        uint4 *dst = reinterpret_cast<uint4 *>(s_t14);
        for (uint32_t i = tid; i < (4u << MICRO_K) / 16; i += S3_THREADS) dst[i] = __ldg(src + i);
    }
    const SyncCtx S{frame, frame_bytes, F0, range_end_bit, nch, tab, work, opaque_shared_addr(s_t14), s_bits, s_wend, s_red,
                    speculative ? 1u : tab->len_gcd, 32u - tab->k2};
    uint32_t bad = 0;
    __syncthreads();

    const unsigned long long g_lo = g_first ? g_first : 1;      // the first group of a stream has an exact start
    uint32_t dense = 0;
    for (unsigned long long g = g_lo + (unsigned long long)blockIdx.x * S3_TEAMS + team; g < g_last;
         g += (unsigned long long)gridDim.x * S3_TEAMS) {
        // the head of a run: given up on, and the group before was not
        if (L.chunkE2[g * GROUP_CHUNKS] != CHUNK_DIRTY) continue;
        if (g > g_lo && L.chunkE2[(g - 1) * GROUP_CHUNKS] == CHUNK_DIRTY) continue;
        for (unsigned long long cur = g;;) {
            const uint32_t s = L.chunkE[cur * GROUP_CHUNKS - 1];    // final: the group before is settled
            sync_group(S, cur, true, s, bad, dense);                // also resets chunkE2 of its chunks
            team_sync(team);                                        // the group's records are in global memory
            const unsigned long long next = cur + 1;
            if (next >= g_last) break;
            const bool dirty = L.chunkE2[next * GROUP_CHUNKS] == CHUNK_DIRTY;
            const bool meets = L.chunkE[next * GROUP_CHUNKS - 1] == (uint32_t)(L.info[next * GROUP_CHUNKS * DEC_THREADS] & 63u);
            if (!dirty && meets) break;
            cur = next;
        }
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
}

// every group must start where the group before it ends; what the parallel repairs left open goes to the serial kernel
__global__ void dec_verify_kernel(const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch,
                                  unsigned long long g_first, unsigned long long g_last)
{
    const unsigned long long g = (g_first ? g_first : 1) + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_last || tab->single_sym) return;
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
dec_write3_kernel(const uint8_t *__restrict__ frame, unsigned long long frame_bytes, unsigned long long F0,
                  const DecodeTable *__restrict__ tab, DecWork *work, unsigned long long nch,
                  unsigned long long c0, unsigned long long c1, unsigned long long n_symbols,
                  uint16_t *__restrict__ out, uint32_t check, uint32_t upw)
{
    extern __shared__ __align__(16) uint32_t w3_smem[];
    uint32_t *s_t14 = w3_smem;                                                  // 2^MICRO_K
    uint16_t *s_leaves = reinterpret_cast<uint16_t *>(s_t14 + (1u << MICRO_K)); // NSYM
    if (tab->single_sym) return;
    DecLayout L(work, nch);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint16_t *sout = s_leaves + NSYM + wid * (W3_WIN + 8);                      // this warp's window
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab->t14);
        uint4 *dst = reinterpret_cast<uint4 *>(s_t14);
        for (uint32_t i = tid; i < (4u << MICRO_K) / 16; i += W3_THREADS) dst[i] = __ldg(src + i);
        src = reinterpret_cast<const uint4 *>(tab->leaves);
        dst = reinterpret_cast<uint4 *>(s_leaves);
        for (uint32_t i = tid; i < NSYM * 2 / 16; i += W3_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
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
        if (cbase >= n_symbols) continue;
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
        uint32_t pos = (c == 0 && t == 0) ? (uint32_t)F0 : (inf & 63u);     // the stream head may sit past bit 63
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
        for (uint32_t w0 = 0; w0 < (uint32_t)total; w0 += W3_WIN) {
            const uint32_t wend = min((uint32_t)total, w0 + W3_WIN);
            const uint32_t o_end = min(my_end, wend);
            if (o < o_end) {
                uint16_t *sp = sout + (o - w0 + mis);
                uint16_t *const sp_end = sout + (o_end - w0 + mis);
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
                        const uint32_t e14 = s_t14[win >> (32 - MICRO_K)];
                        uint32_t len, sym;
                        if (e14 & MICRO_FLAG) {
                            uint32_t leaf;
                            micro_decode(e14, win, len, leaf);
                            sym = s_leaves[leaf];
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
                        *sp++ = (uint16_t)sym;
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
__global__ void idx_total_kernel(DecWork *work, unsigned long long n_symbols)
{
    if (work->result[2] != n_symbols) atomicExch(&work->flags[1], 1ull);
}

int launch_idx_total(Ctx *c, DecWork *work, unsigned long long n_symbols)
{
    HF_PROF(c, "idx_total_kernel"); idx_total_kernel<<<1, 1, 0, c->stream>>>(work, n_symbols);
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
                           uint32_t &bad)
{
    uint16_t *info = L.info + c * DEC_THREADS;
    const uint32_t k2shift = 32u - tab->k2;
    uint32_t q = s;
    long long delta = 0;
    for (uint32_t t = 0; t < DEC_THREADS; t++) {
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
                                unsigned long long nch, unsigned long long g_first, unsigned long long g_last)
{
    const unsigned long long g = (g_first ? g_first : 1) + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_last || tab->single_sym) return;
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
                                       unsigned long long c1)
{
    if (work->flags[0] == 0 || tab->single_sym) return;
    DecLayout L(work, nch);
    uint32_t bad = 0;
    for (unsigned long long c = c0 ? c0 : 1; c < c1; c++) {
        const uint32_t s = L.chunkE[c - 1];
        if (s == (uint32_t)(L.info[c * DEC_THREADS] & 63u)) continue;
        fix_chunk2(tab, frame, frame_bytes, range_end_bit, work, L, c, s, bad);
    }
    if (bad) atomicExch(&work->flags[1], 1ull);
    work->flags[0] = 0;                                 // settled (the next slice starts clean)
}

// chunks [c0, c1) (a slice of the stream, or all of it); everything before c0 is final
int launch_fix2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long F0,
                unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                unsigned long long c0, unsigned long long c1, bool speculative)
{
    if (c1 <= c0) return HF_OK;
    const unsigned long long g_first = c0 / GROUP_CHUNKS, g_last = (c1 + GROUP_CHUNKS - 1) / GROUP_CHUNKS;
    if (g_last <= (g_first ? g_first : 1)) return HF_OK;
    const unsigned long long ng = g_last - (g_first ? g_first : 1);
    HF_CUDA(c, cudaMemsetAsync(&work->flags[3], 0, 8, c->stream));
    HF_PROF(c, "dec_fix2_kernel"); dec_fix2_kernel<<<(unsigned)((ng + 127) / 128), 128, 0, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch, g_first, g_last);
    HF_LAUNCH_CHECK(c);
    unsigned long long grid = (ng + S3_TEAMS - 1) / S3_TEAMS;
    if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "dec_regroup_kernel");
    dec_regroup_kernel<<<(unsigned)grid, S3_THREADS, S3_SMEM, c->stream>>>(frame, frame_bytes, F0, range_end_bit, d_tab, work, nch, g_first, g_last, speculative ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_verify_kernel"); dec_verify_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, c->stream>>>(d_tab, work, nch, g_first, g_last);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "dec_fix2_serial_kernel"); dec_fix2_serial_kernel<<<1, 1, 0, c->stream>>>(frame, frame_bytes, range_end_bit, d_tab, work, nch, c0, c1);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

// -------------------------------------------------------------------------------------------------
// -------------------------------------------------------------------------------------------------
// chunks [c0, c1), c0 a multiple of GROUP_CHUNKS; tail_only ignores the range
int launch_sync2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long F0,
                 unsigned long long range_end_bit, const DecodeTable *d_tab, DecWork *work, unsigned long long nch,
                 unsigned long long c0, unsigned long long c1, bool tail_only)
{
    if (!c->smem_attr[ATTR_SYNC]) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_sync3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3_SMEM));
        HF_CUDA(c, cudaFuncSetAttribute(dec_regroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3_SMEM));
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
    unsigned long long grid = (ngroups - g_first + S3_TEAMS - 1) / S3_TEAMS;
    if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "dec_sync3_kernel");
    dec_sync3_kernel<<<(unsigned)grid, S3_THREADS, S3_SMEM, c->stream>>>(frame, frame_bytes, F0, range_end_bit, d_tab, work, nch,
                                                                       g_first, ngroups, tail_only ? 1u : 0u);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_write2(Ctx *c, const uint8_t *frame, unsigned long long frame_bytes, unsigned long long F0,
                  const DecodeTable *d_tab, DecWork *work, unsigned long long nch, unsigned long long c0,
                  unsigned long long c1, unsigned long long n_symbols, uint16_t *out, bool check)
{
    if (c1 <= c0) return HF_OK;
    if (!c->smem_attr[ATTR_WRITE]) {
        HF_CUDA(c, cudaFuncSetAttribute(dec_write3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
        c->smem_attr[ATTR_WRITE] = true;
    }
    // units per run: a whole chunk per warp when that still gives every warp of the machine 16 runs or more (fewer
    // runs per warp leave the warps that got one less idle at the end)
    uint32_t upw = DEC_THREADS / 32;
    while (upw > 1 && (c1 - c0) * ((DEC_THREADS / 32) / upw) < 16ull * c->sm_count * W3_WARPS) upw >>= 1;
    const unsigned long long nruns = (c1 - c0) * ((DEC_THREADS / 32) / upw);
    unsigned long long grid = (nruns + W3_WARPS - 1) / W3_WARPS;
    if (grid > (unsigned long long)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "dec_write3_kernel");
    dec_write3_kernel<<<(unsigned)grid, W3_THREADS, W3_SMEM, c->stream>>>(frame, frame_bytes, F0, d_tab, work, nch, c0, c1, n_symbols, out, check ? 1u : 0u, upw);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
