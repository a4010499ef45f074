// encode2.cu — warp-independent Huffman encoder.
//
// Replaces populateCWLength + thrust::transform_inclusive_scan + encodeFromCW + the host tail flush
// (/root/reference/Compressor.cu:50-74, :152-313, :541-601, :673-684): the reference materialises 12 bytes of
// scratch per symbol and binary-searches the offsets once per OUTPUT byte.
//
// Work is cut into UNITS of 512 symbols (1 KiB of input, 16 symbols per lane) and GROUPS of 32 units.
//   enc_bits_kernel     bits of every unit (u32) and of every group, from the 64 KiB length plane in shared
//                       memory.  Reads N, writes N / 256.
//                       The group totals are scanned in the same pass (decoupled look-back on two levels,
//                       scan_publish / scan_resolve): every group's start bit is known when the kernel ends.
//   encode2_kernel      one persistent CTA per SM, 20 warps, the 192 KiB code table (24-bit entries in two planes,
//                       index XOR-folded against bank conflicts) in shared memory.  Every WARP packs groups on its
//                       own: unit start = group start + a shuffle scan of the 32 unit counts; the lanes look their
//                       16 codes up, a shuffle scan of the lane totals gives the lane's bit offset, a lane strings its
//                       codes together in registers and stores every 32-bit word that fills up into the warp's zeroed
//                       staging window (plain predicated stores; one red.shared per lane for the leftover), and the
//                       window leaves with aligned 128-bit stores.  Between two units of the common case the partial
//                       last word is carried in a register to the unit that holds its last bit; elsewhere a 32-bit
//                       word belongs to the unit that holds its first bit: the owner completes its last, partial
//                       word by encoding the symbols that FOLLOW the unit, so groups exchange nothing and there is
//                       no CTA barrier after the table is loaded.  Reads N, writes C.
// Codes longer than 23 bits (not in the shared table) and units whose bits exceed the staging window take a
// slow per-symbol path through the global codebook.
// Bits before the start phase in the first byte are preserved (they belong to the header or to the previous shard,
// C:541, C:294-310); the last byte is zero padded (C:597-601).
//
// Algorithmic bytes: N read + C written (traffic 2N + C + N/128).  Roofline: HBM; today bound by shared-memory look-ups / issue.
#include "common.cuh"

namespace hf {

constexpr uint32_t UNIT_SYMS = 512;
constexpr uint32_t GROUP_UNITS = 32;
constexpr uint32_t GROUP_SYMS = UNIT_SYMS * GROUP_UNITS;            // 16,384
constexpr int BITS_THREADS = 512;
constexpr size_t BITS_SMEM = NSYM + 16;                              // lenf plane + the mbarrier of its load
#ifndef HF_E2_WARPS
#define HF_E2_WARPS 20
#endif
constexpr int E2_WARPS = HF_E2_WARPS;
constexpr int E2_THREADS = E2_WARPS * 32;
constexpr uint32_t E2_PLANE_BYTES = NSYM * 3;                       // p16 + p8
constexpr uint32_t E2_WIN = 372;                                    // staging words per warp (multiple of 4)
constexpr uint32_t E2_BAR = E2_PLANE_BYTES + E2_WARPS * E2_WIN * 4;    // the mbarrier of the plane load
constexpr size_t E2_SMEM = E2_BAR + 16;
constexpr unsigned long long NOT_FINAL = ~0ull;

__device__ __forceinline__ uint32_t fold16(uint32_t sym) { return sym ^ (sym >> 8); }      // involution on 16 bits
// the same for the two symbols of a 32-bit input word at once: two instructions instead of four
__device__ __forceinline__ uint32_t fold16x2(uint32_t w) { return w ^ ((w >> 8) & 0x00FF00FFu); }

// shared-memory loads by 32-bit shared address (taken once with __cvta_generic_to_shared)
__device__ __forceinline__ uint32_t lds8(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));        // tables that do not change after the CTA's barrier
    return v;
}

struct Enc2Work {                       // device arrays in ctx->ws
    uint32_t *unit_bits;                // [ngroups * 32]
    unsigned long long *group_start;    // [ngroups]  payload bits before the group
    unsigned long long *desc;           // scan descriptors (zeroed per call): [ngroups] group totals, then [nblocks]
    unsigned long long *bdesc;          // block descriptors, then the group ticket (bdesc[nblocks])
};

// ---- single-pass scan of the group totals: decoupled look-back on two levels --------------------------
// A descriptor is one 64-bit word, flag and value together, so a reader never sees one without the other; zero = not
// there yet.  Every warp publishes DESC_AGG | bits of its group as soon as it has counted it.  Groups form blocks of
// 32; the warp that counted a block's LAST group (its leader) sums the block's totals, publishes DESC_AGG | block total,
// looks back over the blocks before it — 32 block descriptors per round, a lane each, to the nearest one whose
// inclusive sum is known — and publishes DESC_INC | bits up to and including the block.  A group's start is then ONE
// round: the inclusive sum of the block before + the totals of the groups before it in its block.  The warp reads
// those while it counts its NEXT group (scan_peek) and adds them up afterwards (scan_resolve), when the leader has
// long finished: nobody but the leaders (1 warp step in 32) ever waits for a look-back.  (With one level every group
// of a generation of ~7,000 concurrent warps looks back at the same moment and finds only totals: 2 dependent
// rounds per group, 10 % of the kernel.)  Groups are handed out by a ticket, so every group before g belongs to a warp
// that is already counting it or has counted it, and whose waits are all for lower groups: every wait ends.
constexpr unsigned long long DESC_AGG = 1ull << 62, DESC_INC = 2ull << 62, DESC_VAL = DESC_AGG - 1;
constexpr uint32_t SCAN_BLOCK = 32;     // groups per block: one look-back round
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// the warp has counted group g (its total in tot): publish it; the leader of a block also scans the blocks
__device__ __forceinline__ void scan_publish(const Enc2Work &W, uint64_t g, uint64_t ngroups, uint32_t tot, uint32_t lane)
{
    if (lane == 0) st_desc(W.desc + g, DESC_AGG | tot);
    const uint64_t B = g / SCAN_BLOCK;
    const uint32_t r = (uint32_t)(g % SCAN_BLOCK);
    if (r != SCAN_BLOCK - 1 && g + 1 != ngroups) return;
    // leader: the block's total (the groups before mine are being counted by warps that started before me)
    unsigned long long d;
    do d = lane < r ? ld_desc(W.desc + B * SCAN_BLOCK + lane) : DESC_AGG;
    while (__any_sync(0xFFFFFFFFu, d == 0));
    const unsigned long long bt = warp_sum64(d & DESC_VAL) + tot;
    if (B == 0) {
        if (lane == 0) st_desc(W.bdesc, DESC_INC | bt);
        return;
    }
    if (lane == 0) st_desc(W.bdesc + B, DESC_AGG | bt);
    unsigned long long excl = 0;
    for (uint64_t base = B;;) {                                     // this round looks at blocks base - 1 - lane
        uint32_t inc, need;
        do {
            d = base > lane ? ld_desc(W.bdesc + (base - 1 - lane)) : DESC_INC;      // before block 0: an inclusive sum of zero
            inc = __ballot_sync(0xFFFFFFFFu, (d & DESC_INC) != 0);
            need = inc ? ((2u << (__ffs(inc) - 1)) - 1u) : 0xFFFFFFFFu;             // lanes up to the nearest inclusive sum
        } while (__ballot_sync(0xFFFFFFFFu, d == 0) & need);
        excl += warp_sum64(((1u << lane) & need) ? (d & DESC_VAL) : 0ull);
        if (inc) break;
        base -= 32;
    }
    if (lane == 0) st_desc(W.bdesc + B, DESC_INC | (excl + bt));
}
// what my lane adds to the start of group g: lanes below g % 32 the total of a group before it in its block, lane 31
// the inclusive sum of the block before (0 = not there yet; read ahead of time: any later state serves as well)
__device__ __forceinline__ unsigned long long scan_peek(const Enc2Work &W, uint64_t g, uint32_t lane)
{
    const uint64_t B = g / SCAN_BLOCK;
    const uint32_t r = (uint32_t)(g % SCAN_BLOCK);
    if (lane == 31) {
        if (B == 0) return DESC_INC;
        const unsigned long long d = ld_desc(W.bdesc + (B - 1));
        return (d & DESC_INC) ? d : 0ull;
    }
    return lane < r ? ld_desc(W.desc + B * SCAN_BLOCK + lane) : DESC_AGG;
}
__device__ __forceinline__ void scan_resolve(const Enc2Work &W, uint64_t g, uint32_t lane, unsigned long long d)
{
    while (__any_sync(0xFFFFFFFFu, d == 0)) d = scan_peek(W, g, lane);
    const unsigned long long excl = warp_sum64(d & DESC_VAL);
    if (lane == 0) W.group_start[g] = excl;
}

// the 16 symbols of lane `lane` of unit `unit`; 0x10000 = no symbol (past the end)
__device__ __forceinline__ void load_unit(const uint8_t *in_bytes, uint64_t n_sym, bool aligned, uint64_t unit, uint32_t lane,
                                          uint32_t (&sym)[16])
{
    const uint64_t s0 = unit * UNIT_SYMS + lane * 16;
    if (aligned && s0 + 16 <= n_sym) {
        const uint4 a = ld_stream_v4(in_bytes + s0 * 2), b = ld_stream_v4(in_bytes + s0 * 2 + 16);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) { sym[2 * i] = w[i] & 0xFFFFu; sym[2 * i + 1] = w[i] >> 16; }
    } else {
        const uint16_t *in16 = reinterpret_cast<const uint16_t *>(in_bytes);
#pragma unroll
        for (int j = 0; j < 16; j++) sym[j] = s0 + j < n_sym ? (uint32_t)in16[s0 + j] : 0x10000u;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BITS_THREADS, 3)
enc_bits_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb, Enc2Work W,
                uint64_t ngroups, const ShardPlan *__restrict__ plan)
{
    extern __shared__ __align__(16) uint8_t s_len[];           // lenf plane, 64 KiB
    if (plan && plan->status) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t len_a = (uint32_t)__cvta_generic_to_shared(s_len);
    cta_bulk_load(len_a, cb->lenf, NSYM, len_a + NSYM);        // one bulk copy; the mbarrier sits behind the plane
    const bool aligned = ((uintptr_t)in_bytes & 15) == 0, wide = ((uintptr_t)in_bytes & 31) == 0;
    unsigned long long pend_g = 0;                                  // counted and published, its start not summed up yet
    bool have_pend = false;
    const uint64_t nblocks = (ngroups + SCAN_BLOCK - 1) / SCAN_BLOCK;
    for (;;) {
        // the ticket: groups START in ascending order, and a group is being counted from the moment it is handed out.
        // (Taken a few units ahead to hide the atomic's latency, a warp that is held up hands its delay on to the
        // leader that waits for its next group, who hands it on in turn: 1.07 -> 1.65 ms on 4 GiB.)
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(W.bdesc + nblocks, 1ull);
        g = __shfl_sync(0xFFFFFFFFu, g, 0);
        if (g >= ngroups) break;
        unsigned long long pend_d = 0;                              // my descriptor of the pending group's look-back, read
                                                                    // while the last units of this group are counted
        uint32_t mine = 0;
        if (aligned && (g + 1) * GROUP_SYMS <= n_sym) {
            // a whole group of an aligned input: the next unit's symbols are loaded while this one's lengths are
            // looked up, and there is no per-symbol end test (a predicated lookup makes the compiler rebuild the
            // shared base address for every symbol)
            const uint8_t *src = in_bytes + (g * GROUP_SYMS + lane * 16) * 2;
            uint4 a, b;
            ld_stream_2v4(src, wide, a, b);
#pragma unroll 4
            for (uint32_t u = 0; u < GROUP_UNITS; u++) {
                uint4 na = a, nb = b;
                if (u + 1 < GROUP_UNITS) ld_stream_2v4(src + (u + 1) * (UNIT_SYMS * 2), wide, na, nb);
                if (u == GROUP_UNITS - 4) {
                    if (have_pend) pend_d = scan_peek(W, pend_g, lane);
                }
                const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                uint32_t t = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) { const uint32_t f2 = fold16x2(w8[i]); t += lds8(len_a + (f2 & 0xFFFFu)) + lds8(len_a + (f2 >> 16)); }
                t = __reduce_add_sync(0xFFFFFFFFu, t);
                if (lane == u) mine = t;
                a = na; b = nb;
            }
        } else {
            if (have_pend) pend_d = scan_peek(W, pend_g, lane);
            for (uint32_t u = 0; u < GROUP_UNITS; u++) {
                const uint64_t unit = g * GROUP_UNITS + u;
                uint32_t t = 0;
                if ((unit + 1) * UNIT_SYMS <= n_sym) {
                    uint32_t sym[16];
                    load_unit(in_bytes, n_sym, aligned, unit, lane, sym);
#pragma unroll
                    for (int j = 0; j < 16; j++) t += lds8(len_a + fold16(sym[j]));
                    t = __reduce_add_sync(0xFFFFFFFFu, t);
                } else if (unit * UNIT_SYMS < n_sym) {
                    uint32_t sym[16];
                    load_unit(in_bytes, n_sym, aligned, unit, lane, sym);
#pragma unroll
                    for (int j = 0; j < 16; j++) t += sym[j] > 0xFFFFu ? 0u : (uint32_t)s_len[fold16(sym[j])];
                    t = __reduce_add_sync(0xFFFFFFFFu, t);
                }
                if (lane == u) mine = t;
            }
        }
        W.unit_bits[g * GROUP_UNITS + lane] = mine;
        const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, mine);
        scan_publish(W, g, ngroups, tot, lane);
        if (have_pend) scan_resolve(W, pend_g, lane, pend_d);
        pend_g = g;
        have_pend = true;
    }
    if (have_pend) scan_resolve(W, pend_g, lane, scan_peek(W, pend_g, lane));
}

// ---------------------------------------------------------------------------------------------
// predicated shared-memory OR without a branch
__device__ __forceinline__ void red_or_if(bool p, uint32_t saddr, uint32_t v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.shared.or.b32 [%1], %2;\n\t}"
                 :: "r"((uint32_t)p), "r"(saddr), "r"(v) : "memory");
}

// position of the leading one (0xFFFFFFFF for 0): one FLO instead of FLO + a subtraction from 31
__device__ __forceinline__ uint32_t bfind32(uint32_t x)
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

// stores w0 at shared address `a` when nf >= 32 and w1 behind it when nf >= 64, without a branch
__device__ __forceinline__ void sts_if_full(uint32_t nf, uint32_t a, uint32_t w0, uint32_t w1)
{
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ge.u32 p, %0, 32;\n\tsetp.ge.u32 q, %0, 64;\n\t"
                 "@p st.shared.u32 [%1], %2;\n\t@q st.shared.u32 [%1+4], %3;\n\t}"
                 :: "r"(nf), "r"(a), "r"(w0), "r"(w1) : "memory");
}

// (len, code) of one symbol from the shared planes, or from the global codebook when it is not there
__device__ __forceinline__ void lookup_any(const uint16_t *p16, const uint8_t *p8, const Codebook *cb, uint32_t sym,
                                           uint32_t &len, unsigned long long &code)
{
    const uint32_t f = fold16(sym);
    const uint32_t x = (uint32_t)p16[f] | ((uint32_t)p8[f] << 16);
    if (x) { len = 31 - __clz(x); code = x ^ (1u << len); }
    else { len = cb->len[sym]; code = cb->code[sym]; }
}

// One lane completes a partial word — `have` bits of it are there — with the codes of the symbols from sx on (any code
// length; the input may end first).
__device__ __forceinline__ uint32_t complete_word(const uint16_t *p16, const uint8_t *p8, const Codebook *cb, const uint8_t *in_bytes,
                                                  uint64_t n_sym, uint64_t sx, uint32_t have, uint32_t word)
{
    const uint16_t *in16 = reinterpret_cast<const uint16_t *>(in_bytes);
    for (; have < 32 && sx < n_sym; sx++) {
        uint32_t len;
        unsigned long long code;
        lookup_any(p16, p8, cb, in16[sx], len, code);
        if (len) {
            const unsigned long long left = code << (64 - len);    // left aligned
            word |= (uint32_t)(left >> 32) >> have;
            have += len;
        }
    }
    return word;
}

// OR `len` bits of `code` (right aligned) into the window at window bit `pos`; words outside [0, E2_WIN) of the
// window that starts at window word `wbase` are dropped.  Any length up to 64.
__device__ __forceinline__ void put_code_slow(uint32_t sbase, uint32_t wbase, uint32_t pos, unsigned long long code,
                                              uint32_t len)
{
    while (len) {
        const uint32_t w = pos >> 5, sh = pos & 31;
        const uint32_t take = min(len, 32u - sh);
        const uint32_t bits = (uint32_t)(code >> (len - take)) & (take == 32 ? 0xFFFFFFFFu : ((1u << take) - 1u));
        const uint32_t i = w - wbase;
        if (i < E2_WIN) red_or_if(true, sbase + 4u * i, bits << (32 - sh - take));
        pos += take;
        len -= take;
    }
}

// what a warp needs to pack a unit
struct UnitCtx {
    const uint8_t *in_bytes;
    uint64_t n_sym, nunits;
    const Codebook *cb;
    const uint16_t *p16;
    const uint8_t *p8;
    uint32_t *stage;
    uint32_t sbase;
    uint8_t *frame;
    unsigned long long bit0;
    bool aligned;
};

// One unit the general way: ragged or unaligned input, the stream head and tail, codes longer than 23 bits, units
// whose bits exceed the staging window.  (The common case is inlined in the kernel.)
__device__ __noinline__ void encode_unit_general(const UnitCtx &C, uint64_t unit, uint32_t bits, unsigned long long gbit,
                                                 uint32_t lane)
{
    const uint16_t *p16 = C.p16;
    const uint8_t *p8 = C.p8;
    const Codebook *cb = C.cb;
    uint32_t *stage = C.stage;
    const uint32_t sbase = C.sbase;
    uint8_t *frame = C.frame;
    uint32_t *gw = reinterpret_cast<uint32_t *>(frame);
    const uint16_t *in16 = reinterpret_cast<const uint16_t *>(C.in_bytes);
    const uint64_t n_sym = C.n_sym;
    const unsigned long long bit0 = C.bit0;

    const unsigned long long gend = gbit + bits;
    const uint32_t phase = (uint32_t)(gbit & 127);
    const unsigned long long G0 = (gbit - phase) >> 5;      // frame word of window word 0 (multiple of 4)
    const bool head = unit == 0, last_unit = unit + 1 == C.nunits;
    const uint32_t own_lo = head ? (phase >> 5) : ((phase + 31) >> 5);
    const uint32_t end_rel = phase + bits;                  // window bit after my last bit
    const uint32_t own_hi = (end_rel - 1) >> 5;             // bits > 0 here

    uint32_t sym[16];
    load_unit(C.in_bytes, n_sym, C.aligned, unit, lane, sym);
    uint32_t L = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) L += sym[j] > 0xFFFFu ? 0u : (uint32_t)cb->len[sym[j]];
    uint32_t off = L;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, off, o); if (lane >= o) off += y; }
    off -= L;                                               // unit bits before my first symbol

    const uint32_t npass = own_hi / E2_WIN + 1;
    for (uint32_t pass = 0; pass < npass; pass++) {
        const uint32_t wbase = pass * E2_WIN;
        const uint32_t used = min(E2_WIN, own_hi + 1 - wbase);         // words of the window that will be used
        for (uint32_t i = lane; i < (used + 3) / 4; i += 32) reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        {
            uint32_t pos = phase + off;
#pragma unroll 1
            for (int j = 0; j < 16; j++) {
                uint32_t sj = sym[0];
#pragma unroll
                for (int k = 1; k < 16; k++) sj = (k == j) ? sym[k] : sj;  // register select: sym[] stays out of local memory
                if (sj > 0xFFFFu) continue;
                uint32_t len;
                unsigned long long code;
                lookup_any(p16, p8, cb, sj, len, code);
                put_code_slow(sbase, wbase, pos, code, len);
                pos += len;
            }
        }
        __syncwarp();

        // ---- the stream head and my last, partial word (lane 0) ----
        const bool last_pass = pass + 1 == npass;
        unsigned long long fin = NOT_FINAL;
        if (lane == 0) {
            if (pass == 0 && head) {
                // preserve the bits of the first byte that precede the start phase
                const uint32_t b = frame[gbit >> 3];
                const uint32_t keep = b & ~(0xFFu >> (gbit & 7));
                stage[phase >> 5] |= keep << (24 - 8 * (uint32_t)((gbit >> 3) & 3));
            }
            if (last_pass) {
                if (last_unit) fin = gend;
                uint32_t have = end_rel & 31;               // bits of my last word that are mine
                if (have && !last_unit) {
                    // complete the word with the codes of the symbols that follow the unit
                    uint32_t word = 0;
                    uint64_t sx = (unit + 1) * UNIT_SYMS;
                    unsigned long long end = gend;
                    while (have < 32 && sx < n_sym) {
                        uint32_t len;
                        unsigned long long code;
                        lookup_any(p16, p8, cb, in16[sx], len, code);
                        if (len) {
                            const unsigned long long left = code << (64 - len);    // left aligned
                            word |= (uint32_t)(left >> 32) >> have;
                            have += len;
                            end += len;
                        }
                        sx++;
                    }
                    stage[own_hi - wbase] |= word;
                    if (have < 32) fin = end;               // the input ended inside my word: it is the last one
                }
            }
        }
        fin = __shfl_sync(0xFFFFFFFFu, fin, 0);
        __syncwarp();

        // ---- my words of this window leave: 128-bit stores where a whole group is mine, else words / bytes ----
        const unsigned long long W0 = G0 + wbase;
        const uint32_t hi_here = min(own_hi, wbase + E2_WIN - 1) - wbase;      // last window word to store
        const uint32_t lo_here = own_lo > wbase ? own_lo - wbase : 0u;
        for (uint32_t q = (lo_here >> 2) + lane; q <= (hi_here >> 2); q += 32) {
            uint4 o = reinterpret_cast<const uint4 *>(stage)[q];
            o.x = bswap32(o.x); o.y = bswap32(o.y); o.z = bswap32(o.z); o.w = bswap32(o.w);
            const uint32_t w0 = 4 * q;                      // window word of o.x
            const bool whole = w0 >= lo_here && w0 + 3 <= hi_here && !(head && pass == 0 && w0 <= (phase >> 5)) &&
                               !(fin != NOT_FINAL && w0 + 3 + wbase >= own_hi);
            if (whole) {
                st_stream_v4(gw + W0 + w0, o);
            } else {
                const uint32_t vv[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t wk = w0 + k;
                    if (wk < lo_here || wk > hi_here) continue;
                    const unsigned long long fw = W0 + wk;      // frame word
                    uint32_t b_lo = 0, b_hi = 4;                // byte range [b_lo, b_hi) of this word to store
                    if (fw == (bit0 >> 5)) b_lo = (uint32_t)((bit0 >> 3) & 3);         // bytes before the stream are not ours
                    if (fin != NOT_FINAL && wk + wbase == own_hi) b_hi = (uint32_t)(((fin - 1) >> 3) & 3) + 1;   // bytes holding bits
                    if (b_lo == 0 && b_hi == 4) gw[fw] = vv[k];
                    else
                        for (uint32_t b = b_lo; b < b_hi; b++)
                            frame[fw * 4 + b] = (uint8_t)(vv[k] >> (8 * b));    // little-endian view of the swapped word
                }
            }
        }
        __syncwarp();                                       // the window is reused
        // leave it zero, as the inlined path expects to find it
        for (uint32_t i = lane; i < (used + 3) / 4; i += 32) reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(E2_THREADS, 1)
encode2_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb, uint8_t *stream,
               uint64_t start_bit, Enc2Work W, uint64_t ngroups, const ShardPlan *__restrict__ plan)
{
    extern __shared__ __align__(16) uint8_t e2_smem[];
    if (plan) {                                         // the start bit is the device's: sizes never visit the host
        if (plan->status) return;
        stream += plan->local_start_bit >> 3;
        start_bit = plan->local_start_bit & 7;
    }
    const uint16_t *p16 = reinterpret_cast<const uint16_t *>(e2_smem);
    const uint8_t *p8 = e2_smem + NSYM * 2;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t *stage = reinterpret_cast<uint32_t *>(e2_smem + E2_PLANE_BYTES) + wid * E2_WIN;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
    // a warp's window is zero between units: the flush of a unit puts zeros back where it read (one store per vector
    // that leaves anyway, instead of a zeroing loop and a warp barrier before every pack)
    for (uint32_t i = lane; i < E2_WIN / 4; i += 32) reinterpret_cast<uint4 *>(stage)[i] = make_uint4(0, 0, 0, 0);
    {   // p16 | p8 (192 KiB) by one bulk copy
        const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(e2_smem);
        cta_bulk_load(a0, cb->p16, E2_PLANE_BYTES, a0 + E2_BAR);
    }

    // aligned frame: bit 0 of the frame is the 16-byte boundary at or below `stream`
    uint8_t *frame = reinterpret_cast<uint8_t *>((uintptr_t)stream & ~(uintptr_t)15);
    uint32_t *gw = reinterpret_cast<uint32_t *>(frame);
    const unsigned long long bit0 = ((uintptr_t)stream & 15) * 8ull + start_bit;   // first payload bit, frame coordinates
    const bool aligned = ((uintptr_t)in_bytes & 15) == 0, wide = ((uintptr_t)in_bytes & 31) == 0;
    const uint64_t nunits = (n_sym + UNIT_SYMS - 1) / UNIT_SYMS;
    const UnitCtx C{in_bytes, n_sym, nunits, cb, p16, p8, stage, sbase, frame, bit0, aligned};

    for (uint64_t g = (uint64_t)blockIdx.x * E2_WARPS + wid; g < ngroups; g += (uint64_t)gridDim.x * E2_WARPS) {
        const unsigned long long gstart = bit0 + W.group_start[g];
        const uint32_t ub = W.unit_bits[g * GROUP_UNITS + lane];
        uint32_t ux = ub;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, ux, o); if (lane >= o) ux += y; }
        ux -= ub;                                               // bits of the group before unit `lane`

        uint4 cur0 = make_uint4(0, 0, 0, 0), cur1 = cur0;       // my 16 symbols of the unit at hand (when have_cur)
        bool have_cur = false;
        // 32-bit bounds for the unit loop (64-bit unit numbers in it cost registers the kernel does not have):
        // units of this group, units whose successor is whole too, my lane's symbols of the group's first unit
        const uint64_t unit0 = g * GROUP_UNITS;
        const uint32_t nu = (uint32_t)min((uint64_t)GROUP_UNITS, nunits - unit0);
        const uint64_t whole = n_sym / UNIT_SYMS;               // whole units of the input
        const uint32_t fast_end = whole > unit0 + 1 ? (uint32_t)min((uint64_t)GROUP_UNITS, whole - 1 - unit0) : 0u;
        const uint8_t *gin = in_bytes + (unit0 * UNIT_SYMS + lane * 16) * 2;
        const bool first_group = g == 0;
        uint32_t slow = 0;                                      // units of the group that take the general path
        // Between two consecutive units of the common case the last, partial word of the first is CARRIED: lane 0 takes
        // it out of the window and starts the next unit's first word with it, so the word leaves with the unit that
        // holds its last bit.  Everywhere else (the group's ends, units of the general path) a word belongs to the unit
        // that holds its FIRST bit, which completes it with the codes that follow.
        bool carry_in = false;                                  // warp-uniform: lane 0 holds the word before my first bit in cw
        uint32_t cw = 0;
        for (uint32_t u = 0; u < nu; u++) {
            const uint32_t bits = __shfl_sync(0xFFFFFFFFu, ub, u);
            const unsigned long long gbit = gstart + __shfl_sync(0xFFFFFFFFu, ux, u);  // unit's first bit, frame coordinates
            const uint32_t phase = (uint32_t)(gbit & 127);
            // words I own: those whose first bit is mine (the first unit also owns the word the stream starts in, a unit
            // with a carried word the word it starts in)
            const bool first_unit = first_group && u == 0;
            const uint32_t own_lo = (first_unit || carry_in) ? (phase >> 5) : ((phase + 31) >> 5);
            const uint32_t end_rel = phase + bits;                  // window bit after my last bit
            uint32_t own_hi = bits ? ((end_rel - 1) >> 5) : 0u;
            bool fast = bits > 0 && own_hi >= own_lo;               // (else my bits sit in a word the unit before me completes)
            // the common case: whole units of an aligned input, this one and the next, away from the stream's ends
            fast = fast && aligned && !first_unit && u < fast_end && end_rel <= E2_WIN * 32;
            // will my last word be carried?  the next unit must be one of the common case too (but for a code longer than
            // 23 bits, which only its look-ups show: it completes the carried word the slow way then)
            const uint32_t have0 = end_rel & 31;                    // bits of my last word that are mine
            bool carry_out = false;
            if (have0 && u + 1 < nu && u + 1 < fast_end) {
                const uint32_t nbits = __shfl_sync(0xFFFFFFFFu, ub, u + 1);
                carry_out = nbits >= 64 && (end_rel & 127u) + nbits <= E2_WIN * 32;
            }
            uint32_t v[16];
            uint32_t L = 0;
            uint32_t tail_sym = 0;
            if (fast) {
                if (!have_cur) {
                    const uint8_t *src = gin + u * (UNIT_SYMS * 2);
                    ld_stream_2v4(src, wide, cur0, cur1);
                }
                const uint32_t w8[8] = {cur0.x, cur0.y, cur0.z, cur0.w, cur1.x, cur1.y, cur1.z, cur1.w};
                uint32_t zero = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t f2 = fold16x2(w8[j >> 1]);
                    const uint32_t f = (j & 1) ? (f2 >> 16) : (f2 & 0xFFFFu);
                    const uint32_t x = (uint32_t)p16[f] | ((uint32_t)p8[f] << 16);
                    v[j] = x;
                    zero = min(zero, x);
                    L += bfind32(x);                    // (x = 0: the unit takes the general path and L is not used)
                }
                if (__any_sync(0xFFFFFFFFu, zero == 0)) fast = false;   // a code longer than 23 bits
                if (fast) {
                    // my symbols are codes now: their registers take the next unit's symbols, which arrive while this
                    // unit is packed; the next unit's first 16, one per lane, complete my last word when it is not carried
                    const uint8_t *src = gin + (u + 1) * (UNIT_SYMS * 2);
                    ld_stream_2v4(src, wide, cur0, cur1);
                    if (!carry_out) tail_sym = *reinterpret_cast<const uint16_t *>(src - (int)(lane * 32) + (int)((lane & 15) * 2));
                }
            }
            if (!fast) {                                            // left for the loop below: a call in this loop makes
                if (carry_in) {                                     // the compiler keep the loop's state in local memory
                    // the word carried to me after all belongs to the unit before me: completed by one lane, from the input
                    if (lane == 0)
                        gw[gbit >> 5] = bswap32(complete_word(p16, p8, cb, in_bytes, n_sym, (unit0 + u) * UNIT_SYMS, phase & 31u, cw));
                    carry_in = false;
                }
                if (bits > 0 && own_hi >= ((first_unit ? phase : phase + 31) >> 5)) slow |= 1u << u;
                have_cur = false;
                continue;
            }
            uint32_t off = L;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, off, o); if (lane >= o) off += y; }
            off -= L;                                               // unit bits before my first symbol

            // ---- pack (the window is zero: see the flush below).  A lane strings its 16 codes together in registers:
            // pairs of codes are appended to the bits pending in `acc` (`fill` of them, left aligned; the leading
            // `pos & 31` bits of a lane's first word belong to the lanes before it and stay zero — or are the carried
            // word's, in lane 0), and every word that fills up leaves with a plain store — the lane that holds a word's
            // LAST bit is the only one that stores it.  What is left at the end (less than a word) is OR-ed in after the
            // stores: one shared-memory atomic per lane and unit instead of three per pair of codes. ----
            {
                const uint32_t pos = phase + off;
                uint32_t fill = pos & 31, acc = carry_in ? cw : 0u;     // (cw is zero in every lane but lane 0)
                uint32_t wa = sbase + 4u * (pos >> 5);
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const uint32_t x0 = v[j], x1 = v[j + 1];
                    const uint32_t l0 = bfind32(x0), l1 = bfind32(x1);
                    const uint32_t a32 = __funnelshift_lc(0u, x0, 32 - l0);   // left aligned, the leading one falls off
                    const uint32_t b32 = __funnelshift_lc(0u, x1, 32 - l1);
                    const uint32_t hi = a32 | (b32 >> l0);                    // l0 <= 23
                    const uint32_t lo = __funnelshift_r(0u, b32, l0);
                    const uint32_t nf = fill + l0 + l1;                       // < 32 + 46
                    const uint32_t w0 = acc | (hi >> fill);
                    const uint32_t w1 = __funnelshift_r(lo, hi, fill);
                    const uint32_t w2 = __funnelshift_r(0u, lo, fill);
                    sts_if_full(nf, wa, w0, w1);
                    acc = nf >= 64 ? w2 : (nf >= 32 ? w1 : w0);
                    wa += (nf >> 5) << 2;
                    fill = nf & 31;
                }
                __syncwarp();
                red_or_if(fill != 0, wa, acc);
            }
            __syncwarp();
            // ---- my last, partial word: carried to the next unit, or completed with the codes that follow (the next
            // unit's first symbols): lanes 0-15 look one symbol up each and OR their code in where it starts inside the word ----
            cw = 0;
            if (carry_out) {
                if (lane == 0) { cw = stage[own_hi]; stage[own_hi] = 0; }
                own_hi--;                                           // (a unit of the common case has at least 16 words)
            } else if (have0) {
                uint32_t x = 0, l = 0;
                if (lane < 16) {
                    const uint32_t f = fold16(tail_sym);
                    x = (uint32_t)p16[f] | ((uint32_t)p8[f] << 16);
                    l = 31 - __clz(x | 1u);
                }
                uint32_t o = l;
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, o, d); if (lane >= d) o += y; }
                const uint32_t tot16 = __shfl_sync(0xFFFFFFFFu, o, 15);
                const bool odd = __any_sync(0xFFFFFFFFu, lane < 16 && x == 0);     // a code that is not in the shared table
                if (!odd && have0 + tot16 >= 32) {
                    const uint32_t at = have0 + o - l;              // where my code starts in the word
                    red_or_if(lane < 16 && at < 32, sbase + 4u * own_hi, (x << (32 - l)) >> at);    // the leading one falls off
                } else if (lane == 0) {
                    // long codes, or sixteen one- and two-bit codes that do not fill the word: one lane, from the input;
                    // the next unit is whole, so the word does fill up before the input ends
                    stage[own_hi] = complete_word(p16, p8, cb, in_bytes, n_sym, (unit0 + u + 1) * UNIT_SYMS, have0, stage[own_hi]);
                }
            }
            carry_in = carry_out;
            __syncwarp();
            // ---- my words leave: 128-bit stores where a whole group is mine, else single words ----
            const unsigned long long G0 = (gbit - phase) >> 5;      // frame word of window word 0 (multiple of 4)
            // (my first bits may sit in the word before own_lo, which the unit before me completes and stores: vector 0)
            if (lane == 0 && (own_lo >> 2) != 0) reinterpret_cast<uint4 *>(stage)[0] = make_uint4(0, 0, 0, 0);
            for (uint32_t q = (own_lo >> 2) + lane; q <= (own_hi >> 2); q += 32) {
                uint4 o = reinterpret_cast<const uint4 *>(stage)[q];
                reinterpret_cast<uint4 *>(stage)[q] = make_uint4(0, 0, 0, 0);
                o.x = bswap32(o.x); o.y = bswap32(o.y); o.z = bswap32(o.z); o.w = bswap32(o.w);
                const uint32_t w0 = 4 * q;
                if (w0 >= own_lo && w0 + 3 <= own_hi) {
                    st_stream_v4(gw + G0 + w0, o);
                } else {
                    if (w0 >= own_lo && w0 <= own_hi) gw[G0 + w0] = o.x;
                    if (w0 + 1 >= own_lo && w0 + 1 <= own_hi) gw[G0 + w0 + 1] = o.y;
                    if (w0 + 2 >= own_lo && w0 + 2 <= own_hi) gw[G0 + w0 + 2] = o.z;
                    if (w0 + 3 >= own_lo && w0 + 3 <= own_hi) gw[G0 + w0 + 3] = o.w;
                }
            }
            __syncwarp();                                           // the window is reused
            have_cur = true;
        }
        while (slow) {                                          // units are independent of each other: any order will do
            const uint32_t u = __ffs(slow) - 1;
            slow &= slow - 1;
            const uint32_t bits = __shfl_sync(0xFFFFFFFFu, ub, u);
            const unsigned long long gbit = gstart + __shfl_sync(0xFFFFFFFFu, ux, u);
            encode_unit_general(C, unit0 + u, bits, gbit, lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Side index: the format has no offset array, so the decoder normally finds the code word
// boundaries by self-synchronisation.  For streams WE write, the compressor can hand the decoder what its
// synchronisation kernels would compute: one u16 record per 256-bit subsequence of the payload frame,
// (offset of the first code word boundary at or after the subsequence start) | (code words starting in it) << 6.
// A unit writes the records of the subsequences that end inside it with plain stores and adds its share of the two
// it shares with its neighbours with one atomic each (the records are zeroed first).  Reads N, writes C / 16.
constexpr uint32_t IDX_SUB_BITS = 256;                              // = SUB_BITS of the decoder
constexpr uint32_t IDX_SUBS_MAX = 160;                              // a unit spans <= 512 * 64 / 256 + 2 subsequences (+ 1)

__global__ void __launch_bounds__(BITS_THREADS)
enc_index_kernel(const uint8_t *__restrict__ in_bytes, uint64_t n_sym, const Codebook *__restrict__ cb, Enc2Work W,
                 uint64_t ngroups, unsigned long long bit0, uint16_t *__restrict__ rec, unsigned long long n_subs)
{
    extern __shared__ __align__(16) uint8_t s_len[];               // lenf plane, 64 KiB
    __shared__ uint32_t s_acc[BITS_THREADS / 32][IDX_SUBS_MAX];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        const uint32_t len_a = (uint32_t)__cvta_generic_to_shared(s_len);
        cta_bulk_load(len_a, cb->lenf, NSYM, len_a + NSYM);
    }
    const bool aligned = ((uintptr_t)in_bytes & 15) == 0, wide = ((uintptr_t)in_bytes & 31) == 0;
    const uint16_t *in16 = reinterpret_cast<const uint16_t *>(in_bytes);
    uint32_t *acc = s_acc[wid];
    const uint64_t warp0 = (uint64_t)blockIdx.x * (BITS_THREADS / 32) + wid;
    const uint64_t nwarps = (uint64_t)gridDim.x * (BITS_THREADS / 32);
    for (uint64_t g = warp0; g < ngroups; g += nwarps) {
        const unsigned long long gstart = bit0 + W.group_start[g];
        const uint32_t ub = W.unit_bits[g * GROUP_UNITS + lane];
        uint32_t ux = ub;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, ux, o); if (lane >= o) ux += y; }
        ux -= ub;
        uint32_t carry_len = 0;                                     // length of the last code word of the unit before
        uint4 nx0 = make_uint4(0, 0, 0, 0), nx1 = nx0;              // the next unit's symbols, loaded one unit ahead
        bool have_nx = false;
        for (uint32_t u = 0; u < GROUP_UNITS; u++) {
            const uint64_t unit = g * GROUP_UNITS + u;
            if (unit * UNIT_SYMS >= n_sym) break;
            const uint32_t bits = __shfl_sync(0xFFFFFFFFu, ub, u);
            const unsigned long long S = gstart + __shfl_sync(0xFFFFFFFFu, ux, u);     // frame bit of the unit's first code word
            uint32_t len[16];
            if (aligned && (unit + 2) * UNIT_SYMS <= n_sym && u + 1 < GROUP_UNITS) {
                // whole units of an aligned input: mine comes from the registers filled one step ago
                uint4 c0 = nx0, c1 = nx1;
                if (!have_nx) {
                    const uint8_t *src = in_bytes + (unit * UNIT_SYMS + lane * 16) * 2;
                    ld_stream_2v4(src, wide, c0, c1);
                }
                const uint8_t *nsrc = in_bytes + ((unit + 1) * UNIT_SYMS + lane * 16) * 2;
                ld_stream_2v4(nsrc, wide, nx0, nx1);
                have_nx = true;
                const uint32_t w8[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int j = 0; j < 16; j++)
                    len[j] = (uint32_t)s_len[fold16((j & 1) ? (w8[j >> 1] >> 16) : (w8[j >> 1] & 0xFFFFu))];
            } else if (have_nx) {
                // the last unit of the group (or of the whole units): its symbols are already here
                have_nx = false;
                const uint32_t w8[8] = {nx0.x, nx0.y, nx0.z, nx0.w, nx1.x, nx1.y, nx1.z, nx1.w};
#pragma unroll
                for (int j = 0; j < 16; j++)
                    len[j] = (uint32_t)s_len[fold16((j & 1) ? (w8[j >> 1] >> 16) : (w8[j >> 1] & 0xFFFFu))];
            } else {
                uint32_t sym[16];
                load_unit(in_bytes, n_sym, aligned, unit, lane, sym);
#pragma unroll
                for (int j = 0; j < 16; j++) len[j] = sym[j] > 0xFFFFu ? 0u : (uint32_t)s_len[fold16(sym[j])];
            }
            uint32_t L = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) L += len[j];
            uint32_t off = L;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, off, o); if (lane >= o) off += y; }
            off -= L;
            // length of the code word before my first one (the previous lane's last; lane 0: the previous unit's last)
            uint32_t plen = __shfl_up_sync(0xFFFFFFFFu, len[15], 1);
            if (lane == 0) plen = u ? carry_len : (unit ? (uint32_t)s_len[fold16(in16[unit * UNIT_SYMS - 1])] : 0u);
            carry_len = __shfl_sync(0xFFFFFFFFu, len[15], 31);      // (a ragged unit is the last one: never used then)
            if (bits == 0) continue;
            const unsigned long long sub0 = S / IDX_SUB_BITS;       // holds the unit's first bit
            const uint32_t rel0 = (uint32_t)(S - sub0 * IDX_SUB_BITS);              // positions below are relative to sub0
            const uint32_t nsub = (rel0 + bits - 1) / IDX_SUB_BITS + 1;
            // acc[i] = ((unit-relative index of the first code word of subsequence sub0 + i) << 8) | its bit offset + 1,
            // written by the one lane that holds that code word; 0 = none of my code words starts there
            for (uint32_t i = lane; i <= nsub; i += 32) acc[i] = 0;
            __syncwarp();
            {
                uint32_t pos = rel0 + off;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    // first code word of its subsequence: the one before it started in an earlier subsequence
                    // (a lane of a ragged last unit has len 0 from its first missing symbol on: never first)
                    const uint32_t inb = pos % IDX_SUB_BITS;
                    const bool first = len[j] != 0 && (inb < plen || (unit == 0 && lane == 0 && j == 0));
                    if (first) acc[pos / IDX_SUB_BITS] = ((lane * 16 + j) << 8) | ((inb & 63u) + 1u);
                    plen = len[j];
                    pos += len[j];
                }
            }
            __syncwarp();
            // Counts are differences of first-code-word indices.  Subsequences that end inside the unit are complete:
            // plain stores.  The first one when an earlier unit holds its first code word, and the last one (the units
            // that follow add theirs), are shared: 32-bit atomic adds on the zeroed word that holds the 16-bit record
            // (a count never carries into the neighbouring record).
            const uint32_t nsym_unit = (uint32_t)min((uint64_t)UNIT_SYMS, n_sym - unit * UNIT_SYMS);
            for (uint32_t i = lane; i < nsub; i += 32) {
                const uint32_t a = acc[i], nx = i + 1 < nsub ? acc[i + 1] : 0u;
                const unsigned long long sub = sub0 + i;
                if (sub >= n_subs) continue;
                const uint32_t end = nx ? (nx >> 8) : nsym_unit;    // (only the unit's last subsequence can be without a first)
                uint32_t v;
                if (a) v = ((a & 0xFFu) - 1u) | ((end - (a >> 8)) << 6);
                else if (i == 0) v = end << 6;                      // my code words before the first boundary: the owner's
                else continue;                                      // my last bits belong to a code word of the one before
                if (a == 0 || nx == 0)
                    atomicAdd(reinterpret_cast<uint32_t *>(rec + (sub & ~1ull)), v << (16 * (uint32_t)(sub & 1)));
                else
                    rec[sub] = (uint16_t)v;
            }
            __syncwarp();
        }
    }
}

// the workspace arrays of an encode of n_sym symbols (ctx->ws, behind the codebook workspace)
static int enc2_work(Ctx *c, uint64_t n_sym, Enc2Work *W, uint64_t *ngroups_out)
{
    const uint64_t ngroups = (n_sym + GROUP_SYMS - 1) / GROUP_SYMS;
    const size_t off = WS_STAGE_OFFSET;
    const size_t b_units = ((size_t)ngroups * GROUP_UNITS * 4 + 255) & ~(size_t)255;
    const size_t b_gstart = ((size_t)ngroups * 8 + 255) & ~(size_t)255;
    const uint64_t nblocks = (ngroups + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const size_t b_desc = ((size_t)(ngroups + nblocks + 1) * 8 + 255) & ~(size_t)255;
    int rc = ensure_ws(c, off + b_units + b_gstart + b_desc);
    if (rc) return rc;
    uint8_t *p = (uint8_t *)c->ws + off;
    W->unit_bits = reinterpret_cast<uint32_t *>(p); p += b_units;
    W->group_start = reinterpret_cast<unsigned long long *>(p); p += b_gstart;
    W->desc = reinterpret_cast<unsigned long long *>(p);
    W->bdesc = W->desc + ngroups;
    *ngroups_out = ngroups;
    return HF_OK;
}

// Records of the side index for the stream launch_encode has just packed with the same arguments (the unit and
// group bit counts are still in the workspace).  d_rec: n_subs u16 records, zeroed here.
int launch_encode_index(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb, const uint8_t *d_stream,
                        uint64_t start_bit, uint16_t *d_rec, uint64_t n_subs)
{
    const uint64_t n_sym = n_bytes / 2;
    HF_CUDA(c, cudaMemsetAsync(d_rec, 0, n_subs * 2, c->stream));
    if (n_sym == 0) return HF_OK;
    Enc2Work W;
    uint64_t ngroups;
    int rc = enc2_work(c, n_sym, &W, &ngroups);
    if (rc) return rc;
    if (!c->smem_attr[ATTR_INDEX]) {
        HF_CUDA(c, cudaFuncSetAttribute(enc_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BITS_SMEM));
        c->smem_attr[ATTR_INDEX] = true;
    }
    d_stream += start_bit >> 3;
    start_bit &= 7;
    const unsigned long long bit0 = ((uintptr_t)d_stream & 15) * 8ull + start_bit;     // frame bit of the first code word
    const uint64_t bw = BITS_THREADS / 32;
    uint64_t grid = (ngroups + bw - 1) / bw;
    if (grid > (uint64_t)(3 * c->sm_count)) grid = 3 * c->sm_count;
    HF_PROF(c, "enc_index_kernel"); enc_index_kernel<<<(unsigned)grid, BITS_THREADS, BITS_SMEM, c->stream>>>(d_in, n_sym, d_cb, W, ngroups, bit0, d_rec, n_subs);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_encode(Ctx *c, const uint8_t *d_in, uint64_t n_bytes, const Codebook *d_cb, uint8_t *d_stream,
                  uint64_t start_bit, const ShardPlan *plan)
{
    const uint64_t n_sym = n_bytes / 2;
    if (n_sym == 0) return HF_OK;
    if ((uintptr_t)d_in & 1) return set_err(c, HF_ERR_ARG, "hf_encode: input must be 2-byte aligned");
    if ((uintptr_t)d_cb & 15) return set_err(c, HF_ERR_ARG, "hf_encode: codebook must be 16-byte aligned");
    Enc2Work W;
    uint64_t ngroups;
    int rc = enc2_work(c, n_sym, &W, &ngroups);
    if (rc) return rc;

    if (!c->smem_attr[ATTR_ENCODE]) {
        HF_CUDA(c, cudaFuncSetAttribute(encode2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E2_SMEM));
        HF_CUDA(c, cudaFuncSetAttribute(enc_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BITS_SMEM));
        c->smem_attr[ATTR_ENCODE] = true;
    }
    // start_bit may exceed 8: fold whole bytes into the pointer
    d_stream += start_bit >> 3;
    start_bit &= 7;
    const uint64_t bw = BITS_THREADS / 32;
    uint64_t bgrid = (ngroups + bw - 1) / bw;
    if (bgrid > (uint64_t)(3 * c->sm_count)) bgrid = 3 * c->sm_count;
    HF_CUDA(c, cudaMemsetAsync(W.desc, 0, (ngroups + (ngroups + SCAN_BLOCK - 1) / SCAN_BLOCK + 1) * 8, c->stream));   // descriptors + the ticket
    HF_PROF(c, "enc_bits_kernel"); enc_bits_kernel<<<(unsigned)bgrid, BITS_THREADS, BITS_SMEM, c->stream>>>(d_in, n_sym, d_cb, W, ngroups, plan);
    HF_LAUNCH_CHECK(c);
    uint64_t grid = (ngroups + E2_WARPS - 1) / E2_WARPS;
    if (grid > (uint64_t)c->sm_count) grid = c->sm_count;
    HF_PROF(c, "encode2_kernel"); encode2_kernel<<<(unsigned)grid, E2_THREADS, E2_SMEM, c->stream>>>(d_in, n_sym, d_cb, d_stream, start_bit, W, ngroups, plan);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
