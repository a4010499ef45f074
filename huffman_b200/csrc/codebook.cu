// codebook.cu — codebook construction with the reference's exact tie-breaking, and the
// header bit-packer.
//
// Replaces: thrust count_if / sequence / sort_by_key and the host round trips
// (/root/reference/Compressor.cu:378-425), GenerateCL / GenerateCW and the host string
// building (gpuHuffmanConstruction.h:353-494, :551-579), and the per-byte fwrite header
// writer (Compressor.cu:427-487, :637-669).
//
// Result contract: leaves = non-zero bins in ascending (count, symbol);
// two-queue Huffman where the LEAF wins a tie against an internal node and older internal
// nodes precede newer ones; the first node of a pair is the left child and carries bit 1.
//
// Mechanism (ours): everything stays on the device, one stream, no host round trip.
//   1. cb_sort_tree_kernel, one 1024-thread CTA:
//        compaction of the non-zero bins (symbol order) -> stable LSD radix sort on the
//        count (8-bit digits, warp-private digit counters, match_any ranking), so equal
//        counts keep ascending symbol order without relying on a library sort;
//        then the tree by ROUNDS over two queues that never move data: per round all
//        nodes <= (sum of the two smallest) are paired in merged order, each pair found by
//        a merge-path co-rank search (leaf queue first on ties); the new internal nodes
//        come out already sorted.  Rounds ~ 20-50 for real data.
//   2. cb_codes_kernel, one thread per leaf: walk to the root, emit len / code / the encoder's planes,
//      reduce table_bits, payload_bits and maxlen.
//   3. cb_entry_scan_kernel: exclusive scan of the header entry sizes (24 + len).
//   4. header_pack_kernel, one thread per entry: OR the entry's bits into the image.
// Bytes moved are negligible (<= a few MB, L2 resident); the stage is latency bound and
// reported in microseconds, not against the HBM roofline.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace hf {

constexpr int CB_THREADS = 1024;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int SORT_BATCH = 8;                     // keys of a lane in a radix pass (tiles of 256 keys)
constexpr int CB_CLUSTER = 8;                     // CTAs of the sorting cluster (the portable maximum; the scan below assumes 8)

struct CbWork {
    unsigned long long keyA[NSYM], keyB[NSYM];      // counts (sort ping-pong); keyA ends as sorted leaf counts
    uint32_t valA[NSYM], valB[NSYM];                // symbols
    unsigned long long intF[NSYM];                  // internal node counts, creation order
    uint32_t leafPar[NSYM], intPar[NSYM];           // (parent internal index << 1) | is_left, NONE at the root
    uint32_t entry_bits[NSYM];
    uint32_t U;
    uint32_t rounds;
    // exchange areas of the sorting cluster
    alignas(16) uint32_t cta_cnt[256 * CB_CLUSTER];           // keys per (digit, CTA) of a radix pass
    uint32_t cta_nz[CB_CLUSTER];
    unsigned long long cta_max[CB_CLUSTER];
};

size_t cb_work_bytes() { return sizeof(CbWork); }
static_assert(sizeof(CbWork) <= WS_SCRATCH_BYTES, "CbWork lives in the workspace's scratch region");

// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *warp_sums, uint32_t *total)
{   // exclusive scan across the 1024-thread CTA; warp_sums: 33 words of smem
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = warp_sums[lane];
        uint32_t t = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, t, o);
            if (lane >= o) t += y;
        }
        warp_sums[lane] = t - s;                    // exclusive
        if (lane == 31) warp_sums[32] = t;
    }
    __syncthreads();
    uint32_t r = x - v + warp_sums[wid];
    if (total) *total = warp_sums[32];
    __syncthreads();
    return r;
}

// number of elements <= val in the ascending array a[lo, hi); whole CTA cooperates
__device__ __forceinline__ uint32_t block_count_le(const unsigned long long *a, uint32_t lo, uint32_t hi,
                                                   unsigned long long val)
{
    uint32_t L = hi - lo;
    if (L == 0) return 0;
    uint32_t seg = (L + CB_THREADS - 1) / CB_THREADS;
    uint32_t nseg = (L + seg - 1) / seg;
    uint32_t i = threadIdx.x;
    int pred = 0;
    if (i < nseg) {
        uint32_t last = lo + min(L, (i + 1) * seg) - 1;
        pred = a[last] <= val;
    }
    uint32_t full = __syncthreads_count(pred);      // segments entirely <= val (monotone)
    if (full == nseg) return L;
    uint32_t s0 = lo + full * seg;
    uint32_t slen = min(seg, hi - s0);
    pred = (i < slen) ? (a[s0 + i] <= val) : 0;
    uint32_t part = __syncthreads_count(pred);
    return full * seg + part;
}

// ---------------------------------------------------------------------------------
// A thread-block CLUSTER of CB_CLUSTER CTAs (one SM each) compacts and sorts; CTA 0 then builds the tree.  One SM
// stores about one 32-byte sector per cycle, and a radix pass scatters every key to a sector of its own: one CTA
// spent 0.21 of its 0.56 ms in the scatter stores of the four passes (clock64 per phase).  The CTAs exchange their
// digit counts through global memory (L2) and meet at the hardware cluster barrier, twice per pass.
__global__ void __cluster_dims__(CB_CLUSTER, 1, 1) __launch_bounds__(CB_THREADS, 1)
cb_sort_tree_kernel(const unsigned long long *__restrict__ hist, CbWork *w, Codebook *cb)
{
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_cnt[32][257];             // warp-private digit counters / offsets (padded: columns are read too)
    __shared__ uint32_t s_base[256];                // where this CTA's keys of a digit start
    __shared__ unsigned long long s_u64[4];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t cta = cluster.block_rank();
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#ifdef CB_TIMING
    long long tq0 = clock64(), tq1 = 0, tq2 = 0, tqc = 0, tqs = 0, tqx = 0, tqa;
#endif

    // ---- 1. compaction of the non-zero bins, symbol order (C:378-385, C:413-425) ----
    // CTA c takes bins [8192 c, 8192 c + 8192), coalesced: iteration j reads 1,024 of them, a warp's 32 bins give one
    // ballot mask; the masks' population counts are scanned in bin order, the CTA totals are exchanged, and the bins
    // are read once more (L2) to be written at CTA base + mask base + rank.
    constexpr uint32_t NIT = NSYM / CB_THREADS / CB_CLUSTER;    // 8
    uint32_t *s_mask = &s_cnt[0][0];                // [NIT * 32] nonzero masks, index j * 32 + wid = bin order
    uint32_t *s_pref = s_mask + NIT * 32;           // their exclusive prefix
    const uint32_t bin0 = cta * (NSYM / CB_CLUSTER);
    unsigned long long mx = 0;
    unsigned long long hv[NIT];
#pragma unroll
    for (uint32_t j = 0; j < NIT; j++) {
        hv[j] = hist[bin0 + j * CB_THREADS + tid];
        mx = hv[j] > mx ? hv[j] : mx;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, hv[j] != 0);
        if (lane == 0) s_mask[j * 32 + wid] = m;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long y = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
        mx = y > mx ? y : mx;
    }
    if (tid == 0) s_u64[0] = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_u64[0], mx);
    uint32_t nz_cta;
    {
        const uint32_t c0 = tid < NIT * 32 ? __popc(s_mask[tid]) : 0u;
        const uint32_t base = block_excl_scan_u32(c0, s_scan, &nz_cta);        // (contains CTA barriers)
        if (tid < NIT * 32) s_pref[tid] = base;
    }
    if (tid == 0) { w->cta_nz[cta] = nz_cta; w->cta_max[cta] = s_u64[0]; }
    cluster.sync();
    uint32_t U = 0, base_cta = 0;
    mx = 0;
#pragma unroll
    for (uint32_t c = 0; c < CB_CLUSTER; c++) {
        const uint32_t n = w->cta_nz[c];
        const unsigned long long m = w->cta_max[c];
        base_cta += c < cta ? n : 0u;
        U += n;
        mx = m > mx ? m : mx;
    }
#pragma unroll
    for (uint32_t j = 0; j < NIT; j++) {
        if (hv[j]) {
            const uint32_t pos = base_cta + s_pref[j * 32 + wid] + __popc(s_mask[j * 32 + wid] & ((1u << lane) - 1u));
            w->keyA[pos] = hv[j];
            w->valA[pos] = bin0 + j * CB_THREADS + tid;
        }
    }
    int passes = 0;                                 // 8-bit digit passes the largest count needs
    while (passes < 8 && (mx >> (8 * passes)) != 0) passes++;
    cluster.sync();                                 // the keys are complete
#ifdef CB_TIMING
    tq1 = clock64();
#endif

    // ---- 2. stable LSD radix sort by count (C:387-389 semantics) ----
    // 256 tiles (one per warp of the cluster) of `chunk` consecutive keys.  Stable order = digit, then tile: the CTAs
    // exchange their counts per digit (2,048 numbers through L2), every CTA scans them in (digit, CTA) order, and
    // the warps of a CTA share out the CTA's range of a digit in warp order; a warp then walks its tile in order.
    unsigned long long *kin = w->keyA, *kout = w->keyB;
    uint32_t *vin = w->valA, *vout = w->valB;
    constexpr uint32_t NTILE = CB_CLUSTER * 32;
    const uint32_t chunk = (((U + NTILE - 1) / NTILE) + 31) & ~31u;        // keys per tile, a multiple of 32 (<= 256)
    const uint32_t tile = cta * 32 + wid;
    const uint32_t w_lo = min(U, tile * chunk), w_hi = min(U, (tile + 1) * chunk);
    for (int p = 0; p < passes; p++) {
        const int sh = 8 * p;
        for (uint32_t i = tid; i < 32 * 257; i += CB_THREADS) (&s_cnt[0][0])[i] = 0;
        __syncthreads();
#ifdef CB_TIMING
        tqa = clock64();
#endif
        unsigned long long k8[SORT_BATCH];          // my keys of the tile (<= 8 per lane), kept for the scatter
        uint32_t v8[SORT_BATCH];
#pragma unroll
        for (int b = 0; b < SORT_BATCH; b++) {
            const uint32_t i = w_lo + 32 * b + lane;
            k8[b] = i < w_hi ? kin[i] : 0ull;
            v8[b] = i < w_hi ? vin[i] : 0u;
        }
#pragma unroll
        for (int b = 0; b < SORT_BATCH; b++)
            if (w_lo + 32 * b + lane < w_hi) atomicAdd(&s_cnt[wid][(uint32_t)(k8[b] >> sh) & 255u], 1u);
        __syncthreads();
        if (tid < 256) {
            uint32_t tot = 0;
#pragma unroll 8
            for (int q = 0; q < 32; q++) tot += s_cnt[q][tid];
            w->cta_cnt[tid * CB_CLUSTER + cta] = tot;
        }
        cluster.sync();
#ifdef CB_TIMING
        tqc += clock64() - tqa; tqa = clock64();
#endif
        {
            // entries 2 tid and 2 tid + 1 of the (digit, CTA) order: digit tid / 4, CTAs 2 (tid % 4) and + 1
            const uint2 e = reinterpret_cast<const uint2 *>(w->cta_cnt)[tid];
            const uint32_t run = block_excl_scan_u32(e.x + e.y, s_scan, nullptr);
            if (cta == 2 * (tid & 3u)) s_base[tid >> 2] = run;
            else if (cta == 2 * (tid & 3u) + 1) s_base[tid >> 2] = run + e.x;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t run = s_base[tid];
#pragma unroll 8
            for (int q = 0; q < 32; q++) { const uint32_t c = s_cnt[q][tid]; s_cnt[q][tid] = run; run += c; }
        }
        __syncthreads();
#ifdef CB_TIMING
        tqs += clock64() - tqa; tqa = clock64();
#endif
        // scatter, each warp walking its keys in order
#pragma unroll
        for (int b = 0; b < SORT_BATCH; b++) {
            const uint32_t i = w_lo + 32 * b + lane;
            if (w_lo + 32 * b >= w_hi) break;       // uniform
            const bool act = i < w_hi;
            const unsigned long long k = k8[b];
            const uint32_t d = (uint32_t)(k >> sh) & 255u;
            // lanes with my digit: eight independent ballots; the first of them advances the warp's private counter
            // and hands out the base
            uint32_t peers = __ballot_sync(0xFFFFFFFFu, act);
#pragma unroll
            for (int bit = 0; bit < 8; bit++) {
                const uint32_t bl = __ballot_sync(0xFFFFFFFFu, (d >> bit) & 1u);
                peers &= ((d >> bit) & 1u) ? bl : ~bl;
            }
            const uint32_t rank = __popc(peers & ((1u << lane) - 1));
            uint32_t base = 0;
            if (act && rank == 0) base = atomicAdd(&s_cnt[wid][d], __popc(peers));
            base = __shfl_sync(0xFFFFFFFFu, base, act ? __ffs(peers) - 1 : lane);
            if (act) {
                kout[base + rank] = k;
                vout[base + rank] = v8[b];
            }
        }
        cluster.sync();                             // every key of the pass is in place
#ifdef CB_TIMING
        tqx += clock64() - tqa;
#endif
        unsigned long long *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    // publish the order; keep the sorted counts in kin
    for (uint32_t i = cta * CB_THREADS + tid; i < U; i += CB_CLUSTER * CB_THREADS) {
        cb->order[i] = (uint16_t)vin[i];
        w->leafPar[i] = NONE;
        w->intPar[i] = NONE;
    }
    cluster.sync();
    if (cta != 0) return;                           // the tree is one CTA's
    const unsigned long long *leafF = kin;
    if (tid == 0) {
        w->U = U;
        cb->U = U;
        // record which ping-pong buffer holds the sorted leaves for cb_codes_kernel
        w->rounds = (kin == w->keyA) ? 0u : 0x80000000u;
    }
    __syncthreads();

#ifdef CB_TIMING
    tq2 = clock64();
#endif
    // ---- 3. tree by rounds over two queues (h:353-466 result) ----
    unsigned long long *intF = w->intF;
    uint32_t l = 0, h = 0, t = 0;                   // leaf head, internal head, internal tail (uniform)
    uint32_t rounds = 0;
    while ((U - l) + (t - h) > 1) {
        if (tid == 0) {
            // the two smallest of the merged fronts, leaf first on ties
            unsigned long long sum = 0;
            uint32_t ll = l, hh = h;
            for (int j = 0; j < 2; j++) {
                bool takeLeaf = (hh == t) || (ll < U && leafF[ll] <= intF[hh]);
                if (takeLeaf) sum += leafF[ll++]; else sum += intF[hh++];
            }
            s_u64[1] = sum;
        }
        __syncthreads();
        const unsigned long long spec = s_u64[1];
        uint32_t lc = block_count_le(leafF, l, U, spec);
        uint32_t ic = block_count_le(intF, h, t, spec);
        uint32_t m = lc + ic;
        if (m & 1) {
            // leave the LAST element of the merged order for the next round:
            // an internal node when its count >= the last leaf's (leaf first on ties)
            bool lastIsInt = (lc == 0) || (ic > 0 && intF[h + ic - 1] >= leafF[l + lc - 1]);
            if (lastIsInt) ic--; else lc--;
            m--;
        }
        const uint32_t pairs = m >> 1;
        if (pairs == 0) {                           // cannot happen for sorted inputs; never spin on the GPU
            if (tid == 0) cb->status = HF_ERR_INTERNAL;
            break;
        }
        const unsigned long long *A = leafF + l;    // list A: leaves (first on ties)
        const unsigned long long *B = intF + h;     // list B: internal nodes
        for (uint32_t i = tid; i < pairs; i += CB_THREADS) {
            const uint32_t k = 2 * i;
            // co-rank: a = number of A elements among the first k of the merge
            uint32_t lo = k > ic ? k - ic : 0, hi = min(k, lc);
            while (lo < hi) {
                uint32_t a = (lo + hi) >> 1;        // try taking a from A, k - a from B
                // too few from A if B[k-a-1] > A[a]  (A element must come first when <=)
                if (A[a] <= B[k - a - 1]) lo = a + 1; else hi = a;
            }
            uint32_t a = lo, b = k - lo;
            uint32_t node[2];
            unsigned long long f = 0;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                bool takeA = (b >= ic) || (a < lc && A[a] <= B[b]);
                if (takeA) { node[j] = (l + a) | 0x80000000u; f += A[a]; a++; }
                else { node[j] = h + b; f += B[b]; b++; }
            }
            const uint32_t par = t + i;
            intF[par] = f;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                uint32_t link = (par << 1) | (j == 0 ? 1u : 0u);   // first of the pair = left = bit 1
                if (node[j] & 0x80000000u) w->leafPar[node[j] & 0x7FFFFFFFu] = link;
                else w->intPar[node[j]] = link;
            }
        }
        l += lc; h += ic; t += pairs;
        rounds++;
        __syncthreads();
    }
    if (tid == 0) w->rounds |= rounds;
#ifdef CB_TIMING
    if (tid == 0) printf("cb_sort_tree: U %u passes %d rounds %u | compaction %lld sort %lld (count %lld scan %lld scatter %lld) tree %lld clk\n", U, passes, rounds,
                         tq1 - tq0, tq2 - tq1, tqc, tqs, tqx, clock64() - tq2);
#endif
}

// ---------------------------------------------------------------------------------
__global__ void cb_codes_kernel(const unsigned long long *__restrict__ hist, CbWork *__restrict__ w,
                                Codebook *__restrict__ cb)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t U = w->U;
    unsigned long long tb = 0, pb = 0;
    uint32_t mlen = 0;
    if (k < U) {
        const uint32_t *symv = (w->rounds & 0x80000000u) ? w->valB : w->valA;
        const uint32_t sym = symv[k];
        unsigned long long code = 0;
        uint32_t len = 0;
        uint32_t v = w->leafPar[k];
        while (v != NONE) {
            if (len < 64) code |= (unsigned long long)(v & 1u) << len;
            len++;
            v = w->intPar[v >> 1];
        }
        if (len > 64) { atomicExch(&cb->status, (uint32_t)HF_ERR_CODE_TOO_LONG); len = 64; }
        cb->len[sym] = (uint8_t)len;
        cb->code[sym] = code;
        const uint32_t v24 = len <= 23 ? ((1u << len) | (uint32_t)code) : 0u;
        const uint32_t fsym = sym ^ (sym >> 8);                 // the encoder's bank-spreading index (encode.cu fold16)
        cb->p16[fsym] = (uint16_t)v24;
        cb->p8[fsym] = (uint8_t)(v24 >> 16);
        cb->lenf[fsym] = (uint8_t)len;
        w->entry_bits[k] = 24 + len;
        tb = 24 + len;
        pb = hist[sym] * len;
        mlen = len;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        tb += __shfl_xor_sync(0xFFFFFFFFu, tb, o);
        pb += __shfl_xor_sync(0xFFFFFFFFu, pb, o);
        mlen = max(mlen, __shfl_xor_sync(0xFFFFFFFFu, mlen, o));
    }
    if ((threadIdx.x & 31) == 0 && tb) {
        atomicAdd(&cb->table_bits, tb);
        atomicAdd(&cb->payload_bits, pb);
        atomicMax(&cb->maxlen, mlen);
    }
}

__global__ void __launch_bounds__(CB_THREADS, 1)
cb_entry_scan_kernel(const CbWork *__restrict__ w, uint32_t *__restrict__ entry_off)
{
    __shared__ uint32_t s_scan[40];
    const uint32_t U = w->U, tid = threadIdx.x;
    constexpr uint32_t per = NSYM / CB_THREADS;     // 64 consecutive entries per thread, as 16 independent 128-bit loads
    uint4 v[per / 4];
    const uint4 *src = reinterpret_cast<const uint4 *>(w->entry_bits + tid * per);
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < per / 4; j++) {
        const uint32_t k = tid * per + 4 * j;
        uint4 x = k < U ? src[j] : make_uint4(0, 0, 0, 0);          // entries past U hold stale values
        if (k + 1 >= U) x.y = 0;
        if (k + 2 >= U) x.z = 0;
        if (k + 3 >= U) x.w = 0;
        v[j] = x;
        sum += x.x + x.y + x.z + x.w;
    }
    uint32_t run = block_excl_scan_u32(sum, s_scan, nullptr);
    uint4 *dst = reinterpret_cast<uint4 *>(entry_off + tid * per);
#pragma unroll
    for (uint32_t j = 0; j < per / 4; j++) {
        const uint32_t k = tid * per + 4 * j;
        uint4 o;
        o.x = run; run += v[j].x;
        o.y = run; run += v[j].y;
        o.z = run; run += v[j].z;
        o.w = run; run += v[j].w;
        if (k < U) dst[j] = o;                                      // offsets past U are never read
    }
}

// ---------------------------------------------------------------------------------
// header image: bytes 0..2(3), then the MSB-first bit stream.
// The image was zeroed; every field is OR-ed in, 32-bit words in the aligned frame of d_file.
__device__ __forceinline__ void or_bits(uint32_t *words, unsigned long long bitpos, unsigned long long v,
                                        uint32_t nbits)
{   // append the low nbits (<= 64) of v at absolute bit position bitpos (MSB first)
    while (nbits) {
        unsigned long long wi = bitpos >> 5;
        uint32_t room = 32 - (uint32_t)(bitpos & 31);
        uint32_t take = nbits < room ? nbits : room;
        uint32_t chunk = (uint32_t)((v >> (nbits - take)) & (take == 32 ? 0xFFFFFFFFull : ((1ull << take) - 1)));
        atomicOr(&words[wi], bswap32(chunk << (room - take)));
        bitpos += take;
        nbits -= take;
    }
}

__global__ void header_pack_kernel(const Codebook *__restrict__ cb, const uint32_t *__restrict__ entry_off,
                                   uint64_t n_bytes, uint32_t last_byte, const uint8_t *__restrict__ d_last, uint8_t *d_file,
                                   const ShardPlan *__restrict__ plan)
{
    if (plan && plan->status) return;                   // the image does not fit (or the codebook is unusable)
    if (d_last) last_byte = *d_last;
    uint32_t *words = reinterpret_cast<uint32_t *>((uintptr_t)d_file & ~(uintptr_t)3);
    const unsigned long long file_bit0 = ((uintptr_t)d_file & 3) * 8;
    const uint32_t pre = 3 + (uint32_t)(n_bytes & 1);
    const unsigned long long stream0 = file_bit0 + pre * 8ull;
    const uint32_t U = cb->U;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < U) {
        uint32_t sym = cb->order[k];
        uint32_t len = cb->len[sym];
        unsigned long long pos = stream0 + entry_off[k];
        or_bits(words, pos, (sym << 8) | (len & 255u), 24);    // C:463-466: symbol high byte first, then len mod 256
        or_bits(words, pos + 24, cb->code[sym], len);          // C:470-481
    }
    if (k == 0) {
        or_bits(words, file_bit0, U & 0xFF, 8);                // C:434 (65536 -> 0x0000)
        or_bits(words, file_bit0 + 8, (U >> 8) & 0xFF, 8);
        or_bits(words, file_bit0 + 16, n_bytes & 1, 8);        // C:438
        if (n_bytes & 1) or_bits(words, file_bit0 + 24, last_byte & 0xFF, 8);   // C:439-443
        unsigned long long pos = stream0 + cb->table_bits;
        for (int i = 0; i < 8; i++) or_bits(words, pos + 8 * i, (n_bytes >> (8 * i)) & 0xFF, 8);   // C:661-669
    }
}

// dot(hist, len): a shard's payload bit count
__global__ void shard_bits_kernel(const unsigned long long *__restrict__ hist, const Codebook *__restrict__ cb,
                                  unsigned long long *__restrict__ out)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = s < NSYM ? hist[s] * cb->len[s] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

// ---------------------------------------------------------------------------------
// the entry offsets live behind the Codebook object in the same allocation
static __host__ __device__ inline uint32_t *entry_off_of(Codebook *cb)
{
    return reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(cb) + sizeof(Codebook));
}
size_t codebook_alloc_bytes() { return sizeof(Codebook) + NSYM * sizeof(uint32_t); }

int launch_codebook(Ctx *c, const unsigned long long *d_hist, Codebook *d_cb)
{
    int rc = ensure_ws(c, sizeof(CbWork));
    if (rc) return rc;
    CbWork *w = reinterpret_cast<CbWork *>(c->ws);
    HF_CUDA(c, cudaMemsetAsync(d_cb, 0, codebook_alloc_bytes(), c->stream));
    HF_PROF(c, "cb_sort_tree_kernel"); cb_sort_tree_kernel<<<CB_CLUSTER, CB_THREADS, 0, c->stream>>>(d_hist, w, d_cb);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "cb_codes_kernel"); cb_codes_kernel<<<NSYM / 256, 256, 0, c->stream>>>(d_hist, w, d_cb);
    HF_LAUNCH_CHECK(c);
    HF_PROF(c, "cb_entry_scan_kernel"); cb_entry_scan_kernel<<<1, CB_THREADS, 0, c->stream>>>(w, entry_off_of(d_cb));
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_shard_bits(Ctx *c, const unsigned long long *d_hist, const Codebook *d_cb,
                      unsigned long long *d_bits)
{
    HF_CUDA(c, cudaMemsetAsync(d_bits, 0, sizeof(unsigned long long), c->stream));
    HF_PROF(c, "shard_bits_kernel"); shard_bits_kernel<<<NSYM / 256, 256, 0, c->stream>>>(d_hist, d_cb, d_bits);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

int launch_header_pack(Ctx *c, const Codebook *d_cb, uint64_t n_bytes, uint32_t last_byte, const uint8_t *d_last,
                       uint8_t *d_file, uint64_t capacity, const ShardPlan *plan)
{
    // worst-case header: 4 + 65536 * (24 + 64) / 8 + 8 bytes; zero what the capacity allows
    uint64_t bound = 4 + (uint64_t)NSYM * 11 + 8 + 8;
    uint64_t z = capacity < bound ? capacity : bound;
    if (capacity < 12) return set_err(c, HF_ERR_CAPACITY, "hf_header_pack: capacity %llu too small",
                                      (unsigned long long)capacity);
    HF_CUDA(c, cudaMemsetAsync(d_file, 0, z, c->stream));
    HF_PROF(c, "header_pack_kernel"); header_pack_kernel<<<NSYM / 256, 256, 0, c->stream>>>(d_cb, entry_off_of(const_cast<Codebook *>(d_cb)),
                                                         n_bytes, last_byte, d_last, d_file, plan);
    HF_LAUNCH_CHECK(c);
    return HF_OK;
}

}  // namespace hf
