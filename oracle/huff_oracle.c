/*
 * huff_oracle.c — CPU restatement of the reference's results (see huff_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Plain C, 64-bit clean (the reference is only
 * defined for N < 2 GiB; this file is the truth beyond that).
 *
 * Parity pin: see header.  Citations are into /root/reference
 * (C: Compressor.cu, D: Decompressor.cu, h: gpuHuffmanConstruction.h).
 */
#include "huff_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* histogram: C:38-48 (symbol = data[2i] | data[2i+1] << 8)            */
/* ------------------------------------------------------------------ */
void ho_histogram(const uint8_t *in, uint64_t n, uint64_t *hist)
{
    memset(hist, 0, sizeof(uint64_t) * HO_NSYM);
    for (uint64_t i = 0; i + 1 < n; i += 2)
        hist[(uint32_t)in[i] | ((uint32_t)in[i + 1] << 8)]++;
}

/* ------------------------------------------------------------------ */
/* leaf order: C:387-393 (stable radix sort of counts carrying symbol   */
/* ids => ties keep ascending symbol), C:413-425 (non-zero suffix).     */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t f; uint32_t s; } leaf_t;

static int leaf_cmp(const void *a, const void *b)
{
    const leaf_t *x = (const leaf_t *)a, *y = (const leaf_t *)b;
    if (x->f != y->f) return x->f < y->f ? -1 : 1;
    return x->s < y->s ? -1 : (x->s > y->s);
}

static uint32_t sorted_leaves(const uint64_t *hist, leaf_t *lv)
{
    uint32_t U = 0;
    for (uint32_t s = 0; s < HO_NSYM; s++)
        if (hist[s]) { lv[U].f = hist[s]; lv[U].s = s; U++; }
    qsort(lv, U, sizeof(leaf_t), leaf_cmp);
    return U;
}

/* codes from a parent array.  h:468-494 records leaf->root one flag per level
 * (0 = child is the parent's left); h:562-573 maps flag 0 -> '1', 1 -> '0' and
 * reverses, so the path is root->leaf with 1 = left = first of the pair. */
static int finish_codebook(const uint64_t *hist, const leaf_t *lv, uint32_t U,
                           const int32_t *parent, const uint8_t *is_left,
                           ho_codebook_t *cb)
{
    memset(cb, 0, sizeof(*cb));
    cb->U = U;
    for (uint32_t k = 0; k < U; k++) {
        uint32_t s = lv[k].s;
        uint64_t code = 0;
        uint32_t len = 0;
        int32_t node = (int32_t)k;
        while (parent[node] >= 0) {
            if (len >= 64) return HO_ERR_CODE_TOO_LONG;
            code |= (uint64_t)is_left[node] << len;   /* first level seen = last path bit */
            len++;
            node = parent[node];
        }
        cb->order[k] = (uint16_t)s;
        cb->len[s] = (uint8_t)len;
        cb->code[s] = code;
        if (len > cb->maxlen) cb->maxlen = len;
        cb->table_bits += 16 + 8 + len;               /* C:463-481 */
        cb->payload_bits += hist[s] * len;
    }
    return HO_OK;
}

/* ------------------------------------------------------------------ */
/* tree, two-queue form (SURVEY 8.1 result spec).                      */
/* ------------------------------------------------------------------ */
int ho_codebook(const uint64_t *hist, ho_codebook_t *cb)
{
    leaf_t *lv = (leaf_t *)malloc(sizeof(leaf_t) * HO_NSYM);
    uint32_t U = sorted_leaves(hist, lv);
    uint32_t nn = U ? 2 * U - 1 : 0;
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * (nn + 1));
    uint8_t *is_left = (uint8_t *)calloc(nn + 1, 1);
    uint64_t *inf = (uint64_t *)malloc(sizeof(uint64_t) * (U + 1));
    for (uint32_t i = 0; i < nn; i++) parent[i] = -1;

    uint32_t l = 0, h = 0, t = 0;           /* leaf head, internal head, internal tail */
    for (uint32_t m = 0; m + 1 < U; m++) {
        int32_t pick[2];
        uint64_t sum = 0;
        for (int j = 0; j < 2; j++) {
            /* leaf wins ties: carried-over nodes precede this round's new nodes
             * (h:436-439 with KthElement's left-list-first rule, h:163-209) */
            if (h == t || (l < U && lv[l].f <= inf[h])) { pick[j] = (int32_t)l; sum += lv[l].f; l++; }
            else { pick[j] = (int32_t)(U + h); sum += inf[h]; h++; }
        }
        parent[pick[0]] = parent[pick[1]] = (int32_t)(U + t);
        is_left[pick[0]] = 1;                /* h:396-397: left = first of the pair */
        inf[t++] = sum;
    }
    int rc = finish_codebook(hist, lv, U, parent, is_left, cb);
    free(lv); free(parent); free(is_left); free(inf);
    return rc;
}

/* ------------------------------------------------------------------ */
/* tree, round form: a restatement of the mechanism of h:353-466.      */
/* ------------------------------------------------------------------ */
static uint32_t upper_clamped(const uint64_t *f, int32_t size, uint64_t val)
{   /* h:137-151: first index with f > val, but never past size-1; 0 when size <= 0 */
    int32_t lo = 0, hi = size - 1;
    while (lo < hi) {
        int32_t mid = lo + (hi - lo) / 2;
        if (f[mid] <= val) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
}

int ho_codebook_rounds(const uint64_t *hist, ho_codebook_t *cb)
{
    leaf_t *lv = (leaf_t *)malloc(sizeof(leaf_t) * HO_NSYM);
    uint32_t U = sorted_leaves(hist, lv);
    uint32_t nn = U ? 2 * U - 1 : 0;
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * (nn + 1));
    uint8_t *is_left = (uint8_t *)calloc(nn + 1, 1);
    uint64_t *qf = (uint64_t *)malloc(sizeof(uint64_t) * (U + 1));   /* nodeFreq */
    int32_t *qi = (int32_t *)malloc(sizeof(int32_t) * (U + 1));      /* nodeIndex */
    uint64_t *tf = (uint64_t *)malloc(sizeof(uint64_t) * (U + 1));   /* tempFreq */
    int32_t *ti = (int32_t *)malloc(sizeof(int32_t) * (U + 1));      /* tempIndex */
    for (uint32_t i = 0; i < nn; i++) parent[i] = -1;
    for (uint32_t i = 0; i < U; i++) { qf[i] = lv[i].f; qi[i] = (int32_t)i; }

    uint32_t size = U, next = U;
    while (size > 1) {
        uint64_t spec = qf[0] + qf[1];                                   /* h:381 */
        uint32_t pivot = upper_clamped(qf + 2, (int32_t)size - 2, spec) + 2;   /* h:385-386 */
        pivot -= pivot & 1;                                              /* h:387 */
        uint32_t keep = size - pivot, made = pivot >> 1;
        for (uint32_t i = 0; i < keep; i++) { tf[i] = qf[i + pivot]; ti[i] = qi[i + pivot]; }
        for (uint32_t i = 0; i < made; i++) {                            /* h:395-426 */
            int32_t a = qi[2 * i], b = qi[2 * i + 1];
            parent[a] = parent[b] = (int32_t)(next + i);
            is_left[a] = 1;
            tf[keep + i] = qf[2 * i] + qf[2 * i + 1];
            ti[keep + i] = (int32_t)(next + i);
        }
        next += made;
        /* h:436-439 + h:163-209: stable merge, carried-over list first on ties */
        uint32_t x = 0, y = 0, k = 0;
        while (x < keep || y < made) {
            if (y >= made || (x < keep && tf[x] <= tf[keep + y])) { qf[k] = tf[x]; qi[k] = ti[x]; x++; }
            else { qf[k] = tf[keep + y]; qi[k] = ti[keep + y]; y++; }
            k++;
        }
        size = keep + made;
    }
    int rc = finish_codebook(hist, lv, U, parent, is_left, cb);
    free(lv); free(parent); free(is_left); free(qf); free(qi); free(tf); free(ti);
    return rc;
}

/* ------------------------------------------------------------------ */
/* writer: C:637-669 (byte fields merged at the pending bit phase),    */
/* C:470-481 (code bits MSB first), C:597-601 (zero pad).              */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *p; uint64_t cap; uint64_t bitpos; int overflow; } bitw_t;

static void bw_put(bitw_t *w, uint64_t v, uint32_t nbits)
{   /* append the low nbits of v, MSB first */
    while (nbits) {
        uint64_t byte = w->bitpos >> 3;
        uint32_t room = 8 - (uint32_t)(w->bitpos & 7);
        uint32_t take = nbits < room ? nbits : room;
        if (byte >= w->cap) { w->overflow = 1; return; }
        uint32_t chunk = (uint32_t)((v >> (nbits - take)) & ((1u << take) - 1));
        w->p[byte] |= (uint8_t)(chunk << (room - take));
        w->bitpos += take;
        nbits -= take;
    }
}

static uint32_t preamble_bytes(uint64_t n) { return 3 + (uint32_t)(n & 1); }

uint64_t ho_compressed_size(const ho_codebook_t *cb, uint64_t n)
{
    uint64_t bits = cb->table_bits + 64 + cb->payload_bits;
    return preamble_bytes(n) + (bits + 7) / 8;
}

int ho_compress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n)
{
    uint64_t *hist = (uint64_t *)malloc(sizeof(uint64_t) * HO_NSYM);
    ho_codebook_t *cb = (ho_codebook_t *)malloc(sizeof(ho_codebook_t));
    ho_histogram(in, n, hist);
    int rc = ho_codebook(hist, cb);
    free(hist);
    if (rc) { free(cb); return rc; }
    uint64_t total = ho_compressed_size(cb, n);
    if (out_n) *out_n = total;
    if (total > cap) { free(cb); return HO_ERR_CAPACITY; }
    memset(out, 0, total);
    out[0] = (uint8_t)(cb->U & 0xFF);                 /* C:434 (65536 -> 0x0000) */
    out[1] = (uint8_t)((cb->U >> 8) & 0xFF);
    out[2] = (uint8_t)(n & 1);                        /* C:438 */
    if (n & 1) out[3] = in[n - 1];                    /* C:439-443 */
    bitw_t w = { out + preamble_bytes(n), total - preamble_bytes(n), 0, 0 };
    for (uint32_t k = 0; k < cb->U; k++) {            /* C:454-483 */
        uint32_t s = cb->order[k];
        bw_put(&w, s, 16);                            /* C:648-656: high byte first */
        bw_put(&w, cb->len[s], 8);                    /* C:459: length mod 256 */
        bw_put(&w, cb->code[s], cb->len[s]);
    }
    for (int i = 0; i < 8; i++) bw_put(&w, (n >> (8 * i)) & 0xFF, 8);   /* C:661-669 */
    for (uint64_t i = 0; i + 1 < n; i += 2) {         /* C:541-588 */
        uint32_t s = (uint32_t)in[i] | ((uint32_t)in[i + 1] << 8);
        bw_put(&w, cb->code[s], cb->len[s]);
    }
    rc = w.overflow ? HO_ERR_CAPACITY : HO_OK;
    free(cb);
    return rc;
}

/* ------------------------------------------------------------------ */
/* reader: D:68-108 header, D:129-163 tree insert, D:259-291 walk.     */
/* ------------------------------------------------------------------ */
typedef struct { const uint8_t *p; uint64_t nbytes; uint64_t bitpos; int overrun; } bitr_t;

static uint32_t br_bit(bitr_t *r)
{
    uint64_t byte = r->bitpos >> 3;
    if (byte >= r->nbytes) { r->overrun = 1; return 0; }
    uint32_t b = (r->p[byte] >> (7 - (r->bitpos & 7))) & 1;
    r->bitpos++;
    return b;
}

static uint64_t br_get(bitr_t *r, uint32_t nbits)
{
    uint64_t v = 0;
    while (nbits--) v = (v << 1) | br_bit(r);
    return v;
}

typedef struct { int32_t child[2]; int32_t sym; } tnode_t;

typedef struct {
    uint32_t U; int is_odd; uint8_t last; uint64_t n;
    tnode_t *nodes; uint32_t nnodes;
    bitr_t r;                                 /* positioned at the first payload bit */
    int single_zero_len;                      /* U == 1 with a zero-length code (R4) */
} parsed_t;

static int parse_header(const uint8_t *in, uint64_t nbytes, parsed_t *ph)
{
    memset(ph, 0, sizeof(*ph));
    if (nbytes < 3) return HO_ERR_FORMAT;
    uint32_t U = (uint32_t)in[0] | ((uint32_t)in[1] << 8);
    ph->is_odd = in[2] != 0;                                   /* D:76 */
    uint32_t pre = 3 + (ph->is_odd ? 1 : 0);
    if (nbytes < pre) return HO_ERR_FORMAT;
    if (ph->is_odd) ph->last = in[3];                          /* D:77-80 */
    /* D:70-71 maps 0 -> 65536.  Our defined behaviour for N < 2:
     * a file holding nothing but the 64-bit size after the preamble has U = 0. */
    if (U == 0) U = (nbytes - pre == 8) ? 0 : 65536;
    ph->U = U;
    ph->r.p = in + pre; ph->r.nbytes = nbytes - pre; ph->r.bitpos = 0; ph->r.overrun = 0;
    uint64_t cap_nodes = 2 * (uint64_t)(U ? U : 1) + 2;      /* a valid table needs 2U-1 */
    ph->nodes = (tnode_t *)malloc(sizeof(tnode_t) * cap_nodes);
    ph->nnodes = 1;
    ph->nodes[0].child[0] = ph->nodes[0].child[1] = -1; ph->nodes[0].sym = -1;
    for (uint32_t k = 0; k < U; k++) {
        uint32_t sym = (uint32_t)br_get(&ph->r, 16);           /* D:178-182 */
        uint32_t len = (uint32_t)br_get(&ph->r, 8);            /* D:93 */
        if (len == 0) {
            if (U == 1) { ph->single_zero_len = 1; ph->nodes[0].sym = (int32_t)sym; continue; }
            len = 65536;                                       /* D:94-95 */
        }
        int32_t node = 0;
        for (uint32_t i = 0; i < len; i++) {                   /* D:129-163 */
            uint32_t b = br_bit(&ph->r);
            if (ph->r.overrun) return HO_ERR_FORMAT;
            if (ph->nodes[node].child[b] < 0) {
                if (ph->nnodes >= cap_nodes) {                 /* malformed table: grow, bounded */
                    if (cap_nodes > (1u << 24)) return HO_ERR_FORMAT;
                    cap_nodes *= 2;
                    ph->nodes = (tnode_t *)realloc(ph->nodes, sizeof(tnode_t) * cap_nodes);
                }
                ph->nodes[ph->nnodes].child[0] = ph->nodes[ph->nnodes].child[1] = -1;
                ph->nodes[ph->nnodes].sym = -1;
                ph->nodes[node].child[b] = (int32_t)ph->nnodes++;
            }
            node = ph->nodes[node].child[b];
        }
        ph->nodes[node].sym = (int32_t)sym;
    }
    uint64_t n = 0;
    for (int i = 0; i < 8; i++) n |= br_get(&ph->r, 8) << (8 * i);   /* D:243-255 */
    if (ph->r.overrun) return HO_ERR_FORMAT;
    ph->n = n;
    return HO_OK;
}

int ho_decompressed_size(const uint8_t *in, uint64_t nbytes, uint64_t *out_n)
{
    parsed_t ph;
    int rc = parse_header(in, nbytes, &ph);
    if (!rc) *out_n = ph.n;
    free(ph.nodes);
    return rc;
}

int ho_decompress(const uint8_t *in, uint64_t nbytes, uint8_t *out, uint64_t cap, uint64_t *out_n)
{
    parsed_t ph;
    int rc = parse_header(in, nbytes, &ph);
    if (rc) { free(ph.nodes); return rc; }
    if (out_n) *out_n = ph.n;
    if (ph.n > cap) { free(ph.nodes); return HO_ERR_CAPACITY; }
    uint64_t nsym = ph.n / 2;                                   /* D:262 */
    if ((ph.n & 1) != (uint64_t)ph.is_odd) { free(ph.nodes); return HO_ERR_FORMAT; }
    for (uint64_t i = 0; i < nsym; i++) {
        int32_t node = 0;
        if (!ph.single_zero_len) {
            while (ph.nodes[node].child[0] >= 0 || ph.nodes[node].child[1] >= 0) {   /* D:265-282 */
                uint32_t b = br_bit(&ph.r);
                node = ph.nodes[node].child[b];
                if (node < 0 || ph.r.overrun) { free(ph.nodes); return HO_ERR_FORMAT; }
            }
        }
        int32_t sym = ph.nodes[node].sym;
        if (sym < 0) { free(ph.nodes); return HO_ERR_FORMAT; }
        out[2 * i] = (uint8_t)(sym & 0xFF);                     /* D:283: LE pair */
        out[2 * i + 1] = (uint8_t)(sym >> 8);
    }
    if (ph.is_odd) out[ph.n - 1] = ph.last;                     /* D:286-289 */
    free(ph.nodes);
    return HO_OK;
}

/* ------------------------------------------------------------------ */
/* SURVEY 2.3 clean-domain predicate for the reference GPU binary.     */
/* ------------------------------------------------------------------ */
int ho_reference_clean(const uint8_t *in, uint64_t n)
{
    if (n < 4 || n >= (1ull << 31)) return 0;
    uint64_t *hist = (uint64_t *)malloc(sizeof(uint64_t) * HO_NSYM);
    ho_codebook_t *cb = (ho_codebook_t *)malloc(sizeof(ho_codebook_t));
    ho_histogram(in, n, hist);
    int ok = ho_codebook(hist, cb) == HO_OK && cb->U >= 2;
    if (ok) {
        uint64_t hbits = cb->table_bits + 64;
        uint32_t phase = (uint32_t)(hbits & 7);                 /* bitCounter at C:541 */
        uint32_t r = (uint32_t)((hbits + cb->payload_bits) & 7);
        uint64_t ns = n / 2;
        uint32_t s0 = (uint32_t)in[0] | ((uint32_t)in[1] << 8);
        uint32_t sl = (uint32_t)in[2 * (ns - 1)] | ((uint32_t)in[2 * (ns - 1) + 1] << 8);
        if (phase && cb->len[s0] < 8 - phase) ok = 0;           /* R1, C:294-310 */
        if (r && cb->len[sl] < r) ok = 0;                       /* R2, C:227-246 */
    }
    free(hist); free(cb);
    return ok;
}
