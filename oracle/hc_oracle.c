/*
 * hc_oracle.c -- CPU ORACLE (test infrastructure only; see hc_oracle.h).
 *
 * Plain-C restatement of the reference pipeline.  Every function cites the
 * reference file:line (relative to the reference repo root) it follows.  The
 * code is deliberately sequential and simple: it is the checker, never the
 * thing shipped or measured.
 */
#include "hc_oracle.h"

#include <stdlib.h>
#include <string.h>

void hco_free(void *p) { free(p); }

/* ------------------------------------------------------------------ */
/* growable byte vector                                                */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *d; size_t n, cap; } bvec;

static void bv_push(bvec *v, uint8_t b)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 256;
        v->d = (uint8_t *)realloc(v->d, v->cap);
    }
    v->d[v->n++] = b;
}

static void bv_append(bvec *v, const uint8_t *p, size_t n)
{
    for (size_t i = 0; i < n; i++) bv_push(v, p[i]);
}

/* ------------------------------------------------------------------ */
/* differential model -- src/transform.cpp:220-239                      */
/* ------------------------------------------------------------------ */
void hco_diff_apply(uint8_t *v, size_t n)
{
    uint8_t prev = 0;                       /* :222 */
    for (size_t i = 0; i < n; i++) {
        uint8_t cur = v[i];
        v[i] = (uint8_t)(cur - prev);       /* :226 truncated underflow */
        prev = cur;
    }
}

void hco_diff_revert(uint8_t *v, size_t n)
{
    uint8_t prev = 0;                       /* :233 */
    for (size_t i = 0; i < n; i++) {
        v[i] = (uint8_t)(v[i] + prev);      /* :236 truncated overflow */
        prev = v[i];
    }
}

/* ------------------------------------------------------------------ */
/* MNP-5 RLE -- src/transform.cpp:241-279 (encode), :137-159 (decode)   */
/* ------------------------------------------------------------------ */
size_t hco_rle_bound(size_t n) { return n + n / 3 + 4; }

size_t hco_rle_encode(const uint8_t *in, size_t n, uint8_t *out)
{
    size_t m = 0;
    uint8_t match_byte = 0;                 /* :245 */
    int match_count = 0;                    /* :246 */
    for (size_t i = 0; i < n; i++) {
        uint8_t cur = in[i];
        int is_last = (i + 1 == n);
        /* :252 first/reset and last iteration are excluded from matching */
        if (cur == match_byte && match_count != 0 && !is_last) {
            match_count++;
            if (match_count <= 3) {         /* :256 */
                out[m++] = cur;
            } else if (match_count == 258) {/* :259 255 + 3 */
                out[m++] = 255;
                match_count = 0;
            }
        } else {
            if (match_count >= 3)           /* :267 */
                out[m++] = (uint8_t)(match_count - 3);
            out[m++] = cur;
            match_byte = cur;
            match_count = 1;
        }
    }
    return m;
}

/* one decoder step, src/transform.cpp:137-159 */
static void rle_step(bvec *tar, uint8_t *match_byte, int *match_count, uint8_t cur)
{
    if (*match_count == 3) {
        for (int i = 0; i < cur; i++) bv_push(tar, *match_byte);
        *match_count = 0;
    } else {
        bv_push(tar, cur);
        if (*match_byte == cur) {
            (*match_count)++;
        } else {
            *match_byte = cur;
            *match_count = 1;
        }
    }
}

int hco_rle_decode(const uint8_t *in, size_t m, uint8_t **out, size_t *n)
{
    bvec v = {0, 0, 0};
    uint8_t mb = 0;
    int mc = 0;                             /* :285-286 */
    for (size_t i = 0; i < m; i++) rle_step(&v, &mb, &mc, in[i]);
    *out = v.d;
    *n = v.n;
    return 0;
}

/* ------------------------------------------------------------------ */
/* block geometry -- src/transform.cpp:25-62, :410-418                  */
/* ------------------------------------------------------------------ */
uint64_t hco_block_count(uint64_t w, uint64_t h, uint64_t b)
{
    uint64_t bw = w / b + (w % b != 0);
    uint64_t bh = h / b + (h % b != 0);
    return bw * bh;
}

static uint64_t block_base(uint64_t w, uint64_t b, uint64_t idx)
{
    uint64_t in_line = w / b + (w % b != 0);            /* :27 */
    return (idx / in_line) * w * b + (idx % in_line) * b;/* :28-31 */
}

static uint64_t block_size_x(uint64_t w, uint64_t base, uint64_t b)
{
    uint64_t bx = base % w;                             /* :37 */
    return (bx + b > w) ? b - (bx + b - w) : b;         /* :40-42 */
}

static uint64_t block_size_y(uint64_t w, uint64_t h, uint64_t base, uint64_t b)
{
    uint64_t by = base / w;                             /* :54 */
    return (by + b > h) ? b - (by + b - h) : b;         /* :57-59 */
}

/* src/transform.cpp:66-94; hor: row-major inside the block, !hor: column-major */
static void block_vector(const uint8_t *mat, uint64_t w, uint64_t h, uint64_t b,
                         uint64_t idx, int hor, uint8_t *dst, uint64_t *len)
{
    uint64_t base = block_base(w, b, idx);
    uint64_t sx = block_size_x(w, base, b);
    uint64_t sy = block_size_y(w, h, base, b);
    if (!hor) { uint64_t t = sx; sx = sy; sy = t; }     /* :79-81 */
    uint64_t k = 0;
    for (uint64_t y = 0; y < sy; y++)
        for (uint64_t x = 0; x < sx; x++) {
            uint64_t xi = hor ? x : y, yi = hor ? y : x;/* :86-87 */
            dst[k++] = mat[base + yi * w + xi];
        }
    *len = k;
}

/* src/headers.cpp:18-63 */
static void adapt_header(bvec *v, uint64_t w, uint64_t h, uint64_t b,
                         const uint8_t *dirs, uint64_t ndirs)
{
    for (int i = 7; i >= 0; i--) bv_push(v, (uint8_t)(w >> (8 * i)));   /* :27-29 big endian */
    for (int i = 7; i >= 0; i--) bv_push(v, (uint8_t)(h >> (8 * i)));
    for (int i = 7; i >= 0; i--) bv_push(v, (uint8_t)(b >> (8 * i)));
    uint8_t cur = 0;
    uint64_t bits = 0;
    for (uint64_t i = 0; i < ndirs; i++) {                              /* :43-51 */
        cur = (uint8_t)((cur << 1) | (dirs[i] & 1));
        bits++;
        if (bits % 8 == 0) bv_push(v, cur);
    }
    if (bits % 8 != 0) {                                                /* :53-60 */
        do { cur = (uint8_t)(cur << 1); bits++; } while (bits % 8 != 0);
        bv_push(v, cur);
    }
}

/* src/transform.cpp:97-134 */
int hco_adapt_encode_bs(const uint8_t *in, uint64_t w, uint64_t h, uint64_t b,
                        uint8_t **out, size_t *m)
{
    uint64_t nb = hco_block_count(w, h, b);
    uint8_t *dirs = (uint8_t *)malloc(nb ? nb : 1);
    uint64_t maxblk = b * b;
    uint8_t *blk = (uint8_t *)malloc(maxblk ? maxblk : 1);
    uint8_t *hv = (uint8_t *)malloc(hco_rle_bound(maxblk));
    uint8_t *vv = (uint8_t *)malloc(hco_rle_bound(maxblk));
    bvec data = {0, 0, 0};
    for (uint64_t i = 0; i < nb; i++) {
        uint64_t len;
        block_vector(in, w, h, b, i, 1, blk, &len);
        size_t hs = hco_rle_encode(blk, len, hv);                       /* :110 */
        block_vector(in, w, h, b, i, 0, blk, &len);
        size_t vs = hco_rle_encode(blk, len, vv);                       /* :111 */
        if (hs <= vs) { dirs[i] = 1; bv_append(&data, hv, hs); }        /* :114 tie -> horizontal */
        else          { dirs[i] = 0; bv_append(&data, vv, vs); }
    }
    bvec fin = {0, 0, 0};
    adapt_header(&fin, w, h, b, dirs, nb);                              /* :127 */
    bv_append(&fin, data.d, data.n);                                    /* :131 */
    free(dirs); free(blk); free(hv); free(vv); free(data.d);
    *out = fin.d;
    *m = fin.n;
    return 0;
}

/* src/transform.cpp:294-328 */
int hco_adapt_encode(const uint8_t *in, uint64_t w, uint64_t h,
                     uint8_t **out, size_t *m, uint64_t *chosen_b)
{
    uint64_t cur = 8;                                   /* INIT_RLE_BLOCK_SIZE, transform.hpp:17 */
    if (w < cur || h < cur) return 12;                  /* :300-304 */
    uint8_t *best; size_t best_n; uint64_t best_b = cur;
    hco_adapt_encode_bs(in, w, h, cur, &best, &best_n); /* :309 */
    cur *= 2;
    int steps = 1;
    while (steps <= 7 && cur <= w && cur <= h) {        /* :314-315, MAX_RLE_DOUBLING_STEPS 7 */
        uint8_t *c; size_t cn;
        hco_adapt_encode_bs(in, w, h, cur, &c, &cn);
        if (cn < best_n) {                              /* :319 strictly smaller */
            free(best); best = c; best_n = cn; best_b = cur;
        } else {
            free(c);
        }
        cur *= 2;
        steps++;
    }
    *out = best;
    *m = best_n;
    if (chosen_b) *chosen_b = best_b;
    return 0;
}

/* src/transform.cpp:330-361 with src/headers.cpp:65-105, src/transform.cpp:162-216 */
int hco_adapt_decode(const uint8_t *in, size_t m, uint8_t **out, size_t *n)
{
    size_t pos = 0;
    if (m < 24) return 10;                              /* headers.cpp:67-71 */
    uint64_t w = 0, h = 0, b = 0;
    for (int i = 0; i < 8; i++) w = (w << 8) | in[pos++];
    for (int i = 0; i < 8; i++) h = (h << 8) | in[pos++];
    for (int i = 0; i < 8; i++) b = (b << 8) | in[pos++];
    if (b == 0) return 10;                              /* reference divides by zero here (UB) */
    uint64_t nb = hco_block_count(w, h, b);             /* headers.cpp:85 */
    uint8_t *dirs = (uint8_t *)malloc(nb ? nb : 1);
    uint8_t curb = 0;
    for (uint64_t i = 0; i < nb; i++) {                 /* headers.cpp:90-102 */
        if (i % 8 == 0) {
            if (pos >= m) { free(dirs); return 11; }
            curb = in[pos++];
        }
        dirs[i] = (uint8_t)((curb >> (7 - (i % 8))) & 1);
    }
    uint8_t *mat = (uint8_t *)calloc((w * h) != 0 ? w * h : 1, 1);   /* transform.cpp:340 */
    bvec blk = {0, 0, 0};
    for (uint64_t i = 0; i < nb; i++) {
        uint64_t base = block_base(w, b, i);
        uint64_t sx = block_size_x(w, base, b);
        uint64_t sy = block_size_y(w, h, base, b);
        uint64_t req = sx * sy;
        /* revertRLEBlock, transform.cpp:162-187: fresh state per block */
        blk.n = 0;
        uint8_t mb = 0; int mc = 0;
        while (blk.n < req) {
            if (pos >= m) { free(dirs); free(mat); free(blk.d); return 14; }
            rle_step(&blk, &mb, &mc, in[pos++]);
        }
        if (blk.n != req) { free(dirs); free(mat); free(blk.d); return 13; }
        /* insertBlockVector, transform.cpp:191-216 */
        int hor = dirs[i];
        uint64_t ex = sx, ey = sy;
        if (!hor) { ex = sy; ey = sx; }
        uint64_t k = 0;
        for (uint64_t y = 0; y < ey; y++)
            for (uint64_t x = 0; x < ex; x++) {
                uint64_t xi = hor ? x : y, yi = hor ? y : x;
                mat[base + yi * w + xi] = blk.d[k++];
            }
    }
    free(blk.d);
    free(dirs);
    if (pos != m) { free(mat); return 15; }             /* transform.cpp:354-358 */
    *out = mat;
    *n = (size_t)(w * h);
    return 0;
}

/* ------------------------------------------------------------------ */
/* bit writer / reader (src/main.cpp:78-84, :107-112: MSB first)         */
/* ------------------------------------------------------------------ */
typedef struct { bvec v; uint8_t cur; int nb; uint64_t total; } bitw;

static void bw_put(bitw *w, int bit)
{
    w->cur = (uint8_t)((w->cur << 1) | (bit & 1));
    w->nb++;
    w->total++;
    if (w->nb == 8) { bv_push(&w->v, w->cur); w->cur = 0; w->nb = 0; }
}

static void bw_flush(bitw *w)               /* transform.cpp:379-381 zero pad */
{
    while (w->nb != 0) {
        w->cur = (uint8_t)(w->cur << 1);
        w->nb++;
        if (w->nb == 8) { bv_push(&w->v, w->cur); w->cur = 0; w->nb = 0; }
    }
}

typedef struct { const uint8_t *d; uint64_t nbits, pos; } bitr;

static int br_get(bitr *r)                  /* -1 when empty (huffman.cpp:65-67) */
{
    if (r->pos >= r->nbits) return -1;
    int b = (r->d[r->pos >> 3] >> (7 - (r->pos & 7))) & 1;
    r->pos++;
    return b;
}

/* ------------------------------------------------------------------ */
/* FGK tree, mode 0: faithful restatement of src/huffman.cpp            */
/* nodes live in an arena indexed by creation order; -1 == nullptr       */
/* ------------------------------------------------------------------ */
#define FGK_MAX_NODES 513

typedef struct {
    uint16_t num[FGK_MAX_NODES];
    uint64_t freq[FGK_MAX_NODES];
    uint8_t  sym[FGK_MAX_NODES];
    int parent[FGK_MAX_NODES], left[FGK_MAX_NODES], right[FGK_MAX_NODES];
    int n, root, nyt;
    int symnode[256];
    uint64_t levels, swaps;                 /* statistics only */
} ptree;

static int pt_new(ptree *t, uint16_t num, uint8_t sym, int parent)
{
    int i = t->n++;
    t->num[i] = num; t->freq[i] = 0; t->sym[i] = sym;
    t->parent[i] = parent; t->left[i] = -1; t->right[i] = -1;
    return i;
}

static void pt_init(ptree *t)               /* huffman.cpp:23-31 */
{
    t->n = 0;
    for (int i = 0; i < 256; i++) t->symnode[i] = -1;
    t->root = pt_new(t, 512, 0, -1);        /* 2 * MAX_SYMBOLS */
    t->nyt = t->root;
    t->levels = t->swaps = 0;
}

static int pt_is_leaf(const ptree *t, int n) { return t->left[n] < 0; }  /* :15-19 */

/* huffman.cpp:136-155: bits root->node; returns length, code[] MSB(first) at index 0 */
static int pt_code(const ptree *t, int node, uint8_t *code)
{
    int len = 0;
    uint8_t tmp[FGK_MAX_NODES];
    for (int c = node; c != t->root; c = t->parent[c])
        tmp[len++] = (t->left[t->parent[c]] == c) ? 0 : 1;
    for (int i = 0; i < len; i++) code[i] = tmp[len - 1 - i];          /* :153 reverse */
    return len;
}

/* huffman.cpp:157-184 */
static int pt_find_succ(const ptree *t, int node, uint64_t freq)
{
    int succ = -1;
    if (!pt_is_leaf(t, node) && t->freq[node] > freq) {
        int l = pt_find_succ(t, t->left[node], freq);
        int r = pt_find_succ(t, t->right[node], freq);
        if (l >= 0 && r >= 0) succ = (t->num[l] > t->num[r]) ? l : r;   /* :169-173 */
        else succ = (l >= 0) ? l : r;
    } else if (t->freq[node] == freq) {
        succ = node;
    }
    return succ;
}

/* huffman.cpp:186-217 */
static void pt_swap(ptree *t, int a, int b)
{
    uint16_t an = t->num[a]; t->num[a] = t->num[b]; t->num[b] = an;
    int a_left = (t->left[t->parent[a]] == a);
    int b_left = (t->left[t->parent[b]] == b);
    if (a_left) t->left[t->parent[a]] = b; else t->right[t->parent[a]] = b;
    if (b_left) t->left[t->parent[b]] = a; else t->right[t->parent[b]] = a;
    int ap = t->parent[a]; t->parent[a] = t->parent[b]; t->parent[b] = ap;
}

/* huffman.cpp:95-128 */
static void pt_update(ptree *t, uint8_t symbol)
{
    int node = t->symnode[symbol];
    if (node < 0) {                                                    /* :99-111 NYT split */
        int lc = pt_new(t, (uint16_t)(t->num[t->nyt] - 2), 0, t->nyt);
        node = pt_new(t, (uint16_t)(t->num[t->nyt] - 1), symbol, t->nyt);
        t->left[t->nyt] = lc;
        t->right[t->nyt] = node;
        t->nyt = lc;
        t->symnode[symbol] = node;
    }
    while (node != t->root) {                                          /* :113-125 */
        int succ = pt_find_succ(t, t->root, t->freq[node]);
        if (succ >= 0 && succ != t->parent[node] && succ != node) {
            pt_swap(t, node, succ);
            t->swaps++;
        }
        t->freq[node]++;
        node = t->parent[node];
        t->levels++;
    }
    t->freq[node]++;                                                   /* :127 */
}

/* huffman.cpp:37-58 */
static void pt_encode(const ptree *t, uint8_t symbol, bitw *w)
{
    uint8_t code[FGK_MAX_NODES];
    int sn = t->symnode[symbol];
    if (sn < 0) {
        int len = pt_code(t, t->nyt, code);
        for (int i = 0; i < len; i++) bw_put(w, code[i]);
        for (int i = 8; i > 0; i--) bw_put(w, (symbol >> (i - 1)) & 1); /* :46-50 */
    } else {
        int len = pt_code(t, sn, code);
        for (int i = 0; i < len; i++) bw_put(w, code[i]);
    }
}

/* huffman.cpp:60-93 */
static int pt_decode(const ptree *t, bitr *r)
{
    int cur = t->root;
    while (!pt_is_leaf(t, cur)) {
        int b = br_get(r);
        if (b < 0) return -1;
        cur = b ? t->right[cur] : t->left[cur];
    }
    if (cur == t->nyt) {
        int s = 0;
        for (int i = 0; i < 8; i++) {
            int b = br_get(r);
            if (b < 0) return -1;
            s = (s << 1) | b;
        }
        return s;
    }
    return t->sym[cur];
}

/* ------------------------------------------------------------------ */
/* FGK tree, mode 1: node-number indexed arrays.           */
/* slot = node number 0..512; siblings are adjacent (even = left).      */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t w[FGK_MAX_NODES];
    int parent[FGK_MAX_NODES];
    int kid[FGK_MAX_NODES];     /* >=0: slot of left child; -1: leaf */
    int sym[FGK_MAX_NODES];     /* leaf symbol, -1 for NYT/internal */
    int slot_of[256];
    int nyt;
} atree;

static void at_init(atree *t)
{
    for (int i = 0; i < FGK_MAX_NODES; i++) { t->w[i] = 0; t->parent[i] = -1; t->kid[i] = -1; t->sym[i] = -1; }
    for (int i = 0; i < 256; i++) t->slot_of[i] = -1;
    t->nyt = 512;
}

static void at_put_code(const atree *t, int s, bitw *w)
{
    uint8_t tmp[FGK_MAX_NODES]; int l = 0;
    for (; s != 512; s = t->parent[s]) tmp[l++] = (uint8_t)(s & 1);
    for (int i = l - 1; i >= 0; i--) bw_put(w, tmp[i]);
}

static void at_update(atree *t, int symbol)
{
    int s = t->slot_of[symbol];
    if (s < 0) {
        int n = t->nyt;
        t->kid[n] = n - 2; t->sym[n] = -1;
        t->parent[n - 2] = n; t->parent[n - 1] = n;
        t->kid[n - 2] = -1; t->kid[n - 1] = -1;
        t->sym[n - 1] = symbol; t->sym[n - 2] = -1;
        t->w[n - 1] = 0; t->w[n - 2] = 0;
        t->slot_of[symbol] = n - 1;
        t->nyt = n - 2;
        s = n - 1;
    }
    while (s != 512) {
        int l = s;
        while (l < 512 && t->w[l + 1] == t->w[s]) l++;
        if (l != s && l != t->parent[s]) {
            /* exchange the contents (subtrees) of slots s and l */
            int ks = t->kid[s], kl = t->kid[l], ys = t->sym[s], yl = t->sym[l];
            t->kid[s] = kl; t->sym[s] = yl;
            t->kid[l] = ks; t->sym[l] = ys;
            if (kl >= 0) { t->parent[kl] = s; t->parent[kl + 1] = s; }
            else if (yl >= 0) t->slot_of[yl] = s; else t->nyt = s;
            if (ks >= 0) { t->parent[ks] = l; t->parent[ks + 1] = l; }
            else if (ys >= 0) t->slot_of[ys] = l; else t->nyt = l;
            s = l;
        }
        t->w[s]++;
        s = t->parent[s];
    }
    t->w[512]++;
}

static int at_decode(const atree *t, bitr *r)
{
    int s = 512;
    while (t->kid[s] >= 0) {
        int b = br_get(r);
        if (b < 0) return -1;
        s = t->kid[s] + b;
    }
    if (s == t->nyt) {
        int v = 0;
        for (int i = 0; i < 8; i++) {
            int b = br_get(r);
            if (b < 0) return -1;
            v = (v << 1) | b;
        }
        return v;
    }
    return t->sym[s];
}

/* ------------------------------------------------------------------ */
/* Huffman stream drivers -- src/transform.cpp:363-406                   */
/* ------------------------------------------------------------------ */
int hco_fgk_encode(const uint8_t *sym, size_t m, int mode,
                   uint8_t **out, size_t *nbytes, uint64_t *nbits)
{
    bitw w; memset(&w, 0, sizeof w);
    if (mode == 0) {
        ptree *t = (ptree *)malloc(sizeof *t);
        pt_init(t);
        for (size_t i = 0; i < m; i++) {    /* :370-376 encode, append, update */
            pt_encode(t, sym[i], &w);
            pt_update(t, sym[i]);
        }
        free(t);
    } else {
        atree *t = (atree *)malloc(sizeof *t);
        at_init(t);
        for (size_t i = 0; i < m; i++) {
            int s = t->slot_of[sym[i]];
            if (s < 0) {
                at_put_code(t, t->nyt, &w);
                for (int k = 8; k > 0; k--) bw_put(&w, (sym[i] >> (k - 1)) & 1);
            } else {
                at_put_code(t, s, &w);
            }
            at_update(t, sym[i]);
        }
        free(t);
    }
    if (nbits) *nbits = w.total;
    bw_flush(&w);
    if (!w.v.d) w.v.d = (uint8_t *)malloc(1);
    *out = w.v.d;
    *nbytes = w.v.n;
    return 0;
}

int hco_fgk_decode(const uint8_t *bytes, size_t nbytes, uint64_t count, int mode,
                   uint8_t *sym_out)
{
    bitr r = { bytes, (uint64_t)nbytes * 8, 0 };
    if (mode == 0) {
        ptree *t = (ptree *)malloc(sizeof *t);
        pt_init(t);
        for (uint64_t i = 0; i < count; i++) {  /* :391-403 */
            int s = pt_decode(t, &r);
            if (s < 0) { free(t); return 9; }
            pt_update(t, (uint8_t)s);
            sym_out[i] = (uint8_t)s;
        }
        free(t);
    } else {
        atree *t = (atree *)malloc(sizeof *t);
        at_init(t);
        for (uint64_t i = 0; i < count; i++) {
            int s = at_decode(t, &r);
            if (s < 0) { free(t); return 9; }
            at_update(t, s);
            sym_out[i] = (uint8_t)s;
        }
        free(t);
    }
    return 0;
}

int hco_fgk_stats(const uint8_t *sym, size_t m, uint64_t *levels, uint64_t *swaps,
                  uint32_t *max_depth)
{
    ptree *t = (ptree *)malloc(sizeof *t);
    uint8_t code[FGK_MAX_NODES];
    uint32_t md = 0;
    pt_init(t);
    for (size_t i = 0; i < m; i++) {
        int sn = t->symnode[sym[i]];
        int len = pt_code(t, sn < 0 ? t->nyt : sn, code);
        if ((uint32_t)len > md) md = (uint32_t)len;
        pt_update(t, sym[i]);
    }
    *levels = t->levels; *swaps = t->swaps; *max_depth = md;
    free(t);
    return 0;
}

/* ------------------------------------------------------------------ */
/* whole-file pipeline -- src/main.cpp:39-87, :90-128                    */
/* ------------------------------------------------------------------ */
int hco_compress(const uint8_t *in, size_t n, int diff, int adapt, uint64_t width,
                 int mode, uint8_t **out, size_t *outlen)
{
    if (adapt && (n % width) != 0) return 6;            /* main.cpp:54-58 */
    uint64_t height = n / width;                        /* :59 */
    uint8_t *data = (uint8_t *)malloc(n ? n : 1);
    memcpy(data, in, n);
    if (diff) hco_diff_apply(data, n);                  /* :62-64 */
    uint8_t *sym; size_t m;
    if (adapt) {                                        /* :65-66 */
        int rc = hco_adapt_encode(data, width, height, &sym, &m, NULL);
        if (rc) { free(data); return rc; }
    } else {                                            /* :69 */
        sym = (uint8_t *)malloc(hco_rle_bound(n));
        m = hco_rle_encode(data, n, sym);
    }
    free(data);
    uint8_t *bits; size_t nbytes;
    hco_fgk_encode(sym, m, mode, &bits, &nbytes, NULL); /* :71 */
    free(sym);
    uint8_t *o = (uint8_t *)malloc(9 + nbytes);
    for (int i = 0; i < 8; i++) o[i] = (uint8_t)((uint64_t)m >> (8 * i));  /* headers.cpp:112-114 LE */
    o[8] = (uint8_t)(((diff ? 1 : 0) << 7) | ((adapt ? 1 : 0) << 6));      /* headers.cpp:117-122 */
    memcpy(o + 9, bits, nbytes);                        /* main.cpp:78-84 */
    free(bits);
    *out = o;
    *outlen = 9 + nbytes;
    return 0;
}

int hco_decompress(const uint8_t *in, size_t n, int mode, uint8_t **out, size_t *outlen)
{
    if (n < 9) return 8;                                /* main.cpp:93-104 */
    uint64_t count = 0;
    for (int i = 0; i < 8; i++) count |= (uint64_t)in[i] << (8 * i);
    int diff = (in[8] >> 7) & 1, adapt = (in[8] >> 6) & 1;
    /* guard: the reference would try to decode `count` symbols and exit(9) on underrun;
     * every symbol costs >= 1 bit except in a one-leaf tree, so count can exceed the bit
     * count only for degenerate streams; cap allocation by what the bits can express. */
    uint64_t cap = count;
    uint8_t *sym = (uint8_t *)malloc(cap ? cap : 1);
    if (!sym) return 9;
    int rc = hco_fgk_decode(in + 9, n - 9, count, mode, sym);          /* :116 */
    if (rc) { free(sym); return rc; }
    uint8_t *data; size_t dn;
    if (adapt) rc = hco_adapt_decode(sym, count, &data, &dn);          /* :118-119 */
    else       rc = hco_rle_decode(sym, count, &data, &dn);            /* :121 */
    free(sym);
    if (rc) return rc;
    if (diff) hco_diff_revert(data, dn);                               /* :123-125 */
    if (!data) data = (uint8_t *)malloc(1);
    *out = data;
    *outlen = dn;
    return 0;
}
