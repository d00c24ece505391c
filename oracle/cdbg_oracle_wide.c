/*
 * oracle/cdbg_oracle_wide.c -- TEST INFRASTRUCTURE ONLY.  The same restatement as cdbg_oracle.c, for any k <= 127, with
 * k-mers held in four 64-bit words (the reference's LargeInt<N>: G/tools/math/LargeInt.hpp:104-715, little-endian words,
 * comparison most-significant word first :502-509, revcomp :722-735; span dispatch G/tools/math/Integer.hpp:77-85).
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it (a) against cdbg_oracle.c on every k <= 31 golden case and
 * (b) against the real reference binary on the k > 31 golden cases (k = 32, 33, 47, 63, 64, 65, 95, 96, 127).
 *
 * Every stage follows the same reference lines as cdbg_oracle.c (see its header); only the integer type differs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define NW 4
typedef struct { uint64_t w[NW]; } km_t;   /* w[0] least significant; 2 bits per nt, first nt most significant */

typedef struct { uint32_t from, to; uint8_t from_sign, to_sign, pad[2]; } orcw_link;

typedef struct {
    int k;
    uint64_t n_bases, n_instances, n_distinct, n_solid;
    uint64_t *solid_kmers;   /* [n_solid * NW] ascending */
    uint32_t *solid_counts;
    uint64_t n_unitigs, total_nt;
    char *seq; uint64_t *seq_off; uint64_t *kc; uint32_t *abund;
    uint64_t n_links; orcw_link *links;
} orcw_result;

static inline int nt_valid(uint8_t c) {
    switch (c) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1; default: return 0; }
}
static const char NT_CHAR[4] = {'A', 'C', 'T', 'G'};

static inline km_t km_zero(void) { km_t z; memset(&z, 0, sizeof z); return z; }
static inline int km_cmp(const km_t *a, const km_t *b) {
    for (int i = NW - 1; i >= 0; i--) { if (a->w[i] < b->w[i]) return -1; if (a->w[i] > b->w[i]) return 1; }
    return 0;
}
static inline km_t km_mask(km_t x, int n) {   /* keep the low 2n bits */
    int bits = 2 * n;
    for (int i = 0; i < NW; i++) {
        int lo = 64 * i;
        if (bits <= lo) x.w[i] = 0;
        else if (bits < lo + 64) x.w[i] &= ((1ULL << (bits - lo)) - 1);
    }
    return x;
}
static inline km_t km_shl2(km_t x) { for (int i = NW - 1; i > 0; i--) x.w[i] = (x.w[i] << 2) | (x.w[i - 1] >> 62); x.w[0] <<= 2; return x; }
static inline km_t km_shr2(km_t x) { for (int i = 0; i < NW - 1; i++) x.w[i] = (x.w[i] >> 2) | (x.w[i + 1] << 62); x.w[NW - 1] >>= 2; return x; }
static inline void km_or_at(km_t *x, uint64_t code, int nt_pos) { /* nt_pos counted from the least significant nucleotide */
    int bit = 2 * nt_pos; x->w[bit >> 6] |= code << (bit & 63);
}
static inline uint64_t km_nt(const km_t *x, int nt_pos) { int bit = 2 * nt_pos; return (x->w[bit >> 6] >> (bit & 63)) & 3; }
static km_t km_revcomp(km_t x, int n) {
    km_t r = km_zero();
    for (int i = 0; i < n; i++) km_or_at(&r, km_nt(&x, i) ^ 2, n - 1 - i);
    return r;
}
static inline uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static inline uint64_t km_hash(const km_t *x) { uint64_t h = 0; for (int i = 0; i < NW; i++) h = mix64(h ^ x->w[i]); return h; }

/* ---- counting table */
typedef struct { km_t *keys; uint32_t *vals; uint8_t *used; uint64_t cap, n; } ktab;
static int ktab_init(ktab *t, uint64_t cap) {
    t->cap = cap; t->n = 0;
    t->keys = (km_t *)malloc(cap * sizeof(km_t)); t->vals = (uint32_t *)calloc(cap, 4); t->used = (uint8_t *)calloc(cap, 1);
    return (t->keys && t->vals && t->used) ? 0 : -1;
}
static void ktab_free(ktab *t) { free(t->keys); free(t->vals); free(t->used); }
static uint64_t ktab_slot(const ktab *t, const km_t *key) {
    uint64_t i = km_hash(key) & (t->cap - 1);
    while (t->used[i] && km_cmp(&t->keys[i], key) != 0) i = (i + 1) & (t->cap - 1);
    return i;
}
static int ktab_add(ktab *t, const km_t *key) {
    uint64_t s = ktab_slot(t, key);
    if (!t->used[s]) {
        if ((t->n + 1) * 10 > t->cap * 6) {
            ktab n;
            if (ktab_init(&n, t->cap * 2)) return -1;
            for (uint64_t i = 0; i < t->cap; i++) if (t->used[i]) { uint64_t q = ktab_slot(&n, &t->keys[i]); n.keys[q] = t->keys[i]; n.vals[q] = t->vals[i]; n.used[q] = 1; }
            n.n = t->n; ktab_free(t); *t = n;
            s = ktab_slot(t, key);
        }
        t->keys[s] = *key; t->used[s] = 1; t->n++;
    }
    t->vals[s]++;
    return 0;
}

typedef struct { km_t key; uint64_t val; } kvw;
static int kvw_cmp(const void *a, const void *b) {
    const kvw *x = (const kvw *)a, *y = (const kvw *)b;
    int c = km_cmp(&x->key, &y->key);
    if (c) return c;
    return x->val < y->val ? -1 : x->val > y->val;
}

static int count_stage(const uint8_t *bytes, uint64_t n, int k, uint32_t amin, uint32_t amax, orcw_result *res) {
    ktab t;
    if (ktab_init(&t, 1ULL << 12)) return -1;
    km_t fwd = km_zero(), rc = km_zero(); int run = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t c = bytes[i];
        if (!nt_valid(c)) { run = 0; continue; }
        res->n_bases++;
        uint64_t code = (c >> 1) & 3;
        fwd = km_shl2(fwd); fwd.w[0] |= code; fwd = km_mask(fwd, k);
        rc = km_shr2(rc); km_or_at(&rc, code ^ 2, k - 1);
        if (++run >= k) {
            res->n_instances++;
            const km_t *cn = km_cmp(&fwd, &rc) < 0 ? &fwd : &rc;
            if (ktab_add(&t, cn)) return -1;
        }
    }
    res->n_distinct = t.n;
    uint64_t s = 0;
    for (uint64_t i = 0; i < t.cap; i++) if (t.used[i] && t.vals[i] >= amin && t.vals[i] <= amax) s++;
    kvw *a = (kvw *)malloc((s ? s : 1) * sizeof(kvw));
    uint64_t j = 0;
    for (uint64_t i = 0; i < t.cap; i++) if (t.used[i] && t.vals[i] >= amin && t.vals[i] <= amax) { a[j].key = t.keys[i]; a[j].val = t.vals[i]; j++; }
    ktab_free(&t);
    qsort(a, s, sizeof(kvw), kvw_cmp);
    res->n_solid = s;
    res->solid_kmers = (uint64_t *)malloc((s ? s : 1) * NW * 8);
    res->solid_counts = (uint32_t *)malloc((s ? s : 1) * 4);
    for (uint64_t i = 0; i < s; i++) { memcpy(res->solid_kmers + i * NW, a[i].key.w, NW * 8); res->solid_counts[i] = (uint32_t)a[i].val; }
    free(a);
    return 0;
}

#define NOLINK 0xFFFFFFFFu
static uint32_t uf_find(uint32_t *p, uint32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }

static int compact_stage(orcw_result *res) {
    const int k = res->k, k1 = k - 1;
    const uint64_t S = res->n_solid;
    kvw *left = (kvw *)malloc((2 * S + 1) * sizeof(kvw)), *right = (kvw *)malloc((2 * S + 1) * sizeof(kvw));
    uint64_t nl = 0, nr = 0;
    for (uint64_t i = 0; i < S; i++) {
        km_t x; memcpy(x.w, res->solid_kmers + i * NW, NW * 8);
        km_t a = km_shr2(x), ar = km_revcomp(a, k1);
        int c = km_cmp(&a, &ar);
        if (c <= 0) { left[nl].key = a; left[nl].val = 2 * i; nl++; }
        if (c >= 0) { right[nr].key = ar; right[nr].val = 2 * i; nr++; }
        km_t b = km_mask(x, k1), br = km_revcomp(b, k1);
        c = km_cmp(&b, &br);
        if (c <= 0) { right[nr].key = b; right[nr].val = 2 * i + 1; nr++; }
        if (c >= 0) { left[nl].key = br; left[nl].val = 2 * i + 1; nl++; }
    }
    qsort(left, nl, sizeof(kvw), kvw_cmp); qsort(right, nr, sizeof(kvw), kvw_cmp);
    uint32_t *link = (uint32_t *)malloc((2 * S + 1) * 4), *uf = (uint32_t *)malloc((S + 1) * 4);
    for (uint64_t i = 0; i < 2 * S; i++) link[i] = NOLINK;
    for (uint64_t i = 0; i < S; i++) uf[i] = (uint32_t)i;
    uint64_t iL = 0, iR = 0;
    while (iL < nl && iR < nr) {
        int c = km_cmp(&left[iL].key, &right[iR].key);
        if (c == 0) {
            uint64_t cl = 0, cr = 0;
            while (iL + cl < nl && km_cmp(&left[iL + cl].key, &left[iL].key) == 0) cl++;
            while (iR + cr < nr && km_cmp(&right[iR + cr].key, &left[iL].key) == 0) cr++;
            if (cl == 1 && cr == 1) {
                uint32_t eL = (uint32_t)left[iL].val, eR = (uint32_t)right[iR].val;
                uint32_t a = uf_find(uf, eL >> 1), b = uf_find(uf, eR >> 1);
                if (a != b) { link[eL] = eR; link[eR] = eL; uf[a] = b; }
            }
            iL += cl; iR += cr;
        } else if (c < 0) iL++; else iR++;
    }
    free(left); free(right); free(uf);
    uint8_t *seen = (uint8_t *)calloc(S + 1, 1);
    res->seq = (char *)malloc(S * (uint64_t)k + 16);
    res->seq_off = (uint64_t *)malloc((S + 2) * 8);
    res->kc = (uint64_t *)malloc((S + 1) * 8);
    res->abund = (uint32_t *)malloc((S + 1) * 4);
    uint64_t U = 0, nt = 0, na = 0;
    for (uint64_t i = 0; i < S; i++) {
        if (seen[i]) continue;
        uint32_t cur = (uint32_t)i, e = 0;
        for (;;) { uint32_t p = link[2 * cur + e]; if (p == NOLINK) break; cur = p >> 1; e = (p & 1) ^ 1; }
        uint32_t t1 = 2 * cur + e, c2 = cur, e2 = e ^ 1;
        for (;;) { uint32_t p = link[2 * c2 + e2]; if (p == NOLINK) break; c2 = p >> 1; e2 = (p & 1) ^ 1; }
        uint32_t t2 = 2 * c2 + e2, start = t1 < t2 ? t1 : t2;
        uint32_t v = start >> 1, in = start & 1;
        res->seq_off[U] = nt;
        uint64_t kc = 0; int first = 1;
        for (;;) {
            km_t x; memcpy(x.w, res->solid_kmers + (uint64_t)v * NW, NW * 8);
            if (in == 1) x = km_revcomp(x, k);
            seen[v] = 1;
            if (first) { for (int j = 0; j < k; j++) res->seq[nt++] = NT_CHAR[km_nt(&x, k - 1 - j)]; first = 0; }
            else res->seq[nt++] = NT_CHAR[km_nt(&x, 0)];
            res->abund[na++] = res->solid_counts[v];
            kc += res->solid_counts[v];
            uint32_t p = link[2 * v + (in ^ 1)];
            if (p == NOLINK) break;
            v = p >> 1; in = p & 1;
        }
        res->kc[U] = kc;
        U++;
    }
    res->seq_off[U] = nt; res->n_unitigs = U; res->total_nt = nt;
    free(seen); free(link);
    return 0;
}

static km_t encode_str(const char *s, int n) {
    km_t x = km_zero();
    for (int i = 0; i < n; i++) { x = km_shl2(x); x.w[0] |= ((uint8_t)s[i] >> 1) & 3; }
    return x;
}

static int link_stage(orcw_result *res) {
    const int k = res->k, k1 = k - 1;
    const uint64_t U = res->n_unitigs;
    kvw *ext = (kvw *)malloc((2 * U + 1) * sizeof(kvw));
    for (uint64_t u = 0; u < U; u++) {
        const char *s = res->seq + res->seq_off[u]; uint64_t len = res->seq_off[u + 1] - res->seq_off[u];
        km_t b = encode_str(s, k1), br = km_revcomp(b, k1), e = encode_str(s + len - k1, k1), er = km_revcomp(e, k1);
        int cb = km_cmp(&b, &br), ce = km_cmp(&e, &er);
        ext[2 * u].key = cb <= 0 ? b : br;     ext[2 * u].val = (u << 2) | ((uint64_t)(cb > 0) << 1) | 0;
        ext[2 * u + 1].key = ce <= 0 ? e : er; ext[2 * u + 1].val = (u << 2) | ((uint64_t)(ce > 0) << 1) | 1;
    }
    kvw *sorted = (kvw *)malloc((2 * U + 1) * sizeof(kvw));
    memcpy(sorted, ext, 2 * U * sizeof(kvw));
    qsort(sorted, 2 * U, sizeof(kvw), kvw_cmp);
    uint64_t cap = 16, n = 0; orcw_link *L = (orcw_link *)malloc(cap * sizeof(orcw_link));
    for (uint64_t u = 0; u < U; u++) for (int pos = 0; pos < 2; pos++) {
        km_t key = ext[2 * u + pos].key;
        int same = !((ext[2 * u + pos].val >> 1) & 1);
        km_t kr = km_revcomp(key, k1);
        int nevermind = (k1 % 2 == 0) && km_cmp(&kr, &key) == 0;
        uint64_t lo = 0, hi = 2 * U;
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (km_cmp(&sorted[mid].key, &key) < 0) lo = mid + 1; else hi = mid; }
        for (uint64_t j = lo; j < 2 * U && km_cmp(&sorted[j].key, &key) == 0; j++) {
            uint64_t v = sorted[j].val; int erc = (v >> 1) & 1, epos = v & 1; uint32_t eu = (uint32_t)(v >> 2);
            int ok;
            if (pos == 0) ok = (same && epos == 1 && !erc) || (same && epos == 0 && erc) || (!same && epos == 1 && erc) || (!same && epos == 0 && !erc);
            else ok = (same && epos == 0 && !erc) || (same && epos == 1 && erc) || (!same && epos == 0 && erc) || (!same && epos == 1 && !erc);
            if (!(ok || nevermind)) continue;
            if (n == cap) { cap *= 2; L = (orcw_link *)realloc(L, cap * sizeof(orcw_link)); }
            L[n].from = (uint32_t)u; L[n].to = eu; L[n].from_sign = pos == 0 ? '-' : '+'; L[n].to_sign = epos == 1 ? '-' : '+';
            L[n].pad[0] = L[n].pad[1] = 0; n++;
        }
    }
    res->links = L; res->n_links = n;
    free(ext); free(sorted);
    return 0;
}

void orcw_free(orcw_result *r) {
    free(r->solid_kmers); free(r->solid_counts); free(r->seq); free(r->seq_off); free(r->kc); free(r->abund); free(r->links);
    memset(r, 0, sizeof *r);
}

int orcw_build(const uint8_t *bytes, uint64_t n, int k, uint32_t amin, uint32_t amax, int stages, orcw_result *res) {
    memset(res, 0, sizeof *res);
    if (k < 4 || k > 127) return -2;
    res->k = k;
    if (count_stage(bytes, n, k, amin, amax, res)) return -1;
    if (!(stages & 2)) return 0;
    if (compact_stage(res)) return -1;
    if (!(stages & 4)) return 0;
    return link_stage(res);
}
