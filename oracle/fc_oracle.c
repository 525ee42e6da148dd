/*
 * fc_oracle.c -- CPU restatement (plain C11) of FALCON's fc_consensus arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY -- see fc_oracle.h.  Each function cites the reference lines it
 * restates (paths relative to /root/reference).  The code is written from the algorithm's
 * semantics, not transcribed: the trace is indexed directly instead of qsort+bsearch, the MSA
 * working space is allocated per call instead of a 0.88 GB process-lifetime static, etc.  What is
 * kept deliberately is every observable quirk listed in SURVEY.md 8(a)-notes.
 */
#include "fc_oracle.h"
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "fc_oracle: out of memory\n"); abort(); }
    return p;
}
static void *xrealloc(void *p, size_t sz) {
    p = realloc(p, sz ? sz : 1);
    if (!p) { fprintf(stderr, "fc_oracle: out of memory\n"); abort(); }
    return p;
}
void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * K-mer index of the seed.  Restates allocate_kmer_lookup / add_sequence
 * (src/c/kmer_lookup.c:71-88, 140-192): per k-mer the first position and a "next occurrence"
 * chain in ascending position order; positions 0 .. len-K-1 are indexed (loop bound `i < len-K`,
 * kmer_lookup.c:174, i.e. the final k-mer is NOT indexed).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    unsigned K;
    int *first; /* 4^K entries, INT_MAX = absent     (kmer_lookup.start) */
    int *last;  /*                                    (kmer_lookup.last)  */
    int *next;  /* len entries, 0 = end of chain      (seq_addr_array)    */
} seed_index;

static int base_code(char c, int dflt) { /* kmer_lookup.c:159-171 / 234-246 */
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return dflt;
    }
}

static seed_index *seed_index_build(const char *seed, int len, unsigned K) {
    seed_index *ix = xcalloc(1, sizeof *ix);
    size_t nk = (size_t)1 << (2 * K);
    unsigned mask = (unsigned)(nk - 1);
    ix->K = K;
    ix->first = xcalloc(nk, sizeof(int));
    ix->last = xcalloc(nk, sizeof(int));
    ix->next = xcalloc((size_t)len, sizeof(int));
    for (size_t i = 0; i < nk; i++) ix->first[i] = ix->last[i] = INT_MAX;
    if (len < (int)K) return ix; /* reference: unsigned wrap -> out-of-bounds (UB); we index nothing */
    unsigned char *code = xcalloc((size_t)len, 1);
    for (int i = 0; i < len; i++) code[i] = (unsigned char)base_code(seed[i], 0xff); /* :102-107,158-172 */
    unsigned kv = 0;
    for (unsigned i = 0; i < K; i++) kv = (kv << 2) | (code[i] & 3u);               /* :121-138 */
    for (int i = 0; i < len - (int)K; i++) {                                          /* :174-191 */
        if (ix->first[kv] == INT_MAX) {
            ix->first[kv] = i;
        } else {
            ix->next[ix->last[kv]] = i;
        }
        ix->last[kv] = i;
        kv = ((kv << 2) | code[i + K]) & mask; /* raw OR of the code byte, as the reference does */
    }
    free(code);
    return ix;
}
static void seed_index_free(seed_index *ix) {
    free(ix->first); free(ix->last); free(ix->next); free(ix);
}

/* find_kmer_pos_for_seq (kmer_lookup.c:207-286): every (K/2)-th k-mer of the read, all seed
 * occurrences in ascending seed position. */
typedef struct { int n, cap; int *q, *t; } match_list;
static void match_push(match_list *m, int q, int t) {
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 4096;
        m->q = xrealloc(m->q, (size_t)m->cap * sizeof(int));
        m->t = xrealloc(m->t, (size_t)m->cap * sizeof(int));
    }
    m->q[m->n] = q; m->t[m->n] = t; m->n++;
}
static void kmer_matches(const seed_index *ix, const char *read, int len, match_list *m) {
    unsigned K = ix->K;
    m->n = 0;
    if (len < (int)K) return; /* reference UB, see above */
    for (int i = 0; i < len - (int)K; i += (int)(K >> 1)) {               /* :252 */
        unsigned kv = 0;
        for (unsigned b = 0; b < K; b++) kv = (kv << 2) | ((unsigned)base_code(read[i + b], 0) & 3u);
        int pos = ix->first[kv];
        if (pos == INT_MAX) continue;
        for (;;) {                                                        /* :257-282 */
            match_push(m, i, pos);
            int nx = ix->next[pos];
            if (nx <= pos) break;
            pos = nx;
        }
    }
}

/* find_best_aln_range (kmer_lookup.c:294-427), called with (K, 6K, 5) from falcon.c:604 */
static void best_range(const match_list *m, int bin_size, int count_th, orc_pair_info *o) {
    o->s1 = o->e1 = o->s2 = o->e2 = 0; o->score = 0;
    if (m->n == 0) return; /* reference: overflowed calloc size -> NULL -> all loops empty -> zeros */
    long dmin = LONG_MAX, dmax = LONG_MIN;
    for (int i = 0; i < m->n; i++) {
        long d = (long)m->q[i] - (long)m->t[i];
        if (d < dmin) dmin = d;
        if (d > dmax) dmax = d;
    }
    long nbin = (dmax - dmin) / bin_size + 1;
    int *hist = xcalloc((size_t)nbin, sizeof(int));
    for (int i = 0; i < m->n; i++) hist[((long)m->q[i] - m->t[i] - dmin) / bin_size]++;
    long top = 0, top_bin = INT_MAX;                                      /* :357-366 */
    for (int i = 0; i < m->n; i++) {
        long b = ((long)m->q[i] - m->t[i] - dmin) / bin_size;
        if (hist[b] > top) { top = hist[b]; top_bin = b; }
    }
    int *kq = xcalloc((size_t)m->n, sizeof(int)), *kt = xcalloc((size_t)m->n, sizeof(int));
    int nk = 0;
    if (top_bin != INT_MAX && top > count_th) {                           /* :369-383 */
        for (int i = 0; i < m->n; i++) {
            long b = ((long)m->q[i] - m->t[i] - dmin) / bin_size;
            if (labs(b - top_bin) > 5) continue;
            if (hist[b] > count_th) { kq[nk] = m->q[i]; kt[nk] = m->t[i]; nk++; }
        }
    }
    if (nk > 1) {                                                         /* :385-411 */
        o->s1 = o->e1 = kq[0]; o->s2 = o->e2 = kt[0];
        long cur = 0, best = 0; int start = 0;
        for (int i = 1; i < nk; i++) {
            cur += 32 - (kq[i] - kq[i - 1]);
            if (cur < 0) { cur = 0; start = i; }
            else if (cur > best) {
                o->s1 = kq[start]; o->s2 = kt[start]; o->e1 = kq[i]; o->e2 = kt[i];
                best = cur; o->score = best;
            }
        }
    }
    free(hist); free(kq); free(kt);
}

void orc_kmer_range(const char *read, int rlen, const char *seed, int slen, unsigned K,
                    orc_pair_info *out) {
    seed_index *ix = seed_index_build(seed, slen, K);
    match_list m = {0, 0, NULL, NULL};
    kmer_matches(ix, read, rlen, &m);
    memset(out, 0, sizeof *out);
    out->n_match = m.n;
    best_range(&m, (int)K * 6, 5, out);
    free(m.q); free(m.t);
    seed_index_free(ix);
}

/* ------------------------------------------------------------------------------------------
 * Banded O(ND) alignment.  Restates align() (src/c/DW_banded.c:115-330).
 * Furthest-reaching x per diagonal k, indel-only edit graph; band kept within band_tolerance of
 * the best x+y seen (DW_banded.c:227-243); abort when the band is wider than 2*band_tolerance
 * (:184-186) or d reaches max_d = (int)(0.3*(q_len+t_len)) (:149,183).
 * Trace: per d the evaluated cells in ascending k (so (d,k) is found by index, replacing the
 * reference's qsort/bsearch of d_path_data2 records, :260-268).
 * ---------------------------------------------------------------------------------------- */
typedef struct { int x1, x2, pre_k; } cell_t;

int orc_align(const char *q, int q_len, const char *t, int t_len, int band_tolerance,
              int *dist, int *q_e, int *t_e, char *q_aln, char *t_aln, long *cells_out) {
    int max_d = (int)(0.3 * (q_len + t_len));
    int band_size = band_tolerance * 2;
    int *V = xcalloc((size_t)max_d * 2 + 2, sizeof(int));
    int *U = xcalloc((size_t)max_d * 2 + 2, sizeof(int));
    int off = max_d;
    /* per-d directory of the trace */
    long *d_first = xcalloc((size_t)max_d + 1, sizeof(long));
    int *d_mink = xcalloc((size_t)max_d + 1, sizeof(int));
    size_t cap = 1 << 16, ncell = 0;
    cell_t *cells = xcalloc(cap, sizeof(cell_t));

    int best_m = -1, min_k = 0, max_k = 0;
    int aligned = 0, end_d = 0, end_k = 0, x = 0, y = 0;
    *dist = 0; *q_e = 0; *t_e = 0;
    for (int d = 0; d < max_d; d++) {
        if (max_k - min_k > band_size) break;
        d_first[d] = (long)ncell; d_mink[d] = min_k;
        int k;
        for (k = min_k; k <= max_k; k += 2) {
            int pre_k;
            if (k == min_k || (k != max_k && V[k - 1 + off] < V[k + 1 + off])) {
                pre_k = k + 1; x = V[k + 1 + off];
            } else {
                pre_k = k - 1; x = V[k - 1 + off] + 1;
            }
            y = x - k;
            if (ncell == cap) { cap *= 2; cells = xrealloc(cells, cap * sizeof(cell_t)); }
            cells[ncell].x1 = x; cells[ncell].pre_k = pre_k;
            while (x < q_len && y < t_len && q[x] == t[y]) { x++; y++; }
            cells[ncell].x2 = x; ncell++;
            V[k + off] = x; U[k + off] = x + y;
            if (x + y > best_m) best_m = x + y;
            if (x >= q_len || y >= t_len) { aligned = 1; end_d = d; end_k = k; break; }
        }
        if (aligned) break; /* the band update that follows in the reference no longer matters */
        int nmin = max_k, nmax = min_k;
        for (int k2 = min_k; k2 <= max_k; k2 += 2) {
            if (U[k2 + off] >= best_m - band_tolerance) {
                if (k2 < nmin) nmin = k2;
                if (k2 > nmax) nmax = k2;
            }
        }
        max_k = nmax + 1; min_k = nmin - 1;
    }
    if (cells_out) *cells_out = (long)ncell;
    int aln = 0;
    if (aligned) {
        *q_e = x; *t_e = y; *dist = end_d;
        /* traceback (:264-277): collect (x1,y1),(x2,y2) per d from the end cell back to d = 0 */
        int npt = 0;
        int (*pt)[2] = xcalloc((size_t)2 * (end_d + 1) + 2, sizeof *pt);
        int ck = end_k;
        for (int cd = end_d; cd >= 0 && npt < q_len + t_len + 1; cd--) {
            const cell_t *c = &cells[d_first[cd] + (ck - d_mink[cd]) / 2];
            pt[npt][0] = c->x2; pt[npt][1] = c->x2 - ck; npt++;
            pt[npt][0] = c->x1; pt[npt][1] = c->x1 - ck; npt++;
            ck = c->pre_k;
        }
        /* forward walk (:278-319) */
        npt--;
        int cx = pt[npt][0], cy = pt[npt][1];
        while (npt > 0) {
            npt--;
            int nx = pt[npt][0], ny = pt[npt][1];
            if (nx == cx && ny == cy) continue;
            if (nx == cx) {                 /* target-only columns */
                for (int i = 0; i < ny - cy; i++) {
                    if (q_aln) { q_aln[aln + i] = '-'; t_aln[aln + i] = t[cy + i]; }
                }
                aln += ny - cy;
            } else if (ny == cy) {          /* query-only columns */
                for (int i = 0; i < nx - cx; i++) {
                    if (q_aln) { q_aln[aln + i] = q[cx + i]; t_aln[aln + i] = '-'; }
                }
                aln += nx - cx;
            } else {                        /* snake */
                for (int i = 0; i < nx - cx; i++) {
                    if (q_aln) { q_aln[aln + i] = q[cx + i]; t_aln[aln + i] = t[cy + i]; }
                }
                aln += ny - cy;
            }
            cx = nx; cy = ny;
        }
        if (q_aln) { q_aln[aln] = 0; t_aln[aln] = 0; }
        free(pt);
    }
    free(V); free(U); free(d_first); free(d_mink); free(cells);
    return aln;
}

/* ------------------------------------------------------------------------------------------
 * Alignment tags.  Restates get_align_tags (src/c/falcon.c:106-162) with t_offset = 0.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int t_pos; uint8_t delta; char q_base; int p_t_pos; uint8_t p_delta; char p_q_base; } tag_t;
typedef struct { int len; tag_t *tags; } tag_list;

static tag_list make_tags(const char *qa, const char *ta, int n, int s1, int s2) {
    tag_list tl; tl.tags = xcalloc((size_t)n + 1, sizeof(tag_t));
    int i = s1 - 1, j = s2 - 1, jj = 0, p_j = -1, p_jj = 0; char p_b = '.';
    (void)i;
    int k;
    for (k = 0; k < n; k++) {
        if (qa[k] != '-') { i++; jj++; }
        if (ta[k] != '-') { j++; jj = 0; }
        if (j >= 0 && jj < UINT8_MAX && p_jj < UINT8_MAX) {
            tag_t *g = &tl.tags[k];
            g->t_pos = j; g->delta = (uint8_t)jj; g->q_base = qa[k];
            g->p_t_pos = p_j; g->p_delta = (uint8_t)p_jj; g->p_q_base = p_b;
            p_j = j; p_jj = jj; p_b = qa[k];
        } else break;
    }
    tl.len = k;
    return tl;
}

/* ------------------------------------------------------------------------------------------
 * Column store + vote + longest path + backtrack.
 * Restates get_cns_from_align_tags (src/c/falcon.c:308-558) and its store (:64-104,170-303).
 * The per-position delta-group keeps the reference's uint8 `size` / `max_delta` fields and its
 * growth rule (:363-368, 205-218) because they are observable: once max_delta+8 exceeds 255 the
 * stored size wraps and the next growth re-initialises (wipes) live delta columns.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int pt; uint8_t pd; char pb; uint16_t cnt; } link_t;
typedef struct {
    link_t *lk; uint16_t n_link, cap; uint16_t count;
    int best_pt; uint8_t best_pd, best_pb; double score;
} col_t;
typedef struct { col_t (*d)[5]; uint8_t size, max_delta; int alloc; } pos_t;

static void pos_fresh_cols(pos_t *p, int from, int to) { /* allocate_delta_group / realloc_delta_group */
    for (int i = from; i < to; i++)
        for (int b = 0; b < 5; b++) {
            free(p->d[i][b].lk); /* the reference leaks these; same observable state */
            memset(&p->d[i][b], 0, sizeof(col_t));
        }
}
static void pos_init(pos_t *p) {
    p->size = 8; p->max_delta = 0; p->alloc = 8;
    p->d = xcalloc(8, sizeof *p->d);
}
static void pos_grow(pos_t *p, unsigned new_size) {      /* falcon.c:205-218 */
    if ((int)new_size > p->alloc) {
        p->d = xrealloc(p->d, new_size * sizeof *p->d);
        memset(p->d + p->alloc, 0, (new_size - (unsigned)p->alloc) * sizeof *p->d);
        p->alloc = (int)new_size;
    }
    pos_fresh_cols(p, p->size, (int)new_size);           /* i = bs .. es-1 re-initialised */
    p->size = (uint8_t)new_size;                         /* uint8 truncation, as in the reference */
}
static void col_vote(col_t *c, int pt, uint8_t pd, char pb) { /* update_col falcon.c:232-263 */
    c->count++;
    for (int k = 0; k < c->n_link; k++)
        if (c->lk[k].pt == pt && c->lk[k].pd == pd && c->lk[k].pb == pb) { c->lk[k].cnt++; return; }
    if (c->n_link == c->cap) {
        c->cap = c->cap ? (uint16_t)(c->cap * 2) : 8;
        c->lk = xrealloc(c->lk, c->cap * sizeof(link_t));
    }
    c->lk[c->n_link].pt = pt; c->lk[c->n_link].pd = pd; c->lk[c->n_link].pb = pb;
    c->lk[c->n_link].cnt = 1; c->n_link++;
}
static int code5(char c, int dflt) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
                 case '-': return 4; default: return dflt; }
}

static char *consensus_from_tags(tag_list *reads, unsigned n_reads, unsigned t_len, unsigned min_cov,
                                 int **eqv_out) {
    unsigned *coverage = xcalloc(t_len, sizeof(unsigned));
    pos_t *msa = xcalloc((size_t)t_len + 1, sizeof(pos_t));
    for (unsigned i = 0; i <= t_len; i++) pos_init(&msa[i]);

    /* vote (falcon.c:350-382) */
    int t_pos = 0;
    for (unsigned r = 0; r < n_reads; r++) {
        for (int j = 0; j < reads[r].len; j++) {
            const tag_t *g = &reads[r].tags[j];
            unsigned delta = g->delta;
            if (delta == 0) { t_pos = g->t_pos; coverage[t_pos]++; }
            pos_t *p = &msa[t_pos];
            if (delta > p->max_delta) {
                p->max_delta = (uint8_t)delta;
                if (p->max_delta + 4 > p->size) pos_grow(p, (unsigned)p->max_delta + 8);
            }
            int b = code5(g->q_base, -1);
            if (b < 0) { fprintf(stderr, "fc_oracle: non-ACGT base in alignment (reference UB, falcon.c:370-379)\n"); abort(); }
            col_vote(&p->d[delta][b], g->p_t_pos, g->p_delta, g->p_q_base);
        }
    }

    /* longest path over the link DAG (falcon.c:385-477) */
    col_t *g_col = NULL; unsigned g_ck = 0; int g_t = 0; double g_best = -1;
    int best_ck = -1;
    for (unsigned i = 0; i < t_len; i++) {
        for (unsigned j = 0; j <= msa[i].max_delta; j++) {
            for (int kk = 0; kk < 5; kk++) {
                col_t *c = &msa[i].d[j][kk];
                double best = -1;
                for (int ck = 0; ck < c->n_link; ck++) {
                    int pi = c->lk[ck].pt, pj = c->lk[ck].pd, pkk = code5(c->lk[ck].pb, 4);
                    double sc = (double)c->lk[ck].cnt - (double)coverage[i] * 0.5;
                    if (pi != -1) sc += msa[pi].d[pj][pkk].score;
                    if (sc > best) {
                        best = sc; c->best_pt = pi; c->best_pd = (uint8_t)pj; c->best_pb = (uint8_t)pkk;
                        best_ck = ck;
                    }
                }
                c->score = best;
                if (best > g_best) { g_best = best; g_col = c; g_ck = (unsigned)best_ck; g_t = (int)i; }
            }
        }
    }
    if (g_best == -1) { fprintf(stderr, "fc_oracle: no best score (reference asserts, falcon.c:476)\n"); abort(); }

    /* backtrack (falcon.c:479-542).  NOTE the reference seeds `ck` with the best LINK index, not a
     * base code, so the last consensus base is "ACGT-"[link index] (or a stale '$'). */
    char *cns = xcalloc((size_t)t_len * 2 + 1, 1);
    int *eqv = xcalloc((size_t)t_len * 2 + 1, sizeof(int));
    unsigned index = 0; char bb = '$'; int ck = (int)g_ck; int i = g_t; col_t *c = g_col;
    for (;;) {
        if (coverage[i] > min_cov) {
            switch (ck) { case 0: bb = 'A'; break; case 1: bb = 'C'; break; case 2: bb = 'G'; break;
                          case 3: bb = 'T'; break; case 4: bb = '-'; break; }
        } else {
            switch (ck) { case 0: bb = 'a'; break; case 1: bb = 'c'; break; case 2: bb = 'g'; break;
                          case 3: bb = 't'; break; case 4: bb = '-'; break; }
        }
        double score0 = c->score;
        i = c->best_pt;
        if (i == -1 || index >= t_len * 2) break;
        int j = c->best_pd; ck = c->best_pb;
        c = &msa[i].d[j][ck];
        if (bb != '-') { cns[index] = bb; eqv[index] = (int)score0 - (int)c->score; index++; }
    }
    for (unsigned a = 0; a < index / 2; a++) {
        char tc = cns[a]; cns[a] = cns[index - a - 1]; cns[index - a - 1] = tc;
        int te = eqv[a]; eqv[a] = eqv[index - a - 1]; eqv[index - a - 1] = te;
    }
    cns[index] = 0;

    for (unsigned p = 0; p <= t_len; p++) {
        for (int d = 0; d < msa[p].alloc; d++)
            for (int b = 0; b < 5; b++) free(msa[p].d[d][b].lk);
        free(msa[p].d);
    }
    free(msa); free(coverage);
    if (eqv_out) *eqv_out = eqv; else free(eqv);
    return cns;
}

/* ------------------------------------------------------------------------------------------
 * Seed-block driver.  Restates generate_consensus (src/c/falcon.c:562-666).
 * ---------------------------------------------------------------------------------------- */
char *orc_generate_consensus(const char **seqs, unsigned n_seq, unsigned min_cov, unsigned K,
                             double min_idt, orc_pair_info *info, int **eqv_out) {
    double max_diff = 1.0 - min_idt;
    int slen = (int)strlen(seqs[0]);
    seed_index *ix = seed_index_build(seqs[0], slen, K);
    tag_list *tl = xcalloc(n_seq, sizeof(tag_list));
    unsigned n_aln = 0;
    match_list m = {0, 0, NULL, NULL};
    for (unsigned j = 1; j < n_seq; j++) {
        orc_pair_info pi; memset(&pi, 0, sizeof pi);
        int rlen = (int)strlen(seqs[j]);
        kmer_matches(ix, seqs[j], rlen, &m);
        pi.n_match = m.n;
        best_range(&m, (int)K * 6, 5, &pi);
        int sp1 = pi.e1 - pi.s1, sp2 = pi.e2 - pi.s2;
        if (!(sp1 < 100 || sp2 < 100 || abs(sp1 - sp2) > (int)(0.5 * 0.10 * (sp1 + sp2)))) { /* :613-615 */
            pi.passed_filter = 1;
            char *qa = xcalloc((size_t)sp1 + sp2 + 1, 1), *ta = xcalloc((size_t)sp1 + sp2 + 1, 1);
            pi.aln_size = orc_align(seqs[j] + pi.s1, sp1, seqs[0] + pi.s2, sp2, 150,
                                    &pi.dist, &pi.q_e, &pi.t_e, qa, ta, &pi.trace_cells);
            pi.aligned = pi.aln_size > 0;
            if (pi.aln_size > 500 && ((double)pi.dist / (double)pi.aln_size) < max_diff) {    /* :629 */
                pi.accepted = 1;
                tl[n_aln] = make_tags(qa, ta, pi.aln_size, pi.s1, pi.s2);
                pi.n_tags = tl[n_aln].len;
                n_aln++;
            }
            free(qa); free(ta);
        }
        if (info) info[j] = pi;
    }
    char *cns;
    if (n_aln > 0) {
        cns = consensus_from_tags(tl, n_aln, (unsigned)slen, min_cov, eqv_out);
    } else {
        cns = xcalloc(1, 1);
        if (eqv_out) *eqv_out = xcalloc(1, sizeof(int));
    }
    for (unsigned a = 0; a < n_aln; a++) free(tl[a].tags);
    free(tl); free(m.q); free(m.t);
    seed_index_free(ix);
    return cns;
}
