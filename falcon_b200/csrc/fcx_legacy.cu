// fcx_legacy.cu -- the reference's single-call C symbols (falcon_kit/falcon_kit.py:54-122) on top of
// the GPU engine.
//
//   generate_consensus / free_consensus_data : GPU, batch of one seed block (fcx_consensus_blocks).
//   align / free_alignment                   : GPU, batch of one pair (fcx_align_pairs path).
//   allocate_* / init_* / free_* / add_sequence / mask_k_mer / find_kmer_pos_for_seq /
//   find_best_aln_range / find_best_aln_range2 : these operate on CALLER-VISIBLE HOST STRUCTURES
//   (kmer_lookup[], seq_addr_array, kmer_match) whose layout is the ABI itself, so they are host
//   code by definition; they are API surface for the --trim path (consensus.py:48-99), not part of
//   the generate_consensus hot path, which never calls them.
#include "../../include/falcon_b200.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

extern "C" int fcx_internal_want_eqv(fcx_ctx*, int on);
extern "C" int fcx_internal_last_eqv(fcx_ctx*, const int32_t** eqv, uint64_t* n);
extern "C" int fcx_internal_align(fcx_ctx*, const char* q, int q_len, const char* t, int t_len,
                                  int band_tolerance, int get_aln_str, alignment* out);

namespace {
std::mutex g_mu;
fcx_ctx* g_ctx = nullptr;

fcx_ctx* default_ctx() {
    if (!g_ctx) {
        int dev = 0;
        if (const char* s = getenv("FCX_DEVICE")) dev = atoi(s);
        if (fcx_create(dev, &g_ctx) != 0) {
            fprintf(stderr, "CRITICAL ERROR: falcon_b200: %s\n", fcx_last_error(nullptr));
            abort();   // the reference has no error channel here either (it abort()s on failure)
        }
    }
    return g_ctx;
}
[[noreturn]] void die(fcx_ctx* c, const char* what) {
    fprintf(stderr, "CRITICAL ERROR: falcon_b200 %s: %s\n", what, fcx_last_error(c));
    abort();
}
}  // namespace

// ------------------------------------------------------------------ generate_consensus
extern "C" consensus_data* generate_consensus(char** input_seq, unsigned int n_seq, unsigned min_cov,
                                              unsigned K, double min_idt) {
    std::lock_guard<std::mutex> lk(g_mu);
    fflush(stdout);                                   // falcon.c:587
    fcx_ctx* c = default_ctx();
    std::vector<uint64_t> off(n_seq + 1, 0);
    for (unsigned i = 0; i < n_seq; i++) off[i + 1] = off[i] + strlen(input_seq[i]);
    std::vector<char> cat(off[n_seq] + 1);
    for (unsigned i = 0; i < n_seq; i++) memcpy(cat.data() + off[i], input_seq[i], off[i + 1] - off[i]);
    if (fcx_pool_upload(c, cat.data(), off.data(), n_seq)) die(c, "generate_consensus(pool)");
    std::vector<uint32_t> ids(n_seq);
    for (unsigned i = 0; i < n_seq; i++) ids[i] = i;
    uint32_t boff[2] = {0, n_seq};
    const char* bases = nullptr; const uint64_t* ooff = nullptr;
    fcx_internal_want_eqv(c, 1);
    if (fcx_consensus_blocks(c, 1, boff, ids.data(), min_cov, K, min_idt, &bases, &ooff)) die(c, "generate_consensus");
    size_t len = (size_t)(ooff[1] - ooff[0]);
    consensus_data* cd = (consensus_data*)calloc(1, sizeof(consensus_data));
    cd->sequence = (char*)calloc(len + 1, 1);
    cd->eqv = (int*)calloc(len + 1, sizeof(int));
    memcpy(cd->sequence, bases + ooff[0], len);
    const int32_t* eqv = nullptr; uint64_t ne = 0;
    fcx_internal_last_eqv(c, &eqv, &ne);
    if (eqv && ne >= len) memcpy(cd->eqv, eqv, len * sizeof(int));
    return cd;
}

extern "C" void free_consensus_data(consensus_data* cd) {   // falcon.c:776-780
    if (!cd) return;
    free(cd->sequence); free(cd->eqv); free(cd);
}

// ------------------------------------------------------------------ align
extern "C" alignment* align(char* query_seq, seq_coor_t q_len, char* target_seq, seq_coor_t t_len,
                            seq_coor_t band_tolerance, int get_aln_str) {
    std::lock_guard<std::mutex> lk(g_mu);
    fcx_ctx* c = default_ctx();
    alignment* a = (alignment*)calloc(1, sizeof(alignment));
    a->q_aln_str = (char*)calloc((size_t)q_len + t_len + 1, 1);   // DW_banded.c:169-170
    a->t_aln_str = (char*)calloc((size_t)q_len + t_len + 1, 1);
    if (fcx_internal_align(c, query_seq, q_len, target_seq, t_len, band_tolerance, get_aln_str, a)) die(c, "align");
    return a;
}
extern "C" void free_alignment(alignment* a) {                   // DW_banded.c:333-337
    if (!a) return;
    free(a->q_aln_str); free(a->t_aln_str); free(a);
}

// ------------------------------------------------------------------ host k-mer structures
extern "C" kmer_lookup* allocate_kmer_lookup(seq_coor_t size) {
    kmer_lookup* kl = (kmer_lookup*)malloc((size_t)size * sizeof(kmer_lookup));
    init_kmer_lookup(kl, size);
    return kl;
}
extern "C" void init_kmer_lookup(kmer_lookup* kl, seq_coor_t size) {
    for (seq_coor_t i = 0; i < size; i++) { kl[i].start = INT_MAX; kl[i].last = INT_MAX; kl[i].count = 0; }
}
extern "C" void free_kmer_lookup(kmer_lookup* p) { free(p); }
extern "C" seq_array allocate_seq(seq_coor_t size) {
    seq_array sa = (seq_array)malloc((size_t)size);
    init_seq_array(sa, size);
    return sa;
}
extern "C" void init_seq_array(seq_array sa, seq_coor_t size) { memset(sa, 0xff, (size_t)size); }
extern "C" void free_seq_array(seq_array sa) { free(sa); }
extern "C" seq_addr_array allocate_seq_addr(seq_coor_t size) { return (seq_addr_array)calloc((size_t)size, sizeof(seq_addr)); }
extern "C" void free_seq_addr_array(seq_addr_array p) { free(p); }

static inline int code_of(char ch, int dflt) {
    return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : dflt;
}

extern "C" void add_sequence(seq_coor_t start, unsigned int K, char* seq, seq_coor_t seq_len,
                             seq_addr_array sda, seq_array sa, kmer_lookup* lk) {
    const unsigned mask = (K >= 16) ? 0xffffffffu : ((1u << (2 * K)) - 1u);
    for (seq_coor_t i = 0; i < seq_len; i++) { int c = code_of(seq[i], -1); if (c >= 0) sa[start + i] = (base)c; }
    if (seq_len < (seq_coor_t)K) return;
    unsigned kv = 0;
    for (unsigned i = 0; i < K; i++) kv = (kv << 2) | (sa[start + i] & 3u);
    for (seq_coor_t i = 0; i + (seq_coor_t)K < seq_len; i++) {
        kmer_lookup& e = lk[kv];
        if (e.start == INT_MAX) e.start = start + i; else sda[e.last] = start + i;
        e.last = start + i; e.count += 1;
        kv = ((kv << 2) | sa[start + i + K]) & mask;
    }
}

extern "C" void mask_k_mer(seq_coor_t size, kmer_lookup* kl, seq_coor_t threshold) {
    for (seq_coor_t i = 0; i < size; i++)
        if (kl[i].count > threshold) { kl[i].start = INT_MAX; kl[i].last = INT_MAX; }
}

extern "C" kmer_match* find_kmer_pos_for_seq(char* seq, seq_coor_t seq_len, unsigned int K,
                                             seq_addr_array sda, kmer_lookup* lk) {
    std::vector<seq_coor_t> qv, tv;
    const int step = (int)(K >> 1);
    for (seq_coor_t i = 0; i + (seq_coor_t)K < seq_len && step > 0; i += step) {
        unsigned kv = 0;
        for (unsigned b = 0; b < K; b++) kv = (kv << 2) | ((unsigned)code_of(seq[i + b], 0) & 3u);
        seq_coor_t pos = lk[kv].start;
        if (pos == INT_MAX) continue;
        for (;;) {
            qv.push_back(i); tv.push_back(pos);
            seq_coor_t nx = sda[pos];
            if (nx <= pos) break;
            pos = nx;
        }
    }
    kmer_match* m = (kmer_match*)malloc(sizeof(kmer_match));
    m->count = (seq_coor_t)qv.size();
    m->query_pos = (seq_coor_t*)calloc(qv.size() + 1, sizeof(seq_coor_t));
    m->target_pos = (seq_coor_t*)calloc(qv.size() + 1, sizeof(seq_coor_t));
    if (!qv.empty()) {
        memcpy(m->query_pos, qv.data(), qv.size() * sizeof(seq_coor_t));
        memcpy(m->target_pos, tv.data(), tv.size() * sizeof(seq_coor_t));
    }
    return m;
}
extern "C" void free_kmer_match(kmer_match* m) { if (!m) return; free(m->query_pos); free(m->target_pos); free(m); }

extern "C" aln_range* find_best_aln_range(kmer_match* km, seq_coor_t /*K*/, seq_coor_t bin_size, seq_coor_t count_th) {
    aln_range* ar = (aln_range*)calloc(1, sizeof(aln_range));
    const int n = km->count;
    if (n <= 0) return ar;
    long lo = LONG_MAX, hi = LONG_MIN;
    for (int i = 0; i < n; i++) { long d = (long)km->query_pos[i] - km->target_pos[i]; lo = d < lo ? d : lo; hi = d > hi ? d : hi; }
    std::vector<int> hist((size_t)((hi - lo) / bin_size + 1), 0);
    auto bin = [&](int i) { return ((long)km->query_pos[i] - km->target_pos[i] - lo) / bin_size; };
    for (int i = 0; i < n; i++) hist[bin(i)]++;
    long top = 0, top_bin = -1;
    for (int i = 0; i < n; i++) if (hist[bin(i)] > top) { top = hist[bin(i)]; top_bin = bin(i); }
    if (top_bin < 0 || top <= count_th) return ar;
    std::vector<int> kq, kt;
    for (int i = 0; i < n; i++) {
        long b = bin(i);
        if (labs(b - top_bin) > 5) continue;
        if (hist[b] > count_th) { kq.push_back(km->query_pos[i]); kt.push_back(km->target_pos[i]); }
    }
    if (kq.size() > 1) {
        ar->s1 = ar->e1 = kq[0]; ar->s2 = ar->e2 = kt[0];
        long cur = 0, best = 0; size_t from = 0;
        for (size_t i = 1; i < kq.size(); i++) {
            cur += 32 - (kq[i] - kq[i - 1]);
            if (cur < 0) { cur = 0; from = i; }
            else if (cur > best) { best = cur; ar->s1 = kq[from]; ar->s2 = kt[from]; ar->e1 = kq[i]; ar->e2 = kt[i]; ar->score = best; }
        }
    }
    return ar;
}

// kmer_lookup.c:429-585 -- used only by the Python-level get_alignment of the --trim path
extern "C" aln_range* find_best_aln_range2(kmer_match* km, seq_coor_t /*K*/, seq_coor_t /*bin_width*/, seq_coor_t /*count_th*/) {
    aln_range* ar = (aln_range*)calloc(1, sizeof(aln_range));
    const int n = km->count;
    if (n <= 0) return ar;
    std::vector<int> dg((size_t)n);
    int max_q = -1, max_t = -1;
    for (int i = 0; i < n; i++) {
        dg[i] = km->query_pos[i] - km->target_pos[i];
        max_q = max_q > km->query_pos[i] ? max_q : km->query_pos[i];
        max_t = max_t > km->target_pos[i] ? max_q : km->target_pos[i];   // sic (kmer_lookup.c:458)
    }
    std::sort(dg.begin(), dg.end());
    int s = 0, e = 0, max_s = -1, max_e = -1, max_span = -1;
    const int delta = (int)(long)(0.05 * (max_q + max_t));
    for (;;) {
        int d_s = dg[s], d_e = dg[e];
        while (d_e < d_s + delta && e < n - 1) { e++; d_e = dg[e]; }
        if (max_span == -1 || e - s > max_span) { max_span = e - s; max_s = s; max_e = e; }
        s++;
        if (s == n || e == n) break;
    }
    if (max_s == -1 || max_e == -1 || max_e - max_s < 32) return ar;
    const int d_lo = dg[max_s], d_hi = dg[max_e];
    std::vector<int> last_hit((size_t)n, -1), hit_score((size_t)n, 0), hit_count((size_t)n, 0);
    int best_idx = -1, best_score = 0, best_count = 0;
    for (int i = 0; i < n; i++) {
        int cx = km->query_pos[i], cy = km->target_pos[i];
        int d = cx - cy;
        if (d < d_lo || d > d_hi) continue;
        int cand = -1, max_d = 65535;
        for (int j = i - 1; j >= 0; j--) {
            int px = km->query_pos[j], py = km->target_pos[j];
            int dj = px - py;
            if (dj < d_lo || dj > d_hi) continue;
            if (cx - px > 320) break;
            if (cy > py && cx - px + cy - py < max_d && cy - py <= 320) { max_d = cx - px + cy - py; cand = j; }
        }
        if (cand != -1) {
            last_hit[i] = cand;
            hit_score[i] = hit_score[cand] + (64 - max_d);
            hit_count[i] = hit_count[cand] + 1;
            if (hit_score[i] < 0) { hit_score[i] = 0; hit_count[i] = 0; }
        }
        if (hit_score[i] > best_score) { best_score = hit_score[i]; best_count = hit_count[i]; best_idx = i; }
    }
    if (best_idx == -1) return ar;
    ar->score = best_count + 1;
    ar->e1 = km->query_pos[best_idx]; ar->e2 = km->target_pos[best_idx];
    int i = best_idx;
    while (last_hit[i] != -1) i = last_hit[i];
    ar->s1 = km->query_pos[i]; ar->s2 = km->target_pos[i];
    return ar;
}
extern "C" void free_aln_range(aln_range* p) { free(p); }
