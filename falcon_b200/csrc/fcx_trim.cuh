// fcx_trim.cuh -- device side of the --trim path (ref: falcon_kit/mains/consensus.py:48-99,123-158):
//   k_trim_range  per (read, seed) pair: k-mer hits against the seed index with high-count k-mers
//                 masked (mask_k_mer(.., 16), src/c/kmer_lookup.c:195-204 and :253) and the sparse
//                 chaining DP of find_best_aln_range2(K, 400, 25) (kmer_lookup.c:429-585);
//   k_subreads    cuts [s, s + len) out of packed reads into new packed reads (the trimmed reads
//                 that get_consensus_with_trim hands to generate_consensus).
// One warp per pair, persistent warps pulling pairs from a device counter; every warp owns a slice
// of global scratch (match list, diagonal histogram, in-band list, chain arrays).
//
// find_best_aln_range2 restated:
//   (1) d_coor = sorted diagonals q - t.  The two-pointer scan :468-484 looks for the start s whose
//       window [d_s, d_s + delta) holds the most elements, first maximum winning.  With F(d) = number
//       of matches on diagonals < d the scan's e(s) equals min(n - 1, F(d_s + delta)) (both pointers
//       are monotone), and within a run of equal diagonals the first s has the largest e - s, so only
//       the first element of every distinct diagonal can win: a histogram over exact diagonals and
//       its prefix sum replace the qsort.
//   (2) delta = (long)(0.05 * (max_q + max_t)) where max_t follows the reference's recurrence
//       `max_t = max_t > t_i ? max_q : t_i` (kmer_lookup.c:458, sic): a two-state automaton over
//       the match list (the value is always q_i or t_i), evaluated 32 matches per ballot pair.
//   (3) chaining (:505-547): matches inside the diagonal band, in list order; the predecessor of
//       match i is the in-band match j < i with cx - px <= 320, cy > py, cy - py <= 320 minimising
//       cx - px + cy - py, the LARGEST such j on ties (the reference scans j downwards with a
//       strict '<').  Lanes test 32 candidates at once; key = (distance << 5 | lane) min-reduced.
#pragma once

namespace fcx {

struct TrimOut { int32_t s1, e1, s2, e2, score, n_match; };

constexpr int TRIM_MASK_TH = 16;        // consensus.py:60  mask_k_mer(1 << (K * 2), lk_ptr, 16)
constexpr int TRIM_GAP = 320;           // kmer_lookup.c:529,530
constexpr int TRIM_WARPS = 4;

__global__ void __launch_bounds__(TRIM_WARPS * 32)
k_trim_range(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs, uint32_t n_pairs,
             const uint32_t* __restrict__ pool, const uint32_t* __restrict__ ktab,
             const uint32_t* __restrict__ kpos_arena, uint32_t* __restrict__ scratch, uint32_t list_cap,
             uint32_t hist_cap, uint32_t* __restrict__ next_pair, TrimOut* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * TRIM_WARPS + (threadIdx.x >> 5);
    uint32_t* base = scratch + (size_t)gw * ((size_t)5 * list_cap + hist_cap);
    uint32_t* list = base;                        // all matches, list order
    uint32_t* band = base + list_cap;             // matches inside the diagonal band, list order
    int* sc = reinterpret_cast<int*>(base + 2 * (size_t)list_cap);
    int* cn = reinterpret_cast<int*>(base + 3 * (size_t)list_cap);
    int* root = reinterpret_cast<int*>(base + 4 * (size_t)list_cap);
    int* hist = reinterpret_cast<int*>(base + 5 * (size_t)list_cap);
    const unsigned lt = lanemask_lt();
  for (;;) {
    uint32_t p = 0;
    if (lane == 0) p = atomicAdd(next_pair, 1u);
    p = __shfl_sync(FULL, p, 0);
    if (p >= n_pairs) return;
    __syncwarp();
    const PairDesc pd = pairs[p];
    const BlockDesc bd = blocks[pd.block];
    const uint32_t* read = pool + pd.read_woff;
    const uint32_t* tab = ktab + (size_t)pd.block * KTAB;
    const uint32_t* kpos = kpos_arena + bd.kpos_off;
    const int nq = pd.rlen > KMER ? (pd.rlen - KMER + 3) / 4 : 0;      // i = 0,4,.. < rlen-K
    TrimOut r; r.s1 = r.e1 = r.s2 = r.e2 = r.score = 0; r.n_match = 0;
    // ---- match list with masked buckets (kmer_lookup.c:252-283)
    int n = 0, dmin = INT_MAX, dmax = INT_MIN;
    for (int it0 = 0; it0 < nq; it0 += 32) {
        const int it = it0 + lane, i = it * 4;
        uint32_t s = 0, e = 0;
        if (it < nq) {
            const uint32_t kid = fetch16(read, i) & 0xffffu;
            s = kid ? __ldg(tab + kid - 1) : 0u; e = __ldg(tab + kid);
            if (e - s > (uint32_t)TRIM_MASK_TH) e = s;                   // high-count k-mer: masked
        }
        const int c = (int)(e - s);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(FULL, incl, 31);
        int w = n + incl - c;
        for (uint32_t j = s; j < e; j++, w++) {
            const int t = (int)__ldg(kpos + j);
            list[w] = rm_pack(i, t);
            dmin = min(dmin, i - t); dmax = max(dmax, i - t);
        }
        n += total;
    }
    r.n_match = n;
    if (n == 0) { if (lane == 0) out[p] = r; continue; }
    dmin = __reduce_min_sync(FULL, dmin); dmax = __reduce_max_sync(FULL, dmax);
    __syncwarp();
    // ---- max_q, max_t (kmer_lookup.c:455-460): state 1 = "max_t holds q_i", 0 = "max_t holds t_i"
    int st = 0, last_q = -1, last_t = -1;
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        const bool v = e < n;
        const uint32_t m = v ? list[e] : 0u;
        const int qi = rm_q(m), ti = rm_t(m);
        int pq = __shfl_up_sync(FULL, qi, 1), pt = __shfl_up_sync(FULL, ti, 1);
        if (lane == 0) { pq = last_q; pt = last_t; }
        const unsigned A = __ballot_sync(FULL, v && pq > ti), B = __ballot_sync(FULL, v && pt > ti);
        const int cnt = min(32, n - e0);
        for (int l = 0; l < cnt; l++) st = (int)(((st ? A : B) >> l) & 1u);
        last_q = __shfl_sync(FULL, qi, cnt - 1); last_t = __shfl_sync(FULL, ti, cnt - 1);
    }
    const int max_q = last_q, max_t = st ? last_q : last_t;
    const int delta = (int)(long)(0.05 * (max_q + max_t));
    // ---- exact-diagonal histogram and its prefix sum F[r] = matches on diagonals < dmin + r
    const int R = dmax - dmin + 1;
    for (int b = lane; b <= R; b += 32) hist[b] = 0;
    __syncwarp();
    for (int e = lane; e < n; e += 32) { const uint32_t m = list[e]; atomicAdd(&hist[rm_q(m) - rm_t(m) - dmin], 1); }
    __syncwarp();
    {
        int carry = 0;
        for (int b0 = 0; b0 <= R; b0 += 32) {
            const int b = b0 + lane;
            const int c = b < R ? hist[b] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
            if (b <= R) hist[b] = carry + incl - c;
            carry += __shfl_sync(FULL, incl, 31);
        }
    }
    __syncwarp();
    // ---- densest window, first maximum in diagonal order (:468-484)
    int bspan = -1, br = INT_MAX;
    for (int b = lane; b < R; b += 32) {
        const int f = hist[b];
        if (hist[b + 1] == f) continue;                                  // no match on this diagonal
        const int e = min(n - 1, b + delta <= R ? hist[b + delta] : n);
        if (e - f > bspan) { bspan = e - f; br = b; }
    }
    const int top = __reduce_max_sync(FULL, bspan);
    br = __reduce_min_sync(FULL, bspan == top ? br : INT_MAX);
    const int max_s = hist[br], max_e = min(n - 1, br + delta <= R ? hist[br + delta] : n);
    if (max_e - max_s < 32) { if (lane == 0) out[p] = r; continue; }
    const int d_lo = dmin + br;
    int d_hi = dmax;                                                     // diagonal of sorted element max_e
    if ((br + delta <= R ? hist[br + delta] : n) <= n - 1) {
        for (int b0 = br + delta; b0 < R; b0 += 32) {                    // first non-empty diagonal >= d_lo + delta
            const int b = b0 + lane;
            const unsigned hit = __ballot_sync(FULL, b < R && hist[b + 1] != hist[b]);
            if (hit) { d_hi = dmin + b0 + __ffs(hit) - 1; break; }
        }
    }
    // ---- in-band matches, list order
    int nb = 0;
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        uint32_t m = 0; bool keep = false;
        if (e < n) { m = list[e]; const int d = rm_q(m) - rm_t(m); keep = d >= d_lo && d <= d_hi; }
        const unsigned kb = __ballot_sync(FULL, keep);
        if (keep) band[nb + __popc(kb & lt)] = m;
        nb += __popc(kb);
    }
    __syncwarp();
    // ---- chaining (:505-547)
    int best_idx = -1, best_score = 0, best_count = 0;
    for (int i = 0; i < nb; i++) {
        const uint32_t mi = band[i];
        const int cx = rm_q(mi), cy = rm_t(mi);
        int cand = -1, max_d = 65535;
        for (int j0 = i - 1; j0 >= 0; j0 -= 32) {
            const int jj = j0 - lane;
            const bool v = jj >= 0;
            const uint32_t mj = v ? band[jj] : 0u;
            const int px = rm_q(mj), py = rm_t(mj);
            const bool stop = v && cx - px > TRIM_GAP;
            const bool ok = v && !stop && cy > py && cy - py <= TRIM_GAP;
            const int key = ok ? (((cx - px + cy - py) << 5) | lane) : INT_MAX;
            const int kmin = __reduce_min_sync(FULL, key);
            if (kmin != INT_MAX && (kmin >> 5) < max_d) { max_d = kmin >> 5; cand = j0 - (kmin & 31); }
            if (__any_sync(FULL, stop)) break;
        }
        int s = 0, c = 0, rt = i;
        if (cand != -1) {
            s = sc[cand] + (64 - max_d); c = cn[cand] + 1; rt = root[cand];
            if (s < 0) { s = 0; c = 0; }
        }
        if (lane == 0) { sc[i] = s; cn[i] = c; root[i] = rt; }
        if (s > best_score) { best_score = s; best_count = c; best_idx = i; }
        __syncwarp();
    }
    if (best_idx >= 0) {
        const uint32_t me = band[best_idx], ms = band[root[best_idx]];
        r.score = best_count + 1; r.e1 = rm_q(me); r.e2 = rm_t(me); r.s1 = rm_q(ms); r.s2 = rm_t(ms);
    }
    if (lane == 0) out[p] = r;
  }
}

// ------------------------------------------------------------------------------ k_subreads
// One thread per output word: new read r = bases [s, s + len) of the packed read at src_woff.
struct SubRead { uint64_t src_woff, dst_woff; int32_t s, len; };

__global__ void k_subreads(const SubRead* __restrict__ subs, uint32_t n_subs, const uint64_t* __restrict__ word_begin,
                           uint64_t total_words, uint32_t* pool) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_words) return;
    uint32_t lo = 0, hi = n_subs;                  // last r with word_begin[r] <= g
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (word_begin[mid] <= g) lo = mid; else hi = mid; }
    const SubRead sr = subs[lo];
    const int64_t j = (int64_t)(g - word_begin[lo]);               // word index inside the new read
    const int64_t first = j * 16;
    uint32_t w = 0;
    if (first < sr.len) {
        w = fetch16(pool + sr.src_woff, sr.s + (int)first);
        const int left = sr.len - (int)first;
        if (left < 16) w &= (1u << (2 * left)) - 1u;
    }
    pool[sr.dst_woff + (uint64_t)j] = w;
}

}  // namespace fcx
