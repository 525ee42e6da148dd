// fcx_kernels.cuh -- device code of the B200-native fc_consensus engine (sm_100a).
//
// Pipeline per wave of seed blocks (all integer work, HBM/latency bound, no tensor cores):
//   k_pack       ASCII -> 2-bit packed reads (A0 C1 G2 T3, 16 bases / 32-bit word, LSB first)
//   k_index      per seed: K=8 k-mer CSR index           (ref: src/c/kmer_lookup.c:140-192)
//   k_range      per pair: k-mer hits + best range        (ref: kmer_lookup.c:207-286, 294-427,
//                                                               falcon.c:612-619)
//   k_dp         per pair: banded O(ND) forward pass      (ref: src/c/DW_banded.c:115-258)
//   k_traceback  per pair: path walk + forward replay     (ref: DW_banded.c:260-320,
//                                                               falcon.c:106-162)
//   k_consensus  per block: column vote, link-DAG longest path, backtrack
//                                                         (ref: falcon.c:308-558)
//
// Every kernel reproduces the reference's integer/double semantics exactly, including the quirks
// listed in SURVEY.md 8(a)-notes; see DESIGN.md for the data layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fcx {

constexpr int KMER = 8;                 // falcon_kit/mains/consensus.py:270
constexpr int KTAB = 1 << (2 * KMER);   // 65536 buckets
constexpr int BIN_SIZE = 48;            // K * INDEL_ALLOWENCE_0 (falcon.c:602-604)
constexpr int COUNT_TH = 5;             // falcon.c:604
constexpr int BAND_TOL = 150;           // INDEL_ALLOWENCE_2 (falcon.c:624)
constexpr int TRACE_REC_WORDS = 8;      // one 32-byte sector per d step: [min_k, w0..w4, pad, pad]
constexpr int MAX_BAND_WORDS = 5;       // band <= 151 cells -> 5 ballot words
constexpr unsigned FULL = 0xffffffffu;

struct BlockDesc {
    uint64_t seed_woff;   // word offset of the seed in the pool
    uint64_t kpos_off;    // offset (entries) into the kpos arena
    uint64_t rec_off;     // offset (records) into the consensus record arena
    uint64_t cns_off;     // offset (bytes) into the consensus output arena
    uint64_t cov_off;     // offset (entries) into the coverage arena
    uint32_t pair_begin;  // first pair of this block (wave-local pair index)
    uint32_t n_pairs;     // n_seq - 1
    int32_t  slen;        // seed length
    uint32_t rec_cap;     // capacity (records)
};

struct PairDesc {
    uint64_t read_woff;   // word offset of the read in the pool
    uint32_t block;       // wave-local block index
    int32_t  rlen;
};

struct PairRange {        // output of k_range
    int32_t s1, e1, s2, e2;
    int32_t n_match;
    int32_t pass;         // span filters passed (falcon.c:613-619)
};

struct PairAlloc {        // host-computed after k_range
    uint64_t trace_off;   // in 32-byte trace records
    uint64_t xam_off;     // in uint32 entries
    uint64_t path_off;    // in uint32 words
};

struct PairAln {          // output of k_dp / k_traceback
    int32_t aligned, dist, aln_size, q_e, t_e, k_end;
    int32_t accepted;
    int32_t t_cnt;        // target positions carrying tags (after the delta>=255 cut)
    int32_t n_tags;
    int32_t cells;
};

// ------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t fetch16(const uint32_t* __restrict__ w, int pos) {
    int wi = pos >> 4;
    uint32_t lo = __ldg(w + wi), hi = __ldg(w + wi + 1);
    return __funnelshift_r(lo, hi, (pos & 15) << 1);
}
__device__ __forceinline__ int base_at(const uint32_t* __restrict__ w, int pos) {
    return (int)((__ldg(w + (pos >> 4)) >> ((pos & 15) << 1)) & 3u);
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// ------------------------------------------------------------------------------ k_pack
// One thread per output word.  Any byte outside "ACGT" raises the dirty flag (reference UB).
__global__ void k_pack(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ aoff,
                       const uint64_t* __restrict__ woff, const int32_t* __restrict__ rlen,
                       uint32_t n_reads, uint64_t total_words, uint32_t* __restrict__ packed,
                       int* __restrict__ dirty) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_words) return;
    uint32_t lo = 0, hi = n_reads;            // last r with woff[r] <= g
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (woff[mid] <= g) lo = mid; else hi = mid; }
    uint32_t r = lo;
    int64_t p0 = (int64_t)(g - woff[r]) * 16;
    int L = rlen[r];
    const uint8_t* src = ascii + aoff[r];
    uint32_t word = 0; int bad = 0;
#pragma unroll
    for (int b = 0; b < 16; b++) {
        int64_t p = p0 + b;
        if (p < L) {
            uint32_t c = src[p];
            uint32_t v = (c >> 1) & 3u;           // A0 C1 T2 G3
            v ^= (v >> 1);                        // A0 C1 G2 T3
            bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
            word |= v << (2 * b);
        }
    }
    packed[g] = word;
    if (bad) atomicOr(dirty, 1);
}

// ------------------------------------------------------------------------------ k_index
// One CTA per seed.  tab[] (65536 uint32, zeroed by the host) becomes, per bucket, the END offset
// into kpos[]; bucket k spans [k ? tab[k-1] : 0, tab[k]), positions ascending -- the order the
// reference's start/next chain yields (kmer_lookup.c:174-191, 257-282).  Positions 0..slen-9 are
// indexed (loop bound `i < seq_len - K`).
__global__ void __launch_bounds__(256) k_index(const BlockDesc* __restrict__ blocks,
                                               const uint32_t* __restrict__ pool,
                                               uint32_t* __restrict__ ktab,
                                               uint32_t* __restrict__ kpos_arena) {
    const BlockDesc bd = blocks[blockIdx.x];
    const uint32_t* seed = pool + bd.seed_woff;
    uint32_t* tab = ktab + (size_t)blockIdx.x * KTAB;
    uint32_t* kpos = kpos_arena + bd.kpos_off;
    const int n = bd.slen - KMER;
    const int tid = threadIdx.x;
    if (n <= 0) return;
    for (int i = tid; i < n; i += 256) atomicAdd(&tab[fetch16(seed, i) & 0xffffu], 1u);
    __syncthreads();
    // exclusive scan, 256 entries per thread
    __shared__ uint32_t part[256];
    uint32_t sum = 0;
    uint32_t* mine = tab + tid * 256;
    for (int j = 0; j < 256; j++) sum += mine[j];
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) { uint32_t run = 0; for (int j = 0; j < 256; j++) { uint32_t v = part[j]; part[j] = run; run += v; } }
    __syncthreads();
    uint32_t run = part[tid];
    for (int j = 0; j < 256; j++) { uint32_t v = mine[j]; mine[j] = run; run += v; }
    __syncthreads();
    // stable fill by warp 0: tab[k] is the cursor of bucket k and ends as its END offset
    if (tid < 32) {
        volatile uint32_t* vtab = tab;
        const unsigned lt = lanemask_lt();
        for (int b0 = 0; b0 < n; b0 += 32) {
            int i = b0 + tid;
            bool valid = i < n;
            uint32_t kid = valid ? (fetch16(seed, i) & 0xffffu) : (0x10000u + tid);
            unsigned peers = __match_any_sync(FULL, kid);
            int rank = __popc(peers & lt);
            uint32_t cur = valid ? vtab[kid] : 0;
            if (valid) kpos[cur + rank] = (uint32_t)i;
            __syncwarp();
            if (valid && rank == 0) vtab[kid] = cur + __popc(peers);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------ k_range
// One warp per pair.  Restates find_kmer_pos_for_seq + find_best_aln_range(K, 48, 5) +
// the span filters of falcon.c:612-619 without materialising the match list: the list is
// re-walked (order: query position ascending, seed position ascending) for each pass.
//   pass A  count, d_min, d_max                       (kmer_lookup.c:323-343)
//   pass B  48-wide diagonal histogram in shared mem  (:346-355)
//   pass C  arg-max bin, first in match order         (:357-366)
//   pass D  kept matches -> Kadane scan in closed form (:369-411):
//           with S_i = 32*i - q_i (i = index in the kept list) the running score is
//           S_i - min_{j<=i} S_j, a reset happens exactly on a strict new prefix minimum, so
//           (s1,s2) = coordinates of the FIRST arg-min of the prefix and (e1,e2) those of the
//           first i attaining the overall maximum.
constexpr int RANGE_WARPS = 4;
constexpr int RANGE_BINS = 4224;   // >= (99999 + 99999) / 48 + 1

__global__ void __launch_bounds__(RANGE_WARPS * 32)
k_range(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs, uint32_t n_pairs,
        const uint32_t* __restrict__ pool, const uint32_t* __restrict__ ktab,
        const uint32_t* __restrict__ kpos_arena, PairRange* __restrict__ out) {
    extern __shared__ int s_hist_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * RANGE_WARPS + wib;
    if (p >= n_pairs) return;
    int* hist = s_hist_all + wib * RANGE_BINS;
    const PairDesc pd = pairs[p];
    const BlockDesc bd = blocks[pd.block];
    const uint32_t* read = pool + pd.read_woff;
    const uint32_t* tab = ktab + (size_t)pd.block * KTAB;
    const uint32_t* kpos = kpos_arena + bd.kpos_off;
    const int nq = pd.rlen > KMER ? (pd.rlen - KMER + 3) / 4 : 0;   // i = 0,4,.. < rlen-K
    PairRange r; r.s1 = r.e1 = r.s2 = r.e2 = 0; r.n_match = 0; r.pass = 0;

    // ---- pass A
    int cnt = 0, dmin = INT_MAX, dmax = INT_MIN;
    for (int it = lane; it < nq; it += 32) {
        int i = it * 4;
        uint32_t kid = fetch16(read, i) & 0xffffu;
        uint32_t s = kid ? __ldg(tab + kid - 1) : 0u, e = __ldg(tab + kid);
        for (uint32_t j = s; j < e; j++) {
            int d = i - (int)__ldg(kpos + j);
            dmin = min(dmin, d); dmax = max(dmax, d); cnt++;
        }
    }
    cnt = __reduce_add_sync(FULL, cnt);
    dmin = __reduce_min_sync(FULL, dmin);
    dmax = __reduce_max_sync(FULL, dmax);
    r.n_match = cnt;
    if (cnt == 0) { if (lane == 0) out[p] = r; return; }
    const int nbin = (dmax - dmin) / BIN_SIZE + 1;

    // ---- pass B
    for (int b = lane; b < nbin; b += 32) hist[b] = 0;
    __syncwarp();
    for (int it = lane; it < nq; it += 32) {
        int i = it * 4;
        uint32_t kid = fetch16(read, i) & 0xffffu;
        uint32_t s = kid ? __ldg(tab + kid - 1) : 0u, e = __ldg(tab + kid);
        for (uint32_t j = s; j < e; j++) atomicAdd(&hist[(i - (int)__ldg(kpos + j) - dmin) / BIN_SIZE], 1);
    }
    __syncwarp();

    // ---- pass C
    int top = 0;
    for (int b = lane; b < nbin; b += 32) top = max(top, hist[b]);
    top = __reduce_max_sync(FULL, top);
    if (top <= COUNT_TH) { if (lane == 0) out[p] = r; return; }
    int top_bin = -1;
    for (int it0 = 0; it0 < nq && top_bin < 0; it0 += 32) {
        int it = it0 + lane, mybin = -1;
        if (it < nq) {
            int i = it * 4;
            uint32_t kid = fetch16(read, i) & 0xffffu;
            uint32_t s = kid ? __ldg(tab + kid - 1) : 0u, e = __ldg(tab + kid);
            for (uint32_t j = s; j < e; j++) {
                int b = (i - (int)__ldg(kpos + j) - dmin) / BIN_SIZE;
                if (hist[b] == top) { mybin = b; break; }
            }
        }
        unsigned bal = __ballot_sync(FULL, mybin >= 0);
        if (bal) top_bin = __shfl_sync(FULL, mybin, __ffs(bal) - 1);
    }

    // ---- pass D
    long long idx_base = 0;                  // kept elements before this 32-query window
    bool have_min = false; long long c_min = 0; int c_mq = 0, c_mt = 0;   // carry: running prefix min
    long long best = 0; int bs1 = 0, bs2 = 0, be1 = 0, be2 = 0;
    bool first_set = false; int q0 = 0, t0 = 0;
    for (int it0 = 0; it0 < nq; it0 += 32) {
        int it = it0 + lane, i = it * 4;
        int k = 0, tf = 0, tl = 0;
        if (it < nq) {
            uint32_t kid = fetch16(read, i) & 0xffffu;
            uint32_t s = kid ? __ldg(tab + kid - 1) : 0u, e = __ldg(tab + kid);
            for (uint32_t j = s; j < e; j++) {
                int t = (int)__ldg(kpos + j);
                int b = (i - t - dmin) / BIN_SIZE;
                if (abs(b - top_bin) > 5) continue;
                if (hist[b] > COUNT_TH) { if (k == 0) tf = t; tl = t; k++; }
            }
        }
        unsigned has = __ballot_sync(FULL, k > 0);
        if (!has) continue;
        // exclusive prefix of k over lanes
        int incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        int excl = incl - k;
        int total = __shfl_sync(FULL, incl, 31);
        if (!first_set) { int fl = __ffs(has) - 1; q0 = __shfl_sync(FULL, i, fl); t0 = __shfl_sync(FULL, tf, fl); first_set = true; }
        // S of the lane's first kept element; prefix-min scan keeping the earliest arg-min
        long long sv = (k > 0) ? 32ll * (idx_base + excl) - i : LLONG_MAX;
        int src = lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long ov = __shfl_up_sync(FULL, sv, o); int os = __shfl_up_sync(FULL, src, o);
            if (lane >= o && ov <= sv) { sv = ov; src = os; }     // earlier lane wins ties
        }
        // fold the carry (earlier than every lane of this window: wins ties)
        bool from_carry = have_min && c_min <= sv;
        long long mval = from_carry ? c_min : sv;
        int sq = __shfl_sync(FULL, i, src), st = __shfl_sync(FULL, tf, src);
        if (from_carry) { sq = c_mq; st = c_mt; }
        // candidate: the lane's last kept element
        long long cval = (k > 0) ? 32ll * (idx_base + excl + k - 1) - i - mval : -1;
        long long cmax = cval;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { long long v = __shfl_xor_sync(FULL, cmax, o); cmax = v > cmax ? v : cmax; }
        if (cmax > best) {
            unsigned who = __ballot_sync(FULL, cval == cmax);
            int wl = __ffs(who) - 1;
            best = cmax;
            be1 = __shfl_sync(FULL, i, wl); be2 = __shfl_sync(FULL, tl, wl);
            bs1 = __shfl_sync(FULL, sq, wl); bs2 = __shfl_sync(FULL, st, wl);
        }
        // new carry = inclusive state of lane 31
        long long lv = __shfl_sync(FULL, mval, 31); int lq = __shfl_sync(FULL, sq, 31), ltt = __shfl_sync(FULL, st, 31);
        if (lv != LLONG_MAX) { have_min = true; c_min = lv; c_mq = lq; c_mt = ltt; }
        idx_base += total;
    }
    if (idx_base > 1) {
        if (best > 0) { r.s1 = bs1; r.s2 = bs2; r.e1 = be1; r.e2 = be2; }
        else { r.s1 = r.e1 = q0; r.s2 = r.e2 = t0; }
    }
    // span filters, falcon.c:612-619 (double arithmetic kept as written there)
    int sp1 = r.e1 - r.s1, sp2 = r.e2 - r.s2;
    r.pass = !(sp1 < 100 || sp2 < 100 || abs(sp1 - sp2) > (int)(0.5 * 0.10 * (sp1 + sp2)));
    if (lane == 0) out[p] = r;
}

// ------------------------------------------------------------------------------ k_dp
// One warp per pair: furthest-reaching banded O(ND) forward pass (DW_banded.c:149-258).
// Lanes own diagonals k = min_k + 2*lane (+64 per extra band chunk).  V lives in a 512-entry
// shared-memory ring indexed by k (both parities interleaved; the in-place update of
// DW_banded.c:213 only ever reads the opposite parity, written at d-1).  Per step one 32-byte
// trace record is written: [min_k, ballot words of "came from k+1"].  x2/y2 are NOT stored; the
// traceback kernel re-walks the path and recomputes the snakes.
constexpr int DP_WARPS = 8;
constexpr int VRING = 512;

__device__ __forceinline__ int snake(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t,
                                     int qs, int ts, int q_len, int t_len, int& x, int& y) {
    int adv = 0;
    for (;;) {
        int rem = min(q_len - x, t_len - y);
        if (rem <= 0) break;
        uint32_t diff = fetch16(q, qs + x) ^ fetch16(t, ts + y);
        int n = diff ? (__ffs(diff) - 1) >> 1 : 16;
        n = min(n, rem);
        x += n; y += n; adv += n;
        if (n < 16) break;
    }
    return adv;
}

__global__ void __launch_bounds__(DP_WARPS * 32)
k_dp(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs,
     const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs, uint32_t n_pairs,
     const uint32_t* __restrict__ pool, uint32_t* __restrict__ trace_arena, double max_diff,
     PairAln* __restrict__ out) {
    __shared__ int s_V[DP_WARPS][VRING];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * DP_WARPS + wib;
    if (p >= n_pairs) return;
    PairAln res; res.aligned = res.dist = res.aln_size = res.q_e = res.t_e = res.k_end = 0;
    res.accepted = res.t_cnt = res.n_tags = res.cells = 0;
    const PairRange rg = ranges[p];
    if (!rg.pass) { if (lane == 0) out[p] = res; return; }
    const PairDesc pd = pairs[p];
    const uint32_t* q = pool + pd.read_woff;
    const uint32_t* t = pool + blocks[pd.block].seed_woff;
    const int qs = rg.s1, ts = rg.s2, q_len = rg.e1 - rg.s1, t_len = rg.e2 - rg.s2;
    const int max_d = (int)(0.3 * (q_len + t_len));           // DW_banded.c:149
    const int band_size = BAND_TOL * 2;                        // :151
    int* V = s_V[wib];
    uint32_t* trace = trace_arena + allocs[p].trace_off * TRACE_REC_WORDS;

    int best_m = -1, min_k = 0, max_k = 0, cells = 0;
    bool aligned = false; int end_d = 0, end_k = 0, end_x = 0, end_y = 0;
    for (int d = 0; d < max_d; d++) {
        if (max_k - min_k > band_size) break;                  // :184-186
        const int ncell = ((max_k - min_k) >> 1) + 1;
        const int nch = (ncell + 31) >> 5;
        uint32_t* rec = trace + (size_t)d * TRACE_REC_WORDS;
        if (lane == 0) rec[0] = (uint32_t)min_k;
        int step_best = best_m;
        for (int c = 0; c < nch; c++) {
            const int k = min_k + 2 * (lane + 32 * c);
            const bool act = k <= max_k;
            bool up = false; int x = 0, y = 0;
            if (act) {
                if (d == 0) { up = true; x = 0; }              // V[k+1] is calloc'd 0
                else {
                    int vm = V[(k - 1) & (VRING - 1)], vp = V[(k + 1) & (VRING - 1)];
                    up = (k == min_k) || (k != max_k && vm < vp);   // :190
                    x = up ? vp : vm + 1;
                }
                y = x - k;
                snake(q, t, qs, ts, q_len, t_len, x, y);
            }
            unsigned upb = __ballot_sync(FULL, act && up);
            if (lane == 0) rec[1 + c] = upb;
            const bool fin = act && (x >= q_len || y >= t_len);     // :220
            unsigned finb = __ballot_sync(FULL, fin);
            if (act) V[k & (VRING - 1)] = x;
            int u = act ? x + y : INT_MIN;
            step_best = max(step_best, __reduce_max_sync(FULL, u));
            if (finb) {                                         // first k in ascending order wins
                int fl = __ffs(finb) - 1;
                aligned = true; end_d = d; end_k = min_k + 2 * (fl + 32 * c);
                end_x = __shfl_sync(FULL, x, fl); end_y = __shfl_sync(FULL, y, fl);
                cells += fl + 1;
                break;
            }
            cells += __popc(__ballot_sync(FULL, act));
        }
        if (aligned) break;
        best_m = step_best;
        __syncwarp();
        // band update, :227-243
        int nmin = INT_MAX, nmax = INT_MIN;
        const int thr = best_m - BAND_TOL;
        for (int c = 0; c < nch; c++) {
            const int k = min_k + 2 * (lane + 32 * c);
            bool ok = false;
            if (k <= max_k) { int x = V[k & (VRING - 1)]; ok = (2 * x - k) >= thr; }
            unsigned okb = __ballot_sync(FULL, ok);
            if (okb) {
                if (nmin == INT_MAX) nmin = min_k + 2 * (__ffs(okb) - 1 + 32 * c);
                nmax = min_k + 2 * (31 - __clz(okb) + 32 * c);
            }
        }
        max_k = nmax + 1; min_k = nmin - 1;
        __syncwarp();
    }
    if (aligned) {
        res.aligned = 1; res.dist = end_d; res.q_e = end_x; res.t_e = end_y; res.k_end = end_k;
        res.aln_size = (end_x + end_y + end_d) / 2;            // :256, equals the traced length
        res.accepted = (res.aln_size > 500 &&
                        ((double)res.dist / (double)res.aln_size) < max_diff) ? 1 : 0;   // falcon.c:629
    }
    res.cells = cells;
    if (lane == 0) out[p] = res;
}

// ------------------------------------------------------------------------------ k_traceback
// One thread per accepted pair.  (1) walk the trace records backwards collecting the direction
// bit of every step of the optimal path (DW_banded.c:264-277); (2) replay the path forwards,
// recomputing the snakes, and emit per target position y the record
//        xam[y] = (x << 1) | is_match      x = query index when target base y is consumed,
// plus the sentinel xam[t_e] = q_e << 1.  This is get_align_tags (falcon.c:106-162) in closed
// form: the delta-0 tag of column y carries the seed base (match) or '-', and the insertion tags
// delta = 1..n are the query bases x+is_match .. x_next-1.  The reference stops tagging at the
// first column whose delta reaches 255 (falcon.c:138,150-152): t_cnt is cut there and the
// sentinel rewritten so that exactly 254 insertion tags remain.
__global__ void k_traceback(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs,
                            const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs,
                            uint32_t n_pairs, const uint32_t* __restrict__ pool,
                            const uint32_t* __restrict__ trace_arena, uint32_t* __restrict__ path_arena,
                            uint32_t* __restrict__ xam_arena, PairAln* __restrict__ aln) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    PairAln a = aln[p];
    if (!a.accepted) return;
    const PairRange rg = ranges[p];
    const PairDesc pd = pairs[p];
    const uint32_t* q = pool + pd.read_woff;
    const uint32_t* t = pool + blocks[pd.block].seed_woff;
    const int qs = rg.s1, ts = rg.s2, q_len = rg.e1 - rg.s1, t_len = rg.e2 - rg.s2;
    const PairAlloc al = allocs[p];
    const uint32_t* trace = trace_arena + al.trace_off * TRACE_REC_WORDS;
    uint32_t* path = path_arena + al.path_off;
    uint32_t* xam = xam_arena + al.xam_off;
    const int D = a.dist;
    // (1) backwards: bit d of path = step d came from k+1 (a target-only column)
    int k = a.k_end;
    uint32_t acc = 0;
    for (int d = D; d >= 1; d--) {
        const uint32_t* rec = trace + (size_t)d * TRACE_REC_WORDS;
        int idx = (k - (int)rec[0]) >> 1;
        uint32_t up = (rec[1 + (idx >> 5)] >> (idx & 31)) & 1u;
        acc |= up << (d & 31);
        if ((d & 31) == 0 || d == 1) { path[d >> 5] = acc; acc = 0; }
        k += up ? 1 : -1;
    }
    // (2) forwards
    int x = 0, y = 0, run = 0, t_cnt = -1, n_match_cols = 0;
    uint32_t pw = 0;
    for (int d = 0; d <= D; d++) {
        if (d > 0) {
            if ((d & 31) == 0 || d == 1) pw = path[d >> 5];
            if ((pw >> (d & 31)) & 1u) { xam[y] = (uint32_t)x << 1; y++; run = 0; }   // target-only column
            else {
                x++; run++;
                if (run == 255) {            // the 255th consecutive query-only column: tags stop before it
                    t_cnt = y;               // target positions 0..y-1 carry tags
                    xam[y] = (uint32_t)(x - 1) << 1;   // keeps n_ins(y-1) == 254
                    break;
                }
            }
        }
        // snake
        for (;;) {
            int rem = min(q_len - x, t_len - y);
            if (rem <= 0) break;
            uint32_t diff = fetch16(q, qs + x) ^ fetch16(t, ts + y);
            int n = diff ? (__ffs(diff) - 1) >> 1 : 16;
            n = min(n, rem);
            for (int j = 0; j < n; j++) xam[y + j] = ((uint32_t)(x + j) << 1) | 1u;
            x += n; y += n; n_match_cols += n;
            if (n > 0) run = 0;
            if (n < 16) break;
        }
    }
    if (t_cnt < 0) { t_cnt = y; xam[y] = (uint32_t)x << 1; }
    a.t_cnt = t_cnt;
    // tagged columns = target columns + query-only columns among them
    a.n_tags = t_cnt + ((int)(xam[t_cnt] >> 1) - n_match_cols);
    aln[p] = a;
}

// ------------------------------------------------------------------------------ k_consensus
// One warp per seed block, serial over target positions (the longest-path DP of
// falcon.c:405-475 is serial in t_pos).  Per position i:
//   vote   every accepted read covering i contributes its delta-0 tag and its insertion tags
//          (falcon.c:350-382); identical links are merged into a per-position link list in
//          first-appearance order (= ascending accepted-read index, update_col falcon.c:232-263);
//   DP     for delta j = 0..max_delta, base kk = 0..4: best link by strict '>' in list order,
//          score = pred + count - 0.5*coverage kept as an exact integer (x2); columns whose best
//          stays <= -1 keep score -1 and best_p = (0,0,0) as in the reference;
//   global best by strict '>' in (i, j, kk) order, remembering the best LINK INDEX (quirk).
// Then the backtrack of falcon.c:479-542 over the stored column records.
struct CnsRec { int32_t pred; int32_t info; int32_t score2; };   // info = (t_pos << 3) | base
constexpr int CNS_WARPS = 4;
constexpr int LINK_CAP = 512;         // distinct (delta, base, link) entries per position
constexpr int LVL = 255 * 5;

struct CnsOut { int32_t len; int32_t err; };

__global__ void __launch_bounds__(CNS_WARPS * 32)
k_consensus(const BlockDesc* __restrict__ blocks, uint32_t n_blocks, const PairDesc* __restrict__ pairs,
            const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs,
            const PairAln* __restrict__ aln, const uint32_t* __restrict__ pool,
            const uint32_t* __restrict__ xam_arena, CnsRec* __restrict__ rec_arena,
            uint16_t* __restrict__ cov_arena, int32_t* __restrict__ lvl_scratch,
            uint32_t* __restrict__ acc_scratch, uint64_t acc_stride,
            char* __restrict__ cns_arena, int32_t* __restrict__ eqv_arena, unsigned min_cov,
            CnsOut* __restrict__ out) {
    // per-warp link list of the current position
    __shared__ uint32_t s_key[CNS_WARPS][LINK_CAP];    // (delta << 16) | (base << 13) | link
    __shared__ uint16_t s_cnt[CNS_WARPS][LINK_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * CNS_WARPS + wib;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    const uint32_t* seed = pool + bd.seed_woff;
    const int t_len = bd.slen;
    uint32_t* key = s_key[wib]; uint16_t* kcnt = s_cnt[wib];
    CnsRec* recs = rec_arena + bd.rec_off;
    uint16_t* cov = cov_arena + bd.cov_off;
    // per-warp score/record tables of the previous and current position, indexed delta*5+base
    const size_t gw = (size_t)blockIdx.x * CNS_WARPS + wib;
    int32_t* lv_sc[2]  = { lvl_scratch + gw * 4 * LVL, lvl_scratch + gw * 4 * LVL + LVL };
    int32_t* lv_rec[2] = { lvl_scratch + gw * 4 * LVL + 2 * LVL, lvl_scratch + gw * 4 * LVL + 3 * LVL };
    // accepted reads of this block, in order
    uint32_t* acc = acc_scratch + gw * acc_stride;
    int R = 0;
    for (uint32_t j0 = 0; j0 < bd.n_pairs; j0 += 32) {
        uint32_t j = j0 + lane;
        bool ok = j < bd.n_pairs && aln[bd.pair_begin + j].accepted;
        unsigned bal = __ballot_sync(FULL, ok);
        if (ok) acc[R + __popc(bal & lanemask_lt())] = bd.pair_begin + j;
        R += __popc(bal);
    }
    __syncwarp();
    CnsOut co; co.len = 0; co.err = 0;
    char* cns = cns_arena + bd.cns_off;
    int32_t* eqv = eqv_arena + bd.cns_off;
    if (R == 0) { if (lane == 0) { cns[0] = 0; out[b] = co; } return; }     // falcon.c:651-656

    // record 0 is reserved for column (0,0,'A'): the target of floored columns' best_p = (0,0,0)
    if (lane == 0) { recs[0].pred = 0; recs[0].info = 0; recs[0].score2 = -2; }
    uint32_t nrec = 1;
    int g_best2 = -2, g_rec = -1, g_ck = 0, g_t = 0;
    int cur = 0;
    int err = 0;

    for (int i = 0; i < t_len; i++) {
        const int Si = base_at(seed, i);
        const int Sp = i > 0 ? base_at(seed, i - 1) : 0;
        int nlink = 0, coverage = 0, maxd = 0;
        // ------------------------------------------------------------ vote
        for (int c0 = 0; c0 < R; c0 += 32) {
            const int ai = c0 + lane;
            bool act = false; int y = 0; uint32_t pidx = 0; int ts = 0;
            if (ai < R) {
                pidx = acc[ai];
                ts = ranges[pidx].s2;
                y = i - ts;
                act = y >= 0 && y < aln[pidx].t_cnt;
            }
            unsigned actb = __ballot_sync(FULL, act);
            if (!actb) continue;
            coverage += __popc(actb);
            // lane state
            int m = 0, x = 0, nins = 0, b0 = 0; uint32_t lk0 = 0;
            const uint32_t* qr = pool; int qs = 0;
            if (act) {
                const uint32_t* xam = xam_arena + allocs[pidx].xam_off;
                uint32_t c = xam[y], nx = xam[y + 1];
                m = c & 1; x = (int)(c >> 1); nins = (int)(nx >> 1) - x - m;
                b0 = m ? Si : 4;
                qr = pool + pairs[pidx].read_woff; qs = ranges[pidx].s1;
                if (y == 0) lk0 = 0x1fffu;                         // (p_t_pos = -1, 0, '.')
                else {
                    uint32_t pv = xam[y - 1];
                    int pm = pv & 1, px = (int)(pv >> 1), pn = x - px - pm;
                    int pb = pn > 0 ? base_at(qr, qs + x - 1) : (pm ? Sp : 4);
                    lk0 = ((uint32_t)pn << 3) | (uint32_t)pb;
                }
            }
            int lmax = __reduce_max_sync(FULL, act ? nins : 0);
            maxd = max(maxd, lmax);
            // levels 0..lmax of this chunk, merged in (level, lane) order.  Merging level by
            // level is equivalent to read-by-read order: first appearance of a link inside one
            // column is decided by the read index alone.
            for (int lev = 0; lev <= lmax; lev++) {
                bool has = act && nins >= lev;
                uint32_t k = 0xffffffffu - lane;
                if (has) {
                    if (lev == 0) k = ((uint32_t)b0 << 13) | lk0;
                    else {
                        int bb = base_at(qr, qs + x + m + lev - 1);
                        int pb = (lev == 1) ? b0 : base_at(qr, qs + x + m + lev - 2);
                        k = ((uint32_t)lev << 16) | ((uint32_t)bb << 13) | ((uint32_t)(lev - 1) << 3) | (uint32_t)pb;
                    }
                }
                unsigned peers = __match_any_sync(FULL, k);
                bool leader = has && (__popc(peers & lanemask_lt()) == 0);
                int pc = __popc(peers);
                unsigned lead = __ballot_sync(FULL, leader);
                while (lead) {
                    int l = __ffs(lead) - 1; lead &= lead - 1;
                    uint32_t kk = __shfl_sync(FULL, k, l); int cc = __shfl_sync(FULL, pc, l);
                    int found = -1;
                    for (int e0 = 0; e0 < nlink; e0 += 32) {
                        unsigned hb = __ballot_sync(FULL, (e0 + lane < nlink) && key[e0 + lane] == kk);
                        if (hb) { found = e0 + __ffs(hb) - 1; break; }
                    }
                    if (found >= 0) { if (lane == 0) kcnt[found] = (uint16_t)(kcnt[found] + cc); }
                    else if (nlink < LINK_CAP) { if (lane == 0) { key[nlink] = kk; kcnt[nlink] = (uint16_t)cc; } nlink++; }
                    else err = 1;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) cov[i] = (uint16_t)min(coverage, 65535);
        if (coverage == 0) { cur ^= 1; continue; }    // no tags here: max_delta = 0, all columns dead
        // ------------------------------------------------------------ DP over (j, kk)
        int32_t* sc_p = lv_sc[cur ^ 1]; int32_t* rc_p = lv_rec[cur ^ 1];
        int32_t* sc_c = lv_sc[cur];     int32_t* rc_c = lv_rec[cur];
        for (int j = 0; j <= maxd; j++) {
            for (int kk = 0; kk < 5; kk++) {
                // scan the link list for column (j, kk); list order == first-appearance order
                const uint32_t want = ((uint32_t)j << 3) | (uint32_t)kk;       // key >> 13
                bool any = false;
                int seen = 0;          // links of this column seen in earlier list windows
                long long bkey = LLONG_MIN; int brec = -1, bsc2 = 0;
                for (int e0 = 0; e0 < nlink; e0 += 32) {
                    int e = e0 + lane;
                    bool mine = e < nlink && (key[e] >> 13) == want;
                    unsigned mb = __ballot_sync(FULL, mine);
                    if (!mb) continue;
                    any = true;
                    long long v = LLONG_MIN; int prc = -1; int s2v = 0;
                    if (mine) {
                        uint32_t lk = key[e] & 0x1fffu;
                        int cnt = kcnt[e];
                        int s2 = 2 * cnt - coverage;                      // 2*(count - 0.5*cov)
                        if (lk != 0x1fffu) {
                            int pdl = lk >> 3, pbb = lk & 7;
                            const int32_t* sc_src = (j == 0) ? sc_p : sc_c;
                            const int32_t* rc_src = (j == 0) ? rc_p : rc_c;
                            s2 += sc_src[pdl * 5 + pbb];
                            prc = rc_src[pdl * 5 + pbb];
                        }
                        s2v = s2;
                        // order: higher score first, then earlier list position
                        v = (long long)s2 * (1ll << 20) + (long long)(0xfffff - (seen + __popc(mb & lanemask_lt())));
                    }
                    long long wm = v;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { long long ov = __shfl_xor_sync(FULL, wm, o); wm = ov > wm ? ov : wm; }
                    if (wm > bkey) {
                        unsigned who = __ballot_sync(FULL, mine && v == wm);
                        int wl = __ffs(who) - 1;
                        bkey = wm; brec = __shfl_sync(FULL, prc, wl); bsc2 = __shfl_sync(FULL, s2v, wl);
                    }
                    seen += __popc(mb);
                }
                if (!any) { if (lane == 0) { sc_c[j * 5 + kk] = -2; rc_c[j * 5 + kk] = 0; } continue; }  // dead column: score -1
                int best_ck = 0xfffff - (int)(bkey & 0xfffff);
                int col_sc2, col_pred;
                if (bsc2 > -2) { col_sc2 = bsc2; col_pred = brec; }      // recorded only if score > -1
                else { col_sc2 = -2; col_pred = 0; best_ck = -1; }        // floored: best_p = (0,0,0)
                // column record
                uint32_t ridx;
                if (i == 0 && j == 0 && kk == 0) ridx = 0; else { ridx = nrec; nrec++; }
                if (ridx >= bd.rec_cap) { err = 2; ridx = bd.rec_cap - 1; }
                if (lane == 0) {
                    recs[ridx].pred = col_pred; recs[ridx].info = (i << 3) | kk; recs[ridx].score2 = col_sc2;
                    sc_c[j * 5 + kk] = col_sc2; rc_c[j * 5 + kk] = (int32_t)ridx;
                }
                if (col_sc2 > g_best2) { g_best2 = col_sc2; g_rec = (int)ridx; g_ck = best_ck; g_t = i; }
            }
            __syncwarp();
        }
        cur ^= 1;
        __syncwarp();
    }
    // ------------------------------------------------------------ backtrack (falcon.c:479-542)
    if (g_rec < 0) err = 3;                       // reference: assert(g_best_score != -1)
    __syncwarp();
    int index = 0;
    if (lane == 0 && err == 0) {
        char bb = '$'; int ck = g_ck; int i = g_t; int rc = g_rec;
        const unsigned lim = (unsigned)t_len * 2u;
        for (;;) {
            const bool hi = (unsigned)cov[i] > min_cov;
            switch (ck) {
                case 0: bb = hi ? 'A' : 'a'; break;
                case 1: bb = hi ? 'C' : 'c'; break;
                case 2: bb = hi ? 'G' : 'g'; break;
                case 3: bb = hi ? 'T' : 't'; break;
                case 4: bb = '-'; break;
                default: break;
            }
            const CnsRec r = recs[rc];
            if (r.pred == -1 || (unsigned)index >= lim) break;
            const CnsRec pr = recs[r.pred];
            i = pr.info >> 3; ck = pr.info & 7;
            if (bb != '-') { cns[index] = bb; eqv[index] = r.score2 / 2 - pr.score2 / 2; index++; }
            rc = r.pred;
        }
    }
    index = __shfl_sync(FULL, index, 0);
    __syncwarp();
    // reverse in place (falcon.c:533-540)
    for (int a = lane; a < index / 2; a += 32) {
        char tc = cns[a]; cns[a] = cns[index - 1 - a]; cns[index - 1 - a] = tc;
        int te = eqv[a]; eqv[a] = eqv[index - 1 - a]; eqv[index - 1 - a] = te;
    }
    if (lane == 0) { cns[index] = 0; co.len = index; co.err = err; out[b] = co; }
}

}  // namespace fcx
