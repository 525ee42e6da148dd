// fcx_kernels.cuh -- device code of the B200-native fc_consensus engine (sm_100a).
//
// Pipeline per wave of seed blocks (all integer work, issue/latency bound, no tensor cores):
//   k_pack / k_repack_bps / k_subreads   the 2-bit packed read pool (A0 C1 G2 T3, 16 bases / 32-bit word,
//                LSB first) from ASCII, from a Dazzler .bps image, or cut out of packed reads (--trim)
//   k_index      per seed: K=8 k-mer CSR index           (ref: src/c/kmer_lookup.c:140-192)
//   k_range      per pair: k-mer hits + best range        (ref: kmer_lookup.c:207-286, 294-427,
//                                                               falcon.c:612-619)
//   k_dp3        per pair: banded O(ND) forward pass + backward walk of the trace   (fcx_dp.cuh;
//                                                          ref: src/c/DW_banded.c:115-277)
//   k_traceback  per pair: forward replay of the path -> per-column entries
//                                                         (ref: DW_banded.c:284-320, falcon.c:106-162)
//   k_vote       per seed position: column vote          (fcx_vote.cuh; ref: falcon.c:350-382)
//   k_cns_dp     per block: link-DAG longest path, backtrack   (fcx_vote.cuh; ref: falcon.c:405-542)
//   k_trim_range per pair: masked k-mer hits + find_best_aln_range2   (fcx_trim.cuh; --trim)
//   k_dp<>, k_traceback_walk   the round-1 DP kernels (A/B: dp_variant 1 / 2, the latter TMA-staged)
//   k_align1 / k_align1_tb / k_align_batch   align() with an arbitrary band (legacy symbol, stage 2)
//
// Every kernel reproduces the reference's integer/double semantics exactly, including the quirks
// listed in SURVEY.md 8(a)-notes; see DESIGN.md for the data layout.
#pragma once
#include <climits>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

// Launch and dynamic-shared-memory spellings.  FCX_EMU is defined only by tests/emu (a SIMT
// emulator used to check kernel logic on machines without a GPU); the product is built by nvcc.
#ifdef FCX_EMU
#define FCX_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(emu::Dim3((unsigned)(grid)), emu::Dim3((unsigned)(block)), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define FCX_DYN_SHARED(type, name) type* name = reinterpret_cast<type*>(emu::g_cta->dyn_smem)
#define FCX_NOINLINE __attribute__((noinline))
#else
#define FCX_NOINLINE __noinline__
#define FCX_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FCX_DYN_SHARED(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type* name = reinterpret_cast<type*>(name##_raw_)
#endif

namespace fcx {

constexpr int KMER = 8;                 // falcon_kit/mains/consensus.py:270
constexpr int KTAB = 1 << (2 * KMER);   // 65536 buckets
constexpr int BIN_SIZE = 48;            // K * INDEL_ALLOWENCE_0 (falcon.c:602-604)
constexpr int COUNT_TH = 5;             // falcon.c:604
constexpr int BAND_TOL = 150;           // INDEL_ALLOWENCE_2 (falcon.c:624)
constexpr int TRACE_REC_WORDS = 8;      // one 32-byte sector per d step: [min_k, w0..w4, pad, pad]
constexpr unsigned FULL = 0xffffffffu;

struct BlockDesc {
    uint64_t seed_woff;   // word offset of the seed in the pool
    uint64_t kpos_off;    // offset (entries) into the kpos arena
    uint64_t rec_off;     // offset (records) into the consensus record arena
    uint64_t cns_off;     // offset (bytes) into the consensus output arena
    uint32_t pair_begin;  // first pair of this block (wave-local pair index)
    uint32_t n_pairs;     // n_seq - 1
    int32_t  slen;        // seed length
    uint32_t rec_cap;     // capacity (records)
    uint64_t slot_off;    // offset (positions) of this block's vote slots
    uint32_t pad_;
    uint32_t tile_begin;  // first k_vote tile (VOTE_TP positions) of this block
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
    uint64_t xam_off;     // in uint32 entries (a multiple of 4): the pair's ent[] array
    uint64_t path_off;    // in uint32 words
    uint32_t trace_cap;   // trace records available: steps d >= trace_cap are not recorded (such a
                          // pair can no longer be accepted, see fcx_engine.cu)
    uint32_t xck_off;     // in uint32 entries: query-index checkpoints, one per 32 target columns
};

struct PairAln {          // output of k_dp / k_traceback
    int32_t aligned, dist, aln_size, q_e, t_e, k_end;
    int32_t accepted;
    int32_t t_cnt;        // target positions carrying tags (after the delta>=255 cut)
    int32_t n_tags;
    int32_t cells;
};

// per pair, written by k_traceback for accepted pairs (all-zero = not accepted): what k_vote needs
struct VoteMeta { uint64_t ent_off; uint64_t q_woff; int32_t t_start, t_cnt, q_s; uint32_t xck_off; };

// ------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t fetch16(const uint32_t* __restrict__ w, int pos) {
    int wi = pos >> 4;
    uint32_t lo = __ldg(w + wi), hi = __ldg(w + wi + 1);
    return __funnelshift_r(lo, hi, (pos & 15) << 1);
}
// same, with the packed words staged in shared memory (k_dp<true>)
template <bool SM>
__device__ __forceinline__ uint32_t fetch16_t(const uint32_t* __restrict__ w, int pos) {
    if (SM) { const int wi = pos >> 4; return __funnelshift_r(w[wi], w[wi + 1], (pos & 15) << 1); }
    return fetch16(w, pos);
}
__device__ __forceinline__ int base_at(const uint32_t* __restrict__ w, int pos) {
    return (int)((__ldg(w + (pos >> 4)) >> ((pos & 15) << 1)) & 3u);
}
__device__ __forceinline__ unsigned lanemask_lt() {
#ifdef FCX_EMU
    return (1u << (threadIdx.x & 31)) - 1u;
#else
    unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
#endif
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

// ------------------------------------------------------------------------------ k_repack_bps
// Dazzler .bps bytes (4 bases per byte, first base in the top two bits, a0 c1 g2 t3: DAZZ_DB
// DB.c:Compress_Read) -> the engine's packed reads.  Pool entry e = 2r + c: read r forward (c = 0) or
// reverse-complemented (c = 1), cut to `len` bases the way consensus.py:178-179 cuts the text
// (the FIRST len characters of the oriented sequence).  One thread per output word.
__global__ void k_repack_bps(const uint8_t* __restrict__ bps, const uint64_t* __restrict__ boff,
                             const int32_t* __restrict__ rlen, const int32_t* __restrict__ elen,
                             const uint64_t* __restrict__ woff, uint32_t n_entries, uint64_t total_words,
                             uint32_t* __restrict__ packed) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_words) return;
    uint32_t lo = 0, hi = n_entries;          // last e with woff[e] <= g
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (woff[mid] <= g) lo = mid; else hi = mid; }
    const uint32_t e = lo, r = e >> 1;
    const bool comp = (e & 1u) != 0;
    const int full = rlen[r], len = elen[e];
    const uint8_t* src = bps + boff[r];
    const int64_t p0 = (int64_t)(g - woff[e]) * 16;
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const int64_t p = p0 + b;
        if (p < len) {
            const int64_t sp = comp ? (int64_t)full - 1 - p : p;
            uint32_t v = ((uint32_t)src[sp >> 2] >> (6 - 2 * (int)(sp & 3))) & 3u;
            if (comp) v = 3u - v;
            word |= v << (2 * b);
        }
    }
    packed[g] = word;
}

// ------------------------------------------------------------------------------ k_index
// One CTA per seed.  tab[] (65536 uint32, zeroed by the host) becomes, per bucket, the END offset
// into kpos[]; bucket k spans [k ? tab[k-1] : 0, tab[k]), positions ascending -- the order the
// reference's start/next chain yields (kmer_lookup.c:174-191, 257-282).  Positions 0..slen-9 are
// indexed (loop bound `i < seq_len - K`).
constexpr int INDEX_BIG_CAP = 3200;     // buckets of > 32 positions: at most 100000 / 33
constexpr int INDEX_MULTI_CAP = 8192;   // buckets with >= 2 positions listed in shared memory (16 KB)
__global__ void __launch_bounds__(256) k_index(const BlockDesc* __restrict__ blocks,
                                               const uint32_t* __restrict__ pool,
                                               uint32_t* __restrict__ ktab,
                                               uint32_t* __restrict__ kpos_arena,
                                               uint32_t* __restrict__ kbits) {
    const BlockDesc bd = blocks[blockIdx.x];
    const uint32_t* seed = pool + bd.seed_woff;
    uint32_t* tab = ktab + (size_t)blockIdx.x * KTAB;
    uint32_t* kpos = kpos_arena + bd.kpos_off;
    const int n = bd.slen - KMER;
    const int tid = threadIdx.x;
    // bitmap of the non-empty buckets (8 KB per seed): k_range keeps it in shared memory and only
    // goes to the 256 KB table for the ~1 in 4 query k-mers that can have a hit
    uint32_t* bits = kbits + (size_t)blockIdx.x * (KTAB / 32);
    if (n <= 0) { for (int j = tid; j < KTAB / 32; j += 256) bits[j] = 0u; return; }
    // buckets with two or more positions are listed (by the scan below): only they need ordering later
    __shared__ uint16_t s_multi[INDEX_MULTI_CAP];
    __shared__ uint32_t s_nmulti, s_nbig;
    if (tid == 0) { s_nmulti = 0; s_nbig = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += 256) atomicAdd(&tab[fetch16(seed, i) & 0xffffu], 1u);
    __syncthreads();
    // exclusive scan, 256 entries per thread
    __shared__ uint32_t part[256];
    uint32_t sum = 0;
    uint32_t* mine = tab + tid * 256;
    for (int w = 0; w < 8; w++) {
        uint32_t bw = 0;
        for (int j = 0; j < 32; j++) { const uint32_t c = mine[w * 32 + j]; sum += c; bw |= (c ? 1u : 0u) << j; }
        bits[tid * 8 + w] = bw;
    }
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) { uint32_t run = 0; for (int j = 0; j < 256; j++) { uint32_t v = part[j]; part[j] = run; run += v; } }
    __syncthreads();
    uint32_t run = part[tid];
    for (int j = 0; j < 256; j++) {
        const uint32_t v = mine[j]; mine[j] = run; run += v;
        if (v >= 2u) { const uint32_t at = atomicAdd(&s_nmulti, 1u); if (at < INDEX_MULTI_CAP) s_multi[at] = (uint16_t)(tid * 256 + j); }
    }
    __syncthreads();
    // Fill.  tab[k] is the cursor of bucket k and ends as its END offset.  All threads scatter their
    // positions with an atomic cursor (order inside a bucket arbitrary), then every bucket with two or
    // more entries is put into ascending order -- the order the reference's start/next chain yields:
    // small buckets (the rule: a 15 kb seed has < 1 position per bucket on average) by an insertion
    // sort in the thread that owns the bucket, large ones (low-complexity seeds) by a warp that
    // re-scans the seed and writes the positions of that k-mer in order.
    for (int i = tid; i < n; i += 256) kpos[atomicAdd(&tab[fetch16(seed, i) & 0xffffu], 1u)] = (uint32_t)i;
    __shared__ uint16_t s_big[INDEX_BIG_CAP];
    __syncthreads();
    volatile uint32_t* vtab = tab;                 // (the cursors were advanced by atomics: read them past L1)
    volatile uint32_t* vpos = kpos;
    const uint32_t nmulti = s_nmulti;
    const bool listed = nmulti <= INDEX_MULTI_CAP;  // otherwise (long, repetitive seeds) every bucket is looked at
    const int n_todo = listed ? (int)nmulti : KTAB;
    for (int t = tid; t < n_todo; t += 256) {
        const int k = listed ? (int)s_multi[t] : t;
        const uint32_t e = vtab[k], s0 = k ? vtab[k - 1] : 0u;
        const uint32_t c = e - s0;
        if (c < 2) continue;
        if (c > 32) { const uint32_t at = atomicAdd(&s_nbig, 1u); if (at < INDEX_BIG_CAP) s_big[at] = (uint16_t)k; continue; }
        for (uint32_t a = s0 + 1; a < e; a++) {
            const uint32_t v = vpos[a];
            uint32_t b = a;
            while (b > s0 && vpos[b - 1] > v) { vpos[b] = vpos[b - 1]; b--; }
            vpos[b] = v;
        }
    }
    __syncthreads();
    const uint32_t nbig = min(s_nbig, (uint32_t)INDEX_BIG_CAP);     // (more than INDEX_BIG_CAP buckets of > 32 positions
    const int lane = tid & 31;                                      //  need a seed longer than 32 * INDEX_BIG_CAP)
    const unsigned lt = lanemask_lt();
    for (uint32_t j = tid >> 5; j < nbig; j += 8) {
        const uint32_t k = s_big[j];
        uint32_t w = k ? vtab[k - 1] : 0u;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int i = b0 + lane;
            const bool hit = i < n && (fetch16(seed, i) & 0xffffu) == k;
            const unsigned hb = __ballot_sync(FULL, hit);
            if (hit) kpos[w + __popc(hb & lt)] = (uint32_t)i;
            w += __popc(hb);
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
constexpr int RANGE_WARPS_MAX = 16;     // warps per CTA are chosen per wave (shared-memory histogram size)
constexpr int RANGE_BINS = 4224;   // >= (99999 + 99999) / 48 + 1

// Slow path (match list does not fit the per-warp scratch): every pass re-walks the buckets.
__device__ void range_pair_slow(const uint32_t* __restrict__ read, const uint32_t* __restrict__ tab,
                                const uint32_t* __restrict__ kpos, const int nq, int* hist, const int lane,
                                PairRange* __restrict__ out) {
    const int p = 0;               // `out` already points at this pair's slot
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


// Fast path: the match list (query position ascending, seed position ascending -- the order of
// kmer_lookup.c:252-283) is materialised once into a per-warp global scratch, then the histogram,
// arg-max and Kadane passes stream over it with coalesced loads.  Persistent warps: each warp owns
// one scratch slot and loops over pairs.
constexpr int RANGE_LIST_CAP = 16384;   // default capacity of the per-warp match list; the engine raises it for long reads

// A match (query position i = 0, 4, 8, ... < 100000, seed position t < 100000) packed in 32 bits
__device__ __forceinline__ uint32_t rm_pack(int i, int t) { return ((uint32_t)(i >> 2) << 17) | (uint32_t)t; }
__device__ __forceinline__ int rm_q(uint32_t m) { return (int)(m >> 17) << 2; }
__device__ __forceinline__ int rm_t(uint32_t m) { return (int)(m & 0x1ffffu); }

// Persistent CTAs, one seed block at a time: the block's non-empty-bucket bitmap (8 KB, from k_index)
// sits in shared memory, its 256 KB bucket table and position list stay hot in L2 because all
// pairs of the block are looked up back to back by the (up to 16) warps of ONE CTA and only one or
// two CTAs run per SM: the tables in use at any time (~150-300 x 256 KB) fit the 126 MB L2.
__global__ void __launch_bounds__(RANGE_WARPS_MAX * 32)
k_range(const BlockDesc* __restrict__ blocks, uint32_t n_blocks, const PairDesc* __restrict__ pairs,
        const uint32_t* __restrict__ pool, const uint32_t* __restrict__ ktab,
        const uint32_t* __restrict__ kpos_arena, const uint32_t* __restrict__ kbits,
        uint32_t* __restrict__ list_scratch, int list_cap, int bins, PairRange* __restrict__ out) {
    FCX_DYN_SHARED(int, s_dyn_all);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* sbits = reinterpret_cast<uint32_t*>(s_dyn_all);                 // KTAB / 32 words
    int* hist = s_dyn_all + KTAB / 32 + wib * bins;      // bins >= (max read len + max seed len) / 48 + 2 for this wave
    __shared__ uint32_t s_next_pair;
    const int n_warps = (int)(blockDim.x >> 5);
    const uint32_t gw = blockIdx.x * n_warps + wib;
    uint32_t* list = list_scratch + (size_t)gw * list_cap;
    const unsigned lt = lanemask_lt();
  for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const BlockDesc bd = blocks[blk];
    __syncthreads();                                     // every warp is done with the previous bitmap
    for (int j = threadIdx.x; j < KTAB / 32; j += (int)blockDim.x) sbits[j] = __ldg(kbits + (size_t)blk * (KTAB / 32) + j);
    if (threadIdx.x == 0) s_next_pair = 0u;
    __syncthreads();
    const uint32_t* tab = ktab + (size_t)blk * KTAB;
    const uint32_t* kpos = kpos_arena + bd.kpos_off;
    // the block's pairs are handed out dynamically (they differ in read length: a static split leaves
    // warps waiting at the barrier of the next block)
    for (;;) {
        uint32_t pi = 0;
        if (lane == 0) pi = atomicAdd(&s_next_pair, 1u);
        pi = __shfl_sync(FULL, pi, 0);
        if (pi >= bd.n_pairs) break;
        const uint32_t p = bd.pair_begin + pi;
        const PairDesc pd = pairs[p];
        const uint32_t* read = pool + pd.read_woff;
        const int nq = pd.rlen > KMER ? (pd.rlen - KMER + 3) / 4 : 0;   // i = 0,4,.. < rlen-K
        PairRange r; r.s1 = r.e1 = r.s2 = r.e2 = 0; r.n_match = 0; r.pass = 0;
        // ---- materialise the match list
        int n = 0, dmin = INT_MAX, dmax = INT_MIN; bool overflow = false;
        for (int it0 = 0; it0 < nq; it0 += 32) {
            const int it = it0 + lane, i = it * 4;
            uint32_t s = 0, e = 0;
            if (it < nq) {
                const uint32_t kid = fetch16(read, i) & 0xffffu;
                if ((sbits[kid >> 5] >> (kid & 31)) & 1u) { s = kid ? __ldg(tab + kid - 1) : 0u; e = __ldg(tab + kid); }
            }
            const int c = (int)(e - s);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
            const int total = __shfl_sync(FULL, incl, 31);
            if (n + total > list_cap) { overflow = true; break; }
            int w = n + incl - c;
            for (uint32_t j = s; j < e; j++, w++) {
                const int t = (int)__ldg(kpos + j);
                list[w] = rm_pack(i, t);
                dmin = min(dmin, i - t); dmax = max(dmax, i - t);
            }
            n += total;
        }
        if (overflow) { __syncwarp(); range_pair_slow(read, tab, kpos, nq, hist, lane, out + p); __syncwarp(); continue; }
        r.n_match = n;
        if (n == 0) { if (lane == 0) out[p] = r; continue; }
        dmin = __reduce_min_sync(FULL, dmin); dmax = __reduce_max_sync(FULL, dmax);
        const int nbin = (dmax - dmin) / BIN_SIZE + 1;
        for (int b = lane; b < nbin; b += 32) hist[b] = 0;
        __syncwarp();      // also orders the list writes before the reads below
        // ---- histogram (kmer_lookup.c:346-355)
        for (int e0 = 0; e0 < n; e0 += 32) {
            const int e = e0 + lane;
            if (e < n) { const uint32_t m = list[e]; atomicAdd(&hist[(rm_q(m) - rm_t(m) - dmin) / BIN_SIZE], 1); }
        }
        __syncwarp();
        int top = 0;
        for (int b = lane; b < nbin; b += 32) top = max(top, hist[b]);
        top = __reduce_max_sync(FULL, top);
        if (top <= COUNT_TH) { if (lane == 0) out[p] = r; __syncwarp(); continue; }
        // ---- arg-max bin = bin of the first match whose bin holds `top` (:357-366)
        int top_bin = -1;
        for (int e0 = 0; e0 < n && top_bin < 0; e0 += 32) {
            const int e = e0 + lane; int mybin = -1;
            if (e < n) { const uint32_t m = list[e]; const int b = (rm_q(m) - rm_t(m) - dmin) / BIN_SIZE; if (hist[b] == top) mybin = b; }
            const unsigned bal = __ballot_sync(FULL, mybin >= 0);
            if (bal) top_bin = __shfl_sync(FULL, mybin, __ffs(bal) - 1);
        }
        // ---- kept matches + Kadane scan in closed form (:369-411), see the note above
        int idx_base = 0;
        bool have_min = false; int c_min = 0, c_mq = 0, c_mt = 0;
        int best = 0, bs1 = 0, bs2 = 0, be1 = 0, be2 = 0;
        bool first_set = false; int q0 = 0, t0 = 0;
        for (int e0 = 0; e0 < n; e0 += 32) {
            const int e = e0 + lane;
            bool keep = false; int qi = 0, ti = 0;
            if (e < n) {
                const uint32_t m = list[e]; qi = rm_q(m); ti = rm_t(m);
                const int b = (qi - ti - dmin) / BIN_SIZE;
                keep = abs(b - top_bin) <= 5 && hist[b] > COUNT_TH;
            }
            const unsigned kb = __ballot_sync(FULL, keep);
            if (!kb) continue;
            const int excl = __popc(kb & lt), total = __popc(kb);
            if (!first_set) { const int fl = __ffs(kb) - 1; q0 = __shfl_sync(FULL, qi, fl); t0 = __shfl_sync(FULL, ti, fl); first_set = true; }
            int sv = keep ? 32 * (idx_base + excl) - qi : INT_MAX;      // fits: idx < RANGE_LIST_CAP_MAX (2^20), q < 100000
            int src = lane;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ov = __shfl_up_sync(FULL, sv, o), os = __shfl_up_sync(FULL, src, o);
                if (lane >= o && ov <= sv) { sv = ov; src = os; }        // earlier lane wins ties
            }
            const bool from_carry = have_min && c_min <= sv;
            const int mval = from_carry ? c_min : sv;
            int sq = __shfl_sync(FULL, qi, src), st = __shfl_sync(FULL, ti, src);
            if (from_carry) { sq = c_mq; st = c_mt; }
            const int cval = keep ? 32 * (idx_base + excl) - qi - mval : -1;
            const int cmax = __reduce_max_sync(FULL, cval);
            if (cmax > best) {
                const unsigned who = __ballot_sync(FULL, cval == cmax);
                const int wl = __ffs(who) - 1;
                best = cmax;
                be1 = __shfl_sync(FULL, qi, wl); be2 = __shfl_sync(FULL, ti, wl);
                bs1 = __shfl_sync(FULL, sq, wl); bs2 = __shfl_sync(FULL, st, wl);
            }
            const int lv = __shfl_sync(FULL, mval, 31), lq = __shfl_sync(FULL, sq, 31), ltt = __shfl_sync(FULL, st, 31);
            if (lv != INT_MAX) { have_min = true; c_min = lv; c_mq = lq; c_mt = ltt; }
            idx_base += total;
        }
        if (idx_base > 1) {
            if (best > 0) { r.s1 = bs1; r.s2 = bs2; r.e1 = be1; r.e2 = be2; }
            else { r.s1 = r.e1 = q0; r.s2 = r.e2 = t0; }
        }
        // span filters, falcon.c:612-619 (double arithmetic kept as written there)
        const int sp1 = r.e1 - r.s1, sp2 = r.e2 - r.s2;
        r.pass = !(sp1 < 100 || sp2 < 100 || abs(sp1 - sp2) > (int)(0.5 * 0.10 * (sp1 + sp2)));
        if (lane == 0) out[p] = r;
        __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------ k_dp
// One warp per pair: furthest-reaching banded O(ND) forward pass (DW_banded.c:149-258).
// Lanes own diagonals k = min_k + 2*lane (+64 per extra band chunk).  V lives in a 512-entry
// shared-memory ring indexed by k (both parities interleaved; the in-place update of
// DW_banded.c:213 only ever reads the opposite parity, written at d-1).  Per step one 32-byte
// trace record is written: [min_k, ballot words of "came from k+1"].  x2/y2 are NOT stored; the
// traceback kernel re-walks the path and recomputes the snakes.
// Steps whose band holds <= 64 cells (virtually all) run a register-resident fast path with one
// or two cells per lane; wider bands (<= 151 cells) take the generic chunk loop.
constexpr int DP_WARPS = 4;
constexpr int VRING = 512;

template <bool SM = false>
__device__ __forceinline__ void snake(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t,
                                      int qs, int ts, int q_len, int t_len, int& x, int& y) {
    for (;;) {
        int rem = min(q_len - x, t_len - y);
        if (rem <= 0) break;
        uint32_t diff = fetch16_t<SM>(q, qs + x) ^ fetch16_t<SM>(t, ts + y);
        int n = diff ? (__ffs(diff) - 1) >> 1 : 16;
        n = min(n, rem);
        x += n; y += n;
        if (n < 16) break;
    }
}

struct DpCell { int x, y; bool up; };

// predecessor choice of one cell (DW_banded.c:190-197); inactive lanes get a harmless (0,0)
__device__ __forceinline__ DpCell dp_pick(const int* V, int k, int min_k, int max_k, bool act) {
    DpCell c;
    const int vm = V[(k - 1) & (VRING - 1)], vp = V[(k + 1) & (VRING - 1)];
    const bool up = (k == min_k) || (k != max_k && vm < vp);
    c.up = act && up;
    c.x = act ? (up ? vp : vm + 1) : 0;
    c.y = act ? c.x - k : 0;
    return c;
}

// one 16-base compare (DW_banded.c:203-206), straight-line so that all lanes stay converged;
// returns the advance (16 = all compared bases matched and neither end was reached: go on)
template <bool SM = false>
__device__ __forceinline__ int snake16(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t,
                                       int qs, int ts, int q_len, int t_len, bool act, int& x, int& y) {
    const int rem = act ? max(min(q_len - x, t_len - y), 0) : 0;
    const uint32_t diff = fetch16_t<SM>(q, qs + x) ^ fetch16_t<SM>(t, ts + y);
    int n = diff ? (__ffs(diff) - 1) >> 1 : 16;
    n = min(n, rem);
    x += n; y += n;
    return n;
}

// full DP cell for the generic (wide band) path
template <bool SM = false>
__device__ __forceinline__ DpCell dp_cell(const int* V, int d, int k, int min_k, int max_k,
                                          const uint32_t* __restrict__ q, const uint32_t* __restrict__ t,
                                          int qs, int ts, int q_len, int t_len) {
    DpCell c;
    if (d == 0) { c.up = true; c.x = 0; c.y = -k; }        // V[k+1] is calloc'd 0
    else c = dp_pick(V, k, min_k, max_k, true);
    snake<SM>(q, t, qs, ts, q_len, t_len, c.x, c.y);
    return c;
}

// TMA (1-D bulk async copy, cp.async.bulk -> SASS UBLKCP) + mbarrier helpers for the staged variant
#ifndef FCX_EMU
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // init visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(b), "r"(phase) : "memory");
    }
}

#else   // the SIMT emulator has no async proxy: the staged variant is never launched there
__device__ __forceinline__ void mbar_init(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void tma_load_1d(void*, const void*, uint32_t, uint64_t*) {}
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) {}
#endif

// STAGED = true: the two packed spans of the pair are copied into shared memory by one TMA bulk
// copy each (dynamic shared memory: per warp V ring + 2 x stage_words words + an mbarrier) and the
// snakes read shared memory; STAGED = false reads them through the read-only L1 path.
template <bool STAGED>
__global__ void __launch_bounds__(DP_WARPS * 32)
k_dp(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs,
     const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs, uint32_t n_pairs,
     const uint32_t* __restrict__ pool, uint32_t* __restrict__ trace_arena, double max_diff,
     int stage_words, PairAln* __restrict__ out) {
    FCX_DYN_SHARED(unsigned char, s_dyn);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * DP_WARPS + wib;
    if (p >= n_pairs) return;
    const size_t per_warp = (size_t)VRING * 4 + (STAGED ? (size_t)stage_words * 8 + 16 : 0);
    unsigned char* mine = s_dyn + (size_t)wib * per_warp;
    int* V = reinterpret_cast<int*>(mine);
    PairAln res; res.aligned = res.dist = res.aln_size = res.q_e = res.t_e = res.k_end = 0;
    res.accepted = res.t_cnt = res.n_tags = res.cells = 0;
    const PairRange rg = ranges[p];
    if (!rg.pass) { if (lane == 0) out[p] = res; return; }
    const PairDesc pd = pairs[p];
    const uint32_t* q = pool + pd.read_woff;
    const uint32_t* t = pool + blocks[pd.block].seed_woff;
    const int qs = rg.s1, ts = rg.s2, q_len = rg.e1 - rg.s1, t_len = rg.e2 - rg.s2;
    if (STAGED) {
        uint32_t* sq = reinterpret_cast<uint32_t*>(mine + VRING * 4);
        uint32_t* st = sq + stage_words;
        uint64_t* bar = reinterpret_cast<uint64_t*>(st + stage_words);
        // word windows, 16-byte aligned at both ends (+1 word: fetch16 reads one word ahead)
        const int wq0 = (qs >> 4) & ~3, wq1 = (((qs + q_len + 15) >> 4) + 1 + 3) & ~3;
        const int wt0 = (ts >> 4) & ~3, wt1 = (((ts + t_len + 15) >> 4) + 1 + 3) & ~3;
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (uint32_t)((wq1 - wq0) + (wt1 - wt0)) * 4u);
            tma_load_1d(sq, q + wq0, (uint32_t)(wq1 - wq0) * 4u, bar);
            tma_load_1d(st, t + wt0, (uint32_t)(wt1 - wt0) * 4u, bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        q = sq - wq0; t = st - wt0;
    }
    int max_d = (int)(0.3 * (q_len + t_len));                 // DW_banded.c:149
#ifndef FCX_EMU
    asm volatile("" : "+r"(max_d));     // keep the FP64 conversion out of the d loop (ptxas rematerialises it)
#endif
    const int band_size = BAND_TOL * 2;                        // :151
    const PairAlloc al = allocs[p];
    uint32_t* trace = trace_arena + al.trace_off * TRACE_REC_WORDS;
    const int trace_cap = (int)al.trace_cap;

    int best_m = -1, min_k = 0, max_k = 0, cells = 0;
    bool aligned = false; int end_d = 0, end_k = 0, end_x = 0, end_y = 0;
    // ---- d = 0 peeled: the single cell k = 0 starts at (0,0) (V is calloc'd, DW_banded.c:153,190-192)
    if (max_d > 0) {
        int x = 0, y = 0;
        snake<STAGED>(q, t, qs, ts, q_len, t_len, x, y);               // same on every lane
        if (lane == 0) { trace[0] = 0u; trace[1] = 1u; }
        cells = 1;
        if (x >= q_len || y >= t_len) { aligned = true; end_x = x; end_y = y; }
        else { if (lane == 0) V[0] = x; best_m = x + y; min_k = -1; max_k = 1; }
        __syncwarp();
    }
    for (int d = 1; d < max_d && !aligned; d++) {
        if (max_k - min_k > band_size) break;                  // :184-186
        const int ncell = ((max_k - min_k) >> 1) + 1;
        uint32_t* rec = trace + (size_t)d * TRACE_REC_WORDS;
        if (ncell <= 64) {
            // ------------------------------------------------ fast step: <= 2 cells per lane
            const bool two = ncell > 32;
            const int k0 = min_k + 2 * lane, k1 = k0 + 64;
            const bool a0 = k0 <= max_k, a1 = two && k1 <= max_k;
            DpCell c0 = dp_pick(V, k0, min_k, max_k, a0), c1;
            c1.x = c1.y = 0; c1.up = false;
            int n0 = snake16<STAGED>(q, t, qs, ts, q_len, t_len, a0, c0.x, c0.y), n1 = 0;
            if (two) {
                c1 = dp_pick(V, k1, min_k, max_k, a1);
                n1 = snake16<STAGED>(q, t, qs, ts, q_len, t_len, a1, c1.x, c1.y);
            }
            while (__ballot_sync(FULL, n0 == 16 || n1 == 16)) {      // long snakes: rare
                if (n0 == 16) n0 = snake16<STAGED>(q, t, qs, ts, q_len, t_len, true, c0.x, c0.y);
                if (n1 == 16) n1 = snake16<STAGED>(q, t, qs, ts, q_len, t_len, true, c1.x, c1.y);
            }
            const unsigned up0 = __ballot_sync(FULL, c0.up);
            const unsigned up1 = two ? __ballot_sync(FULL, c1.up) : 0u;
            if (lane == 0 && d < trace_cap) {
                *reinterpret_cast<uint2*>(rec) = make_uint2((uint32_t)min_k, up0);
                if (two) rec[2] = up1;
            }
            const unsigned fin0 = __ballot_sync(FULL, a0 && (c0.x >= q_len || c0.y >= t_len));     // :220
            const unsigned fin1 = two ? __ballot_sync(FULL, a1 && (c1.x >= q_len || c1.y >= t_len)) : 0u;
            if (fin0 | fin1) {                                  // first k in ascending order wins
                const int fl = fin0 ? __ffs(fin0) - 1 : __ffs(fin1) - 1;
                aligned = true; end_d = d; end_k = (fin0 ? k0 : k1) - 2 * lane + 2 * fl;
                end_x = __shfl_sync(FULL, fin0 ? c0.x : c1.x, fl);
                end_y = __shfl_sync(FULL, fin0 ? c0.y : c1.y, fl);
                cells += (fin0 ? 0 : 32) + fl + 1;
                break;
            }
            cells += ncell;
            if (a0) V[k0 & (VRING - 1)] = c0.x;
            if (a1) V[k1 & (VRING - 1)] = c1.x;
            const int u0 = a0 ? c0.x + c0.y : INT_MIN, u1 = a1 ? c1.x + c1.y : INT_MIN;
            best_m = max(best_m, __reduce_max_sync(FULL, max(u0, u1)));
            const int thr = best_m - BAND_TOL;                  // band update, :227-243
            const unsigned ok0 = __ballot_sync(FULL, u0 >= thr);
            const unsigned ok1 = two ? __ballot_sync(FULL, u1 >= thr) : 0u;
            const int nmin = ok0 ? min_k + 2 * (__ffs(ok0) - 1) : min_k + 64 + 2 * (__ffs(ok1) - 1);
            const int nmax = ok1 ? min_k + 64 + 2 * (31 - __clz(ok1)) : min_k + 2 * (31 - __clz(ok0));
            max_k = nmax + 1; min_k = nmin - 1;
            __syncwarp();
            continue;
        }
        // ---------------------------------------------------- generic step: up to 151 cells
        const int nch = (ncell + 31) >> 5;
        const bool rec_ok = d < trace_cap;
        if (lane == 0 && rec_ok) rec[0] = (uint32_t)min_k;
        int step_best = best_m;
        for (int c = 0; c < nch; c++) {
            const int k = min_k + 2 * (lane + 32 * c);
            const bool act = k <= max_k;
            DpCell cc; cc.x = cc.y = 0; cc.up = false;
            if (act) cc = dp_cell<STAGED>(V, d, k, min_k, max_k, q, t, qs, ts, q_len, t_len);
            unsigned upb = __ballot_sync(FULL, act && cc.up);
            if (lane == 0 && rec_ok) rec[1 + c] = upb;
            const bool fin = act && (cc.x >= q_len || cc.y >= t_len);
            unsigned finb = __ballot_sync(FULL, fin);
            if (act) V[k & (VRING - 1)] = cc.x;
            int u = act ? cc.x + cc.y : INT_MIN;
            step_best = max(step_best, __reduce_max_sync(FULL, u));
            if (finb) {
                int fl = __ffs(finb) - 1;
                aligned = true; end_d = d; end_k = min_k + 2 * (fl + 32 * c);
                end_x = __shfl_sync(FULL, cc.x, fl); end_y = __shfl_sync(FULL, cc.y, fl);
                cells += fl + 1;
                break;
            }
            cells += __popc(__ballot_sync(FULL, act));
        }
        if (aligned) break;
        best_m = step_best;
        __syncwarp();
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
        if (res.accepted && end_d >= trace_cap) res.accepted = -1;     // cannot happen (bound in fcx_engine.cu); loud if it does
    }
    res.cells = cells;
    if (lane == 0) out[p] = res;
}

__device__ __forceinline__ void warp_walk_back(const uint32_t* trace, uint32_t* __restrict__ path,
                                               const int D, const int k_end, const int lane, int* stage);

}  // namespace fcx

#include "fcx_dp.cuh"

namespace fcx {

// ------------------------------------------------------------------------------ k_traceback
// (1) warp_walk_back: walk the trace records backwards collecting the direction bit of every step
// of the optimal path (DW_banded.c:264-277); (2) k_traceback, one thread per accepted pair: replay
// the path forwards, recomputing the snakes, and emit per target position y
//        ent[y] = VALID | is_match<<30 | n_ins<<22 | first 11 inserted bases (2 bits each)
// This is get_align_tags (falcon.c:106-162) in closed form: the delta-0 tag of column y carries
// the seed base (match) or '-', and the insertion tags delta = 1..n_ins are the query bases
// x+is_match .. x_next-1.  The reference stops tagging at the first column whose delta reaches 255
// (falcon.c:138,150-152): t_cnt is cut there and exactly 254 insertion tags remain.
constexpr uint32_t ENT_VALID = 0x80000000u;
constexpr uint32_t ENT_MATCH = 0x40000000u;
constexpr uint32_t ENT_PLAIN = 0xC0000000u;    // VALID | MATCH, no insertions
constexpr int ENT_INS_INLINE = 11;
__device__ __forceinline__ int ent_nins(uint32_t e) { return (int)((e >> 22) & 0xffu); }
__device__ __forceinline__ int ent_ins(uint32_t e, int k) { return (int)((e >> (2 * k)) & 3u); }

// Work ordering for k_traceback: accepted pairs bucketed by dist (32 steps per bucket, longest
// first), so that the 32 threads of a warp walk paths of similar length.  Three tiny kernels:
// histogram, exclusive scan, scatter.  The order inside a bucket is arbitrary (atomics); results do
// not depend on it.
constexpr int TB_BUCKETS = 2048;
__device__ __forceinline__ int tb_bucket(int dist) { return TB_BUCKETS - 1 - min(dist >> 5, TB_BUCKETS - 1); }

__global__ void k_tb_hist(const PairAln* __restrict__ aln, uint32_t n_pairs, uint32_t* __restrict__ hist) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const PairAln a = aln[p];
    if (a.accepted > 0) atomicAdd(&hist[tb_bucket(a.dist)], 1u);
}
// one CTA of 1024 threads, two buckets each; hist[] becomes the exclusive scan, hist[TB_BUCKETS] the total
__global__ void __launch_bounds__(1024) k_tb_scan(uint32_t* __restrict__ hist) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    const uint32_t a = hist[2 * t], b = hist[2 * t + 1];
    part[t] = a + b;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint32_t excl = part[t] - (a + b);
    hist[2 * t] = excl; hist[2 * t + 1] = excl + a;
    if (t == 1023) hist[TB_BUCKETS] = part[t];
}
__global__ void k_tb_scatter(const PairAln* __restrict__ aln, uint32_t n_pairs, uint32_t* __restrict__ cursor,
                             uint32_t* __restrict__ order) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const PairAln a = aln[p];
    if (a.accepted > 0) order[atomicAdd(&cursor[tb_bucket(a.dist)], 1u)] = p;
}

// Backward walk of one pair by a whole warp (DW_banded.c:264-277): bit d of path[] = "step d came
// from k+1" (a target-only column).  Lane l holds the record of step 32w + l, so one coalesced
// request brings 32 records (1 KB).  The walk is a chain over the CELL INDEX idx = (k - min_k) / 2:
//        up = bit idx of the step's ballot words;   idx' = idx + a + up,
// with a = (min_k(d) - min_k(d-1) - 1) / 2 known per record before the chain starts, so that a step
// of the chain is select / shift / and / add; the per-step operands (w0, w1, a) are staged in 512
// bytes of shared memory and read back as one broadcast LDS.128 per step.
// Called by the DP warp right after the forward pass of an accepted pair (k_dp3), or by
// k_traceback_walk for the round-1 DP kernels.  stage: 128 ints of shared memory owned by the warp.
__device__ __forceinline__ void warp_walk_back(const uint32_t* trace, uint32_t* __restrict__ path,
                                               const int D, const int k_end, const int lane, int* stage) {
    int idx = 0;
    for (int w = D >> 5; w >= 0; w--) {
        const int d_l = 32 * w + lane;
        int4 mine = make_int4(0, 0, 0, 0);                 // w0, w1, a, min_k: a no-op step
        if (d_l >= 1 && d_l <= D) {
            const uint32_t* rec = trace + (size_t)d_l * TRACE_REC_WORDS;
            const uint2 hd = *reinterpret_cast<const uint2*>(rec);
            const int mk_prev = (int)rec[-TRACE_REC_WORDS];
            mine = make_int4((int)hd.y, (int)rec[2], ((int)hd.x - mk_prev - 1) >> 1, (int)hd.x);
        }
        __syncwarp();
        reinterpret_cast<int4*>(stage)[lane] = mine;
        __syncwarp();
        if (w == (D >> 5)) idx = (k_end - reinterpret_cast<const int4*>(stage)[D & 31].w) >> 1;
        const int idx_in = idx;
        uint32_t acc = 0;
        int idx_max = idx;
#pragma unroll
        for (int l = 31; l >= 0; l--) {
            const int4 st = reinterpret_cast<const int4*>(stage)[l];
            const unsigned long long both = ((unsigned long long)(uint32_t)st.y << 32) | (uint32_t)st.x;
            const uint32_t up = (uint32_t)(both >> (idx & 63)) & 1u;
            acc |= up << l;
            idx += st.z + (int)up;
            idx_max = max(idx_max, idx);
        }
        if (idx_max >= 64) {                      // a band wider than 64 cells somewhere in this chunk (rare): redo it
            idx = idx_in; acc = 0;                // with the ballot words beyond the first two read from the record
            for (int l = 31; l >= 0; l--) {
                const int d = 32 * w + l;
                if (d < 1 || d > D) continue;
                const int4 st = reinterpret_cast<const int4*>(stage)[l];
                const uint32_t word = idx < 32 ? (uint32_t)st.x : idx < 64 ? (uint32_t)st.y
                                               : trace[(size_t)d * TRACE_REC_WORDS + 1 + (idx >> 5)];
                const uint32_t up = (word >> (idx & 31)) & 1u;
                acc |= up << l;
                idx += st.z + (int)up;
            }
        }
        if (lane == 0) path[w] = acc;
    }
}

// the same walk for pairs whose forward pass was done by k_dp<> (dp_variant 1 / 2): one warp per
// accepted pair
__global__ void __launch_bounds__(128)
k_traceback_walk(const PairAlloc* __restrict__ allocs, const uint32_t* __restrict__ order,
                 const uint32_t* __restrict__ n_order, const uint32_t* trace_arena,
                 uint32_t* __restrict__ path_arena, const PairAln* __restrict__ aln) {
    __shared__ int s_stage[4][128];
    const int lane = threadIdx.x & 31;
    const uint32_t n = *n_order;
    for (uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5); slot < n; slot += gridDim.x * 4) {
        const uint32_t p = order[slot];
        const PairAlloc al = allocs[p];
        const PairAln a = aln[p];
        warp_walk_back(trace_arena + al.trace_off * TRACE_REC_WORDS, path_arena + al.path_off, a.dist, a.k_end, lane,
                       s_stage[threadIdx.x >> 5]);
    }
}

// One thread per accepted pair: forward replay of the path.  Every target column y gets its entry
// ent[y]; entries are produced in ascending y, collected four at a time in registers and stored as
// one 16-byte vector, so that the array is written densely and exactly once (no pre-fill, no
// partial-sector read-modify-write).  xck[y >> 5] = query index at column y for every 32nd column:
// k_vote needs the query index only to fetch inserted bases beyond the 11 inline ones (xck_lookup).
__global__ void __launch_bounds__(128)
k_traceback(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs,
            const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs,
            const uint32_t* __restrict__ order, const uint32_t* __restrict__ n_order,
            const uint32_t* __restrict__ pool,
            const uint32_t* __restrict__ path_arena,
            uint32_t* __restrict__ xck_arena, uint32_t* __restrict__ ent_arena,
            VoteMeta* __restrict__ vmeta, PairAln* __restrict__ aln) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= *n_order) return;
    const uint32_t p = order[slot];
    PairAln a = aln[p];
    const PairRange rg = ranges[p];
    const PairDesc pd = pairs[p];
    const uint32_t* q = pool + pd.read_woff;
    const uint32_t* t = pool + blocks[pd.block].seed_woff;
    const int qs = rg.s1, ts = rg.s2, q_len = rg.e1 - rg.s1, t_len = rg.e2 - rg.s2;
    const PairAlloc al = allocs[p];
    const uint32_t* path = path_arena + al.path_off;
    uint32_t* xck = xck_arena + al.xck_off;
    uint32_t* ent = ent_arena + al.xam_off;           // 16-byte aligned (xam_off is a multiple of 4)
    const int D = a.dist;
    int x = 0, y = 0, run = 0, t_cnt = -1, n_match_cols = 0;
    uint32_t pend = 0; int pend_x = 0; bool open = false;     // entry of the last target column, still open
    uint32_t pw = 0;
    // x and y creep forward a few bases per step: keep three packed words of each sequence in
    // registers.  Crossing into the next word shifts the window and issues the load of the word
    // after next, which is not needed before the following crossing: the load latency stays off
    // the dependency chain x -> window -> compare -> x.  (Reads are padded: word wi + 2 exists.)
    int wq_i = -4, wt_i = -4; uint32_t q0 = 0, q1 = 0, q2 = 0, t0 = 0, t1 = 0, t2 = 0;
    auto win_q = [&](const int pos) -> uint32_t {
        const int wi = pos >> 4;
        if (wi == wq_i + 1) { q0 = q1; q1 = q2; q2 = __ldg(q + wi + 2); wq_i = wi; }
        else if (wi != wq_i) { q0 = __ldg(q + wi); q1 = __ldg(q + wi + 1); q2 = __ldg(q + wi + 2); wq_i = wi; }
        return __funnelshift_r(q0, q1, (pos & 15) << 1);
    };
    auto win_t = [&](const int pos) -> uint32_t {
        const int wi = pos >> 4;
        if (wi == wt_i + 1) { t0 = t1; t1 = t2; t2 = __ldg(t + wi + 2); wt_i = wi; }
        else if (wi != wt_i) { t0 = __ldg(t + wi); t1 = __ldg(t + wi + 1); t2 = __ldg(t + wi + 2); wt_i = wi; }
        return __funnelshift_r(t0, t1, (pos & 15) << 1);
    };
    // output cursor: yo = number of entries emitted so far (entry yo goes to ent[yo]).  Everything below
    // is straight-line, predicated code: the 32 pairs of a warp take different turns at every step
    // (deletion / insertion, snake lengths), so loops and branches here would serialise the lanes.
    int yo = 0; uint32_t b0 = 0, b1 = 0, b2 = 0;
    auto emit1 = [&](const bool on, const uint32_t e, const int xcol) {      // one entry, if `on`
        const int r = yo & 3;
        if (on && (yo & 31) == 0) xck[yo >> 5] = (uint32_t)xcol;
        if (on && r == 3) *reinterpret_cast<uint4*>(ent + (yo - 3)) = make_uint4(b0, b1, b2, e);
        b0 = (on && r == 0) ? e : b0; b1 = (on && r == 1) ? e : b1; b2 = (on && r == 2) ? e : b2;
        yo += on ? 1 : 0;
    };
    auto plain_run = [&](const int L, const int xc) {                           // 0 <= L <= 16 plain match columns
        const int r = yo & 3, total = r + L;
        const int nx = (yo + 31) & ~31;                                         // a checkpoint column inside the run?
        if (nx < yo + L) xck[nx >> 5] = (uint32_t)(xc + (nx - yo));
        uint32_t* v = ent + (yo - r);
        const bool spill = total >= 4;
        if (spill) *reinterpret_cast<uint4*>(v) = make_uint4(r > 0 ? b0 : ENT_PLAIN, r > 1 ? b1 : ENT_PLAIN, r > 2 ? b2 : ENT_PLAIN, ENT_PLAIN);
        const uint4 pl4 = make_uint4(ENT_PLAIN, ENT_PLAIN, ENT_PLAIN, ENT_PLAIN);
        if (total >= 8) *reinterpret_cast<uint4*>(v + 4) = pl4;
        if (total >= 12) *reinterpret_cast<uint4*>(v + 8) = pl4;
        if (total >= 16) *reinterpret_cast<uint4*>(v + 12) = pl4;
        // what remains open in the register buffer: entries (total & 3) of the last group
        b0 = (spill || (r == 0 && total > 0)) ? ENT_PLAIN : b0;
        b1 = (spill || (r <= 1 && total > 1)) ? ENT_PLAIN : b1;
        b2 = (spill || (r <= 2 && total > 2)) ? ENT_PLAIN : b2;
        yo += L;
    };
    for (int d = 0; d <= D; d++) {
        bool del = false;
        if (d > 0) {
            if ((d & 31) == 0 || d == 1) pw = path[d >> 5];
            del = ((pw >> (d & 31)) & 1u) != 0;                          // target-only column
            const uint32_t qb = win_q(qs + x) & 3u;                      // the query base of a query-only column
            emit1(del && open, pend | ((uint32_t)run << 22), pend_x);    // a new target column closes the open one
            const uint32_t ins_bits = run < ENT_INS_INLINE ? qb << (2 * run) : 0u;
            pend = del ? ENT_VALID : (pend | ins_bits);
            pend_x = del ? x : pend_x; open = true;
            run = del ? 0 : run + 1;
            y += del ? 1 : 0; x += del ? 0 : 1;
            if (run == 255) {            // the 255th consecutive query-only column: tags stop before it
                t_cnt = y;               // target positions 0..y-1 carry tags
                x--; run = 254;          // exactly 254 insertion tags remain on column y-1
                break;
            }
        }
        // snake: interior match columns are plain entries
        const int rem = min(q_len - x, t_len - y);
        uint32_t diff = win_q(qs + x) ^ win_t(ts + y);
        int adv = min((int)((unsigned)(__ffs(diff) - 1) >> 1), max(min(rem, 16), 0));
        while (adv > 0 && (adv & 15) == 0 && adv < rem) {                // long snake: rare
            diff = win_q(qs + x + adv) ^ win_t(ts + y + adv);
            const int nn = min((int)((unsigned)(__ffs(diff) - 1) >> 1), min(rem - adv, 16));
            adv += nn;
            if (nn < 16) break;
        }
        const bool snake_on = adv > 0;
        emit1(snake_on && open, pend | ((uint32_t)run << 22), pend_x);   // the snake's first column closes the open one
        int left = snake_on ? adv - 1 : 0, xc = x;                       // the last match column stays open
        for (; left > 16; left -= 16, xc += 16) plain_run(16, xc);       // (only after a long snake)
        plain_run(left, xc);
        x += adv; y += adv; n_match_cols += adv;
        pend = snake_on ? ENT_PLAIN : pend; pend_x = snake_on ? x - 1 : pend_x; run = snake_on ? 0 : run;
        open = open || snake_on;
    }
    emit1(open, pend | ((uint32_t)run << 22), pend_x);
    {   // the entries of the last, incomplete vector
        const int r = yo & 3, base = yo - r;
        if (r > 0) ent[base] = b0;
        if (r > 1) ent[base + 1] = b1;
        if (r > 2) ent[base + 2] = b2;
    }
    if (t_cnt < 0) t_cnt = y;
    a.t_cnt = t_cnt;
    a.n_tags = t_cnt + (x - n_match_cols);            // target columns + query-only columns among them
    aln[p] = a;
    VoteMeta vm; vm.ent_off = al.xam_off; vm.q_woff = pd.read_woff; vm.t_start = rg.s2; vm.t_cnt = t_cnt; vm.q_s = rg.s1; vm.xck_off = al.xck_off;
    vmeta[p] = vm;                                    // (zero = "not accepted" for every other pair)
}

// query index x at target column y of one read: the checkpoint of the enclosing 32-column group,
// then one base per match column and per inserted base of the columns in between
__device__ __forceinline__ int xck_lookup(const uint32_t* __restrict__ xck, const uint32_t* __restrict__ ent, int y) {
    int x = (int)xck[y >> 5];
    for (int j = y & ~31; j < y; j++) { const uint32_t e = ent[j]; x += ((e & ENT_MATCH) ? 1 : 0) + ent_nins(e); }
    return x;
}

// ------------------------------------------------------------------------------ k_align1 / k_align1_tb
// The legacy single-pair entry align(q, q_len, t, t_len, band_tolerance, get_aln_str)
// (DW_banded.c:115-330) with an arbitrary band: one warp, generic chunk loop, V ring of AL_VRING
// entries, trace records of `rec_words` words per step; then a single-thread traceback that writes
// the two gapped strings (DW_banded.c:264-319).  API surface for falcon_kit.get_alignment /
// graph_to_contig.get_aln_data; the consensus throughput path uses k_dp.
constexpr int AL_VRING = 8192;

__global__ void __launch_bounds__(32)
k_align1(const uint32_t* __restrict__ pool, uint64_t q_woff, uint64_t t_woff, int q_len, int t_len,
         int band_tol, uint32_t* __restrict__ trace, int rec_words, PairAln* __restrict__ out) {
    __shared__ int V[AL_VRING];
    const int lane = threadIdx.x;
    const uint32_t* q = pool + q_woff; const uint32_t* t = pool + t_woff;
    PairAln res; res.aligned = res.dist = res.aln_size = res.q_e = res.t_e = res.k_end = 0;
    res.accepted = res.t_cnt = res.n_tags = res.cells = 0;
    const int max_d = (int)(0.3 * (q_len + t_len));
    const int band_size = band_tol * 2;
    int best_m = -1, min_k = 0, max_k = 0;
    bool aligned = false; int end_d = 0, end_k = 0, end_x = 0, end_y = 0;
    for (int d = 0; d < max_d; d++) {
        if (max_k - min_k > band_size) break;
        const int ncell = ((max_k - min_k) >> 1) + 1;
        const int nch = (ncell + 31) >> 5;
        uint32_t* rec = trace + (size_t)d * rec_words;
        if (lane == 0) rec[0] = (uint32_t)min_k;
        int step_best = best_m;
        for (int c = 0; c < nch; c++) {
            const int k = min_k + 2 * (lane + 32 * c);
            const bool act = k <= max_k;
            int x = 0, y = 0; bool up = false;
            if (act) {
                if (d == 0) up = true;
                else {
                    const int vm = V[(k - 1) & (AL_VRING - 1)], vp = V[(k + 1) & (AL_VRING - 1)];
                    up = (k == min_k) || (k != max_k && vm < vp);
                    x = up ? vp : vm + 1;
                }
                y = x - k;
                snake(q, t, 0, 0, q_len, t_len, x, y);
            }
            const unsigned upb = __ballot_sync(FULL, act && up);
            if (lane == 0) rec[1 + c] = upb;
            const unsigned finb = __ballot_sync(FULL, act && (x >= q_len || y >= t_len));
            if (act) V[k & (AL_VRING - 1)] = x;
            step_best = max(step_best, __reduce_max_sync(FULL, act ? x + y : INT_MIN));
            if (finb) {
                const int fl = __ffs(finb) - 1;
                aligned = true; end_d = d; end_k = min_k + 2 * (fl + 32 * c);
                end_x = __shfl_sync(FULL, x, fl); end_y = __shfl_sync(FULL, y, fl);
                break;
            }
        }
        if (aligned) break;
        best_m = step_best;
        __syncwarp();
        int nmin = INT_MAX, nmax = INT_MIN;
        const int thr = best_m - band_tol;
        for (int c = 0; c < nch; c++) {
            const int k = min_k + 2 * (lane + 32 * c);
            bool ok = false;
            if (k <= max_k) { const int x = V[k & (AL_VRING - 1)]; ok = (2 * x - k) >= thr; }
            const unsigned okb = __ballot_sync(FULL, ok);
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
        res.aln_size = (end_x + end_y + end_d) / 2;
    }
    if (lane == 0) *out = res;
}

// Batched, distance-only form of the same alignment for stage-2 style callers
// (falcon_kit/mains/graph_to_contig.py:50-103 keeps only aln_str_size and dist of
// DWA.align(q[s1:e1], .., t[s2:e2], .., 1500, 1)): one warp per job, no trace.
struct AlignJob { uint64_t q_woff, t_woff; int32_t qs, q_len, ts, t_len; };

__global__ void __launch_bounds__(32)
k_align_batch(const uint32_t* __restrict__ pool, const AlignJob* __restrict__ jobs, uint32_t n_jobs, int band_tol,
              PairAln* __restrict__ out) {
    __shared__ int V[AL_VRING];
    const int lane = threadIdx.x;
    for (uint32_t jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        const AlignJob J = jobs[jb];
        const uint32_t* q = pool + J.q_woff; const uint32_t* t = pool + J.t_woff;
        const int qs = J.qs, ts = J.ts, q_len = J.q_len, t_len = J.t_len;
        PairAln res; res.aligned = res.dist = res.aln_size = res.q_e = res.t_e = res.k_end = 0;
        res.accepted = res.t_cnt = res.n_tags = res.cells = 0;
        const int max_d = (int)(0.3 * (q_len + t_len));
        const int band_size = band_tol * 2;
        int best_m = -1, min_k = 0, max_k = 0;
        bool aligned = false; int end_d = 0, end_k = 0, end_x = 0, end_y = 0;
        __syncwarp();
        for (int d = 0; d < max_d; d++) {
            if (max_k - min_k > band_size) break;
            const int ncell = ((max_k - min_k) >> 1) + 1;
            const int nch = (ncell + 31) >> 5;
            int step_best = best_m;
            for (int c = 0; c < nch; c++) {
                const int k = min_k + 2 * (lane + 32 * c);
                const bool act = k <= max_k;
                int x = 0, y = 0;
                if (act) {
                    if (d > 0) {
                        const int vm = V[(k - 1) & (AL_VRING - 1)], vp = V[(k + 1) & (AL_VRING - 1)];
                        const bool up = (k == min_k) || (k != max_k && vm < vp);
                        x = up ? vp : vm + 1;
                    }
                    y = x - k;
                    snake(q, t, qs, ts, q_len, t_len, x, y);
                }
                const unsigned finb = __ballot_sync(FULL, act && (x >= q_len || y >= t_len));
                if (act) V[k & (AL_VRING - 1)] = x;
                step_best = max(step_best, __reduce_max_sync(FULL, act ? x + y : INT_MIN));
                if (finb) {
                    const int fl = __ffs(finb) - 1;
                    aligned = true; end_d = d; end_k = min_k + 2 * (fl + 32 * c);
                    end_x = __shfl_sync(FULL, x, fl); end_y = __shfl_sync(FULL, y, fl);
                    break;
                }
            }
            if (aligned) break;
            best_m = step_best;
            __syncwarp();
            int nmin = INT_MAX, nmax = INT_MIN;
            const int thr = best_m - band_tol;
            for (int c = 0; c < nch; c++) {
                const int k = min_k + 2 * (lane + 32 * c);
                bool ok = false;
                if (k <= max_k) { const int x = V[k & (AL_VRING - 1)]; ok = (2 * x - k) >= thr; }
                const unsigned okb = __ballot_sync(FULL, ok);
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
            res.aln_size = (end_x + end_y + end_d) / 2;
        }
        if (lane == 0) out[jb] = res;
    }
}

__global__ void k_align1_tb(const uint32_t* __restrict__ pool, uint64_t q_woff, uint64_t t_woff, int q_len, int t_len,
                            const uint32_t* __restrict__ trace, int rec_words, uint32_t* __restrict__ path,
                            const PairAln* __restrict__ alnp, char* __restrict__ q_aln, char* __restrict__ t_aln) {
    const PairAln a = *alnp;
    if (!a.aligned) return;
    const uint32_t* q = pool + q_woff; const uint32_t* t = pool + t_woff;
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const int D = a.dist;
    int k = a.k_end;
    for (int d = D; d >= 1; d--) {
        const uint32_t* rec = trace + (size_t)d * rec_words;
        const int idx = (k - (int)rec[0]) >> 1;
        const uint32_t up = (rec[1 + (idx >> 5)] >> (idx & 31)) & 1u;
        if (up) path[d >> 5] |= 1u << (d & 31);       // path[] is zeroed by the host
        k += up ? 1 : -1;
    }
    int x = 0, y = 0, pos = 0;
    for (int d = 0; d <= D; d++) {
        if (d > 0) {
            if ((path[d >> 5] >> (d & 31)) & 1u) { q_aln[pos] = '-'; t_aln[pos] = ACGT[base_at(t, y)]; y++; pos++; }
            else { q_aln[pos] = ACGT[base_at(q, x)]; t_aln[pos] = '-'; x++; pos++; }
        }
        while (x < q_len && y < t_len && base_at(q, x) == base_at(t, y)) {
            const char c = ACGT[base_at(q, x)];
            q_aln[pos] = c; t_aln[pos] = c; x++; y++; pos++;
        }
    }
    q_aln[pos] = 0; t_aln[pos] = 0;
}

}  // namespace fcx

#include "fcx_vote.cuh"
