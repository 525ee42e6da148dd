// fcx_vote.cuh -- the consensus of one seed block (ref: get_cns_from_align_tags,
// src/c/falcon.c:308-558) as two kernels:
//
//   k_vote     PARALLEL over seed positions.  The column vote of falcon.c:350-382 is independent per
//              target position: one thread owns one position i, visits the accepted reads of the
//              block in order (= the reference's tag order, so "first appearance" of a link is
//              simply the first read that casts it, falcon.c:232-263) and reads their per-position
//              entries ent[] straight from the read-major arrays k_traceback wrote (consecutive
//              threads = consecutive positions = coalesced; no transposed pile-up matrix).  Each
//              thread keeps its position's distinct links -- key (delta, base, predecessor delta,
//              predecessor base), count -- in a small private table in first-appearance order,
//              then writes them, stably sorted by delta, to the position's 128-byte slot.
//   k_cns_dp   SERIAL over seed positions (the longest-path DP of falcon.c:405-475 is a chain in i),
//              ONE THREAD per seed block: per position it reads the handful of links, scores them
//              against the previous column scores, writes one record per live column, keeps the
//              global best, and finally backtracks (falcon.c:479-542).  A position costs a few
//              hundred dependent thread instructions instead of the ~900 warp instructions of the
//              round-1 kernel, which had a whole warp decode, vote and reduce per position.
//
// Exactness notes: scores are exact integers (x2); a column's best link
// is chosen by strict '>' in first-appearance order; a column whose best stays <= -1 keeps score -1
// and best_p = (0,0,0) (record 0 is reserved for column (0,0,'A')); the global best is taken by
// strict '>' in (i, delta, base) order and remembers the LINK INDEX, which the reference then
// (mis)uses as the first base code of the backtrack.
#pragma once

namespace fcx {

// info = (coverage > min_cov) << 31 | t_pos << 3 | base
struct CnsRec { int32_t pred; int32_t info; int32_t score2; int32_t pad; };   // 16 bytes: one vector store / load
struct CnsOut { int32_t len; int32_t err; int32_t start; int32_t positions; };

// link key: delta << 16 | base << 13 | pred  with  pred = pred_delta << 3 | pred_base, or 0x1fff for
// the first tag of a read (p_t_pos = -1, falcon.c:118-120)
constexpr uint32_t LK_START = 0x1fffu;
constexpr int VOTE_TP = 128;         // positions (threads) per CTA
constexpr int VSLOT = 16;            // 8-byte words per position slot: [header, 15 links]
constexpr int VCAP = 160;            // links a thread can hold; more -> error 1 (retry path)

__device__ __forceinline__ uint32_t lk_key(int delta, int base, uint32_t pred) {
    return ((uint32_t)delta << 16) | ((uint32_t)base << 13) | pred;
}

// Slot layout (uint2 units): slot[0] = {n_links << 16 | coverage (<= 65535), overflow offset};
// slot[1..15] = links {key, count}; links 15.. live in the overflow arena at `overflow offset`.
__global__ void __launch_bounds__(VOTE_TP)
k_vote(const BlockDesc* __restrict__ blocks, uint32_t n_blocks, uint32_t n_tiles,
       const VoteMeta* __restrict__ vmeta, const uint32_t* __restrict__ pool,
       const uint32_t* __restrict__ xck_arena, const uint32_t* __restrict__ ent_arena,
       uint2* __restrict__ slot_arena, uint2* __restrict__ ovf_arena, uint32_t ovf_cap,
       uint32_t* __restrict__ ovf_next, int* __restrict__ err_flag) {
    const uint32_t T = blockIdx.x;
    if (T >= n_tiles) return;
    uint32_t lo = 0, hi = n_blocks;                 // last block with tile_begin <= T
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (blocks[mid].tile_begin <= T) lo = mid; else hi = mid; }
    const BlockDesc bd = blocks[lo];
    const int i0 = (int)(T - bd.tile_begin) * VOTE_TP;
    const int i = i0 + (int)threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_seed = i < bd.slen;
    const uint32_t* seed = pool + bd.seed_woff;
    const int Si = in_seed ? base_at(seed, i) : 0;
    const int Sp = (in_seed && i > 0) ? base_at(seed, i - 1) : 0;

    // Link counters of this position.  Every entry carries the index of the read that cast it first
    // (`ord`), because ties between the links of a column are broken by first appearance (falcon.c:232-263,
    // :441); the entries are written out sorted by (delta, ord).
    //   * direct counters for the links that make up almost all votes -- no search:
    //       delta 0: base {seed base, '-'} x predecessor {read start, (0, seed base of i-1), (0, '-')}   6
    //       delta 1: inserted base {A,C,G,T} x predecessor (0, {seed base, '-'})                          8
    //     kept in shared memory, one column of VDIR words per thread: (ord << 16) | count;
    //     the dominant "match after a plain match" counts in a register;
    //   * everything else (a delta-0 link after an insertion, deltas >= 2) in a small private table with
    //     a linear search.
    constexpr int VDIR = 14;
    __shared__ uint32_t s_dir[VDIR * VOTE_TP];
#pragma unroll
    for (int d = 0; d < VDIR; d++) s_dir[d * VOTE_TP + threadIdx.x] = 0u;
    uint32_t key[VCAP]; uint32_t cnt[VCAP];         // generic table: key, (ord << 16) | count
    int n = 0, coverage = 0, maxd = 0; bool overflow = false;
    uint32_t c_dom = 0, o_dom = 0;                  // direct entry 1: (0, seed base) after (0, previous seed base)
    auto vote_direct = [&](const int idx, const uint32_t ord) {
        const uint32_t v = s_dir[idx * VOTE_TP + threadIdx.x];
        s_dir[idx * VOTE_TP + threadIdx.x] = v ? v + 1u : ((ord << 16) | 1u);
    };
    auto vote_generic = [&](const uint32_t k, const uint32_t ord) {
        for (int e = 0; e < n; e++) if (key[e] == k) { cnt[e]++; return; }
        if (n < VCAP) { key[n] = k; cnt[n] = (ord << 16) | 1u; n++; } else overflow = true;
    };

    // The block's reads are taken VOTE_TP at a time: the CTA first lists, in read order, those that are
    // accepted and touch this tile (one read per thread, ordered compaction), then every thread walks the
    // list -- a read costs one broadcast LDS.128 instead of a 32-byte descriptor load, range tests and
    // pointer arithmetic, and reads that do not touch the tile (about half of them) cost nothing.
    __shared__ uint4 s_list[VOTE_TP];              // {ent pointer lo, hi (biased by -t_start), t_start, t_end}
    __shared__ uint32_t s_pair[VOTE_TP];           // the pair index, for the rare lookups beyond the inline bases
    __shared__ uint32_t s_wcnt[VOTE_TP / 32];
    const unsigned lt = lanemask_lt();
    for (uint32_t c0 = 0; c0 < bd.n_pairs; c0 += VOTE_TP) {
        __syncthreads();                           // the previous list is no longer in use
        const uint32_t jj = c0 + threadIdx.x;
        bool touch = false; VoteMeta tv; tv.ent_off = 0; tv.t_start = 0; tv.t_cnt = 0;
        if (jj < bd.n_pairs) {
            tv = vmeta[bd.pair_begin + jj];
            touch = tv.t_cnt != 0 && tv.t_start < i0 + VOTE_TP && tv.t_start + tv.t_cnt > i0;
        }
        const unsigned tm = __ballot_sync(FULL, touch);
        if (lane == 0) s_wcnt[threadIdx.x >> 5] = (uint32_t)__popc(tm);
        __syncthreads();
        uint32_t at = (uint32_t)__popc(tm & lt), n_list = 0;
#pragma unroll
        for (int w = 0; w < VOTE_TP / 32; w++) { const uint32_t c = s_wcnt[w]; if (w < (int)(threadIdx.x >> 5)) at += c; n_list += c; }
        if (touch) {
            const uint64_t ep0 = (uint64_t)(ent_arena + tv.ent_off) - (uint64_t)4 * (uint64_t)(int64_t)tv.t_start;   // &ent[0] - t_start: index by i
            s_list[at] = make_uint4((uint32_t)ep0, (uint32_t)(ep0 >> 32), (uint32_t)tv.t_start, (uint32_t)(tv.t_start + tv.t_cnt));
            s_pair[at] = jj;
        }
        __syncthreads();
      for (uint32_t li = 0; li < n_list; li++) {
        const uint4 L = s_list[li];
        const int t_start = (int)L.z;
        const bool act = in_seed && i >= t_start && i < (int)L.w;
        const uint32_t* ent_i = reinterpret_cast<const uint32_t*>(((uint64_t)L.y << 32) | (uint64_t)L.x);   // ent_i[i] = the read's entry at position i
        const uint32_t ec = act ? __ldg(ent_i + i) : 0u;
        // the read's entry at position i - 1: the left neighbour thread holds it
        uint32_t ep = __shfl_up_sync(FULL, ec, 1);
        if (lane == 0) ep = (act && i > t_start) ? __ldg(ent_i + i - 1) : 0u;
        if (!act) continue;
        coverage++;
        const uint32_t ord = s_pair[li];              // read order = first-appearance order
        const int y = i - t_start;
        const int m = (ec & ENT_MATCH) ? 1 : 0, nins = ent_nins(ec);
        const int b0 = m ? Si : 4;
        // query index at this column: only needed to fetch inserted bases beyond the 11 inline ones
        int x = -1;
        const int pn = y > 0 ? ent_nins(ep) : 0;
        if (pn == 0) {
            // predecessor: the read starts here, or the previous column without insertion
            const int pc = y == 0 ? 0 : ((ep & ENT_MATCH) ? 1 : 2);
            if (m && pc == 1) { if (c_dom == 0) o_dom = ord; c_dom++; }
            else vote_direct((m ? 0 : 3) + pc, ord);
        } else {
            int pb;
            if (pn <= ENT_INS_INLINE) pb = ent_ins(ep, pn - 1);
            else {
                const VoteMeta vm = vmeta[bd.pair_begin + s_pair[li]];
                x = xck_lookup(xck_arena + vm.xck_off, ent_i + t_start, y); pb = base_at(pool + vm.q_woff, vm.q_s + x - 1);
            }
            vote_generic(lk_key(0, b0, ((uint32_t)pn << 3) | (uint32_t)pb), ord);
        }
        if (nins > 0) {
            maxd = max(maxd, nins);
            int pb = ent_ins(ec, 0);
            vote_direct(6 + 2 * pb + (m ? 0 : 1), ord);                       // delta 1
            if (nins > 1) {
                if (nins <= ENT_INS_INLINE) {
                    for (int lev = 2; lev <= nins; lev++) {
                        const int bb = ent_ins(ec, lev - 1);
                        vote_generic(lk_key(lev, bb, ((uint32_t)(lev - 1) << 3) | (uint32_t)pb), ord);
                        pb = bb;
                    }
                } else {
                    const VoteMeta vm = vmeta[bd.pair_begin + s_pair[li]];
                    if (x < 0) x = xck_lookup(xck_arena + vm.xck_off, ent_i + t_start, y);
                    const uint32_t* qr = pool + vm.q_woff;
                    for (int lev = 2; lev <= nins; lev++) {
                        const int bb = lev <= ENT_INS_INLINE ? ent_ins(ec, lev - 1) : base_at(qr, vm.q_s + x + m + lev - 1);
                        vote_generic(lk_key(lev, bb, ((uint32_t)(lev - 1) << 3) | (uint32_t)pb), ord);
                        pb = bb;
                    }
                }
            }
        }
      }
    }
    if (!in_seed) return;
    // ---- gather the direct counters into the table
    if (c_dom) s_dir[1 * VOTE_TP + threadIdx.x] = (o_dom << 16) | c_dom;
#pragma unroll
    for (int d = 0; d < VDIR; d++) {
        const uint32_t v = s_dir[d * VOTE_TP + threadIdx.x];
        if (v == 0u) continue;
        uint32_t k;
        if (d < 6) {
            const int pc = d % 3;
            k = lk_key(0, d < 3 ? Si : 4, pc == 0 ? LK_START : (uint32_t)(pc == 1 ? Sp : 4));
        } else k = lk_key(1, (d - 6) >> 1, (uint32_t)(((d - 6) & 1) ? 4 : Si));
        if (n < VCAP) { key[n] = k; cnt[n] = v; n++; } else overflow = true;
    }
    if (overflow) { atomicMax(err_flag, 1); n = 0; coverage = 0; }
    // ---- write the slot: links sorted by delta (the DP needs a level complete before the next), inside a
    // delta by first appearance
    uint2* slot = slot_arena + (bd.slot_off + (uint64_t)i) * VSLOT;
    uint2* ovf = nullptr; uint32_t ovf_off = 0;
    if (n > VSLOT - 1) {
        ovf_off = atomicAdd(ovf_next, (uint32_t)(n - (VSLOT - 1)));
        if (ovf_off + (uint32_t)(n - (VSLOT - 1)) > ovf_cap) { atomicMax(err_flag, 2); n = 0; coverage = 0; }
        else ovf = ovf_arena + ovf_off;
    }
    slot[0] = make_uint2(((uint32_t)n << 16) | (uint32_t)coverage, ovf_off);
    for (int w = 0; w < n; w++) {                  // selection by (delta, ord): n is ~10
        int best = -1; uint32_t bk = 0xffffffffu;
        for (int e = 0; e < n; e++) {
            const uint32_t sk = (key[e] & 0xffff0000u) | (cnt[e] >> 16);      // delta << 16 | ord
            if (key[e] != 0xffffffffu && sk < bk) { bk = sk; best = e; }      // (two entries of one delta never share ord)
        }
        const uint2 v = make_uint2(key[best], cnt[best] & 0xffffu);
        if (w < VSLOT - 1) slot[1 + w] = v; else ovf[w - (VSLOT - 1)] = v;
        key[best] = 0xffffffffu;                   // emitted
    }
}

// ------------------------------------------------------------------------------ k_cns_dp
// One WARP per seed block, lanes parallel over the LINKS of the current position.  The chain over
// positions is inherent (falcon.c:405-475), so what matters is the latency and the instruction
// count of one link of the chain:
//   * the position slots do not depend on the chain: they are streamed three loads (six positions)
//     ahead through registers, 256 coalesced bytes = two positions per load;
//   * the column (score, record id) pairs of the previous and the current position sit in shared
//     memory (delta levels < CDP_SL; deeper levels, which need an insertion run of >= CDP_SL bases,
//     fall back to a per-block global table);
//   * lane e scores link e; a delta level is closed column by column, LIVE columns only (one
//     warp OR-reduction gives the set): the best link of a column is a warp max-reduction, ties
//     resolved towards the lowest lane = first appearance (the slot keeps links in first-appearance
//     order inside a level);
//   * positions with more than 32 links (deep, noisy pile-ups) take a chunked slow path;
//   * the backtrack walks the record list through a shared-memory window (records are appended in
//     position order, so the predecessor is almost always a few records back).
constexpr int CDP_WARPS = 4;
constexpr int CDP_SL = 12;               // delta levels with shared-memory column tables
constexpr int CDP_LEVELS = 256;          // deltas 0..255 (the tag cut at 255 keeps delta <= 254)
constexpr int CDP_WIN = 256;             // records per backtrack window

struct CdpState {                        // per-block running state of the DP
    uint32_t nrec; int g_best2, g_rec, g_ck, err;
};

__device__ __forceinline__ int2* cdp_tab(int2* s_tab, int2* gtab, const int which, const int idx) {
    return idx < CDP_SL * 5 ? s_tab + which * (CDP_SL * 5) + idx : gtab + which * (CDP_LEVELS * 5) + idx;
}

// Close column (lev, k) of position i: record, column table, global best (falcon.c:420-469).
// All arguments are warp-uniform; lane 0 stores.
__device__ __forceinline__ void cdp_close_column(CdpState& S, const int lane, const int i, const int lev, const int k,
                                                 int col_sc2, int col_pred, int best_ck, const int hi_flag,
                                                 const uint32_t rec_cap, CnsRec* __restrict__ recs,
                                                 int2* s_tab, int2* gtab, const int cur) {
    if (col_sc2 <= -2) { col_sc2 = -2; col_pred = 0; best_ck = -1; }               // floored (falcon.c:447)
    uint32_t ridx;
    if (i == 0 && lev == 0 && k == 0) ridx = 0; else { ridx = S.nrec; S.nrec++; }
    if (ridx >= rec_cap) { S.err = 2; ridx = rec_cap - 1; }
    if (lane == 0) {
        *reinterpret_cast<int4*>(recs + ridx) = make_int4(col_pred, hi_flag | (i << 3) | k, col_sc2, 0);
        *cdp_tab(s_tab, gtab, cur, lev * 5 + k) = make_int2(col_sc2, (int)ridx);
    }
    if (col_sc2 > S.g_best2) { S.g_best2 = col_sc2; S.g_rec = (int)ridx; S.g_ck = best_ck; }
}

// A position with more than 32 links: chunks of 32 links, the per-column state of the open level
// carried across chunks.
__device__ FCX_NOINLINE void cdp_position_slow(CdpState& S, const int lane, const int i, const int n, const int coverage,
                                               const int hi_flag, const uint2* __restrict__ slot, const uint2* __restrict__ ovf,
                                               const uint32_t rec_cap, CnsRec* __restrict__ recs, int2* s_tab, int2* gtab,
                                               const int cur) {
    int best[5], bpred[5], bck[5], nl[5];
    int lev = -1;
#ifdef FCX_EMU
    if (lane == 0 && getenv("FCX_EMU_TRACE_WIDE")) fprintf(stderr, "k_cns_dp: position %d has %d links (chunked path)\n", i, n);
#endif
    auto close_level = [&]() {
#pragma unroll
        for (int k = 0; k < 5; k++)
            if (nl[k] != 0) cdp_close_column(S, lane, i, lev, k, best[k], bpred[k], bck[k], hi_flag, rec_cap, recs, s_tab, gtab, cur);
        __syncwarp();
    };
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        const bool valid = e < n;
        uint2 lk = make_uint2(0u, 0u);
        if (valid) lk = e < VSLOT - 1 ? slot[1 + e] : ovf[e - (VSLOT - 1)];
        const int lev_e = valid ? (int)(lk.x >> 16) : INT_MAX;
        const int kk = (int)((lk.x >> 13) & 7u);
        const uint32_t pred = lk.x & 0x1fffu;
        const int lev_lo = __shfl_sync(FULL, lev_e, 0), lev_hi = __shfl_sync(FULL, lev_e, min(32, n - e0) - 1);
        for (int L = lev_lo; L <= lev_hi; L++) {
            if (L != lev) {
                if (lev >= 0) close_level();
                lev = L;
#pragma unroll
                for (int k = 0; k < 5; k++) { best[k] = INT_MIN; bpred[k] = -1; bck[k] = 0; nl[k] = 0; }
            }
            const bool mine = valid && lev_e == L;
            int s2 = 2 * (int)lk.y - coverage, prj = -1;
            if (mine && pred != LK_START) {
                const int2 v = *cdp_tab(s_tab, gtab, L == 0 ? (cur ^ 1) : cur, (int)(pred >> 3) * 5 + (int)(pred & 7u));
                s2 += v.x; prj = v.y;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const bool in_col = mine && kk == k;
                const unsigned m = __ballot_sync(FULL, in_col);
                if (m == 0u) continue;
                const int cand = in_col ? s2 : INT_MIN;
                const int cb = __reduce_max_sync(FULL, cand);
                if (nl[k] == 0 || cb > best[k]) {                 // strict '>': the first link wins ties
                    const int wl = __ffs(__ballot_sync(FULL, in_col && cand == cb)) - 1;
                    best[k] = cb; bpred[k] = __shfl_sync(FULL, prj, wl);
                    bck[k] = nl[k] + __popc(m & ((1u << wl) - 1u));
                }
                nl[k] += __popc(m);
            }
        }
    }
    if (lev >= 0) close_level();
}

#ifndef CDP_MIN_CTAS
#define CDP_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(CDP_WARPS * 32, CDP_MIN_CTAS)
k_cns_dp(const BlockDesc* __restrict__ blocks, uint32_t n_blocks, const VoteMeta* __restrict__ vmeta,
         const uint2* __restrict__ slot_arena, const uint2* __restrict__ ovf_arena,
         CnsRec* __restrict__ rec_arena, int32_t* __restrict__ lvl_scratch,
         char* __restrict__ cns_arena, int32_t* __restrict__ eqv_arena, int want_eqv, unsigned min_cov,
         CnsOut* __restrict__ out) {
    __shared__ int2 s_tabs[CDP_WARPS][2 * CDP_SL * 5];
    __shared__ int4 s_win[CDP_WARPS][CDP_WIN];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * CDP_WARPS + wib;
    if (b >= n_blocks) return;                    // (no CTA-wide barrier below)
    const BlockDesc bd = blocks[b];
    const int t_len = bd.slen;
    CnsRec* recs = rec_arena + bd.rec_off;
    char* cns = cns_arena + bd.cns_off;
    int32_t* eqv = eqv_arena + bd.cns_off;
    CnsOut co; co.len = 0; co.err = 0; co.start = 0; co.positions = 0;

    // first / last target position carrying tags, and whether anything was accepted (falcon.c:651-656)
    int i_lo = INT_MAX, i_hi = 0, R = 0;
    for (uint32_t j = lane; j < bd.n_pairs; j += 32) {
        const VoteMeta vm = vmeta[bd.pair_begin + j];
        if (vm.t_cnt == 0) continue;
        R++; i_lo = min(i_lo, vm.t_start); i_hi = max(i_hi, vm.t_start + vm.t_cnt);
    }
    R = __reduce_add_sync(FULL, R); i_lo = __reduce_min_sync(FULL, i_lo); i_hi = __reduce_max_sync(FULL, i_hi);
    if (R == 0) { if (lane == 0) { cns[0] = 0; out[b] = co; } return; }
    i_hi = min(i_hi, t_len);
    co.positions = i_hi - i_lo;

    // column (score, record id) of the previous and the current position, indexed delta * 5 + base
    int2* s_tab = s_tabs[wib];
    int2* gtab = reinterpret_cast<int2*>(lvl_scratch + (size_t)b * (4 * CDP_LEVELS * 5));
    int cur = 0;

    // record 0 is reserved for column (0,0,'A'): the target of floored columns' best_p = (0,0,0)
    if (lane == 0) *reinterpret_cast<int4*>(recs) = make_int4(0, 0, -2, 0);
    CdpState S; S.nrec = 1; S.g_best2 = -2; S.g_rec = -1; S.g_ck = 0; S.err = 0;
    const uint2* slots = slot_arena + bd.slot_off * VSLOT;
    // two positions (32 uint2 = 256 bytes) per load: lanes 0-15 position i0, lanes 16-31 position i0 + 1
    auto load2 = [&](const int i0) -> uint2 {
        const int i = i0 + (lane >> 4);
        return i < i_hi ? __ldg(slots + (size_t)i * VSLOT + (lane & 15)) : make_uint2(0u, 0u);
    };
    uint2 v0 = load2(i_lo), v1 = load2(i_lo + 2), v2 = load2(i_lo + 4);

    for (int i0 = i_lo; i0 < i_hi; i0 += 2) {
      const uint2 v = v0; v0 = v1; v1 = v2; v2 = load2(i0 + 6);
#pragma unroll 1
      for (int half = 0; half < 2; half++) {
        const int i = i0 + half;
        if (i >= i_hi) break;
        const int basel = half * 16;
        const uint32_t hx = __shfl_sync(FULL, v.x, basel);
        const int n = (int)(hx >> 16), coverage = (int)(hx & 0xffffu);
        const int hi_flag = ((unsigned)coverage > min_cov) ? (int)0x80000000 : 0;
        if (i == 0 && lane == 0) recs[0].info = hi_flag;
        if (coverage == 0) { cur ^= 1; continue; }
        if (n > 32) {
            const uint32_t hy = __shfl_sync(FULL, v.y, basel);
            cdp_position_slow(S, lane, i, n, coverage, hi_flag, slots + (size_t)i * VSLOT, ovf_arena + hy, bd.rec_cap, recs,
                              s_tab, gtab, cur);
            cur ^= 1;
            continue;
        }
        // links 0..14 of the position sit in the slot; lane e takes link e
        uint32_t kx = __shfl_sync(FULL, v.x, (basel + 1 + lane) & 31), ky = __shfl_sync(FULL, v.y, (basel + 1 + lane) & 31);
        if (n > VSLOT - 1) {
            const uint32_t hy = __shfl_sync(FULL, v.y, basel);
            if (lane >= VSLOT - 1 && lane < n) { const uint2 o = ovf_arena[hy + (uint32_t)(lane - (VSLOT - 1))]; kx = o.x; ky = o.y; }
        }
        const int lev_e = lane < n ? (int)(kx >> 16) : -1;
        const int kk = (int)((kx >> 13) & 7u);
        const uint32_t pred = kx & 0x1fffu;
        const int slotp = (int)(pred >> 3) * 5 + (int)(pred & 7u);
        const int base2 = 2 * (int)ky - coverage;
        const int maxlev = __shfl_sync(FULL, lev_e, n - 1);
        // levels are contiguous: every delta-d tag follows a delta-(d-1) tag of the same read
        const int info0 = hi_flag | (i << 3);
        for (int L = 0; L <= maxlev; L++) {
            const bool mine = lev_e == L;
            int s2 = base2, prj = -1;
            if (mine && pred != LK_START) {
                const int which = L == 0 ? (cur ^ 1) : cur;                               // predecessor columns: position i-1 for delta 0
                const int2 pv = slotp < CDP_SL * 5 ? s_tab[which * (CDP_SL * 5) + slotp] : gtab[which * (CDP_LEVELS * 5) + slotp];
                s2 += pv.x; prj = pv.y;
            }
            unsigned cols = __reduce_or_sync(FULL, mine ? (1u << kk) : 0u);
            const bool reserved0 = i == 0 && L == 0;                                      // column (0,0,'A') owns record 0
            int2* const tab_out = L < CDP_SL ? s_tab + cur * (CDP_SL * 5) + L * 5 : gtab + cur * (CDP_LEVELS * 5) + L * 5;
            while (cols) {                                        // live columns in base order
                const int k = __ffs(cols) - 1;
                cols &= cols - 1u;
                const bool in_col = mine && kk == k;
                const int cand = in_col ? s2 : INT_MIN;
                const int cb = __reduce_max_sync(FULL, cand);     // strict '>' in link order: the FIRST best link wins
                const int wl = __ffs(__ballot_sync(FULL, cand == cb)) - 1;               // (lanes outside the column hold INT_MIN < cb)
                int col_pred = __shfl_sync(FULL, prj, wl), col_sc2 = cb;
                if (cb <= -2) { col_sc2 = -2; col_pred = 0; }                             // floored (falcon.c:447)
                uint32_t ridx = S.nrec;
                if (reserved0 && k == 0) ridx = 0; else S.nrec++;
                if (ridx >= bd.rec_cap) { S.err = 2; ridx = bd.rec_cap - 1; }
                if (lane == 0) {
                    *reinterpret_cast<int4*>(recs + ridx) = make_int4(col_pred, info0 | k, col_sc2, 0);
                    tab_out[k] = make_int2(col_sc2, (int)ridx);
                }
                if (col_sc2 > S.g_best2) {                        // (a floored column never gets here: g_best2 >= -2)
                    S.g_best2 = col_sc2; S.g_rec = (int)ridx;
                    S.g_ck = __popc(__ballot_sync(FULL, in_col) & ((1u << wl) - 1u));     // index of the best link in its column
                }
            }
            __syncwarp();                                         // the next level (or position) reads these columns
        }
        cur ^= 1;
      }
    }
    // ------------------------------------------------------------ backtrack (falcon.c:479-542)
    // The string is produced back to front; it is written from the END of the block's output area
    // towards lower addresses, so no reversal is needed: the consensus starts at cns[start].
    int err = S.err;
    if (S.g_rec < 0) err = 3;                     // reference: assert(g_best_score != -1)
    const int cap = 2 * t_len + 4;
    int pos = cap;                                // one past the last byte written so far
    if (err == 0) {
        int4* win = s_win[wib];
        int wlo = 0, whi = -1;                    // the window holds records wlo..whi
        auto fetch = [&](const int r) -> int4 {
            if (r < wlo || r > whi) {
                __syncwarp();
                whi = r; wlo = max(0, r - (CDP_WIN - 1));
                const int4* src = reinterpret_cast<const int4*>(recs + wlo);
                for (int j = lane; j <= whi - wlo; j += 32) win[j] = src[j];
                __syncwarp();
            }
            return win[r - wlo];
        };
        __syncwarp();                             // lane 0's record stores are visible to every lane
        char bb = '$'; int ck = S.g_ck;           // (a link index >= 5 keeps '$', falcon.c:481)
        unsigned index = 0; const unsigned lim = (unsigned)t_len * 2u;
        int4 r = fetch(S.g_rec);                  // x pred, y info, z score2
        for (;;) {
            // "ACGT-"[ck], lower case below the coverage threshold (falcon.c:497-512)
            if ((unsigned)ck < 5u) bb = (char)(((0x2d54474341ull >> (8 * ck)) & 0xffu) | ((r.y >= 0 && ck < 4) ? 0x20u : 0u));
            if (r.x == -1 || index >= lim) break;
            const int4 pr = fetch(r.x);
            if (bb != '-') {
                pos--;
                if (lane == 0) { cns[pos] = bb; if (want_eqv) eqv[pos] = r.z / 2 - pr.z / 2; }
                index++;
            }
            ck = pr.y & 7;
            r = pr;
        }
        if (lane == 0) cns[cap] = 0;
    }
    co.len = cap - pos; co.start = pos; co.err = err;
    if (lane == 0) out[b] = co;
}

}  // namespace fcx
