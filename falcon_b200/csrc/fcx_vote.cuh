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
struct CnsRec { int32_t pred; int32_t info; int32_t score2; };
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
       const uint32_t* __restrict__ xam_arena, const uint32_t* __restrict__ ent_arena,
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

    uint32_t key[VCAP]; uint32_t cnt[VCAP];         // private link table, first-appearance order
    int n = 0, coverage = 0, maxd = 0; bool overflow = false;
    // the dominant link "match after a plain match" is counted in a register; its table slot is
    // reserved when it first appears so that the order stays the reference's
    const uint32_t k_dom = lk_key(0, Si, (uint32_t)Sp);
    int idx_dom = -1; uint32_t c_dom = 0;
    auto vote = [&](const uint32_t k) {
        if (k == k_dom) {
            if (idx_dom < 0) { if (n < VCAP) { idx_dom = n; key[n] = k; cnt[n] = 0; n++; } else overflow = true; }
            c_dom++;
            return;
        }
        for (int e = 0; e < n; e++) if (key[e] == k) { cnt[e]++; return; }
        if (n < VCAP) { key[n] = k; cnt[n] = 1; n++; } else overflow = true;
    };

    for (uint32_t j = 0; j < bd.n_pairs; j++) {
        const VoteMeta vm = vmeta[bd.pair_begin + j];                    // same for the whole CTA
        if (vm.t_cnt == 0) continue;                                     // pair not accepted
        if (vm.t_start >= i0 + VOTE_TP || vm.t_start + vm.t_cnt <= i0) continue;   // does not touch this tile
        const int y = i - vm.t_start;
        const bool act = in_seed && y >= 0 && y < vm.t_cnt;
        const uint32_t* ent = ent_arena + vm.ent_off;
        const uint32_t ec = act ? __ldg(ent + y) : 0u;
        // the read's entry at position i - 1: the left neighbour thread holds it
        uint32_t ep = __shfl_up_sync(FULL, ec, 1);
        if (lane == 0) ep = (act && y > 0) ? __ldg(ent + y - 1) : 0u;
        if (!act) continue;
        coverage++;
        const int m = (ec & ENT_MATCH) ? 1 : 0, nins = ent_nins(ec);
        const int b0 = m ? Si : 4;
        // query index at this column: only needed to fetch inserted bases beyond the 11 inline ones
        int x = -1;
        const uint32_t* qr = pool + vm.q_woff;
        uint32_t pred = LK_START;
        if (y > 0) {
            const int pn = ent_nins(ep);
            int pb;
            if (pn == 0) pb = (ep & ENT_MATCH) ? Sp : 4;
            else if (pn <= ENT_INS_INLINE) pb = ent_ins(ep, pn - 1);
            else { x = xam_lookup(xam_arena + vm.ent_off, ent, y); pb = base_at(qr, vm.q_s + x - 1); }
            pred = ((uint32_t)pn << 3) | (uint32_t)pb;
        }
        vote(lk_key(0, b0, pred));
        if (nins > 0) {
            maxd = max(maxd, nins);
            if (nins > ENT_INS_INLINE && x < 0) x = xam_lookup(xam_arena + vm.ent_off, ent, y);
            int pb = b0;
            for (int lev = 1; lev <= nins; lev++) {
                const int bb = lev <= ENT_INS_INLINE ? ent_ins(ec, lev - 1) : base_at(qr, vm.q_s + x + m + lev - 1);
                vote(lk_key(lev, bb, ((uint32_t)(lev - 1) << 3) | (uint32_t)pb));
                pb = bb;
            }
        }
    }
    if (!in_seed) return;
    if (idx_dom >= 0) cnt[idx_dom] = c_dom;
    if (overflow) { atomicMax(err_flag, 1); n = 0; coverage = 0; }
    // ---- write the slot: links stably sorted by delta (the DP needs a level complete before the next)
    uint2* slot = slot_arena + (bd.slot_off + (uint64_t)i) * VSLOT;
    uint2* ovf = nullptr; uint32_t ovf_off = 0;
    if (n > VSLOT - 1) {
        ovf_off = atomicAdd(ovf_next, (uint32_t)(n - (VSLOT - 1)));
        if (ovf_off + (uint32_t)(n - (VSLOT - 1)) > ovf_cap) { atomicMax(err_flag, 2); n = 0; coverage = 0; }
        else ovf = ovf_arena + ovf_off;
    }
    slot[0] = make_uint2(((uint32_t)n << 16) | (uint32_t)coverage, ovf_off);
    int w = 0;
    for (int lev = 0; lev <= maxd && w < n; lev++)
        for (int e = 0; e < n; e++)
            if ((int)(key[e] >> 16) == lev) {
                const uint2 v = make_uint2(key[e], cnt[e]);
                if (w < VSLOT - 1) slot[1 + w] = v; else ovf[w - (VSLOT - 1)] = v;
                w++;
            }
}

// ------------------------------------------------------------------------------ k_cns_dp
constexpr int CDP_THREADS = 32;
constexpr int CDP_LEVELS = 256;          // deltas 0..255 (the tag cut at 255 keeps delta <= 254)

__global__ void __launch_bounds__(CDP_THREADS)
k_cns_dp(const BlockDesc* __restrict__ blocks, uint32_t n_blocks, const VoteMeta* __restrict__ vmeta,
         const uint2* __restrict__ slot_arena, const uint2* __restrict__ ovf_arena,
         CnsRec* __restrict__ rec_arena, int32_t* __restrict__ lvl_scratch,
         char* __restrict__ cns_arena, int32_t* __restrict__ eqv_arena, int want_eqv, unsigned min_cov,
         CnsOut* __restrict__ out) {
    const uint32_t b = blockIdx.x * CDP_THREADS + threadIdx.x;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    const int t_len = bd.slen;
    CnsRec* recs = rec_arena + bd.rec_off;
    char* cns = cns_arena + bd.cns_off;
    int32_t* eqv = eqv_arena + bd.cns_off;
    CnsOut co; co.len = 0; co.err = 0; co.start = 0; co.positions = 0;

    // first / last target position carrying tags, and whether anything was accepted (falcon.c:651-656)
    int i_lo = INT_MAX, i_hi = 0, R = 0;
    for (uint32_t j = 0; j < bd.n_pairs; j++) {
        const VoteMeta vm = vmeta[bd.pair_begin + j];
        if (vm.t_cnt == 0) continue;
        R++; i_lo = min(i_lo, vm.t_start); i_hi = max(i_hi, vm.t_start + vm.t_cnt);
    }
    if (R == 0) { cns[0] = 0; out[b] = co; return; }
    i_hi = min(i_hi, t_len);
    co.positions = i_hi - i_lo;

    // column scores / record ids of the previous and the current position, indexed delta * 5 + base
    int32_t* tab = lvl_scratch + (size_t)b * (4 * CDP_LEVELS * 5);
    int32_t* sc[2] = { tab, tab + CDP_LEVELS * 5 };
    int32_t* rc[2] = { tab + 2 * CDP_LEVELS * 5, tab + 3 * CDP_LEVELS * 5 };
    int cur = 0;

    // record 0 is reserved for column (0,0,'A'): the target of floored columns' best_p = (0,0,0)
    recs[0].pred = 0; recs[0].info = 0; recs[0].score2 = -2;
    uint32_t nrec = 1;
    int g_best2 = -2, g_rec = -1, g_ck = 0;
    int err = 0;
    const uint2* slots = slot_arena + bd.slot_off * VSLOT;

    for (int i = i_lo; i < i_hi; i++) {
        const uint2* slot = slots + (size_t)i * VSLOT;
        const uint2 hd = slot[0];
        const int n = (int)(hd.x >> 16), coverage = (int)(hd.x & 0xffffu);
        const int hi_flag = ((unsigned)coverage > min_cov) ? (int)0x80000000 : 0;
        if (i == 0) recs[0].info = hi_flag;
        if (coverage == 0) { cur ^= 1; continue; }
        const uint2* ovf = ovf_arena + hd.y;
        int32_t* psc = sc[cur ^ 1]; int32_t* prc = rc[cur ^ 1];
        int32_t* csc = sc[cur]; int32_t* crc = rc[cur];
        int e = 0;
        int lev = 0;
        while (e < n) {
            // ---- one delta level: links of its (up to five) columns, interleaved in first-appearance order
            int best[5], bpred[5], bck[5], nl[5];
#pragma unroll
            for (int k = 0; k < 5; k++) { best[k] = INT_MIN; bpred[k] = -1; bck[k] = 0; nl[k] = 0; }
            const int32_t* ssc = lev == 0 ? psc : csc;        // predecessor columns: position i-1 for delta 0
            const int32_t* src = lev == 0 ? prc : crc;
            for (; e < n; e++) {
                const uint2 lk = e < VSLOT - 1 ? slot[1 + e] : ovf[e - (VSLOT - 1)];
                if ((int)(lk.x >> 16) != lev) break;
                const int kk = (int)((lk.x >> 13) & 7u);
                const uint32_t pred = lk.x & 0x1fffu;
                int s2 = 2 * (int)lk.y - coverage, prj = -1;
                if (pred != LK_START) {
                    const int slotp = (int)(pred >> 3) * 5 + (int)(pred & 7u);
                    s2 += ssc[slotp]; prj = src[slotp];
                }
#pragma unroll
                for (int k = 0; k < 5; k++)
                    if (k == kk) {
                        if (nl[k] == 0 || s2 > best[k]) { best[k] = s2; bpred[k] = prj; bck[k] = nl[k]; }   // strict '>': first wins
                        nl[k]++;
                    }
            }
            // ---- close the level in base order: records, column scores, global best (falcon.c:420-469)
#pragma unroll
            for (int k = 0; k < 5; k++) {
                if (nl[k] == 0) continue;      // dead column: never referenced (a link's predecessor column
                                               // always carries the previous tag of the same read)
                int col_sc2 = best[k], col_pred = bpred[k], best_ck = bck[k];
                if (col_sc2 <= -2) { col_sc2 = -2; col_pred = 0; best_ck = -1; }               // floored (falcon.c:447)
                uint32_t ridx;
                if (i == 0 && lev == 0 && k == 0) ridx = 0; else { ridx = nrec; nrec++; }
                if (ridx >= bd.rec_cap) { err = 2; ridx = bd.rec_cap - 1; }
                recs[ridx].pred = col_pred; recs[ridx].info = hi_flag | (i << 3) | k; recs[ridx].score2 = col_sc2;
                csc[lev * 5 + k] = col_sc2; crc[lev * 5 + k] = (int32_t)ridx;
                if (col_sc2 > g_best2) { g_best2 = col_sc2; g_rec = (int)ridx; g_ck = best_ck; }
            }
            lev++;      // levels are contiguous: every delta-d tag follows a delta-(d-1) tag of the same read
        }
        cur ^= 1;
    }
    // ------------------------------------------------------------ backtrack (falcon.c:479-542)
    // The string is produced back to front; it is written from the END of the block's output area
    // towards lower addresses, so no reversal is needed: the consensus starts at cns[start].
    if (g_rec < 0) err = 3;                       // reference: assert(g_best_score != -1)
    const int cap = 2 * t_len + 4;
    int pos = cap;                                // one past the last byte written so far
    if (err == 0) {
        char bb = '$'; int ck = g_ck; int rcur = g_rec;
        unsigned index = 0; const unsigned lim = (unsigned)t_len * 2u;
        for (;;) {
            const CnsRec r = recs[rcur];
            const bool hi = r.info < 0;
            switch (ck) {
                case 0: bb = hi ? 'A' : 'a'; break;
                case 1: bb = hi ? 'C' : 'c'; break;
                case 2: bb = hi ? 'G' : 'g'; break;
                case 3: bb = hi ? 'T' : 't'; break;
                case 4: bb = '-'; break;
                default: break;
            }
            if (r.pred == -1 || index >= lim) break;
            const CnsRec pr = recs[r.pred];
            if (bb != '-') {
                pos--; cns[pos] = bb;
                if (want_eqv) eqv[pos] = r.score2 / 2 - pr.score2 / 2;
                index++;
            }
            ck = pr.info & 7;
            rcur = r.pred;
        }
        cns[cap] = 0;
    }
    co.len = cap - pos; co.start = pos; co.err = err;
    out[b] = co;
}

}  // namespace fcx
