// fcx_consensus.cuh -- k_consensus: column vote, link-DAG longest path and backtrack of one seed
// block per warp (ref: get_cns_from_align_tags, src/c/falcon.c:308-558).
//
// The longest-path DP is serial in the target position (falcon.c:405-475), so one warp walks the
// seed left to right; parallelism comes from the thousands of seed blocks in a wave.  The pile-up
// of the block is read from the position-major matrix M[i][j] written by k_transpose: one 32-bit
// entry per (seed position i, pair j) = VALID | is_match | n_ins | first 11 inserted bases.
// Rows are software-pipelined through registers (row i+1 is loaded while row i is processed, row
// i-1 is the previous iteration's registers), so the per-position critical path has no global load.
//
// Per position i:
//   vote   every accepted read covering i contributes its delta-0 tag and its insertion tags
//          (falcon.c:350-382).  FAST PATH (all insertion runs <= FJ): votes go to a dense per-warp
//          table indexed by (column, link) holding a count and the first voting read
//          (= first-appearance order of update_col, falcon.c:232-263, because reads are visited
//          in accepted order), updated with shared-memory atomics (add / min).
//          GENERIC PATH (any longer run; ~2 % of positions): explicit link list in
//          first-appearance order, bases fetched from the packed reads through xam[].
//   DP     for delta j = 0..max_delta, base kk = 0..4: best link by strict '>' in first-appearance
//          order; score = pred + count - 0.5*coverage kept as an exact integer (x2); a column whose
//          best stays <= -1 keeps score -1 and best_p = (0,0,0) as in the reference (:420-464);
//   best   global best by strict '>' in (i, j, kk) order, remembering the best LINK INDEX, which
//          the reference then (mis)uses as the first base code of the backtrack (:466-469,494).
// Then the backtrack of falcon.c:479-542 over the stored column records (windowed through shared
// memory) and the in-place reversal.
#pragma once

namespace fcx {

// info = (coverage > min_cov) << 31 | t_pos << 3 | base
struct CnsRec { int32_t pred; int32_t info; int32_t score2; };
struct CnsOut { int32_t len; int32_t err; int32_t deep_positions; int32_t positions;
                long long cyc_vote, cyc_dp, cyc_generic, cyc_backtrack; };

constexpr int CNS_WARPS = 4;
constexpr int CNS_CTAS_PER_SM = 5;   // 20 resident warps / SM: caps registers at 102 per thread
constexpr int LINK_CAP = 512;     // generic path: distinct (delta, base, link) entries per position
constexpr int LVL = 255 * 5;      // (delta, base) slots of one position
constexpr int RCAP = 160;         // accepted reads whose metadata is cached in shared memory
constexpr int NCHR = 8;           // row chunks (of 32 pairs) pipelined through registers
constexpr int FJ = 7;             // fast path: longest insertion run / predecessor delta
constexpr int T0W = 1 + (FJ + 1) * 5;            // links of a delta-0 column: START + (pd, pb)
constexpr int DENSE0 = 2 * T0W;                  // two live delta-0 columns: seed base, '-'
constexpr int DENSE = DENSE0 + FJ * 20;          // + FJ insertion levels x 4 bases x 5 preds
constexpr int DENSE_PAD = 256;
constexpr int SLV = 8 * 5;                       // (delta, base) slots kept in shared memory
constexpr int BTW = 160;                         // backtrack window (records)
constexpr int NOFIRST = 0x7fffffff;              // "no voter yet" in the dense first[] table
static_assert(DENSE <= DENSE_PAD, "dense table too small");
static_assert(T0W <= 64, "level-0 links: at most two per lane");
static_assert((FJ + 1) * 5 <= SLV, "fast-path levels must live in the shared-memory tables");

struct ReadMeta { uint64_t q_woff; uint32_t xam_off_lo, xam_off_hi; int32_t t_start, t_cnt, q_s; int32_t pad; };

struct CnsWarpSmem {
    int32_t cnt[DENSE_PAD];
    int32_t first[DENSE_PAD];
    int32_t sc[2][SLV];
    int32_t rc[2][SLV];
    union {
        struct { uint32_t key[LINK_CAP]; uint16_t kcnt[LINK_CAP]; } g;   // generic path
        int32_t win[BTW * 3];                                             // backtrack window
    } u;
    ReadMeta meta[RCAP];
};

struct LvlTab {      // score / record tables of one position: first 8 levels in smem, rest in global
    int32_t* s_sc; int32_t* s_rc; int32_t* g_sc; int32_t* g_rc;
    __device__ __forceinline__ int32_t sc(int idx) const { return idx < SLV ? s_sc[idx] : g_sc[idx]; }
    __device__ __forceinline__ int32_t rc(int idx) const { return idx < SLV ? s_rc[idx] : g_rc[idx]; }
    __device__ __forceinline__ void set(int idx, int32_t score2, int32_t rec) const {
        if (idx < SLV) { s_sc[idx] = score2; s_rc[idx] = rec; } else { g_sc[idx] = score2; g_rc[idx] = rec; }
    }
};

template <bool PROF>
__global__ void __launch_bounds__(CNS_WARPS * 32, CNS_CTAS_PER_SM)
k_consensus(const BlockDesc* __restrict__ blocks, uint32_t n_blocks, const PairDesc* __restrict__ pairs,
            const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs,
            const PairAln* __restrict__ aln, const uint32_t* __restrict__ pool,
            const uint32_t* __restrict__ xam_arena, const uint32_t* __restrict__ ent_arena,
            const uint32_t* __restrict__ m_arena,
            CnsRec* __restrict__ rec_arena, int32_t* __restrict__ lvl_scratch,
            ReadMeta* __restrict__ meta_scratch, uint64_t meta_stride,
            char* __restrict__ cns_arena, int32_t* __restrict__ eqv_arena, unsigned min_cov,
            CnsOut* __restrict__ out) {
    __shared__ CnsWarpSmem s_all[CNS_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * CNS_WARPS + wib;
    if (b >= n_blocks) return;
    CnsWarpSmem& sm = s_all[wib];
    const BlockDesc bd = blocks[b];
    const uint32_t* seed = pool + bd.seed_woff;
    const int t_len = bd.slen;
    CnsRec* recs = rec_arena + bd.rec_off;
    const size_t gw = (size_t)blockIdx.x * CNS_WARPS + wib;
    int32_t* g_lvl = lvl_scratch + gw * 4 * LVL;
    const LvlTab tabA = { sm.sc[0], sm.rc[0], g_lvl, g_lvl + LVL };
    const LvlTab tabB = { sm.sc[1], sm.rc[1], g_lvl + 2 * LVL, g_lvl + 3 * LVL };
    const unsigned lt = lanemask_lt();

    // ---- gather the accepted reads of this block, in order (falcon.c:629-635)
    ReadMeta* gmeta = meta_scratch + gw * meta_stride;
    int R = 0;
    for (uint32_t j0 = 0; j0 < bd.n_pairs; j0 += 32) {
        uint32_t j = j0 + lane;
        bool ok = false; ReadMeta rm; rm.pad = 0;
        if (j < bd.n_pairs) {
            const uint32_t p = bd.pair_begin + j;
            const PairAln a = aln[p];
            ok = a.accepted > 0;
            if (ok) {
                const PairRange rg = ranges[p];
                const uint64_t xo = allocs[p].xam_off;
                rm.q_woff = pairs[p].read_woff; rm.xam_off_lo = (uint32_t)xo; rm.xam_off_hi = (uint32_t)(xo >> 32);
                rm.t_start = rg.s2; rm.t_cnt = a.t_cnt; rm.q_s = rg.s1;
            }
        }
        unsigned bal = __ballot_sync(FULL, ok);
        if (ok) {
            int slot = R + __popc(bal & lt);
            if (slot < RCAP) sm.meta[slot] = rm;
            gmeta[slot] = rm;
        }
        R += __popc(bal);
    }
    __syncwarp();
    const ReadMeta* meta = (R <= RCAP) ? sm.meta : gmeta;

    CnsOut co; co.len = 0; co.err = 0; co.deep_positions = 0; co.positions = 0;
    co.cyc_vote = co.cyc_dp = co.cyc_generic = co.cyc_backtrack = 0;
    char* cns = cns_arena + bd.cns_off;
    int32_t* eqv = eqv_arena + bd.cns_off;
    if (R == 0) { if (lane == 0) { cns[0] = 0; out[b] = co; } return; }     // falcon.c:651-656

    // first / last target position carrying tags
    int i_lo = INT_MAX, i_hi = 0;
    for (int a0 = 0; a0 < R; a0 += 32) {
        int ai = a0 + lane;
        if (ai < R) { i_lo = min(i_lo, meta[ai].t_start); i_hi = max(i_hi, meta[ai].t_start + meta[ai].t_cnt); }
    }
    i_lo = __reduce_min_sync(FULL, i_lo); i_hi = __reduce_max_sync(FULL, i_hi);
    i_hi = min(i_hi, t_len);

    for (int e = lane; e < DENSE_PAD; e += 32) { sm.cnt[e] = 0; sm.first[e] = NOFIRST; }
    // record 0 is reserved for column (0,0,'A'): the target of floored columns' best_p = (0,0,0).
    // coverage[0] is 0 unless position 0 is processed below.
    if (lane == 0) { recs[0].pred = 0; recs[0].info = 0; recs[0].score2 = -2; }
    __syncwarp();
    uint32_t nrec = 1;
    int g_best2 = -2, g_rec = -1, g_ck = 0, g_t = 0;
    int cur = 0, err = 0;

    const uint32_t* Mb = m_arena + bd.m_off;
    const int RB = (int)bd.rb_pad;
    long long cyc_vote = 0, cyc_dp = 0, cyc_generic = 0, cyc_backtrack = 0; int n_deep = 0;

    uint32_t ecv[NCHR], epv[NCHR], ecn[NCHR];
#pragma unroll
    for (int c = 0; c < NCHR; c++) {
        epv[c] = 0;        // no read covers i_lo - 1
        ecv[c] = (c * 32 < RB && i_lo < i_hi) ? __ldg(Mb + (size_t)i_lo * RB + c * 32 + lane) : 0u;
        ecn[c] = 0;
    }

    for (int i = i_lo; i < i_hi; i++) {
        // prefetch row i+1 (consumed next iteration)
        if (i + 1 < i_hi) {
#pragma unroll
            for (int c = 0; c < NCHR; c++) if (c * 32 < RB) ecn[c] = __ldg(Mb + (size_t)(i + 1) * RB + c * 32 + lane);
        }
        const int Si = base_at(seed, i);
        const int Sp = i > 0 ? base_at(seed, i - 1) : 0;
        const LvlTab& tp = cur ? tabA : tabB;
        const LvlTab& tc = cur ? tabB : tabA;
        int coverage = 0, maxd = 0;
        bool deep = false;
        long long tk0 = 0;
        if (PROF) tk0 = clock64();
        // =============================================================== fast vote
        // Straight-line over the row chunks (no per-chunk collectives on the critical path): every
        // lane decodes its entry, the dominant link "match after match" (col 0, pd 0, pb = seed
        // base) is counted with one ballot per chunk, everything else goes through shared-memory
        // atomics; coverage / max_delta / deep are reduced once at the end.
        {
            const int idx_maj = 1 + Sp;
            int cov_lane = 0, nins_lane = 0, maj_cnt = 0, maj_first = NOFIRST; bool deep_lane = false;
            // one chunk of 32 pairs: decode the entry of this lane's pair and vote
            auto vote_chunk = [&](const uint32_t ec, const uint32_t ep, const int cbase) {
                const bool act = (ec & ENT_VALID) != 0;
                const int ai = cbase + lane;
                const int m = (ec & ENT_MATCH) ? 1 : 0;
                const int nins = act ? ent_nins(ec) : 0;
                const int b0 = m ? Si : 4;
                int l0 = 0; bool dpl = nins > FJ;
                if (ep & ENT_VALID) {
                    const int pn = ent_nins(ep);
                    dpl = dpl || pn > FJ;
                    const int pb = pn > 0 ? ent_ins(ep, (pn - 1) & 7) : ((ep & ENT_MATCH) ? Sp : 4);
                    l0 = 1 + (pn & 7) * 5 + pb;
                }
                const int idx0 = (m ? 0 : T0W) + l0;
                cov_lane += act ? 1 : 0;
                nins_lane = max(nins_lane, nins);
                deep_lane = deep_lane || (act && dpl);
                const bool vote = act && !dpl;
                const unsigned majb = __ballot_sync(FULL, vote && idx0 == idx_maj);
                maj_cnt += __popc(majb);
                if (majb && maj_first == NOFIRST) maj_first = cbase + __ffs(majb) - 1;
                if (vote && idx0 != idx_maj) { atomicAdd(&sm.cnt[idx0], 1); atomicMin(&sm.first[idx0], ai); }
                if (vote && nins >= 1) {                       // insertion tags (few lanes: divergent is fine)
                    int pb = b0;
                    for (int lev = 1; lev <= nins; lev++) {
                        const int bb = ent_ins(ec, lev - 1);
                        const int idx = DENSE0 + (lev - 1) * 20 + bb * 5 + pb;
                        atomicAdd(&sm.cnt[idx], 1); atomicMin(&sm.first[idx], ai);
                        pb = bb;
                    }
                }
            };
#pragma unroll
            for (int c = 0; c < NCHR; c++) {
                if (c * 32 >= RB) continue;
                vote_chunk(ecv[c], epv[c], c * 32);
            }
            // blocks with more than NCHR*32 pairs: the remaining chunks straight from the matrix
            for (int c0 = NCHR * 32; c0 < RB; c0 += 32) {
                const uint32_t* row = Mb + (size_t)i * RB + c0 + lane;
                vote_chunk(__ldg(row), i > 0 ? __ldg(row - RB) : 0u, c0);
            }
            if (lane == 0 && maj_cnt) { sm.cnt[idx_maj] = maj_cnt; sm.first[idx_maj] = maj_first; }
            deep = __ballot_sync(FULL, deep_lane) != 0;
            coverage = __reduce_add_sync(FULL, cov_lane);
            maxd = __reduce_max_sync(FULL, nins_lane);
            __syncwarp();
        }
        if (PROF) { const long long tk1 = clock64(); cyc_vote += tk1 - tk0; tk0 = tk1; }
        if (!deep) {
            const int hi_flag = ((unsigned)coverage > min_cov) ? (int)0x80000000 : 0;
            if (i == 0 && lane == 0) recs[0].info = hi_flag;
            if (coverage != 0) {
                // =========================================================== fast DP
                // all (j <= maxd, kk) slots start dead (score -1, record 0)
                for (int e = lane; e < (maxd + 1) * 5; e += 32) { tc.s_sc[e] = -2; tc.s_rc[e] = 0; }
                // ---- level 0: the two live columns (seed base Si = A, '-' = B) side by side; each
                // lane owns links `lane` and `lane + 32` (T0W = 41 candidate links per column)
                int cA0 = 0, fA0 = NOFIRST, cB0 = 0, fB0 = NOFIRST, ps0 = 0, pr0 = -1;
                int cA1 = 0, fA1 = NOFIRST, cB1 = 0, fB1 = NOFIRST, ps1 = 0, pr1 = -1;
                {
                    cA0 = sm.cnt[lane]; fA0 = sm.first[lane]; cB0 = sm.cnt[T0W + lane]; fB0 = sm.first[T0W + lane];
                    if (cA0) { sm.cnt[lane] = 0; sm.first[lane] = NOFIRST; }
                    if (cB0) { sm.cnt[T0W + lane] = 0; sm.first[T0W + lane] = NOFIRST; }
                    if (lane > 0) { ps0 = tp.s_sc[lane - 1]; pr0 = tp.s_rc[lane - 1]; }
                    const int l1 = lane + 32;
                    if (l1 < T0W) {
                        cA1 = sm.cnt[l1]; fA1 = sm.first[l1]; cB1 = sm.cnt[T0W + l1]; fB1 = sm.first[T0W + l1];
                        if (cA1) { sm.cnt[l1] = 0; sm.first[l1] = NOFIRST; }
                        if (cB1) { sm.cnt[T0W + l1] = 0; sm.first[T0W + l1] = NOFIRST; }
                        ps1 = tp.s_sc[l1 - 1]; pr1 = tp.s_rc[l1 - 1];
                    }
                }
                __syncwarp();
                const int sA0 = cA0 > 0 ? 2 * cA0 - coverage + ps0 : INT_MIN, sA1 = cA1 > 0 ? 2 * cA1 - coverage + ps1 : INT_MIN;
                const int sB0 = cB0 > 0 ? 2 * cB0 - coverage + ps0 : INT_MIN, sB1 = cB1 > 0 ? 2 * cB1 - coverage + ps1 : INT_MIN;
                const int mxA = __reduce_max_sync(FULL, max(sA0, sA1)), mxB = __reduce_max_sync(FULL, max(sB0, sB1));
                const int fcA = min((cA0 > 0 && sA0 == mxA) ? fA0 : NOFIRST, (cA1 > 0 && sA1 == mxA) ? fA1 : NOFIRST);
                const int fcB = min((cB0 > 0 && sB0 == mxB) ? fB0 : NOFIRST, (cB1 > 0 && sB1 == mxB) ? fB1 : NOFIRST);
                const int fmA = __reduce_min_sync(FULL, fcA), fmB = __reduce_min_sync(FULL, fcB);
                const unsigned wbA = __ballot_sync(FULL, fcA == fmA && fmA != NOFIRST);
                const unsigned wbB = __ballot_sync(FULL, fcB == fmB && fmB != NOFIRST);
                const int ckA = __popc(__ballot_sync(FULL, cA0 > 0 && fA0 < fmA)) + __popc(__ballot_sync(FULL, cA1 > 0 && fA1 < fmA));
                const int ckB = __popc(__ballot_sync(FULL, cB0 > 0 && fB0 < fmB)) + __popc(__ballot_sync(FULL, cB1 > 0 && fB1 < fmB));
                // the winning lane hands over the predecessor record of its winning link
                const int prA_l = (cA0 > 0 && sA0 == mxA && fA0 == fmA) ? pr0 : pr1;
                const int prB_l = (cB0 > 0 && sB0 == mxB && fB0 == fmB) ? pr0 : pr1;
                const int prA = __shfl_sync(FULL, prA_l, wbA ? __ffs(wbA) - 1 : 0);
                const int prB = __shfl_sync(FULL, prB_l, wbB ? __ffs(wbB) - 1 : 0);
#pragma unroll
                for (int col = 0; col < 2; col++) {
                    const int mx = col ? mxB : mxA;
                    if (mx == INT_MIN) continue;                                 // dead column
                    const int kk = col ? 4 : Si;
                    int col_sc2 = mx, col_pred = col ? prB : prA, best_ck = col ? ckB : ckA;
                    if (mx <= -2) { col_sc2 = -2; col_pred = 0; best_ck = -1; }       // floored (falcon.c:447)
                    uint32_t ridx;
                    if (i == 0 && kk == 0) ridx = 0; else { ridx = nrec; nrec++; }
                    if (ridx >= bd.rec_cap) { err = 2; ridx = bd.rec_cap - 1; }
                    if (lane == 0) {
                        recs[ridx].pred = col_pred; recs[ridx].info = hi_flag | (i << 3) | kk; recs[ridx].score2 = col_sc2;
                        tc.s_sc[kk] = col_sc2; tc.s_rc[kk] = (int32_t)ridx;
                    }
                    if (col_sc2 > g_best2) { g_best2 = col_sc2; g_rec = (int)ridx; g_ck = best_ck; g_t = i; }
                }
                __syncwarp();
                // ---- insertion levels: 4 columns x 5 candidate links, lane = base*8 + pred base
                for (int j = 1; j <= maxd; j++) {
                    const int bb = lane >> 3, pb = lane & 7;
                    const int e = DENSE0 + (j - 1) * 20 + bb * 5 + pb;
                    int cntv = 0, fst = NOFIRST;
                    if (pb < 5) { cntv = sm.cnt[e]; fst = sm.first[e]; if (cntv) { sm.cnt[e] = 0; sm.first[e] = NOFIRST; } }
                    const bool valid = cntv > 0;
                    int s2 = INT_MIN, prj = -1;
                    if (valid) { s2 = 2 * cntv - coverage + tc.s_sc[(j - 1) * 5 + pb]; prj = tc.s_rc[(j - 1) * 5 + pb]; }
                    const unsigned vb = __ballot_sync(FULL, valid);
                    if (!vb) continue;
                    // one keyed butterfly over the 8-lane group: highest score, then earliest voter
                    long long kv = valid ? (long long)s2 * 4294967296ll + (long long)(0x7fffffff - fst) : LLONG_MIN;
                    long long km = kv;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) { const long long ov = __shfl_xor_sync(FULL, km, o); km = ov > km ? ov : km; }
                    const int mx = valid || km != LLONG_MIN ? (int)(km >> 32) : INT_MIN;
                    const int fm = 0x7fffffff - (int)(km & 0x7fffffffll);
                    const unsigned wb = __ballot_sync(FULL, valid && kv == km);
                    const unsigned eb = __ballot_sync(FULL, valid && fst < fm && km != LLONG_MIN);       // links before the winner
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const unsigned gm = 0xffu << (8 * g);
                        if (!(vb & gm)) continue;                                   // dead column
                        const int wl = __ffs(wb & gm) - 1;
                        int col_sc2 = __shfl_sync(FULL, mx, wl), col_pred = __shfl_sync(FULL, prj, wl);
                        int best_ck = __popc(eb & gm);
                        if (col_sc2 <= -2) { col_sc2 = -2; col_pred = 0; best_ck = -1; }
                        uint32_t ridx = nrec; nrec++;
                        if (ridx >= bd.rec_cap) { err = 2; ridx = bd.rec_cap - 1; }
                        if (lane == 0) {
                            recs[ridx].pred = col_pred; recs[ridx].info = hi_flag | (i << 3) | g; recs[ridx].score2 = col_sc2;
                            tc.s_sc[j * 5 + g] = col_sc2; tc.s_rc[j * 5 + g] = (int32_t)ridx;
                        }
                        if (col_sc2 > g_best2) { g_best2 = col_sc2; g_rec = (int)ridx; g_ck = best_ck; g_t = i; }
                    }
                    __syncwarp();
                }
            }
            if (PROF) cyc_dp += clock64() - tk0;
        } else {
            // =========================================================== generic path (rare)
            n_deep++;
            for (int e = lane; e < DENSE_PAD; e += 32) { sm.cnt[e] = 0; sm.first[e] = NOFIRST; }
            __syncwarp();
            uint32_t* key = sm.u.g.key; uint16_t* kcnt = sm.u.g.kcnt;
            int nlink = 0; coverage = 0; maxd = 0;
            for (int c0 = 0; c0 < R; c0 += 32) {
                const int ai = c0 + lane;
                bool act = false; int y = 0;
                ReadMeta rm;
                if (ai < R) { rm = meta[ai]; y = i - rm.t_start; act = y >= 0 && y < rm.t_cnt; }
                const unsigned actb = __ballot_sync(FULL, act);
                if (!actb) continue;
                coverage += __popc(actb);
                int m = 0, x = 0, nins = 0, b0 = 0; uint32_t lk0 = 0;
                const uint32_t* qr = pool; int qs = 0;
                if (act) {
                    const uint64_t xo = ((uint64_t)rm.xam_off_hi << 32) | rm.xam_off_lo;
                    const uint32_t* xam = xam_arena + xo;
                    const uint32_t* ent = ent_arena + xo;
                    const uint32_t e = ent[y];
                    m = (e & ENT_MATCH) ? 1 : 0; nins = ent_nins(e);
                    x = xam_lookup(xam, ent, y);
                    b0 = m ? Si : 4;
                    qr = pool + rm.q_woff; qs = rm.q_s;
                    if (y == 0) lk0 = 0x1fffu;                         // (p_t_pos = -1, 0, '.')
                    else {
                        const uint32_t pe = ent[y - 1];
                        const int pm = (pe & ENT_MATCH) ? 1 : 0, pn = ent_nins(pe);
                        const int pb = pn > 0 ? base_at(qr, qs + x - 1) : (pm ? Sp : 4);
                        lk0 = ((uint32_t)pn << 3) | (uint32_t)pb;
                    }
                }
                const int lmax = __reduce_max_sync(FULL, act ? nins : 0);
                maxd = max(maxd, lmax);
                // levels of this chunk merged in (level, lane) order; inside one column the first
                // appearance of a link is decided by the read index alone, so this equals read order
                for (int lev = 0; lev <= lmax; lev++) {
                    const bool has = act && nins >= lev;
                    uint32_t k = 0xffffffffu - lane;
                    if (has) {
                        if (lev == 0) k = ((uint32_t)b0 << 13) | lk0;
                        else {
                            const int bb = base_at(qr, qs + x + m + lev - 1);
                            const int pb = (lev == 1) ? b0 : base_at(qr, qs + x + m + lev - 2);
                            k = ((uint32_t)lev << 16) | ((uint32_t)bb << 13) | ((uint32_t)(lev - 1) << 3) | (uint32_t)pb;
                        }
                    }
                    const unsigned peers = __match_any_sync(FULL, k);
                    const bool leader = has && (peers & lt) == 0;
                    const int pc = __popc(peers);
                    unsigned lead = __ballot_sync(FULL, leader);
                    while (lead) {
                        const int l = __ffs(lead) - 1; lead &= lead - 1;
                        const uint32_t kk = __shfl_sync(FULL, k, l); const int cc = __shfl_sync(FULL, pc, l);
                        int found = -1;
                        for (int e0 = 0; e0 < nlink; e0 += 32) {
                            const unsigned hb = __ballot_sync(FULL, (e0 + lane < nlink) && key[e0 + lane] == kk);
                            if (hb) { found = e0 + __ffs(hb) - 1; break; }
                        }
                        if (found >= 0) { if (lane == 0) kcnt[found] = (uint16_t)(kcnt[found] + cc); }
                        else if (nlink < LINK_CAP) { if (lane == 0) { key[nlink] = kk; kcnt[nlink] = (uint16_t)cc; } nlink++; }
                        else err = 1;
                        __syncwarp();
                    }
                }
            }
            const int hi_flag = ((unsigned)coverage > min_cov) ? (int)0x80000000 : 0;
            if (i == 0 && lane == 0) recs[0].info = hi_flag;
            for (int j = 0; j <= maxd && coverage != 0; j++) {
                for (int kk = 0; kk < 5; kk++) {
                    const uint32_t want = ((uint32_t)j << 3) | (uint32_t)kk;       // key >> 13
                    bool any = false; int seen = 0;
                    long long bkey = LLONG_MIN; int brec = -1, bsc2 = 0;
                    for (int e0 = 0; e0 < nlink; e0 += 32) {
                        const int e = e0 + lane;
                        const bool mine = e < nlink && (key[e] >> 13) == want;
                        const unsigned mb = __ballot_sync(FULL, mine);
                        if (!mb) continue;
                        any = true;
                        long long v = LLONG_MIN; int prj = -1, s2v = 0;
                        if (mine) {
                            const uint32_t lk = key[e] & 0x1fffu;
                            int s2 = 2 * (int)kcnt[e] - coverage;
                            if (lk != 0x1fffu) {
                                const int slot = (int)(lk >> 3) * 5 + (int)(lk & 7);
                                const LvlTab& src = (j == 0) ? tp : tc;
                                s2 += src.sc(slot); prj = src.rc(slot);
                            }
                            s2v = s2;
                            v = (long long)s2 * (1ll << 20) + (long long)(0xfffff - (seen + __popc(mb & lt)));
                        }
                        long long wm = v;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) { const long long ov = __shfl_xor_sync(FULL, wm, o); wm = ov > wm ? ov : wm; }
                        if (wm > bkey) {
                            const unsigned who = __ballot_sync(FULL, mine && v == wm);
                            const int wl = __ffs(who) - 1;
                            bkey = wm; brec = __shfl_sync(FULL, prj, wl); bsc2 = __shfl_sync(FULL, s2v, wl);
                        }
                        seen += __popc(mb);
                    }
                    if (!any) { if (lane == 0) tc.set(j * 5 + kk, -2, 0); continue; }       // dead column
                    int best_ck = 0xfffff - (int)(bkey & 0xfffff);
                    int col_sc2, col_pred;
                    if (bsc2 > -2) { col_sc2 = bsc2; col_pred = brec; }
                    else { col_sc2 = -2; col_pred = 0; best_ck = -1; }
                    uint32_t ridx;
                    if (i == 0 && j == 0 && kk == 0) ridx = 0; else { ridx = nrec; nrec++; }
                    if (ridx >= bd.rec_cap) { err = 2; ridx = bd.rec_cap - 1; }
                    if (lane == 0) {
                        recs[ridx].pred = col_pred; recs[ridx].info = hi_flag | (i << 3) | kk; recs[ridx].score2 = col_sc2;
                        tc.set(j * 5 + kk, col_sc2, (int32_t)ridx);
                    }
                    if (col_sc2 > g_best2) { g_best2 = col_sc2; g_rec = (int)ridx; g_ck = best_ck; g_t = i; }
                }
                __syncwarp();
            }
            __syncwarp();
            if (PROF) cyc_generic += clock64() - tk0;
        }
        cur ^= 1;
#pragma unroll
        for (int c = 0; c < NCHR; c++) { epv[c] = ecv[c]; ecv[c] = ecn[c]; }
    }
    // ------------------------------------------------------------ backtrack (falcon.c:479-542)
    // Records are numbered in DP order, so a predecessor always has a smaller index: the walk moves
    // down through a window of BTW records staged in shared memory (coalesced loads), reloading
    // when it leaves the window; a far jump (floored column -> record 0) falls back to global.
    long long tb0 = 0;
    if (PROF) tb0 = clock64();
    if (g_rec < 0) err = 3;                       // reference: assert(g_best_score != -1)
    __syncwarp();
    int index = 0;
    if (err == 0) {
        char bb = '$'; int ck = g_ck; int rc = g_rec; bool done = false;
        const unsigned lim = (unsigned)t_len * 2u;
        int32_t* win = sm.u.win;
        while (!done) {
            const int lo = max(0, rc - (BTW - 1));
            const int* src = reinterpret_cast<const int*>(recs + lo);
            const int nint = min(BTW, (int)nrec - lo) * 3;
            for (int e = lane; e < nint; e += 32) win[e] = src[e];
            __syncwarp();
            if (lane == 0) {
                for (;;) {
                    const int32_t* r = win + (rc - lo) * 3;      // rc is always inside the window here
                    const bool hi = r[1] < 0;
                    switch (ck) {
                        case 0: bb = hi ? 'A' : 'a'; break;
                        case 1: bb = hi ? 'C' : 'c'; break;
                        case 2: bb = hi ? 'G' : 'g'; break;
                        case 3: bb = hi ? 'T' : 't'; break;
                        case 4: bb = '-'; break;
                        default: break;
                    }
                    const int pred = r[0];
                    if (pred == -1 || (unsigned)index >= lim) { done = true; break; }
                    int pinfo, pscore;
                    if (pred >= lo) { pinfo = win[(pred - lo) * 3 + 1]; pscore = win[(pred - lo) * 3 + 2]; }
                    else { pinfo = recs[pred].info; pscore = recs[pred].score2; }
                    if (bb != '-') { cns[index] = bb; eqv[index] = r[2] / 2 - pscore / 2; index++; }
                    ck = pinfo & 7;
                    const bool leave = pred < lo + 8 && lo > 0;
                    rc = pred;
                    if (pred < lo || leave) break;               // restage the window around the new rc
                }
            }
            done = __shfl_sync(FULL, (int)done, 0) != 0;
            rc = __shfl_sync(FULL, rc, 0);
            __syncwarp();
        }
    }
    index = __shfl_sync(FULL, index, 0);
    __syncwarp();
    for (int a = lane; a < index / 2; a += 32) {       // reverse in place (falcon.c:533-540)
        const char tc2 = cns[a]; cns[a] = cns[index - 1 - a]; cns[index - 1 - a] = tc2;
        const int te = eqv[a]; eqv[a] = eqv[index - 1 - a]; eqv[index - 1 - a] = te;
    }
    if (PROF) cyc_backtrack = clock64() - tb0;
    if (lane == 0) {
        cns[index] = 0; co.len = index; co.err = err; co.deep_positions = n_deep; co.positions = i_hi - i_lo;
        co.cyc_vote = cyc_vote; co.cyc_dp = cyc_dp; co.cyc_generic = cyc_generic; co.cyc_backtrack = cyc_backtrack;
        out[b] = co;
    }
}

}  // namespace fcx
