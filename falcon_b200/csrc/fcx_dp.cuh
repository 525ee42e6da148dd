// fcx_dp.cuh -- k_dp3: the banded O(ND) forward pass (ref: src/c/DW_banded.c:149-258), one warp per
// pair, diagonals PINNED to lanes and the furthest-reaching values V kept in REGISTERS.
//
// Geometry.  At step d every evaluated diagonal has the parity of d (DW_banded.c:188: k runs from
// min_k in steps of 2 and min_k' = new_min_k - 1 flips parity each step).  Write k = 2m + P with
// P = d & 1.  Diagonal index m is owned by lane (m & 31); a lane keeps the previous step's x of its
// diagonals in two registers selected by the "slot" bit (m >> 5) & 1, so a band of up to 64 cells
// (>= 99.9 % of all steps at 15 % read error: the band tolerance of 150 keeps ~31 cells alive) is
// register resident.  The in-place update of V (DW_banded.c:213) only ever reads the OTHER parity,
// written one step earlier, and with k = 2m + P
//        P = 1:  V[k-1] = prev[m]   (own lane)     V[k+1] = prev[m+1] (lane + 1)
//        P = 0:  V[k-1] = prev[m-1] (lane - 1)     V[k+1] = prev[m]   (own lane)
// so a step needs ONE ring shuffle per 32 cells and no shared memory, no __syncwarp, no per-step
// re-mapping of lanes to diagonals.  Cells are visited in "rounds" of 32 in band order
// (c = band index, round = c >> 5); ballots come out in lane (ring) order and are rotated by
// (lo & 31) into band order, which keeps the trace record format of round 1
// ([min_k, "came from k+1" bit per cell in ascending k]) and therefore k_traceback unchanged.
//
// Bands wider than 64 cells (possible up to 151 = band_size / 2 + 1, DW_banded.c:151,184) are rare
// (low-complexity sequence, long indels just before the band check aborts the pair): the warp
// spills its registers to a shared-memory ring indexed by k and finishes the pair in the generic
// chunked loop ("wide mode").
//
// Early exit (DW_banded.c:220-224): the first k in ascending order whose snake reaches either end
// terminates the alignment; cells after it are never evaluated by the reference, which matters
// only for the evaluated-cell counter reported for parity (PairAln::cells).
#pragma once

namespace fcx {

constexpr int DP3_WARPS = 1;          // one warp per CTA: a finished pair frees its slot at once

__device__ __forceinline__ unsigned rot_band(unsigned ballot, int lo) {
    return __funnelshift_r(ballot, ballot, (unsigned)lo);      // bit c = lane (lo + c) & 31
}

// 16 bases of a packed sequence at position pos through an opaque base pointer: one IMAD.WIDE, two
// loads, one funnel shift (the shift amount wraps mod 32, so 2 * pos needs no masking)
__device__ __forceinline__ uint32_t dp3_fetch(const uint32_t* __restrict__ w, int pos) {
    const uint32_t* a = w + (pos >> 4);
    return __funnelshift_r(__ldg(a), __ldg(a + 1), (unsigned)(pos << 1));
}

// One 16-base compare of a cell.  Coordinates are BIASED: X = qs + x is the position in the packed
// read, Y = X - kk (kk = k - (ts - qs)) the position in the packed seed, qe / te the span ends.
// Returns the advance n <= 16 and, through `rem`, the distance to the nearer end before the
// advance: n == rem means an end was reached (DW_banded.c:220), n == 16 < rem that the snake goes
// on (:203-206).
__device__ __forceinline__ int dp3_snake16(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t,
                                           int qe, int te, int X, int kk, int& rem, unsigned& lim) {
    const int Y = X - kk;
    rem = min(qe - X, te - Y);
    const uint32_t diff = dp3_fetch(q, X) ^ dp3_fetch(t, Y);
    const unsigned run = (unsigned)(__ffs(diff) - 1) >> 1;      // diff == 0 -> 0x7fffffff
    lim = min(16u, (unsigned)rem);                              // n == lim: an end reached, or 16 matches and more to compare
    return (int)min(run, lim);
}

__global__ void __launch_bounds__(DP3_WARPS * 32)
k_dp3(const BlockDesc* __restrict__ blocks, const PairDesc* __restrict__ pairs,
      const PairRange* __restrict__ ranges, const PairAlloc* __restrict__ allocs, uint32_t n_pairs,
      const uint32_t* __restrict__ pool, uint32_t* trace_arena, uint32_t trace_stride, uint32_t* __restrict__ path_arena, double max_diff,
      uint32_t* __restrict__ next_pair, PairAln* __restrict__ out) {
    __shared__ int s_V[DP3_WARPS][VRING];          // wide mode only
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // Persistent warps: the grid fills the GPU once and every warp pulls the next pair from a device
  // counter, so a long pair never holds back a CTA slot and short pairs do not pay a CTA launch.
  for (;;) {
    uint32_t p = 0;
    if (lane == 0) p = atomicAdd(next_pair, 1u);
    p = __shfl_sync(FULL, p, 0);
    if (p >= n_pairs) return;
    PairAln res; res.aligned = res.dist = res.aln_size = res.q_e = res.t_e = res.k_end = 0;
    res.accepted = res.t_cnt = res.n_tags = res.cells = 0;
    const PairRange rg = ranges[p];
    if (!rg.pass) { if (lane == 0) out[p] = res; continue; }
    const PairDesc pd = pairs[p];
    const uint32_t* q = pool + pd.read_woff;
    const uint32_t* t = pool + blocks[pd.block].seed_woff;
    const int qs = rg.s1, ts = rg.s2, q_len = rg.e1 - rg.s1, t_len = rg.e2 - rg.s2;
    const int qe = rg.e1, te = rg.e2, dts = ts - qs;
    int max_d = (int)(0.3 * (q_len + t_len));                  // DW_banded.c:149
#ifndef FCX_EMU
    asm volatile("" : "+r"(max_d));     // keep the FP64 conversion out of the d loop
    asm volatile("" : "+l"(q));         // opaque bases: one IMAD.WIDE per fetch instead of
    asm volatile("" : "+l"(t));         // re-deriving pool + offset for every load
#endif
    const PairAlloc al = allocs[p];
    // The trace of a pair is dead as soon as this warp has walked it backwards (below), so it lives in a
    // scratch area owned by the WARP (trace_stride records), reused pair after pair: the trace memory of a
    // wave is (resident warps) x (longest trace), not the sum over all pairs, and mostly stays in L2.
    uint32_t* trace = trace_arena + (size_t)(blockIdx.x * DP3_WARPS + wib) * trace_stride * TRACE_REC_WORDS;
    const int trace_cap = (int)al.trace_cap;
    const int store_cap = lane == 0 ? trace_cap : 0;           // lane 0 writes the trace records
    const int lane_up = (lane + 1) & 31, lane_dn = (lane + 31) & 31;

    int best = 0, cells = 0;            // best = best_m + qs + ts (biased like X + Y)
    bool aligned = false; int end_d = 0, end_k = 0, end_x = 0, end_y = 0;
    int Va = qs, Vb = qs;               // previous step's X of my diagonals, slot 0 / slot 1
    int lo = 0, ncell = 1;              // band of the coming step: diagonals m = lo .. lo + ncell - 1
    int d = 1;
    // ---- d = 0 peeled: the single cell k = 0 starts at (0,0) (V is calloc'd, DW_banded.c:153,190-192)
    if (max_d > 0) {
        int x = 0, y = 0;
        snake(q, t, qs, ts, q_len, t_len, x, y);               // same on every lane
        if (lane == 0 && trace_cap > 0) { trace[0] = 0u; trace[1] = 1u; }
        cells = 1;
        if (x >= q_len || y >= t_len) { aligned = true; end_x = x; end_y = y; }
        else { Va = Vb = qs + x; best = qs + x + ts + y; lo = -1; ncell = 2; }   // d = 1: k = -1, +1 -> m = -1, 0
    }

    // ------------------------------------------------------------------ register-resident steps
    // One step of parity P with one (TWO = 0: <= 32 cells) or two cells per lane.  The four variants
    // are separate straight-line bodies so that the common one-cell step carries no moves or
    // predicates of the two-cell case.  Returns 0: go on, 1: aligned.
    // A second cell per lane is needed from 33 cells on: with 32 cells every lane owns exactly one
    // diagonal, and the two band edges, whose outer neighbours would alias onto the opposite edge's lane,
    // never read them (min_k takes k+1, max_k takes k-1).
    auto step = [&](auto PC, auto TC) -> int {
        constexpr int P = decltype(PC)::value;
        constexpr bool TWO = decltype(TC)::value != 0;
        const int last = ncell - 1;
        cells += ncell;                                         // (corrected on an early exit)
#ifdef FCX_EMU
        if (lane == 0) { static long hist[80]; static bool reg = false; hist[ncell < 79 ? ncell : 79]++;
            if (!reg && getenv("FCX_EMU_NCELL_HIST")) { reg = true; atexit([] { for (int i = 0; i < 80; i++) if (hist[i]) fprintf(stderr, "ncell %d: %ld\n", i, hist[i]); }); } }
#endif
        const int c0 = (lane - lo) & 31;                        // band index of my first cell
        const int m0 = lo + c0;
        // One-cell steps keep the lane's single value in BOTH slots (see the write-back), so they need no
        // slot selection; a two-cell step that follows finds the value in whichever slot it looks.
        const bool s0 = TWO && (m0 & 32) != 0;
        const int W0 = s0 ? Vb : Va;                            // prev[m0]
        const int W1 = s0 ? Va : Vb;                            // prev[m0 + 32] (TWO only)
        int nb0, nb1 = 0;                                       // the neighbours held by other lanes
        if (!TWO) nb0 = __shfl_sync(FULL, W0, P ? lane_up : lane_dn);
        else if (P) { nb0 = __shfl_sync(FULL, c0 == 0 ? W1 : W0, lane_up); nb1 = __shfl_sync(FULL, W1, lane_up); }
        else { nb0 = __shfl_sync(FULL, W0, lane_dn); nb1 = __shfl_sync(FULL, c0 == 31 ? W0 : W1, lane_dn); }
        // predecessor choice, DW_banded.c:190-197: min_k takes k+1, max_k takes k-1, ties take k-1
        const int vm0 = P ? W0 : nb0, vp0 = P ? nb0 : W0;
        // (the edge rules as value substitutions: X = max(V[k+1], V[k-1] + 1) with the missing side at -inf;
        //  ties vm == vp - 1 ... the reference takes k+1 iff V[k-1] < V[k+1], and then V[k+1] >= V[k-1] + 1)
        const bool act0 = TWO || c0 <= last;
        const int vmx0 = c0 == 0 ? INT_MIN : vm0, vpx0 = c0 == last ? INT_MIN : vp0;
        const bool up0 = vmx0 < vpx0;
        int X0 = max(vpx0, vmx0 + 1);
        int kk0 = 2 * m0 + (P - dts);
        if (!act0) { X0 = qe; kk0 = qe - te; }                  // parked on the span ends: rem = 0, n = 0
        int rem0, rem1 = 0, n1 = 0, X1 = 0, kk1 = 0;
        unsigned lim0, lim1 = 1u;
        bool act1 = false, up1 = false;
        int n0 = dp3_snake16(q, t, qe, te, X0, kk0, rem0, lim0);
        X0 += n0;
        if (TWO) {
            const int vm1 = P ? W1 : nb1, vp1 = P ? nb1 : W1;
            act1 = c0 + 32 <= last;
            const int vpx1 = c0 + 32 == last ? INT_MIN : vp1;
            up1 = vm1 < vpx1;
            X1 = max(vpx1, vm1 + 1);
            kk1 = kk0 + 64;
            if (!act1) { X1 = qe; kk1 = qe - te; }
            n1 = dp3_snake16(q, t, qe, te, X1, kk1, rem1, lim1);
            X1 += n1;
        }
        // rare: a snake longer than 16 bases, or an end reached
        if (__any_sync(FULL, (act0 && (unsigned)n0 == lim0) || (TWO && act1 && (unsigned)n1 == lim1))) {
            bool fin0 = act0 && n0 == rem0, fin1 = TWO && act1 && n1 == rem1;
            bool g0 = act0 && n0 == 16 && !fin0, g1 = TWO && act1 && n1 == 16 && !fin1;
            while (__any_sync(FULL, g0 || g1)) {
                if (g0) { n0 = dp3_snake16(q, t, qe, te, X0, kk0, rem0, lim0); X0 += n0; fin0 = n0 == rem0; g0 = n0 == 16 && !fin0; }
                if (g1) { n1 = dp3_snake16(q, t, qe, te, X1, kk1, rem1, lim1); X1 += n1; fin1 = n1 == rem1; g1 = n1 == 16 && !fin1; }
            }
            const unsigned f0 = rot_band(__ballot_sync(FULL, fin0), lo);
            const unsigned f1 = TWO ? rot_band(__ballot_sync(FULL, fin1), lo) : 0u;
            if (f0 | f1) {                                      // first k in ascending order wins (:220)
                const unsigned e0 = rot_band(__ballot_sync(FULL, up0), lo);
                const unsigned e1 = TWO ? rot_band(__ballot_sync(FULL, up1), lo) : 0u;
                if (d < store_cap) {
                    uint32_t* rec = trace + (size_t)d * TRACE_REC_WORDS;
                    rec[0] = (uint32_t)(2 * lo + P); rec[1] = e0; rec[2] = e1;
                }
                const int fc = f0 ? __ffs(f0) - 1 : 32 + __ffs(f1) - 1;
                const int fX = __shfl_sync(FULL, f0 ? X0 : X1, (lo + fc) & 31);
                aligned = true; end_d = d; end_k = 2 * (lo + fc) + P; end_x = fX - qs; end_y = end_x - end_k;
                cells += fc + 1 - ncell;
                return 1;
            }
        }
        // trace record: bit c = "cell c came from k+1" (bits of cells beyond the band are never read)
        const unsigned upb0 = rot_band(__ballot_sync(FULL, up0), lo);
        if (d < store_cap) {
            uint32_t* rec = trace + (size_t)d * TRACE_REC_WORDS;
            *reinterpret_cast<uint2*>(rec) = make_uint2((uint32_t)(2 * lo + P), upb0);
        }
        if (TWO) {
            const unsigned upb1 = rot_band(__ballot_sync(FULL, up1), lo);
            if (d < store_cap) trace[(size_t)d * TRACE_REC_WORDS + 2] = upb1;
        }
        // write back: the slot of m0 gets X0, the other slot X1 -- or X0 too in a one-cell step
        if (TWO) { if (s0) { Vb = X0; Va = X1; } else { Va = X0; Vb = X1; } }
        else { Va = X0; Vb = X0; }
        // band update, DW_banded.c:227-243.  u = X + Y (biased); parked lanes are excluded
        const int u0 = act0 ? 2 * X0 - kk0 : INT_MIN;
        const int u1 = TWO ? (act1 ? 2 * X1 - kk1 : INT_MIN) : INT_MIN;
        best = max(best, __reduce_max_sync(FULL, TWO ? max(u0, u1) : u0));
        const int thr = best - BAND_TOL;
        const unsigned okb0 = rot_band(__ballot_sync(FULL, u0 >= thr), lo);
        int cmin, cmax;
        if (TWO) {
            const unsigned okb1 = rot_band(__ballot_sync(FULL, u1 >= thr), lo);
            cmin = okb0 ? __ffs(okb0) - 1 : 32 + __ffs(okb1) - 1;
            cmax = okb1 ? 63 - __clz(okb1) : 31 - __clz(okb0);
        } else { cmin = __ffs(okb0) - 1; cmax = 31 - __clz(okb0); }
        lo = lo + cmin - 1 + P;                                 // min_k' = nmin - 1 on the other parity
        ncell = cmax - cmin + 2;                                // max_k' = nmax + 1
        if (TWO && ncell <= 32) {
            // the next step has one cell per lane and reads Va only: every lane keeps the value of the
            // diagonal it owns in the narrower band (in both slots)
            const int mn = lo + ((lane - lo) & 31);
            const int v = (mn & 32) ? Vb : Va;
            Va = v; Vb = v;
        }
        return 0;
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    int status = aligned ? 1 : 0;               // 2: leave (band too wide for this mode, or for the algorithm)
    if (status == 0) {
        for (;;) {                              // d is odd at the top: the parity is static in each half
            if (d >= max_d) break;
            if (ncell > 32) { if (ncell > 64) { status = 2; break; } status = step(I1(), I1()); }   // > 64: wide mode
            else status = step(I1(), I0());                                                          //   (or > 151: abort, :184)
            if (status) break;
            d++;
            if (d >= max_d) break;
            if (ncell > 32) { if (ncell > 64) { status = 2; break; } status = step(I0(), I1()); }
            else status = step(I0(), I0());
            if (status) break;
            d++;
        }
    }
    // ------------------------------------------------------------------ wide mode (rare)
    if (status == 2 && ncell <= BAND_TOL + 1) {                 // max_k - min_k <= band_size (:184)
        int* V = s_V[wib];
#ifdef FCX_EMU
        if (lane == 0 && getenv("FCX_EMU_TRACE_WIDE")) fprintf(stderr, "k_dp3: pair %u enters wide mode at d=%d\n", p, d);
#endif
        // Spill.  The band grows by at most one cell per step, so this mode is entered with exactly
        // 65 cells after a step of exactly 64: the previous band is [lo + 1 - P', lo + 64 - P'] with
        // P' the previous parity, one diagonal per (lane, slot).
        {
            const int Pp = (d - 1) & 1;
            const int lo_prev = lo + 1 - Pp;
            const int c0 = (lane - lo_prev) & 31, m0 = lo_prev + c0;
            const bool s0 = (m0 & 32) != 0;
            V[(2 * m0 + Pp) & (VRING - 1)] = (s0 ? Vb : Va) - qs;             // back to unbiased x
            V[(2 * (m0 + 32) + Pp) & (VRING - 1)] = (s0 ? Va : Vb) - qs;
        }
        __syncwarp();
        int best_m = best - qs - ts;
        int min_k = 2 * lo + (d & 1), max_k = min_k + 2 * (ncell - 1);
        for (; d < max_d && !aligned; d++) {
            if (max_k - min_k > 2 * BAND_TOL) break;            // :184-186
            const int nc = ((max_k - min_k) >> 1) + 1;
            const int nch = (nc + 31) >> 5;
            uint32_t* rec = trace + (size_t)d * TRACE_REC_WORDS;
            const bool rec_ok = d < trace_cap;
            if (lane == 0 && rec_ok) rec[0] = (uint32_t)min_k;
            int step_best = best_m;
            for (int c = 0; c < nch; c++) {
                const int k = min_k + 2 * (lane + 32 * c);
                const bool act = k <= max_k;
                DpCell cc = dp_pick(V, k, min_k, max_k, act);
                if (act) snake(q, t, qs, ts, q_len, t_len, cc.x, cc.y);
                const unsigned upb = __ballot_sync(FULL, cc.up);
                if (lane == 0 && rec_ok) rec[1 + c] = upb;
                const unsigned finb = __ballot_sync(FULL, act && (cc.x >= q_len || cc.y >= t_len));
                step_best = max(step_best, __reduce_max_sync(FULL, act ? cc.x + cc.y : INT_MIN));
                if (finb) {                                     // first k in ascending order wins
                    const int fl = __ffs(finb) - 1;
                    aligned = true; end_d = d; end_k = min_k + 2 * (fl + 32 * c);
                    end_x = __shfl_sync(FULL, cc.x, fl); end_y = __shfl_sync(FULL, cc.y, fl);
                    cells += fl + 1;
                    break;
                }
                cells += min(32, nc - 32 * c);
                if (act) V[k & (VRING - 1)] = cc.x;             // other parity than the entries read: safe
            }
            if (aligned) break;
            best_m = step_best;
            __syncwarp();
            int nmin = INT_MAX, nmax = INT_MIN;
            const int thr = best_m - BAND_TOL;
            for (int c = 0; c < nch; c++) {
                const int k = min_k + 2 * (lane + 32 * c);
                bool ok = false;
                if (k <= max_k) { const int x = V[k & (VRING - 1)]; ok = (2 * x - k) >= thr; }
                const unsigned okb = __ballot_sync(FULL, ok);
                if (okb) {
                    if (nmin == INT_MAX) nmin = min_k + 2 * (__ffs(okb) - 1 + 32 * c);
                    nmax = min_k + 2 * (31 - __clz(okb) + 32 * c);
                }
            }
            max_k = nmax + 1; min_k = nmin - 1;
            __syncwarp();
        }
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
    __syncwarp();                       // wide mode: the ring is reused by the next pair; lane 0's trace records are visible
    // the backward walk of an accepted pair, while its last records are still in L2
    if (res.accepted == 1) { warp_walk_back(trace, path_arena + al.path_off, end_d, end_k, lane, s_V[wib]); __syncwarp(); }
  }
}

}  // namespace fcx
