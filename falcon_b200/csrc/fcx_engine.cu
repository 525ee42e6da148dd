// fcx_engine.cu -- host side of libfalcon_b200.so: device memory, wave planning, launches and
// the batched C ABI declared in include/falcon_b200.h.  There is NO CPU path for the arithmetic:
// every stage runs in the CUDA kernels of fcx_kernels.cuh and any CUDA failure is reported (fcx_*)
// or fatal (legacy symbols), never papered over.
#include "fcx_kernels.cuh"
#include "../../include/falcon_b200.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace fcx;

namespace {

struct DevBuf {               // grow-only device buffer
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct HostBuf {              // grow-only pinned host buffer
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct fcx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // pool
    uint32_t n_reads = 0;
    std::vector<uint64_t> h_woff;      // n_reads + 1
    std::vector<int32_t> h_len;
    DevBuf d_pool, d_ascii, d_aoff, d_woff, d_len, d_dirty;
    // wave buffers
    DevBuf d_blocks, d_pairs, d_ranges, d_allocs, d_aln, d_ktab, d_kpos, d_trace, d_path, d_xam, d_ent, d_M,
           d_recs, d_cov, d_lvl, d_acc, d_cns, d_eqv, d_cnsout;
    DevBuf d_rlist;
    int sm_count = 148;
    bool profile = false;
    double prof[8] = {0};
    HostBuf h_ranges, h_aln, h_cns, h_cnsout, h_stage, h_eqv;
    // results
    std::vector<char> out_bases;
    std::vector<uint64_t> out_off;
    std::vector<fcx_pair_info> pair_info;
    std::vector<int32_t> out_eqv;
    bool want_eqv = false;
    bool keep_pair_info = true;
    double times[FCX_T_COUNT] = {0};
    uint64_t counters[FCX_C_COUNT] = {0};
    cudaEvent_t ev[8] = {nullptr};
    cudaEvent_t tev[2] = {nullptr, nullptr};
    size_t arena_budget = (size_t)110 << 30;
    uint32_t max_wave_blocks = 4096;
    uint32_t max_wave_pairs = 1u << 19;
};

static thread_local std::string g_create_err;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            char buf_[512];                                                               \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call,                   \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                         \
            ctx->err = buf_;                                                              \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

extern "C" const char* fcx_version(void) { return "falcon_b200 0.1 sm_100a"; }

extern "C" const char* fcx_last_error(const fcx_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int fcx_create(int device, fcx_ctx** out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                       " (falcon_b200 has no CPU path)";
        return 1;
    }
    if (device < 0 || device >= n) { g_create_err = "bad device ordinal"; return 1; }
    fcx_ctx* ctx = new fcx_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess) {
        g_create_err = "cudaSetDevice/cudaStreamCreate failed";
        delete ctx; return 1;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    for (auto& ev : ctx->tev) cudaEventCreate(&ev);
    if (const char* s = getenv("FCX_ARENA_GB")) ctx->arena_budget = (size_t)atof(s) * ((size_t)1 << 30);
    if (const char* s = getenv("FCX_WAVE_BLOCKS")) ctx->max_wave_blocks = (uint32_t)atoi(s);
    if (const char* s = getenv("FCX_WAVE_PAIRS")) ctx->max_wave_pairs = (uint32_t)atoi(s);
    if (const char* s = getenv("FCX_PROFILE")) ctx->profile = atoi(s) != 0;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    // dynamic shared memory opt-in for k_range
    cudaFuncSetAttribute(k_range, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         RANGE_WARPS * RANGE_BINS * (int)sizeof(int));
    *out = ctx;
    return 0;
}

extern "C" void fcx_destroy(fcx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_pool, &ctx->d_ascii, &ctx->d_aoff, &ctx->d_woff, &ctx->d_len, &ctx->d_dirty,
                      &ctx->d_blocks, &ctx->d_pairs, &ctx->d_ranges, &ctx->d_allocs, &ctx->d_aln,
                      &ctx->d_ktab, &ctx->d_kpos, &ctx->d_trace, &ctx->d_path, &ctx->d_xam, &ctx->d_ent, &ctx->d_M, &ctx->d_rlist, &ctx->d_recs,
                      &ctx->d_cov, &ctx->d_lvl, &ctx->d_acc, &ctx->d_cns, &ctx->d_eqv, &ctx->d_cnsout};
    for (auto* b : bufs) b->release();
    HostBuf* hb[] = {&ctx->h_ranges, &ctx->h_aln, &ctx->h_cns, &ctx->h_cnsout, &ctx->h_stage, &ctx->h_eqv};
    for (auto* b : hb) b->release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->tev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void* fcx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void fcx_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------- pool
extern "C" int fcx_pool_upload(fcx_ctx* ctx, const char* bases, const uint64_t* offsets, uint32_t n_reads) {
    CK(cudaSetDevice(ctx->device));
    ctx->n_reads = 0;
    ctx->h_woff.assign((size_t)n_reads + 1, 0);
    ctx->h_len.assign(n_reads, 0);
    uint64_t w = 0;
    for (uint32_t r = 0; r < n_reads; r++) {
        uint64_t len = offsets[r + 1] - offsets[r];
        if (len >= 100000) { ctx->err = "read longer than 99999 bases (the reference truncates at consensus.py:178-179 and asserts at falcon.c:343)"; return 1; }
        ctx->h_len[r] = (int32_t)len;
        ctx->h_woff[r] = w;
        uint64_t words = (len + 15) / 16 + 1;            // +1 zero pad word: fetch16 reads one word ahead
        w += (words + 3) & ~(uint64_t)3;                 // 16-byte aligned starts
    }
    ctx->h_woff[n_reads] = w;
    const uint64_t total_bytes = offsets[n_reads] - offsets[0];
    CK(ctx->d_pool.reserve((w + 4) * 4));
    CK(ctx->d_ascii.reserve(total_bytes + 16));
    CK(ctx->d_aoff.reserve(((size_t)n_reads + 1) * 8));
    CK(ctx->d_woff.reserve(((size_t)n_reads + 1) * 8));
    CK(ctx->d_len.reserve((size_t)n_reads * 4 + 4));
    CK(ctx->d_dirty.reserve(4));
    std::vector<uint64_t> rel((size_t)n_reads + 1);
    for (uint32_t r = 0; r <= n_reads; r++) rel[r] = offsets[r] - offsets[0];
    CK(cudaMemcpyAsync(ctx->d_ascii.p, bases + offsets[0], total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_aoff.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_woff.p, ctx->h_woff.data(), ctx->h_woff.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_len.p, ctx->h_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_dirty.p, 0, 4, ctx->stream));
    CK(cudaMemsetAsync((char*)ctx->d_pool.p + w * 4, 0, 16, ctx->stream));
    if (w > 0) {
        uint64_t nb = (w + 255) / 256;
        k_pack<<<(unsigned)nb, 256, 0, ctx->stream>>>(ctx->d_ascii.as<uint8_t>(), ctx->d_aoff.as<uint64_t>(),
                                                      ctx->d_woff.as<uint64_t>(), ctx->d_len.as<int32_t>(), n_reads, w,
                                                      ctx->d_pool.as<uint32_t>(), ctx->d_dirty.as<int>());
        CK(cudaGetLastError());
    }
    int dirty = 0;
    CK(cudaMemcpyAsync(&dirty, ctx->d_dirty.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (dirty) { ctx->err = "read pool contains bytes outside upper-case ACGT (reference behaviour undefined: falcon.c:370-379)"; return 2; }
    ctx->n_reads = n_reads;
    return 0;
}

// ---------------------------------------------------------------------------------- waves
namespace {

struct WavePlan { uint32_t b0, b1; };

inline uint64_t max_d_of(int q_len, int t_len) { return (uint64_t)(int)(0.3 * (q_len + t_len)); }

int run_wave(fcx_ctx* ctx, uint32_t b0, uint32_t b1, const uint32_t* block_off, const uint32_t* read_ids,
             unsigned min_cov, double min_idt, uint64_t pair_base) {
    const uint32_t nb = b1 - b0;
    std::vector<BlockDesc> hb(nb);
    uint64_t npairs64 = 0, kpos_total = 0, rec_total = 0, cns_total = 0, cov_total = 0, m_total = 0, tiles = 0;
    uint32_t max_np = 1;
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t lo = block_off[b0 + b], hi = block_off[b0 + b + 1];
        BlockDesc& d = hb[b];
        uint32_t seed = read_ids[lo];
        d.seed_woff = ctx->h_woff[seed];
        d.slen = ctx->h_len[seed];
        d.pair_begin = (uint32_t)npairs64;
        d.n_pairs = hi - lo - 1;
        d.kpos_off = kpos_total; kpos_total += (uint64_t)std::max(d.slen, 1);
        d.rec_cap = (uint32_t)std::max<int64_t>(64, (int64_t)d.slen * 8 + 64);
        d.rec_off = rec_total; rec_total += d.rec_cap;
        d.cns_off = cns_total; cns_total += (uint64_t)d.slen * 2 + 8;
        d.cov_off = cov_total; cov_total += (uint64_t)d.slen + 8;
        d.rb_pad = (d.n_pairs + 31u) & ~31u;
        d.m_off = m_total; m_total += (uint64_t)d.rb_pad * (uint64_t)std::max(d.slen, 1);
        d.tile_begin = (uint32_t)tiles; tiles += (uint64_t)((d.slen + 31) / 32) * (d.rb_pad / 32);
        npairs64 += d.n_pairs;
        max_np = std::max(max_np, d.n_pairs);
    }
    const uint32_t np = (uint32_t)npairs64;
    std::vector<PairDesc> hp(np);
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t lo = block_off[b0 + b];
        for (uint32_t j = 0; j < hb[b].n_pairs; j++) {
            uint32_t rid = read_ids[lo + 1 + j];
            PairDesc& pd = hp[hb[b].pair_begin + j];
            pd.read_woff = ctx->h_woff[rid]; pd.block = b; pd.rlen = ctx->h_len[rid];
        }
    }
    cudaStream_t st = ctx->stream;
    CK(ctx->d_blocks.reserve(nb * sizeof(BlockDesc)));
    CK(ctx->d_pairs.reserve((size_t)std::max(np, 1u) * sizeof(PairDesc)));
    CK(ctx->d_ranges.reserve((size_t)std::max(np, 1u) * sizeof(PairRange)));
    CK(ctx->d_allocs.reserve((size_t)std::max(np, 1u) * sizeof(PairAlloc)));
    CK(ctx->d_aln.reserve((size_t)std::max(np, 1u) * sizeof(PairAln)));
    CK(ctx->d_ktab.reserve((size_t)nb * KTAB * 4));
    CK(ctx->d_kpos.reserve(kpos_total * 4 + 16));
    CK(ctx->d_recs.reserve(rec_total * sizeof(CnsRec)));
    CK(ctx->d_cns.reserve(cns_total));
    CK(ctx->d_eqv.reserve(cns_total * 4));
    CK(ctx->d_cnsout.reserve(nb * sizeof(CnsOut)));
    const uint32_t cns_grid = (nb + CNS_WARPS - 1) / CNS_WARPS;
    CK(ctx->d_lvl.reserve((size_t)cns_grid * CNS_WARPS * 4 * LVL * 4));
    CK(ctx->d_acc.reserve((size_t)cns_grid * CNS_WARPS * max_np * sizeof(ReadMeta)));
    CK(ctx->h_ranges.reserve((size_t)std::max(np, 1u) * sizeof(PairRange)));
    CK(ctx->h_aln.reserve((size_t)std::max(np, 1u) * sizeof(PairAln)));
    CK(ctx->h_cns.reserve(cns_total));
    CK(ctx->h_cnsout.reserve(nb * sizeof(CnsOut)));

    CK(cudaMemcpyAsync(ctx->d_blocks.p, hb.data(), nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    if (np) CK(cudaMemcpyAsync(ctx->d_pairs.p, hp.data(), (size_t)np * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
    const uint32_t* pool = ctx->d_pool.as<uint32_t>();

    // ---- index
    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaMemsetAsync(ctx->d_ktab.p, 0, (size_t)nb * KTAB * 4, st));
    k_index<<<nb, 256, 0, st>>>(ctx->d_blocks.as<BlockDesc>(), pool, ctx->d_ktab.as<uint32_t>(), ctx->d_kpos.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], st));
    ctx->counters[FCX_C_KERNEL_LAUNCHES] += 1;
    // ---- range
    if (np) {
        const unsigned rgrid = std::min<unsigned>((np + RANGE_WARPS - 1) / RANGE_WARPS, (unsigned)ctx->sm_count * 3u);
        CK(ctx->d_rlist.reserve((size_t)rgrid * RANGE_WARPS * RANGE_LIST_CAP * sizeof(int2)));
        k_range<<<rgrid, RANGE_WARPS * 32, RANGE_WARPS * RANGE_BINS * sizeof(int), st>>>(
            ctx->d_blocks.as<BlockDesc>(), ctx->d_pairs.as<PairDesc>(), np, pool, ctx->d_ktab.as<uint32_t>(),
            ctx->d_kpos.as<uint32_t>(), ctx->d_rlist.as<int2>(), ctx->d_ranges.as<PairRange>());
        CK(cudaGetLastError());
        ctx->counters[FCX_C_KERNEL_LAUNCHES] += 1;
        CK(cudaMemcpyAsync(ctx->h_ranges.p, ctx->d_ranges.p, (size_t)np * sizeof(PairRange), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    CK(cudaStreamSynchronize(st));
    // ---- exact per-pair allocations
    std::vector<PairAlloc> ha(np);
    uint64_t trace_recs = 0, xam_n = 0, path_w = 0, dp_pairs = 0, span_bases = 0;
    const PairRange* hr = ctx->h_ranges.as<PairRange>();
    for (uint32_t p = 0; p < np; p++) {
        ha[p].trace_off = trace_recs; ha[p].xam_off = xam_n; ha[p].path_off = path_w;
        if (hr[p].pass) {
            int ql = hr[p].e1 - hr[p].s1, tl = hr[p].e2 - hr[p].s2;
            uint64_t md = max_d_of(ql, tl);
            trace_recs += md + 1; xam_n += (uint64_t)tl + 2; path_w += md / 32 + 2;
            dp_pairs++; span_bases += (uint64_t)ql + tl;
        }
    }
    CK(ctx->d_trace.reserve(trace_recs * TRACE_REC_WORDS * 4 + 64));
    CK(ctx->d_xam.reserve(xam_n * 4 + 64));
    CK(ctx->d_ent.reserve(xam_n * 4 + 64));
    CK(ctx->d_M.reserve(m_total * 4 + 64));
    CK(ctx->d_path.reserve(path_w * 4 + 64));
    if (np) CK(cudaMemcpyAsync(ctx->d_allocs.p, ha.data(), (size_t)np * sizeof(PairAlloc), cudaMemcpyHostToDevice, st));
    // ---- DP
    CK(cudaEventRecord(ctx->ev[3], st));
    if (np) {
        k_dp<<<(np + DP_WARPS - 1) / DP_WARPS, DP_WARPS * 32, 0, st>>>(
            ctx->d_blocks.as<BlockDesc>(), ctx->d_pairs.as<PairDesc>(), ctx->d_ranges.as<PairRange>(),
            ctx->d_allocs.as<PairAlloc>(), np, pool, ctx->d_trace.as<uint32_t>(), 1.0 - min_idt, ctx->d_aln.as<PairAln>());
        CK(cudaGetLastError());
        ctx->counters[FCX_C_KERNEL_LAUNCHES] += 1;
    }
    CK(cudaEventRecord(ctx->ev[4], st));
    // ---- traceback
    if (np) {
        k_traceback<<<(np + 127) / 128, 128, 0, st>>>(
            ctx->d_blocks.as<BlockDesc>(), ctx->d_pairs.as<PairDesc>(), ctx->d_ranges.as<PairRange>(),
            ctx->d_allocs.as<PairAlloc>(), np, pool, ctx->d_trace.as<uint32_t>(), ctx->d_path.as<uint32_t>(),
            ctx->d_xam.as<uint32_t>(), ctx->d_ent.as<uint32_t>(), ctx->d_aln.as<PairAln>());
        CK(cudaGetLastError());
        ctx->counters[FCX_C_KERNEL_LAUNCHES] += 1;
    }
    if (tiles) {
        k_transpose<<<(unsigned)((tiles + TR_WARPS - 1) / TR_WARPS), TR_WARPS * 32, 0, st>>>(
            ctx->d_blocks.as<BlockDesc>(), nb, (uint32_t)tiles, ctx->d_ranges.as<PairRange>(),
            ctx->d_allocs.as<PairAlloc>(), ctx->d_aln.as<PairAln>(), ctx->d_ent.as<uint32_t>(), ctx->d_M.as<uint32_t>());
        CK(cudaGetLastError());
        ctx->counters[FCX_C_KERNEL_LAUNCHES] += 1;
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    // ---- consensus
    {
        auto kfn = ctx->profile ? k_consensus<true> : k_consensus<false>;
        kfn<<<cns_grid, CNS_WARPS * 32, 0, st>>>(
            ctx->d_blocks.as<BlockDesc>(), nb, ctx->d_pairs.as<PairDesc>(), ctx->d_ranges.as<PairRange>(),
            ctx->d_allocs.as<PairAlloc>(), ctx->d_aln.as<PairAln>(), pool, ctx->d_xam.as<uint32_t>(),
            ctx->d_M.as<uint32_t>(), ctx->d_recs.as<CnsRec>(), ctx->d_lvl.as<int32_t>(),
            ctx->d_acc.as<ReadMeta>(), (uint64_t)max_np, ctx->d_cns.as<char>(), ctx->d_eqv.as<int32_t>(), min_cov,
            ctx->d_cnsout.as<CnsOut>());
    }
    CK(cudaGetLastError());
    ctx->counters[FCX_C_KERNEL_LAUNCHES] += 1;
    CK(cudaEventRecord(ctx->ev[6], st));
    CK(cudaMemcpyAsync(ctx->h_cnsout.p, ctx->d_cnsout.p, nb * sizeof(CnsOut), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_cns.p, ctx->d_cns.p, cns_total, cudaMemcpyDeviceToHost, st));
    if (np) CK(cudaMemcpyAsync(ctx->h_aln.p, ctx->d_aln.p, (size_t)np * sizeof(PairAln), cudaMemcpyDeviceToHost, st));
    if (ctx->want_eqv) {
        CK(ctx->h_eqv.reserve(cns_total * 4));
        CK(cudaMemcpyAsync(ctx->h_eqv.p, ctx->d_eqv.p, cns_total * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));

    // ---- collect
    const CnsOut* co = ctx->h_cnsout.as<CnsOut>();
    const char* hc = ctx->h_cns.as<char>();
    for (uint32_t b = 0; b < nb; b++) {
        if (co[b].err) {
            char buf[256];
            snprintf(buf, sizeof buf, "consensus kernel error %d in block %u (1: link table overflow, 2: record overflow, 3: no best score (reference asserts, falcon.c:476))",
                     co[b].err, b0 + b);
            ctx->err = buf; return 3;
        }
        ctx->prof[0] += co[b].deep_positions; ctx->prof[1] += co[b].positions;
        ctx->prof[2] += (double)co[b].cyc_vote; ctx->prof[3] += (double)co[b].cyc_dp;
        ctx->prof[4] += (double)co[b].cyc_generic; ctx->prof[5] += (double)co[b].cyc_backtrack;
        ctx->out_bases.insert(ctx->out_bases.end(), hc + hb[b].cns_off, hc + hb[b].cns_off + co[b].len);
        ctx->out_off.push_back(ctx->out_bases.size());
        if (ctx->want_eqv) {
            const int32_t* he = ctx->h_eqv.as<int32_t>();
            ctx->out_eqv.insert(ctx->out_eqv.end(), he + hb[b].cns_off, he + hb[b].cns_off + co[b].len);
        }
    }
    const PairAln* hal = ctx->h_aln.as<PairAln>();
    uint64_t cells = 0, steps = 0, cols = 0, accepted = 0;
    for (uint32_t p = 0; p < np; p++) {
        cells += (uint64_t)hal[p].cells; accepted += hal[p].accepted;
        if (hal[p].aligned) steps += (uint64_t)hal[p].dist + 1;
        if (hal[p].accepted) cols += (uint64_t)hal[p].aln_size;
        if (ctx->keep_pair_info) {
            fcx_pair_info pi;
            pi.n_match = hr[p].n_match; pi.s1 = hr[p].s1; pi.e1 = hr[p].e1; pi.s2 = hr[p].s2; pi.e2 = hr[p].e2;
            pi.passed_filter = hr[p].pass; pi.aligned = hal[p].aligned; pi.dist = hal[p].dist;
            pi.aln_size = hal[p].aln_size; pi.q_e = hal[p].q_e; pi.t_e = hal[p].t_e;
            pi.accepted = hal[p].accepted; pi.n_tags = hal[p].n_tags; pi.trace_cells = hal[p].cells;
            ctx->pair_info.push_back(pi);
        }
    }
    (void)pair_base;
    ctx->counters[FCX_C_PAIRS] += np; ctx->counters[FCX_C_DP_PAIRS] += dp_pairs;
    ctx->counters[FCX_C_ACCEPTED] += accepted; ctx->counters[FCX_C_TRACE_CELLS] += cells;
    ctx->counters[FCX_C_DP_STEPS] += steps; ctx->counters[FCX_C_ALN_COLS] += cols;
    ctx->counters[FCX_C_SPAN_BASES] += span_bases; ctx->counters[FCX_C_WAVES] += 1;
    float ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->times[FCX_T_INDEX] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->times[FCX_T_RANGE] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); ctx->times[FCX_T_DP] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->times[FCX_T_TRACEBACK] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]); ctx->times[FCX_T_CONSENSUS] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[6]); ctx->times[FCX_T_TOTAL] += ms;
    return 0;
}

}  // namespace

extern "C" int fcx_consensus_blocks(fcx_ctx* ctx, uint32_t n_blocks, const uint32_t* block_off,
                                    const uint32_t* read_ids, unsigned min_cov, unsigned K, double min_idt,
                                    const char** out_bases, const uint64_t** out_off) {
    CK(cudaSetDevice(ctx->device));
    if (K != KMER) { ctx->err = "K must be 8 (falcon_kit/mains/consensus.py:270)"; return 1; }
    ctx->out_bases.clear(); ctx->out_off.clear(); ctx->out_off.push_back(0); ctx->pair_info.clear();
    ctx->out_eqv.clear();
    memset(ctx->times, 0, sizeof ctx->times); memset(ctx->counters, 0, sizeof ctx->counters);
    memset(ctx->prof, 0, sizeof ctx->prof);
    for (uint32_t b = 0; b < n_blocks; b++) {
        if (block_off[b + 1] <= block_off[b]) { ctx->err = "empty block (a block needs at least the seed)"; return 1; }
        if (block_off[b + 1] - block_off[b] > 65000) { ctx->err = "more than 65000 reads in one block"; return 1; }
        for (uint32_t i = block_off[b]; i < block_off[b + 1]; i++)
            if (read_ids[i] >= ctx->n_reads) { ctx->err = "read id outside the uploaded pool"; return 1; }
    }
    // plan waves from upper bounds (exact sizes are computed per wave after k_range)
    uint32_t b = 0; uint64_t pair_base = 0;
    while (b < n_blocks) {
        uint32_t e = b; uint64_t pairs = 0; double bytes = 0;
        while (e < n_blocks) {
            uint32_t lo = block_off[e], hi = block_off[e + 1];
            int slen = ctx->h_len[read_ids[lo]];
            double bb = (double)KTAB * 4 + (double)slen * (4 + 8 * 12 + 2 * 5 + 2) + 4.0 * slen * (((hi - lo - 1) + 31) & ~31u);
            for (uint32_t i = lo + 1; i < hi; i++) {
                int rl = ctx->h_len[read_ids[i]];
                bb += 0.3 * (rl + slen) * 36.0 + 8.0 * (slen + 2) + 128;
            }
            if (e > b && (bytes + bb > (double)ctx->arena_budget || pairs + (hi - lo - 1) > ctx->max_wave_pairs ||
                          e - b >= ctx->max_wave_blocks)) break;
            bytes += bb; pairs += hi - lo - 1; e++;
        }
        int rc = run_wave(ctx, b, e, block_off, read_ids, min_cov, min_idt, pair_base);
        if (rc) return rc;
        pair_base += pairs; b = e;
    }
    *out_bases = ctx->out_bases.data();
    *out_off = ctx->out_off.data();
    return 0;
}

extern "C" int fcx_last_pair_info(fcx_ctx* ctx, fcx_pair_info* out, uint64_t max_pairs, uint64_t* n_pairs) {
    uint64_t n = ctx->pair_info.size();
    if (n_pairs) *n_pairs = n;
    if (out) memcpy(out, ctx->pair_info.data(), std::min(n, max_pairs) * sizeof(fcx_pair_info));
    return 0;
}

extern "C" int fcx_last_stats(fcx_ctx* ctx, double* times_ms, uint64_t* counters) {
    if (times_ms) memcpy(times_ms, ctx->times, sizeof ctx->times);
    if (counters) memcpy(counters, ctx->counters, sizeof ctx->counters);
    return 0;
}

extern "C" int fcx_internal_profile(fcx_ctx* ctx, double* out8) { memcpy(out8, ctx->prof, sizeof ctx->prof); return 0; }

// CUDA-event stopwatch on the engine's stream (bench.py brackets its timed region with it)
extern "C" int fcx_timer_start(fcx_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->tev[0], ctx->stream));
    return 0;
}
extern "C" int fcx_timer_stop(fcx_ctx* ctx, double* ms) {
    CK(cudaEventRecord(ctx->tev[1], ctx->stream));
    CK(cudaEventSynchronize(ctx->tev[1]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, ctx->tev[0], ctx->tev[1]));
    *ms = f;
    return 0;
}

// ---------------------------------------------------------------------------------- internal hooks
extern "C" int fcx_internal_want_eqv(fcx_ctx* ctx, int on) { ctx->want_eqv = on != 0; return 0; }
extern "C" int fcx_internal_last_eqv(fcx_ctx* ctx, const int32_t** eqv, uint64_t* n) {
    *eqv = ctx->out_eqv.data(); *n = ctx->out_eqv.size(); return 0;
}
extern "C" int fcx_internal_align(fcx_ctx* ctx, const char*, int, const char*, int, int, int, alignment*) {
    ctx->err = "align(): single-pair GPU entry not implemented yet";
    return 1;
}
