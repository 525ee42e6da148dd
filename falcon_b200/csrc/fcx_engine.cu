// fcx_engine.cu -- host side of libfalcon_b200.so: device memory, wave planning, launches and
// the batched C ABI declared in include/falcon_b200.h.  There is NO CPU path for the arithmetic:
// every stage runs in the CUDA kernels of fcx_kernels.cuh / fcx_consensus.cuh and any CUDA failure
// is reported (fcx_*) or fatal (legacy symbols), never papered over.
//
// Execution model: a call is cut into WAVES of seed blocks; waves are independent and run on LANES
// (one host thread + CUDA stream + buffer set each).  Several lanes in flight let the latency-bound
// kernels of one wave (k_consensus: one warp per block, a serial chain over the seed) overlap the
// issue-bound kernels of another (k_dp), which is where the throughput comes from.
#include "fcx_kernels.cuh"
#include "fcx_trim.cuh"
#include "../../include/falcon_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
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

struct Lane {                 // one in-flight wave: stream, events, buffers, statistics
    cudaStream_t stream = nullptr;
    cudaStream_t stream_hi = nullptr;   // high priority: the latency-bound consensus kernel
    cudaEvent_t ev[8] = {nullptr};
    DevBuf d_blocks, d_pairs, d_ranges, d_allocs, d_aln, d_ktab, d_kpos, d_trace, d_path, d_xck, d_ent, d_slots,
           d_ovf, d_vmeta, d_rlist, d_recs, d_lvl, d_cns, d_eqv, d_cnsout, d_counter, d_tbhist, d_order, d_kbits;
    HostBuf h_ranges, h_aln, h_cns, h_cnsout, h_eqv;
    std::string err;
    double times[FCX_T_COUNT] = {0};
    uint64_t counters[FCX_C_COUNT] = {0};
    double prof[8] = {0};
    uint32_t ovf_scale = 1;            // size factor of the vote overflow arena (doubled on retry)
    uint32_t rec_scale = 1;            // size factor of the consensus record arena (doubled on retry)
    void release() {
        DevBuf* bufs[] = {&d_blocks, &d_pairs, &d_ranges, &d_allocs, &d_aln, &d_ktab, &d_kpos, &d_trace, &d_path,
                          &d_xck, &d_ent, &d_slots, &d_ovf, &d_vmeta, &d_rlist, &d_recs, &d_lvl, &d_cns, &d_eqv, &d_cnsout,
                          &d_counter, &d_tbhist, &d_order, &d_kbits};
        for (auto* b : bufs) b->release();
        HostBuf* hb[] = {&h_ranges, &h_aln, &h_cns, &h_cnsout, &h_eqv};
        for (auto* b : hb) b->release();
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
        if (stream_hi) { cudaStreamDestroy(stream_hi); stream_hi = nullptr; }
    }
};

struct WaveResult {
    std::vector<char> bases;
    std::vector<uint64_t> lens;
    std::vector<fcx_pair_info> info;
    std::vector<int32_t> eqv;
};

}  // namespace

struct fcx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;     // pool uploads, single-pair align, stopwatch
    std::string err;
    // pool
    uint32_t n_reads = 0;              // committed pool size
    uint32_t n_reserved = 0;           // reads in the reserved layout (fcx_pool_reserve)
    std::vector<uint64_t> h_woff;      // n_reads + 1
    std::vector<int32_t> h_len;
    DevBuf d_pool, d_ascii, d_aoff, d_woff, d_len, d_dirty;
    // single-pair align scratch
    DevBuf d_trace1, d_path1, d_aln1, d_str1;
    std::vector<Lane> lanes;
    int sm_count = 148;
    bool profile = false;
    // results of the last call
    std::vector<char> out_bases;
    std::vector<uint64_t> out_off;
    std::vector<fcx_pair_info> pair_info;
    std::vector<int32_t> out_eqv;
    bool want_eqv = false;
    bool keep_pair_info = true;
    double times[FCX_T_COUNT] = {0};
    uint64_t counters[FCX_C_COUNT] = {0};
    double prof[8] = {0};
    cudaEvent_t tev[2] = {nullptr, nullptr};
    size_t arena_budget = (size_t)160 << 30;
    uint32_t max_wave_blocks = 2960;     // 148 SMs x 20 resident consensus warps (set from the SM count)
    uint32_t max_wave_pairs = 1u << 19;
    uint32_t min_wave_blocks = 384;
    int n_lanes = 3;
    int active_lanes = 0;              // 0 = all
    int dp_variant = 3;                // 3: k_dp3 (default); 1: k_dp (round-1 kernel); 2: k_dp with TMA-staged spans
    uint32_t debug_split_above = 0;    // test hook: pretend waves with more blocks than this do not fit
    bool debug_tiny_capacity = false;  // test hook: start every wave with arenas that are too small
    // --trim: block lists of the last fcx_trim_blocks call
    std::vector<uint32_t> trim_block_off, trim_read_ids;
    DevBuf d_trim_scratch, d_trim_out, d_subs, d_sub_wb;
};

static thread_local std::string g_create_err;

#define CKE(errstr, call)                                                                 \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            char buf_[512];                                                               \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call,                   \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                         \
            (errstr) = buf_;                                                              \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)
#define CK(call) CKE(ctx->err, call)
#define CKL(call) CKE(L.err, call)
// buffer growth inside a wave: an out-of-memory condition is reported as 100 so that the caller
// can split the wave instead of failing
#define CKR(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ == cudaErrorMemoryAllocation) { cudaGetLastError(); L.err = "out of device memory"; return 100; } \
        if (e_ != cudaSuccess) {                                                          \
            char buf_[512];                                                               \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call,                   \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                         \
            L.err = buf_;                                                                 \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

extern "C" const char* fcx_version(void) { return "falcon_b200 0.3 sm_100a"; }

extern "C" int fcx_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

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
    // every CUDA call below is checked: a failure here would otherwise surface later as an
    // unrelated launch or event error
#define CKC(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            g_create_err = std::string(#call " failed: ") + cudaGetErrorString(e_);        \
            fcx_destroy(ctx); return 1;                                                    \
        }                                                                                  \
    } while (0)
    for (auto& ev : ctx->tev) CKC(cudaEventCreate(&ev));
    if (const char* s = getenv("FCX_ARENA_GB")) ctx->arena_budget = (size_t)(atof(s) * (double)((size_t)1 << 30));
    if (const char* s = getenv("FCX_WAVE_BLOCKS")) { ctx->max_wave_blocks = (uint32_t)atoi(s); ctx->min_wave_blocks = std::min(ctx->min_wave_blocks, ctx->max_wave_blocks); }
    if (const char* s = getenv("FCX_WAVE_PAIRS")) ctx->max_wave_pairs = (uint32_t)atoi(s);
    if (const char* s = getenv("FCX_LANES")) ctx->n_lanes = std::max(1, atoi(s));
    if (const char* s = getenv("FCX_PROFILE")) ctx->profile = atoi(s) != 0;
    if (const char* s = getenv("FCX_DP_VARIANT")) ctx->dp_variant = atoi(s);
#ifndef FCX_EMU
    CKC(cudaFuncSetAttribute(k_dp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
#endif
    CKC(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    if (!getenv("FCX_WAVE_BLOCKS")) ctx->max_wave_blocks = (uint32_t)(ctx->sm_count * 20);
    {   // never plan beyond what the device can actually give
        size_t free_b = 0, total_b = 0;
        CKC(cudaMemGetInfo(&free_b, &total_b));
        if (!getenv("FCX_ARENA_GB")) ctx->arena_budget = std::min(ctx->arena_budget, (size_t)((double)free_b * 0.85));
    }
    CKC(cudaFuncSetAttribute(k_range, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (KTAB / 32) * 4 + 4 * RANGE_BINS * (int)sizeof(int) + 24 * 1024));
    ctx->lanes.resize(ctx->n_lanes);
    int prio_lo = 0, prio_hi = 0;
    CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (auto& L : ctx->lanes) {
        CKC(cudaStreamCreateWithPriority(&L.stream, cudaStreamNonBlocking, prio_lo));
        CKC(cudaStreamCreateWithPriority(&L.stream_hi, cudaStreamNonBlocking, prio_hi));
        for (auto& ev : L.ev) CKC(cudaEventCreate(&ev));
    }
#undef CKC
    *out = ctx;
    return 0;
}

extern "C" void fcx_destroy(fcx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&ctx->d_pool, &ctx->d_ascii, &ctx->d_aoff, &ctx->d_woff, &ctx->d_len, &ctx->d_dirty,
                      &ctx->d_trace1, &ctx->d_path1, &ctx->d_aln1, &ctx->d_str1,
                      &ctx->d_trim_scratch, &ctx->d_trim_out, &ctx->d_subs, &ctx->d_sub_wb};
    for (auto* b : bufs) b->release();
    for (auto& L : ctx->lanes) L.release();
    for (auto& ev : ctx->tev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void* fcx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void fcx_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int fcx_set_option(fcx_ctx* ctx, const char* name, double value) {
    std::string n(name);
    if (n == "pair_info") ctx->keep_pair_info = value != 0;
    else if (n == "eqv") ctx->want_eqv = value != 0;
    else if (n == "profile") ctx->profile = value != 0;
    else if (n == "arena_gb") ctx->arena_budget = (size_t)(value * (double)((size_t)1 << 30));
    else if (n == "max_wave_blocks") ctx->max_wave_blocks = (uint32_t)value;
    else if (n == "min_wave_blocks") ctx->min_wave_blocks = (uint32_t)value;
    else if (n == "dp_variant") { if (value != 1 && value != 2 && value != 3) { ctx->err = "dp_variant must be 1, 2 or 3"; return 1; } ctx->dp_variant = (int)value; }
    else if (n == "debug_split_above") ctx->debug_split_above = (uint32_t)value;
    else if (n == "debug_tiny_capacity") ctx->debug_tiny_capacity = value != 0;
    else if (n == "lanes") ctx->active_lanes = value <= 0 ? 0 : std::min((int)value, (int)ctx->lanes.size());
    else { ctx->err = "unknown option: " + n; return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------- pool
// The pool is built in three steps so that several processes / devices can each contribute a part
// (one NCCL broadcast or peer copy per part): reserve (layout from the read lengths),
// upload_part (ASCII of a contiguous range of reads -> packed in place), commit.
extern "C" int fcx_pool_reserve(fcx_ctx* ctx, const uint64_t* offsets, uint32_t n_reads, uint64_t* total_words) {
    CK(cudaSetDevice(ctx->device));
    ctx->n_reads = 0; ctx->n_reserved = 0;
    ctx->h_woff.assign((size_t)n_reads + 1, 0);
    ctx->h_len.assign(n_reads, 0);
    uint64_t w = 0;
    for (uint32_t r = 0; r < n_reads; r++) {
        uint64_t len = offsets[r + 1] - offsets[r];
        // any length may live in the pool (fcx_align_pairs takes contig-sized sequences); the limits
        // of the consensus path -- reads <= 100000 (consensus.py:178-179), seeds < 100000
        // (falcon.c:343) -- are checked per block in fcx_consensus_blocks
        if (len > 0x7fffff00ull) { ctx->err = "sequence longer than 2^31 bases"; return 1; }
        ctx->h_len[r] = (int32_t)len;
        ctx->h_woff[r] = w;
        uint64_t words = (len + 15) / 16 + 1;            // +1 zero pad word: fetch16 reads one word ahead
        w += (words + 3) & ~(uint64_t)3;                 // 16-byte aligned starts
    }
    ctx->h_woff[n_reads] = w;
    CK(ctx->d_pool.reserve((w + 4) * 4));
    CK(ctx->d_dirty.reserve(4));
    CK(cudaMemsetAsync((char*)ctx->d_pool.p + w * 4, 0, 16, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_reserved = n_reads;
    if (total_words) *total_words = w;
    return 0;
}

extern "C" int fcx_pool_upload_part(fcx_ctx* ctx, const char* bases, const uint64_t* offsets, uint32_t first_read,
                                    uint32_t n_part) {
    CK(cudaSetDevice(ctx->device));
    if ((uint64_t)first_read + n_part > ctx->n_reserved) { ctx->err = "fcx_pool_upload_part: range outside the reserved pool"; return 1; }
    if (n_part == 0) return 0;
    for (uint32_t r = 0; r < n_part; r++)
        if ((int64_t)(offsets[r + 1] - offsets[r]) != (int64_t)ctx->h_len[first_read + r]) {
            ctx->err = "fcx_pool_upload_part: read length differs from the reserved layout"; return 1;
        }
    const uint64_t total_bytes = offsets[n_part] - offsets[0];
    const uint64_t w0 = ctx->h_woff[first_read], w1 = ctx->h_woff[first_read + n_part];
    CK(ctx->d_ascii.reserve(total_bytes + 16));
    CK(ctx->d_aoff.reserve(((size_t)n_part + 1) * 8));
    CK(ctx->d_woff.reserve(((size_t)n_part + 1) * 8));
    CK(ctx->d_len.reserve((size_t)n_part * 4 + 4));
    std::vector<uint64_t> rel((size_t)n_part + 1), wrel((size_t)n_part + 1);
    for (uint32_t r = 0; r <= n_part; r++) { rel[r] = offsets[r] - offsets[0]; wrel[r] = ctx->h_woff[first_read + r] - w0; }
    CK(cudaMemcpyAsync(ctx->d_ascii.p, bases + offsets[0], total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_aoff.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_woff.p, wrel.data(), wrel.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_len.p, ctx->h_len.data() + first_read, (size_t)n_part * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_dirty.p, 0, 4, ctx->stream));
    if (w1 > w0) {
        const uint64_t nb = (w1 - w0 + 255) / 256;
        FCX_LAUNCH(k_pack, (unsigned)nb, 256, 0, ctx->stream, ctx->d_ascii.as<uint8_t>(), ctx->d_aoff.as<uint64_t>(),
                   ctx->d_woff.as<uint64_t>(), ctx->d_len.as<int32_t>(), n_part, w1 - w0,
                   ctx->d_pool.as<uint32_t>() + w0, ctx->d_dirty.as<int>());
        CK(cudaGetLastError());
    }
    int dirty = 0;
    CK(cudaMemcpyAsync(&dirty, ctx->d_dirty.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (dirty) { ctx->err = "read pool contains bytes outside upper-case ACGT (reference behaviour undefined: falcon.c:370-379)"; return 2; }
    return 0;
}

// Device view of the packed pool (for a collective or a peer copy issued by the caller): base
// pointer, number of 32-bit words, and the host array of n_reads + 1 word offsets.
extern "C" int fcx_pool_device(fcx_ctx* ctx, void** dev_words, uint64_t* n_words, const uint64_t** word_off) {
    if (!ctx->n_reserved) { ctx->err = "fcx_pool_device: no pool reserved"; return 1; }
    if (dev_words) *dev_words = ctx->d_pool.p;
    if (n_words) *n_words = ctx->h_woff[ctx->n_reserved];
    if (word_off) *word_off = ctx->h_woff.data();
    return 0;
}

extern "C" int fcx_pool_commit(fcx_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->n_reads = ctx->n_reserved;
    return 0;
}

extern "C" int fcx_pool_upload(fcx_ctx* ctx, const char* bases, const uint64_t* offsets, uint32_t n_reads) {
    if (int rc = fcx_pool_reserve(ctx, offsets, n_reads, nullptr)) return rc;
    if (int rc = fcx_pool_upload_part(ctx, bases, offsets, 0, n_reads)) return rc;
    ctx->n_reads = n_reads;
    return 0;
}

// Pool from a Dazzler .bps image (fcx_dazz.cu): 2 entries per read, forward and reverse complement.
extern "C" int fcx_pool_upload_bps(fcx_ctx* ctx, const uint8_t* bps, uint64_t n_bytes, const uint64_t* boff,
                                   const int32_t* rlen, uint32_t n_reads) {
    CK(cudaSetDevice(ctx->device));
    if ((uint64_t)n_reads * 2 > 0xfffffff0ull) { ctx->err = "fcx_pool_upload_bps: too many reads"; return 1; }
    const uint32_t ne = n_reads * 2;
    std::vector<uint64_t> off((size_t)ne + 1, 0);
    std::vector<int32_t> elen(ne);
    for (uint32_t r = 0; r < n_reads; r++) {
        if (rlen[r] < 0 || boff[r] + (uint64_t)(rlen[r] + 3) / 4 > n_bytes) { ctx->err = "fcx_pool_upload_bps: read outside the .bps image"; return 1; }
        const int32_t l = rlen[r] > 100000 ? 99999 : rlen[r];               // consensus.py:178-179
        elen[2 * r] = elen[2 * r + 1] = l;
        off[2 * r + 1] = off[2 * r] + (uint64_t)l; off[2 * r + 2] = off[2 * r + 1] + (uint64_t)l;
    }
    uint64_t total_words = 0;
    if (int rc = fcx_pool_reserve(ctx, off.data(), ne, &total_words)) return rc;
    cudaStream_t st = ctx->stream;
    DevBuf d_bps, d_boff, d_rlen, d_elen, d_woff;
    auto fail = [&](const char* what, cudaError_t e) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        d_bps.release(); d_boff.release(); d_rlen.release(); d_elen.release(); d_woff.release();
        return 1;
    };
    cudaError_t e;
    if ((e = d_bps.reserve(n_bytes + 16)) != cudaSuccess) return fail("device memory for the .bps image", e);
    if ((e = d_boff.reserve((size_t)n_reads * 8 + 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = d_rlen.reserve((size_t)n_reads * 4 + 4)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = d_elen.reserve((size_t)ne * 4 + 4)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = d_woff.reserve(((size_t)ne + 1) * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    // the image is usually an mmap of the .bps file: staged through pinned memory in 64 MB pieces
    {
        HostBuf stage[2];
        const size_t piece = (size_t)64 << 20;
        if ((e = stage[0].reserve(piece)) != cudaSuccess || (e = stage[1].reserve(piece)) != cudaSuccess) {
            stage[0].release(); stage[1].release(); return fail("pinned staging memory", e);
        }
        cudaEvent_t ev[2]; cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]);
        int k = 0;
        for (uint64_t at = 0; at < n_bytes; at += piece, k ^= 1) {
            const size_t n = (size_t)std::min<uint64_t>(piece, n_bytes - at);
            cudaEventSynchronize(ev[k]);                               // the copy that last used this buffer
            memcpy(stage[k].p, bps + at, n);
            e = cudaMemcpyAsync((char*)d_bps.p + at, stage[k].p, n, cudaMemcpyHostToDevice, st);
            cudaEventRecord(ev[k], st);
            if (e != cudaSuccess) break;
        }
        cudaStreamSynchronize(st);
        cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
        stage[0].release(); stage[1].release();
        if (e != cudaSuccess) return fail("upload of the .bps image", e);
    }
    if (n_reads) {
        if ((e = cudaMemcpyAsync(d_boff.p, boff, (size_t)n_reads * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("cudaMemcpyAsync", e);
        if ((e = cudaMemcpyAsync(d_rlen.p, rlen, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("cudaMemcpyAsync", e);
        if ((e = cudaMemcpyAsync(d_elen.p, elen.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("cudaMemcpyAsync", e);
        if ((e = cudaMemcpyAsync(d_woff.p, ctx->h_woff.data(), ((size_t)ne + 1) * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("cudaMemcpyAsync", e);
        if (total_words) {
            FCX_LAUNCH(k_repack_bps, (unsigned)((total_words + 255) / 256), 256, 0, st, d_bps.as<uint8_t>(), d_boff.as<uint64_t>(),
                       d_rlen.as<int32_t>(), d_elen.as<int32_t>(), d_woff.as<uint64_t>(), ne, total_words, ctx->d_pool.as<uint32_t>());
            if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_repack_bps", e);
        }
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail("k_repack_bps", e);
    }
    d_bps.release(); d_boff.release(); d_rlen.release(); d_elen.release(); d_woff.release();
    ctx->n_reads = ne;
    return 0;
}

// ---------------------------------------------------------------------------------- waves
namespace {

inline uint64_t max_d_of(int q_len, int t_len) { return (uint64_t)(int)(0.3 * (q_len + t_len)); }

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int run_wave(fcx_ctx* ctx, Lane& L, uint32_t b0, uint32_t b1, const uint32_t* block_off, const uint32_t* read_ids,
             unsigned min_cov, double min_idt, WaveResult& res) {
    const uint32_t nb = b1 - b0;
    static const bool trace_waves = getenv("FCX_TRACE_WAVES") != nullptr;
    double tl[8] = {0}; tl[0] = now_ms();
    if (ctx->debug_split_above && nb > ctx->debug_split_above) { L.err = "out of device memory (simulated)"; return 100; }
    std::vector<BlockDesc> hb(nb);
    uint64_t npairs64 = 0, kpos_total = 0, rec_total = 0, cns_total = 0, slot_total = 0, tiles = 0;
    uint32_t max_np = 1; int max_slen = 1, max_rlen = 1;
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t lo = block_off[b0 + b], hi = block_off[b0 + b + 1];
        BlockDesc& d = hb[b];
        uint32_t seed = read_ids[lo];
        d.seed_woff = ctx->h_woff[seed];
        d.slen = ctx->h_len[seed];
        d.pair_begin = (uint32_t)npairs64;
        d.n_pairs = hi - lo - 1;
        d.kpos_off = kpos_total; kpos_total += (uint64_t)std::max(d.slen, 1);
        // consensus column records: live (position, delta, base) columns.  Measured ~4 per position
        // at 50x; the number grows with coverage (every column needs a vote), hence the n_pairs term.
        // k_consensus reports an overflow as an error, it never drops a column.
        d.rec_cap = (uint32_t)std::min<int64_t>(0x7fffffff, (int64_t)L.rec_scale * std::max<int64_t>(64, (int64_t)d.slen * (8 + (int64_t)d.n_pairs / 16) + 64));
        if (ctx->debug_tiny_capacity && L.rec_scale == 1) d.rec_cap = 64;
        d.rec_off = rec_total; rec_total += d.rec_cap;
        d.cns_off = cns_total; cns_total += (uint64_t)d.slen * 2 + 8;
        d.pad_ = 0;
        d.slot_off = slot_total; slot_total += (uint64_t)std::max(d.slen, 1);
        d.tile_begin = (uint32_t)tiles; tiles += (uint64_t)((d.slen + VOTE_TP - 1) / VOTE_TP);
        npairs64 += d.n_pairs;
        max_np = std::max(max_np, d.n_pairs);
        max_slen = std::max(max_slen, d.slen);
    }
    const uint32_t np = (uint32_t)npairs64;
    std::vector<PairDesc> hp(np);
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t lo = block_off[b0 + b];
        for (uint32_t j = 0; j < hb[b].n_pairs; j++) {
            uint32_t rid = read_ids[lo + 1 + j];
            PairDesc& pd = hp[hb[b].pair_begin + j];
            pd.read_woff = ctx->h_woff[rid]; pd.block = b; pd.rlen = ctx->h_len[rid];
            max_rlen = std::max(max_rlen, pd.rlen);
        }
    }
    cudaStream_t st = L.stream;
    const size_t np1 = std::max(np, 1u);
    CKR(L.d_blocks.reserve(nb * sizeof(BlockDesc)));
    CKR(L.d_pairs.reserve(np1 * sizeof(PairDesc)));
    CKR(L.d_ranges.reserve(np1 * sizeof(PairRange)));
    CKR(L.d_allocs.reserve(np1 * sizeof(PairAlloc)));
    CKR(L.d_aln.reserve(np1 * sizeof(PairAln)));
    CKR(L.d_ktab.reserve((size_t)nb * KTAB * 4));
    CKR(L.d_kpos.reserve(kpos_total * 4 + 16));
    CKR(L.d_recs.reserve(rec_total * sizeof(CnsRec)));
    CKR(L.d_cns.reserve(cns_total));
    CKR(L.d_eqv.reserve(cns_total * 4));
    CKR(L.d_cnsout.reserve(nb * sizeof(CnsOut)));
    CKR(L.d_lvl.reserve((size_t)nb * 4 * CDP_LEVELS * 5 * 4));
    CKR(L.d_vmeta.reserve(np1 * sizeof(VoteMeta)));
    CKR(L.d_slots.reserve(slot_total * VSLOT * sizeof(uint2) + 64));
    // links beyond the 15 of a position's own slot: rare (deep coverage); sized generously, and an
    // overflow is reported (err 2) so that the caller can retry with a larger arena
    const uint32_t ovf_cap = (uint32_t)std::min<uint64_t>((uint64_t)L.ovf_scale * std::max<uint64_t>(1u << 20, slot_total * 2), 0x7fffffffu);
    CKR(L.d_ovf.reserve((size_t)ovf_cap * sizeof(uint2)));
    const uint32_t ovf_cap_used = (ctx->debug_tiny_capacity && L.ovf_scale == 1) ? 8u : ovf_cap;
    CKL(L.h_ranges.reserve(np1 * sizeof(PairRange)));
    CKL(L.h_aln.reserve(np1 * sizeof(PairAln)));
    CKL(L.h_cns.reserve(cns_total));
    CKL(L.h_cnsout.reserve(nb * sizeof(CnsOut)));

    CKL(cudaMemcpyAsync(L.d_blocks.p, hb.data(), nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    if (np) CKL(cudaMemcpyAsync(L.d_pairs.p, hp.data(), (size_t)np * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
    const uint32_t* pool = ctx->d_pool.as<uint32_t>();

    tl[1] = now_ms();
    // ---- index
    CKL(cudaEventRecord(L.ev[0], st));
    CKL(cudaMemsetAsync(L.d_ktab.p, 0, (size_t)nb * KTAB * 4, st));
    CKR(L.d_kbits.reserve((size_t)nb * (KTAB / 32) * 4));
    FCX_LAUNCH(k_index, nb, 256, 0, st, L.d_blocks.as<BlockDesc>(), pool, L.d_ktab.as<uint32_t>(), L.d_kpos.as<uint32_t>(),
               L.d_kbits.as<uint32_t>());
    CKL(cudaGetLastError());
    CKL(cudaEventRecord(L.ev[1], st));
    L.counters[FCX_C_KERNEL_LAUNCHES] += 1;
    // ---- range
    if (np) {
        // histogram bins actually needed by this wave (diagonal range <= read + seed length)
        const int bins = std::min(RANGE_BINS, (max_rlen + max_slen) / BIN_SIZE + 8);
        // as many warps per CTA as the per-warp histograms allow (all on ONE seed block), at most two
        // CTAs per SM: the 256 KB bucket tables in use at any time stay L2 resident
        int rwarps = RANGE_WARPS_MAX;
        while (rwarps > 4 && (size_t)(KTAB / 32) * 4 + (size_t)rwarps * bins * sizeof(int) > 96 * 1024) rwarps /= 2;
        if (const char* e = getenv("FCX_RANGE_WARPS")) rwarps = std::max(1, std::min(RANGE_WARPS_MAX, atoi(e)));
        const size_t rsmem = (size_t)(KTAB / 32) * 4 + (size_t)rwarps * bins * sizeof(int);
        unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(2, (200 * 1024) / std::max<size_t>(rsmem, 1)));
        if (const char* e = getenv("FCX_RANGE_CTAS")) per_sm = (unsigned)std::max(1, atoi(e));
        const unsigned rgrid = std::min<unsigned>(nb, (unsigned)ctx->sm_count * per_sm);
        // per-warp match list: expected hits = query k-mers x (seed positions per bucket + true hits);
        // a list that still overflows (low-complexity sequence) takes the kernel's re-walking slow path
        const double nq_max = max_rlen / 4.0 + 1;
        int list_cap = (int)std::min<double>(1 << 20, std::max<double>(RANGE_LIST_CAP, 2.0 * nq_max * (max_slen / 65536.0 + 0.4)));
        list_cap = (list_cap + 31) & ~31;
        if (const char* e = getenv("FCX_RANGE_LIST_CAP")) list_cap = std::max(64, atoi(e) & ~31);
        CKR(L.d_rlist.reserve((size_t)rgrid * rwarps * list_cap * sizeof(uint32_t)));
        FCX_LAUNCH(k_range, rgrid, rwarps * 32, rsmem, st,
            L.d_blocks.as<BlockDesc>(), nb, L.d_pairs.as<PairDesc>(), pool, L.d_ktab.as<uint32_t>(),
            L.d_kpos.as<uint32_t>(), L.d_kbits.as<uint32_t>(), L.d_rlist.as<uint32_t>(), list_cap, bins, L.d_ranges.as<PairRange>());
        CKL(cudaGetLastError());
        L.counters[FCX_C_KERNEL_LAUNCHES] += 1;
        CKL(cudaMemcpyAsync(L.h_ranges.p, L.d_ranges.p, (size_t)np * sizeof(PairRange), cudaMemcpyDeviceToHost, st));
    }
    CKL(cudaEventRecord(L.ev[2], st));
    tl[2] = now_ms();
    CKL(cudaStreamSynchronize(st));
    tl[3] = now_ms();
    // ---- exact per-pair allocations
    std::vector<PairAlloc> ha(np);
    uint64_t trace_recs = 0, xam_n = 0, xck_n = 0, path_w = 0, dp_pairs = 0, span_bases = 0; uint32_t max_span = 0, max_trace_cap = 1;
    const PairRange* hr = L.h_ranges.as<PairRange>();
    for (uint32_t p = 0; p < np; p++) {
        ha[p].trace_off = trace_recs; ha[p].xam_off = xam_n; ha[p].path_off = path_w; ha[p].trace_cap = 0; ha[p].xck_off = (uint32_t)xck_n;
        if (hr[p].pass) {
            int ql = hr[p].e1 - hr[p].s1, tl = hr[p].e2 - hr[p].s2;
            // Only accepted pairs are traced back, and acceptance needs D / A < max_diff with
            // A = (q_e + t_e + D) / 2 <= (q + t + D) / 2, i.e. D < max_diff (q + t) / (2 - max_diff):
            // steps beyond that bound need no trace record.
            const double mdiff = std::max(0.0, 1.0 - min_idt);
            uint64_t md = max_d_of(ql, tl);
            if (mdiff < 1.999) md = std::min<uint64_t>(md, (uint64_t)(mdiff * (ql + tl) / (2.0 - mdiff)) + 2);
            ha[p].trace_cap = (uint32_t)(md + 1);
            max_trace_cap = std::max(max_trace_cap, (uint32_t)(md + 1));
            trace_recs += md + 1; xam_n += ((uint64_t)tl + 4 + 3) & ~(uint64_t)3; xck_n += (uint64_t)tl / 32 + 2; path_w += md / 32 + 2;
            dp_pairs++; span_bases += (uint64_t)ql + tl;
            max_span = std::max(max_span, (uint32_t)std::max(ql, tl));
        }
    }
    // k_dp3 walks every trace back inside the DP warp: one scratch trace per resident warp instead of one per pair
    const unsigned dp_grid = std::min<unsigned>((np + DP3_WARPS - 1) / DP3_WARPS, (unsigned)ctx->sm_count * 32u);
    if (ctx->dp_variant == 3) CKR(L.d_trace.reserve((size_t)std::max(dp_grid, 1u) * DP3_WARPS * max_trace_cap * TRACE_REC_WORDS * 4 + 64));
    else CKR(L.d_trace.reserve(trace_recs * TRACE_REC_WORDS * 4 + 64));
    if (xck_n > 0xffffffffull) { L.err = "out of device memory (checkpoint index)"; return 100; }    // split the wave
    CKR(L.d_xck.reserve(xck_n * 4 + 128));
    CKR(L.d_ent.reserve(xam_n * 4 + 128));
    CKR(L.d_path.reserve(path_w * 4 + 64));
    if (np) CKL(cudaMemcpyAsync(L.d_allocs.p, ha.data(), (size_t)np * sizeof(PairAlloc), cudaMemcpyHostToDevice, st));
    tl[4] = now_ms();
    // ---- DP
    CKL(cudaEventRecord(L.ev[3], st));
    if (np) {
        if (ctx->dp_variant == 3) {
            // default: one warp per pair, diagonals pinned to lanes, V in registers (fcx_dp.cuh)
            CKR(L.d_counter.reserve(64));
            CKL(cudaMemsetAsync(L.d_counter.p, 0, 64, st));
            FCX_LAUNCH(k_dp3, dp_grid, DP3_WARPS * 32, 0, st,
                       L.d_blocks.as<BlockDesc>(), L.d_pairs.as<PairDesc>(), L.d_ranges.as<PairRange>(),
                       L.d_allocs.as<PairAlloc>(), np, pool, L.d_trace.as<uint32_t>(), max_trace_cap, L.d_path.as<uint32_t>(), 1.0 - min_idt,
                       L.d_counter.as<uint32_t>(), L.d_aln.as<PairAln>());
        } else {
            // round-1 kernels kept for A/B measurement: shared-memory V ring, lanes re-mapped to the
            // band every step; dp_variant 2 stages the spans through TMA when they fit
            const int stage_words = (int)(((max_span + 15) / 16 + 1 + 8 + 3) & ~3u);
            const size_t smem_staged = (size_t)DP_WARPS * ((size_t)VRING * 4 + (size_t)stage_words * 8 + 16);
            const bool staged = ctx->dp_variant == 2 && smem_staged <= 100 * 1024;
#ifndef FCX_EMU
            if (staged) {
                FCX_LAUNCH(k_dp<true>, (np + DP_WARPS - 1) / DP_WARPS, DP_WARPS * 32, smem_staged, st,
                           L.d_blocks.as<BlockDesc>(), L.d_pairs.as<PairDesc>(), L.d_ranges.as<PairRange>(),
                           L.d_allocs.as<PairAlloc>(), np, pool, L.d_trace.as<uint32_t>(), 1.0 - min_idt, stage_words,
                           L.d_aln.as<PairAln>());
            } else
#endif
            {
                FCX_LAUNCH(k_dp<false>, (np + DP_WARPS - 1) / DP_WARPS, DP_WARPS * 32, (size_t)DP_WARPS * VRING * 4, st,
                           L.d_blocks.as<BlockDesc>(), L.d_pairs.as<PairDesc>(), L.d_ranges.as<PairRange>(),
                           L.d_allocs.as<PairAlloc>(), np, pool, L.d_trace.as<uint32_t>(), 1.0 - min_idt, 0,
                           L.d_aln.as<PairAln>());
            }
        }
        CKL(cudaGetLastError());
        L.counters[FCX_C_KERNEL_LAUNCHES] += 1;
    }
    CKL(cudaEventRecord(L.ev[4], st));
    // ---- traceback
    if (np) {
        CKL(cudaMemsetAsync(L.d_vmeta.p, 0, (size_t)np * sizeof(VoteMeta), st));
        // accepted pairs ordered by dist (longest first): warps of k_traceback walk paths of similar length
        CKR(L.d_tbhist.reserve((TB_BUCKETS + 8) * 4));
        CKR(L.d_order.reserve((size_t)np * 4));
        CKL(cudaMemsetAsync(L.d_tbhist.p, 0, (TB_BUCKETS + 8) * 4, st));
        FCX_LAUNCH(k_tb_hist, (np + 255) / 256, 256, 0, st, L.d_aln.as<PairAln>(), np, L.d_tbhist.as<uint32_t>());
        FCX_LAUNCH(k_tb_scan, 1, 1024, 0, st, L.d_tbhist.as<uint32_t>());
        FCX_LAUNCH(k_tb_scatter, (np + 255) / 256, 256, 0, st, L.d_aln.as<PairAln>(), np, L.d_tbhist.as<uint32_t>(),
                   L.d_order.as<uint32_t>());
        CKL(cudaGetLastError());
        if (ctx->dp_variant != 3) {       // k_dp3 walks back inside the DP warp
            FCX_LAUNCH(k_traceback_walk, (unsigned)std::min<uint32_t>((np + 3) / 4, (uint32_t)ctx->sm_count * 16u), 128, 0, st,
                L.d_allocs.as<PairAlloc>(), L.d_order.as<uint32_t>(), L.d_tbhist.as<uint32_t>() + TB_BUCKETS,
                L.d_trace.as<uint32_t>(), L.d_path.as<uint32_t>(), L.d_aln.as<PairAln>());
            CKL(cudaGetLastError());
            L.counters[FCX_C_KERNEL_LAUNCHES] += 1;
        }
        FCX_LAUNCH(k_traceback, (np + 127) / 128, 128, 0, st,
            L.d_blocks.as<BlockDesc>(), L.d_pairs.as<PairDesc>(), L.d_ranges.as<PairRange>(),
            L.d_allocs.as<PairAlloc>(), L.d_order.as<uint32_t>(), L.d_tbhist.as<uint32_t>() + TB_BUCKETS, pool,
            L.d_path.as<uint32_t>(),
            L.d_xck.as<uint32_t>(), L.d_ent.as<uint32_t>(), L.d_vmeta.as<VoteMeta>(), L.d_aln.as<PairAln>());
        CKL(cudaGetLastError());
        L.counters[FCX_C_KERNEL_LAUNCHES] += 4;
    }
    CKL(cudaEventRecord(L.ev[5], st));
    // ---- consensus: the column vote (parallel over positions), then the serial longest-path DP and
    // backtrack (one thread per seed block) on the lane's high-priority stream, so that its few
    // long-running warps are placed ahead of the next wave's bulk kernels and overlap them
    CKR(L.d_counter.reserve(64));
    CKL(cudaMemsetAsync((char*)L.d_counter.p + 16, 0, 16, st));      // [4] overflow cursor, [5] vote error flag
    if (tiles) {
        FCX_LAUNCH(k_vote, (unsigned)tiles, VOTE_TP, 0, st,
            L.d_blocks.as<BlockDesc>(), nb, (uint32_t)tiles, L.d_vmeta.as<VoteMeta>(), pool, L.d_xck.as<uint32_t>(),
            L.d_ent.as<uint32_t>(), L.d_slots.as<uint2>(), L.d_ovf.as<uint2>(), ovf_cap_used,
            L.d_counter.as<uint32_t>() + 4, L.d_counter.as<int>() + 5);
        CKL(cudaGetLastError());
        L.counters[FCX_C_KERNEL_LAUNCHES] += 1;
    }
    CKL(cudaEventRecord(L.ev[7], st));
    cudaStream_t sh = L.stream_hi;
    CKL(cudaStreamWaitEvent(sh, L.ev[7], 0));
    FCX_LAUNCH(k_cns_dp, (nb + CDP_WARPS - 1) / CDP_WARPS, CDP_WARPS * 32, 0, sh,
        L.d_blocks.as<BlockDesc>(), nb, L.d_vmeta.as<VoteMeta>(), L.d_slots.as<uint2>(), L.d_ovf.as<uint2>(),
        L.d_recs.as<CnsRec>(), L.d_lvl.as<int32_t>(), L.d_cns.as<char>(), L.d_eqv.as<int32_t>(),
        ctx->want_eqv ? 1 : 0, min_cov, L.d_cnsout.as<CnsOut>());
    CKL(cudaGetLastError());
    L.counters[FCX_C_KERNEL_LAUNCHES] += 1;
    CKL(cudaEventRecord(L.ev[6], sh));
    CKL(cudaMemcpyAsync(L.h_cnsout.p, L.d_cnsout.p, nb * sizeof(CnsOut), cudaMemcpyDeviceToHost, sh));
    CKL(cudaMemcpyAsync(L.h_cns.p, L.d_cns.p, cns_total, cudaMemcpyDeviceToHost, sh));
    if (np) CKL(cudaMemcpyAsync(L.h_aln.p, L.d_aln.p, (size_t)np * sizeof(PairAln), cudaMemcpyDeviceToHost, sh));
    if (ctx->want_eqv) {
        CKL(L.h_eqv.reserve(cns_total * 4));
        CKL(cudaMemcpyAsync(L.h_eqv.p, L.d_eqv.p, cns_total * 4, cudaMemcpyDeviceToHost, sh));
    }
    int vote_err = 0;
    CKL(cudaMemcpyAsync(&vote_err, L.d_counter.as<int>() + 5, 4, cudaMemcpyDeviceToHost, sh));
    tl[5] = now_ms();
    CKL(cudaStreamSynchronize(sh));
    tl[6] = now_ms();
    if (vote_err) {
        L.err = vote_err == 1 ? "consensus vote: more than 160 distinct links at one seed position"
                              : "consensus vote: link overflow arena exhausted";
        return vote_err == 2 ? 101 : 3;
    }

    // ---- collect
    const CnsOut* co = L.h_cnsout.as<CnsOut>();
    const char* hc = L.h_cns.as<char>();
    uint64_t out_total = 0;
    for (uint32_t b = 0; b < nb; b++) out_total += (uint64_t)std::max(co[b].len, 0);
    res.bases.reserve(out_total); res.lens.reserve(nb);
    for (uint32_t b = 0; b < nb; b++) {
        if (co[b].err) {
            char buf[256];
            snprintf(buf, sizeof buf, "consensus kernel error %d in block %u (2: record overflow, 3: no best score (reference asserts, falcon.c:476))",
                     co[b].err, b0 + b);
            L.err = buf; return co[b].err == 2 ? 102 : 3;
        }
        L.prof[1] += co[b].positions;
        const uint64_t at = hb[b].cns_off + (uint64_t)co[b].start;       // written back to front (k_cns_dp)
        res.bases.insert(res.bases.end(), hc + at, hc + at + co[b].len);
        res.lens.push_back((uint64_t)co[b].len);
        if (ctx->want_eqv) {
            const int32_t* he = L.h_eqv.as<int32_t>();
            res.eqv.insert(res.eqv.end(), he + at, he + at + co[b].len);
        }
    }
    const PairAln* hal = L.h_aln.as<PairAln>();
    uint64_t cells = 0, steps = 0, cols = 0, accepted = 0;
    if (ctx->keep_pair_info) res.info.reserve(np);
    for (uint32_t p = 0; p < np; p++) {
        if (hal[p].accepted < 0) { L.err = "internal error: accepted alignment beyond its trace bound"; return 4; }
        cells += (uint64_t)hal[p].cells; accepted += hal[p].accepted;
        if (hal[p].aligned) steps += (uint64_t)hal[p].dist + 1;
        if (hal[p].accepted) cols += (uint64_t)hal[p].aln_size;
        if (ctx->keep_pair_info) {
            fcx_pair_info pi;
            pi.n_match = hr[p].n_match; pi.s1 = hr[p].s1; pi.e1 = hr[p].e1; pi.s2 = hr[p].s2; pi.e2 = hr[p].e2;
            pi.passed_filter = hr[p].pass; pi.aligned = hal[p].aligned; pi.dist = hal[p].dist;
            pi.aln_size = hal[p].aln_size; pi.q_e = hal[p].q_e; pi.t_e = hal[p].t_e;
            pi.accepted = hal[p].accepted; pi.n_tags = hal[p].n_tags; pi.trace_cells = hal[p].cells;
            res.info.push_back(pi);
        }
    }
    L.counters[FCX_C_PAIRS] += np; L.counters[FCX_C_DP_PAIRS] += dp_pairs;
    L.counters[FCX_C_ACCEPTED] += accepted; L.counters[FCX_C_TRACE_CELLS] += cells;
    L.counters[FCX_C_DP_STEPS] += steps; L.counters[FCX_C_ALN_COLS] += cols;
    L.counters[FCX_C_SPAN_BASES] += span_bases; L.counters[FCX_C_WAVES] += 1;
    tl[7] = now_ms();
    if (trace_waves)
        fprintf(stderr, "wave lane %d blocks %u..%u pairs %u: t0 %.1f prep %.1f launch1 %.1f wait1 %.1f alloc %.1f launch2 %.1f wait2 %.1f collect %.1f\n",
                (int)(&L - ctx->lanes.data()), b0, b1, np, tl[0], tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3],
                tl[5] - tl[4], tl[6] - tl[5], tl[7] - tl[6]);
    float ms;
    cudaEventElapsedTime(&ms, L.ev[0], L.ev[1]); L.times[FCX_T_INDEX] += ms;
    cudaEventElapsedTime(&ms, L.ev[1], L.ev[2]); L.times[FCX_T_RANGE] += ms;
    cudaEventElapsedTime(&ms, L.ev[3], L.ev[4]); L.times[FCX_T_DP] += ms;
    cudaEventElapsedTime(&ms, L.ev[4], L.ev[5]); L.times[FCX_T_TRACEBACK] += ms;
    cudaEventElapsedTime(&ms, L.ev[5], L.ev[6]); L.times[FCX_T_CONSENSUS] += ms;
    cudaEventElapsedTime(&ms, L.ev[0], L.ev[6]); L.times[FCX_T_TOTAL] += ms;
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
        for (uint32_t i = block_off[b]; i < block_off[b + 1]; i++)
            if (ctx->h_len[read_ids[i]] > 100000) { ctx->err = "read longer than 100000 bases in a seed block (the reference truncates at consensus.py:178-179)"; return 1; }
        if (ctx->h_len[read_ids[block_off[b]]] >= 100000) { ctx->err = "seed of 100000 bases or more (the reference asserts t_len < 100000, falcon.c:343)"; return 1; }
    }
    // ---- plan waves from upper bounds (exact sizes are computed per wave after k_range).
    // Aim for >= 2 waves per lane so that stages of different waves overlap, but keep waves large
    // enough (min_wave_blocks) for the one-warp-per-block consensus kernel to fill the GPU.
    const int nl = ctx->active_lanes > 0 ? ctx->active_lanes : (int)ctx->lanes.size();
    // equal-sized waves: at least one per lane, none larger than max_wave_blocks (a short last wave
    // would still pay the full, latency-bound consensus kernel)
    const int wpl = getenv("FCX_WAVES_PER_LANE") ? std::max(1, atoi(getenv("FCX_WAVES_PER_LANE"))) : 1;
    uint32_t n_waves = std::max<uint32_t>((uint32_t)(wpl * nl), (n_blocks + ctx->max_wave_blocks - 1) / ctx->max_wave_blocks);
    n_waves = ((n_waves + nl - 1) / nl) * nl;                       // a multiple of the lane count
    uint32_t target = (n_blocks + n_waves - 1) / std::max(1u, n_waves);
    target = std::max(target, ctx->min_wave_blocks);
    target = std::min(target, ctx->max_wave_blocks);
    double budget = (double)ctx->arena_budget / nl;
    if (ctx->dp_variant == 3) {              // k_dp3's trace scratch: (resident warps) x (longest trace of the call, upper bound)
        int max_s = 1, max_r = 1;
        for (uint32_t b = 0; b < n_blocks; b++) {
            max_s = std::max(max_s, ctx->h_len[read_ids[block_off[b]]]);
            for (uint32_t i = block_off[b] + 1; i < block_off[b + 1]; i++) max_r = std::max(max_r, ctx->h_len[read_ids[i]]);
        }
        budget -= (double)ctx->sm_count * 32 * DP3_WARPS * (0.3 * ((double)max_s + max_r) + 2) * TRACE_REC_WORDS * 4;
        budget = std::max(budget, 1e9);
    }
    std::vector<std::pair<uint32_t, uint32_t>> waves;
    auto pack = [&](uint32_t tgt) {
        waves.clear();
        for (uint32_t b = 0; b < n_blocks;) {
            uint32_t e = b; uint64_t pairs = 0; double bytes = 0;
            while (e < n_blocks) {
                uint32_t lo = block_off[e], hi = block_off[e + 1];
                int slen = ctx->h_len[read_ids[lo]];
                const double mdiff = std::max(0.0, 1.0 - min_idt);
                const double capfrac = std::min(0.3, mdiff < 1.999 ? mdiff / (2.0 - mdiff) : 0.3);
                double bb = (double)KTAB * 4 + (double)slen * (4 + (8 + (hi - lo - 1) / 16) * 16 + 2 * 5 + 16.0 * VSLOT * 1.5) + 4.0 * CDP_LEVELS * 5 * 4;
                for (uint32_t i = lo + 1; i < hi; i++) {
                    // typical aligned span ~ 0.65 x the shorter sequence (exact sizes follow k_range;
                    // an under-estimate is caught by the out-of-memory split below)
                    const double span = 0.65 * std::min(ctx->h_len[read_ids[i]], slen);
                    bb += (ctx->dp_variant == 3 ? 0.0 : capfrac * 2.0 * span * 33.0) + 4.0 * (span + 8) + 128;      // (k_dp3: per-warp trace scratch, below)
                }
                if (e > b && (bytes + bb > budget || pairs + (hi - lo - 1) > ctx->max_wave_pairs || e - b >= tgt)) break;
                bytes += bb; pairs += hi - lo - 1; e++;
            }
            waves.emplace_back(b, e);
            b = e;
        }
    };
    pack(target);
    // The memory budget may have cut the waves shorter than planned, leaving a small tail wave or a
    // wave count that does not divide among the lanes: repack into equal waves, a multiple of the
    // lane count (a short wave still pays the full latency of the serial consensus kernel).
    if (waves.size() > 1) {
        const uint32_t nw = (uint32_t)(((waves.size() + nl - 1) / nl) * nl);
        const uint32_t tgt2 = std::max(ctx->min_wave_blocks, (n_blocks + nw - 1) / nw);
        if (tgt2 < target || waves.size() % nl) pack(std::min(tgt2, target));
    }
    for (auto& L : ctx->lanes) {
        memset(L.times, 0, sizeof L.times); memset(L.counters, 0, sizeof L.counters); memset(L.prof, 0, sizeof L.prof);
        L.err.clear();
    }
    std::vector<WaveResult> results(waves.size());
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    // a wave that does not fit (code 100) is split in two and retried; results stay in block order
    std::function<int(Lane&, uint32_t, uint32_t, WaveResult&)> run_split =
        [&](Lane& L, uint32_t b0, uint32_t b1, WaveResult& out) -> int {
        int rc = run_wave(ctx, L, b0, b1, block_off, read_ids, min_cov, min_idt, out);
        // capacity heuristics (vote overflow arena, consensus records) are retried with doubled
        // sizes: the reference reallocs, so a deep or noisy block must not fail the call
        for (int tries = 0; (rc == 101 || rc == 102) && tries < 6; tries++) {
            if (rc == 101) L.ovf_scale *= 2; else L.rec_scale *= 2;
            out = WaveResult(); L.err.clear();
            rc = run_wave(ctx, L, b0, b1, block_off, read_ids, min_cov, min_idt, out);
        }
        L.ovf_scale = L.rec_scale = 1;
        if (rc == 101 || rc == 102) return 3;
        if (rc != 100) return rc;
        if (b1 - b0 <= 1) { L.err = "a single seed block does not fit in device memory"; return 1; }
        out = WaveResult();
        const uint32_t mid = b0 + (b1 - b0) / 2;
        WaveResult a, b;
        if ((rc = run_split(L, b0, mid, a))) return rc;
        if ((rc = run_split(L, mid, b1, b))) return rc;
        out.bases = std::move(a.bases); out.bases.insert(out.bases.end(), b.bases.begin(), b.bases.end());
        out.lens = std::move(a.lens); out.lens.insert(out.lens.end(), b.lens.begin(), b.lens.end());
        out.info = std::move(a.info); out.info.insert(out.info.end(), b.info.begin(), b.info.end());
        out.eqv = std::move(a.eqv); out.eqv.insert(out.eqv.end(), b.eqv.begin(), b.eqv.end());
        return 0;
    };
    // results are appended to the call's output in block order by whichever lane thread completes
    // the next wave in line, while the other lanes keep the GPU busy: only the last wave's copy is serial
    {
        uint64_t upper = 0;
        for (uint32_t b = 0; b < n_blocks; b++) upper += (uint64_t)ctx->h_len[read_ids[block_off[b]]] * 2 + 8;
        ctx->out_bases.reserve(upper); ctx->out_off.reserve((size_t)n_blocks + 1);
    }
    std::mutex flush_mtx;
    std::vector<char> done(waves.size(), 0);
    size_t flushed = 0;
    auto flush_ready = [&]() {                       // caller holds flush_mtx
        while (flushed < waves.size() && done[flushed]) {
            WaveResult& r = results[flushed];
            ctx->out_bases.insert(ctx->out_bases.end(), r.bases.begin(), r.bases.end());
            for (uint64_t l : r.lens) ctx->out_off.push_back(ctx->out_off.back() + l);
            if (ctx->keep_pair_info) ctx->pair_info.insert(ctx->pair_info.end(), r.info.begin(), r.info.end());
            if (ctx->want_eqv) ctx->out_eqv.insert(ctx->out_eqv.end(), r.eqv.begin(), r.eqv.end());
            r = WaveResult();
            flushed++;
        }
    };
    auto worker = [&](int li) {
        cudaSetDevice(ctx->device);
        Lane& L = ctx->lanes[li];
        for (;;) {
            size_t w = next.fetch_add(1);
            if (w >= waves.size() || failed.load()) break;
            int rc = run_split(L, waves[w].first, waves[w].second, results[w]);
            if (rc) { failed.store(rc); break; }
            std::lock_guard<std::mutex> g(flush_mtx);
            done[w] = 1;
            flush_ready();
        }
    };
    const int nthreads = (int)std::min<size_t>(waves.size(), (size_t)nl);
    if (nthreads <= 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    if (failed.load()) {
        for (auto& L : ctx->lanes) if (!L.err.empty()) { ctx->err = L.err; break; }
        return failed.load();
    }
    for (auto& L : ctx->lanes) {
        for (int i = 0; i < FCX_T_COUNT; i++) ctx->times[i] += L.times[i];
        for (int i = 0; i < FCX_C_COUNT; i++) ctx->counters[i] += L.counters[i];
        for (int i = 0; i < 8; i++) ctx->prof[i] += L.prof[i];
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

// CUDA-event stopwatch (bench.py brackets its timed region with it).  The events are recorded on
// the engine's main stream; every fcx_* call is synchronous, so all lane work issued between start
// and stop has completed when stop is recorded.
extern "C" int fcx_timer_start(fcx_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(ctx->tev[0], ctx->stream));
    return 0;
}
extern "C" int fcx_timer_stop(fcx_ctx* ctx, double* ms) {
    CK(cudaDeviceSynchronize());
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

// ---------------------------------------------------------------------------------- --trim
// get_consensus_with_trim (falcon_kit/mains/consensus.py:123-158) for a batch of seed blocks: the
// per-read k-mer chaining runs on the device (k_trim_range), the scalar post-processing of
// get_alignment (:62-99) and the read selection (:131-147) are a few integer operations per read
// on the host, and the trimmed reads are cut out of the packed pool on the device (k_subreads) and
// APPENDED to the pool as new reads.  Returns the new block lists (seed first, then the trimmed
// reads, longest alignment first), ready for fcx_consensus_blocks.
extern "C" int fcx_trim_blocks(fcx_ctx* ctx, uint32_t n_blocks, const uint32_t* block_off, const uint32_t* read_ids,
                               int edge_tolerance, int trim_size, unsigned max_n_read, unsigned max_cov_aln,
                               const uint32_t** out_block_off, const uint32_t** out_read_ids, uint32_t* out_n_reads) {
    CK(cudaSetDevice(ctx->device));
    ctx->trim_block_off.assign(1, 0u); ctx->trim_read_ids.clear();
    for (uint32_t b = 0; b < n_blocks; b++) {
        if (block_off[b + 1] <= block_off[b]) { ctx->err = "empty block (a block needs at least the seed)"; return 1; }
        for (uint32_t i = block_off[b]; i < block_off[b + 1]; i++) {
            if (read_ids[i] >= ctx->n_reads) { ctx->err = "read id outside the uploaded pool"; return 1; }
            if (ctx->h_len[read_ids[i]] > 100000) { ctx->err = "read longer than 100000 bases in a seed block"; return 1; }
        }
    }
    Lane& L = ctx->lanes[0];
    cudaStream_t st = L.stream;
    struct Cut { uint32_t src; int32_t s, e; };
    std::vector<SubRead> subs; std::vector<uint64_t> sub_wb;
    std::vector<int32_t> new_len; std::vector<uint64_t> new_woff;
    uint64_t w_next = ctx->h_woff[ctx->n_reads], wb = 0;
    uint32_t next_id = ctx->n_reads;
    const uint32_t chunk = std::max(1u, ctx->max_wave_blocks);
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += chunk) {
        const uint32_t b1 = std::min(n_blocks, b0 + chunk), nb = b1 - b0;
        std::vector<BlockDesc> hb(nb);
        uint64_t npairs64 = 0, kpos_total = 0; int max_slen = 1, max_rlen = 1;
        for (uint32_t b = 0; b < nb; b++) {
            const uint32_t lo = block_off[b0 + b], hi = block_off[b0 + b + 1];
            BlockDesc& d = hb[b];
            memset(&d, 0, sizeof d);
            const uint32_t seed = read_ids[lo];
            d.seed_woff = ctx->h_woff[seed]; d.slen = ctx->h_len[seed];
            d.pair_begin = (uint32_t)npairs64; d.n_pairs = hi - lo - 1;
            d.kpos_off = kpos_total; kpos_total += (uint64_t)std::max(d.slen, 1);
            npairs64 += d.n_pairs; max_slen = std::max(max_slen, d.slen);
        }
        const uint32_t np = (uint32_t)npairs64;
        std::vector<PairDesc> hp(np);
        for (uint32_t b = 0; b < nb; b++)
            for (uint32_t j = 0; j < hb[b].n_pairs; j++) {
                const uint32_t rid = read_ids[block_off[b0 + b] + 1 + j];
                PairDesc& pd = hp[hb[b].pair_begin + j];
                pd.read_woff = ctx->h_woff[rid]; pd.block = b; pd.rlen = ctx->h_len[rid];
                max_rlen = std::max(max_rlen, pd.rlen);
            }
        std::vector<TrimOut> ho(np);
        if (np) {
            CK(L.d_blocks.reserve(nb * sizeof(BlockDesc)));
            CK(L.d_pairs.reserve((size_t)np * sizeof(PairDesc)));
            CK(L.d_ktab.reserve((size_t)nb * KTAB * 4));
            CK(L.d_kpos.reserve(kpos_total * 4 + 16));
            CK(L.d_kbits.reserve((size_t)nb * (KTAB / 32) * 4));
            CK(L.d_counter.reserve(64));
            CK(ctx->d_trim_out.reserve((size_t)np * sizeof(TrimOut)));
            // per-warp scratch: at most TRIM_MASK_TH hits per query k-mer
            const uint32_t list_cap = (uint32_t)(((size_t)TRIM_MASK_TH * ((size_t)max_rlen / 4 + 2) + 63) & ~(size_t)31);
            const uint32_t hist_cap = (uint32_t)(((size_t)max_rlen + max_slen + 72) & ~(size_t)31);
            const size_t per_warp = ((size_t)5 * list_cap + hist_cap) * 4;
            unsigned grid = std::min<unsigned>((np + TRIM_WARPS - 1) / TRIM_WARPS, (unsigned)ctx->sm_count * 4u);
            while (grid > 1 && (size_t)grid * TRIM_WARPS * per_warp > ((size_t)8 << 30)) grid /= 2;
            CK(ctx->d_trim_scratch.reserve((size_t)grid * TRIM_WARPS * per_warp));
            CK(cudaMemcpyAsync(L.d_blocks.p, hb.data(), nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(L.d_pairs.p, hp.data(), (size_t)np * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(L.d_ktab.p, 0, (size_t)nb * KTAB * 4, st));
            CK(cudaMemsetAsync(L.d_counter.p, 0, 64, st));
            FCX_LAUNCH(k_index, nb, 256, 0, st, L.d_blocks.as<BlockDesc>(), ctx->d_pool.as<uint32_t>(), L.d_ktab.as<uint32_t>(),
                       L.d_kpos.as<uint32_t>(), L.d_kbits.as<uint32_t>());
            FCX_LAUNCH(k_trim_range, grid, TRIM_WARPS * 32, 0, st, L.d_blocks.as<BlockDesc>(), L.d_pairs.as<PairDesc>(), np,
                       ctx->d_pool.as<uint32_t>(), L.d_ktab.as<uint32_t>(), L.d_kpos.as<uint32_t>(),
                       ctx->d_trim_scratch.as<uint32_t>(), list_cap, hist_cap, L.d_counter.as<uint32_t>(),
                       ctx->d_trim_out.as<TrimOut>());
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(ho.data(), ctx->d_trim_out.p, (size_t)np * sizeof(TrimOut), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        // ---- get_alignment's scalar tail (consensus.py:62-99) and the selection (:131-147)
        for (uint32_t b = 0; b < nb; b++) {
            const uint32_t lo = block_off[b0 + b];
            const uint32_t seed = read_ids[lo];
            const int len_0 = ctx->h_len[seed];
            std::vector<Cut> cuts;
            for (uint32_t j = 0; j < hb[b].n_pairs; j++) {
                const TrimOut& t = ho[hb[b].pair_begin + j];
                const uint32_t rid = read_ids[lo + 1 + j];
                const int len_1 = ctx->h_len[rid];
                int s1 = t.s1, e1 = t.e1 + KMER + KMER / 2, s0 = t.s2, e0 = t.e2 + KMER + KMER / 2;
                e1 = std::min(e1, len_1); e0 = std::min(e0, len_0);
                int aln_size = 1; long aln_score = 0;
                if (e1 - s1 > 500) { aln_size = std::max(e1 - s1, e0 - s0); aln_score = (long)t.score * 48; }
                if (s1 > edge_tolerance && s0 > edge_tolerance) continue;
                if (len_1 - e1 > edge_tolerance && len_0 - e0 > edge_tolerance) continue;
                if (!(e1 - s1 > 500 && aln_size > 500)) continue;
                if (aln_score > 1000 && e1 - s1 > 500) cuts.push_back(Cut{rid, s1 + trim_size, e1 - trim_size});
            }
            std::stable_sort(cuts.begin(), cuts.end(), [](const Cut& a, const Cut& c) { return a.e - a.s > c.e - c.s; });   // longest first
            size_t keep = cuts.size();
            if (cuts.size() > max_n_read) {                       // get_longest_reads(.., sort=False), consensus.py:26-45
                size_t longest = max_n_read;
                if (max_cov_aln > 0) {
                    longest = 1; long read_cov = 0;
                    for (const Cut& c : cuts) {
                        if (len_0 > 0 && read_cov / len_0 > (long)max_cov_aln) break;
                        longest++; read_cov += std::max(0, c.e - c.s);
                    }
                    longest = std::min<size_t>(longest, max_n_read);
                }
                keep = longest > 0 ? longest - 1 : 0;               // seqs[:longest] includes the seed
            }
            ctx->trim_read_ids.push_back(seed);
            for (size_t c = 0; c < keep; c++) {
                const int len = std::max(0, cuts[c].e - cuts[c].s);   // Python slice semantics: empty if e <= s
                SubRead sr; sr.src_woff = ctx->h_woff[cuts[c].src]; sr.dst_woff = w_next; sr.s = len ? cuts[c].s : 0; sr.len = len;
                const uint64_t words = (((uint64_t)len + 15) / 16 + 1 + 3) & ~(uint64_t)3;
                subs.push_back(sr); sub_wb.push_back(wb);
                new_len.push_back(len); new_woff.push_back(w_next);
                w_next += words; wb += words;
                ctx->trim_read_ids.push_back(next_id++);
            }
            ctx->trim_block_off.push_back((uint32_t)ctx->trim_read_ids.size());
        }
    }
    // ---- append the trimmed reads to the pool
    if (!subs.empty()) {
        const size_t need = (w_next + 4) * 4;
        if (need > ctx->d_pool.cap) {                               // grow, keeping the packed reads
            DevBuf bigger;
            CK(bigger.reserve(need + need / 4));
            CK(cudaMemcpyAsync(bigger.p, ctx->d_pool.p, (ctx->h_woff[ctx->n_reads] + 4) * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));
            ctx->d_pool.release();
            ctx->d_pool = bigger;
        }
        CK(ctx->d_subs.reserve(subs.size() * sizeof(SubRead)));
        CK(ctx->d_sub_wb.reserve((sub_wb.size() + 1) * 8));
        CK(cudaMemcpyAsync(ctx->d_subs.p, subs.data(), subs.size() * sizeof(SubRead), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_sub_wb.p, sub_wb.data(), sub_wb.size() * 8, cudaMemcpyHostToDevice, st));
        FCX_LAUNCH(k_subreads, (unsigned)((wb + 255) / 256), 256, 0, st, ctx->d_subs.as<SubRead>(), (uint32_t)subs.size(),
                   ctx->d_sub_wb.as<uint64_t>(), wb, ctx->d_pool.as<uint32_t>());
        CK(cudaGetLastError());
        CK(cudaMemsetAsync((char*)ctx->d_pool.p + w_next * 4, 0, 16, st));
        CK(cudaStreamSynchronize(st));
        ctx->h_woff.pop_back();
        for (size_t i = 0; i < subs.size(); i++) { ctx->h_woff.push_back(new_woff[i]); ctx->h_len.push_back(new_len[i]); }
        ctx->h_woff.push_back(w_next);
        ctx->n_reads = ctx->n_reserved = next_id;
    }
    *out_block_off = ctx->trim_block_off.data();
    *out_read_ids = ctx->trim_read_ids.data();
    if (out_n_reads) *out_n_reads = ctx->n_reads;
    return 0;
}

// Drop the reads appended by fcx_trim_blocks (or any tail of the pool): the pool keeps its first n reads.
extern "C" int fcx_pool_truncate(fcx_ctx* ctx, uint32_t n_reads) {
    if (n_reads > ctx->n_reads) { ctx->err = "fcx_pool_truncate: the pool holds fewer reads"; return 1; }
    ctx->h_woff.resize((size_t)n_reads + 1);
    ctx->h_len.resize(n_reads);
    ctx->n_reads = ctx->n_reserved = n_reads;
    return 0;
}

// Batched banded alignment of pool sequences (distance only), for stage-2 style callers:
// falcon_kit/mains/graph_to_contig.py:50-103 calls DWA.align(q[s1:e1], e1-s1, t[s2:e2], e2-s2, 1500, 1)
// per contig pair and keeps aln_str_size and dist.
extern "C" int fcx_align_pairs(fcx_ctx* ctx, uint32_t n, const uint32_t* q_ids, const uint32_t* t_ids,
                               const int32_t* ranges, int band_tolerance, fcx_align_result* out) {
    CK(cudaSetDevice(ctx->device));
    if (band_tolerance < 0 || band_tolerance * 2 + 4 > AL_VRING) { ctx->err = "fcx_align_pairs: band_tolerance out of range (<= 4094)"; return 1; }
    if (n == 0) return 0;
    std::vector<AlignJob> jobs(n);
    for (uint32_t i = 0; i < n; i++) {
        if (q_ids[i] >= ctx->n_reads || t_ids[i] >= ctx->n_reads) { ctx->err = "fcx_align_pairs: sequence id outside the uploaded pool"; return 1; }
        const int ql = ctx->h_len[q_ids[i]], tl = ctx->h_len[t_ids[i]];
        int s1 = 0, e1 = ql, s2 = 0, e2 = tl;
        if (ranges) { s1 = ranges[4 * i]; e1 = ranges[4 * i + 1]; s2 = ranges[4 * i + 2]; e2 = ranges[4 * i + 3]; }
        if (s1 < 0 || e1 < s1 || e1 > ql || s2 < 0 || e2 < s2 || e2 > tl) { ctx->err = "fcx_align_pairs: range outside its sequence"; return 1; }
        const long long max_d = (long long)(int)(0.3 * ((e1 - s1) + (e2 - s2)));
        if ((long long)INT_MAX < max_d * (long long)(band_tolerance * 2 + 1) * 2LL) {   // DW_banded.c:158-161
            ctx->err = "fcx_align_pairs: lens are too big (the reference aborts here, DW_banded.c:158-161)"; return 1;
        }
        jobs[i].q_woff = ctx->h_woff[q_ids[i]]; jobs[i].t_woff = ctx->h_woff[t_ids[i]];
        jobs[i].qs = s1; jobs[i].q_len = e1 - s1; jobs[i].ts = s2; jobs[i].t_len = e2 - s2;
    }
    cudaStream_t st = ctx->stream;
    CK(ctx->d_trace1.reserve((size_t)n * sizeof(AlignJob)));
    CK(ctx->d_aln1.reserve((size_t)n * sizeof(PairAln)));
    CK(cudaMemcpyAsync(ctx->d_trace1.p, jobs.data(), (size_t)n * sizeof(AlignJob), cudaMemcpyHostToDevice, st));
    FCX_LAUNCH(k_align_batch, std::min<unsigned>(n, (unsigned)ctx->sm_count * 6u), 32, 0, st, ctx->d_pool.as<uint32_t>(),
               ctx->d_trace1.as<AlignJob>(), n, band_tolerance, ctx->d_aln1.as<PairAln>());
    CK(cudaGetLastError());
    std::vector<PairAln> res(n);
    CK(cudaMemcpyAsync(res.data(), ctx->d_aln1.p, (size_t)n * sizeof(PairAln), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n; i++) {
        out[i].aln_str_size = res[i].aligned ? res[i].aln_size : 0;
        out[i].dist = res[i].aligned ? res[i].dist : 0;
        out[i].aln_q_e = res[i].aligned ? res[i].q_e : 0;
        out[i].aln_t_e = res[i].aligned ? res[i].t_e : 0;
    }
    return 0;
}

// single-pair align(): see k_align1 / k_align1_tb
extern "C" int fcx_internal_align(fcx_ctx* ctx, const char* q, int q_len, const char* t, int t_len,
                                  int band_tolerance, int get_aln_str, alignment* out) {
    CK(cudaSetDevice(ctx->device));
    if (q_len < 0 || t_len < 0 || band_tolerance < 0) { ctx->err = "align(): negative length"; return 1; }
    if (band_tolerance * 2 + 4 > AL_VRING) { ctx->err = "align(): band_tolerance too large for this build"; return 1; }
    const long long max_d = (long long)(int)(0.3 * (q_len + t_len));
    if ((long long)INT_MAX < max_d * (long long)(band_tolerance * 2 + 1) * 2LL) {   // DW_banded.c:158-161
        ctx->err = "align(): lens are too big (the reference aborts here, DW_banded.c:158-161)"; return 1;
    }
    std::vector<char> cat((size_t)q_len + t_len + 1);
    memcpy(cat.data(), q, q_len); memcpy(cat.data() + q_len, t, t_len);
    uint64_t off[3] = {0, (uint64_t)q_len, (uint64_t)q_len + t_len};
    if (int rc = fcx_pool_upload(ctx, cat.data(), off, 2)) return rc;
    const int rec_words = 1 + (band_tolerance + 1 + 31) / 32 + 1;
    cudaStream_t st = ctx->stream;
    CK(ctx->d_trace1.reserve((size_t)(max_d + 1) * rec_words * 4 + 64));
    CK(ctx->d_path1.reserve((size_t)(max_d / 32 + 2) * 4));
    CK(ctx->d_aln1.reserve(sizeof(PairAln)));
    CK(ctx->d_str1.reserve(2 * ((size_t)q_len + t_len + 2)));
    CK(cudaMemsetAsync(ctx->d_path1.p, 0, (size_t)(max_d / 32 + 2) * 4, st));
    const uint32_t* pool = ctx->d_pool.as<uint32_t>();
    FCX_LAUNCH(k_align1, 1, 32, 0, st, pool, ctx->h_woff[0], ctx->h_woff[1], q_len, t_len, band_tolerance,
                               ctx->d_trace1.as<uint32_t>(), rec_words, ctx->d_aln1.as<PairAln>());
    CK(cudaGetLastError());
    char* dq = ctx->d_str1.as<char>(); char* dt = dq + (size_t)q_len + t_len + 2;
    if (get_aln_str > 0) {
        FCX_LAUNCH(k_align1_tb, 1, 1, 0, st, pool, ctx->h_woff[0], ctx->h_woff[1], q_len, t_len, ctx->d_trace1.as<uint32_t>(),
                                     rec_words, ctx->d_path1.as<uint32_t>(), ctx->d_aln1.as<PairAln>(), dq, dt);
        CK(cudaGetLastError());
    }
    PairAln a;
    CK(cudaMemcpyAsync(&a, ctx->d_aln1.p, sizeof a, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out->aln_str_size = 0; out->dist = 0; out->aln_q_s = out->aln_q_e = out->aln_t_s = out->aln_t_e = 0;
    if (a.aligned) {
        out->aln_str_size = a.aln_size; out->dist = a.dist; out->aln_q_e = a.q_e; out->aln_t_e = a.t_e;
        if (get_aln_str > 0 && a.aln_size > 0) {
            CK(cudaMemcpy(out->q_aln_str, dq, (size_t)a.aln_size, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(out->t_aln_str, dt, (size_t)a.aln_size, cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}
