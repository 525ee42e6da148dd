// fcx_multi.cu -- one process, several GPUs: a set of engines that share ONE read
// store and ONE seed-block list.
//
//   pool upload   device d uploads and packs its 1/N of the reads from host memory (N concurrent H2D
//                 copies + k_pack), then every other device receives that part by a peer copy over
//                 NVLink straight into place -- the "one broadcast of the read index": afterwards every
//                 GPU holds the whole 2-bit read store (seed blocks reference reads anywhere in it);
//   consensus     the seed blocks are cut into N contiguous slices of near-equal cost
//                 (pairs x seed length), slice d runs on device d in its own host thread (each engine
//                 still pipelines its waves on its lanes), and the results are concatenated in seed
//                 order -- the ordering contract of exe_pool.imap (falcon_kit/mains/consensus.py:274).
// There is no cross-GPU traffic after the pool upload.  Process-per-GPU jobs (torchrun) use the same
// three-step pool API with an NCCL broadcast instead of the peer copies (bench.py).
#include "../../include/falcon_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct fcx_multi {
    std::vector<int> devices;
    std::vector<fcx_ctx*> eng;
    std::string err;
    std::vector<int32_t> read_len;
    std::vector<char> out_bases;
    std::vector<uint64_t> out_off;
    std::vector<fcx_pair_info> pair_info;
    bool keep_pair_info = true;
    double times[FCX_T_COUNT] = {0};
    uint64_t counters[FCX_C_COUNT] = {0};
    uint64_t peer_bytes = 0;
};

static thread_local std::string g_multi_err;

extern "C" const char* fcx_multi_last_error(const fcx_multi* m) { return m ? m->err.c_str() : g_multi_err.c_str(); }

extern "C" int fcx_multi_create(const int* devices, int n_devices, fcx_multi** out) {
    *out = nullptr;
    if (n_devices <= 0) { g_multi_err = "fcx_multi_create: no devices"; return 1; }
    fcx_multi* m = new fcx_multi();
    for (int i = 0; i < n_devices; i++) {
        fcx_ctx* e = nullptr;
        if (fcx_create(devices[i], &e) != 0) {
            g_multi_err = std::string("fcx_multi_create: device ") + std::to_string(devices[i]) + ": " + fcx_last_error(nullptr);
            for (auto* x : m->eng) fcx_destroy(x);
            delete m; return 1;
        }
        m->devices.push_back(devices[i]); m->eng.push_back(e);
    }
    // peer access for the pool distribution (without it cudaMemcpyPeerAsync stages through the host)
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < n_devices; j++) {
            if (i == j) continue;
            int ok = 0;
            if (cudaDeviceCanAccessPeer(&ok, devices[i], devices[j]) == cudaSuccess && ok) {
                cudaSetDevice(devices[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess) cudaGetLastError();      // already enabled: fine
            }
        }
    *out = m;
    return 0;
}

extern "C" void fcx_multi_destroy(fcx_multi* m) {
    if (!m) return;
    for (auto* e : m->eng) fcx_destroy(e);
    delete m;
}

extern "C" int fcx_multi_device_count(const fcx_multi* m) { return (int)m->eng.size(); }

extern "C" int fcx_multi_set_option(fcx_multi* m, const char* name, double value) {
    if (std::string(name) == "pair_info") m->keep_pair_info = value != 0;
    for (auto* e : m->eng)
        if (fcx_set_option(e, name, value)) { m->err = fcx_last_error(e); return 1; }
    return 0;
}

extern "C" int fcx_multi_pool_upload(fcx_multi* m, const char* bases, const uint64_t* offsets, uint32_t n_reads) {
    const int N = (int)m->eng.size();
    m->read_len.assign(n_reads, 0);
    for (uint32_t r = 0; r < n_reads; r++) m->read_len[r] = (int32_t)(offsets[r + 1] - offsets[r]);
    // layout on every device, then device d packs the reads of part d (cut by bases, not by count)
    std::vector<uint32_t> cut(N + 1, n_reads);
    cut[0] = 0;
    {
        const uint64_t total = offsets[n_reads] - offsets[0];
        uint32_t r = 0;
        for (int d = 1; d < N; d++) {
            const uint64_t target = offsets[0] + total * (uint64_t)d / (uint64_t)N;
            while (r < n_reads && offsets[r] < target) r++;
            cut[d] = r;
        }
    }
    std::vector<int> rc(N, 0);
    {
        std::vector<std::thread> th;
        for (int d = 0; d < N; d++)
            th.emplace_back([&, d]() {
                rc[d] = fcx_pool_reserve(m->eng[d], offsets, n_reads, nullptr);
                if (!rc[d] && cut[d + 1] > cut[d])
                    rc[d] = fcx_pool_upload_part(m->eng[d], bases, offsets + cut[d], cut[d], cut[d + 1] - cut[d]);
            });
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < N; d++) if (rc[d]) { m->err = fcx_last_error(m->eng[d]); return rc[d]; }
    // every device receives the other parts by peer copies (device-to-device, NVLink)
    m->peer_bytes = 0;
    if (N > 1) {
        std::vector<void*> ptr(N, nullptr);
        const uint64_t* woff = nullptr;
        for (int d = 0; d < N; d++) {
            uint64_t nw = 0;
            if (fcx_pool_device(m->eng[d], &ptr[d], &nw, &woff)) { m->err = fcx_last_error(m->eng[d]); return 1; }
        }
        std::vector<cudaStream_t> st(N, nullptr);
        for (int d = 0; d < N; d++) { cudaSetDevice(m->devices[d]); if (cudaStreamCreate(&st[d]) != cudaSuccess) { m->err = "cudaStreamCreate failed"; return 1; } }
        cudaError_t ce = cudaSuccess;
        for (int dst = 0; dst < N && ce == cudaSuccess; dst++) {
            cudaSetDevice(m->devices[dst]);
            for (int src = 0; src < N && ce == cudaSuccess; src++) {
                if (src == dst) continue;
                const uint64_t w0 = woff[cut[src]], w1 = woff[cut[src + 1]];
                if (w1 == w0) continue;
                ce = cudaMemcpyPeerAsync((char*)ptr[dst] + w0 * 4, m->devices[dst], (const char*)ptr[src] + w0 * 4,
                                         m->devices[src], (w1 - w0) * 4, st[dst]);
                m->peer_bytes += (w1 - w0) * 4;
            }
        }
        for (int d = 0; d < N; d++) {
            cudaSetDevice(m->devices[d]);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st[d]);
            cudaStreamDestroy(st[d]);
        }
        if (ce != cudaSuccess) { m->err = std::string("peer copy of the read store failed: ") + cudaGetErrorString(ce); return 1; }
    }
    for (int d = 0; d < N; d++) if (fcx_pool_commit(m->eng[d])) { m->err = fcx_last_error(m->eng[d]); return 1; }
    return 0;
}

extern "C" uint64_t fcx_multi_peer_bytes(const fcx_multi* m) { return m->peer_bytes; }

extern "C" int fcx_multi_consensus_blocks(fcx_multi* m, uint32_t n_blocks, const uint32_t* block_off,
                                          const uint32_t* read_ids, unsigned min_cov, unsigned K, double min_idt,
                                          const char** out_bases, const uint64_t** out_off) {
    const int N = (int)m->eng.size();
    // contiguous slices of near-equal cost = pairs x seed length
    std::vector<double> cum((size_t)n_blocks + 1, 0.0);
    for (uint32_t b = 0; b < n_blocks; b++) {
        const uint32_t lo = block_off[b], hi = block_off[b + 1];
        if (hi <= lo) { m->err = "empty block (a block needs at least the seed)"; return 1; }
        if (read_ids[lo] >= m->read_len.size()) { m->err = "read id outside the uploaded pool"; return 1; }
        cum[b + 1] = cum[b] + (double)(hi - lo) * (double)std::max(1, m->read_len[read_ids[lo]]);
    }
    std::vector<uint32_t> cut(N + 1, n_blocks);
    cut[0] = 0;
    for (int d = 1; d < N; d++) {
        const double target = cum[n_blocks] * d / N;
        uint32_t k = (uint32_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        cut[d] = std::min(n_blocks, std::max(k, cut[d - 1]));
    }
    std::vector<int> rc(N, 0);
    std::vector<const char*> ob(N, nullptr);
    std::vector<const uint64_t*> oo(N, nullptr);
    std::vector<std::vector<uint32_t>> boff(N);
    {
        std::vector<std::thread> th;
        for (int d = 0; d < N; d++)
            th.emplace_back([&, d]() {
                const uint32_t b0 = cut[d], b1 = cut[d + 1];
                boff[d].resize((size_t)(b1 - b0) + 1);
                for (uint32_t b = b0; b <= b1; b++) boff[d][b - b0] = block_off[b] - block_off[b0];
                rc[d] = fcx_consensus_blocks(m->eng[d], b1 - b0, boff[d].data(), read_ids + block_off[b0], min_cov, K,
                                             min_idt, &ob[d], &oo[d]);
            });
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < N; d++) if (rc[d]) { m->err = std::string("device ") + std::to_string(m->devices[d]) + ": " + fcx_last_error(m->eng[d]); return rc[d]; }
    // merge in seed order
    m->out_bases.clear(); m->out_off.assign(1, 0); m->pair_info.clear();
    memset(m->times, 0, sizeof m->times); memset(m->counters, 0, sizeof m->counters);
    for (int d = 0; d < N; d++) {
        const uint32_t nb = cut[d + 1] - cut[d];
        m->out_bases.insert(m->out_bases.end(), ob[d], ob[d] + oo[d][nb]);
        for (uint32_t b = 0; b < nb; b++) m->out_off.push_back(m->out_off.back() + (oo[d][b + 1] - oo[d][b]));
        if (m->keep_pair_info) {
            uint64_t np = 0;
            fcx_last_pair_info(m->eng[d], nullptr, 0, &np);
            const size_t at = m->pair_info.size();
            m->pair_info.resize(at + np);
            fcx_last_pair_info(m->eng[d], m->pair_info.data() + at, np, &np);
        }
        double t[FCX_T_COUNT]; uint64_t c[FCX_C_COUNT];
        fcx_last_stats(m->eng[d], t, c);
        for (int i = 0; i < FCX_T_COUNT; i++) m->times[i] = std::max(m->times[i], t[i]);
        for (int i = 0; i < FCX_C_COUNT; i++) m->counters[i] += c[i];
    }
    *out_bases = m->out_bases.data();
    *out_off = m->out_off.data();
    return 0;
}

extern "C" int fcx_multi_last_pair_info(fcx_multi* m, fcx_pair_info* out, uint64_t max_pairs, uint64_t* n_pairs) {
    const uint64_t n = m->pair_info.size();
    if (n_pairs) *n_pairs = n;
    if (out) memcpy(out, m->pair_info.data(), std::min(n, max_pairs) * sizeof(fcx_pair_info));
    return 0;
}

extern "C" int fcx_multi_last_stats(fcx_multi* m, double* times_ms, uint64_t* counters) {
    if (times_ms) memcpy(times_ms, m->times, sizeof m->times);
    if (counters) memcpy(counters, m->counters, sizeof m->counters);
    return 0;
}
