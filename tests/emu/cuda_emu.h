// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a minimal SIMT emulator that lets the CUDA sources of
// falcon_b200/csrc compile with g++ and run on the authoring container (which has no GPU), so
// that kernel logic can be checked against the oracle before GPU minutes are spent.
//
//   * every CUDA thread of a CTA is a fiber (hand-rolled x86-64 context switch) on ONE OS thread;
//     warp collectives (__shfl_sync, __ballot_sync, __reduce_*_sync, __match_any_sync, __syncwarp)
//     and __syncthreads are rendezvous points: a fiber yields until all named lanes arrived.
//     A collective that can never complete (a lane exited or took another branch) is reported as
//     a deadlock -- it would be undefined behaviour on the GPU as well;
//   * CTAs of one launch are spread over a few OS threads; __shared__ is `static thread_local`;
//   * the CUDA runtime calls the engine makes are shimmed onto malloc / memcpy, launches run
//     synchronously in issue order (which satisfies every stream / event dependency).
//
// The product library (libfalcon_b200.so, built by nvcc) never sees this file: FCX_EMU is defined
// only by tests/emu/Makefile, which builds tests/emu/_build/libfalcon_b200_emu.so, and only the
// `emu`-marked tests load that.  It is NOT a CPU fallback: falcon_b200/binding.py cannot load it.
#pragma once
#ifndef FCX_EMU
#error "cuda_emu.h is only for the FCX_EMU test build"
#endif

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) alignas(n)
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

namespace emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; Dim3() {} Dim3(unsigned x_) : x(x_) {} };

constexpr int MAX_GROUPS = 24;

struct Group {                 // one set of lanes that synchronise with a given mask
    unsigned mask = 0;
    uint64_t slot[2][32];
    uint32_t arrived[2] = {0, 0};
};
struct Warp {
    Group g[MAX_GROUPS];
    int n_groups = 0;
};
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    Dim3 tidx;
    bool done = false;
    uint32_t gen[MAX_GROUPS];
    uint32_t bar_gen = 0;
};
struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    Dim3 bidx, bdim, gdim;
    void* sched_sp = nullptr;
    const std::function<void()>* fn = nullptr;
    char* dyn_smem = nullptr;
    uint32_t bar_arrived = 0;
    bool progress = false;
    size_t n_threads = 0;
};
extern thread_local Cta* g_cta;
extern thread_local Fiber* g_cur;

void yield();                                   // back to the CTA scheduler
void launch(Dim3 grid, Dim3 block, size_t smem, const std::function<void()>& fn);
[[noreturn]] void fail(const char* what);

inline unsigned lane_id() { return g_cur->tidx.x & 31u; }
inline Warp& cur_warp() { return g_cta->warps[g_cur->tidx.x >> 5]; }

// rendezvous of the lanes in `mask`; every lane deposits `v`; returns the group and the phase
// holding everybody's deposits
inline Group& rendezvous(unsigned mask, uint64_t v, int& ph) {
    const unsigned lane = lane_id();
    if (!(mask >> lane & 1u)) fail("collective: calling lane is not in the mask");
    Warp& w = cur_warp();
    int gi = -1;
    for (int i = 0; i < w.n_groups; i++) if (w.g[i].mask == mask) { gi = i; break; }
    if (gi < 0) {
        if (w.n_groups == MAX_GROUPS) fail("collective: too many distinct masks in one warp");
        gi = w.n_groups++;
        w.g[gi].mask = mask; w.g[gi].arrived[0] = w.g[gi].arrived[1] = 0;
        // a group created mid-kernel: every lane starts it at generation 0
    }
    Group& g = w.g[gi];
    const uint32_t k = g_cur->gen[gi]++;
    ph = (int)(k & 1u);
    g.slot[ph][lane] = v;
    g.arrived[ph]++;
    const uint32_t want = (uint32_t)__builtin_popcount(mask) * (k / 2 + 1);
    while (g.arrived[ph] < want) yield();
    g_cta->progress = true;
    return g;
}

}  // namespace emu

#define threadIdx (emu::g_cur->tidx)
#define blockIdx (emu::g_cta->bidx)
#define blockDim (emu::g_cta->bdim)
#define gridDim (emu::g_cta->gdim)
#define warpSize 32

// ---------------------------------------------------------------------------------- intrinsics
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    s &= 31u; return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31u; return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
static inline long long clock64() { return 0; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicMin(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> static inline T atomicMax(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> static inline T atomicCAS(T* p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp;
}

// ---- warp collectives
namespace emu {
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { int ph; emu::rendezvous(mask, 0, ph); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int ph; emu::Group& g = emu::rendezvous(mask, emu::to_bits(v), ph);
    const unsigned lane = emu::lane_id();
    const unsigned s = (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
    if (!(mask >> s & 1u)) return v;      // reading a lane outside the mask: undefined on the GPU
    return emu::from_bits<T>(g.slot[ph][s]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int ph; emu::Group& g = emu::rendezvous(mask, emu::to_bits(v), ph);
    const unsigned lane = emu::lane_id();
    const unsigned base = lane & ~(unsigned)(width - 1);
    if (lane - base < d) return v;
    const unsigned s = lane - d;
    if (!(mask >> s & 1u)) return v;
    return emu::from_bits<T>(g.slot[ph][s]);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int ph; emu::Group& g = emu::rendezvous(mask, emu::to_bits(v), ph);
    const unsigned lane = emu::lane_id();
    const unsigned base = lane & ~(unsigned)(width - 1);
    if (lane + d >= base + (unsigned)width) return v;
    const unsigned s = lane + d;
    if (!(mask >> s & 1u)) return v;
    return emu::from_bits<T>(g.slot[ph][s]);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    int ph; emu::Group& g = emu::rendezvous(mask, emu::to_bits(v), ph);
    const unsigned s = emu::lane_id() ^ (unsigned)x;
    (void)width;
    if (!(mask >> s & 1u)) return v;
    return emu::from_bits<T>(g.slot[ph][s]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    int ph; emu::Group& g = emu::rendezvous(mask, pred ? 1 : 0, ph);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && g.slot[ph][l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    int ph; emu::Group& g = emu::rendezvous(mask, emu::to_bits(v), ph);
    const uint64_t mine = emu::to_bits(v);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && g.slot[ph][l] == mine) r |= 1u << l;
    return r;
}
#define EMU_REDUCE(name, T, init, op)                                                       \
    static inline T name(unsigned mask, T v) {                                              \
        int ph; emu::Group& g = emu::rendezvous(mask, emu::to_bits(v), ph);                 \
        T r = init;                                                                         \
        for (unsigned l = 0; l < 32; l++) if (mask >> l & 1u) { T o = emu::from_bits<T>(g.slot[ph][l]); r = op; } \
        return r;                                                                           \
    }
EMU_REDUCE(__reduce_add_sync, int, 0, r + o)
EMU_REDUCE(__reduce_add_sync, unsigned, 0u, r + o)
EMU_REDUCE(__reduce_min_sync, int, INT_MAX, (o < r ? o : r))
EMU_REDUCE(__reduce_min_sync, unsigned, UINT_MAX, (o < r ? o : r))
EMU_REDUCE(__reduce_max_sync, int, INT_MIN, (o > r ? o : r))
EMU_REDUCE(__reduce_max_sync, unsigned, 0u, (o > r ? o : r))
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, (r | o))
EMU_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, (r & o))
#undef EMU_REDUCE

static inline void __syncthreads() {
    emu::Cta* c = emu::g_cta;
    const uint32_t k = emu::g_cur->bar_gen++;
    c->bar_arrived++;
    const uint32_t want = (uint32_t)c->n_threads * (k + 1);
    while (c->bar_arrived < want) emu::yield();
    c->progress = true;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() {}

// ---------------------------------------------------------------------------------- runtime shim
typedef int cudaError_t;
typedef struct emuStream_* cudaStream_t;
typedef struct emuEvent_* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };

namespace emu { size_t mem_limit(); extern std::atomic<size_t> g_mem_used; void note_alloc(void* p, size_t n); size_t note_free(void* p); }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = getenv("FCX_EMU_DEVICES") ? atoi(getenv("FCX_EMU_DEVICES")) : 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "emu error"; }
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    if (emu::g_mem_used.load() + n > emu::mem_limit()) { *p = nullptr; return cudaErrorMemoryAllocation; }
    *p = aligned_alloc(256, (n + 255) & ~(size_t)255);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xcd, n);                 // poison: device memory is not zero-initialised
    emu::note_alloc(*p, n);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { if (p) { emu::note_free(p); free(p); } return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *t = emu::mem_limit(); *f = *t - std::min(*t, emu::g_mem_used.load()); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = getenv("FCX_EMU_SMS") ? atoi(getenv("FCX_EMU_SMS")) : 4; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* ok, int, int) { *ok = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
