// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): fiber scheduler of the SIMT emulator.
#include "cuda_emu.h"

#include <map>
#include <mutex>
#include <sys/mman.h>

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

thread_local Cta* g_cta = nullptr;
thread_local Fiber* g_cur = nullptr;
std::atomic<size_t> g_mem_used{0};
static std::mutex g_alloc_mu;
static std::map<void*, size_t> g_allocs;

size_t mem_limit() {
    static size_t lim = getenv("FCX_EMU_MEM_GB") ? (size_t)(atof(getenv("FCX_EMU_MEM_GB")) * (double)(1ull << 30)) : (size_t)24 << 30;
    return lim;
}
void note_alloc(void* p, size_t n) { std::lock_guard<std::mutex> lk(g_alloc_mu); g_allocs[p] = n; g_mem_used += n; }
size_t note_free(void* p) {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return 0;
    size_t n = it->second; g_mem_used -= n; g_allocs.erase(it); return n;
}

[[noreturn]] void fail(const char* what) {
    fprintf(stderr, "cuda_emu: %s (block %u thread %u)\n", what, g_cta ? g_cta->bidx.x : 0u, g_cur ? g_cur->tidx.x : 0u);
    abort();
}

void yield() {
    Fiber* f = g_cur;
    emu_switch(&f->sp, g_cta->sched_sp);
}

static void fiber_entry() {
    Cta* c = g_cta;
    (*c->fn)();
    g_cur->done = true;
    c->progress = true;
    emu_switch(&g_cur->sp, c->sched_sp);
    fail("resumed a finished fiber");
}

constexpr size_t STACK_BYTES = 256 * 1024;

struct Worker {                     // per OS thread: fiber stacks, reused across CTAs
    std::vector<char*> stacks;
    std::vector<char> dyn;
    ~Worker() { for (char* s : stacks) munmap(s, STACK_BYTES); }
    char* stack(size_t i) {
        while (stacks.size() <= i) {
            void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) fail("mmap of a fiber stack failed");
            stacks.push_back((char*)p);
        }
        return stacks[i];
    }
};

static void run_cta(Worker& wk, Cta& c) {
    const size_t n = c.n_threads;
    c.fibers.assign(n, Fiber());
    c.warps.assign((n + 31) / 32, Warp());
    c.bar_arrived = 0;
    for (size_t i = 0; i < n; i++) {
        Fiber& f = c.fibers[i];
        f.tidx.x = (unsigned)i; f.done = false; f.bar_gen = 0;
        memset(f.gen, 0, sizeof f.gen);
        f.stack = wk.stack(i);
        void** top = (void**)(f.stack + STACK_BYTES);
        // [r15 r14 r13 r12 rbx rbp ret dummy]: after the pops and the ret, rsp = top - 8 (call-aligned)
        top[-1] = nullptr;
        top[-2] = (void*)&fiber_entry;
        for (int k = 3; k <= 8; k++) top[-k] = nullptr;
        f.sp = (void*)(top - 8);
    }
    g_cta = &c;
    size_t live = n;
    int idle_rounds = 0;
    while (live) {
        c.progress = false;
        for (size_t i = 0; i < n; i++) {
            Fiber& f = c.fibers[i];
            if (f.done) continue;
            g_cur = &f;
            emu_switch(&c.sched_sp, f.sp);
            if (f.done) live--;
        }
        if (!c.progress) { if (++idle_rounds > 2) { g_cur = nullptr; fail("deadlock: a collective or barrier can never complete"); } }
        else idle_rounds = 0;
    }
    g_cur = nullptr; g_cta = nullptr;
}

void launch(Dim3 grid, Dim3 block, size_t smem, const std::function<void()>& fn) {
    const unsigned nblk = grid.x;
    if (nblk == 0 || block.x == 0) return;
    static const unsigned max_thr = getenv("FCX_EMU_THREADS") ? (unsigned)atoi(getenv("FCX_EMU_THREADS")) : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    const unsigned nthr = std::max(1u, std::min(max_thr, nblk));
    std::atomic<unsigned> next(0);
    auto body = [&]() {
        Worker wk;
        wk.dyn.assign(smem + 64, 0);
        Cta c;
        c.bdim = block; c.gdim = grid; c.fn = &fn; c.n_threads = block.x;
        c.dyn_smem = (char*)(((uintptr_t)wk.dyn.data() + 15) & ~(uintptr_t)15);
        for (;;) {
            const unsigned b = next.fetch_add(1);
            if (b >= nblk) break;
            c.bidx.x = b;
            run_cta(wk, c);
        }
    };
    if (nthr == 1) body();
    else {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < nthr; i++) th.emplace_back(body);
        for (auto& t : th) t.join();
    }
}

}  // namespace emu
