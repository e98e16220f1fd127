// emu_lib.cpp — TEST INFRASTRUCTURE: builds the generic window kernel's real CUDA sources for the host on top of
// cuda_emu.h and exports the same C ABI entry points (snn_b200_run_window, snn_b200_workspace_bytes, ...) operating on
// HOST memory.  tests/emu/emu.py routes the host API to it the way oracle/oracle.py routes it to the oracle.
//
//   g++ -O1 -std=c++17 -fPIC -shared -DSNN_EMU -ffp-contract=off -Itests/emu -o tests/emu/libsnn_emu.so tests/emu/emu_lib.cpp -lpthread
#include <execinfo.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <thread>
#include <vector>

#include "cuda_emu.h"

namespace emu {

thread_local Cta *tls_cta = nullptr;
thread_local Fiber *tls_cur = nullptr;

static constexpr size_t STACK_BYTES = 512 * 1024;

// SNN_EMU_SHUFFLE=<seed>: instead of round robin, the next fiber is drawn at random — different interleavings of the
// warps between two synchronisation points, i.e. a poor man's race check for a missing __syncthreads / __syncwarp.
static unsigned long long g_shuffle = 0;

static inline unsigned next_rand(Cta *c) {
    unsigned long long x = c->rng;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    c->rng = x;
    return (unsigned)(x >> 24);
}

void yield() {
    Cta *c = tls_cta;
    Fiber *from = tls_cur;
    int k = c->current;
    if (g_shuffle) {
        k = (int)(next_rand(c) % (unsigned)c->nthreads);
        for (int step = 0; step < c->nthreads && c->fibers[k].done; ++step) k = (k + 1) % c->nthreads;
        if (c->fibers[k].done) return;
    } else {
        for (int step = 0; step < c->nthreads; ++step) {
            k = (k + 1) % c->nthreads;
            if (!c->fibers[k].done) break;
        }
    }
    if (k == c->current) return;   // nobody else is runnable
    c->current = k;
    tls_cur = &c->fibers[k];
    swapcontext(&from->ctx, &c->fibers[k].ctx);
}

static void fiber_main() {
    Fiber *f = tls_cur;
    Cta *c = f->cta;
    c->entry(c->arg);
    f->done = true;
    --c->live;
    if (c->live == 0) {
        setcontext(&c->main_ctx);
    } else {
        int k = c->current;
        do { k = (k + 1) % c->nthreads; } while (c->fibers[k].done);
        c->current = k;
        tls_cur = &c->fibers[k];
        setcontext(&c->fibers[k].ctx);
    }
}

static void run_cta(int bid, int grid, int block, size_t smem_bytes, void (*entry)(void *), void *arg, int bid_y = 0, int grid_y = 1) {
    Cta cta;
    cta.nthreads = block;
    cta.nwarps = (block + WARP - 1) / WARP;
    cta.live = block;
    cta.bidx = {(unsigned)bid, (unsigned)bid_y, 0};
    cta.bdim = dim3(block);
    cta.gdim = dim3(grid, grid_y);
    cta.entry = entry;
    cta.arg = arg;
    cta.rng = (g_shuffle + 1) * 0x9E3779B97F4A7C15ull + (unsigned long long)(bid + 1) * 0xD1B54A32D192ED03ull;
    std::vector<Fiber> fibers(block);
    std::vector<WarpState> warps(cta.nwarps);
    // dynamic shared memory ends at a PROT_NONE guard page: an access beyond the size the host computed for the launch
    // (gen_smem_bytes / smem_layout) faults instead of silently landing in a neighbour
    const size_t page = 4096, smem_al = (smem_bytes + 63) & ~(size_t)63, smem_map = ((smem_al + page - 1) / page + 1) * page;
    char *smem_base = (char *)mmap(nullptr, smem_map, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (smem_base == MAP_FAILED) { perror("emu: mmap smem"); abort(); }
    mprotect(smem_base + smem_map - page, page, PROT_NONE);
    cta.fibers = fibers.data();
    cta.warps = warps.data();
    cta.dyn_smem = (float *)(smem_base + smem_map - page - smem_al);
    // shared memory starts out as garbage on the GPU: poison it (0xFF bytes = NaN floats / -1 integers), so that a kernel
    // relying on zero-initialised shared memory fails here too (the initcheck class of bugs)
    memset(cta.dyn_smem, 0xFF, smem_al);
    memset(cta.static_smem, 0xFF, sizeof(cta.static_smem));
    char *stacks = (char *)mmap(nullptr, STACK_BYTES * block, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == MAP_FAILED) { perror("emu: mmap"); abort(); }
    tls_cta = &cta;
    for (int t = 0; t < block; ++t) {
        Fiber &f = fibers[t];
        f.tid = {(unsigned)t, 0, 0};
        f.lane = t % WARP;
        f.warp = t / WARP;
        f.cta = &cta;
        f.stack = stacks + STACK_BYTES * t;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_main, 0);
    }
    cta.current = 0;
    tls_cur = &fibers[0];
    swapcontext(&cta.main_ctx, &fibers[0].ctx);   // returns when the last fiber finishes
    tls_cur = nullptr;
    tls_cta = nullptr;
    munmap(stacks, STACK_BYTES * block);
    munmap(smem_base, smem_map);
}

// SNN_EMU_TRACE=1: a fault inside an emulated kernel (typically an access beyond a guard page) reports the faulting address,
// the CUDA thread it happened in and a backtrace before the process dies.
static void on_fault(int sig, siginfo_t *si, void *) {
    char buf[256];
    Cta *c = tls_cta; Fiber *f = tls_cur;
    int n = snprintf(buf, sizeof(buf), "\n[emu] signal %d at address %p in block (%u,%u) thread %u\n", sig, si->si_addr, c ? c->bidx.x : 0u,
                     c ? c->bidx.y : 0u, f ? f->tid.x : 0u);
    if (write(2, buf, (size_t)n) < 0) {}
    void *bt[48];
    backtrace_symbols_fd(bt, backtrace(bt, 48), 2);
    _exit(139);
}
static void install_fault_handler() {
    static bool done = false;
    if (done || !getenv("SNN_EMU_TRACE")) return;
    done = true;
    static char altstack[1 << 16];
    stack_t ss; ss.ss_sp = altstack; ss.ss_size = sizeof(altstack); ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa; memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = on_fault; sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr); sigaction(SIGBUS, &sa, nullptr);
}

void run_grid(int grid, int block, size_t smem_bytes, void (*entry)(void *), void *arg) {
    install_fault_handler();
    const char *sh = getenv("SNN_EMU_SHUFFLE");
    g_shuffle = sh ? strtoull(sh, nullptr, 10) : 0ull;
    std::vector<std::thread> ts;
    for (int b = 0; b < grid; ++b) ts.emplace_back([=]() { run_cta(b, grid, block, smem_bytes, entry, arg); });
    for (auto &t : ts) t.join();
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    run_grid_independent((int)grid.x, (int)grid.y, (int)block.x, smem_bytes, [](void *a) { (*(const std::function<void()> *)a)(); }, (void *)&body);
}

void run_grid_independent(int gx, int gy, int block, size_t smem_bytes, void (*entry)(void *), void *arg) {
    const char *sh = getenv("SNN_EMU_SHUFFLE");
    g_shuffle = sh ? strtoull(sh, nullptr, 10) : 0ull;
    const int total = gx * gy, workers = total < 8 ? total : 8;
    std::vector<std::thread> ts;
    for (int w = 0; w < workers; ++w)
        ts.emplace_back([=]() {
            for (int id = w; id < total; id += workers) run_cta(id % gx, gx, block, smem_bytes, entry, arg, id / gx, gy);
        });
    for (auto &t : ts) t.join();
}

}  // namespace emu

// ---- the product's sources, compiled for the host ---------------------------------------------------------------
#include "../../bindsnet_b200/csrc/snn_generic.cu"

// the fused DiehlAndCook2015 window kernel (tier 2, the metric's kernel): its bulk copies / mbarriers / polling loads run on
// the emulation's model of them (cuda_emu.h); the structure check of the static matrices is done on the host here
int snn_verify_structure(const snn_conn_t &C, int n, int32_t *err, cudaStream_t stream);   // snn_ops.cu
#include "../../bindsnet_b200/csrc/snn_fused_dc.cu"

// the column-group kernel (tier 3): emu_dc2.cpp

#include "../../bindsnet_b200/csrc/snn_api.cu"

// ---- the single-operator kernels, the multi-GPU combine, the encoders and the read-out --------------------------------
#include "../../bindsnet_b200/csrc/snn_ops.cu"
#include "../../bindsnet_b200/csrc/snn_encode.cu"
#include "../../bindsnet_b200/csrc/snn_readout.cu"
