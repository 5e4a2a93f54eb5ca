// TEST INFRASTRUCTURE ONLY -- SIMT emulator behind tests/cuemu/include/cuda_runtime.h (see there).
// One CUDA block = blockDim fibers on the calling OS thread; a fiber runs until it reaches a block or warp barrier,
// then the scheduler resumes the next runnable one.  Blocks of a grid run one after the other; launches from
// different OS threads are serialised.  Context switches are a handful of instructions (x86-64 System V only).
#include "include/cuda_runtime.h"
#include <chrono>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#if !defined(__x86_64__)
#error "the test-only CUDA emulator supports x86-64 only"
#endif

namespace cuemu {

extern "C" void cuemu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
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
.size cuemu_switch,.-cuemu_switch
)");

enum Wait { RUN = 0, W_BLOCK = 1, W_WARP = 2, DONE = 3 };

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    ThreadCtx ctx;
    int wait = RUN;
    unsigned warp_mask = 0;
    bool released = false;               // warp barrier: set by the scheduler once every lane of the mask has arrived
    unsigned long long gen = 0;          // generation of the block barrier it waits on
};

struct WarpState {
    unsigned long long slot[32];
    int pred[32];
    unsigned arrived = 0;                // lanes waiting at a warp barrier
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    int live = 0, block_arrived = 0;
    unsigned long long block_gen = 0;
    int red_and = 1, red_or = 0, red_and_out = 1, red_or_out = 0;
    void *sched_sp = nullptr;
    int current = -1;
    const std::function<void()> *body = nullptr;
};

static constexpr size_t kStack = 192 * 1024;
static std::mutex g_launch_mutex;
static thread_local Block *g_blk = nullptr;
static thread_local ThreadCtx g_host_ctx{};
static std::vector<unsigned char> g_dyn_smem;

ThreadCtx &cur() { return g_blk && g_blk->current >= 0 ? g_blk->fibers[g_blk->current].ctx : g_host_ctx; }
void *dyn_smem() { return g_dyn_smem.data(); }
long long clock() { return (long long)__builtin_ia32_rdtsc(); }      // CPU time stamp: section profiles show the distribution of WORK (all threads serialised)

static void yield_to_scheduler() {
    Block &b = *g_blk;
    Fiber &f = b.fibers[b.current];
    cuemu_switch(&f.sp, b.sched_sp);
}

static void fiber_main() {
    Block &b = *g_blk;
    (*b.body)();
    Fiber &f = b.fibers[b.current];
    f.wait = DONE;
    b.live--;
    yield_to_scheduler();
    std::abort();                        // a finished fiber is never resumed
}

static long long g_stat_block_barriers = 0, g_stat_warp_barriers = 0;
void sync_block() {
    Block &b = *g_blk;
    Fiber &f = b.fibers[b.current];
    if (b.current == 0) g_stat_block_barriers++;
    f.wait = W_BLOCK; f.gen = b.block_gen;
    b.block_arrived++;
    yield_to_scheduler();
}

int sync_block_reduce(int pred, int mode) {
    Block &b = *g_blk;
    b.red_and &= (pred != 0); b.red_or |= (pred != 0);
    sync_block();
    return mode == 0 ? b.red_and_out : b.red_or_out;
}

static void warp_barrier(unsigned mask) {
    Block &b = *g_blk;
    Fiber &f = b.fibers[b.current];
    if (b.current == 0) g_stat_warp_barriers++;
    WarpState &w = b.warps[f.ctx.warp];
    f.wait = W_WARP; f.warp_mask = mask; f.released = false;
    w.arrived |= 1u << f.ctx.lane;
    yield_to_scheduler();
}

void sync_warp(unsigned mask) { warp_barrier(mask); }

unsigned long long exchange(unsigned mask, unsigned long long v, int src_lane, bool valid_src) {
    Block &b = *g_blk;
    Fiber &f = b.fibers[b.current];
    WarpState &w = b.warps[f.ctx.warp];
    w.slot[f.ctx.lane] = v;
    warp_barrier(mask);
    unsigned long long r = v;
    if (valid_src && src_lane >= 0 && src_lane < 32 && ((mask >> src_lane) & 1u)) {
        const int idx = f.ctx.warp * 32 + src_lane;
        if (idx < (int)b.fibers.size() && b.fibers[idx].wait != DONE) r = w.slot[src_lane];
    }
    warp_barrier(mask);                  // nobody overwrites a slot before every lane has read
    return r;
}

unsigned ballot(unsigned mask, int pred) {
    Block &b = *g_blk;
    Fiber &f = b.fibers[b.current];
    WarpState &w = b.warps[f.ctx.warp];
    w.pred[f.ctx.lane] = pred != 0;
    warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) {
        const int idx = f.ctx.warp * 32 + l;
        if (((mask >> l) & 1u) && idx < (int)b.fibers.size() && b.fibers[idx].wait != DONE && w.pred[l]) r |= 1u << l;
    }
    warp_barrier(mask);
    return r;
}

// CUEMU_ORDER=reverse runs the fibers of a block from the last thread to the first: a kernel whose result depends on
// the order in which threads reach a point (a missing __syncthreads / __syncwarp) shows up as a difference between
// the two orders.
static bool reverse_order() {
    static const bool r = [] { const char *e = std::getenv("CUEMU_ORDER"); return e && std::string(e) == "reverse"; }();
    return r;
}

static void run_block(Block &b) {
    g_blk = &b;
    const int n = (int)b.fibers.size();
    const bool rev = reverse_order();
    while (b.live > 0) {
        bool progressed = false;
        for (int ii = 0; ii < n; ++ii) {
            const int i = rev ? n - 1 - ii : ii;
            Fiber &f = b.fibers[i];
            if (f.wait == DONE) continue;
            if (f.wait == W_BLOCK) { if (f.gen == b.block_gen) continue; f.wait = RUN; }
            if (f.wait == W_WARP) { if (!f.released) continue; f.wait = RUN; }
            b.current = i;
            cuemu_switch(&b.sched_sp, f.sp);
            b.current = -1;
            progressed = true;
        }
        // release barriers whose participants have all arrived (exited threads do not take part)
        if (b.live > 0 && b.block_arrived == b.live) {
            b.block_arrived = 0; b.block_gen++;
            b.red_and_out = b.red_and; b.red_or_out = b.red_or; b.red_and = 1; b.red_or = 0;
            progressed = true;
        }
        for (size_t wi = 0; wi < b.warps.size(); ++wi) {
            WarpState &w = b.warps[wi];
            if (!w.arrived) continue;
            // a group (all waiters that named the same mask) is released once every live lane of that mask waits in it
            unsigned done_masks[32]; int nd = 0;
            for (int l0 = 0; l0 < 32; ++l0) {
                const int i0 = (int)wi * 32 + l0;
                if (i0 >= n) break;
                const Fiber &f0 = b.fibers[i0];
                if (f0.wait != W_WARP || f0.released) continue;
                const unsigned m = f0.warp_mask;
                bool seen = false;
                for (int k = 0; k < nd; ++k) seen = seen || done_masks[k] == m;
                if (seen) continue;
                done_masks[nd++] = m;
                bool all = true;
                for (int l = 0; l < 32 && all; ++l) {
                    const int idx = (int)wi * 32 + l;
                    if (idx >= n) break;
                    if (!((m >> l) & 1u)) continue;
                    const Fiber &f = b.fibers[idx];
                    if (f.wait == DONE) continue;
                    all = f.wait == W_WARP && !f.released && f.warp_mask == m;
                }
                if (!all) continue;
                for (int l = 0; l < 32; ++l) {
                    const int idx = (int)wi * 32 + l;
                    if (idx >= n) break;
                    Fiber &f = b.fibers[idx];
                    if (((m >> l) & 1u) && f.wait == W_WARP && !f.released && f.warp_mask == m) { f.released = true; w.arrived &= ~(1u << l); }
                }
                progressed = true;
            }
        }
        if (!progressed) {
            std::fprintf(stderr, "cuemu: deadlock (block %u,%u,%u): %d live threads, %d at the block barrier\n", b.fibers[0].ctx.bid.x, b.fibers[0].ctx.bid.y,
                         b.fibers[0].ctx.bid.z, b.live, b.block_arrived);
            std::abort();
        }
    }
    g_blk = nullptr;
}

bool replaying();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    if (replaying()) return;
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    if (g_dyn_smem.size() < smem + 64) g_dyn_smem.resize(smem + 64);
    static thread_local std::vector<char *> stacks;
    while ((int)stacks.size() < nthreads) stacks.push_back((char *)std::aligned_alloc(64, kStack));
    Block b;
    b.body = &body;
    b.fibers.resize(nthreads);
    b.warps.resize((nthreads + 31) / 32);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                b.live = nthreads; b.block_arrived = 0; b.red_and = 1; b.red_or = 0;
                for (auto &w : b.warps) { w.arrived = 0; }
                for (int t = 0; t < nthreads; ++t) {
                    Fiber &f = b.fibers[t];
                    f.stack = stacks[t];
                    f.wait = RUN;
                    f.ctx.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                    f.ctx.bid = uint3{bx, by, bz};
                    f.ctx.bdim = block; f.ctx.gdim = grid;
                    f.ctx.linear = t; f.ctx.lane = t & 31; f.ctx.warp = t >> 5;
                    // initial frame: six callee-saved registers (zero) and the entry point; rsp % 16 == 8 at entry
                    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
                    void **sp = (void **)top;
                    *--sp = nullptr;                       // fake return address slot keeps the ABI alignment
                    *--sp = (void *)&fiber_main;
                    for (int k = 0; k < 6; ++k) *--sp = nullptr;
                    f.sp = sp;
                }
                run_block(b);
            }
    // CUEMU_STATS=1: barriers executed by thread 0 of the last block of every launch (a rough latency-chain length)
    static const bool stats = std::getenv("CUEMU_STATS") != nullptr;
    if (stats) std::fprintf(stderr, "cuemu launch grid %u block %u: block barriers %lld warp barriers %lld (cumulative, thread 0)\n", grid.x * grid.y * grid.z, block.x, g_stat_block_barriers, g_stat_warp_barriers);
}

}  // namespace cuemu

// ---- record / replay: CUEMU_RECORD=file stores every device->host copy; CUEMU_REPLAY=file skips the kernels and serves
// the recorded bytes instead, so the deterministic host glue can be re-run (and profiled) at full speed ----
#include <cstdio>
namespace cuemu {
static FILE *g_rec = nullptr, *g_rep = nullptr;
static bool g_mode_init = false;
static void mode_init() {
    if (g_mode_init) return;
    g_mode_init = true;
    if (const char *p = std::getenv("CUEMU_RECORD")) g_rec = std::fopen(p, "wb");
    if (const char *p = std::getenv("CUEMU_REPLAY")) g_rep = std::fopen(p, "rb");
}
bool replaying() { mode_init(); return g_rep != nullptr; }
void d2h(void *dst, const void *src, size_t n) {
    mode_init();
    if (g_rep) { size_t m = 0; if (std::fread(&m, sizeof(m), 1, g_rep) != 1 || m != n || (n && std::fread(dst, 1, n, g_rep) != n)) { std::fprintf(stderr, "cuemu: replay out of sync (%zu vs %zu)\n", m, n); std::abort(); } return; }
    if (n) std::memmove(dst, src, n);
    if (g_rec) { std::fwrite(&n, sizeof(n), 1, g_rec); if (n) std::fwrite(dst, 1, n, g_rec); std::fflush(g_rec); }
}
void replay_rewind() { if (g_rep) std::rewind(g_rep); }
}
extern "C" void cuemu_replay_rewind() { cuemu::replay_rewind(); }
// ---- runtime stubs ----
struct cuemu_event { std::chrono::steady_clock::time_point t; };
cudaError_t cudaMalloc(void **p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k) { if (k == cudaMemcpyDeviceToHost) cuemu::d2h(dst, src, n); else if (n && !cuemu::replaying()) std::memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, n, k); }
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
    if (cuemu::replaying()) return cudaSuccess;
    for (size_t r = 0; r < height; ++r) std::memmove((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned) { *st = (void *)1; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuemu_event(); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
