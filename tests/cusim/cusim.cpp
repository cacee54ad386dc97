// cusim.cpp -- TEST INFRASTRUCTURE ONLY (see cusim.h).  Fiber scheduler for one CUDA block.
#include "cusim.h"

namespace cusim {
Fiber* cur = nullptr;
Block* blk = nullptr;
uint3 g_blockIdx{0, 0, 0};
dim3 g_blockDim, g_gridDim;
ucontext_t sched_ctx;
long long n_launches = 0;

static const size_t kStack = 512 * 1024;
static const std::function<void()>* g_body = nullptr;

void yield() { swapcontext(&cur->ctx, &sched_ctx); }

static void release_if_complete() {
    if (blk->live > 0 && blk->arrived == blk->live) { blk->arrived = 0; blk->gen++; }
    for (auto& w : blk->warps)
        if (w.live > 0 && w.arrived == w.live) { w.arrived = 0; w.gen++; }
}

void syncthreads() {
    unsigned g = blk->gen;
    blk->arrived++;
    release_if_complete();
    while (blk->gen == g) yield();
}

void syncwarp() {
    Warp& w = warp();
    unsigned g = w.gen;
    w.arrived++;
    release_if_complete();
    while (w.gen == g) yield();
}

unsigned ballot(int pred) {
    // full-warp collectives only: every live lane of the warp must call this (as CUDA requires
    // for a full mask); lanes that never existed contribute 0.
    Warp& w = warp();
    w.slot[lane()] = pred ? 1u : 0u;
    syncwarp();
    unsigned r = 0;
    unsigned base = (cur->linear >> 5) << 5;
    for (unsigned i = 0; i < 32; ++i)
        if (base + i < blk->fibers.size() && w.slot[i]) r |= 1u << i;
    syncwarp();
    return r;
}

static void trampoline() {
    (*g_body)();
    cur->done = true;
    blk->live--;
    blk->warps[cur->linear >> 5].live--;
    release_if_complete();
    swapcontext(&cur->ctx, &sched_ctx);
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    n_launches++;
    g_gridDim = grid;
    g_blockDim = block;
    g_body = &body;
    unsigned nthreads = block.x * block.y * block.z;
    Block b;
    b.fibers.resize(nthreads);
    b.warps.resize((nthreads + 31) / 32);
    b.smem = (char*)aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
    for (auto& f : b.fibers) f.stack = (char*)malloc(kStack);
    blk = &b;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3{bx, by, bz};
                memset(b.smem, 0xFF, smem_bytes);   // poison shared memory
                b.live = nthreads; b.arrived = 0; b.gen = 0;
                for (auto& w : b.warps) { w.live = 0; w.arrived = 0; w.gen = 0; w.ballot_acc = 0; }
                for (unsigned t = 0; t < nthreads; ++t) {
                    Fiber& f = b.fibers[t];
                    f.done = false;
                    f.linear = t;
                    f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    b.warps[t >> 5].live++;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                }
                unsigned remaining = nthreads;
                while (remaining) {
                    remaining = 0;
                    for (unsigned t = 0; t < nthreads; ++t) {
                        Fiber& f = b.fibers[t];
                        if (f.done) continue;
                        cur = &f;
                        swapcontext(&sched_ctx, &f.ctx);
                        if (!f.done) remaining++;
                    }
                }
            }
    for (auto& f : b.fibers) free(f.stack);
    free(b.smem);
    blk = nullptr;
    cur = nullptr;
}
}  // namespace cusim
