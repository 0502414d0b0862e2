// simt.cpp — fiber scheduler of the TEST-ONLY SIMT emulator (see simt.h).
#define TSGPU_SIMT 1
#include <cstdlib>
#include <cstring>
#include "simt.h"

namespace simt {
Block* g_blk = nullptr;
Fiber* g_cur = nullptr;
size_t g_collectives = 0;
static const size_t STACK = 256 * 1024;
// TSGPU_SIMT_ORDER=reverse: lanes run from the highest index down between synchronisation points.  In ascending order
// lane 0 always runs first, which hides a missing __syncwarp() after a "lane 0 writes, everyone reads" section.
static const bool g_reverse = getenv("TSGPU_SIMT_ORDER") && !strncmp(getenv("TSGPU_SIMT_ORDER"), "rev", 3);
// TSGPU_SIMT_ORDER=random[:seed]: the next lane to run is drawn at random at every synchronisation point
static const bool g_random = getenv("TSGPU_SIMT_ORDER") && !strncmp(getenv("TSGPU_SIMT_ORDER"), "random", 6);
static uint32_t g_rng = g_random && strchr(getenv("TSGPU_SIMT_ORDER"), ':') ? (uint32_t)atoi(strchr(getenv("TSGPU_SIMT_ORDER"), ':') + 1) * 2654435761u + 1u : 12345u;
static inline uint32_t next_rand() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5; return g_rng; }

asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
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
.size simt_switch,.-simt_switch
)");

static void fiber_main() {
    (*g_blk->body)();
    Fiber* f = g_cur;
    f->done = true;
    g_blk->alive--;
    Warp& w = g_blk->warps[f->warp];
    w.alive--;
    // a lane leaving may complete a rendezvous the others are parked on
    if (w.alive > 0 && w.count >= w.alive) { w.count = 0; w.gen++; }
    if (g_blk->alive > 0 && g_blk->bar_count >= g_blk->alive) { g_blk->bar_count = 0; g_blk->bar_gen++; }
    void* dummy;
    simt_switch(&dummy, g_blk->sched_sp);
    abort();
}

void yield() {
    // plain round-robin over the whole block: a lane parked on a block barrier must not starve other warps
    Block& b = *g_blk;
    int n = (int)b.fibers.size();
    int me = b.cur;
    const int start = g_random ? (int)(next_rand() % (uint32_t)n) : 0;
    for (int k = 1; k <= n; k++) {
        int t = g_random ? start + k : g_reverse ? me - k : me + k;
        t %= n; if (t < 0) t += n;
        if (t == me) continue;
        if (!b.fibers[t].done) {
            Fiber* from = &b.fibers[me];
            b.cur = t; g_cur = &b.fibers[t];
            simt_switch(&from->sp, b.fibers[t].sp);
            return;
        }
    }
}

static std::vector<uint8_t*> g_stacks;

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, std::function<void()> body) {
    int nthreads = (int)(block.x * block.y * block.z);
    while ((int)g_stacks.size() < nthreads) {
        void* p = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("mmap"); abort(); }
        g_stacks.push_back((uint8_t*)p);
    }
    std::vector<uint8_t> smem(dyn_smem_bytes + 64);
    // TSGPU_SIMT_POISON=1: shared memory starts as garbage for every block, as on the device (a kernel that relies on
    // zeroed shared memory passes the plain emulator and fails on a GPU)
    static const bool poison = getenv("TSGPU_SIMT_POISON") && atoi(getenv("TSGPU_SIMT_POISON")) != 0;
    uint32_t poison_seed = 0x9E3779B9u;
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        if (poison) { for (auto& v : smem) { poison_seed = poison_seed * 1664525u + 1013904223u; v = (uint8_t)(poison_seed >> 24); } }
        Block b;
        b.bid = uint3{bx, by, bz};
        b.bdim = block; b.gdim = grid;
        b.body = &body;
        b.dyn_smem = (uint8_t*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
        b.fibers.resize(nthreads);
        b.warps.resize((nthreads + 31) / 32);
        b.alive = nthreads;
        for (int t = 0; t < nthreads; t++) {
            Fiber& f = b.fibers[t];
            f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
            f.lane = t & 31; f.warp = t >> 5;
            b.warps[f.warp].alive++;
            f.stack = g_stacks[t];
            uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
            uint64_t* sp = (uint64_t*)top;
            *--sp = 0;                          // fake return address of fiber_main
            *--sp = (uint64_t)(uintptr_t)&fiber_main;
            for (int k = 0; k < 6; k++) *--sp = 0;
            f.sp = sp;
        }
        g_blk = &b;
        // run until every fiber is done; each return to the scheduler means one fiber finished
        while (b.alive > 0) {
            int t = -1;
            if (g_reverse) { for (int k = nthreads - 1; k >= 0; k--) if (!b.fibers[k].done) { t = k; break; } }
            else for (int k = 0; k < nthreads; k++) if (!b.fibers[k].done) { t = k; break; }
            if (t < 0) break;
            b.cur = t; g_cur = &b.fibers[t];
            simt_switch(&b.sched_sp, b.fibers[t].sp);
        }
        g_blk = nullptr; g_cur = nullptr;
    }
}
}  // namespace simt
