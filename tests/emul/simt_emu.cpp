// simt_emu.cpp -- TEST INFRASTRUCTURE ONLY: the fiber scheduler behind simt_emu.h.
#include "simt_emu.h"

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64"
#endif

// Cooperative context switch: save the callee-saved registers and the stack pointer of the running
// fiber, load the target's.  (ucontext's swapcontext does a sigprocmask system call per switch.)
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
extern "C" void simt_switch(void** from_sp, void* to_sp);

namespace simt {

Ctx* g_cur = nullptr;
uint3 g_blockIdx = {0, 0, 0};
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber { Ctx ctx; void* sp; char* stack; bool done; };
struct Warp {
    int arrived, gen, live;
    uint64_t slots[2][32];
    bool present[2][32];
    uint64_t result[2];
    int op[2];
};
std::vector<Fiber> fibers;
std::vector<char*> stack_pool;
std::vector<Warp> warps;
int cur_idx = -1, live = 0;
void* sched_sp = nullptr;
const std::function<void()>* body = nullptr;
int bar_arrived = 0, bar_gen = 0, bar_cnt = 0, bar_cnt_res[2] = {0, 0};
unsigned long long progress = 0;

void yield() { simt_switch(&fibers[cur_idx].sp, sched_sp); }

void warp_finish(Warp& W) {
    const int p = W.gen & 1;
    uint64_t ballot = 0, mx = 0; bool all = true, any = false;
    for (int l = 0; l < 32; l++)
        if (W.present[p][l]) {
            const uint64_t v = W.slots[p][l];
            if (v) ballot |= 1ull << l;
            mx = std::max(mx, v); all = all && v != 0; any = any || v != 0;
        }
    switch (W.op[p]) { case 0: W.result[p] = ballot; break; case 1: W.result[p] = mx; break;
                       case 2: W.result[p] = all; break; default: W.result[p] = any; break; }
    for (int l = 0; l < 32; l++) W.present[p ^ 1][l] = false;
    W.arrived = 0; W.gen++; progress++;
}

void fiber_exit() {
    Fiber& f = fibers[cur_idx];
    f.done = true; live--; progress++;
    Warp& W = warps[f.ctx.warp];
    W.live--;
    if (W.arrived > 0 && W.arrived >= W.live) warp_finish(W);
    if (bar_arrived > 0 && bar_arrived >= live) { bar_cnt_res[bar_gen & 1] = bar_cnt; bar_cnt = 0; bar_arrived = 0; bar_gen++; }
    for (;;) simt_switch(&f.sp, sched_sp);
}

void fiber_main() {
    (*body)();
    fiber_exit();
}
}  // namespace

void fiber_yield() { yield(); }

void block_barrier() { block_count(0); }

int block_count(int pred) {
    const int gen = bar_gen;
    bar_cnt += pred ? 1 : 0;
    if (++bar_arrived >= live) { bar_cnt_res[gen & 1] = bar_cnt; bar_cnt = 0; bar_arrived = 0; bar_gen++; progress++; }
    else while (bar_gen == gen) yield();
    return bar_cnt_res[gen & 1];
}

static uint64_t warp_arrive(uint64_t mine, int op, int& p_out) {
    Warp& W = warps[g_cur->warp];
    const int p = W.gen & 1, lane = g_cur->lane, gen = W.gen;
    W.slots[p][lane] = mine; W.present[p][lane] = true; W.op[p] = op;
    if (++W.arrived >= W.live) warp_finish(W);
    else while (W.gen == gen) yield();
    p_out = p;
    return W.result[p];
}

uint64_t warp_exchange(uint64_t mine, int src_lane) {
    int p;
    warp_arrive(mine, 3, p);
    Warp& W = warps[g_cur->warp];
    return (src_lane >= 0 && src_lane < 32 && W.present[p][src_lane]) ? W.slots[p][src_lane] : mine;
}

uint64_t warp_reduce(uint64_t mine, int op) {
    int p;
    return warp_arrive(mine, op, p);
}

void run_grid(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& fn) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) { fprintf(stderr, "simt_emu: bad block size %d\n", nthreads); abort(); }
    while ((int)stack_pool.size() < nthreads) stack_pool.push_back((char*)aligned_alloc(64, kStack));
    std::vector<unsigned char> smem(dyn_smem + 256);
    g_dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    g_blockDim = block; g_gridDim = grid; body = &fn;
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx = uint3{bx, by, bz};
        fibers.assign((size_t)nthreads, Fiber{});
        warps.assign((size_t)((nthreads + 31) / 32), Warp{});
        live = nthreads; bar_arrived = 0; bar_gen = 0; bar_cnt = 0;
        for (int t = 0; t < nthreads; t++) {
            Fiber& f = fibers[(size_t)t];
            f.ctx.linear = t; f.ctx.lane = t & 31; f.ctx.warp = t >> 5;
            f.ctx.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
            f.stack = stack_pool[(size_t)t]; f.done = false;
            warps[(size_t)(t >> 5)].live++;
            // initial frame: six callee-saved registers, then the return address the first switch "returns" to
            uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
            void** slot = (void**)(top - 16);
            *slot = (void*)&fiber_main;
            void** sp = slot - 6;
            for (int k = 0; k < 6; k++) sp[k] = nullptr;
            f.sp = sp;
        }
        unsigned long long last_progress = progress; int idle_passes = 0;
        while (live > 0) {
            for (int t = 0; t < nthreads && live > 0; t++) {
                if (fibers[(size_t)t].done) continue;
                cur_idx = t; g_cur = &fibers[(size_t)t].ctx;
                simt_switch(&sched_sp, fibers[(size_t)t].sp);
            }
            if (progress == last_progress) {
                if (++idle_passes > 100000) {
                    fprintf(stderr, "simt_emu: deadlock in block (%u,%u,%u): %d threads alive, %d at the block barrier\n",
                            bx, by, bz, live, bar_arrived);
                    abort();
                }
            } else { last_progress = progress; idle_passes = 0; }
        }
    }
    g_cur = nullptr; cur_idx = -1; body = nullptr;
}

}  // namespace simt
