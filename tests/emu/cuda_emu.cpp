/* TEST-ONLY: fiber scheduler of the CUDA emulator (see cuda_emu.h). */
#include "cuda_emu.h"

#if !defined(__x86_64__)
#error "the test-only CUDA emulator has an x86-64 context switch"
#endif
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
    movq (%rsi), %rsp
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
Block *g_blk = nullptr;
size_t g_stack_bytes = 256 * 1024;
unsigned long g_progress = 0;
static std::vector<char *> g_stacks;

void fiber_entry()
{
    Block *b = g_blk;
    b->body();
    Fiber &f = b->fibers[(size_t)b->cur];
    f.done = true;
    g_progress++;
    /* a finished thread no longer takes part in barriers */
    int w = b->cur >> 5;
    b->live--;
    b->warp_live[(size_t)w]--;
    if (b->warp_live[(size_t)w] && b->warp_arrived[(size_t)w] == b->warp_live[(size_t)w]) { b->warp_arrived[(size_t)w] = 0; b->warp_gen[(size_t)w]++; }
    if (b->live && b->block_arrived == b->live) { b->block_arrived = 0; b->block_gen++; }
    emu_switch(&f.ctx, &b->sched);
    abort(); /* a finished fiber is never resumed */
}

void launch(emu_dim3 grid, emu_dim3 block, size_t smem, const std::function<void()> &body)
{
    size_t nthreads = (size_t)block.x * block.y * block.z;
    std::vector<char> dyn(smem + 16);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block blk;
                blk.bid = emu_dim3(bx, by, bz); blk.bdim = block; blk.gdim = grid;
                blk.fibers.resize(nthreads);
                size_t nwarps = (nthreads + 31) / 32;
                blk.warp_arrived.assign(nwarps, 0); blk.warp_gen.assign(nwarps, 0); blk.warp_live.assign(nwarps, 0);
                blk.live = (unsigned)nthreads;
                blk.body = body;
                blk.dyn_smem = dyn.data();
                g_blk = &blk;
                for (size_t t = 0; t < nthreads; t++) {
                    Fiber &f = blk.fibers[t];
                    f.tid = emu_dim3((unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y)));
                    if (g_stacks.size() <= t) g_stacks.push_back((char *)malloc(g_stack_bytes));
                    f.stack = g_stacks[t];
                    {   /* initial frame: six callee-saved slots, the entry address `ret` jumps to, a fake return slot */
                        uintptr_t top = ((uintptr_t)f.stack + g_stack_bytes) & ~(uintptr_t)15;
                        void **sp = (void **)(top - 64);
                        for (int k = 0; k < 6; k++) sp[k] = nullptr;
                        sp[6] = (void *)&fiber_entry;
                        sp[7] = nullptr;
                        f.ctx.sp = sp;
                    }
                    blk.warp_live[t >> 5]++;
                }
                size_t remaining = nthreads, stalled = 0;
                while (remaining) {
                    unsigned long before = g_progress;
                    for (size_t t = 0; t < nthreads; t++) {
                        if (blk.fibers[t].done) continue;
                        blk.cur = (int)t;
                        emu_switch(&blk.sched, &blk.fibers[t].ctx);
                        if (blk.fibers[t].done) remaining--;
                    }
                    /* a round in which no barrier opened and no thread finished means every live thread is
                     * parked at a barrier that cannot open (divergent barrier / missing lane) */
                    if (g_progress == before) { if (++stalled > 4) { fprintf(stderr, "[emu] deadlock: block (%u,%u) stuck at a barrier\n", bx, by); abort(); } }
                    else stalled = 0;
                }
                g_blk = nullptr;
            }
}
} // namespace emu
