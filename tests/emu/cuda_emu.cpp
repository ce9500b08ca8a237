// tests/emu/cuda_emu.cpp — TEST INFRASTRUCTURE: fiber scheduler of the lock-step CUDA emulator (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

Cta *g = nullptr;

// void emu_switch(void **save_sp, void *load_sp): saves the callee-saved registers on the current stack, stores the stack
// pointer, loads the other stack and returns into it (x86-64 SysV).
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

static void trampoline()
{
    g->body();
    Fiber *f = g->cur;
    f->state = DONE;
    emu_switch(&f->sp, g->sched_sp);
    abort(); // a finished fiber is never resumed
}

static constexpr size_t STACK = 256 * 1024;

static void run_cta(Cta &c)
{
    g = &c;
    const unsigned nw = (c.nthreads + 31) / 32;
    c.fibers.assign(c.nthreads, Fiber());
    c.warps.assign(nw, Warp());
    for (unsigned t = 0; t < c.nthreads; ++t) {
        Fiber &f = c.fibers[t];
        f.tid = t;
        f.stack = (char *)aligned_alloc(64, STACK);
        uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
        void **sp = (void **)top;
        *--sp = nullptr;               // fake return address of the trampoline
        *--sp = (void *)&trampoline;   // 'ret' of emu_switch jumps here
        for (int i = 0; i < 6; ++i) *--sp = nullptr;
        f.sp = sp;
    }
    unsigned done = 0;
    for (;;) {
        bool progress = false;
        for (auto &f : c.fibers) // lanes that were spinning on a flag look at it again
            if (f.state == YIELDED) {
                f.state = RUNNABLE;
                progress = true;
            }
        for (unsigned w = 0; w < nw; ++w) {
            const unsigned l0 = w * 32, l1 = std::min(l0 + 32, c.nthreads);
            for (;;) {
                bool ran = false;
                for (unsigned t = l0; t < l1; ++t) {
                    Fiber &f = c.fibers[t];
                    if (f.state != RUNNABLE) continue;
                    c.cur = &f;
                    emu_switch(&c.sched_sp, f.sp);
                    ran = true;
                    progress = true;
                    if (f.state == DONE) ++done;
                }
                // release a completed warp collective
                unsigned waiting = 0, alive = 0;
                for (unsigned t = l0; t < l1; ++t) {
                    if (c.fibers[t].state != DONE) ++alive;
                    if (c.fibers[t].state == WAIT_WARP) ++waiting;
                }
                // release the teams (__syncwarp with a partial mask) whose lanes have all arrived
                for (unsigned t = l0; t < l1; ++t) {
                    Fiber &f = c.fibers[t];
                    if (f.state != WAIT_TEAM) continue;
                    bool all = true;
                    for (unsigned u = l0; u < l1; ++u)
                        if ((f.team_mask >> (u - l0)) & 1u) {
                            const Fiber &o = c.fibers[u];
                            if (!(o.state == DONE || (o.state == WAIT_TEAM && o.team_mask == f.team_mask))) all = false;
                        }
                    if (all) {
                        const unsigned m = f.team_mask;
                        for (unsigned u = l0; u < l1; ++u)
                            if (((m >> (u - l0)) & 1u) && c.fibers[u].state == WAIT_TEAM) c.fibers[u].state = RUNNABLE;
                        progress = true;
                        ran = true;
                    }
                }
                if (waiting && waiting == alive) {
                    for (unsigned t = l0; t < l1; ++t)
                        if (c.fibers[t].state == WAIT_WARP) c.fibers[t].state = RUNNABLE;
                    c.warps[w].arrived = 0;
                    progress = true;
                    continue;
                }
                if (!ran) break;
            }
        }
        // named barrier (bar.sync 1, n)
        if (c.named_arrived && c.named_arrived == c.named_need) {
            for (auto &f : c.fibers)
                if (f.state == WAIT_NAMED) f.state = RUNNABLE;
            c.named_arrived = 0;
            progress = true;
        }
        // CTA barrier: every thread that has not exited
        if (c.cta_arrived && c.cta_arrived == c.nthreads - done) {
            c.cta_or_result = c.cta_or;
            c.cta_or = 0;
            c.cta_arrived = 0;
            for (auto &f : c.fibers)
                if (f.state == WAIT_CTA) f.state = RUNNABLE;
            progress = true;
        }
        if (done == c.nthreads) break;
        if (!progress) {
            unsigned ww = 0, wc = 0, wn = 0;
            for (auto &f : c.fibers) {
                ww += f.state == WAIT_WARP;
                wc += f.state == WAIT_CTA;
                wn += f.state == WAIT_NAMED;
            }
            fprintf(stderr, "[cuda_emu] deadlock in CTA %u: %u at a warp collective, %u at __syncthreads, %u at the named barrier, %u done\n",
                    c.bid, ww, wc, wn, done);
            for (unsigned w = 0; w < nw; ++w) {
                unsigned a = 0, b = 0, d = 0, n = 0;
                for (unsigned t = w * 32; t < std::min(w * 32 + 32, c.nthreads); ++t) {
                    a += c.fibers[t].state == WAIT_WARP;
                    b += c.fibers[t].state == WAIT_CTA;
                    n += c.fibers[t].state == WAIT_NAMED;
                    d += c.fibers[t].state == DONE;
                }
                if (a) fprintf(stderr, "  warp %u: %u warp-collective, %u cta, %u named, %u done\n", w, a, b, n, d);
            }
            abort();
        }
    }
    for (auto &f : c.fibers) free(f.stack);
    g = nullptr;
}

void launch(unsigned grid, unsigned threads, std::function<void()> body)
{
    for (unsigned b = 0; b < grid; ++b) {
        Cta c;
        c.nthreads = threads;
        c.bid = b;
        c.grid = grid;
        c.body = body;
        run_cta(c);
    }
}

} // namespace emu
