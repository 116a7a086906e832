// warp_emu.cpp -- TEST ONLY. Runs the 32 lanes of a warp as coroutines in lock step on one host thread, so that the
// product's warp-cooperative device code (csrc/coop.cuh: ballots, shuffles, warp reductions, shared memory between
// lanes) executes unchanged on the CPU and can be checked against the oracle without a GPU.
//
// Every collective (b2rt_emu::exchange) deposits the lane's operand and switches to the scheduler, which resumes the
// lanes round-robin: one scheduler pass advances every lane by exactly one collective, so when a lane resumes all 32
// operands of ITS collective are present (operands alternate between two buffers, because lanes resumed earlier in a
// pass already deposit for the next collective). Lanes that call different collectives, or finish while others still
// wait, are a divergence bug in the code under test and abort.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#if !defined(__x86_64__)
#error "warp_emu.cpp: the context switch is written for x86-64"
#endif

extern "C" void b2_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl b2_emu_switch
.type b2_emu_switch,@function
b2_emu_switch:
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
.size b2_emu_switch,.-b2_emu_switch
)");

namespace b2rt_emu {

struct Warp {
    void* sched_sp = nullptr;
    void* lane_sp[32];
    std::vector<char> stacks;
    int cur = 0;
    bool fin[32];
    uint32_t seq[32];
    uint32_t buf[2][32];
    int op[2][32];
    const std::function<void(uint32_t)>* body = nullptr;
};
static thread_local Warp* g_w = nullptr;

uint32_t lane() { return (uint32_t)g_w->cur; }

uint32_t exchange(uint32_t v, int op, uint32_t arg) {
    Warp* w = g_w;
    const int l = w->cur;
    const int par = (int)(w->seq[l]++ & 1u);
    w->buf[par][l] = v;
    w->op[par][l] = op;
    b2_emu_switch(&w->lane_sp[l], w->sched_sp);
    const uint32_t* b = w->buf[par];
    for (int i = 0; i < 32; ++i)
        if (w->op[par][i] != op) { std::fprintf(stderr, "warp_emu: lanes %d and %d called different collectives (%d vs %d)\n", l, i, op, w->op[par][i]); std::abort(); }
    uint32_t r = 0;
    switch (op) {
        case 0: for (int i = 0; i < 32; ++i) r |= (b[i] & 1u) << i; break;            // ballot
        case 1: r = b[arg & 31u]; break;                                               // shuffle from lane `arg`
        case 2: r = 0xffffffffu; for (int i = 0; i < 32; ++i) r = b[i] < r ? b[i] : r; break;
        case 3: for (int i = 0; i < 32; ++i) r |= b[i]; break;
        default: break;                                                                // 4: barrier only
    }
    return r;
}

static void lane_main() {
    Warp* w = g_w;
    (*w->body)((uint32_t)w->cur);
    w = g_w;
    w->fin[w->cur] = true;
    b2_emu_switch(&w->lane_sp[w->cur], w->sched_sp);
    std::abort();                                                                      // a finished lane is never resumed
}

void run_warp(const std::function<void(uint32_t)>& body) {
    static thread_local Warp warp;
    Warp* w = &warp;
    const size_t STACK = 256u << 10;
    if (w->stacks.empty()) w->stacks.resize(32 * STACK + 64);
    g_w = w;
    w->body = &body;
    for (int l = 0; l < 32; ++l) {
        w->fin[l] = false;
        w->seq[l] = 0;
        // initial frame: six callee-saved registers + the return address (16-byte aligned slot)
        uintptr_t top = ((uintptr_t)(w->stacks.data() + (size_t)(l + 1) * STACK)) & ~(uintptr_t)15;
        void** sp = reinterpret_cast<void**>(top) - 2;      // sp[0] = return address at a 16-byte aligned address
        sp[0] = reinterpret_cast<void*>(&lane_main);
        sp[1] = nullptr;
        sp -= 6;
        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
        w->lane_sp[l] = sp;
    }
    for (;;) {
        int alive = 0;
        for (int l = 0; l < 32; ++l) {
            if (w->fin[l]) continue;
            w->cur = l;
            b2_emu_switch(&w->sched_sp, w->lane_sp[l]);
            if (!w->fin[l]) ++alive;
        }
        if (alive == 0) break;
        bool any_fin = false;
        uint32_t s = 0;
        bool have = false;
        for (int l = 0; l < 32; ++l) {
            if (w->fin[l]) { any_fin = true; continue; }
            if (!have) { s = w->seq[l]; have = true; }
            else if (w->seq[l] != s) { std::fprintf(stderr, "warp_emu: lanes out of step\n"); std::abort(); }
        }
        if (any_fin) { std::fprintf(stderr, "warp_emu: some lanes finished while others wait in a collective\n"); std::abort(); }
    }
}

}  // namespace b2rt_emu
