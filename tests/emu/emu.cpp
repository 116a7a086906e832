// emu.cpp -- TEST ONLY. Compiles the product's traversal primitives (csrc/traverse.cuh)
// and wide-BVH builder (csrc/wide_bvh.cpp) for the host CPU so their logic can be
// checked against the oracle without a GPU. Never linked into libb2rt.so and never
// reachable from the product API.
#include <cstdint>
#include <cstring>
#include <string>
#include <functional>
#include <thread>
#include <vector>
#include "traverse.cuh"
#include "coop.cuh"
#include "wide_bvh.h"

using namespace b2rt;

struct EmuRay { float ox, oy, oz, tmin, dx, dy, dz, tmax; };
struct EmuHit { float t, u, v; uint32_t tri; };
struct EmuStats {
    uint64_t n_wide, n_leaf_blocks, leaf_words, n_children, max_depth_binary, max_depth_wide, stack_bound;
    uint64_t wide_visits, leaf_blocks, leaf_pass, tri_tests, words, overflow, max_stack;
};

namespace b2rt { unsigned long long g_emu_stack_overflows = 0; unsigned long long g_emu_culling_violations = 0; }     // bumped by traverse.cuh's report_stack_overflow() in the host build
static WideBVH g_bvh;
static std::string g_err;

extern "C" const char* emu_build(const void* nodes, uint64_t n_nodes, const void* tris, uint64_t n_tris, EmuStats* st) {
    g_err = build_wide_bvh((const RefNode*)nodes, n_nodes, (const RefTriangle*)tris, n_tris, g_bvh);
    if (!g_err.empty()) return g_err.c_str();
    std::memset(st, 0, sizeof(*st));
    st->n_wide = g_bvh.nodes.size();
    st->n_leaf_blocks = g_bvh.n_leaf_blocks;
    st->leaf_words = g_bvh.leaf.size();
    st->n_children = g_bvh.n_children;
    st->max_depth_binary = g_bvh.max_depth_binary;
    st->max_depth_wide = g_bvh.max_depth_wide;
    st->stack_bound = wide_stack_bound(g_bvh);
    return nullptr;
}

extern "C" void emu_trace(const EmuRay* rays, uint64_t n, EmuHit* hits, uint32_t* occluded, int any, EmuStats* st, uint32_t schedule) {
    const U4* wide = reinterpret_cast<const U4*>(g_bvh.nodes.data());
    const U4* leaf = g_bvh.leaf.data();
    TravCounters total = { 0, 0, 0, 0, 0, 0 };
    uint64_t sums[5] = { 0, 0, 0, 0, 0 };
    for (uint64_t i = 0; i < n; ++i) {
        RayX r = make_ray(rays[i].ox, rays[i].oy, rays[i].oz, rays[i].dx, rays[i].dy, rays[i].dz);
        TravCounters c = { 0, 0, 0, 0, 0, 0 };
        bool overflow = false;
        uint32_t sched = schedule ? schedule + (uint32_t)i * 2654435761u : 0u;   // 0 = leaves first; else a per-ray pseudo-random interleaving
        HitX h = any ? trace_wide<true, true, 256>(wide, leaf, r, rays[i].tmax, &c, &overflow, 0x3F800000u, sched)
                     : trace_wide<false, true, 256>(wide, leaf, r, rays[i].tmax, &c, &overflow, 0x3F800000u, sched);
        if (overflow) st->overflow++;
        if (g_emu_stack_overflows) { st->overflow += g_emu_stack_overflows; g_emu_stack_overflows = 0; }
        if (any) occluded[i] = h.tri != 0xFFFFFFFFu;
        else { hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].tri = h.tri; }
        if (c.max_stack > st->max_stack) st->max_stack = c.max_stack;
        sums[0] += c.wide_nodes; sums[1] += c.leaf_blocks; sums[2] += c.leaf_pass; sums[3] += c.tri_tests; sums[4] += c.words;
    }
    (void)total;
    st->wide_visits += sums[0]; st->leaf_blocks += sums[1]; st->leaf_pass += sums[2]; st->tri_tests += sums[3]; st->words += sums[4];
}

// The kernels' stack placement (HybridStack, traverse.cuh) and the production build's branch-free push (COUNT = false), which
// emu_trace (plain array, counting build) does not run: the first `depth` entries live in a strided "shared" array, the
// rest in the local one. depth: 1, 2, 3 or 12.
template <bool ANY, int DEPTH>
static HitX hybrid_one(const U4* wide, const U4* leaf, const RayX& r, float tmax, uint32_t schedule) {
    Lane<ANY, false, 256> L;
    uint32_t local_stack[256];
    uint32_t shared_cols[DEPTH * 4 + 4];                       // stride 4: this "thread" is column 1 of four
    for (uint32_t& v : shared_cols) v = 0xDEADBEEFu;
    const HybridStack<DEPTH, 4> stack = { shared_cols + 1, local_stack };
    L.start(r, tmax);
    L.overflow = false;
    while (!L.done()) {
        const bool node = L.wants_node(), lf = L.wants_leaf();
        bool do_leaf = lf;
        if (node && lf && schedule) { schedule = schedule * 1664525u + 1013904223u; do_leaf = (schedule >> 16) & 1u; }
        if (do_leaf) { if (L.leaf_step(leaf, stack)) break; }
        else L.node_step(wide, stack, 0x3F800000u);
    }
    for (int i = 0; i < DEPTH * 4 + 4; ++i)                    // the neighbouring columns are untouched
        if ((i & 3) != 1 && shared_cols[i] != 0xDEADBEEFu) { HitX bad; bad.t = -1.0f; bad.u = bad.v = 0.0f; bad.tri = 0xBADBADu; return bad; }
    return L.h;
}
template <bool ANY>
static HitX hybrid_depth(int depth, const U4* wide, const U4* leaf, const RayX& r, float tmax, uint32_t schedule) {
    switch (depth) {
        case 1: return hybrid_one<ANY, 1>(wide, leaf, r, tmax, schedule);
        case 2: return hybrid_one<ANY, 2>(wide, leaf, r, tmax, schedule);
        case 3: return hybrid_one<ANY, 3>(wide, leaf, r, tmax, schedule);
        default: return hybrid_one<ANY, 12>(wide, leaf, r, tmax, schedule);
    }
}
extern "C" void emu_trace_hybrid(const EmuRay* rays, uint64_t n, EmuHit* hits, uint32_t* occluded, int any, int depth, uint32_t schedule) {
    const U4* wide = reinterpret_cast<const U4*>(g_bvh.nodes.data());
    const U4* leaf = g_bvh.leaf.data();
    for (uint64_t i = 0; i < n; ++i) {
        const RayX r = make_ray(rays[i].ox, rays[i].oy, rays[i].oz, rays[i].dx, rays[i].dy, rays[i].dz);
        const uint32_t sched = schedule ? schedule + (uint32_t)i * 2654435761u : 0u;
        const HitX h = any ? hybrid_depth<true>(depth, wide, leaf, r, rays[i].tmax, sched) : hybrid_depth<false>(depth, wide, leaf, r, rays[i].tmax, sched);
        if (any) occluded[i] = h.tri != 0xFFFFFFFFu;
        else { hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].tri = h.tri; }
    }
}

// Children that the exact-arithmetic node test lets through but the fast one culled, over all emulated walks so far (must be 0).
extern "C" uint64_t emu_culling_violations() { return g_emu_culling_violations; }

// Histogram of children per wide node (index 0..8) of the last emu_build.
extern "C" void emu_child_histogram(uint64_t* hist9) {
    for (int i = 0; i < 9; ++i) hist9[i] = 0;
    for (const WideNode& n : g_bvh.nodes) hist9[n.n_children <= 8 ? n.n_children : 8]++;
}

// ---- warp-cooperative tail mode (csrc/coop.cuh) on 32 emulated lanes -------------------------------------------
namespace b2rt_emu { void run_warp(const std::function<void(uint32_t)>& body); }

template <bool ANY>
static void coop_one(const U4* wide, const U4* leaf, const EmuRay& in, uint32_t handoff, uint32_t wide_limit, uint32_t fcap,
                     uint32_t seed, HitX& out, EmuStats* st, uint32_t resume) {
    const RayX r = make_ray(in.ox, in.oy, in.oz, in.dx, in.dy, in.dz);
    Lane<ANY, true, 256> L;
    uint32_t stack[256];
    L.tc = TravCounters{ 0, 0, 0, 0, 0, 0 };
    L.start(r, in.tmax);
    // solo steps first (a random number of them under a random interleaving), then the hand-over
    uint32_t steps = handoff ? seed % handoff : 0u;
    uint32_t sched = seed | 1u;
    while (steps-- && !L.done()) {
        const bool node = L.wants_node(), lf = L.wants_leaf();
        bool do_leaf = lf;
        if (node && lf) { sched = sched * 1664525u + 1013904223u; do_leaf = (sched >> 16) & 1u; }
        if (do_leaf) { if (L.leaf_step(leaf, stack)) break; }
        else L.node_step(wide, stack, 0x3F800000u);
    }
    if (L.done()) { out = L.h; return; }
    std::vector<uint32_t> F(fcap + 64);
    uint32_t n = coop_dump(L, stack, F.data());
    if (resume) {
        // two-step tail: another lane picks the suspended ray up from the record (Lane::resume), walks on alone for a
        // random number of steps and suspends it again for the cooperative finish
        Lane<ANY, true, 256> L2;
        uint32_t stack2[256];
        L2.tc = L.tc;
        L2.start(r, in.tmax);
        L2.h = L.h;
        L2.resume(stack2, F.data(), n);
        steps = (seed >> 8) % resume;
        while (steps-- && !L2.done()) {
            const bool node = L2.wants_node(), lf = L2.wants_leaf();
            bool do_leaf = lf;
            if (node && lf) { sched = sched * 1664525u + 1013904223u; do_leaf = (sched >> 16) & 1u; }
            if (do_leaf) { if (L2.leaf_step(leaf, stack2)) break; }
            else L2.node_step(wide, stack2, 0x3F800000u);
        }
        if (L2.done()) { out = L2.h; return; }
        n = coop_dump(L2, stack2, F.data());
        L.h = L2.h;
    }
    HitX res[32];
    bool ovf[32];
    b2rt_emu::run_warp([&](uint32_t lane) {
        HitX h = L.h;
        TravCounters tc = { 0, 0, 0, 0, 0, 0 };
        bool o = false;
        coop_trace<ANY, true>(wide, leaf, F.data(), n, fcap, wide_limit, r, h, tc, o);
        res[lane] = h;
        ovf[lane] = o;
        if (lane == 0) { st->wide_visits += tc.wide_nodes; st->leaf_blocks += tc.leaf_blocks; st->tri_tests += tc.tri_tests; if (tc.max_stack > st->max_stack) st->max_stack = tc.max_stack; }
    });
    for (int l = 1; l < 32; ++l)
        if (std::memcmp(&res[l], &res[0], sizeof(HitX)) != 0 || ovf[l] != ovf[0]) st->overflow += 1000000;   // lanes must agree
    if (ovf[0]) st->overflow++;
    out = res[0];
}

extern "C" void emu_trace_coop(const EmuRay* rays, uint64_t n, EmuHit* hits, uint32_t* occluded, int any, EmuStats* st,
                               uint32_t handoff, uint32_t wide_limit, uint32_t fcap, uint32_t resume) {
    const U4* wide = reinterpret_cast<const U4*>(g_bvh.nodes.data());
    const U4* leaf = g_bvh.leaf.data();
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 16) nt = 16;
    std::vector<EmuStats> part(nt);
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nt; ++w)
        pool.emplace_back([&, w]() {
            std::memset(&part[w], 0, sizeof(EmuStats));
            for (uint64_t i = n * w / nt; i < n * (w + 1) / nt; ++i) {
                HitX h;
                const uint32_t seed = (uint32_t)i * 2654435761u + 12345u;
                if (any) coop_one<true>(wide, leaf, rays[i], handoff, wide_limit, fcap, seed, h, &part[w], resume);
                else coop_one<false>(wide, leaf, rays[i], handoff, wide_limit, fcap, seed, h, &part[w], resume);
                if (any) occluded[i] = h.tri != 0xFFFFFFFFu;
                else { hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].tri = h.tri; }
            }
        });
    for (auto& t : pool) t.join();
    for (const EmuStats& p : part) {
        st->wide_visits += p.wide_visits; st->leaf_blocks += p.leaf_blocks; st->tri_tests += p.tri_tests; st->overflow += p.overflow;
        if (p.max_stack > st->max_stack) st->max_stack = p.max_stack;
    }
}
