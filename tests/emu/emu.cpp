// emu.cpp -- TEST ONLY. Compiles the product's traversal primitives (csrc/traverse.cuh)
// and wide-BVH builder (csrc/wide_bvh.cpp) for the host CPU so their logic can be
// checked against the oracle without a GPU. Never linked into libb2rt.so and never
// reachable from the product API.
#include <cstdint>
#include <cstring>
#include <string>
#include "traverse.cuh"
#include "wide_bvh.h"

using namespace b2rt;

struct EmuRay { float ox, oy, oz, tmin, dx, dy, dz, tmax; };
struct EmuHit { float t, u, v; uint32_t tri; };
struct EmuStats {
    uint64_t n_wide, n_leaf_blocks, leaf_words, n_children, max_depth_binary, max_depth_wide, stack_bound;
    uint64_t wide_visits, leaf_blocks, leaf_pass, tri_tests, words, overflow, max_stack;
};

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
        if (any) occluded[i] = h.tri != 0xFFFFFFFFu;
        else { hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].tri = h.tri; }
        if (c.max_stack > st->max_stack) st->max_stack = c.max_stack;
        sums[0] += c.wide_nodes; sums[1] += c.leaf_blocks; sums[2] += c.leaf_pass; sums[3] += c.tri_tests; sums[4] += c.words;
    }
    (void)total;
    st->wide_visits += sums[0]; st->leaf_blocks += sums[1]; st->leaf_pass += sums[2]; st->tri_tests += sums[3]; st->words += sums[4];
}

// Histogram of children per wide node (index 0..8) of the last emu_build.
extern "C" void emu_child_histogram(uint64_t* hist9) {
    for (int i = 0; i < 9; ++i) hist9[i] = 0;
    for (const WideNode& n : g_bvh.nodes) hist9[n.n_children <= 8 ? n.n_children : 8]++;
}
