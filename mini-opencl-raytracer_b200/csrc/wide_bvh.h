// wide_bvh.h -- host-side construction of the GPU-resident compressed wide BVH
// from the reference's host-built arrays (CLBVHScene::m_Nodes / m_Triangles,
// produced by CLBVHnode.cpp:185-207). Pure C++ (no CUDA) so that it can also be
// unit-tested without a GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "b2rt_types.h"

namespace b2rt {

struct WideBVH {
    std::vector<WideNode> nodes;      // nodes[0] = root
    std::vector<U4> leaf;             // leaf blocks, 16-byte words
    std::vector<ShadeTri> shade;      // one per reference triangle
    uint64_t n_leaf_blocks = 0;
    uint32_t max_depth_binary = 0;    // deepest binary node (root = 0)
    uint32_t max_depth_wide = 0;      // deepest wide node (root = 0)
    uint32_t max_leaf_records = 0;
    uint64_t n_boxed_blocks = 0;         // leaf blocks that carry an explicit box (LEAF_HAS_BOX)
    uint64_t n_children = 0;          // occupied child slots, for fill statistics
    uint32_t stack_entries = 0;       // most child references any walk can have deferred at once (exact, see build_wide_bvh)
    // refit support (refit.cu): where every child box and leaf block came from in the binary tree
    std::vector<uint32_t> child_bin;  // [8 * wide node + slot] = binary node of that child, 0xFFFFFFFF = empty slot
    std::vector<uint32_t> leaf_dir;   // [2 * block] = binary leaf node, [2 * block + 1] = offset of the block in `leaf` (16-byte words)
};

// Returns an empty string on success, else a description of why the input
// arrays are not a valid flattened BVH (out-of-range offsets, cycles, ...).
std::string build_wide_bvh(const RefNode* nodes, uint64_t n_nodes, const RefTriangle* tris, uint64_t n_tris,
                           WideBVH& out);

// Worst-case traversal stack entries for this tree (see trace_wide): along one root-to-leaf path a walk defers at most
// n_children - 1 references per wide node, so the exact bound is the maximum of that sum over all paths (+1 spare).
inline uint32_t wide_stack_bound(const WideBVH& b) { return b.stack_entries + 1u; }

}  // namespace b2rt
