// wide_bvh.cpp -- collapse the reference's flattened binary BVH (CLLinearBVHNode[],
// DFS order: first child = index+1, second child = offset, leaf iff nPrimitives>0;
// CLBVHnode.cpp:161-183) into up-to-8-wide treelet nodes with 8-bit conservative child
// boxes and per-octant visiting orders, and re-pack leaf triangles as position-only
// records. See b2rt_types.h for the layout and traverse.cuh for why this keeps results
// identical to the reference's walk.
#include "wide_bvh.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <thread>

namespace b2rt {
namespace {

struct Slot {
    enum Kind { EMPTY, LEAF, INTERIOR } kind = EMPTY;
    uint32_t bin = 0;
    uint32_t depth = 0;
};

struct Work { uint32_t bin, wide, depth_bin, depth_wide; };

inline bool same_pos(const RefVec& a, const RefVec& b) {
    return std::memcmp(&a.x, &b.x, 12) == 0;   // bitwise: the two copies must give bit-identical arithmetic
}
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// min / max the way the kernels evaluate them (fminf / fmaxf of three); only used where no operand is NaN
inline float min3(float a, float b, float c) { return std::fmin(std::fmin(a, b), c); }
inline float max3(float a, float b, float c) { return std::fmax(std::fmax(a, b), c); }

// Append the leaf block of binary leaf `n`; returns the number of 16-byte words written.
uint32_t emit_leaf(const RefNode& n, uint32_t bin, const RefTriangle* tris, WideBVH& out) {
    size_t head = out.leaf.size();
    out.leaf_dir.push_back(bin);
    out.leaf_dir.push_back((uint32_t)head);
    uint32_t nrec = 0;
    uint32_t end = n.offset + n.nPrimitives;
    for (uint32_t i = n.offset; i < end;) {
        const RefTriangle& t = tris[i];
        uint32_t flags = REC_SINGLE;
        if (i + 1 < end) {
            const RefTriangle& s = tris[i + 1];
            if (same_pos(s.v1.position, t.v2.position) && same_pos(s.v2.position, t.v3.position) &&
                same_pos(s.v3.position, t.v1.position))
                flags = REC_ROT_LEFT;
            else if (same_pos(s.v1.position, t.v3.position) && same_pos(s.v2.position, t.v1.position) &&
                     same_pos(s.v3.position, t.v2.position))
                flags = REC_ROT_RIGHT;
        }
        out.leaf.push_back(U4{ fbits(t.v1.position.x), fbits(t.v1.position.y), fbits(t.v1.position.z), flags });
        out.leaf.push_back(U4{ fbits(t.v2.position.x), fbits(t.v2.position.y), fbits(t.v2.position.z), 0u });
        out.leaf.push_back(U4{ fbits(t.v3.position.x), fbits(t.v3.position.y), fbits(t.v3.position.z), 0u });
        ++nrec;
        i += flags ? 2u : 1u;
    }
    // one record whose node box is exactly the min / max of its three vertices (what the reference's builder produces):
    // the kernels recompute that box, none is stored. Comparison by value: the sign of a zero never changes RayBounds.
    bool implied = nrec == 1;
    if (implied) {
        const RefTriangle& t = tris[n.offset];
        const float* a = &t.v1.position.x; const float* b = &t.v2.position.x; const float* c = &t.v3.position.x;
        const float* lo = &n.bmin.x; const float* hi = &n.bmax.x;
        for (int k = 0; k < 3 && implied; ++k) {
            if (std::isnan(a[k]) || std::isnan(b[k]) || std::isnan(c[k])) implied = false;
            else implied = min3(a[k], b[k], c[k]) == lo[k] && max3(a[k], b[k], c[k]) == hi[k];
        }
    }
    if (nrec > LEAF_NREC_MASK) nrec = LEAF_NREC_MASK;       // cannot happen: nPrimitives is 16 bits
    out.leaf[head].w |= (nrec << LEAF_NREC_SHIFT) | (implied ? 0u : LEAF_HAS_BOX);
    out.leaf[head + 1].w = n.offset;
    if (!implied) {
        out.leaf.push_back(U4{ fbits(n.bmin.x), fbits(n.bmin.y), fbits(n.bmin.z), 0u });
        out.leaf.push_back(U4{ fbits(n.bmax.x), fbits(n.bmax.y), fbits(n.bmax.z), 0u });
        out.n_boxed_blocks++;
    }
    out.n_leaf_blocks++;
    out.max_leaf_records = std::max(out.max_leaf_records, nrec);
    return (uint32_t)(out.leaf.size() - head);
}

// Largest q in [0,255] with base + q*s <= lo (as real numbers), resp. smallest q with base + q*s >= hi.
int quant_floor(float base, float s, float lo) {
    long double d = ((long double)lo - (long double)base) / (long double)s;
    long double q = std::floor(d);
    int exp_lo, exp_base;
    std::frexp(lo, &exp_lo); std::frexp(base, &exp_base);
    if (std::abs(exp_lo - exp_base) > 36 && lo != 0.0f && base != 0.0f) q -= 1;   // difference was not exact: stay safe
    if (q < 0) q = 0;   // lo >= base always, so 0 is conservative
    if (q > 255) q = 255;
    return (int)q;
}
int quant_ceil(float base, float s, float hi) {
    long double d = ((long double)hi - (long double)base) / (long double)s;
    long double q = std::ceil(d);
    int exp_hi, exp_base;
    std::frexp(hi, &exp_hi); std::frexp(base, &exp_base);
    if (std::abs(exp_hi - exp_base) > 36 && hi != 0.0f && base != 0.0f) q += 1;
    if (q < 0) q = 0;
    return (int)std::min<long double>(q, 1000.0L);   // > 255 tells the caller to widen the grid
}

}  // namespace

namespace {

// Treelet under construction: a small binary tree whose terminals become the child slots.
struct TNode { int left = -1, right = -1; uint32_t bin = 0; uint32_t depth = 0; uint8_t axis = 0; int slot = -1; };

inline float half_area(const RefNode& n) {
    float dx = n.bmax.x - n.bmin.x, dy = n.bmax.y - n.bmin.y, dz = n.bmax.z - n.bmin.z;
    return dx * dy + dx * dz + dy * dz;
}

// Optimal treelet cuts (surface-area heuristic, after Ylitie et al. 2017, restricted to the host's
// fixed binary topology and atomic leaves): the expected number of wide-node visits of a random
// ray is proportional to the summed surface area of the binary nodes at which wide nodes are
// rooted. dist(n,k) = least such sum for the subtree of n when it may occupy at most k child slots
// of its parent wide node:
//   leaf:      dist(n,k) = 0
//   interior:  dist(n,1) = own(n) = area(n) + min_i dist(first,i) + dist(second,8-i)      (n roots a wide node)
//              dist(n,k) = min(dist(n,k-1), min_i dist(first,i) + dist(second,k-i))       (n is opened in place)
// Children follow their parent in the flattened array, so one reverse sweep fills the table.
struct CutPlan {
    std::vector<float> dist;         // [node*8 + (k-1)]
    std::vector<uint8_t> split;      // [node*8 + (k-1)]: slots given to the first child when n is opened with k slots, 0 = closed
    std::vector<uint8_t> open_root;  // [node]: slots given to the first child when n roots a wide node
    uint32_t split_of(uint32_t n, uint32_t k) const { return split[(size_t)n * 8 + (k - 1)]; }
};

void plan_cuts(const RefNode* nodes, uint64_t n_nodes, CutPlan& plan) {
    plan.dist.assign((size_t)n_nodes * 8, 0.0f);
    plan.split.assign((size_t)n_nodes * 8, 0);
    plan.open_root.assign((size_t)n_nodes, 0);
    for (uint64_t idx = n_nodes; idx-- > 0;) {
        const RefNode& n = nodes[idx];
        if (n.nPrimitives > 0) continue;
        const float* a = &plan.dist[(size_t)(idx + 1) * 8];
        const float* b = &plan.dist[(size_t)n.offset * 8];
        float* d = &plan.dist[(size_t)idx * 8];
        uint8_t* sp = &plan.split[(size_t)idx * 8];
        float area = half_area(n);
        if (!(area >= 0.0f)) area = 0.0f;
        // best way to spend k slots on the two children, k = 2..8
        float open_cost[9];
        uint8_t open_split[9];
        for (uint32_t k = 2; k <= 8; ++k) {
            float best = INFINITY;
            uint8_t arg = 1;
            for (uint32_t i = 1; i < k; ++i) {
                float c = a[i - 1] + b[k - i - 1];
                if (c < best) { best = c; arg = (uint8_t)i; }
            }
            open_cost[k] = best;
            open_split[k] = arg;
        }
        plan.open_root[idx] = open_split[8];
        d[0] = area + open_cost[8];
        sp[0] = 0;
        for (uint32_t k = 2; k <= 8; ++k) {
            if (open_cost[k] < d[k - 2]) { d[k - 1] = open_cost[k]; sp[k - 1] = open_split[k]; }
            else { d[k - 1] = d[k - 2]; sp[k - 1] = sp[k - 2]; }
        }
    }
}

// Visiting order of the terminals under t for sign octant `oct` (kernel_bvh.cl:200-207):
// second child first iff the ray direction is negative on the node's split axis.
void visit_order(const TNode* t, int i, uint32_t oct, uint32_t& word, int& rank) {
    if (t[i].left < 0) { word |= (uint32_t)t[i].slot << (4 * rank++); return; }
    bool flip = (oct >> t[i].axis) & 1u;
    visit_order(t, flip ? t[i].right : t[i].left, oct, word, rank);
    visit_order(t, flip ? t[i].left : t[i].right, oct, word, rank);
}

void number_slots(TNode* t, int i, int& next) {
    if (t[i].left < 0) { t[i].slot = next++; return; }
    number_slots(t, t[i].left, next);
    number_slots(t, t[i].right, next);
}

}  // namespace

std::string build_wide_bvh(const RefNode* nodes, uint64_t n_nodes, const RefTriangle* tris, uint64_t n_tris,
                           WideBVH& out) {
    out = WideBVH();
    if (!nodes || n_nodes == 0) return "empty BVH node array";
    if (n_nodes >= 0x7fffffffull || n_tris >= 0xfffffffeull) return "scene too large for 32-bit indices";
    if (!tris && n_tris) return "null triangle array";

    // ---- validate the flattened tree (guards the kernels against wild offsets) -----------
    // It must be a proper pre-order tree: the first child's subtree ends exactly where the second child starts and the
    // root's subtree is the whole array (every node reached exactly once). Children follow their parent, so one reverse
    // sweep computes every subtree's end. Child boxes must lie inside their parent's box: the argument that the leaf
    // gate alone reproduces the reference's culling (traverse.cuh) rests on it.
    {
        std::vector<uint32_t> end(n_nodes);
        auto inside = [](const RefNode& c, const RefNode& p) {
            return c.bmin.x >= p.bmin.x && c.bmin.y >= p.bmin.y && c.bmin.z >= p.bmin.z &&
                   c.bmax.x <= p.bmax.x && c.bmax.y <= p.bmax.y && c.bmax.z <= p.bmax.z;
        };
        for (uint64_t i = n_nodes; i-- > 0;) {
            const RefNode& n = nodes[i];
            if (n.nPrimitives > 0) {
                if ((uint64_t)n.offset + n.nPrimitives > n_tris) return "leaf " + std::to_string(i) + " references triangles out of range";
                end[i] = (uint32_t)(i + 1);
            } else {
                if (i + 1 >= n_nodes || n.offset <= i + 1 || n.offset >= n_nodes)
                    return "interior node " + std::to_string(i) + " has an invalid child offset";
                if (n.axis > 2) return "interior node " + std::to_string(i) + " has an invalid split axis";
                if (end[i + 1] != n.offset)
                    return "interior node " + std::to_string(i) + ": the first child's subtree does not end at the second child (not a pre-order tree)";
                if (!inside(nodes[i + 1], n) || !inside(nodes[n.offset], n))
                    return "interior node " + std::to_string(i) + " does not contain its children's bounds";
                end[i] = end[n.offset];
            }
        }
        if (end[0] != n_nodes) return "node array holds " + std::to_string(n_nodes - end[0]) + " nodes outside the root's subtree";
    }

    // The per-triangle shading records are independent of everything below: a helper thread fills them while this
    // thread plans the cuts and emits the wide nodes.
    out.shade.resize(n_tris);
    ShadeTri* shade = out.shade.data();
    std::thread shade_fill([shade, tris, n_tris]() {
        for (uint64_t i = 0; i < n_tris; ++i) {
            const RefTriangle& t = tris[i];
            ShadeTri& s = shade[i];
            s.n1[0] = t.v1.normal.x; s.n1[1] = t.v1.normal.y; s.n1[2] = t.v1.normal.z; s.mtl = t.mtlIndex;
            s.n2[0] = t.v2.normal.x; s.n2[1] = t.v2.normal.y; s.n2[2] = t.v2.normal.z; s.pad0 = 0;
            s.n3[0] = t.v3.normal.x; s.n3[1] = t.v3.normal.y; s.n3[2] = t.v3.normal.z; s.pad1 = 0;
        }
    });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } join_shade{ shade_fill };

    out.nodes.reserve(n_nodes / 5 + 16);
    out.leaf.reserve((size_t)n_tris * 3 + 16);
    CutPlan plan;
    plan_cuts(nodes, n_nodes, plan);
    std::deque<Work> queue;
    out.nodes.emplace_back();
    queue.push_back(Work{ 0u, 0u, 0u, 0u });

    while (!queue.empty()) {
        Work w = queue.front();
        queue.pop_front();
        out.max_depth_wide = std::max(out.max_depth_wide, w.depth_wide);

        // ---- cut the treelet under binary node w.bin along the optimal partition (see plan_cuts) ---
        TNode t[15];
        int n_t = 1;
        t[0].bin = w.bin; t[0].depth = w.depth_bin;
        if (nodes[w.bin].nPrimitives == 0) {
            struct Open { int t; uint32_t slots; };
            Open todo[16];
            int n_todo = 0;
            todo[n_todo++] = Open{ 0, 8u };
            bool root = true;
            while (n_todo) {
                Open o = todo[--n_todo];
                const RefNode& n = nodes[t[o.t].bin];
                if (n.nPrimitives > 0) continue;                          // binary leaf: a leaf slot
                uint32_t give_first = root ? plan.open_root[t[o.t].bin] : plan.split_of(t[o.t].bin, o.slots);
                root = false;
                if (give_first == 0) continue;                             // stays closed: becomes its own wide node
                t[o.t].axis = n.axis;
                t[o.t].left = n_t; t[o.t].right = n_t + 1;
                t[n_t].bin = t[o.t].bin + 1u; t[n_t].depth = t[o.t].depth + 1u;
                t[n_t + 1].bin = n.offset; t[n_t + 1].depth = t[o.t].depth + 1u;
                todo[n_todo++] = Open{ n_t, give_first };
                todo[n_todo++] = Open{ n_t + 1, o.slots - give_first };
                n_t += 2;
            }
        }
        int next_slot = 0;
        number_slots(t, 0, next_slot);
        Slot slots[8];
        for (int i = 0; i < n_t; ++i) {
            out.max_depth_binary = std::max(out.max_depth_binary, t[i].depth);
            if (t[i].left >= 0) continue;
            Slot& s = slots[t[i].slot];
            s.kind = nodes[t[i].bin].nPrimitives > 0 ? Slot::LEAF : Slot::INTERIOR;
            s.bin = t[i].bin;
            s.depth = t[i].depth;
        }
        const int n_children = next_slot;

        // ---- quantisation frame -----------------------------------------------------------
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (int c = 0; c < n_children; ++c) {
            const RefNode& n = nodes[slots[c].bin];
            const float bl[3] = { n.bmin.x, n.bmin.y, n.bmin.z }, bh[3] = { n.bmax.x, n.bmax.y, n.bmax.z };
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], bl[a]); hi[a] = std::max(hi[a], bh[a]); }
        }
        WideNode wn;
        std::memset(&wn, 0, sizeof(wn));
        wn.n_children = (uint8_t)n_children;
        for (int a = 0; a < 3; ++a) {
            if (!(lo[a] <= hi[a]) || !std::isfinite(lo[a]) || !std::isfinite(hi[a]))
                return "non-finite or inverted bounds under node " + std::to_string(w.bin);
            wn.base[a] = lo[a];
            int e = 1;
            long double extent = (long double)hi[a] - (long double)lo[a];
            if (extent > 0) {
                int k;
                (void)std::frexp((double)(extent / 255.0L), &k);   // extent/255 = m * 2^k, m in [0.5,1)
                e = std::min(std::max(k + 127, 1), 254);
            }
            for (;; ++e) {
                if (e > 230) return "bounds too large to quantise under node " + std::to_string(w.bin);
                float s;
                uint32_t sb = (uint32_t)e << 23;
                std::memcpy(&s, &sb, 4);
                bool ok = true;
                for (int c = 0; c < 8 && ok; ++c) {
                    if (c >= n_children) { wn.qlo[a][c] = 255; wn.qhi[a][c] = 0; continue; }
                    const RefNode& n = nodes[slots[c].bin];
                    float bl = a == 0 ? n.bmin.x : (a == 1 ? n.bmin.y : n.bmin.z);
                    float bh = a == 0 ? n.bmax.x : (a == 1 ? n.bmax.y : n.bmax.z);
                    int ql = quant_floor(lo[a], s, bl), qh = quant_ceil(lo[a], s, bh);
                    if (qh > 255) { ok = false; break; }
                    wn.qlo[a][c] = (uint8_t)ql;
                    wn.qhi[a][c] = (uint8_t)qh;
                }
                if (ok) { wn.exp[a] = (uint8_t)e; break; }
            }
        }

        // ---- visiting order per sign octant ------------------------------------------------------
        for (uint32_t oct = 0; oct < 8; ++oct) {
            uint32_t word = 0;
            int rank = 0;
            visit_order(t, 0, oct, word, rank);
            // unused ranks name an empty slot (7 is empty whenever n_children < 8)
            for (; rank < 8; ++rank) word |= 7u << (4 * rank);
            wn.order[oct] = word;
        }

        // ---- children -----------------------------------------------------------------------
        wn.leaf_base = (uint32_t)out.leaf.size();
        if (out.leaf.size() >= 0x7fffff00ull) return "leaf buffer exceeds 31-bit word offsets";
        bool wrap[8] = { false };
        for (int c = 0; c < n_children; ++c) {
            out.n_children++;
            if (slots[c].kind != Slot::LEAF) continue;
            size_t rel = out.leaf.size() - wn.leaf_base;
            // A treelet that is a single binary leaf (w.bin itself) must emit it here (rel == 0).
            if (rel > META_MAX_LEAF_OFFSET && n_children > 1) { wrap[c] = true; continue; }
            if (rel > META_MAX_LEAF_OFFSET) return "internal error: lone leaf not at offset 0";
            wn.meta[c] = (uint8_t)rel;
            emit_leaf(nodes[slots[c].bin], slots[c].bin, tris, out);
        }
        wn.child_base = (uint32_t)out.nodes.size();
        uint32_t n_interior = 0;
        for (int c = 0; c < n_children; ++c) {
            if (slots[c].kind == Slot::INTERIOR || wrap[c]) {
                wn.meta[c] = (uint8_t)(META_INTERIOR | n_interior++);
                uint32_t idx = (uint32_t)out.nodes.size();
                out.nodes.emplace_back();
                queue.push_back(Work{ slots[c].bin, idx, slots[c].depth, w.depth_wide + 1u });
            }
        }
        out.nodes[w.wide] = wn;
        if (out.child_bin.size() < out.nodes.size() * 8) out.child_bin.resize(out.nodes.size() * 8, 0xFFFFFFFFu);
        for (int c = 0; c < n_children; ++c) out.child_bin[(size_t)w.wide * 8 + c] = slots[c].bin;
    }
    out.child_bin.resize(out.nodes.size() * 8, 0xFFFFFFFFu);
    // The cooperative tail kernel fetches a block's SECOND record before it knows the record count (coop.cuh): keep 64 readable
    // bytes behind the last block.
    for (int i = 0; i < 4; ++i) out.leaf.push_back(U4{ 0u, 0u, 0u, 0u });
    // exact stack bound: children always follow their parent in `nodes`, so one reverse sweep is bottom-up
    {
        std::vector<uint32_t> need(out.nodes.size(), 0);
        for (size_t i = out.nodes.size(); i-- > 0;) {
            const WideNode& n = out.nodes[i];
            uint32_t below = 0;
            for (int c = 0; c < n.n_children; ++c)
                if (n.meta[c] & META_INTERIOR) below = std::max(below, need[n.child_base + (n.meta[c] & 0x7f)]);
            need[i] = (n.n_children ? n.n_children - 1u : 0u) + below;
        }
        out.stack_entries = need.empty() ? 0 : need[0];
    }
    return std::string();
}

}  // namespace b2rt
