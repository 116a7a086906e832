// wide_bvh.cpp -- collapse the reference's flattened binary BVH (CLLinearBVHNode[],
// DFS order: first child = index+1, second child = offset, leaf iff nPrimitives>0;
// CLBVHnode.cpp:161-183) into 8-wide nodes with 8-bit conservative child boxes, and
// re-pack leaf triangles as position-only records. See b2rt_types.h for the layout
// and traverse.cuh for why this keeps results identical to the reference's walk.
#include "wide_bvh.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>

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

// Append the leaf block of binary leaf `n`; returns the number of 16-byte words written.
uint32_t emit_leaf(const RefNode& n, const RefTriangle* tris, WideBVH& out) {
    size_t head = out.leaf.size();
    out.leaf.push_back(U4{ fbits(n.bmin.x), fbits(n.bmin.y), fbits(n.bmin.z), n.offset });
    out.leaf.push_back(U4{ fbits(n.bmax.x), fbits(n.bmax.y), fbits(n.bmax.z), 0u });
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
    out.leaf[head + 1].w = nrec;
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

std::string build_wide_bvh(const RefNode* nodes, uint64_t n_nodes, const RefTriangle* tris, uint64_t n_tris,
                           WideBVH& out) {
    out = WideBVH();
    if (!nodes || n_nodes == 0) return "empty BVH node array";
    if (n_nodes >= 0x7fffffffull || n_tris >= 0xfffffffeull) return "scene too large for 32-bit indices";
    if (!tris && n_tris) return "null triangle array";

    // ---- validate the flattened tree (guards the kernels against wild offsets) -----------
    for (uint64_t i = 0; i < n_nodes; ++i) {
        const RefNode& n = nodes[i];
        if (n.nPrimitives > 0) {
            if ((uint64_t)n.offset + n.nPrimitives > n_tris) return "leaf " + std::to_string(i) + " references triangles out of range";
        } else {
            if (i + 1 >= n_nodes || n.offset <= i + 1 || n.offset >= n_nodes)
                return "interior node " + std::to_string(i) + " has an invalid child offset";
            if (n.axis > 2) return "interior node " + std::to_string(i) + " has an invalid split axis";
        }
    }

    out.shade.resize(n_tris);
    for (uint64_t i = 0; i < n_tris; ++i) {
        const RefTriangle& t = tris[i];
        ShadeTri& s = out.shade[i];
        s.n1[0] = t.v1.normal.x; s.n1[1] = t.v1.normal.y; s.n1[2] = t.v1.normal.z; s.mtl = t.mtlIndex;
        s.n2[0] = t.v2.normal.x; s.n2[1] = t.v2.normal.y; s.n2[2] = t.v2.normal.z; s.pad0 = 0;
        s.n3[0] = t.v3.normal.x; s.n3[1] = t.v3.normal.y; s.n3[2] = t.v3.normal.z; s.pad1 = 0;
    }

    out.nodes.reserve(n_nodes / 3 + 16);
    out.leaf.reserve((size_t)n_tris * 3 + 16);
    std::deque<Work> queue;
    out.nodes.emplace_back();
    queue.push_back(Work{ 0u, 0u, 0u, 0u });

    while (!queue.empty()) {
        Work w = queue.front();
        queue.pop_front();
        out.max_depth_wide = std::max(out.max_depth_wide, w.depth_wide);

        // ---- gather the depth<=3 treelet under binary node w.bin ----------------------------
        Slot slots[8];
        uint8_t axis_of[7] = { 3, 3, 3, 3, 3, 3, 3 };   // 3 = treelet node absent
        struct Item { uint32_t bin, level, path; };
        Item todo[16];
        int n_todo = 0;
        todo[n_todo++] = Item{ w.bin, 0u, 0u };
        while (n_todo) {
            Item it = todo[--n_todo];
            const RefNode& n = nodes[it.bin];
            bool is_leaf = n.nPrimitives > 0;
            out.max_depth_binary = std::max(out.max_depth_binary, w.depth_bin + it.level);
            if (is_leaf || it.level == 3) {
                Slot& s = slots[it.path << (3 - it.level)];
                s.kind = is_leaf ? Slot::LEAF : Slot::INTERIOR;
                s.bin = it.bin;
                s.depth = w.depth_bin + it.level;
                continue;
            }
            uint32_t heap = it.level == 0 ? 0u : (it.level == 1 ? 1u + it.path : 3u + it.path);
            axis_of[heap] = n.axis;
            todo[n_todo++] = Item{ it.bin + 1u, it.level + 1u, (it.path << 1) | 0u };
            todo[n_todo++] = Item{ n.offset, it.level + 1u, (it.path << 1) | 1u };
        }

        // ---- quantisation frame -----------------------------------------------------------
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (const Slot& s : slots) {
            if (s.kind == Slot::EMPTY) continue;
            const RefNode& n = nodes[s.bin];
            const float bl[3] = { n.bmin.x, n.bmin.y, n.bmin.z }, bh[3] = { n.bmax.x, n.bmax.y, n.bmax.z };
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], bl[a]); hi[a] = std::max(hi[a], bh[a]); }
        }
        WideNode wn;
        std::memset(&wn, 0, sizeof(wn));
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
                if (e > 254) return "bounds too large to quantise under node " + std::to_string(w.bin);
                float s;
                uint32_t sb = (uint32_t)e << 23;
                std::memcpy(&s, &sb, 4);
                bool ok = true;
                for (int c = 0; c < 8 && ok; ++c) {
                    if (slots[c].kind == Slot::EMPTY) { wn.qlo[a][c] = 255; wn.qhi[a][c] = 0; continue; }
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

        // ---- children -----------------------------------------------------------------------
        uint32_t mx = 0, my = 0, mz = 0, valid = 0;
        for (int j = 0; j < 7; ++j) {
            if (axis_of[j] == 0) mx |= 1u << j;
            else if (axis_of[j] == 1) my |= 1u << j;
            else if (axis_of[j] == 2) mz |= 1u << j;
        }
        wn.leaf_base = (uint32_t)out.leaf.size();
        if (out.leaf.size() >= 0x7fffff00ull) return "leaf buffer exceeds 31-bit word offsets";
        bool wrap[8] = { false };
        for (int c = 0; c < 8; ++c) {
            if (slots[c].kind == Slot::EMPTY) continue;
            valid |= 1u << c;
            out.n_children++;
            if (slots[c].kind != Slot::LEAF) continue;
            size_t rel = out.leaf.size() - wn.leaf_base;
            // The root treelet may itself be a single leaf (w.bin is a leaf): it must be emitted here.
            if (rel > 255 && !(nodes[w.bin].nPrimitives > 0)) { wrap[c] = true; continue; }
            wn.meta[c] = (uint8_t)rel;
            emit_leaf(nodes[slots[c].bin], tris, out);
        }
        wn.child_base = (uint32_t)out.nodes.size();
        uint32_t imask = 0;
        for (int c = 0; c < 8; ++c) {
            if (slots[c].kind == Slot::INTERIOR || wrap[c]) {
                imask |= 1u << c;
                uint32_t idx = (uint32_t)out.nodes.size();
                out.nodes.emplace_back();
                queue.push_back(Work{ slots[c].bin, idx, slots[c].depth, w.depth_wide + 1u });
            }
        }
        wn.imask = (uint8_t)imask;
        wn.axes = mx | (my << 8) | (mz << 16) | (valid << 24);
        out.nodes[w.wide] = wn;
    }
    return std::string();
}

}  // namespace b2rt
