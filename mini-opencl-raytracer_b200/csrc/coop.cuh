// coop.cuh -- warp-cooperative traversal of ONE ray: the tail mode of trace_persistent (kernels.cu).
//
// A persistent traversal launch ends with a latency-bound tail: when the ray pool is dry, the few rays
// still alive advance one dependent step (a whole 8-child node test or a leaf, plus a memory round trip)
// at a time on one lane each while the other lanes idle. A multi-GPU frame share or a wavefront stage is
// short enough for that tail to dominate. Here all 32 lanes work on ONE ray at a time instead:
//
//  * the ray's pending work -- its queued leaves, current node and short stack -- is laid out as one
//    FRONTIER in shared memory, in the reference's depth-first order (kernel_bvh.cl:182-216), front at the
//    highest index;
//  * leaves at the front are resolved in parallel: one lane per (one-record leaf block, loader copy) evaluates the
//    reference's exact Moller-Trumbore sequence; results are committed leaf by leaf in frontier order,
//    every leaf gated by its exact fp32 box against the `best` of that moment -- the reference's own
//    sequence of decisions, so ties and negative t resolve as in the solo walk (traverse.cuh);
//  * the first B2_COOP_NODES (8) interior entries of the frontier are expanded in one round, four lanes per
//    node, two children per lane in the reference's visiting order (same conservative quantised slab test
//    as test_wide_node); passing children replace their parent in place, keeping the order. Interior
//    nodes behind unresolved leaves are thereby tested with a stale (larger) `best`, which can only let
//    more children through -- allowed for the same reason as the solo walk's speculation.
//
// The same text runs on the CPU in tests/emu (32 coroutine lanes in lock step) against the oracle.
#pragma once
#include "traverse.cuh"

#ifndef B2_COOP_NODES
#define B2_COOP_NODES 8         // interior entries expanded per round: 4 (eight lanes per node, one child per lane) or 8 (four lanes, two children)
#endif
#define COOP_LPN (32 / B2_COOP_NODES)      // lanes per expanded node
#define COOP_CPL (B2_COOP_NODES / 4)       // children per lane

namespace b2rt {

// Monotone float -> uint32 key for a warp min-reduction; -0 is folded onto +0 (they compare equal, and the
// reference keeps the FIRST of equal candidates: ties are then broken by the lowest lane).
B2_HD uint32_t order_key(float t) {
    const uint32_t b = f2bits(xadd(t, 0.0f));
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
B2_HD uint32_t lanes_below(uint32_t l) { return l >= 32u ? 0xffffffffu : (1u << l) - 1u; }
B2_HD uint32_t lanes_upto(uint32_t l) { return 0xffffffffu >> (31u - (l & 31u)); }

// Child slot `slot` of a wide node against [0, best]: the arithmetic of test_wide_node for one child.
B2_HD bool coop_child_test(const U4& w0, const U4& w2, const U4& w3, const U4& w4, uint32_t slot, const RayX& r, float best) {
    const float o[3] = { r.ox, r.oy, r.oz };
    const float inv[3] = { r.ix, r.iy, r.iz };
    const float base[3] = { bits2f(w0.x), bits2f(w0.y), bits2f(w0.z) };
    const uint32_t qlo_a[3] = { w2.x, w2.z, w3.x }, qlo_b[3] = { w2.y, w2.w, w3.y };
    const uint32_t qhi_a[3] = { w3.z, w4.x, w4.z }, qhi_b[3] = { w3.w, w4.y, w4.w };
    if (r.sign & RAY_DEGENERATE) {          // see test_wide_node_robust (warp-uniform: every lane holds the same ray)
        float N = 0.0f, F = best;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float S = bits2f(((w0.w >> (8 * a)) & 0xffu) << 23);
            const float lo = fma_rd(u2f(prmt(qlo_a[a], qlo_b[a], slot) & 0xffu), S, base[a]);
            const float hi = fma_ru(u2f(prmt(qhi_a[a], qhi_b[a], slot) & 0xffu), S, base[a]);
            const bool neg = (r.sign >> a) & 1u;
            N = max_nn(N, xmul(xsub(neg ? hi : lo, o[a]), inv[a]));
            F = min_nn(F, xmul(xsub(neg ? lo : hi, o[a]), inv[a]));
        }
        return F >= N;
    }
    float nn[3], ff[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const uint32_t e = (w0.w >> (8 * a)) & 0xffu;
        const float K = xmul(bits2f((e + 15u) << 23), inv[a]);
        const float A = xmul(xsub(base[a], o[a]), inv[a]);
        const float B = xsub(A, K);
        const float slack = fma_rn(xadd(fabsf(A), fabsf(K)), 9.5367431640625e-7f, 1.0e-35f);
        const bool neg = (r.sign >> a) & 1u;
        const uint32_t qn = prmt(neg ? qhi_a[a] : qlo_a[a], neg ? qhi_b[a] : qlo_b[a], slot) & 0xffu;
        const uint32_t qf = prmt(neg ? qlo_a[a] : qhi_a[a], neg ? qlo_b[a] : qhi_b[a], slot) & 0xffu;
        nn[a] = fma_rn(bits2f(0x3F800000u | (qn << 8)), K, xsub(B, slack));
        ff[a] = fma_rn(bits2f(0x3F800000u | (qf << 8)), K, xadd(B, slack));
    }
    const float N = max_nn(max_nn(nn[0], nn[1]), nn[2]);
    const float Fr = min_nn(min_nn(ff[0], ff[1]), ff[2]);
    const uint32_t bad = f2bits(xsub(Fr, N)) | f2bits(Fr) | f2bits(xsub(best, N));
    return !(bad >> 31);
}

// Lays out a solo lane's pending work as a frontier, back to front: the short stack (oldest first), its register-held
// top, the current node (or leaf waiting for a queue slot), then the queued leaves -- the reverse of the order in which
// the solo walk would get to them. Returns the number of entries.
B2_HD uint32_t coop_dump_fields(int sp, uint32_t top, uint32_t cur, uint32_t leaf1, uint32_t leaf0, const uint32_t* stack, uint32_t* F) {
    uint32_t k = 0;
    for (int i = 0; i < sp; ++i) F[k++] = stack[i];
    if (top != REF_EMPTY) F[k++] = top;
    if (cur != REF_EMPTY) F[k++] = cur;
    if (leaf1 != REF_EMPTY) F[k++] = leaf1;
    if (leaf0 != REF_EMPTY) F[k++] = leaf0;
    return k;
}
template <class LaneT>
B2_HD uint32_t coop_dump(const LaneT& L, const uint32_t* stack, uint32_t* F) {
    static_assert(B2_LEAF_QUEUE == 2, "the hand-over lays out a two-leaf queue");
    return coop_dump_fields(L.sp, L.top, L.cur, L.leaf1, L.leaf0, stack, F);
}

// Finishes the ray whose frontier F[0..n) (front = F[n-1]) and state (r, h) every lane holds identically.
// `fcap` = capacity of F; while n <= wide_limit up to B2_COOP_NODES nodes are expanded per round, beyond it one
// (depth-first: the frontier then grows by at most the solo walk's stack bound). On return h is the final hit on
// every lane; `overflow` is set if F would have overflowed (the ray is then abandoned: the caller reports it).
//
// One round = one memory round trip: the (up to 16) leaves at the front of the frontier and the first Q interior
// entries of the 32-entry window are fetched together; the leaves are resolved and committed in order, then the nodes
// are tested against the `best` just updated, and the window is rewritten in place (leaves consumed, nodes replaced by
// their passing children). Children just found are prefetched towards L2 for the next round.
template <bool ANY, bool COUNT>
B2_HD void coop_trace(const U4* wide, const U4* leaf, uint32_t* F, uint32_t n, uint32_t fcap, uint32_t wide_limit,
                      const RayX& r, HitX& h, TravCounters& tc, bool& overflow) {
    const uint32_t lane = w_lane();
    while (n) {
        if (COUNT) tc.rounds++;
        const uint32_t win = n < 32u ? n : 32u;
        const uint32_t e = lane < win ? F[n - 1u - lane] : REF_EMPTY;
        const bool is_node = lane < win && !(e & REF_LEAF_BIT);
        const uint32_t mnode = w_ballot(is_node);
        const uint32_t lead = mnode ? ctz32(mnode) : win;

        // ---- fetch: leaves at the front, two lanes each = the loader's two copies of a one-record block -------------------
        const uint32_t lg = lane >> 1, copy = lane & 1u;
        uint32_t nl = lead < 16u ? lead : 16u;
        const uint32_t lref = w_shfl(e, lg);
        U4 va = { 0, 0, 0, 0 }, vb = va, vc = va;
        if (lg < nl) {
            const U4* p = leaf + (lref & ~REF_LEAF_BIT);
            va = ld128(p); vb = ld128(p + 1); vc = ld128(p + 2);
        }
        // ---- fetch: the first Q interior entries of the window, COOP_LPN lanes each, every lane COOP_CPL children -------------
        uint32_t Q = n <= wide_limit ? (uint32_t)B2_COOP_NODES : 1u;
        const uint32_t have = popc32(mnode);
        if (Q > have) Q = have;
        // window positions of those Q entries: mexp has their bits, gpos is the one this lane's group expands
        const uint32_t grp = lane / COOP_LPN, jj = lane % COOP_LPN;
        uint32_t mexp = 0, gpos = 0;
        {
            uint32_t mm = mnode;
#pragma unroll
            for (uint32_t g = 0; g < (uint32_t)B2_COOP_NODES; ++g) {
                if (g < Q) { const uint32_t pos = ctz32(mm); mm &= mm - 1u; mexp |= 1u << pos; if (g == grp) gpos = pos; }
            }
        }
        const bool gact = grp < Q;
        const uint32_t node = w_shfl(e, gact ? gpos : 0u);
        U4 w0 = { 0, 0, 0, 0 }, w1 = w0, w2 = w0, w3 = w0, w4 = w0;
        uint32_t order = 0;
        if (gact) {
            const U4* p = wide + (uint32_t)WIDE_NODE_WORDS * node;
            w0 = ld128(p); w1 = ld128(p + 1); w2 = ld128(p + 2); w3 = ld128(p + 3); w4 = ld128(p + 4);
            order = ld32(reinterpret_cast<const uint32_t*>(p + 5) + (r.sign & 7u));
        }

        // ---- leaves: blocks with several records or an explicit box leave the fast path ----------------------------------
        if (nl) {
            const uint32_t big = w_ballot(lg < nl && ((va.w & LEAF_HAS_BOX) || ((va.w >> LEAF_NREC_SHIFT) & LEAF_NREC_MASK) != 1u));   // both lanes of a pair agree
            if (big & 1u) {
                // the first leaf is such a block: every lane walks it like the solo path (same result on every lane)
                const bool got = visit_leaf<COUNT>(leaf, w_shfl(e, 0) & ~REF_LEAF_BIT, r, h, &tc);
                n -= 1u;
                if ((ANY && got) || h.t < 0.0f) { n = 0; break; }
                continue;
            }
            if (big) nl = ctz32(big) >> 1;                                   // stop before the first such block
        }
        bool finished = false;
        if (nl) {
            const bool active = lg < nl;
            const uint32_t flags = va.w & REC_FLAG_MASK;
            const uint32_t id = vb.w + copy;
            const bool valid = active && (copy == 0u || flags != 0u);
            // the block's exact box: min / max of the record's vertices (b2rt_types.h)
            const float lox = min3_nn(bits2f(va.x), bits2f(vb.x), bits2f(vc.x)), hix = max3_nn(bits2f(va.x), bits2f(vb.x), bits2f(vc.x));
            const float loy = min3_nn(bits2f(va.y), bits2f(vb.y), bits2f(vc.y)), hiy = max3_nn(bits2f(va.y), bits2f(vb.y), bits2f(vc.y));
            const float loz = min3_nn(bits2f(va.z), bits2f(vb.z), bits2f(vc.z)), hiz = max3_nn(bits2f(va.z), bits2f(vb.z), bits2f(vc.z));
            float t = 0.0f, u = 0.0f, v = 0.0f;
            bool ok = false;
            if (valid) {
                // copy 1 is the loader's rotated duplicate: (v2,v3,v1) or (v3,v1,v2)
                const bool left = flags == REC_ROT_LEFT;
                const U4 p1 = copy ? (left ? vb : vc) : va, p2 = copy ? (left ? vc : va) : vb, p3 = copy ? (left ? va : vb) : vc;
                ok = tri_eval_exact(r, bits2f(p1.x), bits2f(p1.y), bits2f(p1.z), bits2f(p2.x), bits2f(p2.y), bits2f(p2.z),
                                    bits2f(p3.x), bits2f(p3.y), bits2f(p3.z), t, u, v);
            }
            if (COUNT) {
                const bool g = active && box_gate_exact(r, lox, loy, loz, hix, hiy, hiz, h.t);
                tc.leaf_blocks += nl;
                tc.leaf_pass += popc32(w_ballot(g) & 0x55555555u);
                tc.tri_tests += popc32(w_ballot(valid));                     // evaluations actually made (some behind a gate that fails later)
                tc.words += LEAF_RECORD_WORDS * nl;
            }
            // Commit in frontier order. Only a leaf holding a candidate under the current `best` can change it, and
            // `best` only shrinks: jump from one such leaf to the next, re-evaluating gates and candidates in between.
            uint32_t lo = 0;
            for (;;) {
                const bool gate = active && box_gate_exact(r, lox, loy, loz, hix, hiy, hiz, h.t);
                const bool cand = valid && ok && t < h.t && gate && lg >= lo;
                const uint32_t cm = w_ballot(cand);
                if (!cm) break;
                const uint32_t first = ctz32(cm) >> 1;                        // the first leaf (frontier order) with a candidate
                const bool mine = cand && lg == first;
                const uint32_t key = mine ? order_key(t) : 0xffffffffu;
                const uint32_t kmin = w_redmin(key);
                const uint32_t wl = ctz32(w_ballot(mine && key == kmin));    // lowest lane = lowest triangle id of the minimum
                h.t = bits2f(w_shfl(f2bits(t), wl)); h.u = bits2f(w_shfl(f2bits(u), wl)); h.v = bits2f(w_shfl(f2bits(v), wl));
                h.tri = w_shfl(id, wl);
                lo = first + 1u;
                if (ANY || h.t < 0.0f) { finished = true; break; }
            }
            if (finished || h.t < 0.0f) { n = 0; break; }
        }

        // ---- nodes, against the `best` the leaves of this round left ---------------------------------------------------------
        bool pass[COOP_CPL];
        uint32_t ref[COOP_CPL], pm[COOP_CPL];
#pragma unroll
        for (uint32_t t = 0; t < (uint32_t)COOP_CPL; ++t) {
            pass[t] = false; ref[t] = REF_EMPTY;
            const uint32_t j = jj + COOP_LPN * t;                              // child visited j-th (reference order)
            if (gact) {
                const uint32_t slot = (order >> (4u * j)) & 7u;
                if (j < (w0.w >> 24) && coop_child_test(w0, w2, w3, w4, slot, r, h.t)) {
                    pass[t] = true;
                    const uint32_t mb = prmt(w1.z, w1.w, slot) & 0xffu;
                    ref[t] = mb + ((mb & META_INTERIOR) ? w1.x - (uint32_t)META_INTERIOR : (REF_LEAF_BIT | w1.y));
                    // towards L2 for the next round: the child's node record or leaf block (either may straddle two lines)
                    const U4* c = (ref[t] & REF_LEAF_BIT) ? leaf + (ref[t] & ~REF_LEAF_BIT) : wide + (uint32_t)WIDE_NODE_WORDS * ref[t];
                    prefetch_l2(c); prefetch_l2(c + 6);
                }
            }
            pm[t] = w_ballot(pass[t]);
        }
        // Window entries are rewritten in place: the nl leaves at the front are consumed, every expanded node is replaced by
        // its passing children (in visiting order), everything else keeps its place. What window position `lane` puts out:
        const uint32_t gmask = (1u << COOP_LPN) - 1u;
        uint32_t mine = 0;                                                      // entries this window position contributes
        if (lane >= nl && lane < win) {
            if ((mexp >> lane) & 1u) {
                const uint32_t g = popc32(mexp & lanes_below(lane));           // which group expanded the node at this position
#pragma unroll
                for (uint32_t t = 0; t < (uint32_t)COOP_CPL; ++t) mine += popc32((pm[t] >> (COOP_LPN * g)) & gmask);
            } else mine = 1u;
        }
        // exclusive prefix sum of `mine` over the window (counts are at most 8: four ballots)
        uint32_t before_me = 0, total = 0;
#pragma unroll
        for (uint32_t bit = 0; bit < 4u; ++bit) {
            const uint32_t bm = w_ballot((mine >> bit) & 1u);
            before_me += popc32(bm & lanes_below(lane)) << bit;
            total += popc32(bm) << bit;
        }
        const uint32_t newn = n - win + total;
        if (COUNT) { tc.wide_nodes += Q; tc.words += WIDE_NODE_WORDS * Q; if (newn > tc.max_stack) tc.max_stack = newn; }
        if (newn > fcap) { overflow = true; n = 0; break; }
        const uint32_t K = Q ? top_bit(mexp) + 1u : nl;   // beyond the last expanded entry nothing moves
        const uint32_t node_first = w_shfl(before_me, gact ? gpos : 0u);      // where this group's children start
        w_sync();                                                    // every lane holds its window entry
        if (lane >= nl && lane < K && !((mexp >> lane) & 1u)) F[newn - 1u - before_me] = e;
        uint32_t earlier = 0;                                                   // passing children of this node before child j
#pragma unroll
        for (uint32_t t = 0; t < (uint32_t)COOP_CPL; ++t) {
            const uint32_t gb = (pm[t] >> (COOP_LPN * grp)) & gmask;
            if (pass[t]) F[newn - 1u - (node_first + earlier + popc32(gb & lanes_below(jj)))] = ref[t];
            earlier += popc32(gb);
        }
        w_sync();
        n = newn;
    }
}

}  // namespace b2rt
