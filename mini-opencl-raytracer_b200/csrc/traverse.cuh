// traverse.cuh -- per-ray traversal primitives for the compressed wide BVH.
//
// Written once for the sm_100a kernels (kernels.cu). The same text also
// compiles as plain C++ (tests/emu/, TEST ONLY) so that the builder + traversal
// logic can be checked against the CPU oracle in a container with no GPU; the
// product library never contains or calls that host build.
//
// Exactness contract:
//  * ray set-up, the leaf box gate and the Moller-Trumbore test replay the
//    reference's fp32 operation order with one IEEE rounding per operation
//    (no FMA contraction): InitRay kernel_bvh.cl:42-55, RayBounds :156-169,
//    RayTriangle :98-153. On the device every such op is an explicit
//    __f*_rn intrinsic, which nvcc never contracts.
//  * leaves are entered in exactly the reference's DFS order (near child first
//    by sign[axis], kernel_bvh.cl:200-207) and a leaf's triangles are tested iff
//    the leaf's exact fp32 box passes RayBounds on [0, best]. Because a parent
//    box contains its children and fp32 subtraction/multiplication by a fixed
//    operand are monotone, that is equivalent to the reference's rule "every
//    ancestor passed when it was visited", so interior culling may be anything
//    conservative: here 8-bit quantised boxes evaluated with directed rounding.
#pragma once
#include "b2rt_types.h"

#if defined(__CUDACC__)
#define B2_HD __device__ __forceinline__
#else
#define B2_HD inline
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace b2rt {

// ---- arithmetic primitives ----------------------------------------------------------
#if defined(__CUDACC__)
B2_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
B2_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
B2_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
B2_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
B2_HD float xsqrt(float a) { return __fsqrt_rn(a); }
B2_HD float sub_rd(float a, float b) { return __fsub_rd(a, b); }
B2_HD float sub_ru(float a, float b) { return __fsub_ru(a, b); }
B2_HD float fma_rd(float a, float b, float c) { return __fmaf_rd(a, b, c); }
B2_HD float fma_ru(float a, float b, float c) { return __fmaf_ru(a, b, c); }
B2_HD float mul_rd(float a, float b) { return __fmul_rd(a, b); }
B2_HD float mul_ru(float a, float b) { return __fmul_ru(a, b); }
B2_HD float max_nn(float a, float b) { return fmaxf(a, b); }   // NaN-ignoring
B2_HD float min_nn(float a, float b) { return fminf(a, b); }
B2_HD float u2f(uint32_t v) { return __uint2float_rn(v); }
B2_HD uint32_t byte_of(uint32_t w, uint32_t i) { return __byte_perm(w, 0, 0x4440u + i); }
B2_HD uint32_t popc32(uint32_t v) { return __popc(v); }
B2_HD float bits2f(uint32_t v) { return __uint_as_float(v); }
B2_HD uint32_t f2bits(float v) { return __float_as_uint(v); }
B2_HD U4 ld128(const U4* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    U4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
#else
// Host emulation (tests only). Built with -ffp-contract=off -frounding-math.
B2_HD float xadd(float a, float b) { volatile float r = a + b; return r; }
B2_HD float xsub(float a, float b) { volatile float r = a - b; return r; }
B2_HD float xmul(float a, float b) { volatile float r = a * b; return r; }
B2_HD float xdiv(float a, float b) { volatile float r = a / b; return r; }
B2_HD float xsqrt(float a) { return std::sqrt(a); }
template <class F> B2_HD float with_round(int mode, F f) {
    int old = std::fegetround(); std::fesetround(mode); volatile float r = f(); std::fesetround(old); return r;
}
B2_HD float sub_rd(float a, float b) { return with_round(FE_DOWNWARD, [&] { volatile float x = a, y = b; return x - y; }); }
B2_HD float sub_ru(float a, float b) { return with_round(FE_UPWARD, [&] { volatile float x = a, y = b; return x - y; }); }
B2_HD float fma_rd(float a, float b, float c) { return with_round(FE_DOWNWARD, [&] { volatile float x = a, y = b, z = c; return std::fmaf(x, y, z); }); }
B2_HD float fma_ru(float a, float b, float c) { return with_round(FE_UPWARD, [&] { volatile float x = a, y = b, z = c; return std::fmaf(x, y, z); }); }
B2_HD float mul_rd(float a, float b) { return with_round(FE_DOWNWARD, [&] { volatile float x = a, y = b; return x * y; }); }
B2_HD float mul_ru(float a, float b) { return with_round(FE_UPWARD, [&] { volatile float x = a, y = b; return x * y; }); }
B2_HD float max_nn(float a, float b) { return std::fmax(a, b); }
B2_HD float min_nn(float a, float b) { return std::fmin(a, b); }
B2_HD float u2f(uint32_t v) { return (float)v; }
B2_HD uint32_t byte_of(uint32_t w, uint32_t i) { return (w >> (8 * i)) & 0xffu; }
B2_HD uint32_t popc32(uint32_t v) { return (uint32_t)__builtin_popcount(v); }
B2_HD float bits2f(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
B2_HD uint32_t f2bits(float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; }
B2_HD U4 ld128(const U4* p) { return *p; }
#endif

// OpenCL max()/min() as worded by the spec ("y if x < y, otherwise x"), the
// convention the oracle pins for NaN operands (oracle/ref_build/cl_shim.hpp).
B2_HD float max_cl(float x, float y) { return x < y ? y : x; }
B2_HD float min_cl(float x, float y) { return y < x ? y : x; }

// ---- ray ----------------------------------------------------------------------------
struct RayX {
    float ox, oy, oz;
    float dx, dy, dz;      // normalised exactly like InitRay
    float ix, iy, iz;      // 1/d, may be +-inf
    uint32_t sign;         // bit a: inv[a] < 0
};

// InitRay, kernel_bvh.cl:42-55, with normalize(v) = v / sqrt(dot(v,v)).
B2_HD RayX make_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    RayX r;
    float len = xsqrt(xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz)));
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.dx = xdiv(dx, len); r.dy = xdiv(dy, len); r.dz = xdiv(dz, len);
    r.ix = xdiv(1.0f, r.dx); r.iy = xdiv(1.0f, r.dy); r.iz = xdiv(1.0f, r.dz);
    r.sign = (r.ix < 0 ? 1u : 0u) | (r.iy < 0 ? 2u : 0u) | (r.iz < 0 ? 4u : 0u);
    return r;
}

struct HitX {
    float t, u, v;
    uint32_t tri;          // B2RT_MISS (0xFFFFFFFF) until a triangle is accepted
};

struct TravCounters { uint32_t wide_nodes, leaf_blocks, leaf_pass, tri_tests, words; };

// RayBounds, kernel_bvh.cl:156-169, on an exact fp32 box.
B2_HD bool box_gate_exact(const RayX& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float best) {
    float nx = (r.sign & 1u) ? hix : lox, fx = (r.sign & 1u) ? lox : hix;
    float ny = (r.sign & 2u) ? hiy : loy, fy = (r.sign & 2u) ? loy : hiy;
    float nz = (r.sign & 4u) ? hiz : loz, fz = (r.sign & 4u) ? loz : hiz;
    float t0 = max_cl(0.0f, xmul(xsub(nx, r.ox), r.ix));
    float t1 = min_cl(best, xmul(xsub(fx, r.ox), r.ix));
    t0 = max_cl(t0, xmul(xsub(ny, r.oy), r.iy));
    t1 = min_cl(t1, xmul(xsub(fy, r.oy), r.iy));
    t0 = max_cl(t0, xmul(xsub(nz, r.oz), r.iz));
    t1 = min_cl(t1, xmul(xsub(fz, r.oz), r.iz));
    return t1 >= t0;
}

// RayTriangle, kernel_bvh.cl:98-153, for triangle (a,b,c) = (v1,v2,v3) with id `tri`.
// Accepts iff det >= 1e-8, 0<=u<=1, v>=0, u+v<=1 and t < best (no lower bound on t).
B2_HD void tri_test_exact(const RayX& r, float ax, float ay, float az, float bx, float by, float bz,
                          float cx, float cy, float cz, uint32_t tri, HitX& h) {
    float e1x = xsub(bx, ax), e1y = xsub(by, ay), e1z = xsub(bz, az);
    float e2x = xsub(cx, ax), e2y = xsub(cy, ay), e2z = xsub(cz, az);
    float px = xsub(xmul(r.dy, e2z), xmul(r.dz, e2y));
    float py = xsub(xmul(r.dz, e2x), xmul(r.dx, e2z));
    float pz = xsub(xmul(r.dx, e2y), xmul(r.dy, e2x));
    float det = xadd(xadd(xmul(e1x, px), xmul(e1y, py)), xmul(e1z, pz));
    if (det < 1.0e-8f || -det > 1.0e-8f) return;
    float inv = xdiv(1.0f, det);
    float tx = xsub(r.ox, ax), ty = xsub(r.oy, ay), tz = xsub(r.oz, az);
    float u = xmul(xadd(xadd(xmul(tx, px), xmul(ty, py)), xmul(tz, pz)), inv);
    if (u < 0.0f || u > 1.0f) return;
    float qx = xsub(xmul(ty, e1z), xmul(tz, e1y));
    float qy = xsub(xmul(tz, e1x), xmul(tx, e1z));
    float qz = xsub(xmul(tx, e1y), xmul(ty, e1x));
    float v = xmul(xadd(xadd(xmul(r.dx, qx), xmul(r.dy, qy)), xmul(r.dz, qz)), inv);
    if (v < 0.0f || xadd(u, v) > 1.0f) return;
    float t = xmul(xadd(xadd(xmul(e2x, qx), xmul(e2y, qy)), xmul(e2z, qz)), inv);
    if (t < h.t) { h.t = t; h.u = u; h.v = v; h.tri = tri; }
}

// ---- leaf block ----------------------------------------------------------------------
// Returns true when at least one triangle was accepted by this call.
template <bool COUNT>
B2_HD bool visit_leaf(const U4* leaf, uint32_t offset, const RayX& r, HitX& h, TravCounters* c) {
    const U4* p = leaf + offset;
    U4 h0 = ld128(p), h1 = ld128(p + 1);
    U4 a = ld128(p + 2), b = ld128(p + 3), cc = ld128(p + 4);   // every block has >= 1 record
    uint32_t nrec = h1.w;
    if (COUNT) { c->leaf_blocks++; c->words += LEAF_HEADER_WORDS + LEAF_RECORD_WORDS * nrec; }
    if (!box_gate_exact(r, bits2f(h0.x), bits2f(h0.y), bits2f(h0.z), bits2f(h1.x), bits2f(h1.y), bits2f(h1.z), h.t))
        return false;
    if (COUNT) c->leaf_pass++;
    uint32_t before = h.tri;
    float t_before = h.t;
    uint32_t id = h0.w;
    for (uint32_t k = 0;;) {
        float v1x = bits2f(a.x), v1y = bits2f(a.y), v1z = bits2f(a.z);
        float v2x = bits2f(b.x), v2y = bits2f(b.y), v2z = bits2f(b.z);
        float v3x = bits2f(cc.x), v3y = bits2f(cc.y), v3z = bits2f(cc.z);
        uint32_t flags = a.w;
        tri_test_exact(r, v1x, v1y, v1z, v2x, v2y, v2z, v3x, v3y, v3z, id, h);
        if (COUNT) c->tri_tests++;
        if (flags == REC_ROT_LEFT) {
            tri_test_exact(r, v2x, v2y, v2z, v3x, v3y, v3z, v1x, v1y, v1z, id + 1, h);
            if (COUNT) c->tri_tests++;
        } else if (flags == REC_ROT_RIGHT) {
            tri_test_exact(r, v3x, v3y, v3z, v1x, v1y, v1z, v2x, v2y, v2z, id + 1, h);
            if (COUNT) c->tri_tests++;
        }
        id += flags ? 2u : 1u;
        if (++k >= nrec) break;
        const U4* q = p + LEAF_HEADER_WORDS + LEAF_RECORD_WORDS * k;
        a = ld128(q); b = ld128(q + 1); cc = ld128(q + 2);
    }
    return h.tri != before || h.t != t_before;
}

// ---- wide node ------------------------------------------------------------------------
struct WideHits {
    uint32_t mask;         // bit c: child slot c may intersect [0,best] (conservative)
    uint32_t flips;        // 7 bits: treelet node j is visited second-child-first for this ray
    uint32_t imask, child_base, leaf_base, meta_lo, meta_hi;
};

// Child reference of slot c (interior: wide node index, leaf: REF_LEAF_BIT | block offset).
B2_HD uint32_t child_ref(const WideHits& w, uint32_t c) {
    if ((w.imask >> c) & 1u) return w.child_base + popc32(w.imask & ((1u << c) - 1u));
    uint32_t m = byte_of(c < 4 ? w.meta_lo : w.meta_hi, c & 3u);
    return REF_LEAF_BIT | (w.leaf_base + m);
}

// Slot visited k-th (k = 0 first) when the treelet is walked in the reference's order.
B2_HD uint32_t slot_of_rank(uint32_t flips, uint32_t k) {
    uint32_t c2 = ((k >> 2) & 1u) ^ (flips & 1u);
    uint32_t c1 = ((k >> 1) & 1u) ^ ((flips >> (1u + c2)) & 1u);
    uint32_t c0 = (k & 1u) ^ ((flips >> (3u + 2u * c2 + c1)) & 1u);
    return (c2 << 2) | (c1 << 1) | c0;
}

// Conservative slab test of the 8 quantised child boxes against [0, best].
// For every axis the ray is mirrored so that it travels in +direction:
//   near plane lower bound  n = RD( RD(q_near * S + A_rd) * |inv| )
//   far  plane upper bound  f = RU( RU(q_far  * S + A_ru) * |inv| )
// with S = +-2^e and A = +-(base - o) rounded toward the safe side, so that
// n <= fl((plane_near - o) * inv) and f >= fl((plane_far - o) * inv) of the exact
// test for any exact plane inside the quantised one (monotonicity of fl()).
B2_HD WideHits test_wide_node(const U4* wide, uint32_t index, const RayX& r, float best) {
    const U4* p = wide + 6u * index;
    U4 w0 = ld128(p), w1 = ld128(p + 1), w2 = ld128(p + 2), w3 = ld128(p + 3), w4 = ld128(p + 4), w5 = ld128(p + 5);
    WideHits out;
    out.imask = w0.w >> 24;
    out.child_base = w1.x;
    out.leaf_base = w1.y;
    out.meta_lo = w2.x;
    out.meta_hi = w2.y;
    uint32_t axes = w1.z;
    uint32_t rm = ((r.sign & 1u) ? 0x7fu : 0u) | ((r.sign & 2u) ? 0x7f00u : 0u) | ((r.sign & 4u) ? 0x7f0000u : 0u);
    uint32_t t = axes & rm;
    out.flips = (t | (t >> 8) | (t >> 16)) & 0x7fu;

    float S[3], Ard[3], Aru[3], IA[3];
    uint32_t qn_lo[3], qn_hi[3], qf_lo[3], qf_hi[3];
    const float o[3] = { r.ox, r.oy, r.oz };
    const float inv[3] = { r.ix, r.iy, r.iz };
    const float base[3] = { bits2f(w0.x), bits2f(w0.y), bits2f(w0.z) };
    const uint32_t lo_lo[3] = { w2.z, w3.x, w3.z }, lo_hi[3] = { w2.w, w3.y, w3.w };
    const uint32_t hi_lo[3] = { w4.x, w4.z, w5.x }, hi_hi[3] = { w4.y, w4.w, w5.y };
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float s = bits2f(((w0.w >> (8 * a)) & 0xffu) << 23);
        float d_rd = sub_rd(base[a], o[a]), d_ru = sub_ru(base[a], o[a]);
        bool neg = (r.sign >> a) & 1u;
        S[a] = neg ? -s : s;
        Ard[a] = neg ? -d_ru : d_rd;
        Aru[a] = neg ? -d_rd : d_ru;
        IA[a] = fabsf(inv[a]);
        qn_lo[a] = neg ? hi_lo[a] : lo_lo[a]; qn_hi[a] = neg ? hi_hi[a] : lo_hi[a];
        qf_lo[a] = neg ? lo_lo[a] : hi_lo[a]; qf_hi[a] = neg ? lo_hi[a] : hi_hi[a];
    }
    uint32_t mask = 0;
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c) {
        float t0 = 0.0f, t1 = best;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float qn = u2f(byte_of(c < 4 ? qn_lo[a] : qn_hi[a], c & 3u));
            float qf = u2f(byte_of(c < 4 ? qf_lo[a] : qf_hi[a], c & 3u));
            float n = mul_rd(fma_rd(qn, S[a], Ard[a]), IA[a]);
            float f = mul_ru(fma_ru(qf, S[a], Aru[a]), IA[a]);
            t0 = max_nn(t0, n);
            t1 = min_nn(t1, f);
        }
        if (t1 >= t0) mask |= 1u << c;
    }
    out.mask = mask & (axes >> 24);
    return out;
}

// ---- simple per-ray driver (one thread = one ray; also the host emulation) -------------
// ANY: stop at the first accepted triangle (occlusion query). The persistent
// kernels in kernels.cu re-implement this loop warp-synchronously with the same
// primitives; this form is used by the render megakernel and by tests/emu.
template <bool ANY, bool COUNT, int STACK_CAP>
B2_HD HitX trace_wide(const U4* wide, const U4* leaf, const RayX& r, float tmax, TravCounters* c, bool* overflow) {
    HitX h; h.t = tmax; h.u = 0.0f; h.v = 0.0f; h.tri = 0xFFFFFFFFu;
    uint32_t stack[STACK_CAP];
    int sp = 0;
    uint32_t cur = 0;   // root wide node
    for (;;) {
        if (cur & REF_LEAF_BIT) {
            bool got = visit_leaf<COUNT>(leaf, cur & ~REF_LEAF_BIT, r, h, c);
            if ((ANY && got) || h.t < 0.0f) break;   // best < 0: every later box test fails (A-5)
        } else {
            WideHits w = test_wide_node(wide, cur, r, h.t);
            if (COUNT) { c->wide_nodes++; c->words += 6; }
            // push far -> near so that the nearest child in reference order pops first
            for (int k = 7; k >= 0; --k) {
                uint32_t s = slot_of_rank(w.flips, (uint32_t)k);
                if ((w.mask >> s) & 1u) {
                    if (sp == STACK_CAP) { if (overflow) *overflow = true; break; }
                    stack[sp++] = child_ref(w, s);
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return h;
}

}  // namespace b2rt
