// b2rt_types.h -- GPU-resident data layout shared by the host-side wide-BVH
// builder (wide_bvh.cpp) and the sm_100a kernels (kernels.cu).
//
// Inputs keep the reference's byte layouts (CLshared_structs.hpp:13-88):
//   CLTriangle 256 B, CLLinearBVHNode 48 B, CLMaterial 64 B.
// Derived, traversal-only layout built at upload time:
//   WideNode  112 B  up-to-8-wide node = a treelet of the host's binary BVH; child
//                    boxes quantised to 8 bits per plane, conservatively; the
//                    reference's visiting order stored per ray-sign octant.
//   leaf block 32 B header (the leaf's exact fp32 box, first triangle id, counts)
//                    + 48 B per stored triangle record (positions only).
//   ShadeTri   48 B  the three vertex normals + material index of a triangle.
#pragma once
#include <stdint.h>

namespace b2rt {

// ---- reference layouts (read-only views) ------------------------------------------
struct RefVec { float x, y, z, w; };
struct RefVertex { RefVec position, uv, normal, tangent_s, tangent_t; };          // 80 B
struct RefTriangle { RefVertex v1, v2, v3; uint32_t mtlIndex; uint32_t padding[3]; }; // 256 B
struct RefNode { RefVec bmin, bmax; uint32_t offset; uint16_t nPrimitives; uint8_t axis; uint8_t pad[9]; }; // 48 B
struct RefMaterial { RefVec diffuse, specular, emission; uint32_t type; float roughness, ior; int32_t padding; }; // 64 B
static_assert(sizeof(RefTriangle) == 256, "CLTriangle layout (CLshared_structs.hpp:44-74)");
static_assert(sizeof(RefNode) == 48, "CLLinearBVHNode layout (CLshared_structs.hpp:76-88)");
static_assert(sizeof(RefMaterial) == 64, "CLMaterial layout (CLshared_structs.hpp:13-26)");

// ---- compressed wide BVH -----------------------------------------------------------
// A wide node is a treelet of the host's binary BVH: starting from a binary node, the
// interior child with the largest surface area is opened repeatedly until 8 children
// are reached (or only leaves remain). Child slot c (0..n-1) is the child's position in
// the treelet's depth-first order with the FIRST binary child (`index+1`) before the
// second (`offset`) -- the order the reference walks for a ray whose direction is
// positive on every split axis. For the other seven sign octants the reference swaps
// the two sub-ranges under every treelet node whose split axis has a negative direction
// (kernel_bvh.cl:200-207); the resulting visiting order is precomputed per octant as
// eight 4-bit slot numbers (`order[octant]`, nibble k = slot visited k-th), so the
// kernels can replay the reference's leaf order exactly for any treelet shape.
//
//   word0 = { base.x, base.y, base.z, exp_x | exp_y<<8 | exp_z<<16 | n_children<<24 }
//   word1 = { child_base, leaf_base, meta[0..3], meta[4..7] }
//   word2 = { qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7] }
//   word3 = { qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7] }
//   word4 = { qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7] }
//   word5 = { order[0..3] }   word6 = { order[4..7] }
// plane(axis, q) = base[axis] + q * 2^(exp[axis]-127); child boxes are quantised outward
// (qlo rounded down, qhi rounded up). meta[c]: bit 7 set = interior child, wide node
// index child_base + (meta & 0x7f); clear = leaf block at leaf_base + meta (16-byte
// units). Ranks >= n_children of an order word name an empty slot.
struct alignas(16) WideNode {
    float base[3];          //  0
    uint8_t exp[3];         // 12
    uint8_t n_children;     // 15
    uint32_t child_base;    // 16
    uint32_t leaf_base;     // 20
    uint8_t meta[8];        // 24
    uint8_t qlo[3][8];      // 32  qlo[axis][slot]
    uint8_t qhi[3][8];      // 56
    uint32_t order[8];      // 80  order[sign octant: sx | sy<<1 | sz<<2]
};
static_assert(sizeof(WideNode) == 112, "WideNode must be 112 B (seven 16-byte words)");
enum { WIDE_NODE_WORDS = 7, META_INTERIOR = 0x80, META_MAX_LEAF_OFFSET = 0x7f };

// Leaf block: n_records * 3 sixteen-byte words, optionally followed by two words holding the leaf's box.
//   record = { v1.xyz, w0 } { v2.xyz, w1 } { v3.xyz, 0 }
//   w0 = flags (bits 0-1); in the block's FIRST record additionally n_records << 8 and LEAF_HAS_BOX (bit 31)
//   w1 = (first record only) index of the block's first reference triangle
//   [ { bmin.xyz, 0 } { bmax.xyz, 0 } ]   only when LEAF_HAS_BOX is set
// flags: 0 = one reference triangle (v1,v2,v3);
//        1 = two reference triangles: (v1,v2,v3) then its rotation (v2,v3,v1);
//        2 = two reference triangles: (v1,v2,v3) then its rotation (v3,v1,v2).
// The reference's loader emits every OBJ triangle twice, the second copy
// rotated (CLOBJloader.cpp:102-126), and both copies always share a leaf, so a
// record usually stands for two consecutive reference triangles.
// The reference's builder gives a leaf the union of its triangles' bounds (CLBVHnode.cpp:18-23), i.e. the exact
// component-wise min / max of the vertex positions. A one-record block whose node box equals that min / max (checked when
// the block is written) therefore carries no box at all: the kernels recompute it from the three vertices they load
// anyway (min / max are exact), 48 bytes per leaf instead of 80. Any other block (several records, or a caller-supplied
// tree whose leaf box is something else) keeps its box behind the records and is fetched with one more round trip.
enum { LEAF_RECORD_WORDS = 3, LEAF_BOX_WORDS = 2 };
enum : uint32_t { LEAF_HAS_BOX = 0x80000000u, LEAF_NREC_SHIFT = 8, LEAF_NREC_MASK = 0xffffu, REC_FLAG_MASK = 3u };
enum : uint32_t { REC_SINGLE = 0, REC_ROT_LEFT = 1, REC_ROT_RIGHT = 2 };

struct ShadeTri { float n1[3]; uint32_t mtl; float n2[3]; uint32_t pad0; float n3[3]; uint32_t pad1; }; // 48 B
static_assert(sizeof(ShadeTri) == 48, "ShadeTri layout");

// Stack entries / child references: bit 31 set = leaf block (low 31 bits = offset in
// 16-byte units), clear = wide node index. 0xFFFFFFFF is never produced and means "empty".
enum : uint32_t { REF_LEAF_BIT = 0x80000000u, REF_EMPTY = 0xFFFFFFFFu };

struct alignas(16) U4 { uint32_t x, y, z, w; };   // one 128-bit load

struct SceneView {                 // plain device pointers handed to kernels
    const U4* wide;                // WideNode[] viewed as 7 x U4 each; node 0 is the root
    const U4* leaf;                // leaf blocks
    const ShadeTri* shade;         // per reference triangle
    const RefMaterial* mats;
    const RefTriangle* tris;       // reference-layout copies (binary-walk kernels only)
    const RefNode* nodes;
    uint32_t n_tris, n_nodes, n_mats, n_wide;
    uint32_t one_bits;             // 0x3F800000 as run-time data (see Lane::node_step)
};

// Which pixels (global ids, gid = y*W + x, kernel_bvh.cl:394-395) a frame launch draws: work item i of
// the launch is gid = begin + (i / band) * stride + i % band. A contiguous range has stride == band; a
// multi-GPU rank's share of a frame is every world-th band of `band` pixels.
struct GidMap {
    uint64_t begin;
    uint32_t band, stride;
#if defined(__CUDACC__)
    __host__ __device__
#endif
    uint64_t gid(uint32_t i) const { return begin + (uint64_t)(i / band) * stride + i % band; }
};

// Scalar arguments of KernelEntry (kernel_bvh.cl:421-430) as the kernels receive them.
struct FrameArgs {
    uint32_t width, height, frame_count;
    int32_t bounces, light_type;
    float sky;
    float pos[3], front[3], up[3];
    float angle;            // tanf(0.5f * (45.0f * 3.1415f / 180.0f)) evaluated by the host libm (kernel_bvh.cl:392)
    float* mirror;          // multi-GPU: finished pixels are ALSO stored into this image (the root GPU's, peer-mapped over NVLink); null = none
};

}  // namespace b2rt
