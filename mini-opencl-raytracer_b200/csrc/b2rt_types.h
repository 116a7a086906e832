// b2rt_types.h -- GPU-resident data layout shared by the host-side wide-BVH
// builder (wide_bvh.cpp) and the sm_100a kernels (kernels.cu).
//
// Inputs keep the reference's byte layouts (CLshared_structs.hpp:13-88):
//   CLTriangle 256 B, CLLinearBVHNode 48 B, CLMaterial 64 B.
// Derived, traversal-only layout built at upload time:
//   WideNode   96 B  8-wide node = a binary node of the host BVH collapsed with
//                    its children and grandchildren (a depth<=3 treelet); child
//                    boxes quantised to 8 bits per plane, conservatively.
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
// Child slot c in 0..7 is the 3-bit path from the treelet root: bit2 = first
// split (0 = first child `index+1`, 1 = second child `offset`), bit1 = second
// split, bit0 = third. A binary leaf met above depth 3 keeps the remaining path
// bits zero. The treelet's (up to) 7 binary interior nodes are numbered heap
// style: 0 = root, 1+b2 = depth 1, 3+(b2<<1|b1) = depth 2; their split axes are
// stored as three 7-bit masks so that the reference's near-child-first order
// (kernel_bvh.cl:200-207) can be replayed exactly for any ray sign octant.
struct alignas(32) WideNode {
    float base[3];          //  0  quantisation origin (min corner of the union of the child boxes)
    uint8_t exp[3];         // 12  biased exponents: plane = base + q * 2^(exp-127)
    uint8_t imask;          // 15  bit c: slot c holds an interior child (another WideNode)
    uint32_t child_base;    // 16  index of the first interior child; child of slot c = child_base + popc(imask & ((1<<c)-1))
    uint32_t leaf_base;     // 20  offset (16-byte units) of this node's first leaf block in the leaf buffer
    uint32_t axes;          // 24  mx | my<<8 | mz<<16 | valid<<24 (bit j of m*: treelet node j splits on that axis)
    uint32_t reserved;      // 28
    uint8_t meta[8];        // 32  leaf child: block offset relative to leaf_base in 16-byte units
    uint8_t qlo[3][8];      // 40  quantised min planes: qlo[axis][slot]
    uint8_t qhi[3][8];      // 64  quantised max planes
    uint8_t spare[8];       // 88
};
static_assert(sizeof(WideNode) == 96, "WideNode must be 96 B (three 32-byte sectors)");

// Leaf block header, two 16-byte words followed by n_records * 3 words.
//   word0 = { bmin.x, bmin.y, bmin.z, first_tri }
//   word1 = { bmax.x, bmax.y, bmax.z, n_records }
//   record = { v1.xyz, flags } { v2.xyz, 0 } { v3.xyz, 0 }
// flags: 0 = one reference triangle (v1,v2,v3);
//        1 = two reference triangles: (v1,v2,v3) then its rotation (v2,v3,v1);
//        2 = two reference triangles: (v1,v2,v3) then its rotation (v3,v1,v2).
// The reference's loader emits every OBJ triangle twice, the second copy
// rotated (CLOBJloader.cpp:102-126), and both copies always share a leaf, so a
// record usually stands for two consecutive reference triangles.
enum { LEAF_HEADER_WORDS = 2, LEAF_RECORD_WORDS = 3 };
enum : uint32_t { REC_SINGLE = 0, REC_ROT_LEFT = 1, REC_ROT_RIGHT = 2 };

struct ShadeTri { float n1[3]; uint32_t mtl; float n2[3]; uint32_t pad0; float n3[3]; uint32_t pad1; }; // 48 B
static_assert(sizeof(ShadeTri) == 48, "ShadeTri layout");

// Stack entries / child references: bit 31 set = leaf block (low 31 bits = offset in
// 16-byte units), clear = wide node index. 0xFFFFFFFF is never produced and means "empty".
enum : uint32_t { REF_LEAF_BIT = 0x80000000u, REF_EMPTY = 0xFFFFFFFFu };

struct alignas(16) U4 { uint32_t x, y, z, w; };   // one 128-bit load

struct SceneView {                 // plain device pointers handed to kernels
    const U4* wide;                // WideNode[] viewed as 6 x U4 each; node 0 is the root
    const U4* leaf;                // leaf blocks
    const ShadeTri* shade;         // per reference triangle
    const RefMaterial* mats;
    const RefTriangle* tris;       // reference-layout copies (binary-walk kernels only)
    const RefNode* nodes;
    uint32_t n_tris, n_nodes, n_mats, n_wide;
};

// Scalar arguments of KernelEntry (kernel_bvh.cl:421-430) as the kernels receive them.
struct FrameArgs {
    uint32_t width, height, frame_count;
    int32_t bounces, light_type;
    float sky;
    float pos[3], front[3], up[3];
    float angle;            // tanf(0.5f * (45.0f * 3.1415f / 180.0f)) evaluated by the host libm (kernel_bvh.cl:392)
};

}  // namespace b2rt
