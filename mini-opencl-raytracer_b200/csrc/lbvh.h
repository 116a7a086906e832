// lbvh.h -- device-side binary BVH build (lbvh.cu): Morton sort + Karras hierarchy + bottom-up refit over primitive
// groups. api.cu turns the result into the reference's CLLinearBVHNode array and triangle order.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b2rt {

// Scratch bytes lbvh_build needs for m groups.
size_t lbvh_scratch_bytes(uint32_t m);

// d_group_bounds: m x {min.xyz, max.xyz}. Outputs (device): children of the m-1 internal nodes (>= 0: internal node,
// < 0: ~leaf, leaf = position in Morton order), their boxes (6 floats each) and split axes, and the group id of every
// leaf. Internal node 0 is the root. m >= 2.
cudaError_t lbvh_build(const float* d_group_bounds, uint32_t m, const float scene_lo[3], const float scene_hi[3], void* d_scratch,
                       int2* d_children, float* d_node_bounds, uint8_t* d_axis, uint32_t* d_sorted_group, uint64_t* launches,
                       cudaStream_t st);

}  // namespace b2rt
