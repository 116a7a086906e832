// lbvh.cu -- binary BVH build on the device: a linear BVH (Morton order + Karras 2012) over
// primitive GROUPS, returned to the host in a form that api.cu flattens into the REFERENCE's own
// CLLinearBVHNode array + triangle order, i.e. a drop-in for what CLBVHScene::RecursiveBuild +
// FlattenBVHTree produce (CLBVHnode.cpp:7-183) -- a different tree, same format, same consumers.
//
// A group is a run of consecutive loader triangles with bit-identical centroids: the loader emits every OBJ face
// as 2 (triangle) or 3 (quad) overlapping triangles (CLOBJloader.cpp:102-126) and the reference's builder can never
// separate them (identical centroids), so they stay one leaf here too.
//
//   morton_kernel      30-bit Morton code of each group's centroid inside the scene box
//   split passes       stable radix sort, one bit per pass: block scan of the zero flags, scan of the block sums,
//                      scatter (three small kernels per bit; 90 launches for 30 bits -- the sort of 10^6 keys takes
//                      well under a millisecond of GPU time, launch-bound)
//   karras_kernel      one thread per internal node: range, split, children, parent links, split axis
//   refit_kernel       one thread per leaf walks up; the second arrival at a node unions its children's boxes
#include <cuda_runtime.h>
#include <stdint.h>
#include "lbvh.h"

namespace b2rt {

namespace {

constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ uint32_t expand10(uint32_t v) {      // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(const float* __restrict__ bounds, uint32_t m, float3 lo, float3 inv, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float* b = bounds + 6 * (size_t)i;
    float cx = (0.5f * b[0] + 0.5f * b[3] - lo.x) * inv.x, cy = (0.5f * b[1] + 0.5f * b[4] - lo.y) * inv.y,
          cz = (0.5f * b[2] + 0.5f * b[5] - lo.z) * inv.z;
    uint32_t x = (uint32_t)fminf(fmaxf(cx * 1024.0f, 0.0f), 1023.0f), y = (uint32_t)fminf(fmaxf(cy * 1024.0f, 0.0f), 1023.0f),
             z = (uint32_t)fminf(fmaxf(cz * 1024.0f, 0.0f), 1023.0f);
    keys[i] = (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);      // bit b of the code: axis 2 - b % 3
    vals[i] = i;
}

// Exclusive scan of zero flags (bit `bit` of keys clear) inside each block of SCAN_BLOCK elements; block totals out.
__global__ void __launch_bounds__(SCAN_BLOCK) flag_scan_kernel(const uint32_t* __restrict__ keys, uint32_t m, uint32_t bit,
                                                                uint32_t* __restrict__ excl, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_sums[SCAN_BLOCK / 32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t flag = (i < m && !((keys[i] >> bit) & 1u)) ? 1u : 0u;
    uint32_t v = flag;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v += t; }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += t; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t before = (warp ? warp_sums[warp - 1] : 0u) + v - flag;
    if (i < m) excl[i] = before;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = before + flag;
}

// In-place exclusive scan of up to 2^20 * ... block sums by ONE block (loops over the array); total zeros to sums[n].
__global__ void __launch_bounds__(SCAN_BLOCK) sums_scan_kernel(uint32_t* __restrict__ sums, uint32_t n) {
    __shared__ uint32_t warp_sums[SCAN_BLOCK / 32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < n ? sums[i] : 0u;
        uint32_t v = x;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v += t; }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += t; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = (warp ? warp_sums[warp - 1] : 0u) + v;
        if (i < n) sums[i] = carry + incl - x;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[n] = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) split_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t m,
                                                                    uint32_t bit, const uint32_t* __restrict__ excl,
                                                                    const uint32_t* __restrict__ block_offsets, uint32_t n_blocks,
                                                                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i >= m) return;
    const uint32_t zeros_before = block_offsets[blockIdx.x] + excl[i], total_zeros = block_offsets[n_blocks];
    const uint32_t k = keys[i];
    const uint32_t dst = ((k >> bit) & 1u) ? total_zeros + (i - zeros_before) : zeros_before;
    keys_out[dst] = k;
    vals_out[dst] = vals[i];
}

// Common-prefix length of sorted keys i and j (index tie-break for equal codes), -1 outside the array.
__device__ __forceinline__ int delta(const uint32_t* keys, int m, int i, int j) {
    if (j < 0 || j >= m) return -1;
    const uint32_t a = keys[i], b = keys[j];
    return a != b ? __clz(a ^ b) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

// Karras 2012, one thread per internal node i in [0, m-2]. Child reference: index of an internal node, or ~leaf.
__global__ void karras_kernel(const uint32_t* __restrict__ keys, int m, int2* __restrict__ children, int* __restrict__ parent_internal,
                              int* __restrict__ parent_leaf, uint8_t* __restrict__ axis) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1) return;
    const int d = delta(keys, m, i, i + 1) - delta(keys, m, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, m, i, i - d);
    int lmax = 2;
    while (delta(keys, m, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, m, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, m, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, m, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma, right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    if (left >= 0) parent_internal[left] = i; else parent_leaf[~left] = i;
    if (right >= 0) parent_internal[right] = i; else parent_leaf[~right] = i;
    // split axis: the first bit in which the range's first and last code differ (bit b of the code belongs to axis 2 - b % 3)
    const uint32_t x = keys[lo] ^ keys[hi];
    axis[i] = x ? (uint8_t)(2 - (31 - __clz(x)) % 3) : (uint8_t)0;
    if (i == 0) parent_internal[0] = -1;
}

__global__ void refit_kernel(const float* __restrict__ group_bounds, const uint32_t* __restrict__ sorted_group, int m,
                             const int2* __restrict__ children, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf,
                             float* __restrict__ node_bounds, unsigned int* __restrict__ arrivals) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= m) return;
    int node = parent_leaf[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrivals[node], 1u) == 0u) return;           // the first arrival stops; the second sees both children done
        const int2 c = children[node];
        float b[6];
        for (int side = 0; side < 2; ++side) {
            const int r = side ? c.y : c.x;
            // boxes of internal children were written by other threads of this launch: read them past L1 (__ldcg)
            const float* src = r >= 0 ? node_bounds + 6 * (size_t)r : group_bounds + 6 * (size_t)sorted_group[~r];
            for (int k = 0; k < 3; ++k) {
                const float lo = __ldcg(src + k), hi = __ldcg(src + 3 + k);
                b[k] = side ? fminf(b[k], lo) : lo;
                b[3 + k] = side ? fmaxf(b[3 + k], hi) : hi;
            }
        }
        for (int k = 0; k < 6; ++k) node_bounds[6 * (size_t)node + k] = b[k];
        node = parent_internal[node];
    }
}

}  // namespace

size_t lbvh_scratch_bytes(uint32_t m) {
    const size_t n_blocks = (m + SCAN_BLOCK - 1) / SCAN_BLOCK;
    // keys x2, vals x2, excl, block sums (+1), parents (internal + leaf), arrivals
    return (size_t)m * 4 * 5 + (n_blocks + 1) * 4 + (size_t)m * 4 * 3 + 256;
}

cudaError_t lbvh_build(const float* d_group_bounds, uint32_t m, const float scene_lo[3], const float scene_hi[3], void* d_scratch,
                       int2* d_children, float* d_node_bounds, uint8_t* d_axis, uint32_t* d_sorted_group, uint64_t* launches,
                       cudaStream_t st) {
    if (m < 2) return cudaErrorInvalidValue;
    const uint32_t n_blocks = (m + SCAN_BLOCK - 1) / SCAN_BLOCK;
    uint32_t* p = static_cast<uint32_t*>(d_scratch);
    uint32_t *keys[2] = { p, p + m }, *vals[2] = { p + 2 * (size_t)m, p + 3 * (size_t)m };
    uint32_t* excl = p + 4 * (size_t)m;
    uint32_t* sums = p + 5 * (size_t)m;
    int* parent_internal = reinterpret_cast<int*>(sums + n_blocks + 1);
    int* parent_leaf = parent_internal + m;
    unsigned int* arrivals = reinterpret_cast<unsigned int*>(parent_leaf + m);
    float3 lo = make_float3(scene_lo[0], scene_lo[1], scene_lo[2]);
    float3 inv = make_float3(scene_hi[0] > scene_lo[0] ? 1.0f / (scene_hi[0] - scene_lo[0]) : 0.0f,
                             scene_hi[1] > scene_lo[1] ? 1.0f / (scene_hi[1] - scene_lo[1]) : 0.0f,
                             scene_hi[2] > scene_lo[2] ? 1.0f / (scene_hi[2] - scene_lo[2]) : 0.0f);
    const unsigned g256 = (m + 255u) / 256u;
    morton_kernel<<<g256, 256, 0, st>>>(d_group_bounds, m, lo, inv, keys[0], vals[0]);
    int cur = 0;
    for (uint32_t bit = 0; bit < 30; ++bit) {
        flag_scan_kernel<<<n_blocks, SCAN_BLOCK, 0, st>>>(keys[cur], m, bit, excl, sums);
        sums_scan_kernel<<<1, SCAN_BLOCK, 0, st>>>(sums, n_blocks);
        split_scatter_kernel<<<n_blocks, SCAN_BLOCK, 0, st>>>(keys[cur], vals[cur], m, bit, excl, sums, n_blocks, keys[cur ^ 1], vals[cur ^ 1]);
        cur ^= 1;
    }
    cudaError_t e = cudaMemcpyAsync(d_sorted_group, vals[cur], (size_t)m * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(arrivals, 0, (size_t)m * 4, st);
    if (e != cudaSuccess) return e;
    karras_kernel<<<(m - 1 + 255u) / 256u, 256, 0, st>>>(keys[cur], (int)m, d_children, parent_internal, parent_leaf, d_axis);
    refit_kernel<<<g256, 256, 0, st>>>(d_group_bounds, d_sorted_group, (int)m, d_children, parent_internal, parent_leaf, d_node_bounds, arrivals);
    if (launches) *launches += 1 + 30 * 3 + 2;
    return cudaGetLastError();
}

}  // namespace b2rt
