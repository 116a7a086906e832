/* C API of oracle/_ref/libref_oracle.so: the reference's own loader, BVH
 * builder and kernel_bvh.cl compiled verbatim for the CPU.
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {      /* one ray: origin + (not necessarily unit) direction */
    float ox, oy, oz, pad0;
    float dx, dy, dz, pad1;
} RefRay;

typedef struct {      /* what kernel_bvh.cl:18-27 IntersectData carries */
    int32_t hit;      /* isect.hit */
    int32_t tri;      /* isect.object - triangles, -1 on miss */
    float t;          /* isect.t (100000 on miss) */
    float pos[3];     /* isect.pos (undefined on miss -> zeroed) */
    float normal[3];
    float uv[2];
} RefHit;

/* CLOBJloader::Load(path, maxPrims) + CLBVHScene::CreateBVHTrees(maxPrims)
 * (CLEngineBase.cpp:172-179). Returns an opaque scene or NULL. */
void* ref_scene_load(const char* obj_path, unsigned max_prims);
/* Same build, but from caller-provided pre-loader triangles (256 B each). */
void* ref_scene_from_triangles(const void* tris, uint64_t n_tris, const void* mats, uint64_t n_mats, unsigned max_prims);
void ref_scene_free(void* scene);
uint64_t ref_scene_num_triangles(void* scene);
uint64_t ref_scene_num_nodes(void* scene);
uint64_t ref_scene_num_materials(void* scene);
const void* ref_scene_triangles(void* scene);   /* CLTriangle[], 256 B each */
const void* ref_scene_nodes(void* scene);       /* CLLinearBVHNode[], 48 B each */
const void* ref_scene_materials(void* scene);   /* CLMaterial[], 64 B each */

/* kernel_bvh.cl Intersect() per ray (InitRay normalises dir). nodes visited /
 * triangles tested are not available from the verbatim kernel. */
void ref_intersect(const void* tris, const void* nodes, const RefRay* rays, uint64_t n, RefHit* out, int threads);

/* kernel_bvh.cl KernelEntry() for gid in [gid0, gid1). `result` is the
 * W*H*16-byte accumulation buffer (read-modify-write like the device buffer). */
void ref_render(const void* tris, const void* nodes, const void* mats, void* result,
                uint32_t width, uint32_t height, uint32_t frame_count, uint32_t frame_seed,
                int32_t light_bounces, int32_t light_type, float sky,
                const float* cam_pos, const float* cam_front, const float* cam_up,
                uint64_t gid0, uint64_t gid1, int threads);
#ifdef __cplusplus
}
#endif
