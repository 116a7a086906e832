/* rt_oracle.h -- CPU restatement of the reference's hot path (kernel_bvh.cl).
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle.so; the
 * product (mini-opencl-raytracer_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED against the reference's own sources compiled verbatim
 * (oracle/_ref, see oracle/build_ref.py) -- tests/test_oracle_vs_ref.py checks
 * bit-equality of hits and frames in this container, and the golden vectors in
 * tests/golden/ were generated from oracle/_ref by tests/golden/make_golden.py.
 * The reference ships no tests or golden vectors of its own and
 * no OpenCL implementation exists in this image, so OpenCL built-ins whose
 * precision is implementation-defined follow the convention documented in
 * oracle/ref_build/cl_shim.hpp (IEEE fp32 ops, no FMA, glibc libm).
 *
 * All structs mirror the byte layouts of /root/reference/CLshared_structs.hpp
 * (sizes 256 / 48 / 64 B; offsets in SURVEY.md 8a).
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z, w; } OrVec;                      /* 16-byte float3 */
typedef struct { OrVec position, uv, normal, tangent_s, tangent_t; } OrVertex;   /* 80 B */
typedef struct { OrVertex v1, v2, v3; uint32_t mtlIndex; uint32_t padding[3]; } OrTriangle; /* 256 B */
typedef struct { OrVec bmin, bmax; uint32_t offset; uint16_t nPrimitives; uint8_t axis; uint8_t pad[9]; } OrNode; /* 48 B */
typedef struct { OrVec diffuse, specular, emission; uint32_t type; float roughness, ior; int32_t padding; } OrMaterial; /* 64 B */

/* Ray-stream records shared with the product's C ABI (include/b2rt.h). */
typedef struct { float ox, oy, oz, tmin; float dx, dy, dz, tmax; } OrRay;   /* 32 B */
typedef struct { float t, u, v; uint32_t tri; } OrHit;                       /* 16 B, tri = 0xFFFFFFFF on miss */
typedef struct { float pos[3]; float normal[3]; float uv[2]; } OrHitAttr;   /* what IntersectData also carries */
typedef struct { uint64_t nodes_visited, leaves_entered, tris_tested; } OrCounters;

/* Intersect() (kernel_bvh.cl:171-219) per ray with isect.t initialised to
 * ray.tmax (the reference uses MAX_RENDER_DIST = 100000, kernel_bvh.cl:7,176).
 * ray.tmin is ignored: the reference has no lower bound (kernel_bvh.cl:140).
 * attr / counters may be NULL. threads <= 0 -> all hardware threads. */
void oracle_trace_closest(const OrTriangle* tris, const OrNode* nodes, const OrRay* rays, uint64_t n,
                          OrHit* hits, OrHitAttr* attr, OrCounters* counters, int threads);
/* occluded[i] = Intersect(ray, tmax).hit. */
void oracle_trace_any(const OrTriangle* tris, const OrNode* nodes, const OrRay* rays, uint64_t n,
                      uint32_t* occluded, int threads);
/* KernelEntry() (kernel_bvh.cl:415-456) for gid in [gid0, gid1); `result` is the
 * W*H*16-byte accumulation buffer, read-modify-written like the device buffer. */
void oracle_render(const OrTriangle* tris, const OrNode* nodes, const OrMaterial* mats, float* result,
                   uint32_t width, uint32_t height, uint32_t frame_count,
                   int32_t light_bounces, int32_t light_type, float sky,
                   const float* cam_pos, const float* cam_front, const float* cam_up,
                   uint64_t gid0, uint64_t gid1, int threads);
/* CreateRay() (kernel_bvh.cl:386-403) for gid in [gid0,gid1): the primary ray
 * stream of one frame, as OrRay with tmax = 100000. */
void oracle_camera_rays(uint32_t width, uint32_t height, uint32_t frame_count,
                        const float* cam_pos, const float* cam_front, const float* cam_up,
                        uint64_t gid0, uint64_t gid1, OrRay* rays);
int oracle_hardware_threads(void);
#ifdef __cplusplus
}
#endif
#endif
