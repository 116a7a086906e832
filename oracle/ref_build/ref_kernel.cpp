// Driver TU 2 of oracle/_ref: the reference's kernel_bvh.cl compiled as C++
// (3-rule mechanical transform, see build_ref.py) behind ref_api.h.
// TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "ref_api.h"

namespace ocl {
#include "cl_shim.hpp"
#include "kernel_bvh.inc"   // generated at build time into a temp dir, never committed
}

static_assert(sizeof(ocl::CLTriangle) == 256, "CLTriangle layout");
static_assert(sizeof(ocl::CLLinearBVHNode) == 48, "CLLinearBVHNode layout");
static_assert(sizeof(ocl::CLMaterial) == 64, "CLMaterial layout");

template <class F>
static void parallel_for(uint64_t lo, uint64_t hi, int threads, F f) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > hi - lo) threads = (int)(hi > lo ? hi - lo : 1);
    if (threads == 1) { f(lo, hi); return; }
    std::vector<std::thread> pool;
    uint64_t n = hi - lo;
    for (int i = 0; i < threads; ++i) {
        uint64_t a = lo + n * i / threads, b = lo + n * (i + 1) / threads;
        pool.emplace_back([=] { f(a, b); });
    }
    for (auto& t : pool) t.join();
}

extern "C" void ref_intersect(const void* tris, const void* nodes, const RefRay* rays, uint64_t n,
                              RefHit* out, int threads) {
    ocl::Scene scene;
    std::memset((void*)&scene, 0, sizeof(scene));
    scene.triangles = (ocl::CLTriangle*)tris;
    scene.nodes = (ocl::CLLinearBVHNode*)nodes;
    parallel_for(0, n, threads, [&](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; ++i) {
            ocl::Ray r = ocl::InitRay(ocl::float3(rays[i].ox, rays[i].oy, rays[i].oz),
                                      ocl::float3(rays[i].dx, rays[i].dy, rays[i].dz));
            ocl::IntersectData d = ocl::Intersect(&r, &scene);
            RefHit h;
            std::memset(&h, 0, sizeof(h));
            h.hit = d.hit ? 1 : 0;
            h.t = d.t;
            h.tri = -1;
            if (d.hit) {
                h.tri = (int32_t)(d.object - scene.triangles);
                h.pos[0] = d.pos.x; h.pos[1] = d.pos.y; h.pos[2] = d.pos.z;
                h.normal[0] = d.normal.x; h.normal[1] = d.normal.y; h.normal[2] = d.normal.z;
                h.uv[0] = d.uv.x; h.uv[1] = d.uv.y;
            }
            out[i] = h;
        }
    });
}

extern "C" void ref_render(const void* tris, const void* nodes, const void* mats, void* result,
                           uint32_t width, uint32_t height, uint32_t frame_count, uint32_t frame_seed,
                           int32_t light_bounces, int32_t light_type, float sky,
                           const float* cam_pos, const float* cam_front, const float* cam_up,
                           uint64_t gid0, uint64_t gid1, int threads) {
    ocl::float3 p(cam_pos[0], cam_pos[1], cam_pos[2]);
    ocl::float3 f(cam_front[0], cam_front[1], cam_front[2]);
    ocl::float3 u(cam_up[0], cam_up[1], cam_up[2]);
    parallel_for(gid0, gid1, threads, [&](uint64_t a, uint64_t b) {
        for (uint64_t g = a; g < b; ++g) {
            ocl::g_global_id = (unsigned int)g;
            ocl::KernelEntry((ocl::float3*)result, (ocl::CLTriangle*)tris, (ocl::CLLinearBVHNode*)nodes,
                             (ocl::CLMaterial*)mats, width, height, frame_count, frame_seed,
                             light_bounces, light_type, sky, p, f, u);
        }
    });
}
