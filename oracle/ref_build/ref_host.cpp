// Driver TU 1 of oracle/_ref: exposes the reference's CLOBJloader and
// CLBVHScene (compiled verbatim from /root/reference) behind ref_api.h.
// TEST INFRASTRUCTURE ONLY.
#include "stdafx.h"
#define private public   // CLBVHScene::m_Nodes is private (clBVHnode.h:77)
#include "CLBVHnode.h"
#undef private
#include "CLOBJloader.h"
#include "CLRaytracer.h"
#include "CLEngineBase.h"
#include "ref_api.h"
#include <mutex>

namespace Glaze3D { std::shared_ptr<CLEngineBase> eng; }  // main.cpp:5
using namespace Glaze3D;

static std::mutex g_lock;  // the reference funnels everything through the global `eng`

struct RefScene { std::shared_ptr<CLBVHScene> scene; };

static void install(std::shared_ptr<CLBVHScene> s) {
    eng = std::make_shared<CLEngineBase>();
    eng->render = std::make_shared<CLRaytracer>();
    eng->render->m_Scene = s;
}

extern "C" void* ref_scene_load(const char* obj_path, unsigned max_prims) {
    std::lock_guard<std::mutex> g(g_lock);
    try {
        auto s = std::make_shared<CLBVHScene>();
        install(s);
        CLOBJloader loader;
        loader.Load(obj_path, max_prims);          // CLEngineBase.cpp:176-178
        s->CreateBVHTrees(max_prims);              // CLEngineBase.cpp:179
        eng.reset();
        return new RefScene{s};
    } catch (...) { eng.reset(); return nullptr; }
}

extern "C" void* ref_scene_from_triangles(const void* tris, uint64_t n_tris, const void* mats,
                                          uint64_t n_mats, unsigned max_prims) {
    std::lock_guard<std::mutex> g(g_lock);
    try {
        auto s = std::make_shared<CLBVHScene>();
        install(s);
        const CLTriangle* t = static_cast<const CLTriangle*>(tris);
        s->m_Triangles.assign(t, t + n_tris);
        const CLMaterial* m = static_cast<const CLMaterial*>(mats);
        s->m_Materials.assign(m, m + n_mats);
        s->CreateBVHTrees(max_prims);
        eng.reset();
        return new RefScene{s};
    } catch (...) { eng.reset(); return nullptr; }
}

extern "C" void ref_scene_free(void* p) { delete static_cast<RefScene*>(p); }
extern "C" uint64_t ref_scene_num_triangles(void* p) { return static_cast<RefScene*>(p)->scene->m_Triangles.size(); }
extern "C" uint64_t ref_scene_num_nodes(void* p) { return static_cast<RefScene*>(p)->scene->m_Nodes.size(); }
extern "C" uint64_t ref_scene_num_materials(void* p) { return static_cast<RefScene*>(p)->scene->m_Materials.size(); }
extern "C" const void* ref_scene_triangles(void* p) { return static_cast<RefScene*>(p)->scene->m_Triangles.data(); }
extern "C" const void* ref_scene_nodes(void* p) { return static_cast<RefScene*>(p)->scene->m_Nodes.data(); }
extern "C" const void* ref_scene_materials(void* p) { return static_cast<RefScene*>(p)->scene->m_Materials.data(); }

static_assert(sizeof(CLTriangle) == 256, "CLTriangle layout");
static_assert(sizeof(CLLinearBVHNode) == 48, "CLLinearBVHNode layout");
static_assert(sizeof(CLMaterial) == 64, "CLMaterial layout");
