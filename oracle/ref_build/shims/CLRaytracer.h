// Shim for CLRaytracer.h: the members the loader/builder dereference through
// the global `eng` (CLOBJloader.cpp:12, CLBVHnode.cpp:214-235).
#pragma once
#include "stdafx.h"
#include "CLutils.h"
#include "CLBVHnode.h"
namespace Glaze3D {
class CLRaytracer {
public:
    template <class T> bool SetUniform(int, T&) { return true; }
    std::shared_ptr<CLContext> m_CLContext = std::make_shared<CLContext>();
    std::shared_ptr<CLKernel> m_RenderKernel;
    std::shared_ptr<CLBVHScene> m_Scene;
};
}  // namespace Glaze3D
