// Shim for CLEngineBase.h: only the `render` member is reachable from the
// sources the oracle compiles.
#pragma once
#include "stdafx.h"
namespace Glaze3D {
class CLRaytracer;
class CLEngineBase {
public:
    std::shared_ptr<CLRaytracer> render;
};
}  // namespace Glaze3D
