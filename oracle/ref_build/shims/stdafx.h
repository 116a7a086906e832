// Shim for the reference's Windows/GL precompiled header (stdafx.h:1-110).
#pragma once
#include <CL/cl2.hpp>
#include <glm/glm.hpp>
#include <cmath>
#include <memory>
#include <string>
#include <vector>
namespace Glaze3D {
class CLEngineBase;
extern std::shared_ptr<CLEngineBase> eng;
}
