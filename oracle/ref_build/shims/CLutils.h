// Shim for CLutils.h: the kernel-argument enum (CLutils.h:11-27), a GCC-safe
// CLException (the reference's uses an MSVC-only std::exception ctor,
// CLutils.h:107-114) and a CLContext that only hands out a dummy cl::Context.
#pragma once
#include "stdafx.h"
namespace Glaze3D {
enum class RenderKernelArgument_t : unsigned int {
    BUFFER_OUT, BUFFER_SCENE, BUFFER_NODE, BUFFER_MATERIAL, WIDTH, HEIGHT,
    FRAME_COUNT, FRAME_SEED, LIGHT_BOUNCES, LIGHT_TYPE, SKYBOX_INTENSITY,
    CAMERA_POS, CAMERA_FRONT, CAMERA_UP
};
class CLException : public std::exception {
public:
    CLException(const std::string& m, int code) : msg_(m + " (" + std::to_string(code) + ")") {}
    const char* what() const noexcept override { return msg_.c_str(); }
private:
    std::string msg_;
};
class CLContext {
public:
    const cl::Context& GetContext() const { return ctx_; }
private:
    cl::Context ctx_;
};
class CLKernel {};
}  // namespace Glaze3D
