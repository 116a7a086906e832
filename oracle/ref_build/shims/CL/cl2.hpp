// Shim for <CL/cl2.hpp>: just enough surface for the reference's host sources
// (CLBVHnode.cpp, CLOBJloader.cpp) to compile on a box with no OpenCL SDK.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build) - never part of the product.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <exception>
#include <limits>
#include <memory>
#include <string>
#include <vector>

typedef int cl_int;
#define CL_SUCCESS 0
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_FLT_MAX 340282346638528859811704183484516925440.0f

namespace cl {
struct Context {};
struct Buffer {
    Buffer() {}
    // A "device buffer" in the oracle is just a view of the host array.
    Buffer(const Context&, int, size_t bytes, const void* host, cl_int* err)
        : ptr(host), size(bytes) { if (err) *err = CL_SUCCESS; }
    const void* ptr = nullptr;
    size_t size = 0;
};
}  // namespace cl
