// Shim for <glm/glm.hpp>: the reference only needs glm::vec3 as a POD for the
// float3(glm::vec3) constructor and the camera/light structs.
#pragma once
namespace glm {
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
}  // namespace glm
