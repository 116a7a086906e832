// OpenCL-C-as-C++ shim used ONLY to compile the reference's kernel_bvh.cl on
// the CPU (oracle/_ref). Included inside `namespace ocl { ... }`.
// TEST INFRASTRUCTURE ONLY - never part of the product.
//
// Conventions for built-ins whose precision OpenCL leaves implementation
// defined (there is no OpenCL implementation in this image):
//   + - * / sqrt      : IEEE fp32, one rounding per op (-ffp-contract=off)
//   dot               : x*x' + y*y' + z*z' evaluated left to right
//   cross             : (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   normalize(v)      : v / sqrtf(dot(v,v))  (three divisions)
//   max(x,y)          : x < y ? y : x        (OpenCL 1.2 spec 6.12.4 wording)
//   min(x,y)          : y < x ? y : x
//   pow/cos/sin/tan   : glibc powf/cosf/sinf/tanf
struct alignas(16) float3 {
    float x, y, z, w;
    float3() : x(0), y(0), z(0), w(0) {}
    float3(float s) : x(s), y(s), z(s), w(0) {}  // implicit: `float3 r = 0.0f;`
    float3(float a, float b, float c) : x(a), y(b), z(c), w(0) {}
    float3& operator+=(const float3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    float3& operator*=(const float3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
};
static inline float3 make_float3(float a, float b, float c) { return float3(a, b, c); }

static inline float3 operator+(const float3& a, const float3& b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(const float3& a, const float3& b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(const float3& a, const float3& b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator/(const float3& a, const float3& b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float3 operator-(const float3& a) { return float3(-a.x, -a.y, -a.z); }
static inline float3 operator*(const float3& a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(float s, const float3& a) { return float3(s * a.x, s * a.y, s * a.z); }
static inline float3 operator/(const float3& a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
static inline float3 operator/(float s, const float3& a) { return float3(s / a.x, s / a.y, s / a.z); }
static inline float3 operator+(const float3& a, float s) { return float3(a.x + s, a.y + s, a.z + s); }
static inline float3 operator+(float s, const float3& a) { return float3(s + a.x, s + a.y, s + a.z); }
static inline float3 operator-(const float3& a, float s) { return float3(a.x - s, a.y - s, a.z - s); }
static inline float3 operator-(float s, const float3& a) { return float3(s - a.x, s - a.y, s - a.z); }

static inline float sqrt(float v) { return ::sqrtf(v); }
static inline float cos(float v) { return ::cosf(v); }
static inline float sin(float v) { return ::sinf(v); }
static inline float tan(float v) { return ::tanf(v); }
static inline float fabs(float v) { return ::fabsf(v); }
static inline float pow(float a, float b) { return ::powf(a, b); }
static inline float max(float a, float b) { return a < b ? b : a; }
static inline float min(float a, float b) { return b < a ? b : a; }
static inline float3 pow(const float3& a, float b) { return float3(pow(a.x, b), pow(a.y, b), pow(a.z, b)); }
static inline float3 max(const float3& a, float b) { return float3(max(a.x, b), max(a.y, b), max(a.z, b)); }
static inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 cross(const float3& a, const float3& b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float3 normalize(const float3& v) { return v / sqrt(dot(v, v)); }

#define __global
#define __kernel
static thread_local unsigned int g_global_id = 0;
static inline unsigned int get_global_id(int) { return g_global_id; }
