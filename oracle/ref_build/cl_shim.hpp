// cl_shim.hpp — just enough OpenCL C for g++ to compile the reference's OWN kernel file
// (/root/reference/.../OptimizedConvolution.cl) as C++ and run it on the CPU.  TEST INFRASTRUCTURE.
//
// Nothing here restates the reference's algorithm: argmin rule, matrices, constants, reflect
// indexing, fma nesting and transposition all come from the .cl text, compiled where it lies.
// What the shim has to supply is what OpenCL leaves to the device: the address-space keywords,
// the float4 type, work-item ids, and the builtins pow / cbrt / dot / distance / fma.  Their
// definitions below are the same pins oracle/hq_oracle.c documents ("Device-defined OpenCL
// builtins are pinned"), written independently of it:
//   fma       exact (IEEE fused multiply-add, one rounding)
//   dot       products and sums left to right, each rounded to fp32, no contraction
//   distance  sqrtf of the squared norm accumulated as mul, fma, fma(, fma) over x, y, z(, w)
//   pow(x,y)  the double-precision libm value narrowed to fp32
//   cbrt      the long-double libm value narrowed to fp32 (correctly rounded fp32)
// Compile with -ffp-contract=off.
#pragma once
#include <cmath>

#define __kernel static inline
#define __global
#define __constant const

struct float3 {
    float x, y, z;
};
struct alignas(16) float4 {
    union {
        struct { float x, y, z, w; };
        float3 xyz;  // the one swizzle the kernels use (CIEDE, cl:209)
    };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4& operator+=(const float4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
};

namespace refcl {
extern thread_local int g_global_id;
static inline int get_global_id(int) { return g_global_id; }

static inline float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
static inline float4 fma(const float4& a, const float4& b, const float4& c) {
    return float4(::fmaf(a.x, b.x, c.x), ::fmaf(a.y, b.y, c.y), ::fmaf(a.z, b.z, c.z), ::fmaf(a.w, b.w, c.w));
}
static inline float dot(const float4& a, const float4& b) {
    float s = a.x * b.x;
    s = s + a.y * b.y;
    s = s + a.z * b.z;
    s = s + a.w * b.w;
    return s;
}
static inline float distance(const float3& a, const float3& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return ::sqrtf(::fmaf(dz, dz, ::fmaf(dy, dy, dx * dx)));
}
static inline float distance(const float4& a, const float4& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    return ::sqrtf(::fmaf(dw, dw, ::fmaf(dz, dz, ::fmaf(dy, dy, dx * dx))));
}
static inline float pow(float a, float b) { return (float)::pow((double)a, (double)b); }
static inline float cbrt(float a) { return (float)::cbrtl((long double)a); }
static inline float sqrt(float a) { return ::sqrtf(a); }
}  // namespace refcl
