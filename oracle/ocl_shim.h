// oracle/ocl_shim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Minimal OpenCL-C-on-host shim so that the reference's own kernel source
// (/root/reference/src/kernels/sim_kernels.cl, text after "EndTempDefines%") compiles
// unmodified as C++17 with g++.  The only textual rewrite applied by build_ref.py is the
// OpenCL vector-literal cast "(float3)(" -> "float3(" (same for uint3/int3).
//
// What this header provides is listed in SURVEY.md section 8(c): scalar typedefs, float3/uint3/int3
// with component-wise arithmetic and a broadcast constructor, the OpenCL built-ins the kernels
// call (fma, clamp, min, cross, dot, length, convert_float3, as_uint/as_float,
// vload_half/vstore_half_rte, atomic_xchg, get_global_id) and empty address-space qualifiers.
#pragma once
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef unsigned int uint;
typedef unsigned long ulong;   // OpenCL ulong is 64 bit; LP64 host: unsigned long is 64 bit
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef _Float16 half;         // IEEE binary16; conversions round-to-nearest-even
static_assert(sizeof(ulong) == 8, "LP64 host required");

template <typename T> struct vec3 {
    T x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(T a, T b, T c) : x(a), y(b), z(c) {}
    vec3(T a) : x(a), y(a), z(a) {}  // OpenCL scalar -> vector broadcast
    template <typename A, typename B, typename C> vec3(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
};
typedef vec3<float> float3;
typedef vec3<uint> uint3;
typedef vec3<int> int3;

#define ION_VEC_OP(op)                                                                                       \
    template <typename T> inline vec3<T> operator op(const vec3<T>& a, const vec3<T>& b) {                   \
        return vec3<T>(a.x op b.x, a.y op b.y, a.z op b.z);                                                  \
    }                                                                                                        \
    template <typename T, typename S> inline vec3<T> operator op(const vec3<T>& a, const S b) {              \
        return vec3<T>(a.x op(T) b, a.y op(T) b, a.z op(T) b);                                               \
    }                                                                                                        \
    template <typename T, typename S> inline vec3<T> operator op(const S a, const vec3<T>& b) {              \
        return vec3<T>((T)a op b.x, (T)a op b.y, (T)a op b.z);                                               \
    }
ION_VEC_OP(+)
ION_VEC_OP(-)
ION_VEC_OP(*)
ION_VEC_OP(/)
#undef ION_VEC_OP
template <typename T> inline vec3<T> operator-(const vec3<T>& a) { return vec3<T>(-a.x, -a.y, -a.z); }
template <typename T> inline vec3<T>& operator+=(vec3<T>& a, const vec3<T>& b) {
    a.x += b.x; a.y += b.y; a.z += b.z;
    return a;
}
template <typename T> inline vec3<T>& operator-=(vec3<T>& a, const vec3<T>& b) {
    a.x -= b.x; a.y -= b.y; a.z -= b.z;
    return a;
}

inline float3 convert_float3(const uint3& v) { return float3((float)v.x, (float)v.y, (float)v.z); }
inline float3 convert_float3(const int3& v) { return float3((float)v.x, (float)v.y, (float)v.z); }
inline float3 cross(const float3& a, const float3& b) {  // OpenCL cross(): plain mul/sub, no fma
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const float3& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }

// fma(float,float,float) and sqrt(float) come from <math.h> (libstdc++ exports the float overloads globally)
static_assert(sizeof(decltype(fma(1.0f, 1.0f, 1.0f))) == 4 && sizeof(decltype(sqrt(1.0f))) == 4, "float overloads required");
inline float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline uint as_uint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float as_float(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float vload_half(size_t o, const half* p) { return (float)p[o]; }
inline void vstore_half_rte(float x, size_t o, half* p) { p[o] = (half)x; }
inline float atomic_xchg(volatile float* addr, float v) {  // selects sim.cl:125-126 fallback atomic_add_f
    uint nv = as_uint(v), old;
    __atomic_exchange((volatile uint*)addr, &nv, &old, __ATOMIC_SEQ_CST);
    return as_float(old);
}
#ifndef M_PI_F
#define M_PI_F 3.14159274101257f
#endif

static thread_local size_t ion_shim_gid = 0;
inline size_t get_global_id(uint) { return ion_shim_gid; }

// address-space / kernel qualifiers are no-ops on the host (defined last: keep std headers clean)
#define __kernel
#define kernel
#define __global
#define global
#define printf(...) ((void)0)  // sim.cl:541-542,575-601,628,642,824 debug prints from cell 0 (quirk Q16)
