// cl_shim.hpp -- TEST INFRASTRUCTURE (part of the CPU oracle, never shipped).
//
// A minimal OpenCL-C dialect for g++, so that the reference's own device
// scripts (resources/Scripts/**, read from /root/reference at build time by
// build_ref.py) compile as C++ functions, one call per work-item.  Provides the
// vector value types with the swizzles those scripts use, the handful of
// built-ins they call, and the address-space qualifiers as no-ops.
// The only textual change build_ref.py makes to the sources is turning the
// OpenCL vector literal "(float4)(a, b, c, d)" into "float4(a, b, c, d)",
// which C++ would otherwise parse as a cast of a comma expression.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned int usize; // 32-bit addressing, the reference default (State.cpp:499-502)
typedef int ssize;

#define __kernel extern "C"
#define __global
#define __constant const
#define __local
#define __private

struct float2;
struct float3;
struct float4;
struct float16;

// Swizzle proxy living inside the parent's storage: converts to / assigns from V.
template <class V, int PN, int... I>
struct Swz {
    float d[PN];
    operator V() const { return V(d[I]...); }
    Swz& operator=(const V& v)
    {
        int k = 0;
        ((d[I] = v[k++]), ...);
        return *this;
    }
    Swz& operator+=(const V& v)
    {
        int k = 0;
        ((d[I] += v[k++]), ...);
        return *this;
    }
    Swz& operator-=(const V& v)
    {
        int k = 0;
        ((d[I] -= v[k++]), ...);
        return *this;
    }
    Swz& operator*=(float s)
    {
        ((d[I] *= s), ...);
        return *this;
    }
    Swz& operator/=(float s)
    {
        ((d[I] /= s), ...);
        return *this;
    }
};


// ".TRANSPOSE" (= .s0213 / .s048C159D26AE37BF) is swizzled again by the matrix
// macros of types/{2D,3D}.h, so the transposed views carry their own swizzles.
struct TView4 {
    union {
        float d[4];
        Swz<float2, 4, 0, 2> s01;
        Swz<float2, 4, 1, 3> s23;
        Swz<float2, 4, 0, 1> s02;
        Swz<float2, 4, 2, 3> s13;
        Swz<float4, 4, 0, 2, 1, 3> all;
    };
    operator float4() const;
};
struct TView16 {
    union {
        float d[16];
        Swz<float3, 16, 0, 4, 8> s012;
        Swz<float3, 16, 1, 5, 9> s456;
        Swz<float3, 16, 2, 6, 10> s89A;
        Swz<float3, 16, 0, 1, 2> s048;
        Swz<float3, 16, 4, 5, 6> s159;
        Swz<float3, 16, 8, 9, 10> s26A;
        Swz<float4, 16, 0, 4, 8, 12> s0123;
        Swz<float4, 16, 1, 5, 9, 13> s4567;
        Swz<float4, 16, 2, 6, 10, 14> s89AB;
        Swz<float4, 16, 3, 7, 11, 15> sCDEF;
        Swz<float4, 16, 0, 1, 2, 3> s048C;
        Swz<float4, 16, 4, 5, 6, 7> s159D;
        Swz<float4, 16, 8, 9, 10, 11> s26AE;
        Swz<float4, 16, 12, 13, 14, 15> s37BF;
        Swz<float16, 16, 0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15> all;
    };
    operator float16() const;
};

struct float2 {
    union {
        struct { float x, y; };
        struct { float s0, s1; };
        float d[2];
        Swz<float2, 2, 0, 1> xy;
    };
    float2() : x(0.f), y(0.f) {}
    explicit float2(float a) : x(a), y(a) {}
    float2(float a, float b) : x(a), y(b) {}
    float operator[](int k) const { return d[k]; }
    float& operator[](int k) { return d[k]; }
};

struct float3 {
    union {
        struct { float x, y, z; };
        struct { float s0, s1, s2; };
        float d[4];
        Swz<float3, 4, 0, 1, 2> xyz;
        Swz<float2, 4, 0, 1> xy;
    };
    float3() : x(0.f), y(0.f), z(0.f) {}
    explicit float3(float a) : x(a), y(a), z(a) {}
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int k) const { return d[k]; }
    float& operator[](int k) { return d[k]; }
};

struct float4 {
    union {
        struct { float x, y, z, w; };
        struct { float s0, s1, s2, s3; };
        float d[4];
        Swz<float3, 4, 0, 1, 2> xyz;
        Swz<float2, 4, 0, 1> xy;
        Swz<float2, 4, 0, 1> s01;
        Swz<float2, 4, 2, 3> s23;
        Swz<float2, 4, 0, 2> s02;
        Swz<float2, 4, 1, 3> s13;
        Swz<float2, 4, 0, 3> s03;
        TView4 s0213;
    };
    float4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    explicit float4(float a) : x(a), y(a), z(a), w(a) {}
    float4(float a, float b, float c, float e) : x(a), y(b), z(c), w(e) {}
    float operator[](int k) const { return d[k]; }
    float& operator[](int k) { return d[k]; }
};

struct float16 {
    union {
        struct { float s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sA, sB, sC, sD, sE, sF; };
        float d[16];
        Swz<float3, 16, 0, 1, 2> s012;
        Swz<float3, 16, 4, 5, 6> s456;
        Swz<float3, 16, 8, 9, 10> s89A;
        Swz<float3, 16, 0, 4, 8> s048;
        Swz<float3, 16, 1, 5, 9> s159;
        Swz<float3, 16, 2, 6, 10> s26A;
        Swz<float3, 16, 0, 5, 10> s05A;
        Swz<float4, 16, 0, 1, 2, 3> s0123;
        Swz<float4, 16, 4, 5, 6, 7> s4567;
        Swz<float4, 16, 8, 9, 10, 11> s89AB;
        Swz<float4, 16, 12, 13, 14, 15> sCDEF;
        Swz<float4, 16, 0, 4, 8, 12> s048C;
        Swz<float4, 16, 1, 5, 9, 13> s159D;
        Swz<float4, 16, 2, 6, 10, 14> s26AE;
        Swz<float4, 16, 3, 7, 11, 15> s37BF;
        TView16 s048C159D26AE37BF;
    };
    float16() { for (int k = 0; k < 16; k++) d[k] = 0.f; }
    explicit float16(float a) { for (int k = 0; k < 16; k++) d[k] = a; }
    float16(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
            float a8, float a9, float aA, float aB, float aC, float aD, float aE, float aF)
    {
        const float t[16] = { a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, aA, aB, aC, aD, aE, aF };
        for (int k = 0; k < 16; k++) d[k] = t[k];
    }
    float operator[](int k) const { return d[k]; }
    float& operator[](int k) { return d[k]; }
};

inline TView4::operator float4() const { return float4(d[0], d[2], d[1], d[3]); }
inline TView16::operator float16() const
{
    return float16(d[0], d[4], d[8], d[12], d[1], d[5], d[9], d[13], d[2], d[6], d[10], d[14], d[3], d[7],
                   d[11], d[15]);
}

template <class T> struct vec_size;
template <> struct vec_size<float2> { static constexpr int n = 2; };
template <> struct vec_size<float3> { static constexpr int n = 3; };
template <> struct vec_size<float4> { static constexpr int n = 4; };
template <> struct vec_size<float16> { static constexpr int n = 16; };

// Non-template operators, so that swizzle proxies convert implicitly.
#define CLSHIM_OPS(T)                                                                            \
    inline T operator+(const T& a, const T& b) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] + b[k]; return r; } \
    inline T operator-(const T& a, const T& b) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] - b[k]; return r; } \
    inline T operator*(const T& a, const T& b) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] * b[k]; return r; } \
    inline T operator/(const T& a, const T& b) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] / b[k]; return r; } \
    inline T operator*(const T& a, float s) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] * s; return r; }       \
    inline T operator*(float s, const T& a) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = s * a[k]; return r; }       \
    inline T operator/(const T& a, float s) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] / s; return r; }       \
    inline T operator/(float s, const T& a) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = s / a[k]; return r; }       \
    inline T operator+(const T& a, float s) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] + s; return r; }       \
    inline T operator-(const T& a, float s) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = a[k] - s; return r; }       \
    inline T operator-(const T& a) { T r; for (int k = 0; k < vec_size<T>::n; k++) r[k] = -a[k]; return r; }                   \
    inline T& operator+=(T& a, const T& b) { for (int k = 0; k < vec_size<T>::n; k++) a[k] += b[k]; return a; }                \
    inline T& operator-=(T& a, const T& b) { for (int k = 0; k < vec_size<T>::n; k++) a[k] -= b[k]; return a; }                \
    inline T& operator*=(T& a, const T& b) { for (int k = 0; k < vec_size<T>::n; k++) a[k] *= b[k]; return a; }                \
    inline T& operator*=(T& a, float s) { for (int k = 0; k < vec_size<T>::n; k++) a[k] *= s; return a; }                      \
    inline T& operator/=(T& a, float s) { for (int k = 0; k < vec_size<T>::n; k++) a[k] /= s; return a; }
CLSHIM_OPS(float2)
CLSHIM_OPS(float3)
CLSHIM_OPS(float4)
CLSHIM_OPS(float16)

// dot/length: products accumulated left to right in fp32 (no contraction: the
// build uses -ffp-contract=off, like the restatement it is compared with)
inline float dot(const float2& a, const float2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const float2& a) { return sqrtf(dot(a, a)); }
inline float length(const float3& a) { return sqrtf(dot(a, a)); }
inline float length(const float4& a) { return sqrtf(dot(a, a)); }
inline float3 cross(const float3& a, const float3& b)
{
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float4 cross(const float4& a, const float4& b)
{
    return float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.f);
}
inline float fabs(float x) { return fabsf(x); }
inline float sqrt(float x) { return sqrtf(x); }
inline float pow(float x, float y) { return powf(x, y); }
inline float atan(float x) { return atanf(x); }
inline float log(float x) { return logf(x); }
inline float cos(float x) { return cosf(x); }
inline float sin(float x) { return sinf(x); }
// OpenCL normalize: a vector of the same direction and length 1 (here: v / length(v))
inline float2 normalize(const float2& a) { const float l = length(a); return float2(a.x / l, a.y / l); }
inline float4 normalize(const float4& a)
{
    const float l = length(a);
    return float4(a.x / l, a.y / l, a.z / l, a.w / l);
}
inline float acospi(float x) { return acosf(x) * 0.318309886183790671538f; }
inline float min(float a, float b) { return b < a ? b : a; }   // OpenCL: y if y < x, else x
inline float max(float a, float b) { return a < b ? b : a; }   // OpenCL: y if x < y, else x
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline float sign(float x)
{
    if (x != x) return 0.f;
    return x > 0.f ? 1.f : (x < 0.f ? -1.f : x);
}
inline int isnan(float x) { return x != x; }
inline int isinf(float x) { return std::isinf(x); }

template <class T> struct tvec4 { T x, y, z, w; };
template <class T> struct tvec3 { T x, y, z; };
template <class T> struct tvec2 { T x, y; };
typedef tvec4<unsigned int> uint4;
typedef tvec4<int> int4;
typedef tvec4<usize> usize4;
typedef tvec4<ssize> ssize4;
typedef tvec2<unsigned int> uint2;
typedef tvec2<int> int2;
typedef tvec2<usize> usize2;
typedef tvec3<unsigned int> uint3;
typedef tvec3<int> int3;
typedef tvec3<usize> usize3;

// one work-item at a time; work-groups of one item (LOCAL_MEM_SIZE = 1)
extern thread_local size_t clshim_gid;
inline size_t get_global_id(int) { return clshim_gid; }
inline size_t get_local_id(int) { return 0; }

// "-D" definitions the reference bakes into every program (CalcServer.cpp:240-265,
// basic.xml:119-123): here run-time values set through aqref_set_defs()
struct clshim_defs { float v_h, v_conw, v_conf, v_support, v_dims; };
extern clshim_defs clshim_D;
#define H (clshim_D.v_h)
#define CONW (clshim_D.v_conw)
#define CONF (clshim_D.v_conf)
#define SUPPORT (clshim_D.v_support)
#define DIMS (clshim_D.v_dims)
#define KERNEL_NAME Wendland
#define LOCAL_MEM_SIZE 1
#define __LAP_MONAGHAN__ 1
#define __LAP_MORRIS__ 2
#ifndef __LAP_FORMULATION__ // (build_ref.py VARIANTS: cfd/Interactions@morris.cl)
#define __LAP_FORMULATION__ __LAP_MONAGHAN__
#endif
#define NDEBUG
