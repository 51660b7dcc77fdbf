// clc_prelude.cuh -- the device-side dialect header of the run-time script path (clc.cu).
//
// NOT a header of the library itself: its text is embedded in libaquacuda.so (clc.cu includes it as a
// raw string) and prepended to every user script handed to NVRTC.  It lets an OpenCL C script of the
// reference's kind (resources/Scripts/**, case-local *.cl; compiled there by clBuildProgram,
// aquagpusph/CalcServer/Kernel.cpp:354-420) compile as sm_100a CUDA: address-space qualifiers, the work-item
// functions, vector value types with the swizzles and operators those scripts use, and the few built-ins
// they call.  The only textual change clc.cu makes to a script is the OpenCL vector literal
// "(float4)(a, b, c, d)" -> "float4(a, b, c, d)".
R"CLC(
typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned short ushort;
typedef unsigned char uchar;
// (the unprefixed spellings `kernel` / `global` cannot be offered: CUDA's __global__ is itself the macro
// __attribute__((global)), which a macro named `global` would empty)
#define __kernel extern "C" __global__
#define __global
#define __constant const
#define __local __shared__
#define __private
#define restrict __restrict__
#define CLC __device__ __forceinline__
#ifndef FLT_MAX
#define FLT_MAX 3.402823466e+38f
#endif
#ifndef FLT_MIN
#define FLT_MIN 1.175494351e-38f
#endif
#ifndef FLT_EPSILON
#define FLT_EPSILON 1.192092896e-07f
#endif
#ifndef INFINITY
#define INFINITY (__int_as_float(0x7f800000))
#endif
#ifndef NAN
#define NAN (__int_as_float(0x7fc00000))
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define M_PI_F 3.14159274101257f
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2

CLC uint get_global_id(int) { return blockIdx.x * blockDim.x + threadIdx.x; }
CLC uint get_local_id(int) { return threadIdx.x; }
CLC uint get_group_id(int) { return blockIdx.x; }
CLC uint get_local_size(int) { return blockDim.x; }
CLC uint get_global_size(int) { return gridDim.x * blockDim.x; }
CLC uint get_num_groups(int) { return gridDim.x; }
CLC void barrier(int) { __syncthreads(); }

struct clc_float2;
struct clc_float3;
struct clc_float4;
struct clc_float16;

// a swizzle lives inside its parent's storage: converts to / assigns from the vector type V
template <class V, int PN, int... I>
struct clc_swz {
    float d[PN];
    CLC operator V() const { return V(d[I]...); }
    CLC clc_swz& operator=(const V& v) { int k = 0; ((d[I] = v[k++]), ...); return *this; }
    CLC clc_swz& operator=(const clc_swz& o) { const V v = o; int k = 0; ((d[I] = v[k++]), ...); return *this; }
    CLC clc_swz& operator+=(const V& v) { int k = 0; ((d[I] += v[k++]), ...); return *this; }
    CLC clc_swz& operator-=(const V& v) { int k = 0; ((d[I] -= v[k++]), ...); return *this; }
    CLC clc_swz& operator*=(float s) { ((d[I] *= s), ...); return *this; }
    CLC clc_swz& operator/=(float s) { ((d[I] /= s), ...); return *this; }
};
// ".TRANSPOSE" (types/2D.h: s0213, 3D.h: s048C159D26AE37BF) is swizzled again by the matrix macros
struct clc_tview4 {
    union {
        float d[4];
        clc_swz<clc_float2, 4, 0, 2> s01;
        clc_swz<clc_float2, 4, 1, 3> s23;
        clc_swz<clc_float2, 4, 0, 1> s02;
        clc_swz<clc_float2, 4, 2, 3> s13;
    };
    CLC operator clc_float4() const;
};
struct clc_tview16 {
    union {
        float d[16];
        clc_swz<clc_float3, 16, 0, 4, 8> s012;
        clc_swz<clc_float3, 16, 1, 5, 9> s456;
        clc_swz<clc_float3, 16, 2, 6, 10> s89A;
        clc_swz<clc_float3, 16, 0, 1, 2> s048;
        clc_swz<clc_float3, 16, 4, 5, 6> s159;
        clc_swz<clc_float3, 16, 8, 9, 10> s26A;
        clc_swz<clc_float4, 16, 0, 4, 8, 12> s0123;
        clc_swz<clc_float4, 16, 1, 5, 9, 13> s4567;
        clc_swz<clc_float4, 16, 2, 6, 10, 14> s89AB;
        clc_swz<clc_float4, 16, 3, 7, 11, 15> sCDEF;
        clc_swz<clc_float4, 16, 0, 1, 2, 3> s048C;
        clc_swz<clc_float4, 16, 4, 5, 6, 7> s159D;
        clc_swz<clc_float4, 16, 8, 9, 10, 11> s26AE;
        clc_swz<clc_float4, 16, 12, 13, 14, 15> s37BF;
    };
    CLC operator clc_float16() const;
};

struct __align__(8) clc_float2 {
    union {
        struct { float x, y; };
        struct { float s0, s1; };
        float d[2];
        clc_swz<clc_float2, 2, 0, 1> xy;
        clc_swz<clc_float2, 2, 1, 0> yx;
    };
    CLC clc_float2() : x(0.f), y(0.f) {}
    CLC explicit clc_float2(float a) : x(a), y(a) {}
    CLC clc_float2(float a, float b) : x(a), y(b) {}
    CLC clc_float2& operator=(const clc_float2& o) { for (int k = 0; k < 2; k++) d[k] = o.d[k]; return *this; }
    CLC float operator[](int k) const { return d[k]; }
    CLC float& operator[](int k) { return d[k]; }
};
struct __align__(16) clc_float3 {
    union {
        struct { float x, y, z; };
        struct { float s0, s1, s2; };
        float d[4];
        clc_swz<clc_float3, 4, 0, 1, 2> xyz;
        clc_swz<clc_float2, 4, 0, 1> xy;
    };
    CLC clc_float3() : x(0.f), y(0.f), z(0.f) {}
    CLC explicit clc_float3(float a) : x(a), y(a), z(a) {}
    CLC clc_float3(float a, float b, float c) : x(a), y(b), z(c) {}
    CLC clc_float3(const clc_float2& a, float c) : x(a.x), y(a.y), z(c) {}
    CLC clc_float3& operator=(const clc_float3& o) { for (int k = 0; k < 4; k++) d[k] = o.d[k]; return *this; }
    CLC float operator[](int k) const { return d[k]; }
    CLC float& operator[](int k) { return d[k]; }
};
struct __align__(16) clc_float4 {
    union {
        struct { float x, y, z, w; };
        struct { float s0, s1, s2, s3; };
        float d[4];
        clc_swz<clc_float3, 4, 0, 1, 2> xyz;
        clc_swz<clc_float2, 4, 0, 1> xy;
        clc_swz<clc_float2, 4, 2, 3> zw;
        clc_swz<clc_float2, 4, 0, 1> s01;
        clc_swz<clc_float2, 4, 2, 3> s23;
        clc_swz<clc_float2, 4, 0, 2> s02;
        clc_swz<clc_float2, 4, 1, 3> s13;
        clc_swz<clc_float2, 4, 0, 3> s03;
        clc_tview4 s0213;
    };
    CLC clc_float4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    CLC explicit clc_float4(float a) : x(a), y(a), z(a), w(a) {}
    CLC clc_float4(float a, float b, float c, float e) : x(a), y(b), z(c), w(e) {}
    CLC clc_float4(const clc_float3& a, float e) : x(a.x), y(a.y), z(a.z), w(e) {}
    CLC clc_float4(const clc_float2& a, const clc_float2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    CLC clc_float4& operator=(const clc_float4& o) { for (int k = 0; k < 4; k++) d[k] = o.d[k]; return *this; }
    CLC float operator[](int k) const { return d[k]; }
    CLC float& operator[](int k) { return d[k]; }
};
struct __align__(64) clc_float16 {
    union {
        struct { float s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sA, sB, sC, sD, sE, sF; };
        float d[16];
        clc_swz<clc_float3, 16, 0, 1, 2> s012;
        clc_swz<clc_float3, 16, 4, 5, 6> s456;
        clc_swz<clc_float3, 16, 8, 9, 10> s89A;
        clc_swz<clc_float3, 16, 0, 4, 8> s048;
        clc_swz<clc_float3, 16, 1, 5, 9> s159;
        clc_swz<clc_float3, 16, 2, 6, 10> s26A;
        clc_swz<clc_float3, 16, 0, 5, 10> s05A;
        clc_swz<clc_float4, 16, 0, 1, 2, 3> s0123;
        clc_swz<clc_float4, 16, 4, 5, 6, 7> s4567;
        clc_swz<clc_float4, 16, 8, 9, 10, 11> s89AB;
        clc_swz<clc_float4, 16, 12, 13, 14, 15> sCDEF;
        clc_swz<clc_float4, 16, 0, 4, 8, 12> s048C;
        clc_swz<clc_float4, 16, 1, 5, 9, 13> s159D;
        clc_swz<clc_float4, 16, 2, 6, 10, 14> s26AE;
        clc_swz<clc_float4, 16, 3, 7, 11, 15> s37BF;
        clc_tview16 s048C159D26AE37BF;
    };
    CLC clc_float16() { for (int k = 0; k < 16; k++) d[k] = 0.f; }
    CLC explicit clc_float16(float a) { for (int k = 0; k < 16; k++) d[k] = a; }
    CLC clc_float16(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8,
                    float a9, float aA, float aB, float aC, float aD, float aE, float aF)
    {
        d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3; d[4] = a4; d[5] = a5; d[6] = a6; d[7] = a7;
        d[8] = a8; d[9] = a9; d[10] = aA; d[11] = aB; d[12] = aC; d[13] = aD; d[14] = aE; d[15] = aF;
    }
    CLC clc_float16& operator=(const clc_float16& o) { for (int k = 0; k < 16; k++) d[k] = o.d[k]; return *this; }
    CLC float operator[](int k) const { return d[k]; }
    CLC float& operator[](int k) { return d[k]; }
};
CLC clc_tview4::operator clc_float4() const { return clc_float4(d[0], d[2], d[1], d[3]); }
CLC clc_tview16::operator clc_float16() const
{
    return clc_float16(d[0], d[4], d[8], d[12], d[1], d[5], d[9], d[13], d[2], d[6], d[10], d[14], d[3], d[7],
                       d[11], d[15]);
}

template <class T> struct clc_n;
template <> struct clc_n<clc_float2> { static constexpr int n = 2; };
template <> struct clc_n<clc_float3> { static constexpr int n = 3; };
template <> struct clc_n<clc_float4> { static constexpr int n = 4; };
template <> struct clc_n<clc_float16> { static constexpr int n = 16; };
// non-template operators, so that swizzles convert implicitly
#define CLC_OPS(T)                                                                                              \
    CLC T operator+(const T& a, const T& b) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] + b[k]; return r; } \
    CLC T operator-(const T& a, const T& b) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] - b[k]; return r; } \
    CLC T operator*(const T& a, const T& b) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] * b[k]; return r; } \
    CLC T operator/(const T& a, const T& b) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] / b[k]; return r; } \
    CLC T operator*(const T& a, float s) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] * s; return r; }       \
    CLC T operator*(float s, const T& a) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = s * a[k]; return r; }       \
    CLC T operator/(const T& a, float s) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] / s; return r; }       \
    CLC T operator/(float s, const T& a) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = s / a[k]; return r; }       \
    CLC T operator+(const T& a, float s) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] + s; return r; }       \
    CLC T operator-(const T& a, float s) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] - s; return r; }       \
    CLC T operator-(const T& a) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = -a[k]; return r; }                   \
    CLC T& operator+=(T& a, const T& b) { for (int k = 0; k < clc_n<T>::n; k++) a[k] += b[k]; return a; }                \
    CLC T& operator-=(T& a, const T& b) { for (int k = 0; k < clc_n<T>::n; k++) a[k] -= b[k]; return a; }                \
    CLC T& operator*=(T& a, const T& b) { for (int k = 0; k < clc_n<T>::n; k++) a[k] *= b[k]; return a; }                \
    CLC T& operator*=(T& a, float s) { for (int k = 0; k < clc_n<T>::n; k++) a[k] *= s; return a; }                      \
    CLC T& operator/=(T& a, float s) { for (int k = 0; k < clc_n<T>::n; k++) a[k] /= s; return a; }                      \
    CLC T fabs(const T& a) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = fabsf(a[k]); return r; }                  \
    CLC T min(const T& a, const T& b) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = b[k] < a[k] ? b[k] : a[k]; return r; } \
    CLC T max(const T& a, const T& b) { T r; for (int k = 0; k < clc_n<T>::n; k++) r[k] = a[k] < b[k] ? b[k] : a[k]; return r; }
CLC_OPS(clc_float2)
CLC_OPS(clc_float3)
CLC_OPS(clc_float4)
CLC_OPS(clc_float16)

// products accumulated left to right in fp32 (the scripts are compiled without FMA contraction)
CLC float dot(const clc_float2& a, const clc_float2& b) { return a.x * b.x + a.y * b.y; }
CLC float dot(const clc_float3& a, const clc_float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CLC float dot(const clc_float4& a, const clc_float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
CLC float length(const clc_float2& a) { return sqrtf(dot(a, a)); }
CLC float length(const clc_float3& a) { return sqrtf(dot(a, a)); }
CLC float length(const clc_float4& a) { return sqrtf(dot(a, a)); }
CLC float distance(const clc_float2& a, const clc_float2& b) { return length(a - b); }
CLC float distance(const clc_float4& a, const clc_float4& b) { return length(a - b); }
CLC clc_float3 cross(const clc_float3& a, const clc_float3& b)
{
    return clc_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
CLC clc_float4 cross(const clc_float4& a, const clc_float4& b)
{
    return clc_float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.f);
}
CLC clc_float2 normalize(const clc_float2& a) { const float l = length(a); return clc_float2(a.x / l, a.y / l); }
CLC clc_float3 normalize(const clc_float3& a) { const float l = length(a); return clc_float3(a.x / l, a.y / l, a.z / l); }
CLC clc_float4 normalize(const clc_float4& a)
{
    const float l = length(a);
    return clc_float4(a.x / l, a.y / l, a.z / l, a.w / l);
}
// scalar built-ins with OpenCL's overloaded names (CUDA's own float overloads of sqrt, pow, fabs, ... exist)
CLC float acospi(float x) { return acosf(x) * 0.318309886183790671538f; }
CLC float sign(float x) { return (x != x) ? 0.f : (x > 0.f ? 1.f : (x < 0.f ? -1.f : x)); }
CLC float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
CLC float mix(float a, float b, float t) { return a + (b - a) * t; }
CLC float mad(float a, float b, float c) { return a * b + c; }
CLC float native_sqrt(float x) { return sqrtf(x); }
CLC float rsqrt_(float x) { return 1.f / sqrtf(x); }
CLC uint atomic_add(uint* p, uint v) { return atomicAdd(p, v); }
CLC int atomic_add(int* p, int v) { return atomicAdd(p, v); }
CLC uint atomic_inc(uint* p) { return atomicAdd(p, 1u); }
CLC int atomic_inc(int* p) { return atomicAdd(p, 1); }
CLC uint atomic_min(uint* p, uint v) { return atomicMin(p, v); }
CLC uint atomic_max(uint* p, uint v) { return atomicMax(p, v); }
#define convert_int(x) ((int)(x))
#define convert_uint(x) ((uint)(x))
#define convert_float(x) ((float)(x))
#define convert_usize(x) ((usize)(x))

// integer vectors: plain components (the scripts index them by name only)
template <class T> struct __align__(8) clc_tvec2 {
    T x, y;
    CLC clc_tvec2() : x(0), y(0) {}
    CLC clc_tvec2(T a, T b) : x(a), y(b) {}
    CLC explicit clc_tvec2(T a) : x(a), y(a) {}
};
template <class T> struct clc_tvec3 {
    T x, y, z;
    CLC clc_tvec3() : x(0), y(0), z(0) {}
    CLC clc_tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    CLC explicit clc_tvec3(T a) : x(a), y(a), z(a) {}
};
template <class T> struct __align__(16) clc_tvec4 {
    T x, y, z, w;
    CLC clc_tvec4() : x(0), y(0), z(0), w(0) {}
    CLC clc_tvec4(T a, T b, T c, T e) : x(a), y(b), z(c), w(e) {}
    CLC explicit clc_tvec4(T a) : x(a), y(a), z(a), w(a) {}
};
// (CUDA's own float2 / int4 / ... are built-in types without operators or swizzles: the names
// the scripts use are redirected to the types above)
#define float2 clc_float2
#define float3 clc_float3
#define float4 clc_float4
#define float16 clc_float16
#define int2 clc_tvec2<int>
#define int3 clc_tvec3<int>
#define int4 clc_tvec4<int>
#define uint2 clc_tvec2<uint>
#define uint3 clc_tvec3<uint>
#define uint4 clc_tvec4<uint>
#define long2 clc_tvec2<long>
#define long4 clc_tvec4<long>
#define ulong2 clc_tvec2<ulong>
#define ulong4 clc_tvec4<ulong>
// Tool.cpp:338-343 defines usize / ssize through -D (32-bit indices: State.cpp:499-502)
typedef unsigned int usize;
typedef int ssize;
#define usize2 clc_tvec2<uint>
#define usize3 clc_tvec3<uint>
#define usize4 clc_tvec4<uint>
#define ssize2 clc_tvec2<int>
#define ssize3 clc_tvec3<int>
#define ssize4 clc_tvec4<int>
)CLC"
