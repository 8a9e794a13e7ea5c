// hd.h — shared host/device helpers: vector math that mirrors the reference's
// `Vector` (double, src/vector.h:30-165) and `Color` (float, src/color.h:62-189)
// operation by operation, so that the device intersectors and shaders reproduce the
// reference's arithmetic (same operand order, same float/double mix).
//
// The same headers compile two ways:
//   * nvcc, sm_100a  -> the product (device functions called from the kernels in render.cu)
//   * g++ -DHXR_EMU  -> tests/emu only: a host build of the SAME per-ray functions, used by the
//                       CPU test tier to check logic against the oracle without a GPU. It is
//                       never linked into libhexray_b200.so and is not a fallback.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HXR_HD __host__ __device__ __forceinline__
#define HXR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define HXR_HD inline
#define HXR_HD_NOINLINE inline
#endif

#define HXR_INF 1e99          /* reference INF, src/constants.h:29 */
#define HXR_PI 3.141592653589793238

namespace hxr {

struct d3 {
    double x, y, z;
};

HXR_HD d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
HXR_HD d3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
HXR_HD d3 operator+(const d3& a, const d3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
HXR_HD d3 operator-(const d3& a, const d3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
HXR_HD d3 operator-(const d3& a) { return mk3(-a.x, -a.y, -a.z); }
HXR_HD d3 operator*(const d3& a, double m) { return mk3(a.x * m, a.y * m, a.z * m); }
HXR_HD d3 operator*(double m, const d3& a) { return mk3(a.x * m, a.y * m, a.z * m); }
HXR_HD double dot(const d3& a, const d3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HXR_HD d3 cross(const d3& a, const d3& b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
HXR_HD double length(const d3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
HXR_HD double length_sqr(const d3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
HXR_HD double comp(const d3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// Vector / double  == multiply by the reciprocal (src/vector.h:139-143, 65-68)
HXR_HD d3 div3(const d3& a, double d) { double m = 1.0 / d; return a * m; }
// Vector::normalize() member: scale by 1/length, unconditionally (src/vector.h:69-73)
HXR_HD d3 normalize_m(const d3& a) { double m = 1.0 / length(a); return a * m; }
// free normalize(): leaves vectors whose length is within 1e-6 of 1 untouched (src/vector.h:153-158)
HXR_HD d3 normalize_f(const d3& a)
{
    double len = length(a);
    if (fabs(len - 1.0) < 1e-6) return a;
    return a * (1 / len);
}
HXR_HD double distance3(const d3& a, const d3& b)
{
    return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
// row vector times row-major 3x3 (src/matrix.h:53-60)
HXR_HD d3 mul_vm(const d3& v, const double* m)
{
    return mk3(v.x * m[0] + v.y * m[3] + v.z * m[6],
               v.x * m[1] + v.y * m[4] + v.z * m[7],
               v.x * m[2] + v.y * m[5] + v.z * m[8]);
}
HXR_HD d3 faceforward(const d3& ray, const d3& n) { return dot(ray, n) < 0 ? n : -n; }  // src/vector.h:179-182
HXR_HD d3 reflect(const d3& i, const d3& n) { return 2 * dot(-i, n) * n + i; }          // src/vector.h:184-187
HXR_HD int max_dimension(const d3& a)                                                   // src/vector.h:78-90
{
    double mv = fabs(a.x);
    int md = 0;
    if (fabs(a.y) > mv) { md = 1; mv = fabs(a.y); }
    if (fabs(a.z) > mv) md = 2;
    return md;
}

struct f3 {
    float r, g, b;
};
HXR_HD f3 mkc(float r, float g, float b) { f3 c; c.r = r; c.g = g; c.b = b; return c; }
HXR_HD f3 ldc(const float* p) { return mkc(p[0], p[1], p[2]); }
HXR_HD f3 operator+(const f3& a, const f3& b) { return mkc(a.r + b.r, a.g + b.g, a.b + b.b); }
HXR_HD f3 operator-(const f3& a, const f3& b) { return mkc(a.r - b.r, a.g - b.g, a.b - b.b); }
HXR_HD f3 operator*(const f3& a, const f3& b) { return mkc(a.r * b.r, a.g * b.g, a.b * b.b); }
HXR_HD f3 operator*(const f3& a, float m) { return mkc(a.r * m, a.g * m, a.b * m); }
HXR_HD f3 operator*(float m, const f3& a) { return mkc(a.r * m, a.g * m, a.b * m); }
// Color / float: multiply by reciprocal unless the divider is exactly 1 (src/color.h:182-187)
HXR_HD f3 divc(const f3& a, float d)
{
    if (d == 1) return a;
    float m = 1.0f / d;
    return a * m;
}
HXR_HD float intensity(const f3& a) { return (a.r + a.g + a.b) / 3; }
HXR_HD bool is_zero(const f3& a) { return a.r == 0 && a.g == 0 && a.b == 0; }

// reference Ray (src/vector.h:167-177)
struct Ray {
    d3 o, d;
    int depth;
    unsigned flags;
};
#define HXR_RF_GI_DIFFUSE 0x0002u

// reference IntersectionInfo (src/geometry.h:33-40); geom = index into the geometry table
struct Hit {
    double dist;
    d3 ip, norm, dNdx, dNdy;
    double u, v;
    int geom;
};

}  // namespace hxr
