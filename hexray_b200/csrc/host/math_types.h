// math_types.h — host-side value types of the scene front-end.
// Same semantics as the reference's Vector / Color / Matrix / Transform (src/vector.h,
// src/color.h, src/matrix.h, src/matrix.cpp) — row-vector convention v' = v * M — written
// independently. Only what scene loading and flattening need lives here; per-ray math is
// in csrc/device/hd.h.
#pragma once
#include <cmath>
#include <cstring>

namespace hxr {
namespace host {

const double kPi = 3.141592653589793238;

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double length() const { return std::sqrt(x * x + y * y + z * z); }
    double lengthSqr() const { return x * x + y * y + z * z; }
    void normalize() { double m = 1.0 / length(); x *= m; y *= m; z *= m; }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(const Vec3& a, double m) { return Vec3(a.x * m, a.y * m, a.z * m); }
inline Vec3 operator*(double m, const Vec3& a) { return Vec3(a.x * m, a.y * m, a.z * m); }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b)
{
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double distance(const Vec3& a, const Vec3& b) { return (a - b).length(); }
// free normalize(): vectors within 1e-6 of unit length are returned untouched (src/vector.h:153-158)
inline Vec3 normalized(const Vec3& v)
{
    double len = v.length();
    if (std::fabs(len - 1.0) < 1e-6) return v;
    return v * (1 / len);
}

struct Color3 {
    float r = 0, g = 0, b = 0;
    Color3() {}
    Color3(float a, float b_, float c) : r(a), g(b_), b(c) {}
    float intensity() const { return (r + g + b) / 3; }
};

struct Mat3 {
    double m[3][3];
    Mat3() { identity(); }
    void identity()
    {
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) m[i][j] = i == j ? 1.0 : 0.0;
    }
    static Mat3 zero() { Mat3 r; std::memset(r.m, 0, sizeof r.m); return r; }
};
inline Vec3 operator*(const Vec3& v, const Mat3& a)
{
    return Vec3(v.x * a.m[0][0] + v.y * a.m[1][0] + v.z * a.m[2][0],
                v.x * a.m[0][1] + v.y * a.m[1][1] + v.z * a.m[2][1],
                v.x * a.m[0][2] + v.y * a.m[1][2] + v.z * a.m[2][2]);
}
inline Mat3 operator*(const Mat3& a, const Mat3& b)
{
    Mat3 c = Mat3::zero();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) c.m[i][j] += a.m[i][k] * b.m[k][j];
    return c;
}
inline Mat3 transposed(const Mat3& a)
{
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
inline double determinant(const Mat3& a)
{
    return a.m[0][0] * a.m[1][1] * a.m[2][2] - a.m[0][0] * a.m[1][2] * a.m[2][1] - a.m[0][1] * a.m[1][0] * a.m[2][2] +
           a.m[0][1] * a.m[1][2] * a.m[2][0] + a.m[0][2] * a.m[1][0] * a.m[2][1] - a.m[0][2] * a.m[1][1] * a.m[2][0];
}
// adjugate / determinant; a singular matrix is returned unchanged (src/matrix.cpp:107-117)
inline Mat3 inverse(const Mat3& a)
{
    double D = determinant(a);
    if (std::fabs(D) < 1e-12) return a;
    double rD = 1.0 / D;
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            // cofactor of element (j, i)
            int r0 = (j + 1) % 3, r1 = (j + 2) % 3, c0 = (i + 1) % 3, c1 = (i + 2) % 3;
            if (r0 > r1) { int t = r0; r0 = r1; r1 = t; }
            if (c0 > c1) { int t = c0; c0 = c1; c1 = t; }
            double t = a.m[r0][c0] * a.m[r1][c1] - a.m[r1][c0] * a.m[r0][c1];
            if ((i + j) % 2) t = -t;
            r.m[i][j] = rD * t;
        }
    return r;
}
inline double toRadians(double deg) { return deg / 180.0 * kPi; }
inline Mat3 rotationAroundX(double a)
{
    Mat3 r; double S = std::sin(a), C = std::cos(a);
    r.m[1][1] = C; r.m[2][1] = S; r.m[1][2] = -S; r.m[2][2] = C;
    return r;
}
inline Mat3 rotationAroundY(double a)
{
    Mat3 r; double S = std::sin(a), C = std::cos(a);
    r.m[0][0] = C; r.m[2][0] = -S; r.m[0][2] = S; r.m[2][2] = C;
    return r;
}
inline Mat3 rotationAroundZ(double a)
{
    Mat3 r; double S = std::sin(a), C = std::cos(a);
    r.m[0][0] = C; r.m[1][0] = S; r.m[0][1] = -S; r.m[1][1] = C;
    return r;
}

// scale/rotate post-multiply m in file order; translate accumulates independently
// (src/matrix.cpp:127-152)
struct Transform {
    Vec3 offset;
    Mat3 m, invM, transposedInverse;
    void refresh() { invM = inverse(m); transposedInverse = transposed(invM); }
    void scale(double x, double y, double z)
    {
        Mat3 s = Mat3::zero();
        s.m[0][0] = x; s.m[1][1] = y; s.m[2][2] = z;
        m = m * s;
        refresh();
    }
    void rotate(double yaw, double pitch, double roll)
    {
        m = m * rotationAroundZ(toRadians(roll)) * rotationAroundX(toRadians(pitch)) * rotationAroundY(toRadians(yaw));
        refresh();
    }
    void translate(const Vec3& t) { offset = offset + t; }
    Vec3 transformPoint(const Vec3& p) const { return p * m + offset; }
    Vec3 untransformPoint(const Vec3& p) const { return (p - offset) * invM; }
    Vec3 transformDir(const Vec3& d) const { return normalized(d * m); }
};

}  // namespace host
}  // namespace hxr
