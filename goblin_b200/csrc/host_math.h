// Host-side float math for scene preparation.
//
// Instance matrices, world bounds and BVH boxes must come out bit-identical to
// the reference's, so every function here keeps the reference's operation
// order (no FMA on the host: plain x86-64 SSE2 build).  Citations are to
// /root/reference/src.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>

namespace gb {

constexpr float kPi = 3.14159265358979323f;        // GoblinUtils.h:43-46
constexpr float kTwoPi = 6.28318530718f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
constexpr float kInf = std::numeric_limits<float>::infinity();

struct Vec2 { float x = 0, y = 0; };
struct Vec4 { float x = 0, y = 0, z = 0, w = 0; };

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return (&x)[i]; }
    float& operator[](int i) { return (&x)[i]; }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(const Vec3& a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(const Vec3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, const Vec3& a) { return a * s; }
// GoblinVector.h:166-169: division multiplies by the reciprocal
inline Vec3 operator/(const Vec3& a, float s) { float inv = 1.0f / s; return {a.x * inv, a.y * inv, a.z * inv}; }
inline float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float squaredLength(const Vec3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float length(const Vec3& a) { return std::sqrt(squaredLength(a)); }
inline Vec3 normalize(const Vec3& a) { return a / length(a); }
inline float radians(float deg) { return kPi * (deg / 180.0f); } // GoblinUtils.h:134-136

struct BBox { // GoblinBBox.h:12-20, GoblinBBox.cpp:6-23
    Vec3 pMin{kInf, kInf, kInf};
    Vec3 pMax{-kInf, -kInf, -kInf};
    void expand(const Vec3& p) {
        pMin = Vec3(std::min(pMin.x, p.x), std::min(pMin.y, p.y), std::min(pMin.z, p.z));
        pMax = Vec3(std::max(pMax.x, p.x), std::max(pMax.y, p.y), std::max(pMax.z, p.z));
    }
    void expand(const BBox& b) {
        pMin = Vec3(std::min(pMin.x, b.pMin.x), std::min(pMin.y, b.pMin.y), std::min(pMin.z, b.pMin.z));
        pMax = Vec3(std::max(pMax.x, b.pMax.x), std::max(pMax.y, b.pMax.y), std::max(pMax.z, b.pMax.z));
    }
    int longestAxis() const { // GoblinBBox.cpp:79-88
        Vec3 d = pMax - pMin;
        if (d.x > d.y && d.x > d.z) return 0;
        if (d.y > d.z) return 1;
        return 2;
    }
};
// BBox(p1, p2) constructor: component-wise min / max (GoblinBBox.h:22-25)
inline BBox makeBBox(const Vec3& a, const Vec3& b) {
    BBox r;
    r.pMin = Vec3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z));
    r.pMax = Vec3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z));
    return r;
}

struct Mat4 {
    float m[4][4];
    static Mat4 identity() {
        Mat4 r{};
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0f : 0.0f;
        return r;
    }
};
inline Mat4 mul(const Mat4& a, const Mat4& b) { // GoblinMatrix.cpp:305-314
    Mat4 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
        r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
// inverse(Matrix4*, const Matrix4&), GoblinMatrix.cpp:419-486: cofactor expansion
bool inverse(Mat4* inv, const Mat4& m);

struct Quat { // w + v, GoblinQuaternion.h
    float w = 1, x = 0, y = 0, z = 0;
    Quat() = default;
    Quat(float w_, float x_, float y_, float z_) : w(w_), x(x_), y(y_), z(z_) {}
};
Quat quatFromAxisAngle(const Vec3& axis, float angle);          // GoblinQuaternion.cpp:9-15
Quat quatFromMatrix3(const float R[3][3]);                      // GoblinQuaternion.cpp:21-53
Quat quatMul(const Quat& a, const Quat& b);                     // GoblinQuaternion.h operator*
Quat quatNormalize(const Quat& q);                              // GoblinQuaternion.cpp:94-100
Mat4 quatToMatrix(const Quat& q);                               // GoblinQuaternion.cpp:55-74
Vec3 quatRotate(const Quat& q, const Vec3& p);                  // GoblinQuaternion.cpp:87-93
Quat eulerToQuat(const Vec3& xyzDegrees, const std::string& order); // GoblinQuaternion.cpp:103-151
void coordinateAxises(const Vec3& a1, Vec3* a2, Vec3* a3);      // GoblinUtils.cpp:58-69

struct Transform { // GoblinTransform.cpp:14-21,182-193
    Vec3 position;
    Quat orientation;
    Vec3 scale{1, 1, 1};
    Mat4 matrix = Mat4::identity();
    Mat4 inv = Mat4::identity();
    void update();
    Vec3 onPoint(const Vec3& p) const;
    Vec3 onVector(const Vec3& v) const;
    Vec3 invertPoint(const Vec3& p) const;
    BBox onBBox(const BBox& b) const;
};

} // namespace gb
