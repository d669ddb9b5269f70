#include "host_math.h"

#include <iostream>

namespace gb {

// The reference calls unqualified sin / cos / tan on float arguments from
// inside namespace Goblin, which binds to the C library's double versions;
// the result is then narrowed to float.  Keep that.
static inline float sinD(float a) { return (float)::sin((double)a); }
static inline float cosD(float a) { return (float)::cos((double)a); }

bool inverse(Mat4* out, const Mat4& in) {
    const float (*m)[4] = in.m;
    float m00 = m[0][0], m01 = m[0][1], m02 = m[0][2], m03 = m[0][3];
    float m10 = m[1][0], m11 = m[1][1], m12 = m[1][2], m13 = m[1][3];
    float m20 = m[2][0], m21 = m[2][1], m22 = m[2][2], m23 = m[2][3];
    float m30 = m[3][0], m31 = m[3][1], m32 = m[3][2], m33 = m[3][3];
    float (*r)[4] = out->m;

    // 2x2 minors of rows 2,3
    float a2323 = m22 * m33 - m23 * m32;
    float a1323 = m21 * m33 - m23 * m31;
    float a1223 = m21 * m32 - m22 * m31;
    float a0323 = m20 * m33 - m23 * m30;
    float a0223 = m20 * m32 - m22 * m30;
    float a0123 = m20 * m31 - m21 * m30;

    r[0][0] = +(m11 * a2323 - m12 * a1323 + m13 * a1223);
    r[1][0] = -(m10 * a2323 - m12 * a0323 + m13 * a0223);
    r[2][0] = +(m10 * a1323 - m11 * a0323 + m13 * a0123);
    r[3][0] = -(m10 * a1223 - m11 * a0223 + m12 * a0123);

    float det = m00 * r[0][0] + m01 * r[1][0] + m02 * r[2][0] + m03 * r[3][0];
    if (std::fabs(det) < 1e-6f) { // MATRIX_EPSILON
        return false;
    }
    float invDet = 1.0f / det;

    r[0][1] = -(m01 * a2323 - m02 * a1323 + m03 * a1223);
    r[1][1] = +(m00 * a2323 - m02 * a0323 + m03 * a0223);
    r[2][1] = -(m00 * a1323 - m01 * a0323 + m03 * a0123);
    r[3][1] = +(m00 * a1223 - m01 * a0223 + m02 * a0123);

    // minors of rows 1,3
    float b2313 = m12 * m33 - m13 * m32;
    float b1313 = m11 * m33 - m13 * m31;
    float b1213 = m11 * m32 - m12 * m31;
    float b0313 = m10 * m33 - m13 * m30;
    float b0213 = m10 * m32 - m12 * m30;
    float b0113 = m10 * m31 - m11 * m30;

    r[0][2] = +(m01 * b2313 - m02 * b1313 + m03 * b1213);
    r[1][2] = -(m00 * b2313 - m02 * b0313 + m03 * b0213);
    r[2][2] = +(m00 * b1313 - m01 * b0313 + m03 * b0113);
    r[3][2] = -(m00 * b1213 - m01 * b0213 + m02 * b0113);

    // minors of rows 1,2
    float c2312 = m12 * m23 - m13 * m22;
    float c1312 = m11 * m23 - m13 * m21;
    float c1212 = m11 * m22 - m12 * m21;
    float c0312 = m10 * m23 - m13 * m20;
    float c0212 = m10 * m22 - m12 * m20;
    float c0112 = m10 * m21 - m11 * m20;

    r[0][3] = -(m01 * c2312 - m02 * c1312 + m03 * c1212);
    r[1][3] = +(m00 * c2312 - m02 * c0312 + m03 * c0212);
    r[2][3] = -(m00 * c1312 - m01 * c0312 + m03 * c0112);
    r[3][3] = +(m00 * c1212 - m01 * c0212 + m02 * c0112);

    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[i][j] *= invDet;
    return true;
}

Quat quatFromAxisAngle(const Vec3& axis, float angle) {
    float t = angle * 0.5f;
    Vec3 u = normalize(axis);
    float s = sinD(t);
    Quat q;
    q.w = cosD(t);
    q.x = u.x * s; q.y = u.y * s; q.z = u.z * s;
    return q;
}

Quat quatFromMatrix3(const float R[3][3]) {
    float q[4];
    float trace = R[0][0] + R[1][1] + R[2][2];
    if (trace > 0.0f) {
        float s = std::sqrt(trace + 1.0f);
        q[3] = s * 0.5f;
        float t = 0.5f / s;
        q[0] = (R[2][1] - R[1][2]) * t;
        q[1] = (R[0][2] - R[2][0]) * t;
        q[2] = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        static const int next[3] = {1, 2, 0};
        int j = next[i], k = next[j];
        float s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0f);
        q[i] = s * 0.5f;
        float t = s != 0.0f ? 0.5f / s : s;
        q[3] = (R[k][j] - R[j][k]) * t;
        q[j] = (R[j][i] + R[i][j]) * t;
        q[k] = (R[k][i] + R[i][k]) * t;
    }
    return Quat(q[3], q[0], q[1], q[2]);
}

Quat quatMul(const Quat& a, const Quat& b) {
    Vec3 av(a.x, a.y, a.z), bv(b.x, b.y, b.z);
    float w = a.w * b.w - dot(av, bv);
    Vec3 v = a.w * bv + b.w * av + cross(av, bv);
    return Quat(w, v.x, v.y, v.z);
}

Quat quatNormalize(const Quat& q) {
    float inv = 1.0f / sqrt(q.w * q.w + (q.x * q.x + q.y * q.y + q.z * q.z));
    return Quat(q.w * inv, q.x * inv, q.y * inv, q.z * inv);
}

Mat4 quatToMatrix(const Quat& q) {
    float x2 = 2.0f * q.x, y2 = 2.0f * q.y, z2 = 2.0f * q.z;
    float xx2 = x2 * q.x, xy2 = x2 * q.y, xz2 = x2 * q.z, xw2 = x2 * q.w;
    float yy2 = y2 * q.y, yz2 = y2 * q.z, yw2 = y2 * q.w;
    float zz2 = z2 * q.z, zw2 = z2 * q.w;
    Mat4 r = Mat4::identity();
    r.m[0][0] = 1 - yy2 - zz2; r.m[0][1] = xy2 - zw2;     r.m[0][2] = xz2 + yw2;
    r.m[1][0] = xy2 + zw2;     r.m[1][1] = 1 - xx2 - zz2; r.m[1][2] = yz2 - xw2;
    r.m[2][0] = xz2 - yw2;     r.m[2][1] = yz2 + xw2;     r.m[2][2] = 1 - xx2 - yy2;
    return r;
}

Vec3 quatRotate(const Quat& q, const Vec3& p) {
    Vec3 v(q.x, q.y, q.z);
    Vec3 uv = cross(v, p);
    Vec3 uuv = cross(v, uv);
    uv = uv * (2.0f * q.w);
    uuv = uuv * 2.0f;
    return p + uv + uuv;
}

Quat eulerToQuat(const Vec3& a, const std::string& order) {
    Quat qx = quatFromAxisAngle(Vec3(1, 0, 0), radians(a.x));
    Quat qy = quatFromAxisAngle(Vec3(0, 1, 0), radians(a.y));
    Quat qz = quatFromAxisAngle(Vec3(0, 0, 1), radians(a.z));
    if (order == "xyz") return quatMul(quatMul(qz, qy), qx);
    if (order == "xzy") return quatMul(quatMul(qy, qz), qx);
    if (order == "yxz") return quatMul(quatMul(qz, qx), qy);
    if (order == "yzx") return quatMul(quatMul(qx, qz), qy);
    if (order == "zxy") return quatMul(quatMul(qy, qx), qz);
    if (order == "zyx") return quatMul(quatMul(qx, qy), qz);
    std::cerr << "unrecognized rotation order " << order << ", fall back to XYZ" << std::endl;
    return quatMul(quatMul(qz, qy), qx);
}

void coordinateAxises(const Vec3& a1, Vec3* a2, Vec3* a3) {
    if (std::fabs(a1.x) > std::fabs(a1.y)) {
        float invLen = 1.0f / std::sqrt(a1.x * a1.x + a1.z * a1.z);
        *a2 = Vec3(-a1.z * invLen, 0.0f, a1.x * invLen);
    } else {
        float invLen = 1.0f / std::sqrt(a1.y * a1.y + a1.z * a1.z);
        *a2 = Vec3(0.0f, -a1.z * invLen, a1.y * invLen);
    }
    *a3 = cross(a1, *a2);
}

void Transform::update() {
    Mat4 S = Mat4::identity();
    S.m[0][0] = scale.x; S.m[1][1] = scale.y; S.m[2][2] = scale.z;
    Mat4 R = quatToMatrix(orientation);
    matrix = mul(R, S);
    matrix.m[0][3] = position.x;
    matrix.m[1][3] = position.y;
    matrix.m[2][3] = position.z;
    inverse(&inv, matrix);
}

Vec3 Transform::onPoint(const Vec3& p) const {
    const float (*M)[4] = matrix.m;
    return Vec3(M[0][0] * p.x + M[0][1] * p.y + M[0][2] * p.z + M[0][3],
                M[1][0] * p.x + M[1][1] * p.y + M[1][2] * p.z + M[1][3],
                M[2][0] * p.x + M[2][1] * p.y + M[2][2] * p.z + M[2][3]);
}

Vec3 Transform::onVector(const Vec3& v) const {
    const float (*M)[4] = matrix.m;
    return Vec3(M[0][0] * v.x + M[0][1] * v.y + M[0][2] * v.z,
                M[1][0] * v.x + M[1][1] * v.y + M[1][2] * v.z,
                M[2][0] * v.x + M[2][1] * v.y + M[2][2] * v.z);
}

Vec3 Transform::invertPoint(const Vec3& p) const {
    const float (*M)[4] = inv.m;
    return Vec3(M[0][0] * p.x + M[0][1] * p.y + M[0][2] * p.z + M[0][3],
                M[1][0] * p.x + M[1][1] * p.y + M[1][2] * p.z + M[1][3],
                M[2][0] * p.x + M[2][1] * p.y + M[2][2] * p.z + M[2][3]);
}

BBox Transform::onBBox(const BBox& b) const { // GoblinTransform.cpp:125-135
    BBox rv;
    Vec3 first = onPoint(b.pMin);
    rv.pMin = first; rv.pMax = first;
    rv.expand(onPoint(Vec3(b.pMax.x, b.pMin.y, b.pMin.z)));
    rv.expand(onPoint(Vec3(b.pMin.x, b.pMax.y, b.pMin.z)));
    rv.expand(onPoint(Vec3(b.pMin.x, b.pMin.y, b.pMax.z)));
    rv.expand(onPoint(Vec3(b.pMax.x, b.pMax.y, b.pMin.z)));
    rv.expand(onPoint(Vec3(b.pMax.x, b.pMin.y, b.pMax.z)));
    rv.expand(onPoint(Vec3(b.pMin.x, b.pMax.y, b.pMax.z)));
    rv.expand(onPoint(b.pMax));
    return rv;
}

} // namespace gb
