// Ray / box and ray / primitive tests (the two-level traversal is traverse.cuh).
//
// The arithmetic follows the reference expression by expression so that hit
// ids and distances agree with the CPU build (x86-64 SSE2, no FMA): this file
// is compiled with --fmad=false, IEEE division and square root.
//   slab test            src/GoblinBVH.cpp:156-187
//   closest / any walk   src/GoblinBVH.cpp:234-280 / 189-232
//   instance entry       src/GoblinPrimitive.cpp:99-116, src/GoblinTransform.cpp:137-164
//   triangle             src/GoblinTriangle.cpp:38-80, 127-163
//   sphere               src/GoblinSphere.cpp:12-33, src/GoblinUtils.cpp:93-113
//   disk                 src/GoblinDisk.cpp:12-30
#pragma once
#include "device_scene.h"

namespace gb {

struct TraceStats {
    unsigned int nodes, prims, insts;
};

struct HitRec {
    float t;      // = shrunk ray.maxt
    float b1, b2; // triangle barycentrics
    int inst;     // instance slot (BVH leaf order), -1 = miss
    int prim;     // triangle slot within the model (BVH leaf order)
};

__device__ __forceinline__ float3 make3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return make3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return make3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float sqLen3(float3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ float len3(float3 a) { return sqrtf(sqLen3(a)); }
// Vector3::operator/(float): multiply by the reciprocal (src/GoblinVector.h:166-169)
__device__ __forceinline__ float3 div3(float3 a, float s) { float inv = 1.0f / s; return make3(a.x * inv, a.y * inv, a.z * inv); }
__device__ __forceinline__ float3 normalize3(float3 a) { return div3(a, len3(a)); }

// rows of a 3x4 matrix applied to a point / a vector (Transform::onPoint / onVector)
__device__ __forceinline__ float3 xfPoint(float4 r0, float4 r1, float4 r2, float3 p) {
    return make3(r0.x * p.x + r0.y * p.y + r0.z * p.z + r0.w,
                 r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w,
                 r2.x * p.x + r2.y * p.y + r2.z * p.z + r2.w);
}
__device__ __forceinline__ float3 xfVector(float4 r0, float4 r1, float4 r2, float3 v) {
    return make3(r0.x * v.x + r0.y * v.y + r0.z * v.z,
                 r1.x * v.x + r1.y * v.y + r1.z * v.z,
                 r2.x * v.x + r2.y * v.y + r2.z * v.z);
}

// One node = two 128-bit loads: (bmin.xyz, bmax.x) and (bmax.yz, offset, nprims | axis << 8).
__device__ __forceinline__ bool slabTest(float4 n0, float4 n1, float3 o, float3 invDir, int negX, int negY,
    int negZ, float mint, float maxt) {
    float tMin = ((negX ? n0.w : n0.x) - o.x) * invDir.x;
    float tMax = ((negX ? n0.x : n0.w) - o.x) * invDir.x;
    float tYMin = ((negY ? n1.x : n0.y) - o.y) * invDir.y;
    float tYMax = ((negY ? n0.y : n1.x) - o.y) * invDir.y;
    if (tYMax < tMin || tYMin > tMax) return false;
    if (tYMin > tMin) tMin = tYMin;
    if (tYMax < tMax) tMax = tYMax;
    float tZMin = ((negZ ? n1.y : n0.z) - o.z) * invDir.z;
    float tZMax = ((negZ ? n0.z : n1.y) - o.z) * invDir.z;
    if (tZMax < tMin || tZMin > tMax) return false;
    if (tZMin > tMin) tMin = tZMin;
    if (tZMax < tMax) tMax = tZMax;
    return (tMin < maxt) && (tMax > mint);
}

__device__ __forceinline__ bool triangleTest(float3 p0, float3 e1, float3 e2, float3 o, float3 d, float mint,
    float maxt, float* tOut, float* b1Out, float* b2Out) {
    float3 s1 = cross3(d, e2);
    float divisor = dot3(s1, e1);
    if (divisor == 0.0f) return false;
    float invDivisor = 1.0f / divisor;
    const float fEpsilon = 1e-7f;
    float3 s = o - p0;
    float b1 = dot3(s, s1) * invDivisor;
    if (b1 + fEpsilon < 0.0f || b1 - fEpsilon > 1.0f) return false;
    float3 s2 = cross3(s, e1);
    float b2 = dot3(d, s2) * invDivisor;
    if (b2 + fEpsilon < 0.0f || b1 + b2 - fEpsilon > 1.0f) return false;
    float t = dot3(e2, s2) * invDivisor;
    if (t < mint || t > maxt) return false;
    *tOut = t;
    *b1Out = b1;
    *b2Out = b2;
    return true;
}

__device__ __forceinline__ bool quadraticSolve(float A, float B, float C, float* t1, float* t2) {
    float discriminant = B * B - 4.0f * A * C;
    if (discriminant < 0.0f) return false;
    float rootDiscrim = sqrtf(discriminant);
    float q = B < 0 ? -0.5f * (B - rootDiscrim) : -0.5f * (B + rootDiscrim);
    float a = q / A, b = C / q;
    if (a > b) { float tmp = a; a = b; b = tmp; }
    *t1 = a;
    *t2 = b;
    return true;
}

__device__ __forceinline__ bool sphereTest(float radius, float3 o, float3 d, float mint, float maxt, float* tOut) {
    float A = sqLen3(d);
    float B = 2.0f * dot3(d, o);
    float C = sqLen3(o) - radius * radius;
    float tNear, tFar;
    if (!quadraticSolve(A, B, C, &tNear, &tFar)) return false;
    if (tNear > maxt || tFar < mint) return false;
    float tHit = tNear;
    if (tHit < mint) {
        tHit = tFar;
        if (tHit > maxt) return false;
    }
    *tOut = tHit;
    return true;
}

__device__ __forceinline__ bool diskTest(float radius, float3 o, float3 d, float mint, float maxt, float* tOut) {
    if (fabsf(d.z) < 1e-7f) return false;
    float t = -o.z / d.z;
    float3 p = o + t * d;
    if (t < mint || t > maxt) return false;
    float squareR = p.x * p.x + p.y * p.y;
    if (squareR > radius * radius) return false;
    *tOut = t;
    return true;
}

} // namespace gb
