// Ray / box, ray / primitive tests and the two-level traversal.
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

// Per-thread traversal stack in shared memory, one column per thread so that a
// warp's pushes / pops never bank-conflict: entry k of thread t lives at
// stack[k * blockDim.x + t].
struct SmemStack {
    unsigned int* base; // &stack[threadIdx.x]
    unsigned int stride;
    __device__ __forceinline__ void put(int k, unsigned int v) { base[k * stride] = v; }
    __device__ __forceinline__ unsigned int get(int k) const { return base[k * stride]; }
};

// BVH::intersect / BVH::occluded over one model's triangles, in object space.
// Returns true as soon as something is hit when ANY; otherwise shrinks *maxt
// and records the last accepted triangle (t <= maxt accepts ties, so the later
// primitive in traversal order wins, as in the reference).
template <bool ANY, bool STATS>
__device__ __forceinline__ bool walkModel(const DeviceScene& sc, unsigned int nodeBase, unsigned int triBase,
    float3 o, float3 d, float mint, float* maxt, HitRec* hit, int instSlot, SmemStack st, int sp0,
    TraceStats* stats) {
    float3 invDir = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int negX = d.x < 0.0f, negY = d.y < 0.0f, negZ = d.z < 0.0f;
    unsigned int node = 0;
    int sp = sp0;
    bool found = false;
    const float4* nodes = sc.modelNodes + 2 * (size_t)nodeBase;
    while (true) {
        float4 n0 = __ldg(nodes + 2 * node);
        float4 n1 = __ldg(nodes + 2 * node + 1);
        if (STATS) stats->nodes++;
        bool descend = false;
        if (slabTest(n0, n1, o, invDir, negX, negY, negZ, mint, *maxt)) {
            unsigned int meta = __float_as_uint(n1.w);
            unsigned int nprims = meta & 0xffu;
            unsigned int offset = __float_as_uint(n1.z);
            if (nprims > 0) {
                for (unsigned int i = 0; i < nprims; ++i) {
                    const float4* tr = sc.triRec + 3 * (size_t)(triBase + offset + i);
                    float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
                    if (STATS) stats->prims++;
                    float t, b1, b2;
                    if (triangleTest(make3(a.x, a.y, a.z), make3(a.w, b.x, b.y), make3(b.z, b.w, c.x), o, d,
                            mint, *maxt, &t, &b1, &b2)) {
                        if (ANY) return true;
                        *maxt = t;
                        hit->t = t; hit->b1 = b1; hit->b2 = b2;
                        hit->inst = instSlot;
                        hit->prim = (int)(offset + i);
                        found = true;
                    }
                }
            } else {
                unsigned int axis = (meta >> 8) & 0xffu;
                int neg = axis == 0 ? negX : (axis == 1 ? negY : negZ);
                if (neg) { st.put(sp++, node + 1); node = offset; }
                else { st.put(sp++, offset); node = node + 1; }
                descend = true;
            }
        }
        if (!descend) {
            if (sp == sp0) break;
            node = st.get(--sp);
        }
    }
    return found;
}

// Scene::intersect (ANY = false) / Scene::occluded (ANY = true).
template <bool ANY, bool STATS>
__device__ __forceinline__ bool traceScene(const DeviceScene& sc, float3 o, float3 d, float mint, float maxt,
    HitRec* hit, SmemStack st, TraceStats* stats) {
    hit->inst = -1;
    hit->prim = 0;
    hit->t = maxt;
    hit->b1 = hit->b2 = 0.0f;
    if (sc.nTopNodes == 0) return false;
    float3 invDir = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int negX = d.x < 0.0f, negY = d.y < 0.0f, negZ = d.z < 0.0f;
    unsigned int node = 0;
    int sp = 0;
    bool found = false;
    while (true) {
        float4 n0 = __ldg(sc.topNodes + 2 * node);
        float4 n1 = __ldg(sc.topNodes + 2 * node + 1);
        if (STATS) stats->nodes++;
        bool descend = false;
        if (slabTest(n0, n1, o, invDir, negX, negY, negZ, mint, maxt)) {
            unsigned int meta = __float_as_uint(n1.w);
            unsigned int nprims = meta & 0xffu;
            unsigned int offset = __float_as_uint(n1.z);
            if (nprims > 0) {
                for (unsigned int i = 0; i < nprims; ++i) {
                    unsigned int slot = offset + i;
                    const float4* m = sc.instToObject + 3 * (size_t)slot;
                    float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
                    int4 info = __ldg(sc.instInfo + slot);
                    if (STATS) stats->insts++;
                    // Transform::invertRay: direction is not renormalised, t is shared
                    float3 oo = xfPoint(r0, r1, r2, o);
                    float3 od = xfVector(r0, r1, r2, d);
                    if (info.x == GB_GEOM_MESH) {
                        if (__ldg(sc.instNodeCount + slot) == 0) continue;
                        if (walkModel<ANY, STATS>(sc, (unsigned int)info.y, (unsigned int)info.z, oo, od, mint,
                                &maxt, hit, (int)slot, st, sp, stats)) {
                            if (ANY) return true;
                            found = true;
                        }
                    } else {
                        float t;
                        float radius = __int_as_float(info.w);
                        if (STATS) stats->prims++;
                        bool h = info.x == GB_GEOM_SPHERE ? sphereTest(radius, oo, od, mint, maxt, &t)
                                                          : diskTest(radius, oo, od, mint, maxt, &t);
                        if (h) {
                            if (ANY) return true;
                            maxt = t;
                            hit->t = t; hit->b1 = 0.0f; hit->b2 = 0.0f;
                            hit->inst = (int)slot;
                            hit->prim = 0;
                            found = true;
                        }
                    }
                }
            } else {
                unsigned int axis = (meta >> 8) & 0xffu;
                int neg = axis == 0 ? negX : (axis == 1 ? negY : negZ);
                if (neg) { st.put(sp++, node + 1); node = offset; }
                else { st.put(sp++, offset); node = node + 1; }
                descend = true;
            }
        }
        if (!descend) {
            if (sp == 0) break;
            node = st.get(--sp);
        }
    }
    return found;
}

} // namespace gb
