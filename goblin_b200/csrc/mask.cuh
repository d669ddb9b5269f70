// Mask materials on the device (src/GoblinMaterial.cpp:747-811, src/GoblinPathtracer.cpp:5-48).
//
// A scene with a Mask material makes the path tracer trace with filters: shadow and MIS rays
// stop at OPAQUE primitives only (isOpaque), and the not-opaque ones along the segment attenuate
// it (evalAttenuation: repeated closest-hit over the not-opaque primitives, each contributing
// (1 - alpha) * transparent colour).  Those filtered walks are rare-path work, so they do not go
// through the persistent traversal kernel: each thread walks the reference's own 32-byte nodes
// with the reference's loop (src/GoblinBVH.cpp:189-280: near child first, far child on a todo
// stack, leaves tested as reached), which keeps hits and tie-breaking identical to the CPU build
// without touching the hot kernels.  The filter is applied per instance: the material belongs to
// the model, so this gives the hits of the reference's per-primitive filter.
#pragma once
#include "texture.cuh"
#include "traverse.cuh"

namespace gb {

constexpr int kSimpleStack = 64; // the reference's todo[64] per level
enum { FILTER_NONE = 0, FILTER_OPAQUE = 1, FILTER_NOT_OPAQUE = 2 };

__device__ __forceinline__ bool instanceFilteredOut(const DeviceScene& sc, unsigned int slot, int filter) {
    if (filter == FILTER_NONE) return false;
    const int mat = __ldg(sc.instShade + slot).z;
    const bool opaque = __ldg(sc.matMask + 2 * (size_t)mat).x == 0;
    return (filter == FILTER_OPAQUE) != opaque;
}

// BVH::intersect / BVH::occluded over one level.  leaf(slot) tests primitive `slot` and returns true
// to stop (any-hit).  maxt is taken BY REFERENCE on purpose: the closest-hit caller's leaf functor shrinks
// the same variable when it accepts a hit, and the box tests of the remaining walk must see the shrunk
// value, like ray.maxt in the reference.
template <typename Leaf>
__device__ __forceinline__ bool simpleWalk(const float4* __restrict__ nodes, unsigned int nNodes, float3 o, float3 d,
    float mint, const float& maxt, Leaf leaf) {
    if (nNodes == 0) return false;
    const float3 inv = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const unsigned int neg = signBits(d);
    unsigned int todo[kSimpleStack];
    int top = 0;
    unsigned int node = 0;
    for (;;) {
        const float4 n0 = __ldg(nodes + 2 * (size_t)node), n1 = __ldg(nodes + 2 * (size_t)node + 1);
        bool descend = false;
        if (rootTest(n0, n1, o, inv, neg, mint, maxt)) {
            const unsigned int word = __float_as_uint(n1.w), count = word & 0xffu;
            if (count > 0) {
                const unsigned int first = __float_as_uint(n1.z);
                for (unsigned int k = 0; k < count; ++k) {
                    if (leaf(first + k)) return true;
                }
            } else {
                const unsigned int axis = (word >> 8) & 3u, right = __float_as_uint(n1.z);
                if ((neg >> axis) & 1u) {
                    if (top < kSimpleStack) todo[top++] = node + 1;
                    node = right;
                } else {
                    if (top < kSimpleStack) todo[top++] = right;
                    node = node + 1;
                }
                descend = true;
            }
        }
        if (!descend) {
            if (top == 0) break;
            node = todo[--top];
        }
    }
    return false;
}

// Scene::intersect(ray, ..., filter): closest hit among the instances the filter lets through.
__device__ __noinline__ bool simpleClosest(const DeviceScene& sc, float3 o, float3 d, float mint, float maxtIn, int filter,
    HitRec* hit) {
    float maxt = maxtIn;
    bool found = false;
    hit->inst = -1; hit->prim = 0; hit->t = maxtIn; hit->b1 = hit->b2 = 0.0f;
    simpleWalk(sc.topNodes, sc.nTopNodes, o, d, mint, maxt, [&](unsigned int slot) {
        if (instanceFilteredOut(sc, slot, filter)) return false;
        const float4* m = sc.instToObject + 3 * (size_t)slot;
        const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
        const int4 info = __ldg(sc.instInfo + slot);
        const float3 oo = xfPoint(r0, r1, r2, o), od = xfVector(r0, r1, r2, d);
        if (info.x == GB_GEOM_MESH) {
            const int4 info2 = __ldg(sc.instInfo2 + slot);
            simpleWalk(sc.modelNodes + 2 * (size_t)(unsigned int)info.y, (unsigned int)info2.z, oo, od, mint, maxt,
                [&](unsigned int ts) {
                    const float4* tr = sc.triRec + kTriRecVec4 * ((size_t)(unsigned int)info.z + ts);
                    const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
                    float t, b1, b2;
                    if (triangleTest(make3(a.x, a.y, a.z), make3(a.w, b.x, b.y), make3(b.z, b.w, c.x), oo, od, mint, maxt, &t,
                            &b1, &b2)) {
                        found = true;
                        maxt = t;
                        hit->t = t; hit->b1 = b1; hit->b2 = b2; hit->inst = (int)slot; hit->prim = (int)ts;
                    }
                    return false;
                });
            return false;
        }
        float t;
        const float radius = __int_as_float(info.w);
        if (info.x == GB_GEOM_SPHERE ? sphereTest(radius, oo, od, mint, maxt, &t) : diskTest(radius, oo, od, mint, maxt, &t)) {
            found = true;
            maxt = t;
            hit->t = t; hit->b1 = hit->b2 = 0.0f; hit->inst = (int)slot; hit->prim = 0;
        }
        return false;
    });
    return found;
}

// Scene::occluded(ray, filter)
__device__ __noinline__ bool simpleAny(const DeviceScene& sc, float3 o, float3 d, float mint, float maxt, int filter) {
    return simpleWalk(sc.topNodes, sc.nTopNodes, o, d, mint, maxt, [&](unsigned int slot) {
        if (instanceFilteredOut(sc, slot, filter)) return false;
        const float4* m = sc.instToObject + 3 * (size_t)slot;
        const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
        const int4 info = __ldg(sc.instInfo + slot);
        const float3 oo = xfPoint(r0, r1, r2, o), od = xfVector(r0, r1, r2, d);
        if (info.x == GB_GEOM_MESH) {
            const int4 info2 = __ldg(sc.instInfo2 + slot);
            return simpleWalk(sc.modelNodes + 2 * (size_t)(unsigned int)info.y, (unsigned int)info2.z, oo, od, mint, maxt,
                [&](unsigned int ts) {
                    const float4* tr = sc.triRec + kTriRecVec4 * ((size_t)(unsigned int)info.z + ts);
                    const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
                    float t, b1, b2;
                    return triangleTest(make3(a.x, a.y, a.z), make3(a.w, b.x, b.y), make3(b.z, b.w, c.x), oo, od, mint, maxt, &t,
                        &b1, &b2);
                });
        }
        float t;
        const float radius = __int_as_float(info.w);
        return info.x == GB_GEOM_SPHERE ? sphereTest(radius, oo, od, mint, maxt, &t) : diskTest(radius, oo, od, mint, maxt, &t);
    });
}

// The mask of material `material` at a hit: alpha and transparent colour, constants or texture programs
// (no differentials: evalAttenuation and the bounce rays work on fresh Fragments).
struct MaskEval { float alpha; float3 tc; };
__device__ __noinline__ MaskEval maskAt(const DeviceScene& sc, int material, const HitRec& hit, float3 o, float3 d,
    const Frag& fr) {
    const int4 info = __ldg(sc.matMask + 2 * (size_t)material);         // flag, alpha program, colour program, -
    const float4 val = __ldg(reinterpret_cast<const float4*>(sc.matMask + 2 * (size_t)material + 1)); // alpha, tc rgb
    MaskEval e;
    e.alpha = val.x;
    e.tc = make3(val.y, val.z, val.w);
    if (info.y | info.z) {
        TexFrag tf;
        texFragment(sc, hit, o, d, fr, &tf);
        if (info.y) e.alpha = evalTexture(sc, (unsigned int)info.y, tf).x;
        if (info.z) e.tc = evalTexture(sc, (unsigned int)info.z, tf);
    }
    return e;
}

// PathTracer::evalAttenuation (src/GoblinPathtracer.cpp:21-48)
__device__ __noinline__ float3 evalAttenuation(const DeviceScene& sc, float3 o, float3 d, float mint, float maxt) {
    float3 thr = make3(1.0f, 1.0f, 1.0f);
    float curMint = mint;
    for (int guard = 0; guard < 4096; ++guard) {
        HitRec h;
        if (!simpleClosest(sc, o, d, curMint, maxt, FILTER_NOT_OPAQUE, &h)) break;
        const Frag fr = buildFragment(sc, h, o, d);
        const MaskEval e = maskAt(sc, fr.material, h, o, d, fr);
        thr = mul3(thr, (1.0f - e.alpha) * e.tc); // sampleBSDF(..., BSDFnullptr)
        if (thr.x == 0.0f && thr.y == 0.0f && thr.z == 0.0f) break;
        curMint = h.t + 1e-3f * h.t; // currentRay.mint = currentRay.maxt + epsilon
    }
    return thr;
}

} // namespace gb
