// 4-wide nodes: two binary levels of the reference's tree collapsed into one 128-byte record.
//
// The tree is still the reference's (src/GoblinBVH.cpp:34-151, bit-exact 32-byte nodes) and it is
// still walked in the reference's order (src/GoblinBVH.cpp:234-280: the child on the side the ray
// comes from first, by dirIsNeg[axis]); only the unit of work changes.  An interior node R at an
// even depth becomes a "wide root": its record holds the boxes of R's grandchildren (or of a child
// that is itself a leaf) in the canonical order [LL, LR, RL, RR], each as the reference's own
// 32-byte node layout (bmin[3], bmax[3]) with the last two words replaced by a child reference
// and, in slot 0, the three split axes (R's, its left child's, its right child's).  One traversal
// step fetches the record with 8 x 128-bit loads and runs FOUR independent slab tests -- half the
// dependent fetch round trips per ray of the pair-node walk and four tests' worth of instruction
// level parallelism per step -- then visits the boxes that were hit in exactly the order the
// reference would: near side of R first, and within a side the near grandchild first.
//
// Skipping the intermediate child's own box test cannot change a result: a grandchild's box lies
// inside its parent's, float subtraction and multiplication are monotone, so whenever a
// grandchild's slab test passes the parent's would have passed too (the one exception is the
// reference's NaN case -- a zero direction component with the origin exactly on a plane of the
// PARENT's box -- where the reference prunes a subtree this walk still visits; such a ray may then
// report a hit the reference misses).  The maxt-dependent part of a box test (tMin < maxt) is
// re-evaluated when an entry is popped, against the maxt of that moment, as the reference does.
//
// Everything here is plain arithmetic on plain structs, usable from device code and from the host
// (tests/native/ emulates the walk on the CPU with these same functions to check the logic
// without a GPU; the product itself has no CPU path).
#pragma once
#include <math.h>
#include <stdint.h>

#include "goblin_b200.h"

// GB_SLAB_MINMAX=0: every box test runs the ordered form (A/B builds)
#ifndef GB_SLAB_MINMAX
#define GB_SLAB_MINMAX 1
#endif

#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

namespace gb {

constexpr uint32_t REF_LEAF = 0x80000000u;  // child is a leaf
constexpr uint32_t REF_MULTI = 0x40000000u; // leaf with nprims != 1: index = original node
constexpr uint32_t REF_INDEX = 0x3FFFFFFFu;
constexpr uint32_t REF_NONE = 0xFFFFFFFFu;  // nothing left at this level
constexpr uint32_t REF_POP = 0xFFFFFFFEu;   // take the next entry off the stack
constexpr uint32_t WIDE_NOT_ROOT = 0xFFFFFFFFu; // wideIndex[] of a node that is not a wide root

struct WideChild { // 32 bytes, the reference's node layout with the tail reused
    float lo[3];
    float hi[3];
    uint32_t ref;  // pair-style child reference: wide index, REF_LEAF | slot, REF_LEAF | REF_MULTI | node, REF_POP = empty
    uint32_t meta; // slot 0: axis(R) | axis(left) << 2 | axis(right) << 4
};
struct WideNode {
    WideChild c[4]; // LL, LR, RL, RR
};
static_assert(sizeof(WideNode) == 128, "wide node size");

// Stack entries a walk of a tree of depth `depth` (edges on the longest root-to-leaf path) needs:
// every wide level on the path leaves at most three pending siblings.
GB_HD int wideStackEntries(int depth) { return 3 * ((depth + 1) / 2); }

GB_HD uint32_t wideRefOf(const gb_bvh_node* nodes, const uint32_t* wideIndex, uint32_t node) {
    const gb_bvh_node& nd = nodes[node];
    if (nd.nprims == 0) return wideIndex[node];
    if (nd.nprims == 1) return REF_LEAF | nd.offset;
    return REF_LEAF | REF_MULTI | node;
}

// The record of wide root `i` (an interior node at an even depth).
GB_HD void deriveWideNode(const gb_bvh_node* nodes, const uint32_t* wideIndex, uint32_t i, WideNode* out) {
    const float inf = __builtin_inff();
    for (int k = 0; k < 4; ++k) { // empty slots: a box nothing can enter and a reference that says so
        WideChild& c = out->c[k];
        c.lo[0] = c.lo[1] = c.lo[2] = inf;
        c.hi[0] = c.hi[1] = c.hi[2] = -inf;
        c.ref = REF_POP;
        c.meta = 0u;
    }
    const gb_bvh_node& root = nodes[i];
    uint32_t meta = root.axis & 3u;
    const uint32_t child[2] = {i + 1, root.offset};
    for (int side = 0; side < 2; ++side) {
        const gb_bvh_node& c = nodes[child[side]];
        uint32_t slot[2] = {child[side], 0u};
        int n = 1;
        if (c.nprims == 0) {
            slot[0] = child[side] + 1;
            slot[1] = c.offset;
            n = 2;
            meta |= (uint32_t)(c.axis & 3u) << (2 + 2 * side);
        }
        for (int k = 0; k < n; ++k) {
            const gb_bvh_node& g = nodes[slot[k]];
            WideChild& w = out->c[2 * side + k];
            for (int a = 0; a < 3; ++a) { w.lo[a] = g.bmin[a]; w.hi[a] = g.bmax[a]; }
            w.ref = wideRefOf(nodes, wideIndex, slot[k]);
        }
    }
    out->c[0].meta = meta;
}

// The reference's ordered slab test (src/GoblinBVH.cpp:156-187) on sign-selected bounds, without
// early returns: same comparisons, same NaN behaviour for zero direction components.
GB_HD bool slabOrdered(float nearX, float nearY, float nearZ, float farX, float farY, float farZ, float ox, float oy,
    float oz, float ix, float iy, float iz, float mint, float maxt, float* tEntry) {
    float tMin = (nearX - ox) * ix;
    float tMax = (farX - ox) * ix;
    float tYMin = (nearY - oy) * iy;
    float tYMax = (farY - oy) * iy;
    bool miss = (tYMax < tMin) | (tYMin > tMax);
    tMin = tYMin > tMin ? tYMin : tMin;
    tMax = tYMax < tMax ? tYMax : tMax;
    float tZMin = (nearZ - oz) * iz;
    float tZMax = (farZ - oz) * iz;
    miss |= (tZMax < tMin) | (tZMin > tMax);
    tMin = tZMin > tMin ? tZMin : tMin;
    tMax = tZMax < tMax ? tZMax : tMax;
    *tEntry = tMin;
    return !miss & (tMin < maxt) & (tMax > mint);
}

// The same decision from per-axis min / max of the two plane distances (no sign selection, no ordered early-outs):
// 12 arithmetic + 8 min / max (two of them 3-input on sm_100) + 3 compares per box instead of ~36 instructions.
// For a ray WITHOUT a zero direction component it decides exactly like slabOrdered: the reciprocals are finite,
// no NaN can arise, min / max of the two plane distances IS the sign-selected near / far pair (subtraction and
// multiplication are monotone), and the chain of pairwise interval tests of the reference is the test that the
// three intervals share a point.  A zero component (+0 or -0: the reciprocal is infinite and the sign selection of
// the reference follows the sign bit it reads as "not negative") is where the two differ, so such rays -- flagged
// once per ray and per instance entry, bit 3 of the sign mask -- keep using slabOrdered.
GB_HD bool slabMinMax(float loX, float loY, float loZ, float hiX, float hiY, float hiZ, float ox, float oy, float oz,
    float ix, float iy, float iz, float mint, float maxt, float* tEntry) {
    const float ax = (loX - ox) * ix, bx = (hiX - ox) * ix;
    const float ay = (loY - oy) * iy, by = (hiY - oy) * iy;
    const float az = (loZ - oz) * iz, bz = (hiZ - oz) * iz;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    *tEntry = tn;
    return (tn <= tf) & (tn < maxt) & (tf > mint);
}
constexpr uint32_t NEG_ZERO_COMPONENT = 8u; // bit 3 of the direction sign mask: some component is +-0

// Put the four (reference, entry distance) pairs of a wide node, given in canonical order
// [LL, LR, RL, RR], into the reference's visit order for a ray with direction signs `neg`
// (bit a = direction component a is negative).
GB_HD void wideVisitOrder(uint32_t neg, uint32_t meta, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, float& t0,
    float& t1, float& t2, float& t3) {
    const bool s0 = (neg >> (meta & 3u)) & 1u, s1 = (neg >> ((meta >> 2) & 3u)) & 1u, s2 = (neg >> ((meta >> 4) & 3u)) & 1u;
    uint32_t a; float f;
    if (s1) { a = r0; r0 = r1; r1 = a; f = t0; t0 = t1; t1 = f; }
    if (s2) { a = r2; r2 = r3; r3 = a; f = t2; t2 = t3; t3 = f; }
    if (s0) {
        a = r0; r0 = r2; r2 = a; f = t0; t0 = t2; t2 = f;
        a = r1; r1 = r3; r3 = a; f = t1; t1 = t3; t3 = f;
    }
}

} // namespace gb
