// Persistent two-level BVH traversal with per-lane ray refill.
//
// What is walked is the reference's tree in the reference's order
// (src/GoblinBVH.cpp:189-280: near child first by dirIsNeg[axis], far child
// pushed, leaves tested as they are reached, t <= maxt accepted), so hit ids and
// distances equal the CPU build's.  How it is walked is B200-shaped:
//   * !WIDE (the default): "pair nodes", one 64-byte record per interior node with
//     BOTH child boxes (two 256-bit loads = two 32-byte sectors, two box tests per
//     step).  This is the walk whose box-test count equals the reference's exactly:
//     its STATS instantiation feeds gb_get_counters and the roofline's N_node;
//   * WIDE (GB_TRACE_WIDE): 128-byte 4-wide nodes (wide_node.h) -- two binary levels
//     per step, four slab tests per step, children visited in the reference's order.
//     Built because the dependent-fetch chain looked like the limiter; measured 8 - 18 %
//     SLOWER than the pair walk on every scene: it tests ~30 % more boxes,
//     and the kernels are bound by issue slots and L1 gather wavefronts, both of which
//     scale with boxes tested, not by the latency of the chain;
//   * the far children's entry distances ride on the stack (8-byte entries in a
//     shared-memory column per thread) and are re-checked against the shrunk
//     maxt when popped -- exactly the box test the reference evaluates at pop
//     time, since nothing else in that test depends on maxt;
//   * branch-free slab tests that keep the reference's comparison structure
//     (and therefore its NaN behaviour for zero direction components);
//   * staged scheduling: every loop iteration runs a converged interior stage
//     and pop stage; triangle tests and level changes are batched across the
//     warp (lanes park until enough of them need the same stage);
//   * lanes that finish a ray pull the next one from a global cursor with a
//     warp-aggregated atomic instead of idling until the whole warp is done;
//   * any-hit walks (occluded(): shadow and AO rays) keep the reference's box and
//     triangle tests but not its visiting order, which cannot change their answer:
//     left child first for every lane (GB_ANY_FIXED_ORDER).
#pragma once
#include "rt_core.cuh"
#include "wide_node.h"

namespace gb {

#ifndef GB_STEPS_PAIR
#define GB_STEPS_PAIR 3 // measured (profiles/r02/call5_stdout.txt): 3 beats 2 by 0.4 % (S1) / 1.9 % (S4) / 1.8 % (S3); 1 loses 5 - 8 %, 4 = 3
#endif
#ifndef GB_STEPS_WIDE
#define GB_STEPS_WIDE 2
#endif
#ifndef GB_ANY_FIXED_ORDER
#define GB_ANY_FIXED_ORDER 1 // any-hit walks visit the left child first (0: the reference's near child first, for A/B runs)
#endif
// interior + pop stages between two scheduling checks
constexpr int kStepsPerCheckPair = GB_STEPS_PAIR, kStepsPerCheckWide = GB_STEPS_WIDE;
// scheduling knobs live in DeviceScene::tune (gb_set_tuning): refillBelow = pull new rays when
// fewer lanes than this are busy; leafBatch / levelBatch = run the triangle / level stage once
// this many lanes wait for it; moveFloor = ... or when fewer lanes than this can still move

// One 32-byte sector as two float4: a single 256-bit read-only load where the build allows it.
// L1 policy (GB_NODE_HINT / GB_TRI_HINT, A/B builds): the true L1 hit rate of the node fetches is ~15 % (ncu r02i: the
// 55 % of the 128-bit layout was the second half of a sector hitting behind the first); a ray's first ~8 pair nodes are
// shared by most rays, its triangle records and path state are touched once.  evict_last asks L1 to keep nodes,
// no_allocate keeps the one-touch records out of it.
#ifndef GB_NODE_HINT
#define GB_NODE_HINT 0 // 1: .L1::evict_last on interior records
#endif
#ifndef GB_TRI_HINT
#define GB_TRI_HINT 0  // 1: .L1::no_allocate on triangle records
#endif
struct Sector { float4 lo, hi; };
#define GB_LD256_ASM(QUAL)                                                                                              \
    asm volatile("ld.global.nc" QUAL ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                         \
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) \
                 : "l"(p))
__device__ __forceinline__ Sector ldgSector(const float4* p) { // interior records
    Sector r;
#if GB_LD256
#if GB_NODE_HINT == 1
    GB_LD256_ASM(".L1::evict_last");
#else
    GB_LD256_ASM("");
#endif
#else
    r.lo = __ldg(p);
    r.hi = __ldg(p + 1);
#endif
    return r;
}
__device__ __forceinline__ Sector ldgSectorOnce(const float4* p) { // records a ray touches once (triangles)
    Sector r;
#if GB_LD256
#if GB_TRI_HINT == 1
    GB_LD256_ASM(".L1::no_allocate");
#else
    GB_LD256_ASM("");
#endif
#else
    r.lo = __ldg(p);
    r.hi = __ldg(p + 1);
#endif
    return r;
}

// Per-thread columns in shared memory: entry k of thread t lives at [k * blockDim.x + t],
// so a warp's pushes / pops are contiguous.
// GB_DEBUG_STACK=1 (debug builds, tools/build_variants.sh): every stack access is bounds-checked against the
// column height the host sized from the trees; the first violation is recorded in g_stackViolation (kind, index,
// limit, block) and the access is dropped.  gb_debug_stack_violation reads it back.
#ifndef GB_DEBUG_STACK
#define GB_DEBUG_STACK 0
#endif
#if GB_DEBUG_STACK
__device__ int g_stackViolation[4] = {0, 0, 0, 0};
#endif
struct TravStack {
    uint2* base;
    unsigned int stride;
#if GB_DEBUG_STACK
    int limit;
    __device__ __forceinline__ bool bad(int kind, int k) const {
        if (k >= 0 && k < limit) return false;
        if (atomicCAS(&g_stackViolation[0], 0, kind) == 0) {
            g_stackViolation[1] = k; g_stackViolation[2] = limit; g_stackViolation[3] = (int)blockIdx.x;
        }
        return true;
    }
#endif
    __device__ __forceinline__ void put(int k, unsigned int ref, float tmin) {
#if GB_DEBUG_STACK
        if (bad(1, k)) return;
#endif
        base[k * stride] = make_uint2(ref, __float_as_uint(tmin));
    }
    __device__ __forceinline__ uint2 get(int k) const {
#if GB_DEBUG_STACK
        if (bad(2, k)) return make_uint2(REF_NONE, 0u);
#endif
        return base[k * stride];
    }
};

// The reference's ordered slab test on sign-selected bounds, without early returns (wide_node.h).
__device__ __forceinline__ bool slabNoBranch(float nearX, float nearY, float nearZ, float farX, float farY, float farZ,
    float3 o, float3 inv, float mint, float maxt, float* tEntry) {
    return slabOrdered(nearX, nearY, nearZ, farX, farY, farZ, o.x, o.y, o.z, inv.x, inv.y, inv.z, mint, maxt, tEntry);
}

// One original 32-byte node (two float4) against a ray: used for the root of each level.
__device__ __forceinline__ bool rootTest(float4 n0, float4 n1, float3 o, float3 inv, unsigned int neg, float mint,
    float maxt) {
    bool nx = neg & 1u, ny = neg & 2u, nz = neg & 4u;
    float t;
    return slabNoBranch(nx ? n0.w : n0.x, ny ? n1.x : n0.y, nz ? n1.y : n0.z, nx ? n0.x : n0.w, ny ? n0.y : n1.x,
        nz ? n0.z : n1.y, o, inv, mint, maxt, &t);
}

// bits 0..2: dirIsNeg[axis]; bit 3: some component is +-0 (the ordered slab test must decide, wide_node.h)
__device__ __forceinline__ unsigned int signBits(float3 d) {
    return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u) |
           ((d.x == 0.0f) | (d.y == 0.0f) | (d.z == 0.0f) ? NEG_ZERO_COMPONENT : 0u);
}

// A box stored as the reference's node words (n0 = bmin.xyz, bmax.x; n1 = bmax.yz, ..) against a ray: the ordered
// form when the ray has a zero direction component, the min / max form otherwise.  `ordered` is warp-divergent only
// in the presence of such rays.
__device__ __forceinline__ bool boxTest(float4 n0, float4 n1, float3 o, float3 inv, unsigned int neg, float mint, float maxt,
    float* tEntry) {
#if GB_SLAB_MINMAX
    if (!(neg & NEG_ZERO_COMPONENT)) {
        return slabMinMax(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, o.x, o.y, o.z, inv.x, inv.y, inv.z, mint, maxt, tEntry);
    }
#endif
    const bool nx = neg & 1u, ny = neg & 2u, nz = neg & 4u;
    return slabOrdered(nx ? n0.w : n0.x, ny ? n1.x : n0.y, nz ? n1.y : n0.z, nx ? n0.x : n0.w, ny ? n0.y : n1.x,
        nz ? n0.z : n1.y, o.x, o.y, o.z, inv.x, inv.y, inv.z, mint, maxt, tEntry);
}

// Policy interface (all members __device__):
//   bool fetch(unsigned long long item, float3* o, float3* d, float* mint, float* maxt)
//        -- loads / generates ray `item`; false = nothing to trace for this item
//   void finish(bool done, unsigned long long item, bool found, const HitRec& hit)
//        -- warp-collective: every lane calls it, `done` marks lanes whose ray just ended
//
// s_stack: stackEntries x blockDim.x uint2, s_ray: 6 x blockDim.x floats (the world-space ray
// while a lane is inside an instance).
template <bool ANY, bool STATS, bool WIDE, typename Policy>
__device__ __forceinline__ void persistentTrace(const DeviceScene& sc, Policy& pol, unsigned long long n,
    unsigned long long* head, uint2* s_stack, float* s_ray, TraceStats& ts, unsigned int* raysDone, int stackEntries) {
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31;
#if GB_DEBUG_STACK
    TravStack st{s_stack + threadIdx.x, blockDim.x, stackEntries};
#else
    TravStack st{s_stack + threadIdx.x, blockDim.x};
    (void)stackEntries;
#endif
    float* wr = s_ray + threadIdx.x;
    const unsigned int ws = blockDim.x;

    bool have = false, exhausted = false;
    unsigned long long item = 0;
    float3 o = make3(0, 0, 0), d = make3(0, 0, 0), inv = make3(0, 0, 0);
    unsigned int neg = 0, cur = REF_NONE;
    float mint = 0.0f, maxt = 0.0f;
    int sp = 0, spFloor = 0, level = 0, curSlot = 0;
    const float4* pairs = WIDE ? sc.topWide : sc.topPairs; // interior records of the current level
    unsigned int triBase = 0, nodeBase = 0, instNext = 0, instEnd = 0;
    HitRec hit;
    hit.inst = -1; hit.prim = 0; hit.t = 0.0f; hit.b1 = hit.b2 = 0.0f;
    bool found = false;

    for (;;) {
        // ------------------------------------------------ refill idle lanes
        unsigned int idle = __ballot_sync(FULL, !have);
        if (idle && !exhausted) {
            const unsigned int leader = __ffs(idle) - 1;
            const unsigned int want = __popc(idle);
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(head, (unsigned long long)want);
            base = __shfl_sync(FULL, base, leader);
            if (base + want >= n) exhausted = true;
            if (!have) {
                unsigned long long my = base + __popc(idle & ((1u << lane) - 1u));
                if (my < n) {
                    item = my;
                    have = true;
                    found = false;
                    hit.inst = -1; hit.prim = 0; hit.b1 = hit.b2 = 0.0f;
                    level = 0; sp = 0; spFloor = 0; instNext = instEnd = 0;
                    pairs = WIDE ? sc.topWide : sc.topPairs;
                    cur = REF_NONE;
                    if (pol.fetch(item, &o, &d, &mint, &maxt)) {
                        ++*raysDone;
                        wr[0] = o.x; wr[ws] = o.y; wr[2 * ws] = o.z;
                        wr[3 * ws] = d.x; wr[4 * ws] = d.y; wr[5 * ws] = d.z;
                        inv = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                        neg = signBits(d);
                        if (sc.nTopNodes) {
                            float4 n0 = __ldg(sc.topNodes), n1 = __ldg(sc.topNodes + 1);
                            if (STATS) ts.nodes++;
                            if (rootTest(n0, n1, o, inv, neg, mint, maxt)) cur = sc.topRootRef;
                        }
                    }
                    hit.t = maxt;
                }
            }
        }
        if (__ballot_sync(FULL, have) == 0) break;

        // ------------------------------------------------ traversal burst
        // Every iteration runs a converged interior stage (two box tests for each lane holding an
        // interior node) and a converged pop stage.  Leaf work (triangle tests) and level changes
        // (instance entry / exit, ray end) are longer and rarer, so the lanes that need them park
        // until enough have gathered -- or nobody else can move -- and then run them together.
        // Parking does not reorder anything within a ray: exactness is untouched.
        bool fin = false;
        for (;;) {
            const bool live = have && !fin;
            const bool isLeaf = live && (cur & REF_LEAF) && cur < REF_POP;
            const bool wantTri = isLeaf && level == 1;
            const bool wantLvl = live && (cur == REF_NONE || (isLeaf && level == 0));
            const unsigned int triMask = __ballot_sync(FULL, wantTri);
            const unsigned int lvlMask = __ballot_sync(FULL, wantLvl);
            bool runTri = false, runLvl = false;
            if (triMask | lvlMask) {
                const unsigned int nTri = __popc(triMask), nLvl = __popc(lvlMask);
                const unsigned int nMove = __popc(__ballot_sync(FULL, live)) - nTri - nLvl;
                runTri = nTri >= sc.tune.leafBatch || (nTri && nMove < sc.tune.moveFloor);
                runLvl = nLvl >= sc.tune.levelBatch || (nLvl && nMove < sc.tune.moveFloor);
            }

            if (runTri && wantTri) { // ---- triangle leaf
                unsigned int first = cur & REF_INDEX, count = 1;
                if (cur & REF_MULTI) {
                    const float4 n1 = __ldg(sc.modelNodes + 2 * ((size_t)nodeBase + first) + 1);
                    count = __float_as_uint(n1.w) & 0xffu;
                    first = __float_as_uint(n1.z);
                }
                cur = REF_POP;
                for (unsigned int k = 0; k < count; ++k) {
                    const float4* tr = sc.triRec + kTriRecVec4 * ((size_t)triBase + first + k);
#if GB_LD256
                    const Sector ab = ldgSectorOnce(tr);
                    const float4 a = ab.lo, b = ab.hi, c = __ldg(tr + 2);
#else
                    const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
#endif
                    if (STATS) ts.prims++;
                    float t, b1, b2;
                    if (triangleTest(make3(a.x, a.y, a.z), make3(a.w, b.x, b.y), make3(b.z, b.w, c.x), o, d, mint,
                            maxt, &t, &b1, &b2)) {
                        found = true;
                        if (ANY) { fin = true; break; }
                        maxt = t;
                        hit.t = t; hit.b1 = b1; hit.b2 = b2;
                        hit.inst = curSlot;
                        hit.prim = (int)(first + k);
                    }
                }
            }
            if (runLvl && wantLvl) { // ---- level change: instance exit, instance entry, ray end
                // Everything a ray does in world space between two mesh descents happens in one
                // visit: leave the instance, walk the remaining instances of the leaf, pop the
                // top-level stack, handle the next instance leaf ... until the lane is inside a
                // mesh again, stands on an interior node of the top level, or the ray ends.
                bool iterate = true;
                if (cur == REF_NONE) {
                    if (level == 1) { // this instance is exhausted: back to world space
                        level = 0;
                        spFloor = 0;
                        pairs = WIDE ? sc.topWide : sc.topPairs;
                        o = make3(wr[0], wr[ws], wr[2 * ws]);
                        d = make3(wr[3 * ws], wr[4 * ws], wr[5 * ws]);
                        inv = make3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                        neg = signBits(d);
                    } else {
                        fin = true;
                        iterate = false;
                    }
                } else { // instance leaf of the top level
                    unsigned int first = cur & REF_INDEX, count = 1;
                    if (cur & REF_MULTI) {
                        const float4 n1 = __ldg(sc.topNodes + 2 * (size_t)first + 1);
                        count = __float_as_uint(n1.w) & 0xffu;
                        first = __float_as_uint(n1.z);
                    }
                    instNext = first;
                    instEnd = first + count;
                }
                while (iterate) {
                    bool descended = false;
                    while (instNext < instEnd) {
                        const unsigned int slot = instNext++;
                        const float4* m = sc.instToObject + 3 * (size_t)slot;
                        const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
                        const int4 info = __ldg(sc.instInfo + slot);
                        if (STATS) ts.insts++;
                        // Transform::invertRay: the direction is not renormalised, t is shared
                        const float3 oo = xfPoint(r0, r1, r2, o);
                        const float3 od = xfVector(r0, r1, r2, d);
                        if (info.x == GB_GEOM_MESH) {
                            const int4 info2 = __ldg(sc.instInfo2 + slot);
                            if (info2.z == 0) continue; // empty mesh: BVH::intersect returns false
                            const float3 oinv = make3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
                            const unsigned int oneg = signBits(od);
                            const float4* root = sc.modelNodes + 2 * (size_t)(unsigned int)info.y;
                            const float4 n0 = __ldg(root), n1 = __ldg(root + 1);
                            if (STATS) ts.nodes++;
                            if (!rootTest(n0, n1, oo, oinv, oneg, mint, maxt)) continue;
                            level = 1;
                            spFloor = sp;
                            curSlot = (int)slot;
                            nodeBase = (unsigned int)info.y;
                            triBase = (unsigned int)info.z;
                            pairs = WIDE ? sc.modelWide + 8 * (size_t)(unsigned int)info2.w
                                         : sc.modelPairs + 4 * (size_t)(unsigned int)info2.y;
                            o = oo; d = od; inv = oinv; neg = oneg;
                            cur = (unsigned int)info2.x;
                            descended = true;
                            break;
                        }
                        float t;
                        const float radius = __int_as_float(info.w);
                        if (STATS) ts.prims++;
                        const bool h = info.x == GB_GEOM_SPHERE ? sphereTest(radius, oo, od, mint, maxt, &t)
                                                                : diskTest(radius, oo, od, mint, maxt, &t);
                        if (h) {
                            found = true;
                            if (ANY) { fin = true; break; }
                            maxt = t;
                            hit.t = t; hit.b1 = 0.0f; hit.b2 = 0.0f;
                            hit.inst = (int)slot;
                            hit.prim = 0;
                        }
                    }
                    if (descended || fin) break;
                    // the leaf is done: next entry of the top-level stack
                    cur = REF_NONE;
                    while (sp > 0) {
                        const uint2 e = st.get(--sp);
                        if (STATS) ts.nodes++;
                        if (__uint_as_float(e.y) < maxt) { cur = e.x; break; }
                    }
                    if (cur == REF_NONE) { fin = true; break; }
                    if (!(cur & REF_LEAF)) break; // an interior node of the top level: interior stage
                    unsigned int first = cur & REF_INDEX, count = 1;
                    if (cur & REF_MULTI) {
                        const float4 n1 = __ldg(sc.topNodes + 2 * (size_t)first + 1);
                        count = __float_as_uint(n1.w) & 0xffu;
                        first = __float_as_uint(n1.z);
                    }
                    instNext = first;
                    instEnd = first + count;
                }
            }
#pragma unroll
            for (int rep = 0; rep < (WIDE ? kStepsPerCheckWide : kStepsPerCheckPair); ++rep) {
                if (WIDE) {
                    if (have && !fin && !(cur & REF_LEAF)) { // ---- interior stage: four box tests (wide_node.h)
                        const float4* p = pairs + 8 * (size_t)cur;
                        const Sector sa = ldgSector(p), sb = ldgSector(p + 2), sc4 = ldgSector(p + 4), sd = ldgSector(p + 6);
                        const float4 a0 = sa.lo, a1 = sa.hi, b0 = sb.lo, b1 = sb.hi, c0 = sc4.lo, c1 = sc4.hi, d0 = sd.lo, d1 = sd.hi;
                        // the ordered form for all four (measured: the min / max form with its zero-component fallback
                        // doubles the code of this stage and is 15 % slower here; the pair walk below gains 3 - 6 % from it)
                        const bool nx = neg & 1u, ny = neg & 2u, nz = neg & 4u;
                        float t0, t1, t2, t3;
#define GB_WIDE_BOX(n0, n1, tt)                                                                                   \
    (slabNoBranch(nx ? n0.w : n0.x, ny ? n1.x : n0.y, nz ? n1.y : n0.z, nx ? n0.x : n0.w, ny ? n0.y : n1.x,       \
         nz ? n0.z : n1.y, o, inv, mint, maxt, &tt)                                                                \
            ? __float_as_uint(n1.z)                                                                                \
            : REF_POP)
                        // a child that is missed (or an empty slot, whose reference already says so) becomes REF_POP
                        unsigned int r0 = GB_WIDE_BOX(a0, a1, t0), r1 = GB_WIDE_BOX(b0, b1, t1);
                        unsigned int r2 = GB_WIDE_BOX(c0, c1, t2), r3 = GB_WIDE_BOX(d0, d1, t3);
#undef GB_WIDE_BOX
                        wideVisitOrder(neg, __float_as_uint(a1.w), r0, r1, r2, r3, t0, t1, t2, t3);
                        // nearest hit child next; the others wait on the stack, nearest on top
                        unsigned int next = r3;
                        float tn = t3;
                        if (r2 != REF_POP) { if (next != REF_POP) st.put(sp++, next, tn); next = r2; tn = t2; }
                        if (r1 != REF_POP) { if (next != REF_POP) st.put(sp++, next, tn); next = r1; tn = t1; }
                        if (r0 != REF_POP) { if (next != REF_POP) st.put(sp++, next, tn); next = r0; tn = t0; }
                        cur = next;
                    }
                } else if (have && !fin && !(cur & REF_LEAF)) { // ---- interior stage: two box tests
                    const float4* p = pairs + 4 * (size_t)cur;
                    const Sector s01 = ldgSector(p), s23 = ldgSector(p + 2);
                    const float4 q0 = s01.lo, q1 = s01.hi, q2 = s23.lo;
                    const uint4 q3 = make_uint4(__float_as_uint(s23.hi.x), __float_as_uint(s23.hi.y), __float_as_uint(s23.hi.z), __float_as_uint(s23.hi.w));
                    float tL, tR;
                    // pair record: left box = (q0.xyz | q0.w q1.xy), right box = (q1.zw q2.x | q2.yzw)
                    const bool hitL = boxTest(q0, q1, o, inv, neg, mint, maxt, &tL);
                    const bool hitR = boxTest(make_float4(q1.z, q1.w, q2.x, q2.y), make_float4(q2.z, q2.w, 0.0f, 0.0f), o, inv, neg,
                        mint, maxt, &tR);
#if GB_ANY_FIXED_ORDER
                    // Any-hit walks (shadow, AO): "is some triangle of some reachable leaf hit inside [mint, maxt]" does not
                    // depend on the order of the visits, and maxt never shrinks, so they always go left first: no
                    // ordering selects, and the lanes of a warp agree on the order whatever their directions.  ncu on
                    // k_ao (profiles/r02/r02m_k_ao_*): 17.7 % fewer warp instructions, 17.9 instead of 17.1 active lanes,
                    // 8 % fewer L1 sectors; k_ao - 17 %, k_shadow - 6.5 % (profiles/r02/call21_stdout.txt).  The counting
                    // instantiation keeps the reference's order, so gb_get_counters still reports the reference's tests.
                    if (ANY && !STATS) {
                        if (hitL & hitR) st.put(sp++, q3.y, -INFINITY);
                        cur = hitL ? q3.x : (hitR ? q3.y : REF_POP);
                    } else {
#endif
                    const bool rightFirst = (neg >> (q3.z & 3u)) & 1u; // dirIsNeg[axis]
                    const unsigned int nearRef = rightFirst ? q3.y : q3.x, farRef = rightFirst ? q3.x : q3.y;
                    const bool hitN = rightFirst ? hitR : hitL, hitF = rightFirst ? hitL : hitR;
                    const float tF = rightFirst ? tL : tR;
                    if (STATS) {
                        // the reference tests the near child now and the far child when it is popped
                        ts.nodes++;
                        st.put(sp++, farRef, hitF ? tF : INFINITY);
                        cur = hitN ? nearRef : REF_POP;
                    } else {
                        if (hitN & hitF) st.put(sp++, farRef, tF);
                        cur = hitN ? nearRef : (hitF ? farRef : REF_POP);
                    }
#if GB_ANY_FIXED_ORDER
                    }
#endif
                }
                // ---- pop stage
#if defined(GB_POP_LANE_LOOP) && GB_POP_LANE_LOOP
                // every lane that needs an entry pops until it holds a node or its level is exhausted (a divergent loop)
                if (have && !fin) {
                    while (cur == REF_POP) {
                        if (sp > spFloor) {
                            const uint2 e = st.get(--sp);
                            if (STATS) ts.nodes++;
                            if (__uint_as_float(e.y) < maxt) cur = e.x;
                        } else {
                            cur = REF_NONE;
                            if (level == 1 && spFloor == 0 && instNext >= instEnd) fin = true;
                        }
                    }
                }
#else
                // one entry per trip for every lane that needs one
                while (__any_sync(FULL, have && !fin && cur == REF_POP)) {
                    if (have && !fin && cur == REF_POP) {
                        if (sp > spFloor) {
                            const uint2 e = st.get(--sp);
                            if (STATS) ts.nodes++;
                            if (__uint_as_float(e.y) < maxt) cur = e.x;
                        } else {
                            cur = REF_NONE;
                            // world space with an empty stack, or nothing left in this instance, in
                            // its top-level leaf and on the top-level stack: the ray ends here, no
                            // level change needed
                            if (level == 1 && spFloor == 0 && instNext >= instEnd) fin = true;
#if defined(GB_POP_END_WORLD) && GB_POP_END_WORLD
                            // round 1's dropped variant: also end world-space rays here.  Kept behind a
                            // macro for the fault hunt of tools/fault_hunt.sh; never part of a shipped build.
                            if (level == 0) fin = true;
#endif
                        }
                    }
                }
#endif
            }
            const unsigned int busy = __ballot_sync(FULL, have && !fin);
            if (busy == 0) break;
            if (!exhausted && __popc(busy) < sc.tune.refillBelow) break;
        }
        pol.finish(have && fin, item, found, hit);
        if (fin) have = false;
    }
}

} // namespace gb
