/*
 * TEST INFRASTRUCTURE -- not part of the product, never loaded by it.
 *
 * A CPU emulation of what ONE LANE of the device traversal kernels does in the default (4-wide)
 * walk (goblin_b200/csrc/traverse.cuh, persistentTrace<ANY, false, true>), assembled from the very
 * functions the kernels are compiled from:
 *   goblin_b200/csrc/wide_node.h  node derivation, ordered slab test, visit order, stack bound
 *   goblin_b200/csrc/rt_core.cuh  triangle / sphere / disk tests, instance transforms
 * and from host restatements of the upload-time derivations (leaf-order triangle records, instance
 * tables).  The warp-level scheduling of the kernel (parking, batched stages, lane refill) never
 * reorders work within a ray, so this single-ray state machine has the same results.  The CPU test
 * suite checks it against the oracle's walk of the reference tree (ids, t, epsilon bit for bit) on
 * every test scene: the logic of the wide walk is then known to be right before a GPU sees it, and
 * the GPU tests only have to show that the kernel does what this emulation does.
 *
 * Build: make -C tests/native  ->  tests/native/_build/libwide_walk.so
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "goblin_b200.h"
#include "rt_core.cuh"
#include "wide_node.h"

using namespace gb;

namespace {

struct Tables {
    const gb_scene_desc* d;
    std::vector<WideNode> topWide, modelWide;
    std::vector<uint32_t> modelWideBase, modelRootRef;
    std::vector<float> triRec; // 12 floats per leaf slot: p0, e1, e2, face bits, pad
    uint32_t topRootRef = REF_NONE;
    int topDepth = 0, modelDepth = 0;
};

// depth of every node + numbering of the wide roots, as scanTree does in device.cu
void numberWide(const gb_bvh_node* nodes, uint32_t count, std::vector<uint32_t>& wideIndex, uint32_t* nWide, int* depthOut) {
    wideIndex.assign(count, WIDE_NOT_ROOT);
    std::vector<int> depth(count, 0);
    *nWide = 0;
    *depthOut = 0;
    for (uint32_t i = 0; i < count; ++i) {
        *depthOut = std::max(*depthOut, depth[i]);
        if (nodes[i].nprims == 0) {
            depth[i + 1] = depth[nodes[i].offset] = depth[i] + 1;
            if ((depth[i] & 1) == 0) wideIndex[i] = (*nWide)++;
        }
    }
}

void buildWide(const gb_bvh_node* nodes, uint32_t count, std::vector<WideNode>& out, uint32_t* rootRef, int* depth) {
    std::vector<uint32_t> wideIndex;
    uint32_t nWide = 0;
    numberWide(nodes, count, wideIndex, &nWide, depth);
    size_t base = out.size();
    out.resize(base + nWide);
    for (uint32_t i = 0; i < count; ++i) {
        if (wideIndex[i] != WIDE_NOT_ROOT) deriveWideNode(nodes, wideIndex.data(), i, &out[base + wideIndex[i]]);
    }
    *rootRef = count ? wideRefOf(nodes, wideIndex.data(), 0) : REF_NONE;
}

void buildTables(const gb_scene_desc* d, Tables& t) {
    t.d = d;
    buildWide(d->top_nodes, d->n_top_nodes, t.topWide, &t.topRootRef, &t.topDepth);
    t.modelWideBase.assign(d->n_models, 0u);
    t.modelRootRef.assign(d->n_models, REF_NONE);
    t.triRec.assign(12 * (size_t)d->n_tris, 0.0f);
    for (uint32_t m = 0; m < d->n_models; ++m) {
        const gb_model& md = d->models[m];
        if (md.kind != GB_GEOM_MESH) continue;
        t.modelWideBase[m] = (uint32_t)t.modelWide.size();
        int dm = 0;
        buildWide(d->model_nodes + md.node_offset, md.node_count, t.modelWide, &t.modelRootRef[m], &dm);
        t.modelDepth = std::max(t.modelDepth, dm);
        for (uint32_t k = 0; k < md.tri_count; ++k) { // k_derive_tris
            const uint32_t face = d->model_order[md.tri_offset + k];
            const uint32_t* vi = d->tri_index + 3 * ((size_t)md.tri_offset + face);
            const float* p0 = d->vert_pos + 3 * ((size_t)md.vert_offset + vi[0]);
            const float* p1 = d->vert_pos + 3 * ((size_t)md.vert_offset + vi[1]);
            const float* p2 = d->vert_pos + 3 * ((size_t)md.vert_offset + vi[2]);
            float* r = &t.triRec[12 * ((size_t)md.tri_offset + k)];
            for (int a = 0; a < 3; ++a) { r[a] = p0[a]; r[3 + a] = p1[a] - p0[a]; r[6 + a] = p2[a] - p0[a]; }
            std::memcpy(&r[9], &face, 4);
        }
    }
}

inline float4 row(const float* m, int r) { return make_float4(m[4 * r], m[4 * r + 1], m[4 * r + 2], m[4 * r + 3]); }

inline bool nodeTest(const gb_bvh_node& n, float3 o, float3 inv, uint32_t neg, float mint, float maxt) {
    const bool nx = neg & 1u, ny = neg & 2u, nz = neg & 4u;
    float t;
    return slabOrdered(nx ? n.bmax[0] : n.bmin[0], ny ? n.bmax[1] : n.bmin[1], nz ? n.bmax[2] : n.bmin[2],
        nx ? n.bmin[0] : n.bmax[0], ny ? n.bmin[1] : n.bmax[1], nz ? n.bmin[2] : n.bmax[2], o.x, o.y, o.z, inv.x, inv.y,
        inv.z, mint, maxt, &t);
}
inline uint32_t signs(float3 d) {
    return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u) |
           ((d.x == 0.0f) | (d.y == 0.0f) | (d.z == 0.0f) ? NEG_ZERO_COMPONENT : 0u);
}
// boxTest of traverse.cuh: min / max form unless the ray has a zero direction component
inline bool boxTest(const float lo[3], const float hi[3], float3 o, float3 inv, uint32_t neg, float mint, float maxt, float* t) {
#if GB_SLAB_MINMAX
    if (!(neg & NEG_ZERO_COMPONENT)) return slabMinMax(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], o.x, o.y, o.z, inv.x, inv.y, inv.z, mint, maxt, t);
#endif
    const bool nx = neg & 1u, ny = neg & 2u, nz = neg & 4u;
    return slabOrdered(nx ? hi[0] : lo[0], ny ? hi[1] : lo[1], nz ? hi[2] : lo[2], nx ? lo[0] : hi[0], ny ? lo[1] : hi[1],
        nz ? lo[2] : hi[2], o.x, o.y, o.z, inv.x, inv.y, inv.z, mint, maxt, t);
}

struct Entry { uint32_t ref; float t; };

// One ray through the lane state machine of traverse.cuh.  Returns found; *maxSp = deepest stack use.
bool walk(const Tables& T, const gb_ray& ray, bool any, HitRec* hit, int* maxSp) {
    const gb_scene_desc* d = T.d;
    float3 o = make3(ray.o[0], ray.o[1], ray.o[2]), dir = make3(ray.d[0], ray.d[1], ray.d[2]);
    const float3 wo = o, wd = dir;
    float3 inv = make3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    uint32_t neg = signs(dir);
    const float mint = ray.mint;
    float maxt = ray.maxt;
    std::vector<Entry> st(wideStackEntries(T.topDepth) + wideStackEntries(T.modelDepth) + 2);
    int sp = 0, spFloor = 0, level = 0, curSlot = 0;
    uint32_t cur = REF_NONE, instNext = 0, instEnd = 0, triBase = 0, nodeBase = 0;
    const WideNode* wide = T.topWide.data();
    bool found = false, fin = false;
    hit->inst = -1; hit->prim = 0; hit->t = maxt; hit->b1 = hit->b2 = 0.0f;
    if (d->n_top_nodes && nodeTest(d->top_nodes[0], o, inv, neg, mint, maxt)) cur = T.topRootRef;
    auto push = [&](uint32_t ref, float t) {
        if ((size_t)sp >= st.size()) { *maxSp = 1 << 20; st.resize(st.size() * 2 + 8); } // bound violated: reported
        st[sp++] = Entry{ref, t};
        if (sp > *maxSp) *maxSp = sp;
    };
    while (!fin) {
        const bool isLeaf = (cur & REF_LEAF) && cur < REF_POP;
        if (isLeaf && level == 1) { // ---- triangle leaf
            uint32_t first = cur & REF_INDEX, count = 1;
            if (cur & REF_MULTI) {
                const gb_bvh_node& n = d->model_nodes[(size_t)nodeBase + first];
                count = n.nprims;
                first = n.offset;
            }
            cur = REF_POP;
            for (uint32_t k = 0; k < count && !fin; ++k) {
                const float* r = &T.triRec[12 * ((size_t)triBase + first + k)];
                float t, b1, b2;
                if (triangleTest(make3(r[0], r[1], r[2]), make3(r[3], r[4], r[5]), make3(r[6], r[7], r[8]), o, dir, mint, maxt,
                        &t, &b1, &b2)) {
                    found = true;
                    if (any) { fin = true; break; }
                    maxt = t;
                    hit->t = t; hit->b1 = b1; hit->b2 = b2; hit->inst = curSlot; hit->prim = (int)(first + k);
                }
            }
        } else if (cur == REF_NONE || (isLeaf && level == 0)) { // ---- level change
            bool iterate = true;
            if (cur == REF_NONE) {
                if (level == 1) {
                    level = 0; spFloor = 0; wide = T.topWide.data();
                    o = wo; dir = wd;
                    inv = make3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
                    neg = signs(dir);
                } else { fin = true; iterate = false; }
            } else {
                uint32_t first = cur & REF_INDEX, count = 1;
                if (cur & REF_MULTI) { count = d->top_nodes[first].nprims; first = d->top_nodes[first].offset; }
                instNext = first; instEnd = first + count;
            }
            while (iterate) {
                bool descended = false;
                while (instNext < instEnd) {
                    const uint32_t slot = instNext++;
                    const gb_instance& in = d->instances[d->top_order[slot]];
                    const gb_model& md = d->models[in.model];
                    const float4 r0 = row(in.to_object, 0), r1 = row(in.to_object, 1), r2 = row(in.to_object, 2);
                    const float3 oo = xfPoint(r0, r1, r2, o), od = xfVector(r0, r1, r2, dir);
                    if (md.kind == GB_GEOM_MESH) {
                        if (md.node_count == 0) continue;
                        const float3 oinv = make3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
                        const uint32_t oneg = signs(od);
                        if (!nodeTest(d->model_nodes[md.node_offset], oo, oinv, oneg, mint, maxt)) continue;
                        level = 1; spFloor = sp; curSlot = (int)slot;
                        nodeBase = md.node_offset; triBase = md.tri_offset;
                        wide = T.modelWide.data() + T.modelWideBase[in.model];
                        o = oo; dir = od; inv = oinv; neg = oneg;
                        cur = T.modelRootRef[in.model];
                        descended = true;
                        break;
                    }
                    float t;
                    const bool h = md.kind == GB_GEOM_SPHERE ? sphereTest(md.radius, oo, od, mint, maxt, &t)
                                                             : diskTest(md.radius, oo, od, mint, maxt, &t);
                    if (h) {
                        found = true;
                        if (any) { fin = true; break; }
                        maxt = t;
                        hit->t = t; hit->b1 = hit->b2 = 0.0f; hit->inst = (int)slot; hit->prim = 0;
                    }
                }
                if (descended || fin) break;
                cur = REF_NONE;
                while (sp > 0) {
                    const Entry e = st[--sp];
                    if (e.t < maxt) { cur = e.ref; break; }
                }
                if (cur == REF_NONE) { fin = true; break; }
                if (!(cur & REF_LEAF)) break;
                uint32_t first = cur & REF_INDEX, count = 1;
                if (cur & REF_MULTI) { count = d->top_nodes[first].nprims; first = d->top_nodes[first].offset; }
                instNext = first; instEnd = first + count;
            }
        } else if (!(cur & REF_LEAF)) { // ---- interior stage: four box tests
            const WideNode& w = wide[cur];
            uint32_t r[4]; float t[4];
            for (int k = 0; k < 4; ++k) {
                const WideChild& c = w.c[k];
                r[k] = boxTest(c.lo, c.hi, o, inv, neg, mint, maxt, &t[k]) ? c.ref : REF_POP;
            }
            wideVisitOrder(neg, w.c[0].meta, r[0], r[1], r[2], r[3], t[0], t[1], t[2], t[3]);
            uint32_t next = r[3]; float tn = t[3];
            for (int k = 2; k >= 0; --k) {
                if (r[k] != REF_POP) { if (next != REF_POP) push(next, tn); next = r[k]; tn = t[k]; }
            }
            cur = next;
        } else { // ---- pop stage (cur == REF_POP)
            if (sp > spFloor) {
                const Entry e = st[--sp];
                if (e.t < maxt) cur = e.ref;
            } else {
                cur = REF_NONE;
                if (level == 1 && spFloor == 0 && instNext >= instEnd) fin = true;
            }
        }
    }
    return found;
}

} // namespace

extern "C" {

// closest != null: gb_hit per ray as gb_trace_closest reports it; occluded != null: any-hit flags.
// max_stack: deepest stack use over the batch; stack_bound: what the device allocates per thread.
int ww_trace(const gb_scene_desc* d, const gb_ray* rays, size_t n, gb_hit* closest, uint8_t* occluded, int* max_stack,
    int* stack_bound) {
    Tables T;
    buildTables(d, T);
    *stack_bound = wideStackEntries(T.topDepth) + wideStackEntries(T.modelDepth) + 2;
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<int> deepest(nt, 0);
    std::vector<std::thread> pool;
    for (unsigned k = 0; k < nt; ++k) {
        pool.emplace_back([&, k]() {
            for (size_t i = k; i < n; i += nt) {
                HitRec h;
                if (closest) {
                    gb_hit out;
                    if (walk(T, rays[i], false, &h, &deepest[k])) {
                        const gb_model& md = d->models[d->instances[d->top_order[h.inst]].model];
                        out.t = h.t;
                        out.eps = 1e-3f * h.t;
                        out.inst = (int)d->top_order[h.inst];
                        out.prim = 0;
                        if (md.kind == GB_GEOM_MESH) std::memcpy(&out.prim, &T.triRec[12 * ((size_t)md.tri_offset + h.prim) + 9], 4);
                    } else {
                        out.t = 0.0f; out.eps = 0.0f; out.inst = -1; out.prim = -1;
                    }
                    closest[i] = out;
                }
                if (occluded) occluded[i] = walk(T, rays[i], true, &h, &deepest[k]) ? 1 : 0;
            }
        });
    }
    for (auto& th : pool) th.join();
    *max_stack = 0;
    for (int v : deepest) *max_stack = std::max(*max_stack, v);
    return 0;
}

} // extern "C"
