// sm_100a kernels, the device context and the GPU half of the C ABI.
//
// Execution model (DESIGN.md has the long form):
//   * gb_trace_*: one persistent kernel; warps pull 32-ray batches from a
//     global cursor, each thread walks the reference's two-level BVH with a
//     private stack column in shared memory, nodes arrive as two 128-bit
//     read-only loads.
//   * gb_render: a wavefront path tracer over "waves" of a few million camera
//     samples that live in SoA path-state arrays:
//       raygen -> [extend -> shade<material> x3 -> shadow] x (depth-1)
//              -> extend -> emission-only shade -> film accumulate
//     extend and shadow are the persistent traversal kernels; shade kernels are
//     binned by material through index queues that the extend kernel fills
//     with warp-aggregated atomics; the film kernel splats a pixel tile in
//     shared memory and flushes it with one vector atomic per pixel.
// No tensor cores, no TMA: traversal is a dependent gather of 32-byte nodes,
// not a dense contraction (north_star).
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "capi_util.h"
#include "device_scene.h"
#include "image_io.h"
#include "rt_core.cuh"
#include "traverse.cuh"
#include "shade.cuh"
#include "texture.cuh"
#include "mask.cuh"

namespace gb {

#ifndef GB_TRACE_BLOCK
#define GB_TRACE_BLOCK 128
#define GB_TRACE_MIN_BLOCKS 7
#endif
#ifndef GB_WIDE_MIN_BLOCKS
#define GB_WIDE_MIN_BLOCKS 5 // the 4-wide step holds four boxes at once: ~100 registers
#endif
constexpr int kTraceBlock = GB_TRACE_BLOCK;   // threads per traversal block
// resident blocks per SM the register allocation must allow: 7 x 128 threads -> 73 registers,
// the most the traversal loop can use without spilling
constexpr int kTraceMinBlocks = GB_TRACE_MIN_BLOCKS;
constexpr int kShadeBlock = 128;
constexpr int kMaxStack = 64;      // entries per thread; the reference's todo[64] per level
constexpr size_t kMaxTraceSmem = 200 * 1024;
constexpr size_t kMaxWideSmem = 56 * 1024; // beyond this the 4-wide walk would run fewer than 4 CTAs per SM: walk pair-wise
constexpr int kCtrStride = 16;     // counters per bounce
// per-bounce counters; the *_HEAD cursors are 64-bit (two slots, 8-byte aligned).  Row b holds what bounce b's kernels
// consume, except C_SHADOW / C_SHADOW_HEAD: the shadow segments of bounce b are counted in row b + 1, next to the
// C_EXTEND its shade kernels fill at the same time (k_shade reserves both with one 64-bit atomic)
enum { C_EXTEND = 0, C_SHADOW = 1, C_MAT0 = 2 /* .. 5: one per GB_MAT_* */, C_EXTEND_HEAD = 6, C_SHADOW_HEAD = 8, C_AO_HEAD = 10 };
// traversal statistics are kept apart for closest-hit and any-hit walks (S_ANY_BASE + ...)
enum { S_RAYS_CLOSEST = 0, S_RAYS_ANY = 1, S_NODES = 2, S_PRIMS = 3, S_INSTS = 4, S_SAMPLES = 5, S_ANY_BASE = 8, S_COUNT = 16 };

struct PathState {
    float4* rayO;    // o.xyz, mint
    float4* rayD;    // d.xyz, unused
    float4* hit;     // t, b1, b2, unused
    int2* hitId;     // instance slot (-1 = miss), triangle slot
    float4* thr;     // throughput rgb
    float4* L;       // accumulated radiance rgb, AO: unoccluded count in w
    float4* pend;    // pending emission weight rgb, __int_as_float(light) in w
    float4* shO;     // shadow ray o.xyz, mint
    float4* shD;     // shadow ray d.xyz, maxt
    float4* shC;     // contribution rgb, __int_as_float(path)
    unsigned int* qExtend[2];
    unsigned int* qMat[GB_MAT_COUNT];
    unsigned int* aoCount;
};

struct WaveParams {
    unsigned int nPaths;
    int y0;          // first sample-range row of the wave
    int width;       // sample-range width
    int rows;
    int sppBegin;    // first sample index of this call
    int nSpp;        // samples per pixel in this wave
    int sppTotal;    // samples per pixel of the whole job
    int root;        // sqrt(sppTotal): image-plane strata per axis
    int maxDepth;
    int aoSamples;
    int aoRoot;
};

__device__ __forceinline__ unsigned long long sampleIdOf(const DeviceScene& sc, const WaveParams& wp, unsigned int i,
    int* px, int* py, int* s, unsigned long long* pixel = nullptr) {
    unsigned int pix = i / (unsigned int)wp.nSpp;
    unsigned int k = i - pix * (unsigned int)wp.nSpp;
    unsigned int row = pix / (unsigned int)wp.width;
    unsigned int col = pix - row * (unsigned int)wp.width;
    *px = sc.sx0 + (int)col;
    *py = wp.y0 + (int)row;
    *s = wp.sppBegin + (int)k;
    unsigned long long pixelIndex = (unsigned long long)(*py - sc.sy0) * (unsigned long long)wp.width + col;
    if (pixel) *pixel = pixelIndex;
    return pixelIndex * (unsigned long long)wp.sppTotal + (unsigned long long)(*s);
}

// Sampler::requestSamples' image-plane stratification (GoblinSampler.cpp:
// 142-143,192-195 with stratifiedUniform2D(buffer, 1)): sample s of a pixel sits
// in cell (s % root, s / root) of a root x root jittered grid.
__device__ __forceinline__ void imagePosition(const WaveParams& wp, int px, int py, int s, float4 u, bool tableDriven,
    float* imageX, float* imageY) {
    if (tableDriven) { *imageX = u.x; *imageY = u.y; return; }
    float sub = 1.0f / (float)wp.root;
    int cx = s % wp.root, cy = s / wp.root;
    *imageX = (float)px + ((float)cx + u.x) * sub;
    *imageY = (float)py + ((float)cy + u.y) * sub;
}

// Shared tail of the traversal kernels: ray count and (optional) traversal statistics, one
// 64-bit atomic per warp.
template <bool ANY, bool STATS>
__device__ __forceinline__ void flushStats(unsigned long long* stats, unsigned int done, const TraceStats& tsIn) {
    const unsigned int lane = threadIdx.x & 31;
    TraceStats ts = tsIn;
    unsigned int total = done;
    for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
    if (lane == 0 && total) atomicAdd(stats + (ANY ? S_RAYS_ANY : S_RAYS_CLOSEST), (unsigned long long)total);
    if (STATS) {
        for (int off = 16; off > 0; off >>= 1) {
            ts.nodes += __shfl_xor_sync(0xffffffffu, ts.nodes, off);
            ts.prims += __shfl_xor_sync(0xffffffffu, ts.prims, off);
            ts.insts += __shfl_xor_sync(0xffffffffu, ts.insts, off);
        }
        if (lane == 0) {
            atomicAdd(stats + (ANY ? S_ANY_BASE : 0) + S_NODES, (unsigned long long)ts.nodes);
            atomicAdd(stats + (ANY ? S_ANY_BASE : 0) + S_PRIMS, (unsigned long long)ts.prims);
            atomicAdd(stats + (ANY ? S_ANY_BASE : 0) + S_INSTS, (unsigned long long)ts.insts);
        }
    }
}

#define GB_TRACE_SMEM(stackEntries)                                                    \
    extern __shared__ uint2 s_dyn[];                                                   \
    uint2* s_stack = s_dyn;                                                            \
    float* s_ray = reinterpret_cast<float*>(s_dyn + (size_t)(stackEntries) * blockDim.x)

// ------------------------------------------------------------ stand-alone trace
template <bool ANY>
struct TracePolicy {
    const gb_ray* rays;
    gb_hit* hits;
    unsigned char* occluded;
    const DeviceScene* sc;
    __device__ __forceinline__ bool fetch(unsigned long long i, float3* o, float3* d, float* mint, float* maxt) {
        const float4* r = reinterpret_cast<const float4*>(rays + i);
        float4 a = __ldg(r), b = __ldg(r + 1);
        *o = make3(a.x, a.y, a.z);
        *d = make3(a.w, b.x, b.y);
        *mint = b.z;
        *maxt = b.w;
        return true;
    }
    __device__ __forceinline__ void finish(bool done, unsigned long long i, bool found, const HitRec& h) {
        if (!done) return;
        if (ANY) {
            occluded[i] = found ? 1 : 0;
            return;
        }
        gb_hit out;
        if (found) {
            int4 sh = __ldg(sc->instShade + h.inst);
            int4 info = __ldg(sc->instInfo + h.inst);
            out.t = h.t;
            out.eps = 1e-3f * h.t;
            out.inst = sh.x;
            out.prim = 0;
            if (info.x == GB_GEOM_MESH) {
                out.prim = (int)__float_as_uint(__ldg(sc->triRec + kTriRecVec4 * (size_t)(info.z + h.prim) + 2).y);
            }
        } else {
            out.t = 0.0f; out.eps = 0.0f; out.inst = -1; out.prim = -1;
        }
        *reinterpret_cast<float4*>(hits + i) = *reinterpret_cast<float4*>(&out);
    }
};

// MODE of every traversal kernel: which walk of the reference's tree it runs (traverse.cuh)
enum { WALK_WIDE = 0,   // 4-wide nodes (GB_TRACE_WIDE)
       WALK_PAIR = 1,   // pair nodes, box tests exactly where the reference evaluates them: the default
       WALK_STATS = 2 };// the pair walk with the traversal counters on
#define GB_WALK_FLAGS(MODE) constexpr bool STATS = (MODE) == WALK_STATS, WIDE = (MODE) == WALK_WIDE
// the counting walk is not timed: one CTA per SM less buys it the registers to compile without spills (k_extend<STATS>
// spilled a predicate pair at 72 registers; see the fault hunt in DESIGN.md 4 for why no spill is left in these kernels)
constexpr int traceMinBlocks(int mode) {
#if defined(GB_STATS_SAME_BOUNDS) && GB_STATS_SAME_BOUNDS // the fault hunt's control: the counting walk at 7 CTAs / SM again, spill included
    return mode == WALK_WIDE ? GB_WIDE_MIN_BLOCKS : kTraceMinBlocks;
#else
    return mode == WALK_WIDE ? GB_WIDE_MIN_BLOCKS : (mode == WALK_STATS ? kTraceMinBlocks - 1 : kTraceMinBlocks);
#endif
}

template <bool ANY, int MODE>
__global__ void __launch_bounds__(kTraceBlock, traceMinBlocks(MODE))
k_trace(DeviceScene sc, const gb_ray* __restrict__ rays, unsigned long long n, gb_hit* __restrict__ hits,
    unsigned char* __restrict__ occluded, unsigned long long* head, unsigned long long* stats, int stackEntries) {
    GB_WALK_FLAGS(MODE);
    GB_TRACE_SMEM(stackEntries);
    TracePolicy<ANY> pol{rays, hits, occluded, &sc};
    TraceStats ts{0, 0, 0};
    unsigned int done = 0;
    persistentTrace<ANY, STATS, WIDE>(sc, pol, n, head, s_stack, s_ray, ts, &done, stackEntries);
    flushStats<ANY, STATS>(stats, done, ts);
}

// ------------------------------------------------------------------- raygen
__global__ void k_raygen(DeviceScene sc, PathState ps, WaveParams wp, SampleSource src, unsigned int* ctr,
    unsigned long long* stats) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        ctr[C_EXTEND] = wp.nPaths;
        atomicAdd(stats + S_SAMPLES, (unsigned long long)wp.nPaths);
    }
    if (i >= wp.nPaths) return;
    int px, py, s;
    unsigned long long pixel;
    unsigned long long id = sampleIdOf(sc, wp, i, &px, &py, &s, &pixel);
    float4 u = src.cameraBlock(id, i, pixel, (unsigned int)s, sc.camera.lens_radius > 0.0f);
    float imageX, imageY;
    imagePosition(wp, px, py, s, u, src.table != nullptr, &imageX, &imageY);
    float3 o, d;
    cameraRay(sc, imageX, imageY, u.z, u.w, &o, &d);
    ps.rayO[i] = make_float4(o.x, o.y, o.z, cameraMint(sc)); // ray->mint
    ps.rayD[i] = make_float4(d.x, d.y, d.z, 0.0f);
    ps.thr[i] = make_float4(1.0f, 1.0f, 1.0f, 1.0f); // w: no real bounce yet (firstBounce)
    ps.L[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    ps.pend[i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
    if (ps.aoCount) ps.aoCount[i] = 0u;
}

// ------------------------------------------------------------------- extend
// Scene::intersect for every queued path, then bin the hits by material.
struct ExtendPolicy {
    const DeviceScene* sc;
    PathState ps;
    const unsigned int* queue;
    unsigned int* ctr;
    int fixedBin; // >= 0: every hit goes to this bin (AO; scenes with one material type), no lookup
    unsigned int path; // per lane: the path the lane's ray belongs to
    __device__ __forceinline__ bool fetch(unsigned long long j, float3* o, float3* d, float* mint, float* maxt) {
        path = queue ? __ldg(queue + j) : (unsigned int)j;
        float4 a = ps.rayO[path], b = ps.rayD[path];
        *o = make3(a.x, a.y, a.z);
        *d = make3(b.x, b.y, b.z);
        *mint = a.w;
        *maxt = INFINITY;
        return true;
    }
    __device__ __forceinline__ void finish(bool done, unsigned long long, bool found, const HitRec& h) {
        const unsigned int lane = threadIdx.x & 31;
        int bin = -1;
        if (done) {
            ps.hit[path] = make_float4(h.t, h.b1, h.b2, 0.0f);
            ps.hitId[path] = make_int2(h.inst, h.prim);
            if (found) {
                if (fixedBin >= 0) bin = fixedBin;
                else {
                    int mat = __ldg(sc->instShade + h.inst).z;
                    bin = __float_as_int(__ldg(&sc->materials[mat].kdType).w);
                }
            }
        }
        // warp-aggregated append to the material queues
#pragma unroll
        for (int m = 0; m < GB_MAT_COUNT; ++m) {
            unsigned int mask = __ballot_sync(0xffffffffu, bin == m);
            if (mask) {
                unsigned int leader = __ffs(mask) - 1;
                unsigned int start = 0;
                if (lane == leader) start = atomicAdd(ctr + C_MAT0 + m, __popc(mask));
                start = __shfl_sync(0xffffffffu, start, leader);
                if (bin == m) ps.qMat[m][start + __popc(mask & ((1u << lane) - 1))] = path;
            }
        }
    }
};

template <int MODE>
__global__ void __launch_bounds__(kTraceBlock, traceMinBlocks(MODE))
k_extend(DeviceScene sc, PathState ps, const unsigned int* __restrict__ queue, unsigned int* ctr, int fixedBin,
    unsigned long long* stats, int stackEntries) {
    GB_WALK_FLAGS(MODE);
    GB_TRACE_SMEM(stackEntries);
    ExtendPolicy pol{&sc, ps, queue, ctr, fixedBin, 0u};
    TraceStats ts{0, 0, 0};
    unsigned int done = 0;
    persistentTrace<false, STATS, WIDE>(sc, pol, (unsigned long long)ctr[C_EXTEND],
        reinterpret_cast<unsigned long long*>(ctr + C_EXTEND_HEAD), s_stack, s_ray, ts, &done, stackEntries);
    flushStats<false, STATS>(stats, done, ts);
}

// ------------------------------------------------------------------- shadow
// Scene::occluded for every queued shadow segment; unoccluded segments add
// their (already weighted) contribution to the owning path.
struct ShadowPolicy {
    PathState ps;
    __device__ __forceinline__ bool fetch(unsigned long long j, float3* o, float3* d, float* mint, float* maxt) {
        float4 a = ps.shO[j], b = ps.shD[j];
        *o = make3(a.x, a.y, a.z);
        *d = make3(b.x, b.y, b.z);
        *mint = a.w;
        *maxt = b.w;
        return true;
    }
    __device__ __forceinline__ void finish(bool done, unsigned long long j, bool found, const HitRec&) {
        if (!done || found) return;
        float4 c = ps.shC[j];
        unsigned int i = (unsigned int)__float_as_int(c.w);
        float4 L = ps.L[i]; // one shadow segment per path per bounce: no race
        L.x += c.x; L.y += c.y; L.z += c.z;
        ps.L[i] = L;
    }
};

template <int MODE>
__global__ void __launch_bounds__(kTraceBlock, traceMinBlocks(MODE))
k_shadow(DeviceScene sc, PathState ps, unsigned int* ctr, unsigned long long* stats, int stackEntries) {
    GB_WALK_FLAGS(MODE);
    GB_TRACE_SMEM(stackEntries);
    ShadowPolicy pol{ps};
    TraceStats ts{0, 0, 0};
    unsigned int done = 0;
    persistentTrace<true, STATS, WIDE>(sc, pol, (unsigned long long)ctr[C_SHADOW],
        reinterpret_cast<unsigned long long*>(ctr + C_SHADOW_HEAD), s_stack, s_ray, ts, &done, stackEntries);
    flushStats<true, STATS>(stats, done, ts);
}

// -------------------------------------------------------------------- shade
// resident CTAs the Lambert / Blinn variants are compiled for: 80 registers and a 72-byte spill frame.  Measured round 2
// (profiles/r02/call17_stdout.txt): 5 CTAs (95 registers, no spills) the same within 2 %, 7 CTAs (72 registers) 3 - 10 % slower.
constexpr int kShadeHeavyBlocks = 6;

// One bounce of PathTracer::Li (GoblinPathtracer.cpp:76-172) for the paths whose
// hit carries material MAT.  `bounce` is the reference's loop variable; with
// emissionOnly the kernel only resolves the pending BSDF-sampled emission
// (the reference's trace #4 of the last iteration).
template <int MAT, bool ML, bool TEX>
__global__ void __launch_bounds__(kShadeBlock, (MAT == GB_MAT_LAMBERT || MAT == GB_MAT_BLINN) ? kShadeHeavyBlocks : 8)
k_shade(DeviceScene sc, PathState ps, WaveParams wp, SampleSource src, int bounce, int emissionOnly,
    unsigned int* ctr, unsigned int* ctrNext, unsigned int* qNext) {
    const unsigned int n = ctr[C_MAT0 + MAT];
    const unsigned int lane = threadIdx.x & 31;
    const unsigned int stride = gridDim.x * blockDim.x;
    // The kernel is a chain of dependent fetches (queue entry -> path state -> instance -> triangle -> material) at 24
    // warps per SM, and ncu puts a fifth of its stall samples on the first two links (profiles/r02/r02j_*).  The queue
    // entries are therefore read two trips ahead: when a warp starts a trip its entry is already in a register.  Asking
    // the next trip's path state into L2 as well (prefetch.global.L2, five lines per path) made the kernel 13 - 20 %
    // slower on every scene (profiles/r02/call17_stdout.txt) and is not done.
    const unsigned int jFirst = ((blockIdx.x * blockDim.x + threadIdx.x) & ~31u) + lane;
    unsigned int iCur = jFirst < n ? __ldg(ps.qMat[MAT] + jFirst) : 0u;
    unsigned int iNext = jFirst + stride < n ? __ldg(ps.qMat[MAT] + jFirst + stride) : 0u;
    for (unsigned int j0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; j0 < n; j0 += stride) {
        unsigned int j = j0 + lane;
        bool alive = false;   // continues to the next extend
        bool shadow = false;  // emits a shadow segment
        const unsigned int i = iCur;
        {
            const unsigned int j2 = j + 2u * stride; // a wave holds at most 2^28 paths (gb_set_wave_paths): no wrap-around
            iCur = iNext;
            iNext = j2 < n ? __ldg(ps.qMat[MAT] + j2) : 0u;
        }
        float3 shO = make3(0, 0, 0), shD = make3(0, 0, 0), shC = make3(0, 0, 0);
        float shMint = 0.0f, shMaxt = 0.0f;
        if (j < n) {
            float4 ro = ps.rayO[i], rd = ps.rayD[i], hv = ps.hit[i];
            int2 hid = ps.hitId[i];
            HitRec h;
            h.t = hv.x; h.b1 = hv.y; h.b2 = hv.z; h.inst = hid.x; h.prim = hid.y;
            float3 o = make3(ro.x, ro.y, ro.z), d = make3(rd.x, rd.y, rd.z);
            Frag fr = buildFragment(sc, h, o, d);
            float3 wo = -d;
            // emitted radiance seen through this segment: the only thing that touches L here, so the
            // 32 bytes of its read-modify-write are spent on emitter hits alone
            if (fr.areaLight >= 0) {
                float4 lc = __ldg(&sc.lights[fr.areaLight].colorType);
                bool facing = dot3(fr.n, wo) > 0.0f; // AreaLight::L
                if (bounce == 0) { // Li += intersection.Le(-ray.d), GoblinPathtracer.cpp:67
                    if (facing) {
                        float4 Lacc = ps.L[i];
                        Lacc.x += lc.x; Lacc.y += lc.y; Lacc.z += lc.z;
                        ps.L[i] = Lacc;
                    }
                } else if (!(TEX && sc.matMask)) { // with masks the MIS ray is traced on its own (k_mis_mask)
                    float4 pd = ps.pend[i]; // BSDF-sampled MIS term, GoblinPathtracer.cpp:148-155
                    if (__float_as_int(pd.w) == fr.areaLight && facing) {
                        float4 Lacc = ps.L[i];
                        Lacc.x += pd.x * lc.x; Lacc.y += pd.y * lc.y; Lacc.z += pd.z * lc.z;
                        ps.L[i] = Lacc;
                    }
                }
            }
            if (!emissionOnly) {
                int px, py, s;
                unsigned long long pixel;
                unsigned long long id = sampleIdOf(sc, wp, i, &px, &py, &s, &pixel);
                // which of the bounce's dimensions this material and the scene's lights can read at all: light position only
                // with area / image lights and a BSDF that is not a delta, its component only with mesh emitters, the BSDF
                // component for glass (and masks), its direction for Lambert / Blinn, the light pick with several lights
                unsigned int need = sc.nLights > 1u ? 1u << DIM_PICK : 0u;
                if (MAT == GB_MAT_LAMBERT || MAT == GB_MAT_BLINN) {
                    need |= 1u << DIM_BSDF_UV;
                    if (sc.hasAreaLight | sc.hasEnvLight) need |= 1u << DIM_LIGHT_UV;
                    if (ML) need |= 1u << DIM_LIGHT_COMP;
                }
                if (MAT == GB_MAT_TRANSPARENT || TEX) need |= 1u << DIM_BSDF_COMP;
                float4 uA, uB;
                src.bounceBlocks(id, i, pixel, (unsigned int)s, (unsigned int)bounce, need, &uA, &uB);
                float pickPdf;
                int li = pickLight(sc, uB.z, &pickPdf);
                float4 tv = ps.thr[i];
                float3 thr = make3(tv.x, tv.y, tv.z);
                float eps = 1e-3f * h.t;
                const DeviceMaterial& mat = sc.materials[fr.material];
                DeviceMaterial m;
                m.kdType = __ldg(&mat.kdType);
                m.ktEta = __ldg(&mat.ktEta);
                if (TEX) { // some material slot of this scene is a procedural texture
                    float imageX = 0.0f, imageY = 0.0f;
                    float4 u0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (bounce == 0) { // only the camera ray carries differentials
                        u0 = src.cameraBlock(id, i, pixel, (unsigned int)s, sc.camera.lens_radius > 0.0f);
                        imagePosition(wp, px, py, s, u0, src.table != nullptr, &imageX, &imageY);
                    }
                    // bump / normal maps rewrite the frame: it comes back by value
                    float3 nOut = fr.n, dpduOut = fr.dpdu;
                    applyTextures(sc, fr.material, MAT, h, o, d, fr, bounce == 0, imageX, imageY, u0.z, u0.w, &m, &nOut, &dpduOut);
                    fr.n = nOut;
                    fr.dpdu = dpduOut;
                }
                // Mask around this material (MaskMaterial, GoblinMaterial.cpp:747-811): alpha scales the
                // masked BSDF, 1 - alpha goes straight through
                bool masked = false;
                MaskEval me;
                me.alpha = 1.0f; me.tc = make3(0.0f, 0.0f, 0.0f);
                if (TEX && sc.matMask && __ldg(sc.matMask + 2 * (size_t)fr.material).x) {
                    masked = true;
                    me = maskAt(sc, fr.material, h, o, d, fr);
                }
                if (MAT == GB_MAT_LAMBERT || MAT == GB_MAT_BLINN) { // specular BSDFs evaluate to black: no light sample survives
                    LightSampleResult ls = sampleLight<ML>(sc, li, fr.p, eps, uA.x, uA.y, uA.z);
                    if (!isBlack(ls.L) && ls.pdf > 0.0f) {
                        float3 f = MAT == GB_MAT_LAMBERT ? lambertEval(m, fr.n, wo, ls.wi) : blinnEval(m, fr.n, wo, ls.wi);
                        if (TEX && masked) f = f * me.alpha;
                        if (!isBlack(f)) {
                            float3 c = mul3(f, ls.L) * absdot3(fr.n, ls.wi);
                            if (!ls.delta) {
                                float bsdfPdf = MAT == GB_MAT_LAMBERT ? lambertPdf(fr.n, wo, ls.wi) : blinnPdf(m, fr.n, wo, ls.wi);
                                if (TEX && masked) bsdfPdf = me.alpha * bsdfPdf;
                                c = c * powerHeuristic(ls.pdf, bsdfPdf);
                            }
                            c = div3(c, ls.pdf);
                            shC = div3(mul3(thr, c), pickPdf);
                            shO = fr.p; shD = ls.wi; shMint = eps; shMaxt = ls.maxt;
                            shadow = true;
                        }
                    }
                }
                BsdfSample bs;
                bool nullSampled = false;
                if (TEX && masked && !(uA.w < me.alpha)) { // the index-matched pass-through (BSDFnullptr)
                    bs.f = (1.0f - me.alpha) * me.tc;
                    bs.wi = -normalize3(wo);
                    bs.pdf = 1.0f - me.alpha;
                    bs.specular = false;
                    nullSampled = true;
                    shadow = false; // the reference `continue`s: this bounce's direct light is dropped
                } else {
                    bs = sampleBsdf(m, MAT, fr, wo, uA.w, uB.x, uB.y);
                    if (TEX && masked) { bs.f = bs.f * me.alpha; bs.pdf *= me.alpha; }
                }
                float4 pendOut = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
                if (!nullSampled && !isBlack(bs.f) && bs.pdf > 0.0f) {
                    float fWeight = 1.0f;
                    if (!bs.specular) fWeight = powerHeuristic(bs.pdf, lightPdf<ML>(sc, li, fr.p, bs.wi));
                    int ltype = __float_as_int(__ldg(&sc.lights[li].colorType).w);
                    if (ltype == GB_LIGHT_AREA) {
                        float3 w = div3(bs.f * absdot3(bs.wi, fr.n) * fWeight, bs.pdf);
                        float3 pw = div3(mul3(thr, w), pickPdf);
                        pendOut = make_float4(pw.x, pw.y, pw.z, __int_as_float(li));
                    } else if (ML && ltype == GB_LIGHT_IBL) { // f * tr * light->Le(r) * fWeight / bsdfPdf: no cosine
                        float3 w = div3(bs.f * fWeight, bs.pdf);
                        float3 pw = div3(mul3(thr, w), pickPdf);
                        pendOut = make_float4(pw.x, pw.y, pw.z, __int_as_float(li));
                    }
                }
                if (!(isBlack(bs.f) || bs.pdf == 0.0f)) {
                    // thr.w: the path has only punched through masks so far (the reference's firstBounce)
                    float first = 0.0f;
                    if (TEX && nullSampled) {
                        thr = mul3(thr, div3(bs.f, bs.pdf)); // throughput *= f / bsdfPdf, no cosine
                        first = tv.w;
                    } else {
                        float3 w = div3(bs.f * absdot3(bs.wi, fr.n), bs.pdf);
                        thr = mul3(thr, w);
                    }
                    ps.thr[i] = make_float4(thr.x, thr.y, thr.z, first);
                    if (sc.hasAreaLight | sc.hasEnvLight) ps.pend[i] = pendOut; // nobody reads it otherwise
                    ps.rayO[i] = make_float4(fr.p.x, fr.p.y, fr.p.z, eps);
                    ps.rayD[i] = make_float4(bs.wi.x, bs.wi.y, bs.wi.z, 0.0f);
                    alive = true;
                }
            }
        }
        // warp-aggregated appends to the next extend queue and to the shadow queue.  Their two counters sit side by side
        // in the next bounce's row (C_EXTEND, C_SHADOW: one aligned 64-bit word), so one atomic reserves both ranges and the
        // warp waits for one round trip to L2 instead of two.  Neither half can carry into the other (at most 2^28 paths a wave).
        const unsigned int aliveMask = __ballot_sync(0xffffffffu, alive);
        const unsigned int shadowMask = __ballot_sync(0xffffffffu, shadow);
        if (aliveMask | shadowMask) {
            unsigned long long start = 0;
            if (lane == 0) {
                start = atomicAdd(reinterpret_cast<unsigned long long*>(ctrNext + C_EXTEND),
                    (unsigned long long)__popc(aliveMask) | ((unsigned long long)__popc(shadowMask) << 32));
            }
            start = __shfl_sync(0xffffffffu, start, 0);
            if (alive) qNext[(unsigned int)start + __popc(aliveMask & ((1u << lane) - 1))] = i;
            if (shadow) {
                unsigned int k = (unsigned int)(start >> 32) + __popc(shadowMask & ((1u << lane) - 1));
                ps.shO[k] = make_float4(shO.x, shO.y, shO.z, shMint);
                ps.shD[k] = make_float4(shD.x, shD.y, shD.z, shMaxt);
                ps.shC[k] = make_float4(shC.x, shC.y, shC.z, __int_as_float((int)i));
            }
        }
    }
}

// ------------------------------------------------------------- environment
// Scenes with an image based light: what a ray that left the scene picks up.  Camera rays add
// Scene::evalEnvironmentLight (GoblinScene.cpp:89-95: the sum of Le over all lights,
// GoblinPathtracer.cpp:61-65); BSDF-sampled rays add the MIS-weighted term of the light that was
// picked for their bounce (GoblinPathtracer.cpp:156-160), whose weight the shade kernel left in
// `pend`.  One thread per path of the extend queue that has just been traced.
__global__ void k_miss(DeviceScene sc, PathState ps, const unsigned int* __restrict__ queue, const unsigned int* ctr,
    int bounce) {
    const unsigned int n = ctr[C_EXTEND];
    for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const unsigned int i = queue ? __ldg(queue + j) : j;
        if (ps.hitId[i].x >= 0) continue;
        const float4 rd = ps.rayD[i];
        const float3 d = make3(rd.x, rd.y, rd.z);
        float4 L = ps.L[i];
        if (bounce == 0) {
            for (unsigned int l = 0; l < sc.nLights; ++l) {
                if (__float_as_int(__ldg(&sc.lights[l].colorType).w) != GB_LIGHT_IBL) continue;
                const float3 le = iblLe(sc, sc.lights[l], d);
                L.x += le.x; L.y += le.y; L.z += le.z;
            }
        } else if (sc.matMask) {
            // masks: the MIS term was settled by k_mis_mask; a path that only punched through masks
            // since the camera still sees the environment (firstBounce, GoblinPathtracer.cpp:124-128)
            const float4 tv = ps.thr[i];
            if (tv.w == 0.0f) continue;
            for (unsigned int l = 0; l < sc.nLights; ++l) {
                if (__float_as_int(__ldg(&sc.lights[l].colorType).w) != GB_LIGHT_IBL) continue;
                const float3 le = iblLe(sc, sc.lights[l], d);
                L.x += tv.x * le.x; L.y += tv.y * le.y; L.z += tv.z * le.z;
            }
        } else {
            const float4 pd = ps.pend[i];
            const int li = __float_as_int(pd.w);
            if (li < 0 || __float_as_int(__ldg(&sc.lights[li].colorType).w) != GB_LIGHT_IBL) continue;
            const float3 le = iblLe(sc, sc.lights[li], d);
            L.x += pd.x * le.x; L.y += pd.y * le.y; L.z += pd.z * le.z;
        }
        ps.L[i] = L;
    }
}

// --------------------------------------------------------------------- masks
// Scenes with a Mask material (mask.cuh).  Shadow segments: occluded by OPAQUE primitives only, then
// attenuated by the not-opaque ones (GoblinPathtracer.cpp:96-99).
__global__ void k_shadow_mask(DeviceScene sc, PathState ps, const unsigned int* ctr, unsigned long long* stats) {
    const unsigned int n = ctr[C_SHADOW];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + S_RAYS_ANY, (unsigned long long)n);
    for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const float4 a = ps.shO[j], b = ps.shD[j];
        const float3 o = make3(a.x, a.y, a.z), d = make3(b.x, b.y, b.z);
        if (simpleAny(sc, o, d, a.w, b.w, FILTER_OPAQUE)) continue;
        const float3 tr = evalAttenuation(sc, o, d, a.w, b.w);
        const float4 c = ps.shC[j];
        const unsigned int i = (unsigned int)__float_as_int(c.w);
        float4 L = ps.L[i];
        L.x += c.x * tr.x; L.y += c.y * tr.y; L.z += c.z * tr.z;
        ps.L[i] = L;
    }
}
// The MIS ray of a bounce (GoblinPathtracer.cpp:142-161): closest OPAQUE hit, attenuation up to it, then
// the picked light's emission there -- or the environment map when nothing opaque is hit.  One thread per
// path that continues; `pend` holds the weight the shade kernel prepared.
__global__ void k_mis_mask(DeviceScene sc, PathState ps, const unsigned int* __restrict__ queue, const unsigned int* ctrNext,
    unsigned long long* stats) {
    const unsigned int n = ctrNext[C_EXTEND];
    for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const unsigned int i = __ldg(queue + j);
        const float4 pd = ps.pend[i];
        const int li = __float_as_int(pd.w);
        if (li < 0) continue;
        const float4 ro = ps.rayO[i], rd = ps.rayD[i];
        const float3 o = make3(ro.x, ro.y, ro.z), d = make3(rd.x, rd.y, rd.z);
        atomicAdd(stats + S_RAYS_CLOSEST, 1ull);
        HitRec h;
        const bool found = simpleClosest(sc, o, d, ro.w, INFINITY, FILTER_OPAQUE, &h);
        const int ltype = __float_as_int(__ldg(&sc.lights[li].colorType).w);
        float3 le = make3(0.0f, 0.0f, 0.0f);
        if (found) {
            const Frag fr = buildFragment(sc, h, o, d);
            if (fr.areaLight == li && ltype == GB_LIGHT_AREA && dot3(fr.n, -d) > 0.0f) {
                const float4 lc = __ldg(&sc.lights[li].colorType);
                le = make3(lc.x, lc.y, lc.z);
            }
        } else if (ltype == GB_LIGHT_IBL) {
            le = iblLe(sc, sc.lights[li], d);
        }
        if (le.x == 0.0f && le.y == 0.0f && le.z == 0.0f) continue;
        const float3 tr = evalAttenuation(sc, o, d, ro.w, found ? h.t : INFINITY);
        float4 L = ps.L[i];
        L.x += pd.x * le.x * tr.x; L.y += pd.y * le.y * tr.y; L.z += pd.z * le.z * tr.z;
        ps.L[i] = L;
    }
}

// ----------------------------------------------------------------------- AO
// AORenderer::Li (GoblinAO.cpp:12-37): work item = (hit path, occlusion ray).
// The hit frame of every AO path, once (instead of once per occlusion ray): position + epsilon,
// tangent, bitangent, normal go to path-state arrays the AO integrator does not otherwise use.
template <bool TEX> // TEX: the scene has textured materials, i.e. possibly bump / normal maps
__global__ void k_ao_frames(DeviceScene sc, PathState ps, WaveParams wp, SampleSource src, const unsigned int* ctr) {
    const unsigned int n = ctr[C_MAT0];
    for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        unsigned int i = __ldg(ps.qMat[0] + q);
        float4 ro = ps.rayO[i], rd = ps.rayD[i], hv = ps.hit[i];
        int2 hid = ps.hitId[i];
        HitRec h;
        h.t = hv.x; h.b1 = hv.y; h.b2 = hv.z; h.inst = hid.x; h.prim = hid.y;
        Frag fr = buildFragment(sc, h, make3(ro.x, ro.y, ro.z), make3(rd.x, rd.y, rd.z));
        if (TEX) { // bump / normal maps
            float3 nOut = fr.n, dpduOut = fr.dpdu;
            perturbOnly(sc, h, make3(ro.x, ro.y, ro.z), make3(rd.x, rd.y, rd.z), fr, &nOut, &dpduOut);
            fr.n = nOut;
            fr.dpdu = dpduOut;
        }
        ShadeFrame sf = makeFrame(fr);
        ps.shO[i] = make_float4(fr.p.x, fr.p.y, fr.p.z, 1e-3f * h.t); // Ray(fragment.getPosition(), dir, epsilon)
        int px, py, s;
        unsigned long long pixel;
        sampleIdOf(sc, wp, i, &px, &py, &s, &pixel);
        // .w: the sub-cell this camera sample takes in every AO direction's stratum (SampleSource::aoCell)
        ps.thr[i] = make_float4(sf.t.x, sf.t.y, sf.t.z, __uint_as_float(src.aoCell(pixel, (unsigned int)s)));
        ps.pend[i] = make_float4(sf.b.x, sf.b.y, sf.b.z, 0.0f);
        ps.shD[i] = make_float4(sf.n.x, sf.n.y, sf.n.z, 0.0f);
    }
}

struct AOPolicy {
    const DeviceScene* sc;
    PathState ps;
    WaveParams wp;
    SampleSource src;
    unsigned int path;
    __device__ __forceinline__ bool fetch(unsigned long long j, float3* o, float3* d, float* mint, float* maxt) {
        unsigned int q = (unsigned int)(j / (unsigned int)wp.aoSamples);
        unsigned int a = (unsigned int)(j - (unsigned long long)q * (unsigned int)wp.aoSamples);
        unsigned int i = __ldg(ps.qMat[0] + q);
        path = i;
        float4 po = ps.shO[i], ft = ps.thr[i], fb = ps.pend[i], fn = ps.shD[i];
        int px, py, s;
        unsigned long long pixel;
        unsigned long long id = sampleIdOf(*sc, wp, i, &px, &py, &s, &pixel);
        float2 u = src.aoPair(id, i, pixel, __float_as_uint(ft.w), a);
        if (!src.table) { // the reference stratifies the AO directions on a root x root grid
            float sub = 1.0f / (float)wp.aoRoot;
            u.x = ((float)(a % (unsigned int)wp.aoRoot) + u.x) * sub;
            u.y = ((float)(a / (unsigned int)wp.aoRoot) + u.y) * sub;
        }
        ShadeFrame sf;
        sf.t = make3(ft.x, ft.y, ft.z);
        sf.b = make3(fb.x, fb.y, fb.z);
        sf.n = make3(fn.x, fn.y, fn.z);
        *o = make3(po.x, po.y, po.z);
        *d = shadeToWorld(sf, uniformSampleHemisphere(u.x, u.y));
        *mint = po.w;
        *maxt = INFINITY;
        return true;
    }
    __device__ __forceinline__ void finish(bool done, unsigned long long, bool found, const HitRec&) {
        if (done && !found) atomicAdd(ps.aoCount + path, 1u);
    }
};

template <int MODE>
__global__ void __launch_bounds__(kTraceBlock, traceMinBlocks(MODE))
k_ao(DeviceScene sc, PathState ps, WaveParams wp, SampleSource src, unsigned int* ctr, unsigned long long* stats,
    int stackEntries) {
    GB_WALK_FLAGS(MODE);
    GB_TRACE_SMEM(stackEntries);
    AOPolicy pol{&sc, ps, wp, src, 0u};
    TraceStats ts{0, 0, 0};
    unsigned int done = 0;
    const unsigned long long n = (unsigned long long)ctr[C_MAT0] * (unsigned long long)wp.aoSamples;
    persistentTrace<true, STATS, WIDE>(sc, pol, n, reinterpret_cast<unsigned long long*>(ctr + C_AO_HEAD), s_stack, s_ray,
        ts, &done, stackEntries);
    flushStats<true, STATS>(stats, done, ts);
}

__global__ void k_ao_finish(PathState ps, WaveParams wp) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wp.nPaths) return;
    float v = 0.0f;
    if (ps.hitId[i].x >= 0) v = (float)ps.aoCount[i] / (float)wp.aoSamples; // (n - occluded) / n
    ps.L[i] = make_float4(v, v, v, 0.0f);
}

// --------------------------------------------------------------------- film
// ImageTile::addSample (GoblinFilm.cpp:61-90) + Film::mergeTile (:140-153) as a
// warp-per-sample-pixel gather: the samples of one sample-range pixel can only
// touch the (2 ceil(wx) + 1) x (2 ceil(wy) + 1) film pixels around it, so lane c
// owns candidate pixel c, the warp streams the pixel's samples 32 at a time
// (each lane loads one L and regenerates its image position from Philox, then
// the values are broadcast with shuffles) and every lane accumulates w * L and
// w for its own pixel in registers.  One 128-bit atomic per touched film pixel
// per sample pixel replaces 64 scalar atomics per sample.
// The inclusion test is the reference's: x0 = ceil(dx - w) <= x <= floor(dx + w)
// clamped to the crop window; for integer x that is dx - w <= x <= dx + w.
constexpr int kFilmBlock = 256;
#ifndef GB_FILM_LDS
#define GB_FILM_LDS 1
#endif
#ifndef GB_FILM_WARP
#define GB_FILM_WARP 1 // k_film_warp for candidate windows of at most 32 pixels
#endif

__global__ void __launch_bounds__(kFilmBlock)
k_film(DeviceScene sc, PathState ps, WaveParams wp, SampleSource src, float4* film, int radX, int radY,
    int invExact) {
    __shared__ float s_table[256];
#if GB_FILM_LDS
    __shared__ float4 s_stage[(kFilmBlock / 32) * 64];
#endif
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_table[k] = __ldg(sc.filterTable + k);
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31;
    const unsigned int warpsPerBlock = blockDim.x >> 5;
    const unsigned int nPix = (unsigned int)wp.width * (unsigned int)wp.rows;
    const int dX = 2 * radX + 1, dY = 2 * radY + 1;
    const int nCand = dX * dY;
    const int cropX1 = sc.xstart + sc.xcount - 1, cropY1 = sc.ystart + sc.ycount - 1;
    const float wX = sc.filterWidthX, wY = sc.filterWidthY;
    const float sX = 16.0f / wX, sY = 16.0f / wY; // exact when the width is a power of two
    for (unsigned int pix = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); pix < nPix;
         pix += gridDim.x * warpsPerBlock) {
        const unsigned int row = pix / (unsigned int)wp.width, col = pix - row * (unsigned int)wp.width;
        const int px = sc.sx0 + (int)col, py = wp.y0 + (int)row;
        const unsigned int base = pix * (unsigned int)wp.nSpp;
        for (int c0 = 0; c0 < nCand; c0 += 32) { // one pass per 32 candidate pixels (one pass for widths <= 2)
            const int c = c0 + (int)lane;
            const int cx = px - radX + c % dX, cy = py - radY + c / dX;
            const bool mine = c < nCand && cx >= sc.xstart && cx <= cropX1 && cy >= sc.ystart && cy <= cropY1;
            const float fx = (float)cx, fy = (float)cy;
            float aR = 0.0f, aG = 0.0f, aB = 0.0f, aW = 0.0f;
            for (int k0 = 0; k0 < wp.nSpp; k0 += 32) {
                const int k = k0 + (int)lane;
                float dImageX = 0.0f, dImageY = 0.0f;
                float4 L = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                bool ok = false;
                if (k < wp.nSpp) {
                    const unsigned int i = base + (unsigned int)k;
                    L = ps.L[i];
                    ok = !(L.x != L.x || L.y != L.y || L.z != L.z); // NaN samples are discarded, weight included
                    int qx, qy, s;
                    unsigned long long id = sampleIdOf(sc, wp, i, &qx, &qy, &s);
                    float4 u = src.block(id, i, 0);
                    float imageX, imageY;
                    imagePosition(wp, qx, qy, s, u, src.table != nullptr, &imageX, &imageY);
                    dImageX = imageX - 0.5f;
                    dImageY = imageY - 0.5f;
                }
                const unsigned int okMask = __ballot_sync(0xffffffffu, ok);
                const int cnt = min(32, wp.nSpp - k0);
#if GB_FILM_LDS
                // the 32 samples of this pass go through shared memory: two 128-bit broadcast reads per sample instead of
                // five shuffles (the loop is bound by the shuffle / shared-memory pipe, one warp instruction per clock per SM)
                float4* stage = s_stage + (threadIdx.x >> 5) * 64;
                __syncwarp();
                stage[2 * lane] = make_float4(dImageX, dImageY, 0.0f, 0.0f);
                stage[2 * lane + 1] = make_float4(L.x, L.y, L.z, 0.0f);
                __syncwarp();
#endif
                for (int j = 0; j < cnt; ++j) {
#if GB_FILM_LDS
                    if (!((okMask >> j) & 1u)) continue;
                    const float4 sp = stage[2 * j], sl = stage[2 * j + 1];
                    const float sx = sp.x, sy = sp.y, lr = sl.x, lg = sl.y, lb = sl.z;
#else
                    const float sx = __shfl_sync(0xffffffffu, dImageX, j);
                    const float sy = __shfl_sync(0xffffffffu, dImageY, j);
                    const float lr = __shfl_sync(0xffffffffu, L.x, j);
                    const float lg = __shfl_sync(0xffffffffu, L.y, j);
                    const float lb = __shfl_sync(0xffffffffu, L.z, j);
                    if (!((okMask >> j) & 1u)) continue;
#endif
                    if (mine && fx >= sx - wX && fx <= sx + wX && fy >= sy - wY && fy <= sy + wY) {
                        // FilterTable::evaluate: nearest lower entry of the 16 x 16 table
                        const float tx = invExact ? fabsf((fx - sx) * sX) : fabsf(16 * (fx - sx) / wX);
                        const float ty = invExact ? fabsf((fy - sy) * sY) : fabsf(16 * (fy - sy) / wY);
                        const int ix = min((int)floorf(tx), 15), iy = min((int)floorf(ty), 15);
                        const float w = s_table[iy * 16 + ix];
                        aR += w * lr; aG += w * lg; aB += w * lb; aW += w;
                    }
                }
            }
            if (mine && (aW != 0.0f || aR != 0.0f || aG != 0.0f || aB != 0.0f)) {
                atomicAdd(film + (size_t)cy * sc.xres + cx, make_float4(aR, aG, aB, aW));
            }
        }
    }
}

// The same gather for filters whose candidate window fits one warp ((2 ceil(wx) + 1) (2 ceil(wy) + 1) <= 32: every
// width up to 2, the reference's defaults included).  k_film spends ~50 warp instructions per sample, most of them on
// the inclusion test and its bookkeeping, repeated by all 32 lanes for every sample (ncu: 83 % of the SM's issue
// slots, profiles/r02).  Here the lane that loads a sample also computes the reference's pixel range
// [ceil(dx - w), floor(dx + w)] x [ceil(dy - w), floor(dy + w)] (GoblinFilm.cpp:66-69) once and publishes it as a 32-bit
// mask over the warp's candidate pixels (0 for a NaN sample), so a consumer lane tests one bit.  The weights, the order
// of the additions and therefore the film are those of k_film, bit for bit.
template <bool INV_EXACT>
__global__ void __launch_bounds__(kFilmBlock)
k_film_warp(DeviceScene sc, PathState ps, WaveParams wp, SampleSource src, float4* film, int radX, int radY) {
    __shared__ float s_table[256];
    __shared__ float4 s_stage[(kFilmBlock / 32) * 64];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_table[k] = __ldg(sc.filterTable + k);
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31;
    const unsigned int warpsPerBlock = blockDim.x >> 5;
    const unsigned int nPix = (unsigned int)wp.width * (unsigned int)wp.rows;
    const int dX = 2 * radX + 1, dY = 2 * radY + 1;
    const int rx = (int)lane % dX, ry = (int)lane / dX; // this lane's pixel inside the candidate window
    const int cropX1 = sc.xstart + sc.xcount - 1, cropY1 = sc.ystart + sc.ycount - 1;
    const float wX = sc.filterWidthX, wY = sc.filterWidthY;
    const float sX = 16.0f / wX, sY = 16.0f / wY; // exact when the width is a power of two
    float4* stage = s_stage + (threadIdx.x >> 5) * 64;
    // the table's shared-memory address, made opaque so that it lives in a register (the compiler otherwise rebuilds
    // it from the CTA's shared window, three instructions, at every lookup)
    unsigned int tableAt = (unsigned int)__cvta_generic_to_shared(s_table);
    asm volatile("" : "+r"(tableAt));
    for (unsigned int pix = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); pix < nPix;
         pix += gridDim.x * warpsPerBlock) {
        const unsigned int row = pix / (unsigned int)wp.width, col = pix - row * (unsigned int)wp.width;
        const int px = sc.sx0 + (int)col, py = wp.y0 + (int)row;
        const int wx0 = px - radX, wy0 = py - radY; // corner of the candidate window
        const unsigned int base = pix * (unsigned int)wp.nSpp;
        const int cx = wx0 + rx, cy = wy0 + ry;
        const bool mine = ry < dY && cx >= sc.xstart && cx <= cropX1 && cy >= sc.ystart && cy <= cropY1;
        const float fx = (float)cx, fy = (float)cy;
        float aR = 0.0f, aG = 0.0f, aB = 0.0f, aW = 0.0f;
        for (int k0 = 0; k0 < wp.nSpp; k0 += 32) {
            const int k = k0 + (int)lane;
            float dImageX = 0.0f, dImageY = 0.0f;
            float4 L = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            unsigned int inside = 0u; // candidate pixels this sample reaches
            if (k < wp.nSpp) {
                const unsigned int i = base + (unsigned int)k;
                L = ps.L[i];
                // sampleIdOf(i) without its divisions: the warp knows the pixel, the lane the sample
                const int s = wp.sppBegin + k;
                const unsigned long long id = ((unsigned long long)(py - sc.sy0) * (unsigned long long)wp.width + col) *
                        (unsigned long long)wp.sppTotal + (unsigned long long)s;
                float4 u = src.block(id, i, 0);
                float imageX, imageY;
                imagePosition(wp, px, py, s, u, src.table != nullptr, &imageX, &imageY);
                dImageX = imageX - 0.5f;
                dImageY = imageY - 0.5f;
                if (!(L.x != L.x || L.y != L.y || L.z != L.z)) { // NaN samples are discarded, weight included
                    // relative to the window, clipped to it (k_film's candidates are the window's pixels too)
                    const int x0 = max((int)ceilf(dImageX - wX) - wx0, 0), x1 = min((int)floorf(dImageX + wX) - wx0, dX - 1);
                    const int y0 = max((int)ceilf(dImageY - wY) - wy0, 0), y1 = min((int)floorf(dImageY + wY) - wy0, dY - 1);
                    if (x0 <= x1) {
                        const unsigned int span = ((2u << (x1 - x0)) - 1u) << x0;
                        for (int y = y0; y <= y1; ++y) inside |= span << (y * dX);
                    }
                }
            }
            const int cnt = min(32, wp.nSpp - k0);
            __syncwarp();
            stage[2 * lane] = make_float4(dImageX, dImageY, __uint_as_float(inside), 0.0f);
            stage[2 * lane + 1] = make_float4(L.x, L.y, L.z, 0.0f);
            __syncwarp();
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float4 sp = stage[2 * j];
                if ((__float_as_uint(sp.z) >> lane) & 1u) {
                    const float4 sl = stage[2 * j + 1];
                    // FilterTable::evaluate: nearest lower entry of the 16 x 16 table
                    const float tx = INV_EXACT ? fabsf((fx - sp.x) * sX) : fabsf(16 * (fx - sp.x) / wX);
                    const float ty = INV_EXACT ? fabsf((fy - sp.y) * sY) : fabsf(16 * (fy - sp.y) / wY);
                    const int ix = min((int)floorf(tx), 15), iy = min((int)floorf(ty), 15);
                    float w;
                    asm("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(tableAt + 4u * (unsigned int)(iy * 16 + ix)));
                    aR += w * sl.x; aG += w * sl.y; aB += w * sl.z; aW += w;
                }
            }
        }
        if (mine && (aW != 0.0f || aR != 0.0f || aG != 0.0f || aB != 0.0f)) {
            atomicAdd(film + (size_t)cy * sc.xres + cx, make_float4(aR, aG, aB, aW));
        }
    }
}

// ------------------------------------------------------------ image post-processing
// Film::writeImage's normalisation: Color::operator/(float) = multiply by 1 / weight.
__global__ void k_resolve(const float4* __restrict__ film, unsigned int n, float* __restrict__ rgb) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = film[i];
    const float inv = 1.0f / p.w;
    rgb[3 * i] = p.x * inv;
    rgb[3 * i + 1] = p.y * inv;
    rgb[3 * i + 2] = p.z * inv;
}

// Goblin::bloom (src/GoblinImageIO.cpp:169-218): every pixel gathers its (2 fw - 1)^2 neighbourhood
// (itself excluded) with the radial weight table, in the reference's row-major order so that the
// float sums are the reference's bit for bit, then blends.  One thread per pixel; a warp covers 32
// consecutive x, so each tap is one coalesced row segment served from L1 / L2.
__global__ void k_bloom(const float* __restrict__ in, float* __restrict__ out, int w, int h, int fw,
    const float* __restrict__ table, float bloomWeight) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int x0 = max(0, x - fw + 1), x1 = min(x + fw - 1, w - 1);
    const int y0 = max(0, y - fw + 1), y1 = min(y + fw - 1, h - 1);
    float r = 0.0f, g = 0.0f, b = 0.0f, weightSum = 0.0f;
    for (int py = y0; py <= y1; ++py) {
        const int fy = abs(py - y);
        for (int px = x0; px <= x1; ++px) {
            const int fx = abs(px - x);
            if (fx == 0 && fy == 0) continue;
            const float wgt = __ldg(table + fy * fw + fx);
            const float* c = in + 3 * ((size_t)py * w + px);
            r += c[0] * wgt;
            g += c[1] * wgt;
            b += c[2] * wgt;
            weightSum += wgt;
        }
    }
    const float inv = 1.0f / weightSum; // Color::operator/=(float)
    r *= inv; g *= inv; b *= inv;
    const size_t i = 3 * ((size_t)y * w + x);
    const float keep = 1.0f - bloomWeight;
    out[i] = in[i] * keep + r * bloomWeight;
    out[i + 1] = in[i + 1] * keep + g * bloomWeight;
    out[i + 2] = in[i + 2] * keep + b * bloomWeight;
}

__global__ void k_copy_L(PathState ps, unsigned int n, float* out) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 L = ps.L[i];
    out[3 * i] = L.x; out[3 * i + 1] = L.y; out[3 * i + 2] = L.z;
}

__global__ void k_camera_rays(DeviceScene sc, const float* samples, unsigned int n, gb_ray* rays) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 o, d;
    cameraRay(sc, samples[4 * i], samples[4 * i + 1], samples[4 * i + 2], samples[4 * i + 3], &o, &d);
    gb_ray r;
    r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z;
    r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
    r.mint = cameraMint(sc);
    r.maxt = INFINITY;
    rays[i] = r;
}

} // namespace gb

// =========================================================================
// host side of the device context
// =========================================================================
using namespace gb;

struct gb_context {
    int device = 0;
    int numSMs = 0;
    cudaStream_t stream = nullptr;
    bool overlapTails = true;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    bool haveScene = false;
    DeviceScene sc{};
    // Every scene array lives in one device allocation, filled through one pinned staging buffer.  There are two
    // such pairs ("slots"): gb_upload_scene_async stages and copies the NEXT scene into the idle slot on its own
    // stream while kernels still read the current one; the second slot only exists once that call has been used.
    struct SceneSlot {
        char* dev = nullptr;
        char* host = nullptr;
        size_t devCap = 0, hostCap = 0;
        bool pinned = false;
        cudaEvent_t uploaded = nullptr; // copy + derive kernels of the last upload into this slot are done
        cudaEvent_t lastUse = nullptr;  // everything queued on the context's stream while this slot was current
        int* deriveError = nullptr;     // device flag of the derive kernels (bad index in the scene arrays)
        int* deriveErrorHost = nullptr; // pinned
        bool checkPending = false;      // an asynchronous upload whose deriveErrorHost has not been looked at yet
    } slots[2];
    int slot = 0;                   // the slot ctx->sc points into
    cudaStream_t copyStream = nullptr;
    size_t uploadBytes = 0;
    float* filmHost = nullptr;       // pinned staging for film downloads
    size_t filmHostPixels = 0;
    gb_render_setting setting{};
    int stackEntries = 0; // per-thread traversal stack entries this scene needs: the 4-wide walk's ...
    int stackEntriesPair = 0; // ... and the pair walk's (one per level)
    float4* film = nullptr;
    size_t filmPixels = 0;
    gb_film_desc filmDesc{};
    // Wavefront buffers.  A render is cut into waves of camera samples; two "lanes" of waves run side by side on
    // their own streams with their own path state, so that the tail of one lane's kernel (the last CTAs of a
    // persistent grid on a mostly idle machine, a launch of a few thousand paths late in a path) is filled by the
    // other lane's kernels.  Lane 0 runs on the context's stream.
    struct WaveLane {
        cudaStream_t stream = nullptr;   // lane 0: the context's stream
        cudaStream_t stream2 = nullptr;  // the shadow kernel of bounce b runs here, beside the extend of bounce b + 1
        cudaEvent_t evFork = nullptr, evJoin = nullptr, evDone = nullptr;
        size_t capacity = 0;
        PathState ps{};
        std::vector<void*> allocs;
        unsigned int* ctr = nullptr; // (kMaxDepthCtr) x kCtrStride
    } lanes[4];
    int waveLanes = 2;           // lanes a render uses (gb_set_tuning values[5], 1 .. 4; 1 = one wave at a time)
    cudaEvent_t evLaneStart = nullptr;
    unsigned long long* traceHead = nullptr;
    unsigned long long* stats = nullptr;
    bool statsOn = false;
    uint64_t launches = 0;
    void* comm = nullptr;   // ncclComm_t of the film all-reduce (film_comm.inl); owned unless attached
    bool commOwned = false;
    int commRanks = 0;
    bool wideFits = false;  // the 4-wide walk's stack fits in shared memory for this scene
    int traceMode = GB_TRACE_PAIR;
    // launch geometry of the traversal kernels, keyed by kernel function: depends on the kernel, the
    // stack size of the uploaded scene and the tuning only (invalidated by gb_upload_scene / gb_set_tuning)
    std::vector<std::pair<const void*, int>> gridCache;
    size_t maxWavePaths = 32u << 20;
    TraceTuning tune{20u, 6u, 4u, 10u};
    int blocksPerSM = 0; // 0 = as many as fit
    bool generalFilm = false;    // gb_set_tuning values[7]: k_film also where k_film_warp applies (tests)
    unsigned int matBins = 0;    // bit m: some material of the scene has type m, so shade bin m can be non-empty
    bool hasMeshLight = false; // the scene has a mesh emitter: shade kernels with the GeometrySet loop
    // optional per-kernel-class timing (CUDA event pairs on the context's stream)
    bool timingOn = false;
    bool syncLaunches = false; // GB_SYNC_LAUNCHES=1
    std::vector<cudaEvent_t> evPool;
    size_t evUsed = 0;
    std::vector<int> evClass; // class of event pair k (events 2k, 2k + 1)
    double classMs[GB_K_COUNT] = {};
    uint64_t classLaunches[GB_K_COUNT] = {};
};

namespace {

constexpr int kMaxDepthCtr = 66;

#define GB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            return gb::failWith(GB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                                      \
    } while (0)

void freeScene(gb_context* ctx) {
    for (gb_context::SceneSlot& sl : ctx->slots) {
        if (sl.dev) cudaFree(sl.dev);
        if (sl.host) { if (sl.pinned) cudaFreeHost(sl.host); else std::free(sl.host); }
        sl.dev = nullptr;
        sl.host = nullptr;
        sl.devCap = sl.hostCap = 0;
        sl.checkPending = false;
    }
    if (ctx->film) cudaFree(ctx->film);
    ctx->film = nullptr;
    ctx->filmPixels = 0;
    ctx->haveScene = false;
}

void freeWave(gb_context::WaveLane& lane) {
    for (void* p : lane.allocs) cudaFree(p);
    lane.allocs.clear();
    lane.capacity = 0;
}

// Kernel-class timing: an event pair around each launch, resolved at collect time.
struct KernelTick {
    gb_context* ctx;
    cudaStream_t on;
    int cls;
    bool live = false;
    KernelTick(gb_context* c, int cls_, cudaStream_t s = nullptr) : ctx(c), on(s ? s : c->stream), cls(cls_) {
        if (!ctx->timingOn) return;
        if (ctx->evUsed + 2 > ctx->evPool.size()) {
            for (int k = 0; k < 2; ++k) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreate(&e) != cudaSuccess) return;
                ctx->evPool.push_back(e);
            }
        }
        cudaEventRecord(ctx->evPool[ctx->evUsed], on);
        ctx->evClass.push_back(cls);
        live = true;
    }
    ~KernelTick() {
        if (ctx->syncLaunches) { // GB_SYNC_LAUNCHES=1 (fault hunting): wait for the launch and name the class that failed
            const cudaError_t e = cudaStreamSynchronize(on);
            if (e != cudaSuccess) std::fprintf(stderr, "[goblin_b200] kernel class %d failed: %s\n", cls, cudaGetErrorString(e));
        }
        if (!live) return;
        cudaEventRecord(ctx->evPool[ctx->evUsed + 1], on);
        ctx->evUsed += 2;
    }
};

void collectTimes(gb_context* ctx) { // the stream must be idle
    for (size_t k = 0; 2 * k + 1 < ctx->evUsed; ++k) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->evPool[2 * k], ctx->evPool[2 * k + 1]) == cudaSuccess) {
            ctx->classMs[ctx->evClass[k]] += ms;
            ctx->classLaunches[ctx->evClass[k]]++;
        }
    }
    ctx->evUsed = 0;
    ctx->evClass.clear();
}

template <typename T>
int waveAlloc(gb_context::WaveLane& lane, T** out, size_t n) {
    void* p = nullptr;
    GB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    lane.allocs.push_back(p);
    *out = static_cast<T*>(p);
    return GB_OK;
}

int ensureWave(gb_context::WaveLane& lane, size_t paths) {
    if (paths <= lane.capacity) return GB_OK;
    GB_CUDA(cudaStreamSynchronize(lane.stream)); // kernels of an earlier render may still use the old buffers
    freeWave(lane);
    PathState& ps = lane.ps;
    int rc;
#define WA(field, type) if ((rc = waveAlloc<type>(lane, &ps.field, paths)) != GB_OK) return rc
    WA(rayO, float4); WA(rayD, float4); WA(hit, float4); WA(hitId, int2); WA(thr, float4); WA(L, float4);
    WA(pend, float4); WA(shO, float4); WA(shD, float4); WA(shC, float4);
    WA(qExtend[0], unsigned int); WA(qExtend[1], unsigned int);
    WA(qMat[0], unsigned int); WA(qMat[1], unsigned int); WA(qMat[2], unsigned int); WA(qMat[3], unsigned int);
    WA(aoCount, unsigned int);
#undef WA
    lane.capacity = paths;
    return GB_OK;
}

// per thread: stackEntries 8-byte stack entries + the 6-float world-space ray
size_t traceSmemFor(int stackEntries) { return ((size_t)stackEntries * sizeof(uint2) + 6 * sizeof(float)) * kTraceBlock; }
int stackEntriesOf(const gb_context* ctx, int mode); // by walk (defined below, next to walkMode)
size_t traceSmem(const gb_context* ctx, int mode) { return traceSmemFor(stackEntriesOf(ctx, mode)); }

template <typename K>
int setupTraceKernel(gb_context* ctx, K kernel, int mode, int* grid) {
    const void* key = reinterpret_cast<const void*>(kernel);
    for (const auto& e : ctx->gridCache) {
        if (e.first == key) { *grid = e.second; return GB_OK; }
    }
    size_t smem = traceSmem(ctx, mode);
    GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSM = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kTraceBlock, smem));
    if (perSM < 1) return gb::failWith(GB_ERR_LIMIT, "traversal kernel does not fit on an SM");
    if (ctx->blocksPerSM > 0) perSM = std::min(perSM, ctx->blocksPerSM);
    *grid = perSM * ctx->numSMs; // persistent: exactly one resident wave of CTAs
    ctx->gridCache.emplace_back(key, *grid);
    return GB_OK;
}

int stackEntriesOf(const gb_context* ctx, int mode) { return mode == WALK_WIDE ? ctx->stackEntries : ctx->stackEntriesPair; }

// which walk the traversal kernels of this context run right now
int walkMode(const gb_context* ctx) {
    if (ctx->statsOn) return WALK_STATS;
    return ctx->traceMode == GB_TRACE_WIDE && ctx->wideFits ? WALK_WIDE : WALK_PAIR;
}

} // namespace

extern "C" {

int gb_device_count(int* count) {
    if (!count) return gb::failWith(GB_ERR_INVALID, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return gb::failWith(GB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return GB_OK;
}

int gb_create(int device, gb_context** out) {
    if (!out) return gb::failWith(GB_ERR_INVALID, "null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        return gb::failWith(GB_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
            cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return gb::failWith(GB_ERR_INVALID, "device index out of range");
    GB_CUDA(cudaSetDevice(device));
    gb_context* ctx = new gb_context();
    ctx->device = device;
    cudaDeviceProp prop;
    GB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->numSMs = prop.multiProcessorCount;
    GB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    GB_CUDA(cudaEventCreateWithFlags(&ctx->evLaneStart, cudaEventDisableTiming));
    for (int l = 0; l < 4; ++l) {
        gb_context::WaveLane& lane = ctx->lanes[l];
        if (l == 0) lane.stream = ctx->stream;
        else GB_CUDA(cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking));
        GB_CUDA(cudaStreamCreateWithFlags(&lane.stream2, cudaStreamNonBlocking));
        GB_CUDA(cudaEventCreateWithFlags(&lane.evFork, cudaEventDisableTiming));
        GB_CUDA(cudaEventCreateWithFlags(&lane.evJoin, cudaEventDisableTiming));
        GB_CUDA(cudaEventCreateWithFlags(&lane.evDone, cudaEventDisableTiming));
        GB_CUDA(cudaMalloc((void**)&lane.ctr, kMaxDepthCtr * kCtrStride * sizeof(unsigned int)));
    }
    if (const char* e = std::getenv("GB_WAVE_LANES")) ctx->waveLanes = std::min(4, std::max(1, std::atoi(e)));
    ctx->syncLaunches = std::getenv("GB_SYNC_LAUNCHES") != nullptr;
    ctx->overlapTails = std::getenv("GB_NO_OVERLAP") == nullptr;
    GB_CUDA(cudaEventCreate(&ctx->evStart));
    GB_CUDA(cudaEventCreate(&ctx->evStop));
    GB_CUDA(cudaMalloc((void**)&ctx->traceHead, 64));
    GB_CUDA(cudaMalloc((void**)&ctx->stats, S_COUNT * sizeof(unsigned long long)));
    GB_CUDA(cudaMemset(ctx->stats, 0, S_COUNT * sizeof(unsigned long long)));
    GB_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    for (gb_context::SceneSlot& sl : ctx->slots) {
        GB_CUDA(cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming));
        GB_CUDA(cudaEventCreateWithFlags(&sl.lastUse, cudaEventDisableTiming));
        GB_CUDA(cudaMalloc((void**)&sl.deriveError, sizeof(int)));
        GB_CUDA(cudaMallocHost((void**)&sl.deriveErrorHost, sizeof(int)));
        *sl.deriveErrorHost = 0;
    }
    *out = ctx;
    return GB_OK;
}

int gb_comm_destroy(gb_context* ctx);

int gb_destroy(gb_context* ctx) {
    if (!ctx) return GB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    gb_comm_destroy(ctx);
    cudaStreamSynchronize(ctx->copyStream);
    freeScene(ctx);
    for (gb_context::WaveLane& lane : ctx->lanes) {
        if (lane.stream) cudaStreamSynchronize(lane.stream);
        if (lane.stream2) cudaStreamSynchronize(lane.stream2);
        freeWave(lane);
        cudaFree(lane.ctr);
    }
    cudaFree(ctx->traceHead);
    cudaFree(ctx->stats);
    cudaStreamSynchronize(ctx->copyStream);
    for (gb_context::SceneSlot& sl : ctx->slots) {
        cudaFree(sl.deriveError);
        cudaFreeHost(sl.deriveErrorHost);
        cudaEventDestroy(sl.uploaded);
        cudaEventDestroy(sl.lastUse);
    }
    cudaStreamDestroy(ctx->copyStream);
    if (ctx->filmHost) cudaFreeHost(ctx->filmHost);
    cudaEventDestroy(ctx->evStart);
    cudaEventDestroy(ctx->evStop);
    for (cudaEvent_t e : ctx->evPool) cudaEventDestroy(e);
    for (int l = 0; l < 4; ++l) {
        gb_context::WaveLane& lane = ctx->lanes[l];
        if (lane.stream2) cudaStreamDestroy(lane.stream2);
        if (lane.evFork) cudaEventDestroy(lane.evFork);
        if (lane.evJoin) cudaEventDestroy(lane.evJoin);
        if (lane.evDone) cudaEventDestroy(lane.evDone);
        if (l > 0 && lane.stream) cudaStreamDestroy(lane.stream);
    }
    cudaEventDestroy(ctx->evLaneStart);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GB_OK;
}

} // extern "C"

// gb_upload_scene: validate the flattened scene, derive the device layouts (leaf-ordered instance
// tables, 48-byte triangle records, 64-byte pair nodes) directly inside one pinned staging arena
// and move it to the GPU with a single copy.  Device and staging arenas are kept across calls and
// only grow, host loops run on all cores: re-uploading a scene costs the fill plus one H2D copy.
namespace {

template <typename F>
void hostParallelFor(size_t n, size_t grain, F body) { // body(begin, end)
    unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (n <= grain || nt == 1) { body(0, n); return; }
    nt = (unsigned)std::min<size_t>(nt, (n + grain - 1) / grain);
    std::vector<std::thread> pool;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        size_t b = t * per, e = std::min(n, b + per);
        if (b >= e) break;
        pool.emplace_back([=]() { body(b, e); });
    }
    for (auto& th : pool) th.join();
}

// Structure check, depth and numbering of a pre-order node array.  Children always follow their parent in the
// reference's layout (left = i + 1, right = secondChildOffset > i + 1), so a subtree is a contiguous index range
// and the pair / wide index of a node is simply the number of interior nodes / wide roots before it.
//   scanRange: the sequential scan of one subtree's range [begin, end) whose root sits at depth `rootDepth`.
//   scanTree:  small trees: one scanRange.  Large trees (the 20 M-node tree of the 10 M-triangle scene took
//              120 ms of every upload in one thread): the top of the tree is expanded sequentially into a few
//              hundred disjoint subtrees, those are scanned in parallel, the numbering is a parallel prefix sum.
// A malformed array can never be read out of bounds: every child index is checked against the range of the subtree
// it must lie in, and every node must be reached exactly once, by its own parent.
bool scanRange(const gb_bvh_node* nodes, uint32_t begin, uint32_t end, int rootDepth, uint64_t primLimit, uint8_t* depth,
    uint8_t* reached, int* deepestOut) {
    if (reached[begin]) return false;
    reached[begin] = 1;
    depth[begin] = (uint8_t)rootDepth;
    int deepest = rootDepth;
    for (uint32_t i = begin; i < end; ++i) {
        if (!reached[i]) return false;
        const gb_bvh_node& nd = nodes[i];
        deepest = std::max(deepest, (int)depth[i]);
        if (nd.nprims == 0) {
            if (nd.axis > 2 || i + 1 >= end || nd.offset <= i + 1 || nd.offset >= end) return false;
            if (depth[i] >= 2 * kMaxStack) return false;
            if (reached[i + 1] || reached[nd.offset]) return false;
            reached[i + 1] = reached[nd.offset] = 1;
            depth[i + 1] = depth[nd.offset] = (uint8_t)(depth[i] + 1);
        } else {
            if ((uint64_t)nd.offset + nd.nprims > primLimit) return false;
            if (nd.nprims == 1 ? nd.offset > REF_INDEX : i > REF_INDEX) return false;
        }
    }
    *deepestOut = deepest;
    return true;
}

bool scanTree(const gb_bvh_node* nodes, uint32_t count, uint64_t primLimit, int* depthOut,
    std::vector<uint32_t>& pairIndex, uint32_t* nPairsOut, std::vector<uint32_t>& wideIndex, uint32_t* nWideOut) {
    *depthOut = 0;
    *nPairsOut = 0;
    *nWideOut = 0;
    pairIndex.resize(count);
    wideIndex.resize(count);
    if (count == 0) return true;
    std::vector<uint8_t> depth(count), reached(count, 0);
    int deepest = 0;
    const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    struct Range { uint32_t begin, end; int depth; };
    std::vector<Range> ranges{{0u, count, 0}};
    if (count >= (1u << 18) && nt > 1) {
        // expand the largest range until there are enough subtrees to balance the threads
        while (ranges.size() < 16 * (size_t)nt) {
            size_t big = 0;
            for (size_t k = 1; k < ranges.size(); ++k) if (ranges[k].end - ranges[k].begin > ranges[big].end - ranges[big].begin) big = k;
            const Range r = ranges[big];
            if (r.end - r.begin < (1u << 14)) break;
            const uint32_t i = r.begin;
            const gb_bvh_node& nd = nodes[i];
            if (nd.nprims != 0) break; // cannot happen for a range this large in a valid tree; scanRange will say so
            if (nd.axis > 2 || i + 1 >= r.end || nd.offset <= i + 1 || nd.offset >= r.end || r.depth >= 2 * kMaxStack) return false;
            if (reached[i]) return false;
            reached[i] = 1;
            depth[i] = (uint8_t)r.depth;
            deepest = std::max(deepest, r.depth);
            ranges[big] = Range{i + 1, nd.offset, r.depth + 1};
            ranges.push_back(Range{nd.offset, r.end, r.depth + 1});
        }
    }
    std::atomic<bool> ok{true};
    std::atomic<size_t> next{0};
    std::vector<int> deepestOf(ranges.size(), 0);
    auto worker = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= ranges.size() || !ok.load()) return;
            if (!scanRange(nodes, ranges[k].begin, ranges[k].end, ranges[k].depth, primLimit, depth.data(), reached.data(), &deepestOf[k])) ok = false;
        }
    };
    if (ranges.size() == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    if (!ok.load()) return false;
    for (int v : deepestOf) deepest = std::max(deepest, v);
    // numbering: exclusive prefix sums, in index order, of "interior" and of "interior at an even depth" (a wide root)
    const size_t chunks = count >= (1u << 18) ? 4 * (size_t)nt : 1;
    const size_t per = (count + chunks - 1) / chunks;
    std::vector<uint32_t> cPairs(chunks + 1, 0u), cWide(chunks + 1, 0u);
    hostParallelFor(chunks, 1, [&](size_t cb, size_t ce) {
        for (size_t c = cb; c < ce; ++c) {
            uint32_t np = 0, nw = 0;
            for (size_t i = c * per, e = std::min<size_t>(count, i + per); i < e; ++i) {
                if (!reached[i]) { ok = false; break; }
                if (nodes[i].nprims == 0) { ++np; nw += (depth[i] & 1u) == 0u; }
            }
            cPairs[c + 1] = np;
            cWide[c + 1] = nw;
        }
    });
    if (!ok.load()) return false;
    for (size_t c = 0; c < chunks; ++c) { cPairs[c + 1] += cPairs[c]; cWide[c + 1] += cWide[c]; }
    hostParallelFor(chunks, 1, [&](size_t cb, size_t ce) {
        for (size_t c = cb; c < ce; ++c) {
            uint32_t np = cPairs[c], nw = cWide[c];
            for (size_t i = c * per, e = std::min<size_t>(count, i + per); i < e; ++i) {
                const bool interior = nodes[i].nprims == 0;
                pairIndex[i] = interior ? np++ : 0u;
                wideIndex[i] = interior && (depth[i] & 1u) == 0u ? nw++ : WIDE_NOT_ROOT;
            }
        }
    });
    if (cPairs[chunks] > REF_INDEX) return false;
    *depthOut = deepest;
    *nPairsOut = cPairs[chunks];
    *nWideOut = cWide[chunks];
    return true;
}

inline uint32_t refOf(const gb_bvh_node* nodes, const uint32_t* pairIndex, uint32_t node) {
    const gb_bvh_node& nd = nodes[node];
    if (nd.nprims == 0) return pairIndex[node];
    if (nd.nprims == 1) return REF_LEAF | nd.offset;
    return REF_LEAF | REF_MULTI | node;
}

// Postfix programs for the textured material slots: a texture's children come before it, so the
// device evaluates a program left to right on a value stack.  Validates the part of the texture
// table the materials reach (types, child indices pointing at EARLIER entries: no cycles, float
// children where a float is read) and the stack / length bounds of the evaluator.
bool compileTexturePrograms(const gb_scene_desc* d, std::vector<int4>* matTex, std::vector<int4>* matMask,
    std::vector<unsigned int>* prog, std::string* err) {
    auto emit = [&](int root, bool wantFloat, int* offset) -> bool {
        std::vector<unsigned int> out;
        int depth = 0, maxDepth = 0;
        bool ok = true;
        // recursion depth is bounded by the index order (children < parent)
        std::function<void(int, bool)> rec = [&](int t, bool isFloat) {
            if (!ok) return;
            if (t < 0 || (uint32_t)t >= d->n_textures || out.size() > 4096) { ok = false; return; }
            const gb_texture& g = d->textures[t];
            if ((g.is_float != 0) != isFloat) { ok = false; return; }
            if (g.type == GB_TEX_CHECKERBOARD || g.type == GB_TEX_SCALE) {
                if (g.child[0] >= t || g.child[1] >= t) { ok = false; return; }
                rec(g.child[0], isFloat);
                rec(g.child[1], g.type == GB_TEX_SCALE ? true : isFloat);
                if (!ok) return;
                if (g.type == GB_TEX_CHECKERBOARD && g.mapping != GB_MAPPING_UV && g.mapping != GB_MAPPING_SPHERICAL) { ok = false; return; }
                --depth; // two values in, one out
            } else if (g.type == GB_TEX_CONSTANT || g.type == GB_TEX_IMAGE) {
                if (g.type == GB_TEX_IMAGE) {
                    if (g.n_levels < 1 || g.first_level < 0 || (uint64_t)g.first_level + (uint64_t)g.n_levels > d->n_image_levels ||
                        g.image_filter < GB_FILTER_NEAREST || g.image_filter > GB_FILTER_EWA ||
                        g.address_mode < GB_ADDRESS_REPEAT || g.address_mode > GB_ADDRESS_BORDER ||
                        (g.mapping != GB_MAPPING_UV && g.mapping != GB_MAPPING_SPHERICAL)) { ok = false; return; }
                    for (int l = 0; l < g.n_levels; ++l) {
                        const gb_image_level& il = d->image_levels[g.first_level + l];
                        if (il.width < 1 || il.height < 1 || il.texel_offset > 0xffffffffull ||
                            il.texel_offset + (uint64_t)il.width * (uint64_t)il.height > d->n_image_texels) { ok = false; return; }
                    }
                }
                ++depth;
                maxDepth = std::max(maxDepth, depth);
            } else {
                ok = false;
                return;
            }
            out.push_back((unsigned int)t);
        };
        rec(root, wantFloat);
        if (!ok) { *err = "malformed texture table (type, child order or format)"; return false; }
        if (maxDepth > kTexStack) { *err = "texture nesting exceeds the evaluator's stack"; return false; }
        if (prog->empty()) prog->push_back(0u); // offset 0 means "constant"
        *offset = (int)prog->size();
        prog->push_back((unsigned int)out.size());
        prog->insert(prog->end(), out.begin(), out.end());
        return true;
    };
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const gb_material& mm = d->materials[m];
        int4 slots = make_int4(0, 0, 0, 0);
        if (mm.kd_tex && !emit(mm.kd_tex - 1, false, &slots.x)) return false;
        if (mm.kt_tex && mm.type == GB_MAT_TRANSPARENT && !emit(mm.kt_tex - 1, false, &slots.y)) return false;
        if (mm.exponent_tex && mm.type == GB_MAT_BLINN && !emit(mm.exponent_tex - 1, true, &slots.z)) return false;
        int4 slots2 = make_int4(0, 0, 0, 0);
        if (mm.bump_tex && !emit(mm.bump_tex - 1, true, &slots.w)) return false;
        if (mm.normal_tex && !emit(mm.normal_tex - 1, false, &slots2.x)) return false;
        (*matTex)[2 * (size_t)m] = slots;
        (*matTex)[2 * (size_t)m + 1] = slots2;
        if (mm.mask) { // MaskMaterial: alpha (float) and transparent colour, constants or programs
            int4 info = make_int4(1, 0, 0, 0);
            if (mm.alpha_tex && !emit(mm.alpha_tex - 1, true, &info.y)) return false;
            if (mm.transparent_tex && !emit(mm.transparent_tex - 1, false, &info.z)) return false;
            float v[4] = {mm.alpha, mm.transparent_color[0], mm.transparent_color[1], mm.transparent_color[2]};
            int4 bits;
            std::memcpy(&bits, v, 16);
            (*matMask)[2 * (size_t)m] = info;
            (*matMask)[2 * (size_t)m + 1] = bits;
            if (prog->empty()) prog->push_back(0u); // mask scenes run the texture-capable shade kernels
        }
    }
    return true;
}

} // namespace (reopened below: kernels have external linkage)

// Pair nodes from the reference's 32-byte nodes (fillPairs, on the device): one thread per node, interior
// nodes write the 64-byte record of their two children at the pair index the host numbered them with.
__global__ void k_derive_pairs(const float4* __restrict__ nodes, const unsigned int* __restrict__ pairIndex,
    unsigned int count, float4* __restrict__ out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 n1 = __ldg(nodes + 2 * (size_t)i + 1);
    const unsigned int word = __float_as_uint(n1.w); // nprims | axis << 8
    if ((word & 0xffu) != 0u) return;
    const unsigned int right = __float_as_uint(n1.z), left = i + 1;
    auto refOfNode = [&](unsigned int node, float4 c1) -> unsigned int {
        const unsigned int w = __float_as_uint(c1.w) & 0xffu;
        if (w == 0u) return __ldg(pairIndex + node);
        if (w == 1u) return REF_LEAF | __float_as_uint(c1.z);
        return REF_LEAF | REF_MULTI | node;
    };
    const float4 l0 = __ldg(nodes + 2 * (size_t)left), l1 = __ldg(nodes + 2 * (size_t)left + 1);
    const float4 r0 = __ldg(nodes + 2 * (size_t)right), r1 = __ldg(nodes + 2 * (size_t)right + 1);
    float4* q = out + 4 * (size_t)__ldg(pairIndex + i);
    q[0] = make_float4(l0.x, l0.y, l0.z, l0.w);
    q[1] = make_float4(l1.x, l1.y, r0.x, r0.y);
    q[2] = make_float4(r0.z, r0.w, r1.x, r1.y);
    q[3] = make_float4(__uint_as_float(refOfNode(left, l1)), __uint_as_float(refOfNode(right, r1)),
        __uint_as_float((word >> 8) & 0xffu), 0.0f);
}

// 4-wide nodes (wide_node.h) from the same nodes: one thread per node, wide roots (interior, even depth: the host
// numbered them) gather their grandchildren.
__global__ void k_derive_wide(const gb_bvh_node* __restrict__ nodes, const unsigned int* __restrict__ wideIndex,
    unsigned int count, float4* __restrict__ out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned int w = wideIndex[i];
    if (w == WIDE_NOT_ROOT) return;
    WideNode wn;
    deriveWideNode(nodes, wideIndex, i, &wn);
    float4* q = out + 8 * (size_t)w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const WideChild& c = wn.c[k];
        q[2 * k] = make_float4(c.lo[0], c.lo[1], c.lo[2], c.hi[0]);
        q[2 * k + 1] = make_float4(c.hi[1], c.hi[2], __uint_as_float(c.ref), __uint_as_float(c.meta));
    }
}

// Triangle test records (p0, e1, e2, face) and shading records (vertex normals, uvs) in BVH leaf order,
// one thread per leaf slot: e1 = p1 - p0, e2 = p2 - p0 are the reference's per-test subtractions
// (src/GoblinTriangle.cpp:51-54), done once.  Index errors are reported through `error`.
__global__ void k_derive_tris(const unsigned int* __restrict__ order, const unsigned int* __restrict__ triIndex,
    const float* __restrict__ pos, const float* __restrict__ nrm, const float* __restrict__ uv, unsigned int triCount,
    unsigned int vertCount, float4* __restrict__ triRec, float4* __restrict__ triShade, int* error) {
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= triCount) return;
    const unsigned int face = __ldg(order + k);
    if (face >= triCount) { *error = 1; return; }
    const unsigned int v0 = __ldg(triIndex + 3 * (size_t)face), v1 = __ldg(triIndex + 3 * (size_t)face + 1),
                       v2 = __ldg(triIndex + 3 * (size_t)face + 2);
    if (v0 >= vertCount || v1 >= vertCount || v2 >= vertCount) { *error = 2; return; }
    const float* p0 = pos + 3 * (size_t)v0;
    const float* p1 = pos + 3 * (size_t)v1;
    const float* p2 = pos + 3 * (size_t)v2;
    const float p0x = __ldg(p0), p0y = __ldg(p0 + 1), p0z = __ldg(p0 + 2);
    const float e1x = __ldg(p1) - p0x, e1y = __ldg(p1 + 1) - p0y, e1z = __ldg(p1 + 2) - p0z;
    const float e2x = __ldg(p2) - p0x, e2y = __ldg(p2 + 1) - p0y, e2z = __ldg(p2 + 2) - p0z;
    float4* r = triRec + kTriRecVec4 * (size_t)k;
    r[0] = make_float4(p0x, p0y, p0z, e1x);
    r[1] = make_float4(e1y, e1z, e2x, e2y);
    r[2] = make_float4(e2z, __uint_as_float(face), 0.0f, 0.0f);
    if (kTriRecVec4 == 4) r[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float* n0 = nrm + 3 * (size_t)v0;
    const float* n1 = nrm + 3 * (size_t)v1;
    const float* n2 = nrm + 3 * (size_t)v2;
    const float* t0 = uv + 2 * (size_t)v0;
    const float* t1 = uv + 2 * (size_t)v1;
    const float* t2 = uv + 2 * (size_t)v2;
    float4* q = triShade + 4 * (size_t)k;
    q[0] = make_float4(__ldg(n0), __ldg(n0 + 1), __ldg(n0 + 2), __ldg(n1));
    q[1] = make_float4(__ldg(n1 + 1), __ldg(n1 + 2), __ldg(n2), __ldg(n2 + 1));
    q[2] = make_float4(__ldg(n2 + 2), __ldg(t0), __ldg(t0 + 1), __ldg(t1));
    q[3] = make_float4(__ldg(t1 + 1), __ldg(t2), __ldg(t2 + 1), 0.0f);
}

namespace {

struct Arena { // offsets into the staging / device arena, 256-byte aligned
    size_t size = 0;
    size_t take(size_t bytes) {
        size_t off = size;
        size = (size + std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
        return off;
    }
};

} // namespace

namespace {

// An asynchronous upload reports a bad index found by its derive kernels at the next call that waits for the device.
int checkDeferredUpload(gb_context* ctx) {
    for (gb_context::SceneSlot& sl : ctx->slots) {
        if (!sl.checkPending || cudaEventQuery(sl.uploaded) != cudaSuccess) continue;
        sl.checkPending = false;
        const int code = *sl.deriveErrorHost;
        if (code != 0) {
            ctx->haveScene = false;
            return gb::failWith(GB_ERR_INVALID, code == 1 ? "model_order entry out of range (asynchronous upload)"
                                                          : "vertex index out of range (asynchronous upload)");
        }
    }
    return GB_OK;
}

} // namespace

// async = false: gb_upload_scene (waits for the device before and after, clears the film, one slot).
// async = true:  gb_upload_scene_async (include/goblin_b200.h): the idle slot, the copy stream, no wait.
static int uploadScene(gb_context* ctx, const gb_scene_desc* d, bool async) {
    if (!ctx || !d) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    if (!async) {
        GB_CUDA(cudaStreamSynchronize(ctx->stream));
        GB_CUDA(cudaStreamSynchronize(ctx->copyStream));
        ctx->haveScene = false;
    }
    const int k = async && ctx->haveScene ? 1 - ctx->slot : ctx->slot; // the slot this upload fills
    gb_context::SceneSlot& slot = ctx->slots[k];
    const uint32_t nInst = d->n_instances;
    const gb_film_desc& f = d->film;
    if (f.xres <= 0 || f.yres <= 0) return gb::failWith(GB_ERR_INVALID, "bad film resolution");
    // GB_UPLOAD_TIMING=1 prints where the host time of an upload goes (e2e overhead analysis)
    static const bool timing = std::getenv("GB_UPLOAD_TIMING") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[gb_upload_scene] %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    // ---- validate, measure tree depth (stack need), number the pair nodes
    int topDepth = 0, modelDepth = 0;
    std::vector<uint32_t> topPairIndex, topWideIndex;
    uint32_t nTopPairs = 0, nTopWide = 0;
    if (!scanTree(d->top_nodes, d->n_top_nodes, nInst, &topDepth, topPairIndex, &nTopPairs, topWideIndex, &nTopWide)) {
        return gb::failWith(GB_ERR_INVALID, "malformed top-level BVH");
    }
    std::vector<std::vector<uint32_t>> modelPairIndex(d->n_models), modelWideIndex(d->n_models);
    std::vector<uint32_t> modelPairBase(d->n_models, 0u), modelPairCount(d->n_models, 0u);
    std::vector<uint32_t> modelWideBase(d->n_models, 0u), modelWideCount(d->n_models, 0u);
    uint64_t nModelPairs = 0, nModelWide = 0;
    for (uint32_t m = 0; m < d->n_models; ++m) {
        const gb_model& md = d->models[m];
        if (md.material < 0 || (uint32_t)md.material >= d->n_materials) return gb::failWith(GB_ERR_INVALID, "model material out of range");
        if (md.kind != GB_GEOM_MESH) continue;
        if ((uint64_t)md.node_offset + md.node_count > d->n_model_nodes || (uint64_t)md.tri_offset + md.tri_count > d->n_tris ||
            (uint64_t)md.vert_offset + md.vert_count > d->n_verts) {
            return gb::failWith(GB_ERR_INVALID, "model ranges exceed the scene arrays");
        }
        int dm = 0;
        if (!scanTree(d->model_nodes + md.node_offset, md.node_count, md.tri_count, &dm, modelPairIndex[m], &modelPairCount[m],
                modelWideIndex[m], &modelWideCount[m])) {
            return gb::failWith(GB_ERR_INVALID, "malformed model BVH");
        }
        modelDepth = std::max(modelDepth, dm);
        if (nModelPairs > 0xffffffffull) return gb::failWith(GB_ERR_LIMIT, "model BVHs exceed the 32-bit pair index");
        modelPairBase[m] = (uint32_t)nModelPairs;
        nModelPairs += modelPairCount[m];
        modelWideBase[m] = (uint32_t)nModelWide;
        nModelWide += modelWideCount[m];
    }
    // The pair walk (push far / go near) keeps at most one entry per level, the 4-wide walk at most three per
    // wide level; both levels of the scene share one column.  A scene too deep for the wide walk's column in
    // shared memory is walked pair-wise; one too deep for that as well is refused -- before anything is touched.
    const int pairEntries = topDepth + modelDepth + 2;
    const int wideEntries = std::max(pairEntries, wideStackEntries(topDepth) + wideStackEntries(modelDepth) + 2);
    if (pairEntries > 2 * kMaxStack) {
        return gb::failWith(GB_ERR_LIMIT, "BVH deeper than the traversal stack (reference: todo[64] per level)");
    }
    if (traceSmemFor(pairEntries) > kMaxTraceSmem) return gb::failWith(GB_ERR_LIMIT, "BVH too deep for the shared-memory stack");
    size_t maxWideSmem = kMaxWideSmem;
    if (const char* e = std::getenv("GB_MAX_WIDE_SMEM")) maxWideSmem = (size_t)std::strtoull(e, nullptr, 10); // tests: force the fallback
    const bool wideFits = traceSmemFor(wideEntries) <= maxWideSmem;
    const int stackEntries = wideFits ? wideEntries : pairEntries;
    const int stackEntriesPair = pairEntries;
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        if (d->materials[m].type < 0 || d->materials[m].type >= GB_MAT_COUNT) return gb::failWith(GB_ERR_INVALID, "unsupported material type");
    }
    std::vector<int> slotOf(nInst, -1); // original instance index -> leaf slot
    for (uint32_t s = 0; s < nInst; ++s) {
        uint32_t id = d->top_order[s];
        if (id >= nInst || slotOf[id] >= 0) return gb::failWith(GB_ERR_INVALID, "top_order is not a permutation");
        slotOf[id] = (int)s;
        const gb_instance& in = d->instances[id];
        if (in.model < 0 || (uint32_t)in.model >= d->n_models) return gb::failWith(GB_ERR_INVALID, "instance model out of range");
    }
    lap("validate");
    // ---- arena layout
    Arena ar;
    const size_t oTopNodes = ar.take(32 * (size_t)d->n_top_nodes);
    const size_t oModelNodes = ar.take(32 * (size_t)d->n_model_nodes);
    // what the derive kernels read (the reference's own arrays + the pair numbering): uploaded
    const size_t oRawTopPairIdx = ar.take(4 * (size_t)d->n_top_nodes);
    const size_t oRawModelPairIdx = ar.take(4 * (size_t)d->n_model_nodes);
    const size_t oRawTopWideIdx = ar.take(4 * (size_t)d->n_top_nodes);
    const size_t oRawModelWideIdx = ar.take(4 * (size_t)d->n_model_nodes);
    const size_t oRawOrder = ar.take(4 * (size_t)d->n_tris);
    const size_t oRawTriIndex = ar.take(12 * (size_t)d->n_tris);
    const size_t oRawPos = ar.take(12 * (size_t)d->n_verts);
    const size_t oRawNrm = ar.take(12 * (size_t)d->n_verts);
    const size_t oRawUv = ar.take(8 * (size_t)d->n_verts);
    const size_t oInstToObject = ar.take(48 * (size_t)nInst);
    const size_t oInstToWorld = ar.take(48 * (size_t)nInst);
    const size_t oInstInfo = ar.take(16 * (size_t)nInst);
    const size_t oInstInfo2 = ar.take(16 * (size_t)nInst);
    const size_t oInstShade = ar.take(16 * (size_t)nInst);
    const size_t oModelShade = ar.take(16 * (size_t)d->n_models);
    const size_t oMaterials = ar.take(sizeof(DeviceMaterial) * (size_t)d->n_materials);
    // procedural textures: postfix programs of the textured material slots
    std::vector<int4> matTex(2 * (size_t)d->n_materials, make_int4(0, 0, 0, 0));
    std::vector<unsigned int> texProg;
    std::vector<int4> matMask(2 * (size_t)d->n_materials, make_int4(0, 0, 0, 0));
    bool hasMask = false;
    for (uint32_t m = 0; m < d->n_materials; ++m) hasMask = hasMask || d->materials[m].mask != 0;
    {
        std::string terr;
        if (!compileTexturePrograms(d, &matTex, &matMask, &texProg, &terr)) return gb::failWith(GB_ERR_INVALID, terr);
    }
    const bool hasTextures = !texProg.empty();
    if (hasMask && (topDepth > kSimpleStack || modelDepth > kSimpleStack)) {
        return gb::failWith(GB_ERR_LIMIT, "mask scenes walk each BVH level with the reference's todo[64]: a tree is deeper");
    }
    const size_t oMatMask = ar.take(hasMask ? 32 * (size_t)d->n_materials : 0);
    const size_t oMatTex = ar.take(hasTextures ? 32 * (size_t)d->n_materials : 0);
    const size_t oTexProg = ar.take(hasTextures ? 4 * texProg.size() : 0);
    const size_t oTexNodes = ar.take(hasTextures ? 16 * (size_t)kTexNodeVec4 * d->n_textures : 0);
    const size_t oTexLevels = ar.take(hasTextures ? 16 * (size_t)d->n_image_levels : 0);
    const size_t oImageTexels = ar.take(16 * (size_t)d->n_image_texels);
    const size_t oLightDist = ar.take(4 * (size_t)d->n_light_dist);
    const size_t oLights = ar.take(sizeof(DeviceLight) * (size_t)d->n_lights);
    const size_t oLightPower = ar.take(4 * (size_t)d->n_lights);
    const size_t oLightCdf = ar.take(4 * ((size_t)d->n_lights + 1));
    const size_t oLightTris = ar.take(96 * (size_t)d->n_light_tri_area);
    const size_t oLightTriCdf = ar.take(4 * (size_t)d->n_light_tri_cdf);
    const size_t oFilter = ar.take(4 * 256);
    // everything above crosses PCIe; the arrays below are derived from it on the device
    // (k_derive_pairs, k_derive_tris): only the device arena holds them
    const size_t uploadSize = ar.size;
    const size_t oTopPairs = ar.take(64 * (size_t)nTopPairs);
    const size_t oModelPairs = ar.take(64 * (size_t)nModelPairs);
    const size_t oTopWide = ar.take(128 * (size_t)nTopWide);
    const size_t oModelWide = ar.take(128 * (size_t)nModelWide);
    const size_t oTriRec = ar.take(16 * (size_t)kTriRecVec4 * (size_t)d->n_tris);
    const size_t oTriShade = ar.take(64 * (size_t)d->n_tris);
    // the staging buffer of this slot is free once its previous copy has left it; its device arena once the kernels
    // that read it have finished (two uploads ago in a pipeline: both long done)
    GB_CUDA(cudaEventSynchronize(slot.uploaded));
    if (ar.size > slot.devCap) { // both arenas persist and only grow
        GB_CUDA(cudaEventSynchronize(slot.lastUse));
        if (slot.dev) cudaFree(slot.dev);
        slot.dev = nullptr; slot.devCap = 0;
        GB_CUDA(cudaMalloc((void**)&slot.dev, ar.size));
        slot.devCap = ar.size;
    }
    if (uploadSize > slot.hostCap) {
        if (slot.host) { if (slot.pinned) cudaFreeHost(slot.host); else std::free(slot.host); }
        slot.host = nullptr; slot.hostCap = 0;
        void* hp = nullptr;
        if (cudaMallocHost(&hp, uploadSize) == cudaSuccess) { slot.pinned = true; }
        else {
            cudaGetLastError();
            hp = std::malloc(uploadSize);
            slot.pinned = false;
            if (!hp) return gb::failWith(GB_ERR_INVALID, "out of host memory for the staging arena");
        }
        slot.host = static_cast<char*>(hp);
        slot.hostCap = uploadSize;
    }
    char* H = slot.host;
    lap("layout");
    // ---- fill the staging arena
    std::memcpy(H + oTopNodes, d->top_nodes, 32 * (size_t)d->n_top_nodes);
    hostParallelFor(d->n_model_nodes, 1u << 15, [&](size_t b, size_t e) {
        std::memcpy(H + oModelNodes + 32 * b, d->model_nodes + b, 32 * (e - b));
    });
    // the pair numbering of every interior node and the meshes' own arrays: the derive kernels turn
    // them into pair nodes and leaf-order triangle records on the device
    if (d->n_top_nodes) {
        std::memcpy(H + oRawTopPairIdx, topPairIndex.data(), 4 * (size_t)d->n_top_nodes);
        std::memcpy(H + oRawTopWideIdx, topWideIndex.data(), 4 * (size_t)d->n_top_nodes);
    }
    const uint32_t topRootRef = d->n_top_nodes ? refOf(d->top_nodes, topPairIndex.data(), 0) : REF_NONE;
    std::vector<uint32_t> modelRootRef(d->n_models, REF_NONE);
    int4* modelShade = reinterpret_cast<int4*>(H + oModelShade);
    for (uint32_t m = 0; m < d->n_models; ++m) {
        const gb_model& md = d->models[m];
        modelShade[m] = make_int4((int)md.vert_offset, (int)md.tri_offset, (md.has_normal ? 1 : 0) | (md.has_uv ? 2 : 0), 0);
        if (md.kind != GB_GEOM_MESH) continue;
        const gb_bvh_node* nodes = d->model_nodes + md.node_offset;
        if (md.node_count) {
            std::memcpy(H + oRawModelPairIdx + 4 * (size_t)md.node_offset, modelPairIndex[m].data(), 4 * (size_t)md.node_count);
            std::memcpy(H + oRawModelWideIdx + 4 * (size_t)md.node_offset, modelWideIndex[m].data(), 4 * (size_t)md.node_count);
            modelRootRef[m] = refOf(nodes, modelPairIndex[m].data(), 0); // pair 0 = wide root 0 when the root is interior
        }
        std::vector<uint32_t>().swap(modelPairIndex[m]);
        std::vector<uint32_t>().swap(modelWideIndex[m]);
    }
    if (d->n_tris) hostParallelFor(d->n_tris, 1u << 15, [&](size_t b, size_t e) {
        std::memcpy(H + oRawOrder + 4 * b, d->model_order + b, 4 * (e - b));
        std::memcpy(H + oRawTriIndex + 12 * b, d->tri_index + 3 * b, 12 * (e - b));
    });
    if (d->n_verts) hostParallelFor(d->n_verts, 1u << 15, [&](size_t b, size_t e) {
        std::memcpy(H + oRawPos + 12 * b, d->vert_pos + 3 * b, 12 * (e - b));
        std::memcpy(H + oRawNrm + 12 * b, d->vert_nrm + 3 * b, 12 * (e - b));
        std::memcpy(H + oRawUv + 8 * b, d->vert_uv + 2 * b, 8 * (e - b));
    });
    float4* instToObject = reinterpret_cast<float4*>(H + oInstToObject);
    float4* instToWorld = reinterpret_cast<float4*>(H + oInstToWorld);
    int4* instInfo = reinterpret_cast<int4*>(H + oInstInfo);
    int4* instInfo2 = reinterpret_cast<int4*>(H + oInstInfo2);
    int4* instShade = reinterpret_cast<int4*>(H + oInstShade);
    bool hasArea = false;
    for (uint32_t s = 0; s < nInst; ++s) { // instances in BVH leaf order
        const uint32_t id = d->top_order[s];
        const gb_instance& in = d->instances[id];
        const gb_model& md = d->models[in.model];
        for (int r = 0; r < 3; ++r) {
            instToObject[3 * (size_t)s + r] = make_float4(in.to_object[4 * r], in.to_object[4 * r + 1], in.to_object[4 * r + 2], in.to_object[4 * r + 3]);
            instToWorld[3 * (size_t)s + r] = make_float4(in.to_world[4 * r], in.to_world[4 * r + 1], in.to_world[4 * r + 2], in.to_world[4 * r + 3]);
        }
        int radiusBits;
        std::memcpy(&radiusBits, &md.radius, 4);
        instInfo[s] = make_int4(md.kind, (int)md.node_offset, (int)md.tri_offset, radiusBits);
        instInfo2[s] = make_int4((int)modelRootRef[in.model], (int)modelPairBase[in.model],
            md.kind == GB_GEOM_MESH ? (int)md.node_count : 0, (int)modelWideBase[in.model]);
        instShade[s] = make_int4((int)id, in.model, md.material, md.area_light);
        if (md.area_light >= 0) hasArea = true;
    }
    DeviceMaterial* mats = reinterpret_cast<DeviceMaterial*>(H + oMaterials);
    unsigned int matBins = 0;
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const gb_material& mm = d->materials[m];
        matBins |= 1u << mm.type;
        float tb;
        std::memcpy(&tb, &mm.type, 4);
        mats[m].kdType = make_float4(mm.kd[0], mm.kd[1], mm.kd[2], tb);
        // mirror keeps k in ktEta.x (its Kt is unused)
        if (mm.type == GB_MAT_BLINN) {
            float fb;
            std::memcpy(&fb, &mm.fresnel, 4);
            mats[m].ktEta = make_float4(mm.k, mm.exponent, fb, mm.eta);
        } else if (mm.type == GB_MAT_MIRROR) mats[m].ktEta = make_float4(mm.k, 0.0f, 0.0f, mm.eta);
        else mats[m].ktEta = make_float4(mm.kt[0], mm.kt[1], mm.kt[2], mm.eta);
    }
    if (hasMask) std::memcpy(H + oMatMask, matMask.data(), 16 * matMask.size());
    if (hasTextures) {
        std::memcpy(H + oMatTex, matTex.data(), 16 * matTex.size());
        std::memcpy(H + oTexProg, texProg.data(), 4 * texProg.size());
        float4* tn = reinterpret_cast<float4*>(H + oTexNodes);
        for (uint32_t t = 0; t < d->n_textures; ++t) {
            const gb_texture& gt = d->textures[t];
            float4* q = tn + (size_t)kTexNodeVec4 * t;
            float tb;
            std::memcpy(&tb, &gt.type, 4);
            q[0] = make_float4(gt.type == GB_TEX_IMAGE ? gt.max_anisotropy : gt.value[0], gt.value[1], gt.value[2], tb);
            int opts[4] = {gt.type == GB_TEX_IMAGE ? gt.image_filter : gt.filter, gt.mapping, gt.address_mode, gt.first_level};
            std::memcpy(&q[1], opts, 16);
            int img[4] = {gt.n_levels, gt.is_float, 0, 0};
            std::memcpy(&q[6], img, 16);
            q[2] = make_float4(gt.map_scale[0], gt.map_scale[1], gt.map_offset[0], gt.map_offset[1]);
            for (int r = 0; r < 3; ++r) q[3 + r] = make_float4(gt.to_tex[4 * r], gt.to_tex[4 * r + 1], gt.to_tex[4 * r + 2], gt.to_tex[4 * r + 3]);
        }
    }
    if (d->n_image_texels) std::memcpy(H + oImageTexels, d->image_texels, 16 * (size_t)d->n_image_texels);
    if (d->n_light_dist) std::memcpy(H + oLightDist, d->light_dist, 4 * (size_t)d->n_light_dist);
    if (hasTextures) {
        int4* tl = reinterpret_cast<int4*>(H + oTexLevels);
        for (uint32_t l = 0; l < d->n_image_levels; ++l) {
            const gb_image_level& il = d->image_levels[l];
            tl[l] = make_int4(il.width, il.height, (int)(uint32_t)il.texel_offset, 0);
        }
    }
    DeviceLight* lights = reinterpret_cast<DeviceLight*>(H + oLights);
    bool hasMeshLight = false, hasEnvLight = false;
    for (uint32_t l = 0; l < d->n_lights; ++l) {
        const gb_light& gl = d->lights[l];
        DeviceLight& dl = lights[l];
        float tb, kb, sb;
        int t = gl.type, k = gl.geom_kind;
        int slot = (gl.instance >= 0 && (uint32_t)gl.instance < nInst) ? slotOf[gl.instance] : -1;
        std::memcpy(&tb, &t, 4);
        std::memcpy(&kb, &k, 4);
        std::memcpy(&sb, &slot, 4);
        dl.colorType = make_float4(gl.color[0], gl.color[1], gl.color[2], tb);
        dl.posRadius = make_float4(gl.position[0], gl.position[1], gl.position[2], gl.radius);
        dl.dirCos = make_float4(gl.direction[0], gl.direction[1], gl.direction[2], gl.cos_theta_max);
        if (gl.type == GB_LIGHT_AREA && gl.geom_kind == GB_GEOM_MESH) {
            // face-order records of the emitting mesh for GeometrySet::sample / pdf
            if (gl.model < 0 || (uint32_t)gl.model >= d->n_models || d->models[gl.model].kind != GB_GEOM_MESH) {
                return gb::failWith(GB_ERR_INVALID, "mesh area light without a mesh model");
            }
            const gb_model& md = d->models[gl.model];
            if ((uint64_t)gl.area_offset + md.tri_count > d->n_light_tri_area ||
                (uint64_t)gl.cdf_offset + md.tri_count + 1 > d->n_light_tri_cdf) {
                return gb::failWith(GB_ERR_INVALID, "mesh area light ranges exceed light_tri_area / light_tri_cdf");
            }
            float4* lt = reinterpret_cast<float4*>(H + oLightTris) + 6 * (size_t)gl.area_offset;
            for (uint32_t face = 0; face < md.tri_count; ++face) {
                const uint32_t* vi = d->tri_index + 3 * ((size_t)md.tri_offset + face);
                for (int c = 0; c < 3; ++c) {
                    if (vi[c] >= md.vert_count) return gb::failWith(GB_ERR_INVALID, "vertex index out of range");
                    const float* pp = d->vert_pos + 3 * ((size_t)md.vert_offset + vi[c]);
                    const float* pn = d->vert_nrm + 3 * ((size_t)md.vert_offset + vi[c]);
                    lt[6 * (size_t)face + c] = make_float4(pp[0], pp[1], pp[2], c == 0 ? d->light_tri_area[gl.area_offset + face] : 0.0f);
                    lt[6 * (size_t)face + 3 + c] = make_float4(pn[0], pn[1], pn[2], 0.0f);
                }
            }
            uint32_t w[4] = {gl.area_offset, md.tri_count, md.has_normal ? 1u : 0u, gl.cdf_offset};
            std::memcpy(&dl.dirCos, w, 16);
            hasMeshLight = true;
        }
        if (gl.type == GB_LIGHT_IBL) {
            const uint64_t texels = (uint64_t)gl.image_width * (uint64_t)gl.image_height;
            const uint64_t dw = (uint64_t)gl.dist_width, dh = (uint64_t)gl.dist_height;
            if (gl.image_width <= 0 || gl.image_height <= 0 || gl.dist_width <= 0 || gl.dist_height <= 0 ||
                gl.image_offset + texels > d->n_image_texels || gl.image_offset > 0xffffffffull ||
                gl.dist_offset + dw * dh + (dw + 1) * dh + 2 * dh + 2 > d->n_light_dist || gl.dist_offset > 0xffffffffull) {
                return gb::failWith(GB_ERR_INVALID, "image based light ranges exceed image_texels / light_dist");
            }
            int dims[4] = {gl.image_width, gl.image_height, gl.dist_width, gl.dist_height};
            uint32_t offs[4] = {(uint32_t)gl.image_offset, (uint32_t)gl.dist_offset, 0u, 0u};
            std::memcpy(&dl.posRadius, dims, 16);
            std::memcpy(&dl.dirCos, offs, 16);
            hasMeshLight = true; // the shade-kernel variants that carry the rare light kinds
            hasEnvLight = true;
        }
        dl.misc = make_float4(gl.cos_falloff_start, gl.area, kb, sb);
        for (int r = 0; r < 3; ++r) {
            dl.toWorld[r] = make_float4(gl.to_world[4 * r], gl.to_world[4 * r + 1], gl.to_world[4 * r + 2], gl.to_world[4 * r + 3]);
            dl.toObject[r] = make_float4(gl.to_object[4 * r], gl.to_object[4 * r + 1], gl.to_object[4 * r + 2], gl.to_object[4 * r + 3]);
        }
    }
    if (d->n_lights) {
        std::memcpy(H + oLightPower, d->light_power, 4 * (size_t)d->n_lights);
        std::memcpy(H + oLightCdf, d->light_cdf, 4 * ((size_t)d->n_lights + 1));
    }
    std::memcpy(H + oFilter, f.filter_table, 4 * 256);
    if (d->n_light_tri_cdf) std::memcpy(H + oLightTriCdf, d->light_tri_cdf, 4 * (size_t)d->n_light_tri_cdf);
    // CDF1D::mIntegral
    float integral = 0.0f;
    if (d->n_lights) {
        float dx = 1.0f / d->n_lights, acc = 0.0f;
        for (uint32_t l = 0; l < d->n_lights; ++l) acc = acc + d->light_power[l] * dx;
        integral = acc;
    }
    lap("fill");
    // ---- one host-to-device copy
    // on the copy stream, behind the last kernel that read this slot's device arena
    cudaStream_t cs = ctx->copyStream;
    GB_CUDA(cudaStreamWaitEvent(cs, slot.lastUse, 0));
    GB_CUDA(cudaMemcpyAsync(slot.dev, H, uploadSize, cudaMemcpyHostToDevice, cs));
    ctx->uploadBytes = uploadSize;
    char* D = slot.dev;
    // ---- derive the traversal / shading records on the device, at HBM speed instead of PCIe speed
    GB_CUDA(cudaMemsetAsync(slot.deriveError, 0, sizeof(int), cs));
    if (nTopPairs) {
        k_derive_pairs<<<(d->n_top_nodes + 255) / 256, 256, 0, cs>>>(reinterpret_cast<const float4*>(D + oTopNodes),
            reinterpret_cast<const unsigned int*>(D + oRawTopPairIdx), d->n_top_nodes, reinterpret_cast<float4*>(D + oTopPairs));
        ctx->launches++;
    }
    if (nTopWide) {
        k_derive_wide<<<(d->n_top_nodes + 255) / 256, 256, 0, cs>>>(reinterpret_cast<const gb_bvh_node*>(D + oTopNodes),
            reinterpret_cast<const unsigned int*>(D + oRawTopWideIdx), d->n_top_nodes, reinterpret_cast<float4*>(D + oTopWide));
        ctx->launches++;
    }
    for (uint32_t m = 0; m < d->n_models; ++m) {
        const gb_model& md = d->models[m];
        if (md.kind != GB_GEOM_MESH) continue;
        if (modelPairCount[m]) {
            k_derive_pairs<<<(md.node_count + 255) / 256, 256, 0, cs>>>(
                reinterpret_cast<const float4*>(D + oModelNodes) + 2 * (size_t)md.node_offset,
                reinterpret_cast<const unsigned int*>(D + oRawModelPairIdx) + md.node_offset, md.node_count,
                reinterpret_cast<float4*>(D + oModelPairs) + 4 * (size_t)modelPairBase[m]);
            ctx->launches++;
        }
        if (modelWideCount[m]) {
            k_derive_wide<<<(md.node_count + 255) / 256, 256, 0, cs>>>(
                reinterpret_cast<const gb_bvh_node*>(D + oModelNodes) + md.node_offset,
                reinterpret_cast<const unsigned int*>(D + oRawModelWideIdx) + md.node_offset, md.node_count,
                reinterpret_cast<float4*>(D + oModelWide) + 8 * (size_t)modelWideBase[m]);
            ctx->launches++;
        }
        if (md.tri_count) {
            k_derive_tris<<<(md.tri_count + 255) / 256, 256, 0, cs>>>(
                reinterpret_cast<const unsigned int*>(D + oRawOrder) + md.tri_offset,
                reinterpret_cast<const unsigned int*>(D + oRawTriIndex) + 3 * (size_t)md.tri_offset,
                reinterpret_cast<const float*>(D + oRawPos) + 3 * (size_t)md.vert_offset,
                reinterpret_cast<const float*>(D + oRawNrm) + 3 * (size_t)md.vert_offset,
                reinterpret_cast<const float*>(D + oRawUv) + 2 * (size_t)md.vert_offset, md.tri_count, md.vert_count,
                reinterpret_cast<float4*>(D + oTriRec) + kTriRecVec4 * (size_t)md.tri_offset,
                reinterpret_cast<float4*>(D + oTriShade) + 4 * (size_t)md.tri_offset, slot.deriveError);
            ctx->launches++;
        }
    }
    GB_CUDA(cudaMemcpyAsync(slot.deriveErrorHost, slot.deriveError, sizeof(int), cudaMemcpyDeviceToHost, cs));
    GB_CUDA(cudaEventRecord(slot.uploaded, cs));
    DeviceScene sc{};
    sc.tune = ctx->tune;
    sc.topNodes = reinterpret_cast<const float4*>(D + oTopNodes);
    sc.modelNodes = reinterpret_cast<const float4*>(D + oModelNodes);
    sc.topPairs = reinterpret_cast<const float4*>(D + oTopPairs);
    sc.modelPairs = reinterpret_cast<const float4*>(D + oModelPairs);
    sc.topWide = reinterpret_cast<const float4*>(D + oTopWide);
    sc.modelWide = reinterpret_cast<const float4*>(D + oModelWide);
    sc.instToObject = reinterpret_cast<const float4*>(D + oInstToObject);
    sc.instToWorld = reinterpret_cast<const float4*>(D + oInstToWorld);
    sc.instInfo = reinterpret_cast<const int4*>(D + oInstInfo);
    sc.instInfo2 = reinterpret_cast<const int4*>(D + oInstInfo2);
    sc.instShade = reinterpret_cast<const int4*>(D + oInstShade);
    sc.triRec = reinterpret_cast<const float4*>(D + oTriRec);
    sc.modelShade = reinterpret_cast<const int4*>(D + oModelShade);
    sc.triShade = reinterpret_cast<const float4*>(D + oTriShade);
    sc.materials = reinterpret_cast<const DeviceMaterial*>(D + oMaterials);
    sc.matTex = hasTextures ? reinterpret_cast<const int4*>(D + oMatTex) : nullptr;
    sc.texProg = hasTextures ? reinterpret_cast<const unsigned int*>(D + oTexProg) : nullptr;
    sc.texNodes = hasTextures ? reinterpret_cast<const float4*>(D + oTexNodes) : nullptr;
    sc.texLevels = hasTextures ? reinterpret_cast<const int4*>(D + oTexLevels) : nullptr;
    sc.matMask = hasMask ? reinterpret_cast<const int4*>(D + oMatMask) : nullptr;
    sc.lights = reinterpret_cast<const DeviceLight*>(D + oLights);
    sc.lightPower = reinterpret_cast<const float*>(D + oLightPower);
    sc.lightCdf = reinterpret_cast<const float*>(D + oLightCdf);
    sc.lightTris = reinterpret_cast<const float4*>(D + oLightTris);
    sc.lightTriCdf = reinterpret_cast<const float*>(D + oLightTriCdf);
    sc.filterTable = reinterpret_cast<const float*>(D + oFilter);
    sc.nTopNodes = d->n_top_nodes;
    sc.topRootRef = topRootRef;
    sc.nInstances = nInst;
    sc.nLights = d->n_lights;
    sc.lightIntegral = integral;
    sc.hasAreaLight = hasArea ? 1u : 0u;
    sc.hasEnvLight = hasEnvLight ? 1u : 0u;
    sc.imageTexels = reinterpret_cast<const float4*>(D + oImageTexels);
    sc.lightDist = reinterpret_cast<const float*>(D + oLightDist);
    sc.camera = d->camera;
    sc.xres = f.xres; sc.yres = f.yres;
    sc.xstart = f.xstart; sc.xcount = f.xcount; sc.ystart = f.ystart; sc.ycount = f.ycount;
    sc.sx0 = f.sx0; sc.sx1 = f.sx1; sc.sy0 = f.sy0; sc.sy1 = f.sy1;
    sc.invXRes = 1.0f / (float)f.xres; // Film::mInvXRes
    sc.invYRes = 1.0f / (float)f.yres;
    sc.filterWidthX = f.filter_width[0];
    sc.filterWidthY = f.filter_width[1];
    const size_t filmPixels = (size_t)f.xres * f.yres;
    bool newFilm = false;
    if (filmPixels != ctx->filmPixels || !ctx->film) {
        GB_CUDA(cudaStreamSynchronize(ctx->stream)); // kernels may still add to the old film
        if (ctx->film) cudaFree(ctx->film);
        ctx->film = nullptr;
        ctx->filmPixels = 0;
        GB_CUDA(cudaMalloc((void**)&ctx->film, filmPixels * sizeof(float4)));
        ctx->filmPixels = filmPixels;
        newFilm = true;
    }
    if (!async) {
        GB_CUDA(cudaStreamSynchronize(cs)); // the staging arena may be refilled after this
        if (*slot.deriveErrorHost == 1) return gb::failWith(GB_ERR_INVALID, "model_order entry out of range");
        if (*slot.deriveErrorHost == 2) return gb::failWith(GB_ERR_INVALID, "vertex index out of range");
        slot.checkPending = false;
    } else {
        // what is queued on the context's stream so far read the current slot: the NEXT upload into it waits for this
        // marker; what is queued from now on reads the new slot, once its copy and derive kernels are through
        GB_CUDA(cudaEventRecord(ctx->slots[ctx->slot].lastUse, ctx->stream));
        GB_CUDA(cudaStreamWaitEvent(ctx->stream, slot.uploaded, 0));
        slot.checkPending = true;
    }
    // the synchronous call hands back an empty film (as it always did); the asynchronous one leaves the film alone:
    // in a pipeline the previous frame is still waiting there to be downloaded
    if (!async || newFilm) GB_CUDA(cudaMemsetAsync(ctx->film, 0, ctx->filmPixels * sizeof(float4), ctx->stream));
    ctx->filmDesc = d->film;
    ctx->slot = k;
    ctx->sc = sc;
    ctx->setting = d->setting;
    ctx->matBins = matBins;
    ctx->hasMeshLight = hasMeshLight;
    ctx->wideFits = wideFits;
    ctx->stackEntries = stackEntries;
    ctx->stackEntriesPair = stackEntriesPair;
    ctx->gridCache.clear();
    lap("copy+sync");
    ctx->haveScene = true; // every check has passed
    return GB_OK;
}

extern "C" int gb_upload_scene(gb_context* ctx, const gb_scene_desc* d) { return uploadScene(ctx, d, false); }

extern "C" int gb_upload_scene_async(gb_context* ctx, const gb_scene_desc* d) { return uploadScene(ctx, d, true); }

extern "C" {

static int launchTrace(gb_context* ctx, bool any, const gb_ray* d_rays, size_t n, gb_hit* d_hits, unsigned char* d_occ) {
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    const int mode = walkMode(ctx);
    const size_t smem = traceSmem(ctx, mode);
    const int stackEntries = stackEntriesOf(ctx, mode);
    int grid = 0, rc;
    GB_CUDA(cudaMemsetAsync(ctx->traceHead, 0, 8, ctx->stream));
    GB_CUDA(cudaEventRecord(ctx->evStart, ctx->stream));
#define LAUNCH(ANYV, MODEV)                                                                        \
    do {                                                                                           \
        if ((rc = setupTraceKernel(ctx, k_trace<ANYV, MODEV>, MODEV, &grid)) != GB_OK) return rc;   \
        KernelTick tick(ctx, GB_K_TRACE);                                                           \
        k_trace<ANYV, MODEV><<<grid, kTraceBlock, smem, ctx->stream>>>(ctx->sc, d_rays, (unsigned long long)n, d_hits, \
            d_occ, ctx->traceHead, ctx->stats, stackEntries);                                       \
    } while (0)
    if (any) {
        if (mode == WALK_WIDE) LAUNCH(true, WALK_WIDE); else if (mode == WALK_PAIR) LAUNCH(true, WALK_PAIR); else LAUNCH(true, WALK_STATS);
    } else {
        if (mode == WALK_WIDE) LAUNCH(false, WALK_WIDE); else if (mode == WALK_PAIR) LAUNCH(false, WALK_PAIR); else LAUNCH(false, WALK_STATS);
    }
#undef LAUNCH
    ctx->launches++;
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaEventRecord(ctx->evStop, ctx->stream));
    return GB_OK;
}

int gb_trace_closest_device(gb_context* ctx, const gb_ray* d_rays, size_t n, gb_hit* d_hits) {
    if (!ctx || (n && (!d_rays || !d_hits))) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    return launchTrace(ctx, false, d_rays, n, d_hits, nullptr);
}

int gb_trace_any_device(gb_context* ctx, const gb_ray* d_rays, size_t n, uint8_t* d_occ) {
    if (!ctx || (n && (!d_rays || !d_occ))) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    return launchTrace(ctx, true, d_rays, n, nullptr, d_occ);
}

int gb_trace_closest(gb_context* ctx, const gb_ray* rays, size_t n, gb_hit* hits) {
    if (!ctx || (n && (!rays || !hits))) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    if (n == 0) return GB_OK;
    GB_CUDA(cudaSetDevice(ctx->device));
    gb_ray* d_rays = nullptr;
    gb_hit* d_hits = nullptr;
    GB_CUDA(cudaMalloc((void**)&d_rays, n * sizeof(gb_ray)));
    cudaError_t e = cudaMalloc((void**)&d_hits, n * sizeof(gb_hit));
    if (e != cudaSuccess) { cudaFree(d_rays); return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e)); }
    int rc = GB_OK;
    e = cudaMemcpyAsync(d_rays, rays, n * sizeof(gb_ray), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = launchTrace(ctx, false, d_rays, n, d_hits, nullptr);
    if (e == cudaSuccess && rc == GB_OK) e = cudaMemcpyAsync(hits, d_hits, n * sizeof(gb_hit), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == GB_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rays);
    cudaFree(d_hits);
    if (rc != GB_OK) return rc;
    if (e != cudaSuccess) return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e));
    return GB_OK;
}

int gb_trace_any(gb_context* ctx, const gb_ray* rays, size_t n, uint8_t* occluded) {
    if (!ctx || (n && (!rays || !occluded))) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    if (n == 0) return GB_OK;
    GB_CUDA(cudaSetDevice(ctx->device));
    gb_ray* d_rays = nullptr;
    unsigned char* d_occ = nullptr;
    GB_CUDA(cudaMalloc((void**)&d_rays, n * sizeof(gb_ray)));
    cudaError_t e = cudaMalloc((void**)&d_occ, n);
    if (e != cudaSuccess) { cudaFree(d_rays); return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e)); }
    int rc = GB_OK;
    e = cudaMemcpyAsync(d_rays, rays, n * sizeof(gb_ray), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = launchTrace(ctx, true, d_rays, n, nullptr, d_occ);
    if (e == cudaSuccess && rc == GB_OK) e = cudaMemcpyAsync(occluded, d_occ, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == GB_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rays);
    cudaFree(d_occ);
    if (rc != GB_OK) return rc;
    if (e != cudaSuccess) return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e));
    return GB_OK;
}

int gb_camera_rays(gb_context* ctx, const float* samples, size_t n, gb_ray* rays) {
    if (!ctx || (n && (!samples || !rays))) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    if (n == 0) return GB_OK;
    GB_CUDA(cudaSetDevice(ctx->device));
    float* d_s = nullptr;
    gb_ray* d_r = nullptr;
    GB_CUDA(cudaMalloc((void**)&d_s, n * 4 * sizeof(float)));
    cudaError_t e = cudaMalloc((void**)&d_r, n * sizeof(gb_ray));
    if (e != cudaSuccess) { cudaFree(d_s); return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(d_s, samples, n * 4 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_camera_rays<<<(unsigned int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->sc, d_s, (unsigned int)n, d_r);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rays, d_r, n * sizeof(gb_ray), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_s);
    cudaFree(d_r);
    if (e != cudaSuccess) return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e));
    return GB_OK;
}

// One wave of the wavefront integrator: nPaths camera samples start to finish.
// table != nullptr: explicit sample values (gb_li); the wave is then a flat
// list of samples and the film is not touched.
static int runWave(gb_context* ctx, gb_context::WaveLane& lane, const WaveParams& wp, const SampleSource& src, int method,
    bool toFilm) {
    PathState& ps = lane.ps;
    cudaStream_t st = lane.stream;
    const int mode = walkMode(ctx);
    const size_t smem = traceSmem(ctx, mode);
    const int stackEntries = stackEntriesOf(ctx, mode);
    const unsigned int n = wp.nPaths;
    int rc, grid = 0;
    GB_CUDA(cudaMemsetAsync(lane.ctr, 0, kMaxDepthCtr * kCtrStride * sizeof(unsigned int), st));
    PathState psRay = ps;
    if (method != GB_METHOD_AO) psRay.aoCount = nullptr;
    {
        KernelTick tick(ctx, GB_K_RAYGEN, st);
        k_raygen<<<(n + 255) / 256, 256, 0, st>>>(ctx->sc, psRay, wp, src, lane.ctr, ctx->stats);
    }
    ctx->launches++;
    const int shadeGrid = ctx->numSMs * 8;
    // the one bin every hit of this scene lands in, if its materials are all of one type (else -1: look it up per hit)
    const int onlyBin = (ctx->matBins & (ctx->matBins - 1u)) == 0u && ctx->matBins ? __builtin_ctz(ctx->matBins) : -1;
    auto extend = [&](int b, int fixedBin) -> int {
        unsigned int* c = lane.ctr + b * kCtrStride;
        const unsigned int* q = b == 0 ? nullptr : ps.qExtend[b & 1];
        KernelTick tick(ctx, GB_K_EXTEND, st);
#define GB_EXTEND(MODEV)                                                                                            \
    do {                                                                                                            \
        if ((rc = setupTraceKernel(ctx, k_extend<MODEV>, MODEV, &grid)) != GB_OK) return rc;                         \
        k_extend<MODEV><<<grid, kTraceBlock, smem, st>>>(ctx->sc, ps, q, c, fixedBin, ctx->stats, stackEntries);     \
    } while (0)
        if (mode == WALK_WIDE) GB_EXTEND(WALK_WIDE); else if (mode == WALK_PAIR) GB_EXTEND(WALK_PAIR); else GB_EXTEND(WALK_STATS);
#undef GB_EXTEND
        ctx->launches++;
        return GB_OK;
    };
    if (method == GB_METHOD_AO) {
        if ((rc = extend(0, 0)) != GB_OK) return rc; // AO reads the hits in path order: one bin
        {
            KernelTick tick(ctx, GB_K_OTHER, st);
            if (ctx->sc.matTex) k_ao_frames<true><<<ctx->numSMs * 8, 256, 0, st>>>(ctx->sc, ps, wp, src, lane.ctr);
            else k_ao_frames<false><<<ctx->numSMs * 8, 256, 0, st>>>(ctx->sc, ps, wp, src, lane.ctr);
            ctx->launches++;
        }
        {
            KernelTick tick(ctx, GB_K_AO, st);
#define GB_AO(MODEV)                                                                                                  \
    do {                                                                                                              \
        if ((rc = setupTraceKernel(ctx, k_ao<MODEV>, MODEV, &grid)) != GB_OK) return rc;                               \
        k_ao<MODEV><<<grid, kTraceBlock, smem, st>>>(ctx->sc, ps, wp, src, lane.ctr, ctx->stats, stackEntries);         \
    } while (0)
            if (mode == WALK_WIDE) GB_AO(WALK_WIDE); else if (mode == WALK_PAIR) GB_AO(WALK_PAIR); else GB_AO(WALK_STATS);
#undef GB_AO
        }
        {
            KernelTick tick(ctx, GB_K_OTHER, st);
            k_ao_finish<<<(n + 255) / 256, 256, 0, st>>>(ps, wp);
        }
        ctx->launches += 2;
    } else if (ctx->sc.nLights > 0) { // no lights: Li returns black before tracing (GoblinPathtracer.cpp:53-56)
        const int depth = wp.maxDepth;
        // Tail overlap: the shadow kernel of bounce b and the extend kernel of bounce b + 1 touch disjoint
        // data (one adds to L, the other writes hits), and both are single-wave persistent grids whose last
        // CTAs run on a mostly idle machine.  The shadow kernel goes to a second stream, so the next
        // extend's CTAs start as its CTAs retire; the shade kernels of bounce b + 1 (which reuse the shadow
        // queue and may touch L) wait for it.  Not with an environment light (k_miss adds to L right after
        // the extend) and not while counting.
        const bool overlap = ctx->overlapTails && !ctx->statsOn && !ctx->sc.hasEnvLight && !ctx->sc.matMask;
        bool pendingJoin = false;
        cudaError_t joinStatus = cudaSuccess;
        auto join = [&]() {
            if (pendingJoin) {
                const cudaError_t e = cudaStreamWaitEvent(st, lane.evJoin, 0);
                if (e != cudaSuccess) joinStatus = e;
            }
            pendingJoin = false;
        };
        // every way out of the bounce loop (errors included) makes the context's stream wait for a shadow
        // kernel still queued on the second one: the next wave's memset / raygen must not overtake it
        struct JoinGuard {
            decltype(join)& f;
            ~JoinGuard() { f(); }
        } joinGuard{join};
        for (int b = 0; b < depth; ++b) {
            // the last extend only feeds the BSDF-sampled emission term; skip it when no area light exists
            const bool last = b == depth - 1;
            if (last && b > 0 && !ctx->sc.hasAreaLight && !ctx->sc.hasEnvLight) break;
            if ((rc = extend(b, onlyBin)) != GB_OK) return rc;
            if (ctx->sc.hasEnvLight) { // rays that left the scene pick up the environment map
                KernelTick tick(ctx, GB_K_SHADE, st);
                k_miss<<<shadeGrid, 256, 0, st>>>(ctx->sc, ps, b == 0 ? nullptr : ps.qExtend[b & 1], lane.ctr + b * kCtrStride, b);
                ctx->launches++;
            }
            unsigned int* c = lane.ctr + b * kCtrStride;
            unsigned int* cn = lane.ctr + (b + 1) * kCtrStride;
            unsigned int* qn = ps.qExtend[(b + 1) & 1];
            int eo = last ? 1 : 0;
            join(); // the previous bounce's shadow kernel is done with the shadow queue and L
            {
                KernelTick tick(ctx, GB_K_SHADE, st);
#define GB_SHADE(MATV, MLV)                                                                                           \
    do {                                                                                                              \
        if (ctx->matBins & (1u << MATV)) {                                                                            \
            k_shade<MATV, MLV, false><<<shadeGrid, kShadeBlock, 0, st>>>(ctx->sc, ps, wp, src, b, eo, c, cn, qn);     \
            ctx->launches++;                                                                                          \
        }                                                                                                             \
    } while (0)
#define GB_SHADE_TEX(MATV)                                                                                            \
    do {                                                                                                              \
        if (ctx->matBins & (1u << MATV)) {                                                                            \
            k_shade<MATV, true, true><<<shadeGrid, kShadeBlock, 0, st>>>(ctx->sc, ps, wp, src, b, eo, c, cn, qn);     \
            ctx->launches++;                                                                                          \
        }                                                                                                             \
    } while (0)
                // one launch per material type the scene holds (a bin nobody can land in needs no kernel)
                if (ctx->sc.matTex) { // a textured material slot: the variants with the texture evaluator
                    GB_SHADE_TEX(GB_MAT_LAMBERT); GB_SHADE_TEX(GB_MAT_MIRROR); GB_SHADE_TEX(GB_MAT_TRANSPARENT);
                    GB_SHADE_TEX(GB_MAT_BLINN);
                } else if (ctx->hasMeshLight) {
                    GB_SHADE(GB_MAT_LAMBERT, true); GB_SHADE(GB_MAT_MIRROR, true); GB_SHADE(GB_MAT_TRANSPARENT, true);
                    GB_SHADE(GB_MAT_BLINN, true);
                } else {
                    GB_SHADE(GB_MAT_LAMBERT, false); GB_SHADE(GB_MAT_MIRROR, false); GB_SHADE(GB_MAT_TRANSPARENT, false);
                    GB_SHADE(GB_MAT_BLINN, false);
                }
#undef GB_SHADE
#undef GB_SHADE_TEX
            }
            if (!last && ctx->sc.matMask) { // masks: filtered shadow / MIS traces with attenuation (mask.cuh)
                KernelTick tick(ctx, GB_K_SHADOW, st);
                k_shadow_mask<<<shadeGrid, 128, 0, st>>>(ctx->sc, ps, cn, ctx->stats); // row b + 1: see C_SHADOW
                if (ctx->sc.hasAreaLight | ctx->sc.hasEnvLight) k_mis_mask<<<shadeGrid, 128, 0, st>>>(ctx->sc, ps, qn, cn, ctx->stats);
                ctx->launches += 2;
            } else if (!last) {
                cudaStream_t ss = overlap ? lane.stream2 : st;
                if (overlap) {
                    GB_CUDA(cudaEventRecord(lane.evFork, st));
                    GB_CUDA(cudaStreamWaitEvent(ss, lane.evFork, 0));
                }
                {
                    KernelTick tick(ctx, GB_K_SHADOW, ss);
#define GB_SHADOW(MODEV)                                                                                      \
    do {                                                                                                      \
        if ((rc = setupTraceKernel(ctx, k_shadow<MODEV>, MODEV, &grid)) != GB_OK) return rc;                   \
        k_shadow<MODEV><<<grid, kTraceBlock, smem, ss>>>(ctx->sc, ps, cn, ctx->stats, stackEntries); /* row b + 1 */ \
    } while (0)
                    if (mode == WALK_WIDE) GB_SHADOW(WALK_WIDE); else if (mode == WALK_PAIR) GB_SHADOW(WALK_PAIR); else GB_SHADOW(WALK_STATS);
#undef GB_SHADOW
                }
                if (overlap) {
                    GB_CUDA(cudaEventRecord(lane.evJoin, ss));
                    pendingJoin = true;
                }
                ctx->launches++;
            }
        }
        join(); // the film kernel (or the caller) reads L
        if (joinStatus != cudaSuccess) return gb::failWith(GB_ERR_CUDA, std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(joinStatus));
    }
    if (toFilm) {
        const float wx = ctx->sc.filterWidthX, wy = ctx->sc.filterWidthY;
        const int radX = (int)ceilf(wx), radY = (int)ceilf(wy);
        auto pow2 = [](float v) { int e; return std::frexp(v, &e) == 0.5f; };
        const int invExact = pow2(wx) && pow2(wy) ? 1 : 0; // 16 * x / w == x * (16 / w) bit for bit
        const unsigned int nPix = (unsigned int)wp.width * (unsigned int)wp.rows;
        const unsigned int warpsPerBlock = kFilmBlock / 32;
        unsigned int blocks = std::min<unsigned int>((nPix + warpsPerBlock - 1) / warpsPerBlock, (unsigned int)ctx->numSMs * 16u);
        {
            KernelTick tick(ctx, GB_K_FILM, st);
            if (GB_FILM_WARP && !ctx->generalFilm && (2 * radX + 1) * (2 * radY + 1) <= 32) { // the candidate window fits one warp
                if (invExact) k_film_warp<true><<<blocks, kFilmBlock, 0, st>>>(ctx->sc, ps, wp, src, ctx->film, radX, radY);
                else k_film_warp<false><<<blocks, kFilmBlock, 0, st>>>(ctx->sc, ps, wp, src, ctx->film, radX, radY);
            } else {
                k_film<<<blocks, kFilmBlock, 0, st>>>(ctx->sc, ps, wp, src, ctx->film, radX, radY, invExact);
            }
        }
        ctx->launches++;
    }
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

int gb_render(gb_context* ctx, const gb_render_params* p) {
    if (!ctx || !p) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    GB_CUDA(cudaSetDevice(ctx->device));
    int method = p->method >= 0 ? p->method : ctx->setting.method;
    if (method != GB_METHOD_PATH_TRACING && method != GB_METHOD_AO) {
        return gb::failWith(GB_ERR_INVALID, "render_method is outside the accelerated path (path_tracing and ao only)");
    }
    int depth = p->max_ray_depth > 0 ? p->max_ray_depth : ctx->setting.max_ray_depth;
    if (depth < 1) depth = 1;
    if (depth > kMaxDepthCtr - 2) return gb::failWith(GB_ERR_LIMIT, "max_ray_depth above 64");
    int ao = p->ao_sample_num > 0 ? p->ao_sample_num : ctx->setting.ao_sample_num;
    if (ao < 1) ao = 1;
    { int r = (int)ceilf(sqrtf((float)ao)); ao = r * r; } // SampleQuota::requestTwoDQuota rounds up to a square
    int sppTotal = p->spp_total;
    int root = (int)ceilf(sqrtf((float)sppTotal)); // roundToSquare, GoblinUtils.h:126-132
    if (sppTotal < 1 || root * root != sppTotal) {
        return gb::failWith(GB_ERR_INVALID, "spp_total must be a positive perfect square (sample_per_pixel rounded up)");
    }
    if (p->spp_begin < 0 || p->spp_end > sppTotal || p->spp_begin > p->spp_end) {
        return gb::failWith(GB_ERR_INVALID, "bad sample index range");
    }
    const DeviceScene& sc = ctx->sc;
    const int width = sc.sx1 - sc.sx0, height = sc.sy1 - sc.sy0;
    if (width <= 0 || height <= 0 || p->spp_begin == p->spp_end) return GB_OK;
    SampleSource src;
    src.table = nullptr;
    src.rowFloats = 0;
    src.key = make_uint2((unsigned int)p->seed, (unsigned int)(p->seed >> 32));
    src.spp = (unsigned int)sppTotal;
    src.root = (unsigned int)root;
    src.invSpp = 1.0f / (float)sppTotal;
    src.invRoot = 1.0f / (float)root;
    GB_CUDA(cudaEventRecord(ctx->evStart, ctx->stream));
    // waves: whole rows x a slice of the sample indices, at most maxWavePaths paths
    int sppChunk = p->spp_end - p->spp_begin;
    if ((size_t)sppChunk * width > ctx->maxWavePaths) sppChunk = std::max<int>(1, (int)(ctx->maxWavePaths / width));
    int rowsPerWave = std::max<int>(1, (int)(ctx->maxWavePaths / ((size_t)sppChunk * width)));
    rowsPerWave = std::min(rowsPerWave, height);
    // Two lanes of waves side by side (gb_context::WaveLane): with a single wave the rows are cut in two.
    const int nLanes = (method == GB_METHOD_PATH_TRACING || method == GB_METHOD_AO) && !ctx->statsOn
        ? std::max(1, std::min(std::min(ctx->waveLanes, 4), height)) : 1;
    if (nLanes > 1 && p->spp_end - p->spp_begin <= sppChunk) rowsPerWave = std::min(rowsPerWave, (height + nLanes - 1) / nLanes);
    int rc = GB_OK;
    for (int l = 0; l < nLanes; ++l) {
        if ((rc = ensureWave(ctx->lanes[l], (size_t)rowsPerWave * width * sppChunk)) != GB_OK) return rc;
    }
    if (nLanes > 1) { // the other lanes start where the context's stream is now (behind a film clear, an upload ...)
        GB_CUDA(cudaEventRecord(ctx->evLaneStart, ctx->stream));
        for (int l = 1; l < nLanes; ++l) GB_CUDA(cudaStreamWaitEvent(ctx->lanes[l].stream, ctx->evLaneStart, 0));
    }
    int wave = 0;
    for (int s0 = p->spp_begin; s0 < p->spp_end && rc == GB_OK; s0 += sppChunk) {
        int ns = std::min(sppChunk, p->spp_end - s0);
        for (int y = 0; y < height && rc == GB_OK; y += rowsPerWave) {
            WaveParams wp;
            wp.rows = std::min(rowsPerWave, height - y);
            wp.y0 = sc.sy0 + y;
            wp.width = width;
            wp.sppBegin = s0;
            wp.nSpp = ns;
            wp.sppTotal = sppTotal;
            wp.root = root;
            wp.maxDepth = depth;
            wp.aoSamples = ao;
            wp.aoRoot = std::max(1, (int)sqrtf((float)ao));
            wp.nPaths = (unsigned int)((size_t)wp.rows * width * ns);
            rc = runWave(ctx, ctx->lanes[wave % nLanes], wp, src, method, true);
            ++wave;
        }
    }
    for (int l = 1; l < nLanes; ++l) { // the context's stream continues when every lane is through, errors included
        cudaEventRecord(ctx->lanes[l].evDone, ctx->lanes[l].stream);
        cudaStreamWaitEvent(ctx->stream, ctx->lanes[l].evDone, 0);
    }
    if (rc != GB_OK) return rc;
    GB_CUDA(cudaEventRecord(ctx->evStop, ctx->stream));
    return GB_OK;
}

int gb_li(gb_context* ctx, const float* samples, size_t n, size_t row_floats, float* out_rgb) {
    if (!ctx || (n && (!samples || !out_rgb))) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    if (n == 0) return GB_OK;
    GB_CUDA(cudaSetDevice(ctx->device));
    int method = ctx->setting.method;
    if (method != GB_METHOD_PATH_TRACING && method != GB_METHOD_AO) return gb::failWith(GB_ERR_INVALID, "unsupported render_method");
    int depth = std::max(1, ctx->setting.max_ray_depth);
    int ao = std::max(1, ctx->setting.ao_sample_num);
    size_t need = 4 + (method == GB_METHOD_AO ? 2 * (size_t)ao : 7 * (size_t)depth);
    if (row_floats < need) return gb::failWith(GB_ERR_INVALID, "sample rows too short for this integrator");
    if (depth > kMaxDepthCtr - 2) return gb::failWith(GB_ERR_LIMIT, "max_ray_depth above 64");
    float* d_s = nullptr;
    float* d_out = nullptr;
    GB_CUDA(cudaMalloc((void**)&d_s, n * row_floats * sizeof(float)));
    cudaError_t e = cudaMalloc((void**)&d_out, n * 3 * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_s); return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e)); }
    int rc = GB_OK;
    e = cudaMemcpyAsync(d_s, samples, n * row_floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    size_t chunk = std::min<size_t>(n, ctx->maxWavePaths);
    if (e == cudaSuccess) rc = ensureWave(ctx->lanes[0], chunk);
    for (size_t off = 0; e == cudaSuccess && rc == GB_OK && off < n; off += chunk) {
        size_t cnt = std::min(chunk, n - off);
        WaveParams wp{};
        wp.nPaths = (unsigned int)cnt;
        wp.y0 = ctx->sc.sy0; wp.width = (int)cnt; wp.rows = 1;
        wp.sppBegin = 0; wp.nSpp = 1; wp.sppTotal = 1; wp.root = 1;
        wp.maxDepth = depth; wp.aoSamples = ao; wp.aoRoot = 1;
        SampleSource src;
        src.table = d_s + off * row_floats;
        src.rowFloats = (unsigned int)row_floats;
        src.key = make_uint2(0, 0);
        src.spp = src.root = 0u;
        src.invSpp = src.invRoot = 0.0f;
        rc = runWave(ctx, ctx->lanes[0], wp, src, method, false);
        if (rc == GB_OK) {
            k_copy_L<<<(unsigned int)((cnt + 255) / 256), 256, 0, ctx->stream>>>(ctx->lanes[0].ps, (unsigned int)cnt, d_out + 3 * off);
            ctx->launches++;
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess && rc == GB_OK) e = cudaMemcpyAsync(out_rgb, d_out, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == GB_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_s);
    cudaFree(d_out);
    if (rc != GB_OK) return rc;
    if (e != cudaSuccess) return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e));
    return GB_OK;
}

int gb_film_clear(gb_context* ctx) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaMemsetAsync(ctx->film, 0, ctx->filmPixels * sizeof(float4), ctx->stream));
    return GB_OK;
}

int gb_film_download(gb_context* ctx, float* rgbw) {
    if (!ctx || !rgbw) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    GB_CUDA(cudaSetDevice(ctx->device));
    // through a pinned staging buffer: a DMA at PCIe speed plus a host memcpy beats the driver's
    // staged copy into pageable memory (0.57 -> 0.2 ms for a 512 x 384 film)
    if (ctx->filmHostPixels < ctx->filmPixels) {
        if (ctx->filmHost) cudaFreeHost(ctx->filmHost);
        ctx->filmHost = nullptr;
        ctx->filmHostPixels = 0;
        if (cudaMallocHost((void**)&ctx->filmHost, ctx->filmPixels * sizeof(float4)) == cudaSuccess) ctx->filmHostPixels = ctx->filmPixels;
        else cudaGetLastError();
    }
    float* dst = ctx->filmHost ? ctx->filmHost : rgbw;
    GB_CUDA(cudaMemcpyAsync(dst, ctx->film, ctx->filmPixels * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (dst != rgbw) std::memcpy(rgbw, dst, ctx->filmPixels * sizeof(float4));
    return checkDeferredUpload(ctx);
}

int gb_film_upload(gb_context* ctx, const float* rgbw) {
    if (!ctx || !rgbw) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaMemcpyAsync(ctx->film, rgbw, ctx->filmPixels * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GB_OK;
}

int gb_film_device_ptr(gb_context* ctx, void** ptr, size_t* n_floats) {
    if (!ctx || !ptr || !n_floats) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    *ptr = ctx->film;
    *n_floats = ctx->filmPixels * 4;
    return GB_OK;
}

int gb_film_resolve(gb_context* ctx, float* rgb) {
    if (!ctx || !rgb) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    GB_CUDA(cudaSetDevice(ctx->device));
    const int w = ctx->sc.xres, h = ctx->sc.yres;
    const size_t n = ctx->filmPixels;
    const gb_film_desc& fd = ctx->filmDesc;
    const bool doBloom = fd.bloom_radius > 0.0f && fd.bloom_weight > 0.0f; // Film::writeImage, GoblinFilm.cpp:187-189
    const int fw = doBloom ? gb::bloomFilterWidth(fd.bloom_radius, w, h) : 0;
    float *d_rgb = nullptr, *d_out = nullptr, *d_tab = nullptr;
    GB_CUDA(cudaMalloc((void**)&d_rgb, n * 3 * sizeof(float)));
    auto release = [&]() { cudaFree(d_rgb); cudaFree(d_out); cudaFree(d_tab); };
    k_resolve<<<(unsigned int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->film, (unsigned int)n, d_rgb);
    ctx->launches++;
    const float* result = d_rgb;
    if (doBloom) {
        // fw == 0 (a radius below two pixels) divides 0 by 0 in the reference's table: NaN everywhere;
        // the same arithmetic runs here
        std::vector<float> table((size_t)std::max(fw, 1) * std::max(fw, 1), 0.0f);
        gb::bloomFilterTable(fw, table.data());
        cudaError_t e = cudaMalloc((void**)&d_out, n * 3 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_tab, table.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { release(); return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e)); }
        dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
        k_bloom<<<grd, blk, 0, ctx->stream>>>(d_rgb, d_out, w, h, fw, d_tab, fd.bloom_weight);
        ctx->launches++;
        result = d_out;
    }
    cudaError_t e = cudaMemcpyAsync(rgb, result, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    release();
    if (e != cudaSuccess) return gb::failWith(GB_ERR_CUDA, cudaGetErrorString(e));
    return GB_OK;
}

int gb_film_write(gb_context* ctx, const char* path) {
    if (!ctx || !path) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    std::vector<float> host(ctx->filmPixels * 3);
    int rc = gb_film_resolve(ctx, host.data());
    if (rc != GB_OK) return rc;
    std::string err;
    if (!gb::writeRgb(path, host.data(), ctx->sc.xres, ctx->sc.yres, ctx->filmDesc.tone_mapping != 0, &err)) {
        return gb::failWith(GB_ERR_IO, err);
    }
    return GB_OK;
}

int gb_synchronize(gb_context* ctx) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    return checkDeferredUpload(ctx);
}

int gb_stream(gb_context* ctx, void** cuda_stream) {
    if (!ctx || !cuda_stream) return gb::failWith(GB_ERR_INVALID, "null argument");
    *cuda_stream = ctx->stream;
    return GB_OK;
}

int gb_enable_counters(gb_context* ctx, int on) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    ctx->statsOn = on != 0;
    return GB_OK;
}

int gb_get_counters(gb_context* ctx, gb_counters* out) {
    if (!ctx || !out) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[S_COUNT];
    GB_CUDA(cudaMemcpy(h, ctx->stats, sizeof h, cudaMemcpyDeviceToHost));
    out->camera_samples = h[S_SAMPLES];
    out->rays_closest = h[S_RAYS_CLOSEST];
    out->rays_any = h[S_RAYS_ANY];
    out->nodes_visited = h[S_NODES] + h[S_ANY_BASE + S_NODES];
    out->prims_tested = h[S_PRIMS] + h[S_ANY_BASE + S_PRIMS];
    out->instances_entered = h[S_INSTS] + h[S_ANY_BASE + S_INSTS];
    out->nodes_visited_any = h[S_ANY_BASE + S_NODES];
    out->prims_tested_any = h[S_ANY_BASE + S_PRIMS];
    out->instances_entered_any = h[S_ANY_BASE + S_INSTS];
    out->kernel_launches = ctx->launches;
    return GB_OK;
}

int gb_reset_counters(gb_context* ctx) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    GB_CUDA(cudaMemset(ctx->stats, 0, S_COUNT * sizeof(unsigned long long)));
    ctx->launches = 0;
    return GB_OK;
}

int gb_enable_kernel_timing(gb_context* ctx, int on) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    collectTimes(ctx);
    ctx->timingOn = on != 0;
    return GB_OK;
}

int gb_get_kernel_times(gb_context* ctx, gb_kernel_times* out) {
    if (!ctx || !out) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    collectTimes(ctx);
    for (int k = 0; k < GB_K_COUNT; ++k) {
        out->ms[k] = ctx->classMs[k];
        out->launches[k] = ctx->classLaunches[k];
    }
    return GB_OK;
}

int gb_reset_kernel_times(gb_context* ctx) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    collectTimes(ctx);
    for (int k = 0; k < GB_K_COUNT; ++k) { ctx->classMs[k] = 0.0; ctx->classLaunches[k] = 0; }
    return GB_OK;
}

int gb_set_tuning(gb_context* ctx, const int* values, int n) {
    if (!ctx || (n && !values)) return gb::failWith(GB_ERR_INVALID, "null argument");
    unsigned int* t[4] = {&ctx->tune.refillBelow, &ctx->tune.leafBatch, &ctx->tune.levelBatch, &ctx->tune.moveFloor};
    for (int k = 0; k < n && k < 8; ++k) {
        if (values[k] < 0 || values[k] > 33) return gb::failWith(GB_ERR_INVALID, "tuning value out of range");
        if (k >= 4) continue;
        *t[k] = (unsigned int)values[k];
    }
    if (n > 4) ctx->blocksPerSM = values[4];
    if (n > 5) ctx->waveLanes = std::min(4, std::max(1, values[5]));
    if (n > 6) ctx->overlapTails = values[6] != 0;
    if (n > 7) ctx->generalFilm = values[7] != 0;
    ctx->sc.tune = ctx->tune;
    ctx->gridCache.clear();
    return GB_OK;
}

int gb_set_trace_mode(gb_context* ctx, int mode) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (mode != GB_TRACE_WIDE && mode != GB_TRACE_PAIR) return gb::failWith(GB_ERR_INVALID, "unknown trace mode");
    ctx->traceMode = mode;
    return GB_OK;
}

int gb_get_trace_mode(gb_context* ctx, int* mode) {
    if (!ctx || !mode) return gb::failWith(GB_ERR_INVALID, "null argument");
    // what the kernels actually run: a scene too deep for the wide walk's shared-memory stack is walked pair-wise
    *mode = ctx->traceMode == GB_TRACE_WIDE && (!ctx->haveScene || ctx->wideFits) ? GB_TRACE_WIDE : GB_TRACE_PAIR;
    return GB_OK;
}

int gb_debug_stack_violation(gb_context* ctx, int* out4) {
    if (!ctx || !out4) return gb::failWith(GB_ERR_INVALID, "null argument");
#if GB_DEBUG_STACK
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaDeviceSynchronize());
    GB_CUDA(cudaMemcpyFromSymbol(out4, gb::g_stackViolation, 4 * sizeof(int)));
    return GB_OK;
#else
    out4[0] = out4[1] = out4[2] = out4[3] = -1; // not a debug build
    return GB_OK;
#endif
}

int gb_upload_bytes(gb_context* ctx, size_t* bytes) {
    if (!ctx || !bytes) return gb::failWith(GB_ERR_INVALID, "null argument");
    *bytes = ctx->uploadBytes;
    return GB_OK;
}

int gb_set_wave_paths(gb_context* ctx, size_t max_paths) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (max_paths < 1024 || max_paths > (size_t)1 << 28) return gb::failWith(GB_ERR_INVALID, "wave size out of range");
    ctx->maxWavePaths = max_paths;
    return GB_OK;
}

int gb_last_kernel_ms(gb_context* ctx, float* ms) {
    if (!ctx || !ms) return gb::failWith(GB_ERR_INVALID, "null argument");
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_CUDA(cudaEventSynchronize(ctx->evStop));
    GB_CUDA(cudaEventElapsedTime(ms, ctx->evStart, ctx->evStop));
    return GB_OK;
}

} // extern "C"

#include "film_comm.inl"
