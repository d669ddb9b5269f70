/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product.
 *
 * ref_tool: a harness that links the UNMODIFIED reference objects (compiled
 * from /root/reference/src by oracle/Makefile, outputs in oracle/_ref/) and
 * drives the reference's own code to
 *   dump    flattened BVH nodes / primitive order / instance matrices / mesh
 *           arrays / light CDF / camera / filter table
 *   camrays Camera::generateRay on given image/lens samples
 *   trace   Scene-level closest-hit / any-hit on a ray batch, with the hit
 *           instance and primitive identified
 *   li      PathTracer::Li / AORenderer::Li on explicit sample values
 *   render  timed RenderContext::render() with Scene::intersect/occluded call
 *           counters (linker --wrap) and a raw float dump of the film
 *   post    Goblin::bloom / Goblin::toneMapping / the PPM writer on a raw rgb image
 *   session load the scene once, then run any of the above from stdin (large scenes)
 * Only this TU is compiled with -fno-access-control so it can read private
 * members; it observes the reference, it never reimplements it.
 *
 * Output container ("GBAR"): a flat sequence of named arrays, see put_array().
 */
#include "GoblinContextLoader.h"
#include "GoblinRenderContext.h"
#include "GoblinBVH.h"
#include "GoblinModel.h"
#include "GoblinPrimitive.h"
#include "GoblinPolygonMesh.h"
#include "GoblinTriangle.h"
#include "GoblinSphere.h"
#include "GoblinDisk.h"
#include "GoblinCamera.h"
#include "GoblinFilm.h"
#include "GoblinFilter.h"
#include "GoblinLight.h"
#include "GoblinSampler.h"
#include "GoblinPathtracer.h"
#include "GoblinAO.h"
#include "GoblinRay.h"
#include "GoblinThreadLocalStorage.h"
#include "GoblinImageIO.h"
#include "GoblinTexture.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <sstream>
#include <thread>

using namespace Goblin;

// ---------------------------------------------------------------- GBAR io
enum { DT_F32 = 0, DT_U32 = 1, DT_I32 = 2, DT_U8 = 3 };

static void put_array(FILE* f, const char* name, int dtype,
    const std::vector<uint64_t>& dims, const void* data) {
    static const size_t esize[4] = {4, 4, 4, 1};
    uint32_t nl = (uint32_t)strlen(name);
    fwrite("GBAR", 1, 4, f);
    fwrite(&nl, 4, 1, f);
    fwrite(name, 1, nl, f);
    uint32_t dt = dtype, nd = (uint32_t)dims.size();
    fwrite(&dt, 4, 1, f);
    fwrite(&nd, 4, 1, f);
    uint64_t n = 1;
    for (uint64_t d : dims) { fwrite(&d, 8, 1, f); n *= d; }
    if (n) fwrite(data, esize[dtype], n, f);
}

static void put_f32(FILE* f, const std::string& name, const std::vector<float>& v,
    std::vector<uint64_t> dims = {}) {
    if (dims.empty()) dims = {v.size()};
    put_array(f, name.c_str(), DT_F32, dims, v.data());
}
static void put_u32(FILE* f, const std::string& name, const std::vector<uint32_t>& v,
    std::vector<uint64_t> dims = {}) {
    if (dims.empty()) dims = {v.size()};
    put_array(f, name.c_str(), DT_U32, dims, v.data());
}
static void put_i32(FILE* f, const std::string& name, const std::vector<int32_t>& v,
    std::vector<uint64_t> dims = {}) {
    if (dims.empty()) dims = {v.size()};
    put_array(f, name.c_str(), DT_I32, dims, v.data());
}

static std::vector<float> read_f32_file(const char* path) {
    std::vector<float> out;
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz / 4);
    if (fread(out.data(), 4, out.size(), f) != out.size()) { exit(2); }
    fclose(f);
    return out;
}

// ------------------------------------------------- call counters (--wrap)
static std::atomic<uint64_t> g_intersectCalls(0), g_occludedCalls(0);
struct TLCount {
    uint64_t i = 0, o = 0;
    ~TLCount() { g_intersectCalls += i; g_occludedCalls += o; }
    void flush() { g_intersectCalls += i; g_occludedCalls += o; i = o = 0; }
};
static thread_local TLCount tl_count;

// optional recording of every ray handed to Scene::intersect / occluded
struct RecordedRay { float v[8]; uint32_t kind; };
static bool g_record = false;
static thread_local std::vector<RecordedRay>* tl_rec = nullptr;
// per-thread recording caps per kind (0 = intersect, 1 = occluded); ~0 = unlimited
static uint64_t g_recordCap[2] = {~0ull, ~0ull};
static thread_local uint64_t tl_recorded[2] = {0, 0};
static std::vector<RecordedRay> g_recorded;

extern "C" {
bool __real__ZNK6Goblin5Scene9intersectERKNS_3RayEPfPNS_12IntersectionEPFbPKNS_9PrimitiveES3_E(
    const Scene*, const Ray&, float*, Intersection*, IntersectFilter);
bool __real__ZNK6Goblin5Scene8occludedERKNS_3RayEPFbPKNS_9PrimitiveES3_E(
    const Scene*, const Ray&, IntersectFilter);

bool __wrap__ZNK6Goblin5Scene9intersectERKNS_3RayEPfPNS_12IntersectionEPFbPKNS_9PrimitiveES3_E(
    const Scene* s, const Ray& r, float* e, Intersection* is, IntersectFilter f) {
    tl_count.i++;
    if (g_record && tl_rec && tl_recorded[0] < g_recordCap[0]) {
        RecordedRay rr = {{r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.mint, r.maxt}, 0u};
        tl_rec->push_back(rr);
        tl_recorded[0]++;
    }
    return __real__ZNK6Goblin5Scene9intersectERKNS_3RayEPfPNS_12IntersectionEPFbPKNS_9PrimitiveES3_E(
        s, r, e, is, f);
}
bool __wrap__ZNK6Goblin5Scene8occludedERKNS_3RayEPFbPKNS_9PrimitiveES3_E(
    const Scene* s, const Ray& r, IntersectFilter f) {
    tl_count.o++;
    if (g_record && tl_rec && tl_recorded[1] < g_recordCap[1]) {
        RecordedRay rr = {{r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.mint, r.maxt}, 1u};
        tl_rec->push_back(rr);
        tl_recorded[1]++;
    }
    return __real__ZNK6Goblin5Scene8occludedERKNS_3RayEPFbPKNS_9PrimitiveES3_E(s, r, f);
}
}

// ----------------------------------------------------- scene introspection
struct SceneView {
    RenderContext* ctx = nullptr;
    const Scene* scene = nullptr;
    // original instance order: [camera lens] + JSON instances + area-light
    // instances (GoblinContextLoader.cpp:161-163, 381-383, 437-439)
    std::vector<const InstancedPrimitive*> instances;
    std::map<const Primitive*, int> instanceId;
    std::vector<const Model*> models; // distinct, first-appearance order
    std::map<const Primitive*, int> modelId;
};

static int geometryKind(const Geometry* g) {
    if (dynamic_cast<const PolygonMesh*>(g)) return 0;
    if (dynamic_cast<const Sphere*>(g)) return 1;
    if (dynamic_cast<const Disk*>(g)) return 2;
    if (dynamic_cast<const Triangle*>(g)) return 3;
    return -1;
}

static void buildView(SceneView& v, RenderContext* ctx) {
    v.ctx = ctx;
    v.scene = ctx->mScene.get();
    const Scene* s = v.scene;
    std::map<const Primitive*, bool> seen;
    for (Primitive* p : s->mPrimitives) {
        const InstancedPrimitive* ip = dynamic_cast<const InstancedPrimitive*>(p);
        if (ip) { v.instances.push_back(ip); seen[ip] = true; }
    }
    // area-light instances are only reachable through the scene BVH
    for (const Light* l : s->mLights) {
        const AreaLight* al = dynamic_cast<const AreaLight*>(l);
        if (!al) continue;
        for (const Primitive* p : s->mBVH.mRefinedPrimitives) {
            if (seen.count(p)) continue;
            const InstancedPrimitive* ip = dynamic_cast<const InstancedPrimitive*>(p);
            if (!ip) continue;
            const Model* m = dynamic_cast<const Model*>(ip->mPrimitive);
            if (m && m->mAreaLight == al) {
                v.instances.push_back(ip);
                seen[ip] = true;
                break;
            }
        }
    }
    if (v.instances.size() != s->mBVH.mRefinedPrimitives.size()) {
        fprintf(stderr, "instance recovery mismatch %zu vs %zu\n",
            v.instances.size(), s->mBVH.mRefinedPrimitives.size());
        exit(3);
    }
    for (size_t i = 0; i < v.instances.size(); ++i) {
        v.instanceId[v.instances[i]] = (int)i;
        const Primitive* m = v.instances[i]->mPrimitive;
        if (!v.modelId.count(m)) {
            v.modelId[m] = (int)v.models.size();
            v.models.push_back(dynamic_cast<const Model*>(m));
        }
    }
}

static void putNodes(FILE* f, const std::string& name, const BVH& bvh) {
    static_assert(sizeof(CompactBVHNode) == 32, "node size");
    std::vector<uint64_t> dims = {bvh.mBVHNodes.size(), 32};
    put_array(f, name.c_str(), DT_U8, dims, bvh.mBVHNodes.data());
}

static int lightIndexOf(const Scene* s, const Light* l) {
    for (size_t i = 0; i < s->mLights.size(); ++i) {
        if (s->mLights[i] == l) return (int)i;
    }
    return -1;
}

static int cmdDump(SceneView& v, const char* outPath) {
    FILE* f = fopen(outPath, "wb");
    if (!f) return 2;
    const Scene* s = v.scene;
    // --- top level
    putNodes(f, "top.nodes", s->mBVH);
    std::vector<uint32_t> order;
    for (const Primitive* p : s->mBVH.mRefinedPrimitives) {
        order.push_back((uint32_t)v.instanceId[p]);
    }
    put_u32(f, "top.order", order);
    // --- instances
    std::vector<float> fwd, inv, aabb;
    std::vector<int32_t> imodel;
    for (const InstancedPrimitive* ip : v.instances) {
        const Matrix4& M = ip->mToWorld.getMatrix();
        const Matrix4& I = ip->mToWorld.getInverse();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) fwd.push_back(M[r][c]);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) inv.push_back(I[r][c]);
        BBox b = ip->getAABB();
        aabb.insert(aabb.end(), {b.pMin.x, b.pMin.y, b.pMin.z, b.pMax.x, b.pMax.y, b.pMax.z});
        imodel.push_back(v.modelId[ip->mPrimitive]);
    }
    put_f32(f, "inst.fwd", fwd, {v.instances.size(), 16});
    put_f32(f, "inst.inv", inv, {v.instances.size(), 16});
    put_f32(f, "inst.aabb", aabb, {v.instances.size(), 6});
    put_i32(f, "inst.model", imodel);
    // --- models
    std::vector<int32_t> mkind, mlight;
    std::vector<float> mradius;
    for (size_t mi = 0; mi < v.models.size(); ++mi) {
        const Model* m = v.models[mi];
        int kind = geometryKind(m->mGeometry);
        mkind.push_back(kind);
        mlight.push_back(m->mAreaLight ? lightIndexOf(s, m->mAreaLight) : -1);
        float radius = 0.0f;
        if (kind == 1) radius = static_cast<const Sphere*>(m->mGeometry)->mRadius;
        if (kind == 2) radius = static_cast<const Disk*>(m->mGeometry)->mRadius;
        mradius.push_back(radius);
        std::ostringstream pre;
        pre << "model" << mi << ".";
        if (kind == 0) {
            const PolygonMesh* mesh = static_cast<const PolygonMesh*>(m->mGeometry);
            if (m->mBVH) {
                putNodes(f, pre.str() + "nodes", *m->mBVH);
                std::vector<uint32_t> tri;
                for (const Primitive* p : m->mBVH->mRefinedPrimitives) {
                    const Model* leaf = static_cast<const Model*>(p);
                    tri.push_back((uint32_t)static_cast<const Triangle*>(leaf->mGeometry)->mIndex);
                }
                put_u32(f, pre.str() + "order", tri);
            }
            std::vector<float> pos, nrm, uv;
            for (const Vertex& vx : mesh->mVertices) {
                pos.insert(pos.end(), {vx.position.x, vx.position.y, vx.position.z});
                nrm.insert(nrm.end(), {vx.normal.x, vx.normal.y, vx.normal.z});
                uv.insert(uv.end(), {vx.texC.x, vx.texC.y});
            }
            std::vector<uint32_t> idx;
            for (const TriangleIndex& t : mesh->mTriangles) {
                idx.insert(idx.end(), {t.v[0], t.v[1], t.v[2]});
            }
            put_f32(f, pre.str() + "pos", pos, {mesh->mVertices.size(), 3});
            put_f32(f, pre.str() + "nrm", nrm, {mesh->mVertices.size(), 3});
            put_f32(f, pre.str() + "uv", uv, {mesh->mVertices.size(), 2});
            put_u32(f, pre.str() + "idx", idx, {mesh->mTriangles.size(), 3});
            std::vector<int32_t> flags = {mesh->hasNormal() ? 1 : 0, mesh->hasTexCoord() ? 1 : 0};
            put_i32(f, pre.str() + "flags", flags);
            BBox ob = mesh->getObjectBound();
            put_f32(f, pre.str() + "bound",
                {ob.pMin.x, ob.pMin.y, ob.pMin.z, ob.pMax.x, ob.pMax.y, ob.pMax.z});
        }
    }
    put_i32(f, "model.kind", mkind);
    put_i32(f, "model.light", mlight);
    put_f32(f, "model.radius", mradius);
    // --- lights
    if (s->mPowerDistribution) {
        put_f32(f, "light.power", s->mPowerDistribution->mFunction);
        put_f32(f, "light.cdf", s->mPowerDistribution->mCDF);
    }
    // image based lights: radiance pyramid, orientation, sampling tables (include/goblin_b200.h layout)
    for (size_t li = 0; li < s->getLights().size(); ++li) {
        const ImageBasedLight* ibl = dynamic_cast<const ImageBasedLight*>(s->getLights()[li]);
        if (!ibl) continue;
        std::string pre = "ibl" + std::to_string(li) + ".";
        const MIPMap<Color>* mm = ibl->mRadiance;
        std::vector<int32_t> sizes;
        for (int l = 0; l < mm->mLevelsNum; ++l) {
            const ImageBuffer<Color>* b = mm->mPyramid[l];
            sizes.push_back(b->width); sizes.push_back(b->height);
            std::vector<float> px((size_t)b->width * b->height * 4);
            memcpy(px.data(), b->image, px.size() * 4);
            put_f32(f, pre + "level" + std::to_string(l), px);
        }
        put_i32(f, pre + "sizes", sizes);
        put_f32(f, pre + "average", {ibl->mAverageRadiance.r, ibl->mAverageRadiance.g, ibl->mAverageRadiance.b});
        const Matrix4& M = ibl->mToWorld.getMatrix();
        const Matrix4& I = ibl->mToWorld.getInverse();
        std::vector<float> fw, inv;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) { fw.push_back(M[r][c]); inv.push_back(I[r][c]); }
        put_f32(f, pre + "to_world", fw);
        put_f32(f, pre + "to_object", inv);
        const CDF2D* d2 = ibl->mDistribution;
        std::vector<float> table;
        int h = (int)d2->mConditionalDist.size(), w = d2->mConditionalDist[0]->mCount;
        for (int r = 0; r < h; ++r) table.insert(table.end(), d2->mConditionalDist[r]->mFunction.begin(), d2->mConditionalDist[r]->mFunction.end());
        for (int r = 0; r < h; ++r) table.insert(table.end(), d2->mConditionalDist[r]->mCDF.begin(), d2->mConditionalDist[r]->mCDF.end());
        table.insert(table.end(), d2->mMarginalDist->mFunction.begin(), d2->mMarginalDist->mFunction.end());
        table.insert(table.end(), d2->mMarginalDist->mCDF.begin(), d2->mMarginalDist->mCDF.end());
        table.push_back(d2->mMarginalDist->mIntegral);
        put_f32(f, pre + "dist", table);
        put_i32(f, pre + "dist_size", {w, h});
    }
    // --- camera / film / filter
    const CameraPtr cam = s->getCamera();
    Film* film = cam->getFilm();
    const PerspectiveCamera* pc = dynamic_cast<const PerspectiveCamera*>(cam.get());
    std::vector<float> camv = {cam->mPosition.x, cam->mPosition.y, cam->mPosition.z,
        cam->mOrientation.w, cam->mOrientation.v.x, cam->mOrientation.v.y, cam->mOrientation.v.z,
        cam->mProj[0][0], cam->mProj[1][1],
        pc ? pc->mLensRadius : 0.0f, pc ? pc->mFocalDistance : 0.0f};
    put_f32(f, "camera", camv);
    SampleRange sr;
    film->getSampleRange(sr);
    std::vector<int32_t> filmv = {film->mXRes, film->mYRes, film->mXStart, film->mXCount,
        film->mYStart, film->mYCount, sr.xStart, sr.xEnd, sr.yStart, sr.yEnd};
    put_i32(f, "film", filmv);
    std::vector<float> table(film->mCachedFilter.mTable,
        film->mCachedFilter.mTable + FILTER_TABLE_WIDTH * FILTER_TABLE_WIDTH);
    put_f32(f, "filter.table", table);
    put_f32(f, "filter.width", {film->mCachedFilter.mFilterWidth.x, film->mCachedFilter.mFilterWidth.y});
    fclose(f);
    return 0;
}

// -------------------------------------------------- tagged top-level trace
static thread_local int tl_lastInstance = -1;

struct TagPrim : public Primitive {
    const Primitive* inner;
    int id;
    TagPrim(const Primitive* p, int i) : inner(p), id(i) {}
    bool intersect(const Ray& r, float* eps, Intersection* is, IntersectFilter f) const override {
        bool h = inner->intersect(r, eps, is, f);
        if (h) tl_lastInstance = id;
        return h;
    }
    bool occluded(const Ray& r, IntersectFilter f) const override { return inner->occluded(r, f); }
    BBox getAABB() const override { return inner->getAABB(); }
};

struct Tracer {
    std::vector<TagPrim*> tags;
    std::unique_ptr<BVH> bvh;
    const SceneView& v;
    explicit Tracer(const SceneView& view) : v(view) {
        PrimitiveList list;
        for (size_t i = 0; i < v.instances.size(); ++i) {
            tags.push_back(new TagPrim(v.instances[i], (int)i));
            list.push_back(tags.back());
        }
        // same class, same inputs, same order as Scene::Scene (GoblinScene.cpp:15)
        bvh.reset(new BVH(list, 1, "equal_count"));
        if (bvh->mBVHNodes.size() != v.scene->mBVH.mBVHNodes.size() ||
            memcmp(bvh->mBVHNodes.data(), v.scene->mBVH.mBVHNodes.data(),
                bvh->mBVHNodes.size() * sizeof(CompactBVHNode)) != 0) {
            fprintf(stderr, "tagged BVH differs from scene BVH\n");
            exit(3);
        }
    }
    int primIndex(int inst, const Primitive* leaf) const {
        const Model* m = static_cast<const Model*>(v.instances[inst]->mPrimitive);
        if (m->mRefinedModels.empty()) return 0;
        const Model* first = &m->mRefinedModels.front();
        const Model* l = static_cast<const Model*>(leaf);
        long i = l - first;
        if (i < 0 || i >= (long)m->mRefinedModels.size()) return -2;
        return (int)static_cast<const Triangle*>(l->mGeometry)->mIndex;
    }
};

static int cmdTrace(SceneView& v, const char* raysPath, const char* outPath, bool withFrag = true) {
    std::vector<float> rays = read_f32_file(raysPath);
    size_t n = rays.size() / 8;
    Tracer tr(v);
    std::vector<int32_t> hit(n), inst(n), prim(n), occ(n);
    std::vector<float> t(n), eps(n), frag(withFrag ? n * 14 : 0);
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (unsigned k = 0; k < nt; ++k) {
        pool.emplace_back([&, k]() {
            for (size_t i = k; i < n; i += nt) {
                const float* r = &rays[8 * i];
                Ray ray(Vector3(r[0], r[1], r[2]), Vector3(r[3], r[4], r[5]), r[6], r[7]);
                // any-hit first (does not modify the ray)
                occ[i] = v.scene->occluded(ray) ? 1 : 0;
                Intersection is;
                float e = 0.0f;
                tl_lastInstance = -1;
                bool h = tr.bvh->intersect(ray, &e, &is, nullptr);
                hit[i] = h ? 1 : 0;
                inst[i] = h ? tl_lastInstance : -1;
                prim[i] = h ? tr.primIndex(tl_lastInstance, is.primitive) : -1;
                t[i] = h ? ray.maxt : 0.0f;
                eps[i] = h ? e : 0.0f;
                if (h && withFrag) {
                    const Fragment& fr = is.fragment;
                    float* o = &frag[14 * i];
                    const Vector3& p = fr.getPosition();
                    const Vector3& nn = fr.getNormal();
                    const Vector2& uv = fr.getUV();
                    const Vector3& du = fr.getDPDU();
                    const Vector3& dv = fr.getDPDV();
                    float vals[14] = {p.x, p.y, p.z, nn.x, nn.y, nn.z, uv.x, uv.y,
                        du.x, du.y, du.z, dv.x, dv.y, dv.z};
                    memcpy(o, vals, sizeof(vals));
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    FILE* f = fopen(outPath, "wb");
    if (!f) return 2;
    put_i32(f, "hit", hit);
    put_i32(f, "inst", inst);
    put_i32(f, "prim", prim);
    put_i32(f, "occluded", occ);
    put_f32(f, "t", t);
    put_f32(f, "eps", eps);
    if (withFrag) put_f32(f, "frag", frag, {n, 14});
    fclose(f);
    return 0;
}

// samples file for camrays: n x 4 floats (imageX, imageY, lensU1, lensU2)
static int cmdCamRays(SceneView& v, const char* samplesPath, const char* outPath) {
    std::vector<float> sm = read_f32_file(samplesPath);
    size_t n = sm.size() / 4;
    std::vector<float> out(n * 8);
    const CameraPtr cam = v.scene->getCamera();
    for (size_t i = 0; i < n; ++i) {
        Sample s;
        s.imageX = sm[4 * i];
        s.imageY = sm[4 * i + 1];
        s.lensU1 = sm[4 * i + 2];
        s.lensU2 = sm[4 * i + 3];
        RayDifferential ray;
        cam->generateRay(s, &ray);
        float vals[8] = {ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, ray.mint, ray.maxt};
        memcpy(&out[8 * i], vals, sizeof(vals));
    }
    FILE* f = fopen(outPath, "wb");
    if (!f) return 2;
    put_f32(f, "rays", out, {n, 8});
    fclose(f);
    return 0;
}

/*
 * li: evaluate Renderer::Li of the scene's own integrator on explicit sample
 * values.  Row layout (floats): imageX, imageY, lensU1, lensU2, then
 *   path_tracing: per bounce b < max_ray_depth: lightComp, lightU0, lightU1,
 *                 bsdfComp, bsdfU0, bsdfU1, pick          (7 * depth floats)
 *   ao:           ao_sample_num pairs (u0, u1)
 * The values are routed to the slots the integrator itself reads
 * (GoblinPathtracer.cpp:78-81 via GoblinLight.cpp:28-33, GoblinMaterial.cpp:31-37).
 */
// samplesPath "@N:seed[:capI:capO]": N rows of mt19937 uniforms generated here (image position scaled to the
// film) instead of a file -- for exporting large ray batches from the reference's own integrator; with caps the
// run stops recording (and each thread stops sampling) once that many intersect / occluded rays are held.
static int cmdLi(SceneView& v, const char* samplesPath, const char* outPath, bool record, bool recordThreads = false) {
    Renderer* r = v.ctx->mRenderer.get();
    PathTracer* pt = dynamic_cast<PathTracer*>(r);
    AORenderer* ao = dynamic_cast<AORenderer*>(r);
    if (!pt && !ao) { fprintf(stderr, "li: unsupported integrator\n"); return 2; }
    SampleQuota quota;
    r->querySampleQuota(v.ctx->mScene, &quota);
    // the AO renderer shoots mAOSampleIndex.sampleNum rays: ao_sample_num rounded up to a square
    const int aoRays = ao ? (int)ao->mAOSampleIndex.sampleNum : 0;
    size_t row = 4 + (pt ? 7 * (size_t)pt->mMaxRayDepth : 2 * (size_t)aoRays);
    std::vector<float> sm;
    size_t n = 0;
    const bool generated = samplesPath[0] == '@';
    unsigned long long genSeed = 0, capI = 0, capO = 0;
    if (generated) {
        unsigned long long nn = 0;
        sscanf(samplesPath + 1, "%llu:%llu:%llu:%llu", &nn, &genSeed, &capI, &capO);
        n = (size_t)nn;
    } else {
        sm = read_f32_file(samplesPath);
        n = sm.size() / row;
    }
    std::vector<float> out(generated ? 0 : n * 3);
    std::vector<uint32_t> calls(generated ? 0 : n * 2);
    // --record keeps the rays in sample order (one thread); --record-mt records on all cores (any order)
    unsigned nt = record && !recordThreads ? 1u : std::max(1u, std::thread::hardware_concurrency());
    g_record = record;
    g_recordCap[0] = capI ? (capI + nt - 1) / nt : ~0ull;
    g_recordCap[1] = capO ? (capO + nt - 1) / nt : ~0ull;
    const float filmW = (float)v.scene->getCamera()->getFilm()->mXRes, filmH = (float)v.scene->getCamera()->getFilm()->mYRes;
    std::vector<std::vector<RecordedRay>> recs(nt);
    std::vector<std::thread> pool;
    const CameraPtr cam = v.scene->getCamera();
    for (unsigned k = 0; k < nt; ++k) {
        pool.emplace_back([&, k]() {
            tl_rec = &recs[k];
            tl_recorded[0] = tl_recorded[1] = 0;
            RNG rng;
            Sample s;
            s.allocateQuota(quota);
            std::mt19937 gen((unsigned)(genSeed * 7919ull + k));
            std::vector<float> rowBuf(row);
            for (size_t i = k; i < n; i += nt) {
                if (generated) {
                    if (capI && capO && tl_recorded[0] >= g_recordCap[0] && tl_recorded[1] >= g_recordCap[1]) break;
                    for (size_t c = 0; c < row; ++c) rowBuf[c] = (float)(gen() >> 8) * (1.0f / 16777216.0f);
                    rowBuf[0] *= filmW;
                    rowBuf[1] *= filmH;
                }
                const float* u = generated ? rowBuf.data() : &sm[row * i];
                s.imageX = u[0]; s.imageY = u[1]; s.lensU1 = u[2]; s.lensU2 = u[3];
                if (pt) {
                    for (int b = 0; b < pt->mMaxRayDepth; ++b) {
                        const float* ub = u + 4 + 7 * b;
                        const LightSampleIndex& li = pt->mLightSampleIndexes[b];
                        const BSDFSampleIndex& bi = pt->mBSDFSampleIndexes[b];
                        s.u1D[li.componentIndex][0] = ub[0];
                        s.u2D[li.geometryIndex][0] = ub[1];
                        s.u2D[li.geometryIndex][1] = ub[2];
                        s.u1D[bi.componentIndex][0] = ub[3];
                        s.u2D[bi.directionIndex][0] = ub[4];
                        s.u2D[bi.directionIndex][1] = ub[5];
                        s.u1D[pt->mPickLightSampleIndexes[b].offset][0] = ub[6];
                    }
                } else {
                    for (int a = 0; a < aoRays; ++a) {
                        s.u2D[ao->mAOSampleIndex.offset][2 * a] = u[4 + 2 * a];
                        s.u2D[ao->mAOSampleIndex.offset][2 * a + 1] = u[4 + 2 * a + 1];
                    }
                }
                RayDifferential ray;
                float w = cam->generateRay(s, &ray);
                uint64_t i0 = tl_count.i, o0 = tl_count.o;
                Color L = r->Li(v.ctx->mScene, ray, s, rng, nullptr);
                if (generated) continue;
                calls[2 * i] = (uint32_t)(tl_count.i - i0);
                calls[2 * i + 1] = (uint32_t)(tl_count.o - o0);
                out[3 * i] = w * L.r; out[3 * i + 1] = w * L.g; out[3 * i + 2] = w * L.b;
            }
            tl_count.flush();
        });
    }
    for (auto& th : pool) th.join();
    FILE* f = fopen(outPath, "wb");
    if (!f) return 2;
    if (!generated) {
        put_f32(f, "L", out, {n, 3});
        put_u32(f, "calls", calls, {n, 2});
    }
    g_record = false;
    g_recordCap[0] = g_recordCap[1] = ~0ull;
    if (record) {
        std::vector<float> rr;
        std::vector<uint32_t> kind;
        for (auto& vec : recs) for (auto& x : vec) {
            rr.insert(rr.end(), x.v, x.v + 8);
            kind.push_back(x.kind);
        }
        put_f32(f, "rays", rr, {kind.size(), 8});
        put_u32(f, "kind", kind);
    }
    fclose(f);
    return 0;
}

static int cmdRender(SceneView& v, const char* outPath, int seed, int threads, int spp) {
    Renderer* r = v.ctx->mRenderer.get();
    if (threads > 0) r->mThreadNum = threads;
    if (spp > 0) r->mSamplePerPixel = spp;
    srand((unsigned)seed); // RNG seeds come from rand() (GoblinUtils.cpp:20, GoblinRenderer.cpp:19)
    Film* film = v.scene->getCamera()->getFilm();
    SampleRange sr;
    film->getSampleRange(sr);
    int root = 1;
    int sppSq = roundToSquare(r->mSamplePerPixel, &root);
    uint64_t samples = (uint64_t)(sr.xEnd - sr.xStart) * (uint64_t)(sr.yEnd - sr.yStart) * sppSq;
    unsigned cores = std::min<unsigned>(std::thread::hardware_concurrency(),
        r->mThreadNum == 0 ? ~0u : (unsigned)r->mThreadNum);
    // a session renders the same context repeatedly: Film::mergeTile accumulates, so start from an empty film
    for (size_t i = 0; i < (size_t)film->mXRes * film->mYRes; ++i) {
        film->mPixels[i].color = Color(0.0f, 0.0f, 0.0f);
        film->mPixels[i].weight = 0.0f;
    }
    g_intersectCalls = 0; g_occludedCalls = 0;
    auto t0 = std::chrono::steady_clock::now();
    v.ctx->render();
    auto t1 = std::chrono::steady_clock::now();
    tl_count.flush();
    double sec = std::chrono::duration<double>(t1 - t0).count();
    if (outPath && strcmp(outPath, "-") != 0) {
        FILE* f = fopen(outPath, "wb");
        if (!f) return 2;
        size_t npx = (size_t)film->mXRes * film->mYRes;
        std::vector<float> rgbw(npx * 4);
        for (size_t i = 0; i < npx; ++i) {
            rgbw[4 * i] = film->mPixels[i].color.r;
            rgbw[4 * i + 1] = film->mPixels[i].color.g;
            rgbw[4 * i + 2] = film->mPixels[i].color.b;
            rgbw[4 * i + 3] = film->mPixels[i].weight;
        }
        put_f32(f, "film", rgbw, {(uint64_t)film->mYRes, (uint64_t)film->mXRes, 4});
        fclose(f);
    }
    uint64_t ic = g_intersectCalls.load(), oc = g_occludedCalls.load();
    printf("\nREF_RESULT {\"seconds\": %.6f, \"camera_samples\": %llu, \"spp\": %d, "
        "\"intersect_calls\": %llu, \"occluded_calls\": %llu, \"cores\": %u, "
        "\"msamples_per_s\": %.6f, \"mrays_per_s\": %.6f}\n",
        sec, (unsigned long long)samples, sppSq, (unsigned long long)ic,
        (unsigned long long)oc, cores, samples / sec * 1e-6, (ic + oc) / sec * 1e-6);
    return 0;
}

static void usage() {
    fprintf(stderr,
        "usage: ref_tool dump    scene.json out.gbar\n"
        "       ref_tool camrays scene.json samples.f32 out.gbar\n"
        "       ref_tool trace   scene.json rays.f32 out.gbar [--no-frag]\n"
        "       ref_tool li      scene.json samples.f32 out.gbar [--record|--record-mt]\n"
        "       ref_tool session scene.json   (commands on stdin, one per line, without the scene path)\n"
        "       ref_tool render  scene.json film.gbar|- [--seed S] [--threads T] [--spp N]\n"
        "       ref_tool post    in.f32 out.f32|out.ppm W H bloomRadius bloomWeight tone\n");
}

// post in.f32 out.f32|out.ppm W H bloomRadius bloomWeight tone: the reference's image
// post-processing (src/GoblinImageIO.cpp:169-237) on W*H rgb floats; a .ppm output goes
// through Goblin::writeImage (tone mapping there), anything else is the raw float result
static int cmdPost(int argc, char** argv) {
    if (argc < 9) return 1;
    int w = atoi(argv[4]), h = atoi(argv[5]);
    float radius = (float)atof(argv[6]), weight = (float)atof(argv[7]);
    bool tone = atoi(argv[8]) != 0;
    std::vector<float> in((size_t)w * h * 3);
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(in.data(), 4, in.size(), f) != in.size()) return 2;
    fclose(f);
    std::vector<Color> c((size_t)w * h);
    for (size_t i = 0; i < c.size(); ++i) c[i] = Color(in[3 * i], in[3 * i + 1], in[3 * i + 2]);
    if (radius > 0.0f && weight > 0.0f) bloom(c.data(), w, h, radius, weight); // Film::writeImage
    std::string out = argv[3];
    if (out.size() > 4 && out.substr(out.size() - 4) == ".ppm") {
        return writeImage(out, c.data(), w, h, tone) ? 0 : 3;
    }
    if (tone) toneMapping(c.data(), w, h);
    for (size_t i = 0; i < c.size(); ++i) { in[3 * i] = c[i].r; in[3 * i + 1] = c[i].g; in[3 * i + 2] = c[i].b; }
    f = fopen(out.c_str(), "wb");
    if (!f) return 3;
    fwrite(in.data(), 4, in.size(), f);
    fclose(f);
    return 0;
}

static std::streambuf* g_cout = nullptr;
static std::ostringstream g_sink;

// one command against a loaded scene; a = {cmd, args...}
static int runCommand(SceneView& view, const std::vector<std::string>& a) {
    const std::string& cmd = a[0];
    if (cmd == "dump" && a.size() >= 2) return cmdDump(view, a[1].c_str());
    if (cmd == "camrays" && a.size() >= 3) return cmdCamRays(view, a[1].c_str(), a[2].c_str());
    if (cmd == "trace" && a.size() >= 3) return cmdTrace(view, a[1].c_str(), a[2].c_str(), !(a.size() >= 4 && a[3] == "--no-frag"));
    if (cmd == "li" && a.size() >= 3) {
        bool rec = a.size() >= 4 && (a[3] == "--record" || a[3] == "--record-mt");
        return cmdLi(view, a[1].c_str(), a[2].c_str(), rec, rec && a[3] == "--record-mt");
    }
    if (cmd == "render" && a.size() >= 2) {
        int seed = 1, threads = 0, spp = 0;
        for (size_t i = 2; i + 1 < a.size(); i += 2) {
            if (a[i] == "--seed") seed = atoi(a[i + 1].c_str());
            else if (a[i] == "--threads") threads = atoi(a[i + 1].c_str());
            else if (a[i] == "--spp") spp = atoi(a[i + 1].c_str());
        }
        std::cout.rdbuf(g_sink.rdbuf()); // progress + "write image" chatter
        int rc = cmdRender(view, a[1].c_str(), seed, threads, spp);
        std::cout.rdbuf(g_cout);
        return rc;
    }
    return -1;
}

int main(int argc, char** argv) {
    if (argc < 3) { usage(); return 1; }
    std::string cmd = argv[1];
    if (cmd == "post") return cmdPost(argc, argv);
    if (cmd != "session" && argc < 4) { usage(); return 1; }
    // the loader echoes every parameter to stdout; silence it while loading
    g_cout = std::cout.rdbuf();
    std::cout.rdbuf(g_sink.rdbuf());
    RenderContext* ctx = ContextLoader::load(argv[2]);
    std::cout.rdbuf(g_cout);
    if (!ctx) { fprintf(stderr, "failed to load %s\n", argv[2]); return 2; }
    SceneView view;
    buildView(view, ctx);
    if (cmd == "session") {
        // session scene.json: the scene is loaded once (a 10 M-triangle scene takes the reference a minute),
        // then one command per line of stdin, same words as the one-shot forms without the scene path
        printf("SESSION_READY\n");
        fflush(stdout);
        std::string line;
        while (std::getline(std::cin, line)) {
            std::istringstream is(line);
            std::vector<std::string> a;
            for (std::string w; is >> w;) a.push_back(w);
            if (a.empty()) continue;
            if (a[0] == "quit") break;
            int rc = runCommand(view, a);
            printf("SESSION_DONE %d %s\n", rc, a[0].c_str());
            fflush(stdout);
            if (rc != 0) return rc < 0 ? 1 : rc;
        }
        return 0;
    }
    std::vector<std::string> a = {cmd};
    for (int i = 3; i < argc; ++i) a.push_back(argv[i]);
    int rc = runCommand(view, a);
    if (rc < 0) { usage(); return 1; }
    return rc;
}
