/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product.
 *
 * goblin_oracle: a plain C++ CPU restatement of the reference's path-tracing
 * hot path (bachi95/Goblin), written from the reference sources, function by
 * function, with the file:line each one follows.  It exists to check the CUDA
 * path; nothing under goblin_b200/ links, loads or calls it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it.
 *
 * Parity is PINNED: tests/test_oracle_port.py checks this file against golden
 * vectors produced by the unmodified reference itself (oracle/_ref/ref_tool,
 * tests/golden/make_golden.py): hit ids / t / epsilon bit-exact, per-sample
 * Li, the number of Scene::intersect / Scene::occluded calls per sample, and
 * camera rays.
 *
 * It consumes the flattened scene of include/goblin_b200.h (gb_scene_desc,
 * itself checked bit-exactly against the reference's dump) but none of the
 * product's derived device layouts: triangles are fetched through tri_index /
 * vert_pos and edges recomputed per test as the reference does.
 *
 * Known, documented divergence: a triangle whose uv
 * determinant is 0 makes the reference read the *stale* contents of the
 * caller's Fragment (src/GoblinTriangle.cpp:113-117); product and oracle both
 * use coordinateAxises(normal) for dpdu there.
 *
 * Build: make -C oracle port  ->  oracle/_build/libgoblin_oracle.so
 * Compiled for generic x86-64 at -O2: SSE2 scalar float, no FMA contraction,
 * like the reference build of oracle/Makefile.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "goblin_b200.h"

namespace {

const float PI = 3.14159265358979323f;          // src/GoblinUtils.h:43-46
const float TWO_PI = 6.28318530718f;
const float INV_PI = 0.31830988618379067154f;
const float INV_TWOPI = 0.15915494309189533577f;
const float INF = std::numeric_limits<float>::infinity();

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
// Vector3::operator/(float) multiplies by the reciprocal (src/GoblinVector.h:166-169)
inline V3 operator/(V3 a, float s) { float inv = 1.0f / s; return V3(a.x * inv, a.y * inv, a.z * inv); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float absdot(V3 a, V3 b) { return std::fabs(dot(a, b)); }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float sqLen(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float len(V3 a) { return std::sqrt(sqLen(a)); }
inline V3 normalize(V3 a) { return a / len(a); }
inline bool isBlack(V3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; }

struct Ray {
    V3 o, d;
    float mint, maxt;
};

// Transform::onPoint / onVector / invertPoint / invertVector / onNormal on the
// 3x4 rows of gb_instance / gb_light (src/GoblinTransform.cpp:97-164)
inline V3 xfPoint(const float* m, V3 p) {
    return V3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
        m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
inline V3 xfVector(const float* m, V3 v) {
    return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
        m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
inline V3 xfNormal(const float* inv, V3 n) { // transpose(inverse) * n
    return V3(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
        inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
}

// what Triangle / Sphere / Disk::intersect leave in the Fragment, as far as the
// path tracer reads it (position, normal, dpdu)
struct Frag {
    V3 p, n, dpdu;
    // what only the procedural textures read (src/GoblinGeometry.h Fragment)
    float u = 0.0f, v = 0.0f;
    V3 dpdv, dpdx, dpdy;
    float dudx = 0.0f, dvdx = 0.0f, dudy = 0.0f, dvdy = 0.0f;
};

// RayDifferential's auxiliary rays (src/GoblinRay.h), set by the camera only
struct RayDiff {
    bool has = false;
    V3 dxO, dxD, dyO, dyD;
};

struct Isect {
    int inst = -1;  // instance index (scene order)
    int prim = 0;   // face index within the mesh, 0 for sphere / disk
    Frag frag;
};

struct Stats {
    uint64_t nodes = 0, prims = 0, insts = 0, closest = 0, any = 0;
    uint64_t nodesAny = 0, primsAny = 0, instsAny = 0; // the any-hit walks' share
    uint64_t refIntersect = 0, refOccluded = 0; // calls the reference would have made
};

// ---- sampling maps: src/GoblinSampler.cpp:449-602, src/GoblinUtils.cpp:58-69
V3 uniformSampleCone(float u1, float u2, float cosThetaMax, V3 x, V3 y, V3 z) {
    float cosTheta = 1.0f - u1 + u1 * cosThetaMax;
    float sinTheta = sqrtf(std::max(0.0f, 1.0f - cosTheta * cosTheta));
    float phi = TWO_PI * u2;
    return x * sinTheta * (float)cos(phi) + y * sinTheta * (float)sin(phi) + z * cosTheta;
}
float uniformConePdf(float cosThetaMax) { return 1.0f / (TWO_PI * (1.0f - cosThetaMax)); }
V3 uniformSampleSphere(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float sinTheta = sqrtf(std::max(0.0f, 1.0f - z * z));
    float phi = TWO_PI * u2;
    return V3(sinTheta * (float)cos(phi), sinTheta * (float)sin(phi), z);
}
V3 uniformSampleHemisphere(float u1, float u2) {
    float sinTheta = sqrtf(std::max(0.0f, 1.0f - u1 * u1));
    float phi = TWO_PI * u2;
    return V3(sinTheta * (float)cos(phi), sinTheta * (float)sin(phi), u1);
}
V3 cosineSampleHemisphere(float u1, float u2) {
    float sinTheta = sqrtf(u1);
    float cosTheta = sqrtf(std::max(0.0f, 1.0f - u1));
    float phi = TWO_PI * u2;
    return V3(sinTheta * (float)cos(phi), sinTheta * (float)sin(phi), cosTheta);
}
void uniformSampleDisk(float u1, float u2, float* ox, float* oy) {
    float r, theta;
    float x = 2.0f * u1 - 1.0f;
    float y = 2.0f * u2 - 1.0f;
    if (x + y > 0) {
        if (x > y) { r = x; theta = 0.25f * PI * (y / x); }
        else { r = y; theta = 0.25f * PI * (2.0f - x / y); }
    } else {
        if (x < y) { r = -x; theta = 0.25f * PI * (4.0f + y / x); }
        else {
            r = -y;
            if (y != 0.0f) theta = 0.25f * PI * (6.0f - x / y);
            else theta = 0.0f;
        }
    }
    *ox = r * (float)cos(theta);
    *oy = r * (float)sin(theta);
}
void coordinateAxises(V3 a1, V3* a2, V3* a3) {
    if (fabsf(a1.x) > fabsf(a1.y)) {
        float invLen = 1.0f / sqrtf(a1.x * a1.x + a1.z * a1.z);
        *a2 = V3(-a1.z * invLen, 0.0f, a1.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a1.y * a1.y + a1.z * a1.z);
        *a2 = V3(0.0f, -a1.z * invLen, a1.y * invLen);
    }
    *a3 = cross(a1, *a2);
}
float powerHeuristic(float fPdf, float gPdf) { // nF = nG = 1, src/GoblinSampler.h:286-290
    float f = fPdf, g = gPdf;
    return (f * f) / (f * f + g * g);
}

// ---- primitive tests
// quadratic, src/GoblinUtils.cpp:93-113
bool quadratic(float A, float B, float C, float* t1, float* t2) {
    float discriminant = B * B - 4.0f * A * C;
    if (discriminant < 0.0f) return false;
    float rootDiscrim = std::sqrt(discriminant);
    float q;
    if (B < 0) q = -0.5f * (B - rootDiscrim);
    else q = -0.5f * (B + rootDiscrim);
    *t1 = q / A;
    *t2 = C / q;
    if (*t1 > *t2) std::swap(*t1, *t2);
    return true;
}

// Sphere::intersect, src/GoblinSphere.cpp:12-79 (object space, origin centred)
bool sphereIntersect(float radius, Ray& ray, float* epsilon, Frag* frag) {
    float A = sqLen(ray.d);
    float B = 2.0f * dot(ray.d, ray.o);
    float C = sqLen(ray.o) - radius * radius;
    float tNear, tFar;
    if (!quadratic(A, B, C, &tNear, &tFar)) return false;
    if (tNear > ray.maxt || tFar < ray.mint) return false;
    float tHit = tNear;
    if (tHit < ray.mint) {
        tHit = tFar;
        if (tHit > ray.maxt) return false;
    }
    ray.maxt = tHit;
    *epsilon = 1e-3f * tHit;
    V3 pHit = ray.o + tHit * ray.d;
    frag->p = pHit;
    frag->n = normalize(pHit);
    frag->dpdu = V3(-TWO_PI * pHit.y, TWO_PI * pHit.x, 0.0f);
    // uv and dpdv, src/GoblinSphere.cpp:62-77
    float phi = (float)atan2(pHit.y, pHit.x);
    if (phi < 0.0f) phi += TWO_PI;
    frag->u = phi * INV_TWOPI;
    float theta = (float)acos(pHit.z / radius);
    frag->v = theta * INV_PI;
    float invR = 1.0f / (float)sqrt(pHit.x * pHit.x + pHit.y * pHit.y);
    float cosPhi = pHit.x * invR;
    float sinPhi = pHit.y * invR;
    frag->dpdv = PI * V3(pHit.z * cosPhi, pHit.z * sinPhi, -radius * (float)sin(theta));
    return true;
}
// Sphere::occluded, src/GoblinSphere.cpp:81-98: the same acceptance test
bool sphereOccluded(float radius, const Ray& ray) {
    float A = sqLen(ray.d);
    float B = 2.0f * dot(ray.d, ray.o);
    float C = sqLen(ray.o) - radius * radius;
    float tNear, tFar;
    if (!quadratic(A, B, C, &tNear, &tFar)) return false;
    if (tNear > ray.maxt || tFar < ray.mint) return false;
    float tHit = tNear;
    if (tHit < ray.mint) {
        tHit = tFar;
        if (tHit > ray.maxt) return false;
    }
    return true;
}

// Disk::intersect / occluded, src/GoblinDisk.cpp:12-64
bool diskIntersect(float radius, Ray& ray, float* epsilon, Frag* frag) {
    if (std::fabs(ray.d.z) < 1e-7f) return false;
    float t = -ray.o.z / ray.d.z;
    V3 p = ray.o + t * ray.d;
    if (t < ray.mint || t > ray.maxt) return false;
    float squareR = p.x * p.x + p.y * p.y;
    if (squareR > radius * radius) return false;
    ray.maxt = t;
    *epsilon = 1e-3f * t;
    frag->p = p;
    frag->n = V3(0.0f, 0.0f, 1.0f);
    frag->dpdu = V3(-TWO_PI * p.y, TWO_PI * p.x, 0.0f);
    // uv and dpdv, src/GoblinDisk.cpp:50-61
    float r = (float)sqrt(squareR);
    float phi = (float)atan2(p.y, p.x);
    if (phi < 0.0f) phi += TWO_PI;
    frag->u = phi * INV_TWOPI;
    frag->v = r / radius;
    frag->dpdv = V3(radius * p.x / r, radius * p.y / r, 0.0f);
    return true;
}
bool diskOccluded(float radius, const Ray& ray) {
    if (std::fabs(ray.d.z) < 1e-7f) return false;
    float t = -ray.o.z / ray.d.z;
    V3 p = ray.o + t * ray.d;
    if (t < ray.mint || t > ray.maxt) return false;
    return p.x * p.x + p.y * p.y <= radius * radius;
}

struct Oracle {
    const gb_scene_desc* d;
    explicit Oracle(const gb_scene_desc* desc) : d(desc) {}

    V3 vpos(const gb_model& m, uint32_t v) const {
        const float* p = d->vert_pos + 3 * ((size_t)m.vert_offset + v);
        return V3(p[0], p[1], p[2]);
    }
    V3 vnrm(const gb_model& m, uint32_t v) const {
        const float* p = d->vert_nrm + 3 * ((size_t)m.vert_offset + v);
        return V3(p[0], p[1], p[2]);
    }

    // Triangle::intersect, src/GoblinTriangle.cpp:38-125
    bool triangleIntersect(const gb_model& m, uint32_t face, Ray& ray, float* epsilon, Frag* frag) const {
        const uint32_t* ti = d->tri_index + 3 * ((size_t)m.tri_offset + face);
        V3 p0 = vpos(m, ti[0]), p1 = vpos(m, ti[1]), p2 = vpos(m, ti[2]);
        V3 e1 = p1 - p0;
        V3 e2 = p2 - p0;
        V3 s1 = cross(ray.d, e2);
        float divisor = dot(s1, e1);
        if (divisor == 0.0f) return false;
        float invDivisor = 1.0f / divisor;
        float fEpsilon = 1e-7f;
        V3 s = ray.o - p0;
        float b1 = dot(s, s1) * invDivisor;
        if (b1 + fEpsilon < 0.0f || b1 - fEpsilon > 1.0f) return false;
        V3 s2 = cross(s, e1);
        float b2 = dot(ray.d, s2) * invDivisor;
        if (b2 + fEpsilon < 0.0f || b1 + b2 - fEpsilon > 1.0f) return false;
        float t = dot(e2, s2) * invDivisor;
        if (t < ray.mint || t > ray.maxt) return false;
        float b0 = 1.0f - b1 - b2;
        ray.maxt = t;
        *epsilon = 1e-3f * t;
        V3 position = ray.o + t * ray.d;
        V3 normal;
        if (m.has_normal) normal = normalize(b0 * vnrm(m, ti[0]) + b1 * vnrm(m, ti[1]) + b2 * vnrm(m, ti[2]));
        else normal = normalize(cross(e1, e2));
        float uv[3][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {0.0f, 1.0f}};
        if (m.has_uv) {
            for (int k = 0; k < 3; ++k) {
                const float* t2 = d->vert_uv + 2 * ((size_t)m.vert_offset + ti[k]);
                uv[k][0] = t2[0];
                uv[k][1] = t2[1];
            }
        }
        float du1 = uv[1][0] - uv[0][0];
        float dv1 = uv[1][1] - uv[0][1];
        float du2 = uv[2][0] - uv[0][0];
        float dv2 = uv[2][1] - uv[0][1];
        float determinant = du1 * dv2 - dv1 * du2;
        V3 dpdu, dpdv;
        if (determinant == 0.0f) {
            // divergence D1: the reference reads the stale output fragment here
            coordinateAxises(normal, &dpdu, &dpdv);
        } else {
            float invDet = 1.0f / determinant;
            dpdu = invDet * (dv2 * e1 - dv1 * e2);
            dpdv = invDet * (-du2 * e1 + du1 * e2);
        }
        frag->p = position;
        frag->n = normal;
        frag->dpdu = dpdu;
        frag->dpdv = dpdv;
        // Vector2 uv(b0 * uvs[0] + b1 * uvs[1] + b2 * uvs[2]), src/GoblinTriangle.cpp:105
        frag->u = b0 * uv[0][0] + b1 * uv[1][0] + b2 * uv[2][0];
        frag->v = b0 * uv[0][1] + b1 * uv[1][1] + b2 * uv[2][1];
        return true;
    }
    // Triangle::occluded, src/GoblinTriangle.cpp:127-163
    bool triangleOccluded(const gb_model& m, uint32_t face, const Ray& ray) const {
        const uint32_t* ti = d->tri_index + 3 * ((size_t)m.tri_offset + face);
        V3 p0 = vpos(m, ti[0]), p1 = vpos(m, ti[1]), p2 = vpos(m, ti[2]);
        V3 e1 = p1 - p0;
        V3 e2 = p2 - p0;
        V3 s1 = cross(ray.d, e2);
        float divisor = dot(s1, e1);
        if (divisor == 0.0f) return false;
        float invDivisor = 1.0f / divisor;
        float fEpsilon = 1e-7f;
        V3 s = ray.o - p0;
        float b1 = dot(s, s1) * invDivisor;
        if (b1 + fEpsilon < 0.0f || b1 - fEpsilon > 1.0f) return false;
        V3 s2 = cross(s, e1);
        float b2 = dot(ray.d, s2) * invDivisor;
        if (b2 + fEpsilon < 0.0f || b1 + b2 - fEpsilon > 1.0f) return false;
        float t = dot(e2, s2) * invDivisor;
        if (t < ray.mint || t > ray.maxt) return false;
        return true;
    }

    // the ordered slab test, src/GoblinBVH.cpp:156-187
    static bool slab(const gb_bvh_node& n, const Ray& ray, V3 invDir, const uint32_t neg[3]) {
        const float* b[2] = {n.bmin, n.bmax};
        float tMin = (b[neg[0]][0] - ray.o.x) * invDir.x;
        float tMax = (b[1 - neg[0]][0] - ray.o.x) * invDir.x;
        float tYMin = (b[neg[1]][1] - ray.o.y) * invDir.y;
        float tYMax = (b[1 - neg[1]][1] - ray.o.y) * invDir.y;
        if (tYMax < tMin || tYMin > tMax) return false;
        if (tYMin > tMin) tMin = tYMin;
        if (tYMax < tMax) tMax = tYMax;
        float tZMin = (b[neg[2]][2] - ray.o.z) * invDir.z;
        float tZMax = (b[1 - neg[2]][2] - ray.o.z) * invDir.z;
        if (tZMax < tMin || tZMin > tMax) return false;
        if (tZMin > tMin) tMin = tZMin;
        if (tZMax < tMax) tMax = tZMax;
        return (tMin < ray.maxt) && (tMax > ray.mint);
    }

    // BVH::intersect / BVH::occluded (src/GoblinBVH.cpp:234-280 / 189-232) with the
    // leaf action supplied by the caller.  leaf(index) returns true on a hit.
    template <bool ANY, typename Leaf>
    bool walk(const gb_bvh_node* nodes, uint32_t nNodes, const Ray& ray, Stats& st, Leaf leaf) const {
        if (nNodes == 0) return false;
        V3 invDir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        uint32_t neg[3] = {ray.d.x < 0.0f, ray.d.y < 0.0f, ray.d.z < 0.0f};
        uint32_t nodeNum = 0, todoOffset = 0;
        uint32_t todo[64];
        bool hit = false;
        while (true) {
            const gb_bvh_node& node = nodes[nodeNum];
            st.nodes++;
            if (slab(node, ray, invDir, neg)) {
                if (node.nprims > 0) {
                    for (uint32_t i = 0; i < node.nprims; ++i) {
                        if (leaf(node.offset + i)) {
                            if (ANY) return true;
                            hit = true;
                        }
                    }
                    if (todoOffset == 0) break;
                    nodeNum = todo[--todoOffset];
                } else {
                    if (neg[node.axis]) {
                        todo[todoOffset++] = nodeNum + 1;
                        nodeNum = node.offset;
                    } else {
                        todo[todoOffset++] = node.offset;
                        nodeNum = nodeNum + 1;
                    }
                }
            } else {
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            }
        }
        return hit;
    }

    // Scene::intersect -> BVH -> InstancedPrimitive::intersect -> Model::intersect
    // (src/GoblinScene.cpp:75-83, src/GoblinPrimitive.cpp:103-112, src/GoblinModel.cpp:39-52)
    // Material::perturb -> BumpShaders::evaluate, src/GoblinMaterial.cpp:221-281 (Scene::intersect applies it
    // to every hit it returns, src/GoblinScene.cpp:78-81)
    void perturb(const gb_material& m, Frag& f) const {
        if (m.bump_tex) {
            V3 p = f.p, n = f.n;
            float bumpD = texLookup(m.bump_tex - 1, f).x;
            float du = 0.002f;
            Frag fdu = f;
            fdu.p = p + du * f.dpdu;
            fdu.u = f.u + du; // uv + Vector2(du, 0)
            fdu.v = f.v + 0.0f;
            float bumpDdu = texLookup(m.bump_tex - 1, fdu).x;
            V3 bumpDPDU = f.dpdu + (bumpDdu - bumpD) / du * n;
            float dv = 0.002f;
            Frag fdv = f;
            fdv.p = p + dv * f.dpdv;
            fdv.u = f.u + 0.0f;
            fdv.v = f.v + dv;
            float bumpDdv = texLookup(m.bump_tex - 1, fdv).x;
            V3 bumpDPDV = f.dpdv + (bumpDdv - bumpD) / dv * n;
            V3 bumpN = normalize(cross(bumpDPDU, bumpDPDV));
            if (dot(bumpN, n) < 0.0f) bumpN = bumpN * -1.0f;
            f.n = bumpN;
            f.dpdu = bumpDPDU;
            f.dpdv = bumpDPDV;
        }
        if (m.normal_tex) {
            V3 c = texLookup(m.normal_tex - 1, f);
            V3 nShade = 2.0f * c - V3(1.0f, 1.0f, 1.0f);
            V3 nWorld = normalize(shadeToWorld(f, nShade));
            if (dot(nWorld, f.n) < 0.0f) nWorld = nWorld * -1.0f;
            f.n = nWorld;
        }
    }
    // filter: the IntersectFilter of src/GoblinPathtracer.cpp:5-11 (0 = none, 1 = isOpaque, 2 = notOpaque).
    // The reference applies it per refined Model, after walking a mesh's BVH (src/GoblinModel.cpp:44);
    // the material belongs to the model, so skipping the whole instance gives the same hits.
    bool filteredOut(const gb_model& m, int filter) const {
        if (!filter) return false;
        const bool opaque = !d->materials[m.material].mask;
        return (filter == 1) != opaque;
    }
    bool intersect(Ray& ray, float* epsilon, Isect* isect, Stats& st, int filter = 0) const {
        if (!intersectGeometry(ray, epsilon, isect, st, filter)) return false;
        // *fragment = Fragment(position, normal, uv, dpdu, dpdv): a fresh Fragment has no differentials
        isect->frag.dpdx = isect->frag.dpdy = V3();
        isect->frag.dudx = isect->frag.dvdx = isect->frag.dudy = isect->frag.dvdy = 0.0f;
        const gb_material& m = d->materials[d->models[d->instances[isect->inst].model].material];
        if (m.bump_tex | m.normal_tex) perturb(m, isect->frag);
        return true;
    }
    bool intersectGeometry(Ray& ray, float* epsilon, Isect* isect, Stats& st, int filter) const {
        st.closest++;
        return walk<false>(d->top_nodes, d->n_top_nodes, ray, st, [&](uint32_t slot) {
            uint32_t ii = d->top_order[slot];
            const gb_instance& in = d->instances[ii];
            const gb_model& m = d->models[in.model];
            if (filteredOut(m, filter)) return false;
            st.insts++;
            Ray r{xfPoint(in.to_object, ray.o), xfVector(in.to_object, ray.d), ray.mint, ray.maxt}; // invertRay
            bool hit = false;
            if (m.kind == GB_GEOM_MESH) {
                hit = walk<false>(d->model_nodes + m.node_offset, m.node_count, r, st, [&](uint32_t ts) {
                    uint32_t face = d->model_order[m.tri_offset + ts];
                    st.prims++;
                    if (triangleIntersect(m, face, r, epsilon, &isect->frag)) {
                        isect->prim = (int)face;
                        return true;
                    }
                    return false;
                });
            } else {
                st.prims++;
                hit = m.kind == GB_GEOM_SPHERE ? sphereIntersect(m.radius, r, epsilon, &isect->frag)
                                               : diskIntersect(m.radius, r, epsilon, &isect->frag);
                if (hit) isect->prim = 0;
            }
            if (hit) {
                isect->inst = (int)ii;
                // Fragment::transform, src/GoblinGeometry.cpp:31-37
                isect->frag.p = xfPoint(in.to_world, isect->frag.p);
                isect->frag.n = normalize(xfNormal(in.to_object, isect->frag.n));
                isect->frag.dpdu = xfVector(in.to_world, isect->frag.dpdu);
                isect->frag.dpdv = xfVector(in.to_world, isect->frag.dpdv);
                ray.maxt = r.maxt;
            }
            return hit;
        });
    }

    // Scene::occluded (src/GoblinScene.cpp:85-87)
    bool occluded(const Ray& ray, Stats& st, int filter = 0) const {
        st.any++;
        const uint64_t n0 = st.nodes, p0 = st.prims, i0 = st.insts;
        bool occ = occludedWalk(ray, st, filter);
        st.nodesAny += st.nodes - n0;
        st.primsAny += st.prims - p0;
        st.instsAny += st.insts - i0;
        return occ;
    }
    bool occludedWalk(const Ray& ray, Stats& st, int filter) const {
        return walk<true>(d->top_nodes, d->n_top_nodes, ray, st, [&](uint32_t slot) {
            const gb_instance& in = d->instances[d->top_order[slot]];
            const gb_model& m = d->models[in.model];
            if (filteredOut(m, filter)) return false;
            st.insts++;
            Ray r{xfPoint(in.to_object, ray.o), xfVector(in.to_object, ray.d), ray.mint, ray.maxt};
            if (m.kind == GB_GEOM_MESH) {
                return walk<true>(d->model_nodes + m.node_offset, m.node_count, r, st, [&](uint32_t ts) {
                    st.prims++;
                    return triangleOccluded(m, d->model_order[m.tri_offset + ts], r);
                });
            }
            st.prims++;
            return m.kind == GB_GEOM_SPHERE ? sphereOccluded(m.radius, r) : diskOccluded(m.radius, r);
        });
    }

    // ---- camera: PerspectiveCamera::generateRay, src/GoblinCamera.cpp:97-148
    static V3 quatRotate(const float q[4], V3 p) { // Quaternion::operator*(Vector3), src/GoblinQuaternion.cpp:86-92
        V3 v(q[1], q[2], q[3]);
        V3 uv = cross(v, p);
        V3 uuv = cross(v, uv);
        uv = uv * (2.0f * q[0]);
        uuv = uuv * 2.0f;
        return p + uv + uuv;
    }
    // the dx / dy auxiliary rays of PerspectiveCamera::generateRay, src/GoblinCamera.cpp:103-142
    RayDiff cameraRayDiff(float imageX, float imageY, float lensU1, float lensU2) const {
        const gb_camera& c = d->camera;
        float invXRes = 1.0f / (float)d->film.xres;
        float invYRes = 1.0f / (float)d->film.yres;
        float xNDC = +2.0f * imageX * invXRes - 1.0f;
        float yNDC = -2.0f * imageY * invYRes + 1.0f;
        float dxNDC = +2.0f * (imageX + 1.0f) * invXRes - 1.0f;
        float dyNDC = -2.0f * (imageY + 1.0f) * invYRes + 1.0f;
        float xView = xNDC / c.proj00;
        float yView = yNDC / c.proj11;
        V3 viewDir(xView, yView, 1.0f);
        V3 dxViewDir(dxNDC / c.proj00, yView, 1.0f);
        V3 dyViewDir(xView, dyNDC / c.proj11, 1.0f);
        V3 pos(c.position[0], c.position[1], c.position[2]);
        RayDiff rd;
        rd.has = true;
        if (c.orthographic) { // OrthographicCamera::generateRay, src/GoblinCamera.cpp:301-329
            rd.dxO = pos + quatRotate(c.orientation, V3(0.5f * c.film_width * dxNDC, 0.5f * c.film_height * yNDC, 0.0f));
            rd.dyO = pos + quatRotate(c.orientation, V3(0.5f * c.film_width * xNDC, 0.5f * c.film_height * dyNDC, 0.0f));
            rd.dxD = rd.dyD = quatRotate(c.orientation, V3(0.0f, 0.0f, 1.0f));
            return rd;
        }
        if (c.lens_radius == 0.0f) {
            rd.dxO = rd.dyO = pos;
            rd.dxD = quatRotate(c.orientation, normalize(dxViewDir));
            rd.dyD = quatRotate(c.orientation, normalize(dyViewDir));
        } else {
            float ft = c.focal_distance / viewDir.z;
            V3 pDxFocus = dxViewDir * ft;
            V3 pDyFocus = dyViewDir * ft;
            float lx, ly;
            uniformSampleDisk(lensU1, lensU2, &lx, &ly);
            V3 viewOrigin(c.lens_radius * lx, c.lens_radius * ly, 0.0f);
            rd.dxO = rd.dyO = quatRotate(c.orientation, viewOrigin) + pos;
            rd.dxD = quatRotate(c.orientation, normalize(pDxFocus - viewOrigin));
            rd.dyD = quatRotate(c.orientation, normalize(pDyFocus - viewOrigin));
        }
        return rd;
    }

    // Intersection::computeUVDifferential, src/GoblinPrimitive.cpp:32-97
    static void computeUVDifferential(Frag& fragment, const RayDiff& ray) {
        float dudx, dvdx, dudy, dvdy;
        dudx = dvdx = dudy = dvdy = 0.0f;
        if (ray.has) {
            V3 p = fragment.p;
            V3 n = fragment.n;
            float minusD = dot(p, n);
            float tdx = (minusD - dot(ray.dxO, n)) / dot(ray.dxD, n);
            float tdy = (minusD - dot(ray.dyO, n)) / dot(ray.dyD, n);
            if (!(tdx != tdx) && !(tdy != tdy)) {
                V3 pdx = ray.dxO + tdx * ray.dxD;
                V3 pdy = ray.dyO + tdy * ray.dyD;
                V3 dpdx = pdx - p;
                V3 dpdy = pdy - p;
                fragment.dpdx = dpdx;
                fragment.dpdy = dpdy;
                int axis[2];
                if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) { axis[0] = 1; axis[1] = 2; }
                else if (fabsf(n.y) > fabsf(n.z)) { axis[0] = 0; axis[1] = 2; }
                else { axis[0] = 0; axis[1] = 1; }
                auto comp = [](V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); };
                float A[2][2] = {{comp(fragment.dpdu, axis[0]), comp(fragment.dpdv, axis[0])},
                                 {comp(fragment.dpdu, axis[1]), comp(fragment.dpdv, axis[1])}};
                float Bx[2] = {comp(dpdx, axis[0]), comp(dpdx, axis[1])};
                if (!solve2x2(A, Bx, &dudx, &dvdx)) dudx = dvdx = 0.0f;
                float By[2] = {comp(dpdy, axis[0]), comp(dpdy, axis[1])};
                if (!solve2x2(A, By, &dudy, &dvdy)) dudy = dvdy = 0.0f;
            }
        }
        fragment.dudx = dudx; fragment.dvdx = dvdx; fragment.dudy = dudy; fragment.dvdy = dvdy;
    }
    // solve2x2LinearSystem, src/GoblinUtils.h:151-165
    static bool solve2x2(const float A[2][2], const float B[2], float* x, float* y) {
        float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
        if (fabsf(det) < 1e-10f) return false;
        *x = (+A[1][1] * B[0] - A[0][1] * B[1]) / det;
        *y = (-A[1][0] * B[0] + A[0][0] * B[1]) / det;
        if (*x != *x || *y != *y) return false;
        return true;
    }

    // ---- procedural textures, src/GoblinTexture.cpp:292-427
    struct TexCoord { float s, t, dsdx, dtdx, dsdy, dtdy; };
    static int floorInt(float f) { return (int)floor(f); }
    void pointToST(const gb_texture& t, V3 p, float* s, float* tt) const { // SphericalMapping::pointToST
        V3 v = normalize(xfPoint(t.to_tex, p));
        float theta = (float)acos(clampf(v.z, -1.0f, 1.0f));
        float phi = (float)atan2(v.y, v.x);
        phi = phi < 0.0f ? phi + TWO_PI : phi;
        *s = phi * INV_TWOPI;
        *tt = theta * INV_PI;
    }
    TexCoord mapTexture(const gb_texture& t, const Frag& f) const {
        TexCoord tc;
        if (t.mapping == GB_MAPPING_SPHERICAL) { // SphericalMapping::map
            float s, tt;
            pointToST(t, f.p, &s, &tt);
            tc.s = s; tc.t = tt;
            float sdx, tdx, sdy, tdy;
            pointToST(t, f.p + f.dpdx, &sdx, &tdx);
            pointToST(t, f.p + f.dpdy, &sdy, &tdy);
            float dsdx = sdx - s;
            if (dsdx > 0.5f) dsdx -= 1.0f; else if (dsdx < -0.5f) dsdx += 1.0f;
            float dsdy = sdy - s;
            if (dsdy > 0.5f) dsdy -= 1.0f; else if (dsdy < -0.5f) dsdy += 1.0f;
            tc.dsdx = dsdx; tc.dtdx = tdx - tt; tc.dsdy = dsdy; tc.dtdy = tdy - tt;
        } else { // UVMapping::map
            tc.s = t.map_scale[0] * f.u + t.map_offset[0];
            tc.t = t.map_scale[1] * f.v + t.map_offset[1];
            tc.dsdx = t.map_scale[0] * f.dudx;
            tc.dtdx = t.map_scale[1] * f.dvdx;
            tc.dsdy = t.map_scale[0] * f.dudy;
            tc.dtdy = t.map_scale[1] * f.dvdy;
        }
        return tc;
    }
    static float integrateChecker(float x) {
        float xHalf = 0.5f * x;
        return (float)floor(xHalf) + 2.0f * std::max(xHalf - (float)floor(xHalf) - 0.5f, 0.0f);
    }
    // ---- image textures: MIPMap<T> lookups, src/GoblinTexture.cpp:10-37,82-288
    V3 imgTexel(const gb_texture& t, int level, int s, int tt) const { // ImageBuffer::texel
        const gb_image_level& im = d->image_levels[t.first_level + level];
        if (t.address_mode == GB_ADDRESS_CLAMP) {
            s = std::min(std::max(s, 0), im.width - 1);
            tt = std::min(std::max(s, 0), im.height - 1); // sic: the reference clamps s into t
        } else if (t.address_mode == GB_ADDRESS_BORDER) {
            if (s < 0 || tt < 0 || s >= im.width || tt >= im.height) return V3();
        } else {
            s = s % im.width;
            tt = tt % im.height;
            if (s < 0) s += im.width;
            if (tt < 0) tt += im.height;
        }
        const float* c = d->image_texels + 4 * ((size_t)im.texel_offset + (size_t)tt * im.width + s);
        return V3(c[0], c[1], c[2]);
    }
    V3 imgBilinear(const gb_texture& t, int level, float s, float tt) const { // MIPMap::lookup(level, s, t, m)
        // the reference clamps to [0, mLevelsNum] and then indexes one past the pyramid at the top;
        // product and oracle clamp to the last level
        level = std::min(std::max(level, 0), t.n_levels - 1);
        const gb_image_level& im = d->image_levels[t.first_level + level];
        float sRes = s * im.width - 0.5f;
        float tRes = tt * im.height - 0.5f;
        int s0 = floorInt(sRes);
        float ds = sRes - (float)s0;
        int t0 = floorInt(tRes);
        float dt = tRes - (float)t0;
        return (1.0f - ds) * (1.0f - dt) * imgTexel(t, level, s0, t0) + (ds) * (1.0f - dt) * imgTexel(t, level, s0 + 1, t0) +
            (1.0f - ds) * (dt) * imgTexel(t, level, s0, t0 + 1) + (ds) * (dt) * imgTexel(t, level, s0 + 1, t0 + 1);
    }
    // log2 of a float binds to the float overload in the reference's translation units (sqrt, pow,
    // sin, acos, atan2, exp and floor bind to the double ones)
    V3 imgTrilinear(const gb_texture& t, float s, float tt, float width) const {
        float level = t.n_levels - 1 + log2f(std::max(width, 1e-8f));
        int iLevel = floorInt(level);
        if (iLevel < 0) return imgBilinear(t, 0, s, tt);
        if (iLevel >= t.n_levels - 1) return imgBilinear(t, t.n_levels - 1, s, tt);
        float delta = level - (float)iLevel;
        return (1.0f - delta) * imgBilinear(t, iLevel, s, tt) + (delta) * imgBilinear(t, iLevel + 1, s, tt);
    }
    V3 imgEWA(const gb_texture& t, int level, float s, float tt, float A, float B, float C) const { // MIPMap::EWA
        const gb_image_level& im = d->image_levels[t.first_level + level];
        float sRes = (float)im.width;
        float tRes = (float)im.height;
        s = s * im.width - 0.5f;
        tt = tt * im.height - 0.5f;
        A = A / (sRes * sRes);
        B = B / (sRes * tRes);
        C = C / (tRes * tRes);
        float invDet = 1.0f / (-B * B + 4.0f * A * C);
        float offsetS = 2.0f * (float)sqrt(C * invDet);
        float offsetT = 2.0f * (float)sqrt(A * invDet);
        int s0 = (int)ceil(s - offsetS);
        int s1 = floorInt(s + offsetS);
        int t0 = (int)ceil(tt - offsetT);
        int t1 = floorInt(tt + offsetT);
        float weightSum = 0.0f;
        V3 result;
        for (int is = s0; is <= s1; ++is) {
            for (int it = t0; it <= t1; ++it) {
                float ss = is - s;
                float tt2 = it - tt;
                float r2 = A * ss * ss + B * ss * tt2 + C * tt2 * tt2;
                if (r2 <= 1.0f) {
                    size_t lutIndex = (size_t)floorInt(r2 * 128);
                    size_t li = std::min(lutIndex, (size_t)127);
                    float r2l = float(li) / float(127); // initEWALut: expf(-2 r2) - expf(-2)
                    float weight = expf(-2.0f * r2l) - expf(-2.0f);
                    result = result + imgTexel(t, level, is, it) * weight;
                    weightSum += weight;
                }
            }
        }
        if (weightSum > 0.0f) {
            // Color::operator/=(float) multiplies by the reciprocal; MIPMap<float> divides
            result = t.is_float ? V3(result.x / weightSum, result.y / weightSum, result.z / weightSum) : result / weightSum;
        } else {
            result = imgTexel(t, level, (int)s, (int)tt);
        }
        return result;
    }
    V3 imgLookup(const gb_texture& t, const TexCoord& tc) const { // MIPMap::lookup(tc, filter, address)
        float s = tc.s, tt = tc.t;
        if (t.image_filter == GB_FILTER_BILINEAR || t.image_filter == GB_FILTER_TRILINEAR) {
            float width = std::max(std::max(fabsf(tc.dsdx), fabsf(tc.dtdx)), std::max(fabsf(tc.dsdy), fabsf(tc.dtdy)));
            if (t.image_filter == GB_FILTER_TRILINEAR) return imgTrilinear(t, s, tt, width);
            float level = t.n_levels - 1 + log2f(std::max(width, 1e-8f));
            return imgBilinear(t, floorInt(level + 0.5f), s, tt); // roundInt
        }
        if (t.image_filter != GB_FILTER_EWA) return imgBilinear(t, 0, s, tt); // lookupNearest
        float ds0 = tc.dsdx, dt0 = tc.dtdx, ds1 = tc.dsdy, dt1 = tc.dtdy; // lookupEWA
        float majorLength = (float)sqrt(ds0 * ds0 + dt0 * dt0);
        float minorLength = (float)sqrt(ds1 * ds1 + dt1 * dt1);
        if (majorLength < minorLength) {
            std::swap(ds0, ds1);
            std::swap(dt0, dt1);
            std::swap(majorLength, minorLength);
        }
        if (minorLength * t.max_anisotropy < majorLength && minorLength > 0.0f) {
            float scale = majorLength / (minorLength * t.max_anisotropy);
            minorLength *= scale;
            ds1 *= scale;
            dt1 *= scale;
        }
        float A = dt0 * dt0 + dt1 * dt1;
        float B = -2.0f * (ds0 * dt0 + ds1 * dt1);
        float C = ds0 * ds0 + ds1 * ds1;
        float F = A * C - 0.25f * B * B;
        if (minorLength == 0.0f || F <= 0.0f) return imgTrilinear(t, s, tt, minorLength);
        float invF = 1.0f / F;
        A *= invF;
        B *= invF;
        C *= invF;
        float level = t.n_levels - 1 + log2f(minorLength);
        int iLevel = floorInt(level);
        if (iLevel < 0) return imgBilinear(t, 0, s, tt);
        if (iLevel >= t.n_levels - 1) return imgBilinear(t, t.n_levels - 1, s, tt);
        float delta = level - (float)iLevel;
        return (1.0f - delta) * imgEWA(t, iLevel, s, tt, A, B, C) + (delta) * imgEWA(t, iLevel + 1, s, tt, A, B, C);
    }

    // Texture<T>::lookup; float textures live in .x
    V3 texLookup(int index, const Frag& f) const {
        const gb_texture& t = d->textures[index];
        if (t.type == GB_TEX_IMAGE) return imgLookup(t, mapTexture(t, f)); // ImageTexture::lookup
        if (t.type == GB_TEX_SCALE) { // mScale->lookup(f) * mTexture->lookup(f)
            float sc = texLookup(t.child[1], f).x;
            return texLookup(t.child[0], f) * sc;
        }
        if (t.type != GB_TEX_CHECKERBOARD) return V3(t.value[0], t.value[1], t.value[2]);
        TexCoord tc = mapTexture(t, f);
        float s = tc.s, tt = tc.t;
        if (!t.filter) {
            return (floorInt(s) + floorInt(tt)) % 2 == 0 ? texLookup(t.child[0], f) : texLookup(t.child[1], f);
        }
        float ds = std::max(fabsf(tc.dsdx), fabsf(tc.dsdy));
        float dt = std::max(fabsf(tc.dtdx), fabsf(tc.dtdy));
        float s0 = s - ds, s1 = s + ds, t0 = tt - dt, t1 = tt + dt;
        if (floorInt(s0) == floorInt(s1) && floorInt(t0) == floorInt(t1)) {
            return (floorInt(s) + floorInt(tt)) % 2 == 0 ? texLookup(t.child[0], f) : texLookup(t.child[1], f);
        }
        float sTex2Ratio = (integrateChecker(s1) - integrateChecker(s0)) / (2.0f * ds);
        float tTex2Ratio = (integrateChecker(t1) - integrateChecker(t0)) / (2.0f * dt);
        float tex2Area = sTex2Ratio + tTex2Ratio - 2.0f * sTex2Ratio * tTex2Ratio;
        if (ds > 1.0f || dt > 1.0f) tex2Area = 0.5f;
        return (1.0f - tex2Area) * texLookup(t.child[0], f) + tex2Area * texLookup(t.child[1], f);
    }
    // the material as the BSDF code reads it at this fragment: textured slots looked up
    gb_material resolveMaterial(const gb_material& m, const Frag& f) const {
        gb_material r = m;
        if (m.kd_tex) { V3 c = texLookup(m.kd_tex - 1, f); r.kd[0] = c.x; r.kd[1] = c.y; r.kd[2] = c.z; }
        if (m.kt_tex) { V3 c = texLookup(m.kt_tex - 1, f); r.kt[0] = c.x; r.kt[1] = c.y; r.kt[2] = c.z; }
        if (m.exponent_tex) r.exponent = texLookup(m.exponent_tex - 1, f).x;
        return r;
    }

    Ray cameraRay(float imageX, float imageY, float lensU1, float lensU2) const {
        const gb_camera& c = d->camera;
        float invXRes = 1.0f / (float)d->film.xres;
        float invYRes = 1.0f / (float)d->film.yres;
        float xNDC = +2.0f * imageX * invXRes - 1.0f;
        float yNDC = -2.0f * imageY * invYRes + 1.0f;
        float xView = xNDC / c.proj00;
        float yView = yNDC / c.proj11;
        V3 viewDir(xView, yView, 1.0f);
        V3 pos(c.position[0], c.position[1], c.position[2]);
        Ray ray;
        if (c.orthographic) { // OrthographicCamera::generateRay: mint 0
            ray.o = pos + quatRotate(c.orientation, V3(0.5f * c.film_width * xNDC, 0.5f * c.film_height * yNDC, 0.0f));
            ray.d = quatRotate(c.orientation, V3(0.0f, 0.0f, 1.0f));
            ray.mint = 0.0f;
            ray.maxt = INF;
            return ray;
        }
        if (c.lens_radius == 0.0f) {
            ray.o = pos;
            ray.d = quatRotate(c.orientation, normalize(viewDir));
        } else {
            float ft = c.focal_distance / viewDir.z;
            V3 pFocus = viewDir * ft;
            float lx, ly;
            uniformSampleDisk(lensU1, lensU2, &lx, &ly);
            V3 viewOrigin(c.lens_radius * lx, c.lens_radius * ly, 0.0f);
            ray.o = quatRotate(c.orientation, viewOrigin) + pos;
            ray.d = quatRotate(c.orientation, normalize(pFocus - viewOrigin));
        }
        ray.mint = 1e-3f;
        ray.maxt = INF;
        return ray;
    }

    // ---- Fragment::getWorldToShade, src/GoblinGeometry.cpp:17-29 (transposed use)
    static V3 shadeToWorld(const Frag& f, V3 v) {
        V3 n = f.n;
        V3 t = normalize(f.dpdu - n * dot(f.dpdu, n));
        V3 b = cross(n, t);
        return V3(t.x * v.x + b.x * v.y + n.x * v.z, t.y * v.x + b.y * v.y + n.y * v.z,
            t.z * v.x + b.z * v.y + n.z * v.z);
    }

    // ---- materials: src/GoblinMaterial.cpp:285-480, 647-726
    static float clampf(float f, float lo, float hi) { return f < lo ? lo : (f > hi ? hi : f); }
    static float fresnelDieletric(float cosi, float etai, float etat) {
        cosi = clampf(cosi, -1.0f, 1.0f);
        float sint = (etai / etat) * std::sqrt(std::max(0.0f, 1.0f - cosi * cosi));
        if (sint >= 1.0f) return 1.0f;
        float cost = std::sqrt(std::max(0.0f, 1 - sint * sint));
        cosi = std::fabs(cosi);
        float rParl = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
        float rPerp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
        return (rParl * rParl + rPerp * rPerp) / 2.0f;
    }
    static float fresnelConductor(float cosi, float eta, float k) {
        float tmp = (eta * eta + k * k);
        float cosi2 = cosi * cosi;
        float rParl2 = (tmp * cosi2 - 2.0f * eta * cosi + 1.0f) / (tmp * cosi2 + 2.0f * eta * cosi + 1.0f);
        float rPerp2 = (tmp - 2.0f * eta * cosi + cosi2) / (tmp + 2.0f * eta * cosi + cosi2);
        return (rParl2 + rPerp2) * 0.5f;
    }
    static float specularReflectDieletric(V3 n, V3 wo, V3* wi, float etai, float etat) {
        float cosi = dot(n, wo);
        float ei = etai, et = etat;
        bool entering = cosi > 0.0f;
        if (!entering) {
            std::swap(ei, et);
            n = -n;
            cosi = -cosi;
        }
        float f = fresnelDieletric(cosi, ei, et);
        *wi = 2 * cosi * n - wo;
        return f / cosi;
    }
    static float specularReflectConductor(V3 n, V3 wo, V3* wi, float eta, float k) {
        float cosi = dot(n, wo);
        if (cosi <= 0.0f) return 0.0f;
        float f = fresnelConductor(cosi, eta, k);
        *wi = 2 * cosi * n - wo;
        return f / cosi;
    }
    static float specularRefract(V3 n, V3 wo, V3* wi, float etao, float etai) { // mode = BSDFRadiance
        float coso = dot(n, wo);
        float et = etao, ei = etai;
        bool entering = coso > 0.0f;
        if (!entering) {
            std::swap(ei, et);
            n = -n;
            coso = -coso;
        }
        float f = fresnelDieletric(coso, et, ei);
        if (f == 1.0f) return 0.0f;
        float eta = et / ei;
        *wi = normalize(n * (eta * coso - std::sqrt(std::max(0.0f, 1.0f - eta * eta * (1.0f - coso * coso)))) -
            eta * wo);
        return eta * eta * (1.0f - f) / absdot(*wi, n);
    }
    static V3 rgb(const float* c) { return V3(c[0], c[1], c[2]); }

    // BlinnMaterial::bsdf / pdf, src/GoblinMaterial.cpp:540-573, 628-644
    static V3 blinnBsdf(const gb_material& m, const Frag& fr, V3 wo, V3 wi) {
        V3 n = fr.n;
        if (!(dot(n, wo) * dot(n, wi) > 0.0f)) return V3(); // getSampleType + matchType(Glossy | Reflection)
        float cosi = absdot(n, wi);
        float coso = absdot(n, wo);
        if (cosi == 0.0f || coso == 0.0f) return V3();
        V3 wh = normalize(wo + wi);
        float cosh = absdot(n, wh);
        float exp = m.exponent;
        float D = (exp + 2.0f) * INV_TWOPI * (float)pow(cosh, exp);
        float woDotWh = absdot(wo, wh);
        float G = std::min(1.0f, std::min(2.0f * cosh * coso / woDotWh, 2.0f * cosh * cosi / woDotWh));
        float F = m.fresnel == GB_FRESNEL_CONDUCTOR ? fresnelConductor(woDotWh, m.eta, m.k)
                                                    : fresnelDieletric(woDotWh, 1.0f, m.eta);
        return rgb(m.kd) * D * G * F / (4.0f * cosi * coso);
    }
    static float blinnPdf(const gb_material& m, const Frag& fr, V3 wo, V3 wi) {
        if (!(dot(wo, fr.n) * dot(wi, fr.n) > 0.0f)) return 0.0f;
        V3 wh = normalize(wo + wi);
        float cosThetah = absdot(wh, fr.n);
        float exp = m.exponent;
        return (exp + 1.0f) * (float)pow(cosThetah, exp) / (TWO_PI * 4.0f * dot(wo, wh));
    }

    // Material::bsdf: Lambert evaluates Kd / pi on the reflection side, the specular ones are black
    V3 bsdf(const gb_material& m, const Frag& fr, V3 wo, V3 wi) const {
        if (m.type == GB_MAT_BLINN) return blinnBsdf(m, fr, wo, wi);
        if (m.type != GB_MAT_LAMBERT) return V3();
        if (dot(fr.n, wo) * dot(fr.n, wi) > 0.0f) return rgb(m.kd) * INV_PI;
        return V3();
    }
    float bsdfPdf(const gb_material& m, const Frag& fr, V3 wo, V3 wi) const {
        if (m.type == GB_MAT_BLINN) return blinnPdf(m, fr, wo, wi);
        if (m.type != GB_MAT_LAMBERT) return 0.0f;
        return dot(wo, fr.n) * dot(wi, fr.n) > 0.0f ? absdot(fr.n, wi) * INV_PI : 0.0f;
    }
    V3 sampleBSDF(const gb_material& m, const Frag& fr, V3 wo, float uComp, float u1, float u2, V3* wi,
        float* pdf, bool* specular) const {
        *specular = m.type == GB_MAT_MIRROR || m.type == GB_MAT_TRANSPARENT;
        if (m.type == GB_MAT_BLINN) { // BlinnMaterial::sampleBSDF, src/GoblinMaterial.cpp:596-622
            float exp = m.exponent;
            float cosTheta = (float)pow(u1, 1.0f / (exp + 1.0f));
            float sinTheta = sqrtf(std::max(0.0f, 1.0f - cosTheta * cosTheta));
            float phi = u2 * TWO_PI;
            V3 whLocal(sinTheta * (float)cos(phi), sinTheta * (float)sin(phi), cosTheta);
            if (dot(wo, fr.n) < 0.0f) whLocal = whLocal * -1.0f;
            V3 wh = shadeToWorld(fr, whLocal);
            *wi = -wo + 2.0f * dot(wo, wh) * wh;
            *pdf = blinnPdf(m, fr, wo, *wi);
            return blinnBsdf(m, fr, wo, *wi);
        }
        if (m.type == GB_MAT_LAMBERT) {
            V3 wiLocal = cosineSampleHemisphere(u1, u2);
            if (dot(wo, fr.n) < 0.0f) wiLocal = wiLocal * -1.0f;
            *wi = shadeToWorld(fr, wiLocal);
            *pdf = bsdfPdf(m, fr, wo, *wi);
            return rgb(m.kd) * INV_PI;
        }
        if (m.type == GB_MAT_MIRROR) {
            V3 f = rgb(m.kd) * specularReflectConductor(fr.n, wo, wi, m.eta, m.k);
            *pdf = 1.0f;
            return f;
        }
        V3 wReflect, wRefract;
        float reflect = specularReflectDieletric(fr.n, wo, &wReflect, 1.0f, m.eta);
        float refract = specularRefract(fr.n, wo, &wRefract, 1.0f, m.eta);
        float reflectChance = reflect * absdot(wReflect, fr.n);
        if (uComp < reflectChance) {
            *wi = wReflect;
            *pdf = reflectChance;
            return rgb(m.kd) * reflect;
        }
        *wi = wRefract;
        *pdf = 1.0f - reflectChance;
        return rgb(m.kt) * refract;
    }

    // ---- lights: src/GoblinLight.cpp:87-99, 145-154, 225-237, 277-287, 368-394, 456-460
    // CDF1D::sampleDiscrete, src/GoblinSampler.cpp:333-342
    int pickLight(float u, float* pdf) const {
        const float* cdf = d->light_cdf;
        int n = (int)d->n_lights;
        const float* lb = std::lower_bound(cdf, cdf + n + 1, u);
        int offset = std::max(0, (int)(lb - cdf - 1));
        if (offset > n - 1) offset = n - 1; // unreachable for u < 1
        // mIntegral = mCDF[n] before normalisation (src/GoblinSampler.cpp:318-326)
        float dx = 1.0f / n, acc = 0.0f;
        for (int i = 0; i < n; ++i) acc = acc + d->light_power[i] * dx;
        *pdf = (d->light_power[offset] / acc) * dx;
        return offset;
    }
    // Geometry::pdf, src/GoblinGeometry.cpp:44-62, for the sphere / disk emitters
    static float genericPdf(const gb_light& l, V3 p, V3 wi) {
        Ray ray{p, wi, 1e-3f, INF};
        float eps;
        Frag fr;
        bool hit = l.geom_kind == GB_GEOM_SPHERE ? sphereIntersect(l.radius, ray, &eps, &fr)
                                                 : diskIntersect(l.radius, ray, &eps, &fr);
        if (!hit) return 0.0f;
        float pdf = sqLen(p - fr.p) / (l.area * absdot(-wi, fr.n));
        if (std::isinf(pdf)) pdf = 0.0f;
        return pdf;
    }
    // ---- mesh emitters: GeometrySet over the mesh's triangles (src/GoblinLight.cpp:289-343)
    struct TriVerts { V3 p0, p1, p2; };
    TriVerts lightFace(const gb_model& m, uint32_t face) const {
        const uint32_t* ti = d->tri_index + 3 * ((size_t)m.tri_offset + face);
        return TriVerts{vpos(m, ti[0]), vpos(m, ti[1]), vpos(m, ti[2])};
    }
    static float faceArea(const TriVerts& t) { return 0.5f * len(cross(t.p1 - t.p0, t.p2 - t.p0)); } // Triangle::area
    float meshSumArea(const gb_model& m) const {
        float sum = 0.0f;
        for (uint32_t f = 0; f < m.tri_count; ++f) sum += faceArea(lightFace(m, f));
        return sum;
    }
    // CDF1D over the face areas + sampleDiscrete (src/GoblinSampler.cpp:312-342)
    uint32_t meshPickFace(const gb_model& m, float u) const {
        const uint32_t n = m.tri_count;
        std::vector<float> cdf(n + 1, 0.0f);
        const float dx = 1.0f / n;
        for (uint32_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + faceArea(lightFace(m, i - 1)) * dx;
        const float integral = cdf[n];
        for (uint32_t i = 1; i < n + 1; ++i) cdf[i] /= integral;
        const float* lb = std::lower_bound(cdf.data(), cdf.data() + n + 1, u);
        int offset = std::max(0, (int)(lb - cdf.data() - 1));
        return (uint32_t)std::min(offset, (int)n - 1);
    }
    float meshPdf(const gb_model& m, float sumArea, V3 p, V3 wi) const {
        float pdf = 0.0f;
        for (uint32_t f = 0; f < m.tri_count; ++f) { // Geometry::pdf per face, area weighted, in face order
            Ray ray{p, wi, 1e-3f, INF};
            float eps;
            Frag fr;
            float gp = 0.0f;
            if (triangleIntersect(m, f, ray, &eps, &fr)) {
                gp = sqLen(p - fr.p) / (faceArea(lightFace(m, f)) * absdot(-wi, fr.n));
                if (std::isinf(gp)) gp = 0.0f;
            }
            pdf += faceArea(lightFace(m, f)) * gp;
        }
        pdf /= sumArea;
        return pdf;
    }

    // GeometrySet::pdf over the light's single shape (src/GoblinLight.cpp:336-343, src/GoblinSphere.cpp:138-149)
    static float shapePdf(const gb_light& l, V3 p, V3 wi) {
        float gpdf;
        if (l.geom_kind == GB_GEOM_SPHERE) {
            float squaredDistance = sqLen(p);
            float squaredRadius = l.radius * l.radius;
            if (squaredDistance - squaredRadius < 1e-4f) gpdf = genericPdf(l, p, wi);
            else {
                float sinThetaMax2 = squaredRadius / squaredDistance;
                float cosThetaMax = std::sqrt(std::max(0.0f, 1.0f - sinThetaMax2));
                gpdf = uniformConePdf(cosThetaMax);
            }
        } else {
            gpdf = genericPdf(l, p, wi);
        }
        float pdf = 0.0f;
        pdf += l.area * gpdf;
        pdf /= l.area;
        return pdf;
    }
    // ---- ImageBasedLight, src/GoblinLight.cpp:464-629
    // MIPMap::lookup(0, s, t) with repeat addressing, src/GoblinTexture.cpp:10-37,274-288
    V3 iblLookup(const gb_light& l, float s, float t) const {
        const float* img = d->image_texels + 4 * (size_t)l.image_offset;
        const int w = l.image_width, h = l.image_height;
        float sRes = s * w - 0.5f;
        float tRes = t * h - 0.5f;
        int s0 = (int)floor(sRes);
        float ds = sRes - (float)s0;
        int t0 = (int)floor(tRes);
        float dt = tRes - (float)t0;
        auto texel = [&](int ss, int tt) {
            ss = ss % w;
            tt = tt % h;
            if (ss < 0) ss += w;
            if (tt < 0) tt += h;
            const float* c = img + 4 * ((size_t)tt * w + ss);
            return V3(c[0], c[1], c[2]);
        };
        return (1.0f - ds) * (1.0f - dt) * texel(s0, t0) + (ds) * (1.0f - dt) * texel(s0 + 1, t0) +
            (1.0f - ds) * (dt) * texel(s0, t0 + 1) + (ds) * (dt) * texel(s0 + 1, t0 + 1);
    }
    // Light::Le(ray): black for every light but the image based one (:511-518)
    V3 lightLe(const gb_light& l, V3 dir) const {
        if (l.type != GB_LIGHT_IBL) return V3();
        V3 w = xfVector(l.to_object, dir); // mToWorld.invertVector
        float theta = (float)acos(clampf(w.z, -1.0f, 1.0f));
        float phi = (float)atan2(w.y, w.x);
        phi = phi < 0.0f ? phi + TWO_PI : phi;
        return iblLookup(l, phi * INV_TWOPI, theta * INV_PI);
    }
    // CDF1D::sampleContinuous, src/GoblinSampler.cpp:344-357
    static float cdfSampleContinuous(const float* func, const float* cdf, int n, float integral, float u, float* pdf,
        int* index) {
        const float* lb = std::lower_bound(cdf, cdf + n + 1, u);
        int offset = std::max(0, (int)(lb - cdf - 1));
        float dd = (u - cdf[offset]) / (cdf[offset + 1] - cdf[offset]);
        *pdf = func[offset] / integral;
        if (index) *index = offset;
        return ((float)offset + dd) / n;
    }
    struct Dist2D { const float *rowF, *rowC, *margF, *margC; float margI; int w, h; };
    Dist2D dist2D(const gb_light& l) const {
        Dist2D t;
        t.w = l.dist_width; t.h = l.dist_height;
        t.rowF = d->light_dist + l.dist_offset;
        t.rowC = t.rowF + (size_t)t.w * t.h;
        t.margF = t.rowC + (size_t)(t.w + 1) * t.h;
        t.margC = t.margF + t.h;
        t.margI = t.margC[t.h + 1];
        return t;
    }

    V3 sampleL(const gb_light& l, V3 p, float epsilon, float uComp, float u1, float u2, V3* wi, float* pdf,
        Ray* shadow) const {
        shadow->o = p;
        shadow->mint = epsilon;
        shadow->maxt = INF;
        *pdf = 1.0f;
        V3 color = rgb(l.color);
        if (l.type == GB_LIGHT_IBL) { // ImageBasedLight::sampleL, :520-545; CDF2D::sampleContinuous
            Dist2D t = dist2D(l);
            float pdfRow, pdfCol;
            int row;
            float v = cdfSampleContinuous(t.margF, t.margC, t.h, t.margI, u2, &pdfRow, &row);
            float uu = cdfSampleContinuous(t.rowF + (size_t)row * t.w, t.rowC + (size_t)row * (t.w + 1), t.w,
                t.margF[row], u1, &pdfCol, nullptr);
            float pdfST = pdfRow * pdfCol;
            float theta = v * PI;
            float phi = uu * TWO_PI;
            float cosTheta = (float)cos(theta), sinTheta = (float)sin(theta);
            float cosPhi = (float)cos(phi), sinPhi = (float)sin(phi);
            V3 wLocal(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
            *wi = xfVector(l.to_world, wLocal);
            *pdf = pdfST / (TWO_PI * PI * sinTheta); // the sinTheta == 0 guard above it is overwritten
            shadow->d = *wi;
            return iblLookup(l, uu, v);
        }
        if (l.type == GB_LIGHT_POINT || l.type == GB_LIGHT_SPOT) {
            V3 dir = rgb(l.position) - p;
            *wi = normalize(dir);
            shadow->d = *wi;
            float squaredDistance = sqLen(dir);
            shadow->maxt = std::sqrt(squaredDistance) - epsilon;
            if (l.type == GB_LIGHT_POINT) return color / squaredDistance;
            // SpotLight::falloff(-wi), src/GoblinLight.cpp:277-287
            float cosTheta = dot(-*wi, rgb(l.direction));
            float falloff;
            if (cosTheta < l.cos_theta_max) falloff = 0.0f;
            else if (cosTheta > l.cos_falloff_start) falloff = 1.0f;
            else {
                float delta = (cosTheta - l.cos_theta_max) / (l.cos_falloff_start - l.cos_theta_max);
                falloff = delta * delta * delta * delta;
            }
            return falloff * color / squaredDistance;
        }
        if (l.type == GB_LIGHT_DIRECTIONAL) {
            *wi = -rgb(l.direction);
            shadow->d = *wi;
            return color;
        }
        // AreaLight::sampleL
        V3 pLocal = xfPoint(l.to_object, p);
        V3 nsLocal, psLocal;
        if (l.geom_kind == GB_GEOM_MESH) { // GeometrySet::sample -> Triangle::sample, src/GoblinTriangle.cpp:165-178
            const gb_model& m = d->models[l.model];
            TriVerts t = lightFace(m, meshPickFace(m, uComp));
            float u1root = sqrtf(u1); // uniformSampleTriangle, src/GoblinSampler.cpp:420-424
            float b0 = 1.0f - u1root, b1 = u1root * u2;
            nsLocal = normalize(cross(t.p1 - t.p0, t.p2 - t.p0));
            psLocal = b0 * t.p0 + b1 * t.p1 + (1.0f - b0 - b1) * t.p2;
        } else if (l.geom_kind == GB_GEOM_SPHERE) { // Sphere::sample(p, u1, u2, n), src/GoblinSphere.cpp:108-136
            float squaredRadius = l.radius * l.radius;
            float squaredDistance = sqLen(pLocal);
            if (squaredDistance - squaredRadius < 1e-4f) {
                nsLocal = uniformSampleSphere(u1, u2);
                psLocal = l.radius * nsLocal;
            } else {
                V3 zAxis = normalize(-pLocal);
                V3 xAxis, yAxis;
                coordinateAxises(zAxis, &xAxis, &yAxis);
                float sinThetaMax2 = squaredRadius / squaredDistance;
                float cosThetaMax = std::sqrt(std::max(0.0f, 1.0f - sinThetaMax2));
                Ray ray{pLocal, uniformSampleCone(u1, u2, cosThetaMax, xAxis, yAxis, zAxis), 1e-3f, INF};
                Frag fr;
                float eps;
                V3 pHit;
                if (sphereIntersect(l.radius, ray, &eps, &fr)) pHit = fr.p;
                else pHit = ray.o + (std::sqrt(squaredDistance) * cosThetaMax) * ray.d;
                nsLocal = normalize(pHit);
                psLocal = pHit;
            }
        } else { // Disk::sample, src/GoblinDisk.cpp:78-82
            nsLocal = V3(0.0f, 0.0f, 1.0f);
            float px, py;
            uniformSampleDisk(u1, u2, &px, &py);
            psLocal = V3(l.radius * px, l.radius * py, 0.0f);
        }
        V3 wiLocal = normalize(psLocal - pLocal);
        if (l.geom_kind == GB_GEOM_MESH) {
            const gb_model& m = d->models[l.model];
            *pdf = meshPdf(m, meshSumArea(m), pLocal, wiLocal);
        } else {
            *pdf = shapePdf(l, pLocal, wiLocal);
        }
        V3 ps = xfPoint(l.to_world, psLocal);
        V3 ns = normalize(xfNormal(l.to_object, nsLocal));
        *wi = normalize(ps - p);
        shadow->d = *wi;
        shadow->maxt = len(ps - p) - epsilon;
        return dot(ns, -*wi) > 0.0f ? color : V3(); // AreaLight::L
    }
    float lightPdf(const gb_light& l, V3 p, V3 wi) const { // Light::pdf / AreaLight::pdf
        if (l.type == GB_LIGHT_IBL) { // ImageBasedLight::pdf, :618-629; CDF2D::pdf
            V3 wiLocal = xfVector(l.to_object, wi);
            float theta = (float)acos(clampf(wiLocal.z, -1.0f, 1.0f));
            float sinTheta = (float)sin(theta);
            if (sinTheta == 0.0f) return 0.0f;
            float phi = (float)atan2(wiLocal.y, wiLocal.x);
            phi = phi < 0.0f ? phi + TWO_PI : phi;
            Dist2D t = dist2D(l);
            float uu = phi * INV_TWOPI, v = theta * INV_PI;
            int row = std::min(std::max((int)floor(t.h * v), 0), t.h - 1);
            int col = std::min(std::max((int)floor(t.w * uu), 0), t.w - 1);
            float integral = t.margI * t.margF[row];
            float pdf2 = integral == 0.0f ? 0.0f : t.margF[row] * t.rowF[(size_t)row * t.w + col] / integral;
            return pdf2 / (TWO_PI * PI * sinTheta);
        }
        if (l.type != GB_LIGHT_AREA) return 0.0f;
        if (l.geom_kind == GB_GEOM_MESH) {
            const gb_model& m = d->models[l.model];
            return meshPdf(m, meshSumArea(m), xfPoint(l.to_object, p), xfVector(l.to_object, wi));
        }
        return shapePdf(l, xfPoint(l.to_object, p), xfVector(l.to_object, wi));
    }
    // Intersection::Le, src/GoblinPrimitive.cpp:8-14
    V3 emitted(const Isect& is, V3 outDirection) const {
        int al = d->models[d->instances[is.inst].model].area_light;
        if (al < 0) return V3();
        return dot(is.frag.n, outDirection) > 0.0f ? rgb(d->lights[al].color) : V3();
    }

    // ---- PathTracer::Li, src/GoblinPathtracer.cpp:50-179.  u: 7 floats per bounce
    // (light component, light u0, u1, bsdf component, bsdf u0, u1, pick light).
    // With no BSDFnullptr material in the supported set, trace #4 == trace #6 and
    // evalAttenuation == 1; those calls are counted (refIntersect) but not traced.
    template <typename U>
    V3 liPath(Ray ray, RayDiff rayDiff, int maxDepth, U u, Stats& st) const {
        if (d->n_lights == 0) return V3();
        V3 Li;
        float epsilon;
        Isect is;
        st.refIntersect++;
        if (!intersect(ray, &epsilon, &is, st)) { // Scene::evalEnvironmentLight, src/GoblinScene.cpp:89-95
            for (uint32_t i = 0; i < d->n_lights; ++i) Li = Li + lightLe(d->lights[i], ray.d);
            return Li;
        }
        Li = Li + emitted(is, -ray.d);
        V3 throughput(1.0f, 1.0f, 1.0f);
        for (int bounces = 0; bounces < maxDepth - 1; ++bounces) {
            float ub[7];
            u(bounces, ub);
            float pickLightPdf;
            int li = pickLight(ub[6], &pickLightPdf);
            const gb_light& light = d->lights[li];
            V3 Ld;
            computeUVDifferential(is.frag, rayDiff); // src/GoblinPathtracer.cpp:77
            const Frag& fragment = is.frag;
            const gb_material material =
                resolveMaterial(d->materials[d->models[d->instances[is.inst].model].material], fragment);
            V3 wo = -ray.d;
            V3 wi;
            V3 p = fragment.p;
            V3 n = fragment.n;
            float lightPdfV, bsdfPdfV;
            Ray shadowRay;
            V3 L = sampleL(light, p, epsilon, ub[0], ub[1], ub[2], &wi, &lightPdfV, &shadowRay);
            if (!isBlack(L) && lightPdfV > 0.0f) {
                V3 f = bsdf(material, fragment, wo, wi);
                if (!isBlack(f)) {
                    st.refOccluded++;
                    if (!occluded(shadowRay, st)) {
                        st.refIntersect++; // evalAttenuation(shadowRay): one fruitless notOpaque walk
                        if (light.type != GB_LIGHT_AREA && light.type != GB_LIGHT_IBL) { // light->isDelta()
                            Ld = Ld + f * L * absdot(n, wi) / lightPdfV;
                        } else {
                            bsdfPdfV = bsdfPdf(material, fragment, wo, wi);
                            float lWeight = powerHeuristic(lightPdfV, bsdfPdfV);
                            Ld = Ld + f * L * absdot(n, wi) * lWeight / lightPdfV;
                        }
                    }
                }
            }
            bool specular;
            V3 f = sampleBSDF(material, fragment, wo, ub[3], ub[4], ub[5], &wi, &bsdfPdfV, &specular);
            bool haveNext = false, nextHit = false;
            Isect next;
            float nextEps = 0.0f;
            Ray r{p, wi, epsilon, INF};
            if (!isBlack(f) && bsdfPdfV > 0.0f) {
                float fWeight = 1.0f;
                if (!specular) fWeight = powerHeuristic(bsdfPdfV, lightPdf(light, p, wi));
                st.refIntersect += 2; // trace #4 (isOpaque) + evalAttenuation
                haveNext = true;
                nextHit = intersect(r, &nextEps, &next, st);
                if (nextHit) {
                    int al = d->models[d->instances[next.inst].model].area_light;
                    if (al == li && light.type == GB_LIGHT_AREA) {
                        V3 Le = emitted(next, -wi);
                        if (!isBlack(Le)) Ld = Ld + f * Le * absdot(wi, n) * fWeight / bsdfPdfV;
                    }
                } else { // the radiance contribution from IBL (no cosine factor in the reference)
                    Ld = Ld + f * lightLe(light, wi) * fWeight / bsdfPdfV;
                }
            }
            Li = Li + throughput * Ld / pickLightPdf;
            if (isBlack(f) || bsdfPdfV == 0.0f) break;
            throughput = throughput * f * absdot(wi, n) / bsdfPdfV;
            st.refIntersect++; // trace #6: the same ray, no filter
            if (!haveNext) { // f != black but pdf < 0 cannot happen; keep the reference's order anyway
                nextHit = intersect(r, &nextEps, &next, st);
            }
            if (!nextHit) break;
            ray = r;
            rayDiff.has = false; // RayDifferential(p, wi, epsilon): no auxiliary rays after the camera
            is = next;
            epsilon = nextEps;
        }
        return Li;
    }

    // ---- MaskMaterial (src/GoblinMaterial.cpp:747-811) around the masked material's record
    struct MaskEval { float alpha; V3 tc; };
    MaskEval maskEval(const gb_material& m, const Frag& f) const {
        MaskEval e;
        e.alpha = m.alpha_tex ? texLookup(m.alpha_tex - 1, f).x : m.alpha;
        e.tc = m.transparent_tex ? texLookup(m.transparent_tex - 1, f) : rgb(m.transparent_color);
        return e;
    }
    bool sceneHasMask() const {
        for (uint32_t i = 0; i < d->n_materials; ++i) if (d->materials[i].mask) return true;
        return false;
    }
    // PathTracer::evalAttenuation, src/GoblinPathtracer.cpp:21-48: the product of the index-matched
    // transmittances of the not-opaque surfaces along a segment
    V3 evalAttenuation(const Ray& ray, Stats& st) const {
        V3 throughput(1.0f, 1.0f, 1.0f);
        float maxt = ray.maxt;
        Ray currentRay = ray;
        float epsilon;
        Isect is;
        while (true) {
            st.refIntersect++;
            if (!intersect(currentRay, &epsilon, &is, st, 2)) break;
            Frag fr = is.frag;
            computeUVDifferential(fr, RayDiff()); // a fresh Fragment: no differentials
            const gb_material& m = d->materials[d->models[d->instances[is.inst].model].material];
            MaskEval e = maskEval(m, fr);
            throughput = throughput * ((1.0f - e.alpha) * e.tc); // sampleBSDF(..., BSDFnullptr)
            if (throughput.x == 0.0f && throughput.y == 0.0f && throughput.z == 0.0f) break;
            currentRay.mint = currentRay.maxt + epsilon;
            currentRay.maxt = maxt;
        }
        return throughput;
    }
    // PathTracer::Li for scenes with a Mask material: every filtered trace and attenuation walk of the
    // reference is executed as written (src/GoblinPathtracer.cpp:50-179)
    template <typename U>
    V3 liPathMask(Ray ray, RayDiff rayDiff, int maxDepth, U u, Stats& st) const {
        if (d->n_lights == 0) return V3();
        V3 Li;
        float epsilon;
        Isect is;
        auto envLight = [&](V3 dir) { V3 e; for (uint32_t i = 0; i < d->n_lights; ++i) e = e + lightLe(d->lights[i], dir); return e; };
        st.refIntersect++;
        if (!intersect(ray, &epsilon, &is, st)) return envLight(ray.d);
        Li = Li + emitted(is, -ray.d);
        V3 throughput(1.0f, 1.0f, 1.0f);
        bool firstBounce = true;
        for (int bounces = 0; bounces < maxDepth - 1; ++bounces) {
            float ub[7];
            u(bounces, ub);
            float pickLightPdf;
            int li = pickLight(ub[6], &pickLightPdf);
            const gb_light& light = d->lights[li];
            V3 Ld;
            computeUVDifferential(is.frag, rayDiff);
            const Frag& fragment = is.frag;
            const gb_material& raw = d->materials[d->models[d->instances[is.inst].model].material];
            const gb_material material = resolveMaterial(raw, fragment);
            const bool mask = raw.mask != 0;
            MaskEval me{1.0f, V3()};
            if (mask) me = maskEval(raw, fragment);
            V3 wo = -ray.d;
            V3 wi;
            V3 p = fragment.p;
            V3 n = fragment.n;
            float lightPdfV, bsdfPdfV;
            Ray shadowRay;
            V3 L = sampleL(light, p, epsilon, ub[0], ub[1], ub[2], &wi, &lightPdfV, &shadowRay);
            if (!isBlack(L) && lightPdfV > 0.0f) {
                V3 f = bsdf(material, fragment, wo, wi);
                if (mask) f = f * me.alpha; // alpha * mMaskedMaterial->bsdf
                if (!isBlack(f) && (st.refOccluded++, !occluded(shadowRay, st, 1))) {
                    V3 tr = evalAttenuation(shadowRay, st);
                    if (light.type != GB_LIGHT_AREA && light.type != GB_LIGHT_IBL) {
                        Ld = Ld + f * tr * L * absdot(n, wi) / lightPdfV;
                    } else {
                        bsdfPdfV = bsdfPdf(material, fragment, wo, wi);
                        if (mask) bsdfPdfV = me.alpha * bsdfPdfV; // MaskMaterial::pdf, both parts requested
                        float lWeight = powerHeuristic(lightPdfV, bsdfPdfV);
                        Ld = Ld + f * tr * L * absdot(n, wi) * lWeight / lightPdfV;
                    }
                }
            }
            bool specular = false, nullSampled = false;
            V3 f;
            if (mask && !(ub[3] < me.alpha)) { // the index-matched pass-through
                f = (1.0f - me.alpha) * me.tc;
                wi = -normalize(wo);
                bsdfPdfV = 1.0f - me.alpha;
                nullSampled = true;
            } else {
                f = sampleBSDF(material, fragment, wo, ub[3], ub[4], ub[5], &wi, &bsdfPdfV, &specular);
                if (mask) { f = f * me.alpha; bsdfPdfV *= me.alpha; }
            }
            if (!isBlack(f) && bsdfPdfV > 0.0f) {
                if (nullSampled) {
                    throughput = throughput * (f / bsdfPdfV);
                    ray = Ray{p, wi, epsilon, INF};
                    rayDiff.has = false;
                    st.refIntersect++;
                    if (!intersect(ray, &epsilon, &is, st)) {
                        if (firstBounce) Li = Li + throughput * envLight(ray.d);
                        break;
                    }
                    continue;
                }
                float fWeight = 1.0f;
                if (!specular) fWeight = powerHeuristic(bsdfPdfV, lightPdf(light, p, wi));
                Isect lightIs;
                float lightEps;
                Ray r{p, wi, epsilon, INF};
                st.refIntersect++;
                if (intersect(r, &lightEps, &lightIs, st, 1)) {
                    V3 tr = evalAttenuation(r, st); // r.maxt has shrunk to the opaque hit
                    int al = d->models[d->instances[lightIs.inst].model].area_light;
                    if (al == li && light.type == GB_LIGHT_AREA) {
                        V3 Le = emitted(lightIs, -wi);
                        if (!isBlack(Le)) Ld = Ld + f * tr * Le * absdot(wi, n) * fWeight / bsdfPdfV;
                    }
                } else {
                    V3 tr = evalAttenuation(r, st);
                    Ld = Ld + f * tr * lightLe(light, wi) * fWeight / bsdfPdfV;
                }
            }
            Li = Li + throughput * Ld / pickLightPdf;
            if (isBlack(f) || bsdfPdfV == 0.0f) break;
            throughput = throughput * f * absdot(wi, n) / bsdfPdfV;
            ray = Ray{p, wi, epsilon, INF};
            rayDiff.has = false;
            st.refIntersect++;
            if (!intersect(ray, &epsilon, &is, st)) break;
            firstBounce = false;
        }
        return Li;
    }

    // ---- AORenderer::Li, src/GoblinAO.cpp:12-37.  u(a, out[2])
    template <typename U>
    V3 liAO(Ray ray, int aoSamples, U u, Stats& st) const {
        float epsilon;
        Isect is;
        st.refIntersect++;
        if (!intersect(ray, &epsilon, &is, st)) return V3();
        uint32_t occludedNum = 0;
        for (int n = 0; n < aoSamples; ++n) {
            float uv[2];
            u(n, uv);
            V3 sampleDir = uniformSampleHemisphere(uv[0], uv[1]);
            V3 dir = shadeToWorld(is.frag, sampleDir);
            Ray occludeRay{is.frag.p, dir, epsilon, INF};
            st.refOccluded++;
            if (occluded(occludeRay, st)) occludedNum++;
        }
        float v = (float)(aoSamples - occludedNum) / (float)aoSamples;
        return V3(v, v, v);
    }
};

// Philox4x32-10 with the product's counter layout (goblin_b200/csrc/shade.cuh): the
// sampler is the one component north_star replaces, so the oracle's render()
// draws the same numbers to make films comparable pixel by pixel.
struct Philox {
    static void gen(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
            uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
            uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += W0;
            k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    static float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
};

void block(uint64_t seed, uint64_t sampleId, uint32_t blk, float out[4]) {
    uint32_t r[4];
    Philox::gen((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)sampleId, (uint32_t)(sampleId >> 32), blk, 0u, r);
    for (int k = 0; k < 4; ++k) out[k] = Philox::u01(r[k]);
}

// The product's counter-form restatement of Sampler::requestSamples' per-pixel stratification + shuffle
// (src/GoblinSampler.cpp:108-197; goblin_b200/csrc/shade.cuh): sample s of a pixel takes stratum pi(s) of every
// dimension, pi = Kensler's hash permutation keyed by (seed, pixel, dimension), the Philox value is the jitter.
uint32_t permuteIndex(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1u;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8; i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
uint32_t strataKey(uint64_t seed, uint64_t pixel, uint32_t dim) {
    uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u) ^ ((uint32_t)pixel * 0x85EBCA6Bu) ^
                 ((uint32_t)(pixel >> 32) * 0xC2B2AE35u) ^ (dim * 0x27D4EB2Fu);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
struct Strata {
    uint64_t seed, pixel;
    uint32_t s, spp, root;
    // GO_NO_STRATA=1: plain Philox values (what round 1 drew), for the variance comparison in the tests
    bool on() const { static const bool off = std::getenv("GO_NO_STRATA") != nullptr; return spp > 1u && !off; }
    float one(float u, uint32_t dim) const {
        uint32_t k = permuteIndex(s, spp, strataKey(seed, pixel, dim));
        return std::min(((float)k + u) * (1.0f / (float)spp), 0.99999994f);
    }
    void two(float* u0, float* u1, uint32_t dim) const {
        uint32_t k = permuteIndex(s, spp, strataKey(seed, pixel, dim));
        uint32_t cy = k / root, cx = k - cy * root;
        float inv = 1.0f / (float)root;
        *u0 = std::min(((float)cx + *u0) * inv, 0.99999994f);
        *u1 = std::min(((float)cy + *u1) * inv, 0.99999994f);
    }
};
enum { DIM_LIGHT_COMP = 0, DIM_LIGHT_UV = 1, DIM_BSDF_COMP = 2, DIM_BSDF_UV = 3, DIM_PICK = 4, DIM_PER_BOUNCE = 5,
       DIM_LENS = 0x10000, DIM_AO = 0x20000 };

unsigned threadCount(int threads) {
    if (threads > 0) return (unsigned)threads;
    return std::max(1u, std::thread::hardware_concurrency());
}

template <typename F>
void parallelFor(size_t n, int threads, F body) { // body(begin, end, threadIndex)
    unsigned nt = (unsigned)std::min<size_t>(threadCount(threads), std::max<size_t>(n, 1));
    if (nt <= 1) { body(0, n, 0u); return; }
    std::atomic<size_t> next(0);
    const size_t chunk = std::max<size_t>(1, n / (nt * 16));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) {
        pool.emplace_back([&, t]() {
            while (true) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                body(b, std::min(n, b + chunk), t);
            }
        });
    }
    for (auto& th : pool) th.join();
}

void addStats(gb_counters* out, const std::vector<Stats>& per, uint64_t samples, uint64_t* refCalls) {
    Stats s;
    for (const Stats& p : per) {
        s.nodes += p.nodes; s.prims += p.prims; s.insts += p.insts; s.closest += p.closest; s.any += p.any;
        s.refIntersect += p.refIntersect; s.refOccluded += p.refOccluded;
        s.nodesAny += p.nodesAny; s.primsAny += p.primsAny; s.instsAny += p.instsAny;
    }
    if (out) {
        out->camera_samples += samples;
        out->rays_closest += s.closest;
        out->rays_any += s.any;
        out->nodes_visited += s.nodes;
        out->prims_tested += s.prims;
        out->instances_entered += s.insts;
        out->nodes_visited_any += s.nodesAny;
        out->prims_tested_any += s.primsAny;
        out->instances_entered_any += s.instsAny;
    }
    if (refCalls) { refCalls[0] += s.refIntersect; refCalls[1] += s.refOccluded; }
}

} // namespace

extern "C" {

int go_hardware_threads(void) { return (int)threadCount(0); }

// the stratum sample s of `pixel` takes in dimension `dim`, for s = 0 .. spp - 1 (tests: it must be a permutation)
int go_strata(uint64_t seed, uint64_t pixel, uint32_t dim, uint32_t spp, uint32_t* out) {
    for (uint32_t s = 0; s < spp; ++s) out[s] = permuteIndex(s, spp, strataKey(seed, pixel, dim));
    return 0;
}

// Scene::intersect on a ray batch; hits as gb_hit (inst = scene instance index, prim = face index)
int go_trace_closest(const gb_scene_desc* d, const gb_ray* rays, size_t n, gb_hit* hits, int threads,
    gb_counters* counters) {
    Oracle o(d);
    std::vector<Stats> per(threadCount(threads));
    parallelFor(n, threads, [&](size_t b, size_t e, unsigned t) {
        for (size_t i = b; i < e; ++i) {
            Ray r{V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
                rays[i].mint, rays[i].maxt};
            float eps = 0.0f;
            Isect is;
            if (o.intersect(r, &eps, &is, per[t])) {
                hits[i].t = r.maxt; hits[i].eps = eps; hits[i].inst = is.inst; hits[i].prim = is.prim;
            } else {
                hits[i].t = 0.0f; hits[i].eps = 0.0f; hits[i].inst = -1; hits[i].prim = -1;
            }
        }
    });
    addStats(counters, per, 0, nullptr);
    return 0;
}

// world-space fragment of each closest hit: p(3) n(3) dpdu(3); zeros on a miss
int go_trace_fragments(const gb_scene_desc* d, const gb_ray* rays, size_t n, float* frags) {
    Oracle o(d);
    Stats st;
    for (size_t i = 0; i < n; ++i) {
        Ray r{V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
            rays[i].mint, rays[i].maxt};
        float eps;
        Isect is;
        float* f = frags + 9 * i;
        std::memset(f, 0, 9 * sizeof(float));
        if (o.intersect(r, &eps, &is, st)) {
            f[0] = is.frag.p.x; f[1] = is.frag.p.y; f[2] = is.frag.p.z;
            f[3] = is.frag.n.x; f[4] = is.frag.n.y; f[5] = is.frag.n.z;
            f[6] = is.frag.dpdu.x; f[7] = is.frag.dpdu.y; f[8] = is.frag.dpdu.z;
        }
    }
    return 0;
}

int go_trace_any(const gb_scene_desc* d, const gb_ray* rays, size_t n, uint8_t* occ, int threads,
    gb_counters* counters) {
    Oracle o(d);
    std::vector<Stats> per(threadCount(threads));
    parallelFor(n, threads, [&](size_t b, size_t e, unsigned t) {
        for (size_t i = b; i < e; ++i) {
            Ray r{V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
                rays[i].mint, rays[i].maxt};
            occ[i] = o.occluded(r, per[t]) ? 1 : 0;
        }
    });
    addStats(counters, per, 0, nullptr);
    return 0;
}

int go_camera_rays(const gb_scene_desc* d, const float* samples, size_t n, gb_ray* rays) {
    Oracle o(d);
    for (size_t i = 0; i < n; ++i) {
        const float* s = samples + 4 * i;
        Ray r = o.cameraRay(s[0], s[1], s[2], s[3]);
        rays[i].o[0] = r.o.x; rays[i].o[1] = r.o.y; rays[i].o[2] = r.o.z;
        rays[i].d[0] = r.d.x; rays[i].d[1] = r.d.y; rays[i].d[2] = r.d.z;
        rays[i].mint = r.mint; rays[i].maxt = r.maxt;
    }
    return 0;
}

// Renderer::Li on explicit sample rows (layout of oracle/ref/ref_tool.cpp "li").
// ref_calls (optional): n x 2, the Scene::intersect / Scene::occluded calls the reference makes.
int go_li(const gb_scene_desc* d, const float* samples, size_t n, size_t row, float* out_rgb, uint32_t* ref_calls,
    int threads, gb_counters* counters) {
    Oracle o(d);
    const int method = d->setting.method;
    const int depth = std::max(1, d->setting.max_ray_depth);
    const int ao = std::max(1, d->setting.ao_sample_num);
    const bool hasMask = o.sceneHasMask();
    size_t need = 4 + (method == GB_METHOD_AO ? 2 * (size_t)ao : 7 * (size_t)depth);
    if (row < need) return 1;
    std::vector<Stats> per(threadCount(threads));
    parallelFor(n, threads, [&](size_t b, size_t e, unsigned t) {
        for (size_t i = b; i < e; ++i) {
            const float* s = samples + row * i;
            Ray ray = o.cameraRay(s[0], s[1], s[2], s[3]);
            Stats& st = per[t];
            uint64_t i0 = st.refIntersect, o0 = st.refOccluded;
            V3 L;
            if (method == GB_METHOD_AO) {
                L = o.liAO(ray, ao, [&](int a, float* uv) { uv[0] = s[4 + 2 * a]; uv[1] = s[4 + 2 * a + 1]; }, st);
            } else {
                auto ubs = [&](int bn, float* ub) { std::memcpy(ub, s + 4 + 7 * bn, 7 * sizeof(float)); };
                L = hasMask ? o.liPathMask(ray, o.cameraRayDiff(s[0], s[1], s[2], s[3]), depth, ubs, st)
                            : o.liPath(ray, o.cameraRayDiff(s[0], s[1], s[2], s[3]), depth, ubs, st);
            }
            out_rgb[3 * i] = L.x; out_rgb[3 * i + 1] = L.y; out_rgb[3 * i + 2] = L.z;
            if (ref_calls) {
                ref_calls[2 * i] = (uint32_t)(st.refIntersect - i0);
                ref_calls[2 * i + 1] = (uint32_t)(st.refOccluded - o0);
            }
        }
    });
    addStats(counters, per, n, nullptr);
    return 0;
}

// Renderer::render + ImageTile::addSample + Film::mergeTile
// (src/GoblinRenderer.cpp:29-52,99-126, src/GoblinFilm.cpp:61-90,140-153) for the sample indices
// [spp_begin, spp_end) of every pixel of the sample range, accumulated into film (yres x xres x 4:
// weighted r, g, b, weight).  Sample values: Philox, product layout (see struct Philox).
// pixel_stride > 1 renders only every pixel_stride-th sample-range pixel (bounded CPU baselines).
int go_render(const gb_scene_desc* d, const gb_render_params* p, float* film, int threads, int pixel_stride,
    gb_counters* counters, uint64_t* ref_calls) {
    Oracle o(d);
    const bool hasMask = o.sceneHasMask();
    const gb_film_desc& f = d->film;
    int method = p->method >= 0 ? p->method : d->setting.method;
    int depth = p->max_ray_depth > 0 ? p->max_ray_depth : d->setting.max_ray_depth;
    if (depth < 1) depth = 1;
    int ao = p->ao_sample_num > 0 ? p->ao_sample_num : d->setting.ao_sample_num;
    if (ao < 1) ao = 1;
    { int r = (int)std::ceil(std::sqrt((float)ao)); ao = r * r; } // SampleQuota::requestTwoDQuota, src/GoblinSampler.cpp:29-33
    const int sppTotal = p->spp_total;
    const int root = (int)std::ceil(std::sqrt((float)sppTotal));
    if (sppTotal < 1 || root * root != sppTotal) return 1;
    if (p->spp_begin < 0 || p->spp_end > sppTotal || p->spp_begin > p->spp_end) return 1;
    const int aoRoot = std::max(1, (int)sqrtf((float)ao));
    const int width = f.sx1 - f.sx0, height = f.sy1 - f.sy0;
    if (width <= 0 || height <= 0) return 0;
    if (pixel_stride < 1) pixel_stride = 1;
    const unsigned nt = threadCount(threads);
    std::vector<Stats> per(nt);
    std::vector<std::vector<float>> tiles(nt); // RenderingTLS: one full-frame tile per worker
    std::vector<uint64_t> samplesDone(nt, 0);
    const size_t nPixels = (size_t)width * height;
    parallelFor(nPixels, threads, [&](size_t b, size_t e, unsigned t) {
        std::vector<float>& tile = tiles[t];
        if (tile.empty()) tile.assign((size_t)f.xres * f.yres * 4, 0.0f);
        Stats& st = per[t];
        for (size_t pix = b; pix < e; ++pix) {
            if (pix % (size_t)pixel_stride) continue;
            int px = f.sx0 + (int)(pix % (size_t)width), py = f.sy0 + (int)(pix / (size_t)width);
            for (int s = p->spp_begin; s < p->spp_end; ++s) {
                uint64_t id = (uint64_t)pix * (uint64_t)sppTotal + (uint64_t)s;
                float u0[4];
                block(p->seed, id, 0, u0);
                const Strata strata{p->seed, (uint64_t)pix, (uint32_t)s, (uint32_t)sppTotal, (uint32_t)root};
                if (strata.on() && d->camera.lens_radius > 0.0f) strata.two(&u0[2], &u0[3], DIM_LENS);
                // Sampler::requestSamples' image stratification (src/GoblinSampler.cpp:142,192-195, 290-307)
                float sub = 1.0f / (float)root;
                float imageX = (float)px + ((float)(s % root) + u0[0]) * sub;
                float imageY = (float)py + ((float)(s / root) + u0[1]) * sub;
                Ray ray = o.cameraRay(imageX, imageY, u0[2], u0[3]);
                V3 L;
                if (method == GB_METHOD_AO) {
                    L = o.liAO(ray, ao, [&](int a, float* uv) {
                        float r4[4];
                        block(p->seed, id, 1u + ((uint32_t)a >> 1), r4);
                        float ux = (a & 1) ? r4[2] : r4[0], uy = (a & 1) ? r4[3] : r4[1];
                        if (strata.on()) { // one permutation per camera sample, a hashed cyclic shift per AO ray (shade.cuh aoCell / aoPair)
                            const uint32_t k = permuteIndex((uint32_t)s, (uint32_t)sppTotal, strataKey(p->seed, (uint64_t)pix, DIM_AO));
                            const uint32_t h = strataKey(p->seed, (uint64_t)pix, DIM_AO + 1u + (uint32_t)a);
                            uint32_t cy = k / (uint32_t)root, cx = k - cy * (uint32_t)root;
                            cx += (uint32_t)(((uint64_t)h * (uint32_t)root) >> 32);
                            cy += (uint32_t)(((uint64_t)(uint32_t)(h * 0x9E3779B1u) * (uint32_t)root) >> 32);
                            if (cx >= (uint32_t)root) cx -= (uint32_t)root;
                            if (cy >= (uint32_t)root) cy -= (uint32_t)root;
                            const float inv = 1.0f / (float)root;
                            ux = std::min(((float)cx + ux) * inv, 0.99999994f);
                            uy = std::min(((float)cy + uy) * inv, 0.99999994f);
                        }
                        float asub = 1.0f / (float)aoRoot; // stratifiedUniform2D over the pixel's AO rays
                        uv[0] = ((float)(a % aoRoot) + ux) * asub;
                        uv[1] = ((float)(a / aoRoot) + uy) * asub;
                    }, st);
                } else {
                    auto ubs = [&](int bn, float* ub) {
                        float a4[4], b4[4];
                        block(p->seed, id, 1u + 2u * (uint32_t)bn, a4);
                        block(p->seed, id, 2u + 2u * (uint32_t)bn, b4);
                        ub[0] = a4[0]; ub[1] = a4[1]; ub[2] = a4[2]; ub[3] = a4[3];
                        ub[4] = b4[0]; ub[5] = b4[1]; ub[6] = b4[2];
                        if (strata.on()) {
                            const uint32_t d0 = DIM_PER_BOUNCE * (uint32_t)bn;
                            ub[0] = strata.one(ub[0], d0 + DIM_LIGHT_COMP);
                            strata.two(&ub[1], &ub[2], d0 + DIM_LIGHT_UV);
                            ub[3] = strata.one(ub[3], d0 + DIM_BSDF_COMP);
                            strata.two(&ub[4], &ub[5], d0 + DIM_BSDF_UV);
                            ub[6] = strata.one(ub[6], d0 + DIM_PICK);
                        }
                    };
                    L = hasMask ? o.liPathMask(ray, o.cameraRayDiff(imageX, imageY, u0[2], u0[3]), depth, ubs, st)
                                : o.liPath(ray, o.cameraRayDiff(imageX, imageY, u0[2], u0[3]), depth, ubs, st);
                }
                samplesDone[t]++;
                // ImageTile::addSample, src/GoblinFilm.cpp:61-90
                if (std::isnan(L.x) || std::isnan(L.y) || std::isnan(L.z)) continue;
                float dImageX = imageX - 0.5f, dImageY = imageY - 0.5f;
                int x0 = (int)std::ceil(dImageX - f.filter_width[0]);
                int x1 = (int)std::floor(dImageX + f.filter_width[0]);
                int y0 = (int)std::ceil(dImageY - f.filter_width[1]);
                int y1 = (int)std::floor(dImageY + f.filter_width[1]);
                x0 = std::max(x0, f.xstart);
                x1 = std::min(x1, f.xstart + f.xcount - 1);
                y0 = std::max(y0, f.ystart);
                y1 = std::min(y1, f.ystart + f.ycount - 1);
                for (int y = y0; y <= y1; ++y) {
                    for (int x = x0; x <= x1; ++x) {
                        // FilterTable::evaluate, src/GoblinFilm.cpp:29-37
                        int iy = std::min((int)std::floor(std::fabs(16 * ((float)y - dImageY) / f.filter_width[1])), 15);
                        int ix = std::min((int)std::floor(std::fabs(16 * ((float)x - dImageX) / f.filter_width[0])), 15);
                        float w = f.filter_table[iy * 16 + ix];
                        float* px4 = &tile[4 * ((size_t)y * f.xres + x)];
                        px4[0] += w * L.x; px4[1] += w * L.y; px4[2] += w * L.z; px4[3] += w;
                    }
                }
            }
        }
    });
    uint64_t total = 0;
    for (unsigned t = 0; t < nt; ++t) { // Film::mergeTile
        total += samplesDone[t];
        if (tiles[t].empty()) continue;
        for (size_t k = 0; k < tiles[t].size(); ++k) film[k] += tiles[t][k];
    }
    addStats(counters, per, total, ref_calls);
    return 0;
}

} // extern "C"
