// Shading-side device code: Philox sampler, camera rays, hit-frame
// reconstruction, BSDFs, light sampling.  Each function names the reference
// code it restates (paths relative to /root/reference/src).
#pragma once
#include "rt_core.cuh"

namespace gb {

#define GB_PI 3.14159265358979323f
#define GB_TWO_PI 6.28318530718f
#define GB_INV_PI 0.31830988618379067154f
#define GB_INV_TWOPI 0.15915494309189533577f

// ----------------------------------------------------------------- sampler
// Counter-based Philox4x32-10 replaces GoblinSampler's per-tile mt19937 tables
// (north_star).  key = seed, counter = (sample id lo, hi, block, 0); one call
// yields four 24-bit uniforms.
struct Philox {
    __device__ static __forceinline__ uint4 gen(uint2 key, uint4 ctr) {
        const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            unsigned int hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
            unsigned int hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
            ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
            key.x += W0;
            key.y += W1;
        }
        return ctr;
    }
    __device__ static __forceinline__ float u01(unsigned int x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
};

// Where the sample values of one camera sample come from: Philox in a render,
// or an explicit table (gb_li, the per-sample parity surface).
// Dimension layout: block 0 = image x, image y, lens u1, lens u2;
// path tracing bounce b: block 1 + 2b = light component, light u0, light u1, bsdf component
//                        block 2 + 2b = bsdf u0, bsdf u1, pick light, (spare)
// ambient occlusion ray a: pair a of blocks 1.. (two rays per block)
// Stratification across the samples of one pixel, in counter form.  The reference fills, per pixel, every 1-D
// dimension with sample_per_pixel jittered strata and every 2-D dimension with a root x root jittered grid, then
// shuffles each dimension independently so that the samples of the pixel see the strata in random order
// (Sampler::requestSamples, stratifiedUniform1D / 2D, shuffle: src/GoblinSampler.cpp:108-197,276-333).  A table per
// pixel does not exist here: sample s of a pixel takes stratum pi(s), pi a pseudo-random permutation of [0, spp)
// keyed by (seed, pixel, dimension), and the Philox value becomes the jitter inside that stratum.  The permutation is
// Kensler's hash-based `permute` (Correlated Multi-Jittered Sampling, Pixar TM 13-01): a cycle-walking bijection on
// the next power of two.  Same estimator, same expectation; what it buys is the reference's variance per sample.
__host__ __device__ __forceinline__ unsigned int permuteIndex(unsigned int i, unsigned int l, unsigned int p) {
    unsigned int w = l - 1u;
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
__host__ __device__ __forceinline__ unsigned int strataKey(unsigned int k0, unsigned int k1, unsigned long long pixel, unsigned int dim) {
    unsigned int h = k0 ^ (k1 * 0x9E3779B1u) ^ ((unsigned int)pixel * 0x85EBCA6Bu) ^ ((unsigned int)(pixel >> 32) * 0xC2B2AE35u) ^
                     (dim * 0x27D4EB2Fu);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
// dimension numbers: 5 per bounce (light component, light position, BSDF component, BSDF direction, light pick),
// then the lens and the ambient-occlusion rays
enum { DIM_LIGHT_COMP = 0, DIM_LIGHT_UV = 1, DIM_BSDF_COMP = 2, DIM_BSDF_UV = 3, DIM_PICK = 4, DIM_PER_BOUNCE = 5,
       DIM_LENS = 0x10000, DIM_AO = 0x20000 };
#define GB_ONE_MINUS_EPS 0.99999994f

struct SampleSource {
    const float* table; // explicit rows or nullptr
    unsigned int rowFloats;
    uint2 key;
    // samples per pixel of the whole job and its square root (0 = no stratification: explicit rows, 1 spp)
    unsigned int spp, root;
    float invSpp, invRoot;
    __device__ __forceinline__ float4 block(unsigned long long sampleId, unsigned int path, unsigned int blk) const {
        if (table) {
            const float* row = table + (size_t)path * rowFloats;
            if (blk == 0) return make_float4(row[0], row[1], row[2], row[3]);
            // path tracing rows: 7 floats per bounce after the first 4
            unsigned int b = (blk - 1) >> 1;
            const float* u = row + 4 + 7 * b;
            if ((blk - 1) & 1) return make_float4(u[4], u[5], u[6], 0.0f);
            return make_float4(u[0], u[1], u[2], u[3]);
        }
        uint4 r = Philox::gen(key, make_uint4((unsigned int)sampleId, (unsigned int)(sampleId >> 32), blk, 0u));
        return make_float4(Philox::u01(r.x), Philox::u01(r.y), Philox::u01(r.z), Philox::u01(r.w));
    }
    __device__ __forceinline__ bool stratified() const { return !table && spp > 1u; }
    // jitter u -> a point of stratum pi(s) of spp strata
    __device__ __forceinline__ float strat1(float u, unsigned long long pixel, unsigned int s, unsigned int dim) const {
        const unsigned int k = permuteIndex(s, spp, strataKey(key.x, key.y, pixel, dim));
        return fminf(((float)k + u) * invSpp, GB_ONE_MINUS_EPS);
    }
    // jitter (u0, u1) -> a point of cell pi(s) of the root x root grid
    __device__ __forceinline__ void strat2(float* u0, float* u1, unsigned long long pixel, unsigned int s, unsigned int dim) const {
        const unsigned int k = permuteIndex(s, spp, strataKey(key.x, key.y, pixel, dim));
        const unsigned int cy = k / root, cx = k - cy * root;
        *u0 = fminf(((float)cx + *u0) * invRoot, GB_ONE_MINUS_EPS);
        *u1 = fminf(((float)cy + *u1) * invRoot, GB_ONE_MINUS_EPS);
    }
    // block 0 of a camera sample: image jitter (its strata are laid out by imagePosition) and the lens sample
    __device__ __forceinline__ float4 cameraBlock(unsigned long long sampleId, unsigned int path, unsigned long long pixel,
        unsigned int s, bool lens) const {
        float4 u = block(sampleId, path, 0);
        if (lens && stratified()) strat2(&u.z, &u.w, pixel, s, DIM_LENS);
        return u;
    }
    // the seven numbers of bounce b: (light component, light u0, u1, BSDF component) and (BSDF u0, u1, light pick).
    // `need`: the dimensions the caller's material / lights can read (bit = DIM_*); a dimension nothing reads is left
    // as the plain Philox value -- unobservable, and each permutation costs ~50 integer instructions per path.
    __device__ __forceinline__ void bounceBlocks(unsigned long long sampleId, unsigned int path, unsigned long long pixel,
        unsigned int s, unsigned int bounce, unsigned int need, float4* uA, float4* uB) const {
        *uA = block(sampleId, path, 1u + 2u * bounce);
        *uB = block(sampleId, path, 2u + 2u * bounce);
        if (!stratified()) return;
        const unsigned int d0 = DIM_PER_BOUNCE * bounce;
        if (need & (1u << DIM_LIGHT_COMP)) uA->x = strat1(uA->x, pixel, s, d0 + DIM_LIGHT_COMP);
        if (need & (1u << DIM_LIGHT_UV)) strat2(&uA->y, &uA->z, pixel, s, d0 + DIM_LIGHT_UV);
        if (need & (1u << DIM_BSDF_COMP)) uA->w = strat1(uA->w, pixel, s, d0 + DIM_BSDF_COMP);
        if (need & (1u << DIM_BSDF_UV)) strat2(&uB->x, &uB->y, pixel, s, d0 + DIM_BSDF_UV);
        if (need & (1u << DIM_PICK)) uB->z = strat1(uB->z, pixel, s, d0 + DIM_PICK);
    }
    // ambient occlusion: ray a of a camera sample; the caller places the pair in cell a of the pixel's aoRoot x aoRoot
    // grid of directions, this sub-stratifies the cell over the samples of the pixel as the reference does
    // (stratifiedUniform2D(buffer, n): n strata x sample_per_pixel sub-strata).  One permutation per camera sample
    // (aoCell, computed once per hit by k_ao_frames) gives the sample its sub-cell; ray a shifts it by its own hashed
    // offset, cyclically per axis -- for every a still a bijection between the pixel's samples and the sub-cells, at a
    // hash instead of a permutation per ray (25 rays per sample: the per-ray permutation cost 8 % of k_ao).
    __device__ __forceinline__ unsigned int aoCell(unsigned long long pixel, unsigned int s) const {
        if (!stratified()) return 0u;
        const unsigned int k = permuteIndex(s, spp, strataKey(key.x, key.y, pixel, DIM_AO));
        const unsigned int cy = k / root;
        return (cy << 16) | (k - cy * root);
    }
    __device__ __forceinline__ float2 aoPair(unsigned long long sampleId, unsigned int path, unsigned long long pixel,
        unsigned int cell, unsigned int a) const {
        if (table) {
            const float* row = table + (size_t)path * rowFloats;
            return make_float2(row[4 + 2 * a], row[4 + 2 * a + 1]);
        }
        uint4 r = Philox::gen(key, make_uint4((unsigned int)sampleId, (unsigned int)(sampleId >> 32), 1u + (a >> 1), 0u));
        float2 u = (a & 1) ? make_float2(Philox::u01(r.z), Philox::u01(r.w)) : make_float2(Philox::u01(r.x), Philox::u01(r.y));
        if (stratified()) {
            const unsigned int h = strataKey(key.x, key.y, pixel, DIM_AO + 1u + a);
            unsigned int cx = (cell & 0xffffu) + __umulhi(h, root), cy = (cell >> 16) + __umulhi(h * 0x9E3779B1u, root);
            cx = cx >= root ? cx - root : cx;
            cy = cy >= root ? cy - root : cy;
            u.x = fminf(((float)cx + u.x) * invRoot, GB_ONE_MINUS_EPS);
            u.y = fminf(((float)cy + u.y) * invRoot, GB_ONE_MINUS_EPS);
        }
        return u;
    }
};

// ------------------------------------------------------------ sampling maps
// GoblinSampler.cpp:449-602, GoblinUtils.cpp:58-69
__device__ __forceinline__ float3 cosineSampleHemisphere(float u1, float u2) {
    float sinTheta = sqrtf(u1);
    float cosTheta = sqrtf(fmaxf(0.0f, 1.0f - u1));
    float phi = GB_TWO_PI * u2;
    float s, c;
    sincosf(phi, &s, &c);
    return make3(sinTheta * c, sinTheta * s, cosTheta);
}
__device__ __forceinline__ float3 uniformSampleHemisphere(float u1, float u2) {
    float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - u1 * u1));
    float phi = GB_TWO_PI * u2;
    float s, c;
    sincosf(phi, &s, &c);
    return make3(sinTheta * c, sinTheta * s, u1);
}
__device__ __forceinline__ float3 uniformSampleSphere(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float phi = GB_TWO_PI * u2;
    float s, c;
    sincosf(phi, &s, &c);
    return make3(sinTheta * c, sinTheta * s, z);
}
__device__ __forceinline__ float3 uniformSampleCone(float u1, float u2, float cosThetaMax, float3 x, float3 y, float3 z) {
    float cosTheta = 1.0f - u1 + u1 * cosThetaMax;
    float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
    float phi = GB_TWO_PI * u2;
    float s, c;
    sincosf(phi, &s, &c);
    return x * sinTheta * c + y * sinTheta * s + z * cosTheta;
}
__device__ __forceinline__ float uniformConePdf(float cosThetaMax) { return 1.0f / (GB_TWO_PI * (1.0f - cosThetaMax)); }
__device__ __forceinline__ float2 uniformSampleDisk(float u1, float u2) { // Shirley-Chiu concentric map
    float r, theta;
    float x = 2.0f * u1 - 1.0f;
    float y = 2.0f * u2 - 1.0f;
    if (x + y > 0) {
        if (x > y) { r = x; theta = 0.25f * GB_PI * (y / x); }
        else { r = y; theta = 0.25f * GB_PI * (2.0f - x / y); }
    } else {
        if (x < y) { r = -x; theta = 0.25f * GB_PI * (4.0f + y / x); }
        else {
            r = -y;
            theta = y != 0.0f ? 0.25f * GB_PI * (6.0f - x / y) : 0.0f;
        }
    }
    float s, c;
    sincosf(theta, &s, &c);
    return make_float2(r * c, r * s);
}
__device__ __forceinline__ void coordinateAxises(float3 a1, float3* a2, float3* a3) {
    if (fabsf(a1.x) > fabsf(a1.y)) {
        float invLen = 1.0f / sqrtf(a1.x * a1.x + a1.z * a1.z);
        *a2 = make3(-a1.z * invLen, 0.0f, a1.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a1.y * a1.y + a1.z * a1.z);
        *a2 = make3(0.0f, -a1.z * invLen, a1.y * invLen);
    }
    *a3 = cross3(a1, *a2);
}
__device__ __forceinline__ float powerHeuristic(float pdfA, float pdfB) { // nA = nB = 1, GoblinSampler.h:286-290
    return pdfA * pdfA / (pdfA * pdfA + pdfB * pdfB);
}
__device__ __forceinline__ float absdot3(float3 a, float3 b) { return fabsf(dot3(a, b)); }
__device__ __forceinline__ bool isBlack(float3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; }
__device__ __forceinline__ float3 mul3(float3 a, float3 b) { return make3(a.x * b.x, a.y * b.y, a.z * b.z); }

// ------------------------------------------------------------------ camera
// Quaternion::operator*(Vector3), GoblinQuaternion.cpp:87-93
__device__ __forceinline__ float3 quatRotate(const float q[4], float3 p) {
    float3 v = make3(q[1], q[2], q[3]);
    float3 uv = cross3(v, p);
    float3 uuv = cross3(v, uv);
    uv = uv * (2.0f * q[0]);
    uuv = uuv * 2.0f;
    return p + uv + uuv;
}

// PerspectiveCamera::generateRay, GoblinCamera.cpp:97-148 (the differential
// rays are not generated: every supported texture is constant)
__device__ __forceinline__ void cameraRay(const DeviceScene& sc, float imageX, float imageY, float lensU1,
    float lensU2, float3* o, float3* d) {
    const gb_camera& cam = sc.camera;
    float xNDC = +2.0f * imageX * sc.invXRes - 1.0f;
    float yNDC = -2.0f * imageY * sc.invYRes + 1.0f;
    float xView = xNDC / cam.proj00;
    float yView = yNDC / cam.proj11;
    float3 viewDir = make3(xView, yView, 1.0f);
    float3 pos = make3(cam.position[0], cam.position[1], cam.position[2]);
    if (cam.orthographic) { // OrthographicCamera::generateRay, GoblinCamera.cpp:301-329
        float3 pView = make3(0.5f * cam.film_width * xNDC, 0.5f * cam.film_height * yNDC, 0.0f);
        *o = pos + quatRotate(cam.orientation, pView);
        *d = quatRotate(cam.orientation, make3(0.0f, 0.0f, 1.0f));
        return;
    }
    if (cam.lens_radius == 0.0f) {
        *o = pos;
        *d = quatRotate(cam.orientation, normalize3(viewDir));
    } else {
        float ft = cam.focal_distance / viewDir.z;
        float3 pFocus = viewDir * ft;
        float2 ls = uniformSampleDisk(lensU1, lensU2);
        float3 viewOrigin = make3(cam.lens_radius * ls.x, cam.lens_radius * ls.y, 0.0f);
        *o = quatRotate(cam.orientation, viewOrigin) + pos;
        *d = quatRotate(cam.orientation, normalize3(pFocus - viewOrigin));
    }
}

// ray->mint of a camera ray: 1e-3 (perspective, GoblinCamera.cpp:143), 0 (orthographic, :324)
__device__ __forceinline__ float cameraMint(const DeviceScene& sc) { return sc.camera.orthographic ? 0.0f : 1e-3f; }

// ------------------------------------------------------------ hit fragment
struct Frag {
    float3 p;    // world position
    float3 n;    // world shading normal
    float3 dpdu; // world
    int material;
    int areaLight;
};

// Rebuilds what Triangle/Sphere/Disk::intersect wrote into the Fragment and
// what InstancedPrimitive::intersect then transformed to world space
// (GoblinTriangle.cpp:76-123, GoblinSphere.cpp:33-79, GoblinDisk.cpp:29-64,
// GoblinGeometry.cpp:31-37).  o, d: the world-space ray that produced `hit`.
__device__ __forceinline__ Frag buildFragment(const DeviceScene& sc, const HitRec& hit, float3 o, float3 d) {
    Frag f;
    unsigned int slot = (unsigned int)hit.inst;
    const float4* mo = sc.instToObject + 3 * (size_t)slot;
    const float4* mw = sc.instToWorld + 3 * (size_t)slot;
    float4 i0 = __ldg(mo), i1 = __ldg(mo + 1), i2 = __ldg(mo + 2);
    float4 w0 = __ldg(mw), w1 = __ldg(mw + 1), w2 = __ldg(mw + 2);
    int4 info = __ldg(sc.instInfo + slot);
    int4 sh = __ldg(sc.instShade + slot);
    f.material = sh.z;
    f.areaLight = sh.w;
    float3 oo = xfPoint(i0, i1, i2, o);
    float3 od = xfVector(i0, i1, i2, d);
    float3 pObj = oo + hit.t * od; // ray(t) in object space
    float3 nObj, dpduObj;
    if (info.x == GB_GEOM_MESH) {
        const float4* tr = sc.triRec + kTriRecVec4 * (size_t)(info.z + hit.prim);
        float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
        float3 e1 = make3(a.w, b.x, b.y), e2 = make3(b.z, b.w, c.x);
        int4 ms = __ldg(sc.modelShade + sh.y);
        float b1 = hit.b1, b2 = hit.b2;
        float b0 = 1.0f - b1 - b2;
        const float4* ts = sc.triShade + 4 * (size_t)(info.z + hit.prim);
        if (ms.z & 1) {
            float4 s0 = __ldg(ts), s1 = __ldg(ts + 1);
            float n2z = __ldg(ts + 2).x;
            float3 nn = b0 * make3(s0.x, s0.y, s0.z) + b1 * make3(s0.w, s1.x, s1.y) + b2 * make3(s1.z, s1.w, n2z);
            nObj = normalize3(nn);
        } else {
            nObj = normalize3(cross3(e1, e2));
        }
        float du1 = 1.0f, dv1 = 0.0f, du2 = 0.0f, dv2 = 1.0f;
        if (ms.z & 2) {
            float4 s2 = __ldg(ts + 2), s3 = __ldg(ts + 3);
            float u0 = s2.y, vv0 = s2.z;
            du1 = s2.w - u0; dv1 = s3.x - vv0;
            du2 = s3.y - u0; dv2 = s3.z - vv0;
        }
        float determinant = du1 * dv2 - dv1 * du2;
        if (determinant == 0.0f) {
            // the reference reads a stale fragment here (GoblinTriangle.cpp:113-117,
            // reachable only with degenerate vt); use an arbitrary tangent instead
            float3 t2v;
            coordinateAxises(nObj, &dpduObj, &t2v);
        } else {
            float invDet = 1.0f / determinant;
            dpduObj = invDet * (dv2 * e1 - dv1 * e2);
        }
    } else if (info.x == GB_GEOM_SPHERE) {
        nObj = normalize3(pObj);
        dpduObj = make3(-GB_TWO_PI * pObj.y, GB_TWO_PI * pObj.x, 0.0f);
    } else {
        nObj = make3(0.0f, 0.0f, 1.0f);
        dpduObj = make3(-GB_TWO_PI * pObj.y, GB_TWO_PI * pObj.x, 0.0f);
    }
    // Fragment::transform: point by M, normal by inverse-transpose (then
    // normalised), tangent by M
    f.p = xfPoint(w0, w1, w2, pObj);
    float3 nw = make3(i0.x * nObj.x + i1.x * nObj.y + i2.x * nObj.z,
                      i0.y * nObj.x + i1.y * nObj.y + i2.y * nObj.z,
                      i0.z * nObj.x + i1.z * nObj.y + i2.z * nObj.z);
    f.n = normalize3(nw);
    f.dpdu = xfVector(w0, w1, w2, dpduObj);
    return f;
}

// Fragment::getWorldToShade (GoblinGeometry.cpp:17-29): rows t, b, n.
struct ShadeFrame { float3 t, b, n; };
__device__ __forceinline__ ShadeFrame makeFrame(const Frag& f) {
    ShadeFrame s;
    s.n = f.n;
    s.t = normalize3(f.dpdu - f.n * dot3(f.dpdu, f.n));
    s.b = cross3(s.n, s.t);
    return s;
}
__device__ __forceinline__ float3 shadeToWorld(const ShadeFrame& s, float3 v) { // transpose(worldToShade) * v
    return make3(s.t.x * v.x + s.b.x * v.y + s.n.x * v.z,
                 s.t.y * v.x + s.b.y * v.y + s.n.y * v.z,
                 s.t.z * v.x + s.b.z * v.y + s.n.z * v.z);
}

// ------------------------------------------------------------------- BSDFs
// GoblinMaterial.cpp:285-304 (helpers), 306-416 (specular), 437-480 (Lambert),
// 647-726 (Transparent / Mirror sampling)
__device__ __forceinline__ float clampf(float f, float lo, float hi) { return f < lo ? lo : (f > hi ? hi : f); }

__device__ __forceinline__ float fresnelDieletric(float cosi, float etai, float etat) {
    cosi = clampf(cosi, -1.0f, 1.0f);
    float sint = (etai / etat) * sqrtf(fmaxf(0.0f, 1.0f - cosi * cosi));
    if (sint >= 1.0f) return 1.0f;
    float cost = sqrtf(fmaxf(0.0f, 1 - sint * sint));
    cosi = fabsf(cosi);
    float rParl = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
    float rPerp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
    return (rParl * rParl + rPerp * rPerp) / 2.0f;
}
__device__ __forceinline__ float fresnelConductor(float cosi, float eta, float k) {
    float tmp = (eta * eta + k * k);
    float cosi2 = cosi * cosi;
    float rParl2 = (tmp * cosi2 - 2.0f * eta * cosi + 1.0f) / (tmp * cosi2 + 2.0f * eta * cosi + 1.0f);
    float rPerp2 = (tmp - 2.0f * eta * cosi + cosi2) / (tmp + 2.0f * eta * cosi + cosi2);
    return (rParl2 + rPerp2) * 0.5f;
}
__device__ __forceinline__ float specularReflectDieletric(float3 n, float3 wo, float3* wi, float etai, float etat) {
    float cosi = dot3(n, wo);
    float ei = etai, et = etat;
    if (!(cosi > 0.0f)) {
        float tmp = ei; ei = et; et = tmp;
        n = -n;
        cosi = -cosi;
    }
    float f = fresnelDieletric(cosi, ei, et);
    *wi = 2 * cosi * n - wo;
    return f / cosi;
}
__device__ __forceinline__ float specularReflectConductor(float3 n, float3 wo, float3* wi, float eta, float k) {
    float cosi = dot3(n, wo);
    if (cosi <= 0.0f) return 0.0f;
    float f = fresnelConductor(cosi, eta, k);
    *wi = 2 * cosi * n - wo;
    return f / cosi;
}
__device__ __forceinline__ float specularRefract(float3 n, float3 wo, float3* wi, float etao, float etai) {
    float coso = dot3(n, wo);
    float et = etao, ei = etai;
    if (!(coso > 0.0f)) {
        float tmp = ei; ei = et; et = tmp;
        n = -n;
        coso = -coso;
    }
    float f = fresnelDieletric(coso, et, ei);
    if (f == 1.0f) return 0.0f; // total internal reflection
    float eta = et / ei;
    *wi = normalize3(n * (eta * coso - sqrtf(fmaxf(0.0f, 1.0f - eta * eta * (1.0f - coso * coso)))) - eta * wo);
    return eta * eta * (1.0f - f) / absdot3(*wi, n); // BSDFRadiance
}

struct BsdfSample {
    float3 f;
    float3 wi;
    float pdf;
    bool specular;
};

__device__ __forceinline__ float3 lambertEval(const DeviceMaterial& m, float3 n, float3 wo, float3 wi) {
    // getSampleType + matchType: only the reflection hemisphere matches Diffuse | Reflection
    if (dot3(n, wo) * dot3(n, wi) > 0.0f) {
        return make3(m.kdType.x * GB_INV_PI, m.kdType.y * GB_INV_PI, m.kdType.z * GB_INV_PI);
    }
    return make3(0.0f, 0.0f, 0.0f);
}
__device__ __forceinline__ float lambertPdf(float3 n, float3 wo, float3 wi) {
    return dot3(wo, n) * dot3(wi, n) > 0.0f ? absdot3(n, wi) * GB_INV_PI : 0.0f;
}

// BlinnMaterial (Torrance-Sparrow with a Blinn distribution), GoblinMaterial.cpp:540-644.
// DeviceMaterial packing for blinn: kdType = (Kg, type), ktEta = (k, exponent, fresnel type, eta).
__device__ __forceinline__ float3 blinnEval(const DeviceMaterial& m, float3 n, float3 wo, float3 wi) {
    if (!(dot3(n, wo) * dot3(n, wi) > 0.0f)) return make3(0.0f, 0.0f, 0.0f); // Glossy | Reflection only
    float cosi = absdot3(n, wi);
    float coso = absdot3(n, wo);
    if (cosi == 0.0f || coso == 0.0f) return make3(0.0f, 0.0f, 0.0f);
    float3 wh = normalize3(wo + wi);
    float cosh = absdot3(n, wh);
    float e = m.ktEta.y;
    float D = (e + 2.0f) * GB_INV_TWOPI * powf(cosh, e);
    float woDotWh = absdot3(wo, wh);
    float G = fminf(1.0f, fminf(2.0f * cosh * coso / woDotWh, 2.0f * cosh * cosi / woDotWh));
    float F = __float_as_int(m.ktEta.z) == GB_FRESNEL_CONDUCTOR ? fresnelConductor(woDotWh, m.ktEta.w, m.ktEta.x)
                                                                : fresnelDieletric(woDotWh, 1.0f, m.ktEta.w);
    float scale = D * G * F;
    float denom = 4.0f * cosi * coso;
    // Color * D * G * F / (4 cosi coso): Color::operator* per factor, then operator/(float)
    float3 c = make3(m.kdType.x * D, m.kdType.y * D, m.kdType.z * D);
    c = c * G;
    c = c * F;
    (void)scale;
    return div3(c, denom);
}
__device__ __forceinline__ float blinnPdf(const DeviceMaterial& m, float3 n, float3 wo, float3 wi) {
    if (!(dot3(wo, n) * dot3(wi, n) > 0.0f)) return 0.0f;
    float3 wh = normalize3(wo + wi);
    float cosThetah = absdot3(wh, n);
    float e = m.ktEta.y;
    return (e + 1.0f) * powf(cosThetah, e) / (GB_TWO_PI * 4.0f * dot3(wo, wh));
}

__device__ __forceinline__ BsdfSample sampleBsdf(const DeviceMaterial& m, int type, const Frag& fr, float3 wo,
    float uComp, float u1, float u2) {
    BsdfSample s;
    s.f = make3(0.0f, 0.0f, 0.0f);
    s.wi = make3(0.0f, 0.0f, 0.0f);
    s.pdf = 0.0f;
    s.specular = type == GB_MAT_MIRROR || type == GB_MAT_TRANSPARENT;
    if (type == GB_MAT_BLINN) {
        float e = m.ktEta.y;
        float cosTheta = powf(u1, 1.0f / (e + 1.0f));
        float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
        float phi = u2 * GB_TWO_PI;
        float sp, cp;
        sincosf(phi, &sp, &cp);
        float3 whLocal = make3(sinTheta * cp, sinTheta * sp, cosTheta);
        if (dot3(wo, fr.n) < 0.0f) whLocal = whLocal * -1.0f;
        ShadeFrame sf = makeFrame(fr);
        float3 wh = shadeToWorld(sf, whLocal);
        s.wi = -wo + (2.0f * dot3(wo, wh)) * wh;
        s.pdf = blinnPdf(m, fr.n, wo, s.wi);
        s.f = blinnEval(m, fr.n, wo, s.wi);
    } else if (type == GB_MAT_LAMBERT) {
        float3 wiLocal = cosineSampleHemisphere(u1, u2);
        if (dot3(wo, fr.n) < 0.0f) wiLocal = wiLocal * -1.0f;
        ShadeFrame sf = makeFrame(fr);
        s.wi = shadeToWorld(sf, wiLocal);
        s.pdf = lambertPdf(fr.n, wo, s.wi);
        s.f = make3(m.kdType.x * GB_INV_PI, m.kdType.y * GB_INV_PI, m.kdType.z * GB_INV_PI);
    } else if (type == GB_MAT_MIRROR) {
        float3 wi = make3(0.0f, 0.0f, 0.0f);
        float r = specularReflectConductor(fr.n, wo, &wi, m.ktEta.w, m.ktEta.x /* k */);
        s.f = make3(m.kdType.x * r, m.kdType.y * r, m.kdType.z * r);
        s.wi = wi;
        s.pdf = 1.0f;
    } else { // transparent: choose reflect / refract with the Fresnel factor
        float3 wReflect = make3(0.0f, 0.0f, 0.0f), wRefract = make3(0.0f, 0.0f, 0.0f);
        float etat = m.ktEta.w;
        float reflect = specularReflectDieletric(fr.n, wo, &wReflect, 1.0f, etat);
        float refract = specularRefract(fr.n, wo, &wRefract, 1.0f, etat);
        float reflectChance = reflect * absdot3(wReflect, fr.n);
        if (uComp < reflectChance) {
            s.f = make3(m.kdType.x * reflect, m.kdType.y * reflect, m.kdType.z * reflect);
            s.wi = wReflect;
            s.pdf = reflectChance;
        } else {
            s.f = make3(m.ktEta.x * refract, m.ktEta.y * refract, m.ktEta.z * refract);
            s.wi = wRefract;
            s.pdf = 1.0f - reflectChance;
        }
    }
    return s;
}

// ------------------------------------------------------------------ lights
// CDF1D::sampleDiscrete (GoblinSampler.cpp:333-342): std::lower_bound on the CDF
__device__ __forceinline__ int pickLight(const DeviceScene& sc, float u, float* pdf) {
    int n = (int)sc.nLights + 1;
    int lo = 0, count = n;
    while (count > 0) { // first element not less than u
        int step = count >> 1;
        if (__ldg(sc.lightCdf + lo + step) < u) { lo += step + 1; count -= step + 1; }
        else count = step;
    }
    int offset = max(0, lo - 1);
    // u beyond the last entry (cannot happen for u < 1 with a normalised CDF) would index past the end
    offset = min(offset, (int)sc.nLights - 1);
    *pdf = (__ldg(sc.lightPower + offset) / sc.lightIntegral) * (1.0f / sc.nLights);
    return offset;
}

// Geometry::pdf (GoblinGeometry.cpp:44-62) for a sphere / disk emitter, in the
// light's local space.  wi need not be unit length (AreaLight::pdf passes the
// inverse-transformed world direction as is).
__device__ __forceinline__ float genericShapePdf(int kind, float radius, float area, float3 p, float3 wi) {
    float t;
    float3 nHit;
    float3 pHit;
    if (kind == GB_GEOM_SPHERE) {
        if (!sphereTest(radius, p, wi, 1e-3f, INFINITY, &t)) return 0.0f;
        pHit = p + t * wi;
        nHit = normalize3(pHit);
    } else {
        if (!diskTest(radius, p, wi, 1e-3f, INFINITY, &t)) return 0.0f;
        pHit = p + t * wi;
        nHit = make3(0.0f, 0.0f, 1.0f);
    }
    float pdf = sqLen3(p - pHit) / (area * absdot3(-wi, nHit));
    if (isinf(pdf)) pdf = 0.0f;
    return pdf;
}
// GeometrySet::pdf over a mesh emitter (GoblinLight.cpp:336-343): every face's Geometry::pdf
// (GoblinGeometry.cpp:44-62 through Triangle::intersect, GoblinTriangle.cpp:38-125), area
// weighted, summed in face order.  O(faces) per evaluation, exactly like the reference.
// (kept out of line: mesh emitters are rare and their loop must not cost the common shade path registers)
__device__ __noinline__ float meshLightPdf(const DeviceScene& sc, unsigned int triBase, unsigned int triCount,
    bool hasNormal, float sumArea, float3 p, float3 wi) {
    float pdf = 0.0f;
    for (unsigned int i = 0; i < triCount; ++i) {
        const float4* lt = sc.lightTris + 6 * (size_t)(triBase + i);
        const float4 a = __ldg(lt), b = __ldg(lt + 1), c = __ldg(lt + 2);
        const float3 p0 = make3(a.x, a.y, a.z), p1 = make3(b.x, b.y, b.z), p2 = make3(c.x, c.y, c.z);
        const float3 e1 = p1 - p0, e2 = p2 - p0;
        float t, b1, b2, gp = 0.0f;
        if (triangleTest(p0, e1, e2, p, wi, 1e-3f, INFINITY, &t, &b1, &b2)) {
            float3 nHit;
            if (hasNormal) {
                const float4 n0 = __ldg(lt + 3), n1 = __ldg(lt + 4), n2 = __ldg(lt + 5);
                const float b0 = 1.0f - b1 - b2;
                nHit = normalize3(b0 * make3(n0.x, n0.y, n0.z) + b1 * make3(n1.x, n1.y, n1.z) + b2 * make3(n2.x, n2.y, n2.z));
            } else {
                nHit = normalize3(cross3(e1, e2));
            }
            const float3 pHit = p + t * wi;
            gp = sqLen3(p - pHit) / (a.w * absdot3(-wi, nHit));
            if (isinf(gp)) gp = 0.0f;
        }
        pdf += a.w * gp;
    }
    pdf /= sumArea;
    return pdf;
}

// GeometrySet::sample -> Triangle::sample (GoblinLight.cpp:308-320, GoblinTriangle.cpp:165-178):
// face by area CDF (CDF1D::sampleDiscrete = std::lower_bound), point by uniformSampleTriangle.
__device__ __noinline__ void sampleMeshEmitter(const DeviceScene& sc, unsigned int triBase, unsigned int triCount,
    unsigned int cdfBase, float uComp, float u1, float u2, float3* ps, float3* ns) {
    const float* cdf = sc.lightTriCdf + cdfBase;
    int lo = 0, count = (int)triCount + 1;
    while (count > 0) {
        int step = count >> 1;
        if (__ldg(cdf + lo + step) < uComp) { lo += step + 1; count -= step + 1; }
        else count = step;
    }
    const int face = min(max(0, lo - 1), (int)triCount - 1);
    const float4* lt = sc.lightTris + 6 * (size_t)(triBase + (unsigned int)face);
    const float4 a = __ldg(lt), b = __ldg(lt + 1), c = __ldg(lt + 2);
    const float3 p0 = make3(a.x, a.y, a.z), p1 = make3(b.x, b.y, b.z), p2 = make3(c.x, c.y, c.z);
    const float u1root = sqrtf(u1); // uniformSampleTriangle
    const float b0 = 1.0f - u1root, b1 = u1root * u2;
    *ns = normalize3(cross3(p1 - p0, p2 - p0));
    *ps = b0 * p0 + b1 * p1 + (1.0f - b0 - b1) * p2;
}

// Sphere::pdf (GoblinSphere.cpp:138-149) / Disk -> Geometry::pdf, then
// GeometrySet::pdf's area weighting (GoblinLight.cpp:336-343) for one shape.
__device__ __forceinline__ float shapePdf(int kind, float radius, float area, float3 p, float3 wi) {
    float pdf;
    if (kind == GB_GEOM_SPHERE) {
        float squaredDistance = sqLen3(p);
        float squaredRadius = radius * radius;
        if (squaredDistance - squaredRadius < 1e-4f) {
            pdf = genericShapePdf(kind, radius, area, p, wi);
        } else {
            float sinThetaMax2 = squaredRadius / squaredDistance;
            float cosThetaMax = sqrtf(fmaxf(0.0f, 1.0f - sinThetaMax2));
            pdf = uniformConePdf(cosThetaMax);
        }
    } else {
        pdf = genericShapePdf(kind, radius, area, p, wi);
    }
    float sum = 0.0f;
    sum += area * pdf;
    sum /= area;
    return sum;
}

// ---- ImageBasedLight (GoblinLight.cpp:464-629).  DeviceLight packing for GB_LIGHT_IBL: posRadius =
// int bits (image width, image height, distribution width, distribution height), dirCos = uint bits
// (first texel in sc.imageTexels, first float in sc.lightDist, 0, 0), toWorld / toObject = orientation.
// MIPMap::lookup(0, s, t), repeat addressing (GoblinTexture.cpp:10-37,274-288)
__device__ __noinline__ float3 iblLookup(const float4* __restrict__ img, int w, int h, float s, float t) {
    const float sRes = s * w - 0.5f;
    const float tRes = t * h - 0.5f;
    const int s0 = (int)floorf(sRes);
    const float ds = sRes - (float)s0;
    const int t0 = (int)floorf(tRes);
    const float dt = tRes - (float)t0;
    int sa = s0 % w, sb = (s0 + 1) % w, ta = t0 % h, tb = (t0 + 1) % h;
    if (sa < 0) sa += w;
    if (sb < 0) sb += w;
    if (ta < 0) ta += h;
    if (tb < 0) tb += h;
    const float4 a = __ldg(img + (size_t)ta * w + sa), b = __ldg(img + (size_t)ta * w + sb);
    const float4 c = __ldg(img + (size_t)tb * w + sa), d = __ldg(img + (size_t)tb * w + sb);
    const float wa = (1.0f - ds) * (1.0f - dt), wb = ds * (1.0f - dt), wc = (1.0f - ds) * dt, wd = ds * dt;
    return make3(((a.x * wa + b.x * wb) + c.x * wc) + d.x * wd, ((a.y * wa + b.y * wb) + c.y * wc) + d.y * wd,
        ((a.z * wa + b.z * wb) + c.z * wc) + d.z * wd);
}
// ImageBasedLight::Le(ray): the map looked up along a world direction
__device__ __noinline__ float3 iblLe(const DeviceScene& sc, const DeviceLight& l, float3 dir) {
    const float4 i0 = __ldg(&l.toObject[0]), i1 = __ldg(&l.toObject[1]), i2 = __ldg(&l.toObject[2]);
    const float4 dims = __ldg(&l.posRadius), offs = __ldg(&l.dirCos);
    const float3 w = xfVector(i0, i1, i2, dir); // mToWorld.invertVector
    const float theta = acosf(fminf(fmaxf(w.z, -1.0f), 1.0f));
    float phi = atan2f(w.y, w.x);
    phi = phi < 0.0f ? phi + GB_TWO_PI : phi;
    return iblLookup(sc.imageTexels + __float_as_uint(offs.x), __float_as_int(dims.x), __float_as_int(dims.y),
        phi * GB_INV_TWOPI, theta * GB_INV_PI);
}
// std::lower_bound over a CDF of n + 1 entries, then CDF1D::sampleContinuous (GoblinSampler.cpp:344-357)
__device__ __forceinline__ float cdfSampleContinuous(const float* __restrict__ func, const float* __restrict__ cdf, int n,
    float integral, float u, float* pdf, int* index) {
    int lo = 0, len = n + 1;
    while (len > 0) { // first entry that is not < u
        const int half = len >> 1;
        if (__ldg(cdf + lo + half) < u) { lo += half + 1; len -= half + 1; }
        else len = half;
    }
    const int offset = max(0, lo - 1);
    const float c0 = __ldg(cdf + offset), c1 = __ldg(cdf + offset + 1);
    const float d = (u - c0) / (c1 - c0);
    *pdf = __ldg(func + offset) / integral;
    *index = offset;
    return ((float)offset + d) / (float)n;
}

struct LightSampleResult {
    float3 L;
    float3 wi;
    float pdf;
    float maxt; // shadow ray [eps, maxt]
    bool delta;
};

// Light::sampleL for each light type (GoblinLight.cpp:87-99 point, 145-154
// directional, 225-237 spot, 368-394 area) and Light::isDelta.
// ML: the scene has mesh emitters (compiled out otherwise: their loop would cost every shade
// kernel registers).
template <bool ML>
__device__ __forceinline__ LightSampleResult sampleLight(const DeviceScene& sc, int li, float3 p, float eps,
    float uComp, float u1, float u2) {
    LightSampleResult r;
    const DeviceLight& l = sc.lights[li];
    float4 ct = __ldg(&l.colorType);
    int type = __float_as_int(ct.w);
    float3 color = make3(ct.x, ct.y, ct.z);
    r.delta = type != GB_LIGHT_AREA && type != GB_LIGHT_IBL;
    r.pdf = 1.0f;
    r.maxt = INFINITY;
    if (ML && type == GB_LIGHT_IBL) { // ImageBasedLight::sampleL + CDF2D::sampleContinuous
        const float4 dims = __ldg(&l.posRadius), offs = __ldg(&l.dirCos);
        const int dw = __float_as_int(dims.z), dh = __float_as_int(dims.w);
        const float* rowF = sc.lightDist + __float_as_uint(offs.y);
        const float* rowC = rowF + (size_t)dw * dh;
        const float* margF = rowC + (size_t)(dw + 1) * dh;
        const float* margC = margF + dh;
        float pdfRow, pdfCol;
        int row, col;
        const float v = cdfSampleContinuous(margF, margC, dh, __ldg(margC + dh + 1), u2, &pdfRow, &row);
        const float uu = cdfSampleContinuous(rowF + (size_t)row * dw, rowC + (size_t)row * (dw + 1), dw, __ldg(margF + row), u1,
            &pdfCol, &col);
        const float pdfST = pdfRow * pdfCol;
        float sinTheta, cosTheta, sinPhi, cosPhi;
        sincosf(v * GB_PI, &sinTheta, &cosTheta);
        sincosf(uu * GB_TWO_PI, &sinPhi, &cosPhi);
        const float4 w0 = __ldg(&l.toWorld[0]), w1 = __ldg(&l.toWorld[1]), w2 = __ldg(&l.toWorld[2]);
        r.wi = xfVector(w0, w1, w2, make3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));
        r.pdf = pdfST / (GB_TWO_PI * GB_PI * sinTheta);
        r.L = iblLookup(sc.imageTexels + __float_as_uint(offs.x), __float_as_int(dims.x), __float_as_int(dims.y), uu, v);
        return r;
    }
    if (type == GB_LIGHT_POINT || type == GB_LIGHT_SPOT) {
        float4 pr = __ldg(&l.posRadius);
        float3 dir = make3(pr.x, pr.y, pr.z) - p;
        r.wi = normalize3(dir);
        float squaredDistance = sqLen3(dir);
        r.maxt = sqrtf(squaredDistance) - eps;
        float scale = 1.0f;
        if (type == GB_LIGHT_SPOT) { // SpotLight::falloff(-wi)
            float4 dc = __ldg(&l.dirCos);
            float cosFalloffStart = __ldg(&l.misc).x;
            float cosTheta = dot3(-r.wi, make3(dc.x, dc.y, dc.z));
            if (cosTheta < dc.w) scale = 0.0f;
            else if (cosTheta > cosFalloffStart) scale = 1.0f;
            else {
                float delta = (cosTheta - dc.w) / (cosFalloffStart - dc.w);
                scale = delta * delta * delta * delta;
            }
            // falloff * mIntensity / squaredDistance
            float3 c = color * scale;
            r.L = div3(c, squaredDistance);
        } else {
            r.L = div3(color, squaredDistance);
        }
    } else if (type == GB_LIGHT_DIRECTIONAL) {
        float4 dc = __ldg(&l.dirCos);
        r.wi = -make3(dc.x, dc.y, dc.z);
        r.L = color;
    } else {
        float4 pr = __ldg(&l.posRadius);
        float4 misc = __ldg(&l.misc);
        float radius = pr.w, area = misc.y;
        int kind = __float_as_int(misc.z);
        float4 w0 = __ldg(&l.toWorld[0]), w1 = __ldg(&l.toWorld[1]), w2 = __ldg(&l.toWorld[2]);
        float4 i0 = __ldg(&l.toObject[0]), i1 = __ldg(&l.toObject[1]), i2 = __ldg(&l.toObject[2]);
        float3 pLocal = xfPoint(i0, i1, i2, p);
        float3 nsLocal, psLocal;
        unsigned int mTriBase = 0, mTriCount = 0;
        bool mHasNormal = false;
        if (ML && kind == GB_GEOM_MESH) {
            const float4 dc = __ldg(&l.dirCos);
            mTriBase = __float_as_uint(dc.x);
            mTriCount = __float_as_uint(dc.y);
            mHasNormal = __float_as_uint(dc.z) != 0u;
            sampleMeshEmitter(sc, mTriBase, mTriCount, __float_as_uint(dc.w), uComp, u1, u2, &psLocal, &nsLocal);
        } else if (kind == GB_GEOM_SPHERE) { // Sphere::sample(p, u1, u2, n), GoblinSphere.cpp:108-136
            float squaredRadius = radius * radius;
            float squaredDistance = sqLen3(pLocal);
            if (squaredDistance - squaredRadius < 1e-4f) {
                nsLocal = uniformSampleSphere(u1, u2);
                psLocal = radius * nsLocal;
            } else {
                float3 zAxis = normalize3(-pLocal);
                float3 xAxis, yAxis;
                coordinateAxises(zAxis, &xAxis, &yAxis);
                float sinThetaMax2 = squaredRadius / squaredDistance;
                float cosThetaMax = sqrtf(fmaxf(0.0f, 1.0f - sinThetaMax2));
                float3 dir = uniformSampleCone(u1, u2, cosThetaMax, xAxis, yAxis, zAxis);
                float t;
                float3 pHit;
                if (sphereTest(radius, pLocal, dir, 1e-3f, INFINITY, &t)) pHit = pLocal + t * dir;
                else pHit = pLocal + (sqrtf(squaredDistance) * cosThetaMax) * dir;
                nsLocal = normalize3(pHit);
                psLocal = pHit;
            }
        } else { // Disk::sample, GoblinDisk.cpp:66-70
            nsLocal = make3(0.0f, 0.0f, 1.0f);
            float2 pxy = uniformSampleDisk(u1, u2);
            psLocal = make3(radius * pxy.x, radius * pxy.y, 0.0f);
        }
        float3 wiLocal = normalize3(psLocal - pLocal);
        r.pdf = (ML && kind == GB_GEOM_MESH) ? meshLightPdf(sc, mTriBase, mTriCount, mHasNormal, area, pLocal, wiLocal)
                                     : shapePdf(kind, radius, area, pLocal, wiLocal);
        float3 ps = xfPoint(w0, w1, w2, psLocal);
        float3 nw = make3(i0.x * nsLocal.x + i1.x * nsLocal.y + i2.x * nsLocal.z,
                          i0.y * nsLocal.x + i1.y * nsLocal.y + i2.y * nsLocal.z,
                          i0.z * nsLocal.x + i1.z * nsLocal.y + i2.z * nsLocal.z);
        float3 ns = normalize3(nw);
        r.wi = normalize3(ps - p);
        r.maxt = len3(ps - p) - eps;
        r.L = dot3(ns, -r.wi) > 0.0f ? color : make3(0.0f, 0.0f, 0.0f); // AreaLight::L
    }
    return r;
}

// Light::pdf(p, wi): 0 for delta lights (GoblinLight.h:110-112), AreaLight::pdf otherwise (GoblinLight.cpp:456-460)
template <bool ML>
__device__ __forceinline__ float lightPdf(const DeviceScene& sc, int li, float3 p, float3 wi) {
    const DeviceLight& l = sc.lights[li];
    int type = __float_as_int(__ldg(&l.colorType).w);
    if (ML && type == GB_LIGHT_IBL) { // ImageBasedLight::pdf + CDF2D::pdf
        const float4 j0 = __ldg(&l.toObject[0]), j1 = __ldg(&l.toObject[1]), j2 = __ldg(&l.toObject[2]);
        const float3 wl = xfVector(j0, j1, j2, wi);
        const float theta = acosf(fminf(fmaxf(wl.z, -1.0f), 1.0f));
        const float sinTheta = sinf(theta);
        if (sinTheta == 0.0f) return 0.0f;
        float phi = atan2f(wl.y, wl.x);
        phi = phi < 0.0f ? phi + GB_TWO_PI : phi;
        const float4 dims = __ldg(&l.posRadius), offs = __ldg(&l.dirCos);
        const int dw = __float_as_int(dims.z), dh = __float_as_int(dims.w);
        const float* rowF = sc.lightDist + __float_as_uint(offs.y);
        const float* margF = rowF + (size_t)dw * dh + (size_t)(dw + 1) * dh;
        const float margI = __ldg(margF + dh + dh + 1);
        const int row = min(max((int)floorf(dh * (theta * GB_INV_PI)), 0), dh - 1);
        const int col = min(max((int)floorf(dw * (phi * GB_INV_TWOPI)), 0), dw - 1);
        const float mf = __ldg(margF + row);
        const float integral = margI * mf;
        const float pdf2 = integral == 0.0f ? 0.0f : mf * __ldg(rowF + (size_t)row * dw + col) / integral;
        return pdf2 / (GB_TWO_PI * GB_PI * sinTheta);
    }
    if (type != GB_LIGHT_AREA) return 0.0f;
    float4 pr = __ldg(&l.posRadius);
    float4 misc = __ldg(&l.misc);
    float4 i0 = __ldg(&l.toObject[0]), i1 = __ldg(&l.toObject[1]), i2 = __ldg(&l.toObject[2]);
    float3 pLocal = xfPoint(i0, i1, i2, p);
    float3 wiLocal = xfVector(i0, i1, i2, wi);
    if (ML && __float_as_int(misc.z) == GB_GEOM_MESH) {
        const float4 dc = __ldg(&l.dirCos);
        return meshLightPdf(sc, __float_as_uint(dc.x), __float_as_uint(dc.y), __float_as_uint(dc.z) != 0u, misc.y, pLocal,
            wiLocal);
    }
    return shapePdf(__float_as_int(misc.z), pr.w, misc.y, pLocal, wiLocal);
}

} // namespace gb
