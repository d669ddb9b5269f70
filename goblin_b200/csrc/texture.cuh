// Procedural textures on the device (src/GoblinTexture.cpp:292-427): constant,
// checkerboard (point-sampled or box-filtered with the primary ray's uv
// differentials), scale; uv and spherical mappings.  A material slot that is
// not a plain constant points at a postfix program over the texture table
// (children before parents, compiled at upload), evaluated on a small value
// stack.  Everything here is out of line and only reached when the scene has
// a textured material (DeviceScene::matTex != nullptr), so scenes with constant
// textures pay one uniform branch per shaded hit.
#pragma once
#include "shade.cuh"

namespace gb {

constexpr int kTexStack = 8;     // value stack of the postfix evaluation (checked at upload)
constexpr int kTexNodeVec4 = 7;  // float4 per texture node

// What the textures read of the Fragment beyond buildFragment's p / n / dpdu.
struct TexFrag {
    float3 p, n, dpdu, dpdv;
    float u, v;
    float3 dpdx, dpdy;
    float dudx, dvdx, dudy, dvdy;
};

// The uv and dpdv that Triangle / Sphere / Disk::intersect store (GoblinTriangle.cpp:96-121,
// GoblinSphere.cpp:62-77, GoblinDisk.cpp:50-61), dpdv taken to world space like
// Fragment::transform does (GoblinGeometry.cpp:31-37).
__device__ __noinline__ void texFragment(const DeviceScene& sc, const HitRec& hit, float3 o, float3 d, const Frag& fr,
    TexFrag* tf) {
    const unsigned int slot = (unsigned int)hit.inst;
    const float4* mo = sc.instToObject + 3 * (size_t)slot;
    const float4* mw = sc.instToWorld + 3 * (size_t)slot;
    const float4 i0 = __ldg(mo), i1 = __ldg(mo + 1), i2 = __ldg(mo + 2);
    const float4 w0 = __ldg(mw), w1 = __ldg(mw + 1), w2 = __ldg(mw + 2);
    const int4 info = __ldg(sc.instInfo + slot);
    const int4 sh = __ldg(sc.instShade + slot);
    const float3 oo = xfPoint(i0, i1, i2, o);
    const float3 od = xfVector(i0, i1, i2, d);
    const float3 pObj = oo + hit.t * od;
    float3 dpdvObj;
    if (info.x == GB_GEOM_MESH) {
        const float4* tr = sc.triRec + kTriRecVec4 * (size_t)(info.z + hit.prim);
        const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
        const float3 e1 = make3(a.w, b.x, b.y), e2 = make3(b.z, b.w, c.x);
        const int4 ms = __ldg(sc.modelShade + sh.y);
        const float b1 = hit.b1, b2 = hit.b2;
        const float b0 = 1.0f - b1 - b2;
        float u0 = 0.0f, v0 = 0.0f, u1 = 1.0f, v1 = 0.0f, u2 = 0.0f, v2 = 1.0f;
        if (ms.z & 2) {
            const float4* ts = sc.triShade + 4 * (size_t)(info.z + hit.prim);
            const float4 s2 = __ldg(ts + 2), s3 = __ldg(ts + 3);
            u0 = s2.y; v0 = s2.z; u1 = s2.w; v1 = s3.x; u2 = s3.y; v2 = s3.z;
        }
        tf->u = b0 * u0 + b1 * u1 + b2 * u2;
        tf->v = b0 * v0 + b1 * v1 + b2 * v2;
        const float du1 = u1 - u0, dv1 = v1 - v0, du2 = u2 - u0, dv2 = v2 - v0;
        const float determinant = du1 * dv2 - dv1 * du2;
        if (determinant == 0.0f) { // divergence D1: an arbitrary frame around the normal
            float3 t1; // the same frame buildFragment picks for dpdu
            float3 nObj = normalize3(cross3(e1, e2));
            if (ms.z & 1) {
                const float4* ts = sc.triShade + 4 * (size_t)(info.z + hit.prim);
                const float4 q0 = __ldg(ts), q1 = __ldg(ts + 1);
                const float n2z = __ldg(ts + 2).x;
                nObj = normalize3(b0 * make3(q0.x, q0.y, q0.z) + b1 * make3(q0.w, q1.x, q1.y) + b2 * make3(q1.z, q1.w, n2z));
            }
            coordinateAxises(nObj, &t1, &dpdvObj);
        } else {
            const float invDet = 1.0f / determinant;
            dpdvObj = invDet * (-du2 * e1 + du1 * e2);
        }
    } else if (info.x == GB_GEOM_SPHERE) {
        const float radius = __int_as_float(info.w);
        float phi = atan2f(pObj.y, pObj.x);
        if (phi < 0.0f) phi += GB_TWO_PI;
        tf->u = phi * GB_INV_TWOPI;
        const float theta = acosf(pObj.z / radius);
        tf->v = theta * GB_INV_PI;
        const float invR = 1.0f / sqrtf(pObj.x * pObj.x + pObj.y * pObj.y);
        const float cosPhi = pObj.x * invR, sinPhi = pObj.y * invR;
        dpdvObj = GB_PI * make3(pObj.z * cosPhi, pObj.z * sinPhi, -radius * sinf(theta));
    } else {
        const float radius = __int_as_float(info.w);
        const float r = sqrtf(pObj.x * pObj.x + pObj.y * pObj.y);
        float phi = atan2f(pObj.y, pObj.x);
        if (phi < 0.0f) phi += GB_TWO_PI;
        tf->u = phi * GB_INV_TWOPI;
        tf->v = r / radius;
        dpdvObj = make3(radius * pObj.x / r, radius * pObj.y / r, 0.0f);
    }
    tf->p = fr.p;
    tf->n = fr.n;
    tf->dpdu = fr.dpdu;
    tf->dpdv = xfVector(w0, w1, w2, dpdvObj);
    tf->dpdx = tf->dpdy = make3(0.0f, 0.0f, 0.0f);
    tf->dudx = tf->dvdx = tf->dudy = tf->dvdy = 0.0f;
}

// solve2x2LinearSystem, GoblinUtils.h:151-164
__device__ __forceinline__ bool solve2x2(float a00, float a01, float a10, float a11, float b0, float b1, float* x,
    float* y) {
    const float det = a00 * a11 - a01 * a10;
    if (fabsf(det) < 1e-10f) return false;
    *x = (+a11 * b0 - a01 * b1) / det;
    *y = (-a10 * b0 + a00 * b1) / det;
    return !(isnan(*x) || isnan(*y));
}
__device__ __forceinline__ float comp3(float3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// The camera's dx / dy auxiliary rays (PerspectiveCamera::generateRay, GoblinCamera.cpp:103-142) and
// Intersection::computeUVDifferential (GoblinPrimitive.cpp:32-97) for a primary hit.
__device__ __noinline__ void texDifferentials(const DeviceScene& sc, float imageX, float imageY, float lensU1,
    float lensU2, TexFrag* tf) {
    const gb_camera& cam = sc.camera;
    const float xNDC = +2.0f * imageX * sc.invXRes - 1.0f;
    const float yNDC = -2.0f * imageY * sc.invYRes + 1.0f;
    const float dxNDC = +2.0f * (imageX + 1.0f) * sc.invXRes - 1.0f;
    const float dyNDC = -2.0f * (imageY + 1.0f) * sc.invYRes + 1.0f;
    const float xView = xNDC / cam.proj00, yView = yNDC / cam.proj11;
    const float3 viewDir = make3(xView, yView, 1.0f);
    const float3 dxViewDir = make3(dxNDC / cam.proj00, yView, 1.0f);
    const float3 dyViewDir = make3(xView, dyNDC / cam.proj11, 1.0f);
    const float3 pos = make3(cam.position[0], cam.position[1], cam.position[2]);
    float3 oAux, oAuxY, dxD, dyD;
    if (cam.orthographic) { // OrthographicCamera::generateRay: shifted origins, one direction
        oAux = pos + quatRotate(cam.orientation, make3(0.5f * cam.film_width * dxNDC, 0.5f * cam.film_height * yNDC, 0.0f));
        oAuxY = pos + quatRotate(cam.orientation, make3(0.5f * cam.film_width * xNDC, 0.5f * cam.film_height * dyNDC, 0.0f));
        dxD = dyD = quatRotate(cam.orientation, make3(0.0f, 0.0f, 1.0f));
    } else if (cam.lens_radius == 0.0f) {
        oAux = pos;
        dxD = quatRotate(cam.orientation, normalize3(dxViewDir));
        dyD = quatRotate(cam.orientation, normalize3(dyViewDir));
    } else {
        const float ft = cam.focal_distance / viewDir.z;
        const float2 ls = uniformSampleDisk(lensU1, lensU2);
        const float3 viewOrigin = make3(cam.lens_radius * ls.x, cam.lens_radius * ls.y, 0.0f);
        oAux = quatRotate(cam.orientation, viewOrigin) + pos;
        dxD = quatRotate(cam.orientation, normalize3(dxViewDir * ft - viewOrigin));
        dyD = quatRotate(cam.orientation, normalize3(dyViewDir * ft - viewOrigin));
    }
    if (!cam.orthographic) oAuxY = oAux; // perspective: both auxiliary rays start at the ray origin
    const float3 p = tf->p, n = tf->n;
    const float minusD = dot3(p, n);
    const float tdx = (minusD - dot3(oAux, n)) / dot3(dxD, n);
    const float tdy = (minusD - dot3(oAuxY, n)) / dot3(dyD, n);
    if (isnan(tdx) || isnan(tdy)) return;
    const float3 dpdx = (oAux + tdx * dxD) - p;
    const float3 dpdy = (oAuxY + tdy * dyD) - p;
    tf->dpdx = dpdx;
    tf->dpdy = dpdy;
    int a0, a1;
    if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) { a0 = 1; a1 = 2; }
    else if (fabsf(n.y) > fabsf(n.z)) { a0 = 0; a1 = 2; }
    else { a0 = 0; a1 = 1; }
    const float A00 = comp3(tf->dpdu, a0), A01 = comp3(tf->dpdv, a0), A10 = comp3(tf->dpdu, a1), A11 = comp3(tf->dpdv, a1);
    float x, y;
    if (solve2x2(A00, A01, A10, A11, comp3(dpdx, a0), comp3(dpdx, a1), &x, &y)) { tf->dudx = x; tf->dvdx = y; }
    if (solve2x2(A00, A01, A10, A11, comp3(dpdy, a0), comp3(dpdy, a1), &x, &y)) { tf->dudy = x; tf->dvdy = y; }
}

__device__ __forceinline__ int floorInt(float f) { return (int)floorf(f); }
__device__ __forceinline__ float integrateChecker(float x) {
    const float xHalf = 0.5f * x;
    return floorf(xHalf) + 2.0f * fmaxf(xHalf - floorf(xHalf) - 0.5f, 0.0f);
}
// SphericalMapping::pointToST, GoblinTexture.cpp:339-346
__device__ __forceinline__ void pointToST(float4 r0, float4 r1, float4 r2, float3 p, float* s, float* t) {
    const float3 v = normalize3(xfPoint(r0, r1, r2, p));
    const float theta = acosf(clampf(v.z, -1.0f, 1.0f));
    float phi = atan2f(v.y, v.x);
    phi = phi < 0.0f ? phi + GB_TWO_PI : phi;
    *s = phi * GB_INV_TWOPI;
    *t = theta * GB_INV_PI;
}

// ---- image textures: MIPMap<T> lookups (GoblinTexture.cpp:10-37,82-288).  Node record of an image
// texture: [0] = (max anisotropy, -, -, type), [1] int bits = (filter, mapping, address mode, first level),
// [2] = uv scale / offset, [3..5] = spherical rows with int bits (level count, is_float) in [5].w ... kept in
// TexImage below.  Level table: int4 (width, height, first texel, 0) per level.
struct TexImage {
    const int4* levels;     // this texture's level 0
    const float4* texels;
    int nLevels, address;
    bool isFloat;
};
__device__ __forceinline__ float3 imgTexel(const TexImage& im, int level, int s, int t) { // ImageBuffer::texel
    const int4 l = __ldg(im.levels + level);
    if (im.address == GB_ADDRESS_CLAMP) {
        s = min(max(s, 0), l.x - 1);
        t = min(max(s, 0), l.y - 1); // sic: the reference clamps s into t
    } else if (im.address == GB_ADDRESS_BORDER) {
        if (s < 0 || t < 0 || s >= l.x || t >= l.y) return make3(0.0f, 0.0f, 0.0f);
    } else {
        s = s % l.x;
        t = t % l.y;
        if (s < 0) s += l.x;
        if (t < 0) t += l.y;
    }
    const float4 c = __ldg(im.texels + (size_t)(unsigned int)l.z + (size_t)t * l.x + s);
    return make3(c.x, c.y, c.z);
}
__device__ __noinline__ float3 imgBilinear(const TexImage& im, int level, float s, float t) { // MIPMap::lookup(level, s, t, m)
    level = min(max(level, 0), im.nLevels - 1); // the reference indexes one past the pyramid at the top
    const int4 l = __ldg(im.levels + level);
    const float sRes = s * l.x - 0.5f;
    const float tRes = t * l.y - 0.5f;
    const int s0 = floorInt(sRes);
    const float ds = sRes - (float)s0;
    const int t0 = floorInt(tRes);
    const float dt = tRes - (float)t0;
    return (1.0f - ds) * (1.0f - dt) * imgTexel(im, level, s0, t0) + (ds) * (1.0f - dt) * imgTexel(im, level, s0 + 1, t0) +
        (1.0f - ds) * (dt) * imgTexel(im, level, s0, t0 + 1) + (ds) * (dt) * imgTexel(im, level, s0 + 1, t0 + 1);
}
__device__ __forceinline__ float3 imgTrilinear(const TexImage& im, float s, float t, float width) {
    const float level = im.nLevels - 1 + log2f(fmaxf(width, 1e-8f));
    const int iLevel = floorInt(level);
    if (iLevel < 0) return imgBilinear(im, 0, s, t);
    if (iLevel >= im.nLevels - 1) return imgBilinear(im, im.nLevels - 1, s, t);
    const float delta = level - (float)iLevel;
    return (1.0f - delta) * imgBilinear(im, iLevel, s, t) + (delta) * imgBilinear(im, iLevel + 1, s, t);
}
__device__ __noinline__ float3 imgEWA(const TexImage& im, int level, float s, float t, float A, float B, float C) {
    const int4 l = __ldg(im.levels + level);
    const float sRes = (float)l.x, tRes = (float)l.y;
    s = s * l.x - 0.5f;
    t = t * l.y - 0.5f;
    A = A / (sRes * sRes);
    B = B / (sRes * tRes);
    C = C / (tRes * tRes);
    const float invDet = 1.0f / (-B * B + 4.0f * A * C);
    const float offsetS = 2.0f * sqrtf(C * invDet);
    const float offsetT = 2.0f * sqrtf(A * invDet);
    const int s0 = (int)ceilf(s - offsetS), s1 = floorInt(s + offsetS);
    const int t0 = (int)ceilf(t - offsetT), t1 = floorInt(t + offsetT);
    float weightSum = 0.0f;
    float3 result = make3(0.0f, 0.0f, 0.0f);
    for (int is = s0; is <= s1; ++is) {
        for (int it = t0; it <= t1; ++it) {
            const float ss = is - s, tt = it - t;
            const float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
            if (r2 <= 1.0f) {
                const int lutIndex = min(floorInt(r2 * 128.0f), 127);
                // MIPMap::initEWALut: expf(-2 r2') - expf(-2) on 128 entries
                const float weight = expf(-2.0f * ((float)lutIndex / 127.0f)) - expf(-2.0f);
                result = result + imgTexel(im, level, is, it) * weight;
                weightSum += weight;
            }
        }
    }
    if (weightSum > 0.0f) {
        // Color::operator/=(float) multiplies by the reciprocal; MIPMap<float> divides
        if (im.isFloat) result = make3(result.x / weightSum, result.y / weightSum, result.z / weightSum);
        else result = result * (1.0f / weightSum);
    } else {
        result = imgTexel(im, level, (int)s, (int)t);
    }
    return result;
}
// MIPMap::lookup(tc, filter, address)
__device__ __noinline__ float3 imgLookup(const TexImage& im, int filter, float maxAniso, float s, float t, float dsdx,
    float dtdx, float dsdy, float dtdy) {
    if (filter == GB_FILTER_BILINEAR || filter == GB_FILTER_TRILINEAR) {
        const float width = fmaxf(fmaxf(fabsf(dsdx), fabsf(dtdx)), fmaxf(fabsf(dsdy), fabsf(dtdy)));
        if (filter == GB_FILTER_TRILINEAR) return imgTrilinear(im, s, t, width);
        const float level = im.nLevels - 1 + log2f(fmaxf(width, 1e-8f));
        return imgBilinear(im, floorInt(level + 0.5f), s, t);
    }
    if (filter != GB_FILTER_EWA) return imgBilinear(im, 0, s, t); // lookupNearest
    float ds0 = dsdx, dt0 = dtdx, ds1 = dsdy, dt1 = dtdy; // lookupEWA
    float majorLength = sqrtf(ds0 * ds0 + dt0 * dt0);
    float minorLength = sqrtf(ds1 * ds1 + dt1 * dt1);
    if (majorLength < minorLength) {
        float x = ds0; ds0 = ds1; ds1 = x;
        x = dt0; dt0 = dt1; dt1 = x;
        x = majorLength; majorLength = minorLength; minorLength = x;
    }
    if (minorLength * maxAniso < majorLength && minorLength > 0.0f) {
        const float scale = majorLength / (minorLength * maxAniso);
        minorLength *= scale;
        ds1 *= scale;
        dt1 *= scale;
    }
    float A = dt0 * dt0 + dt1 * dt1;
    float B = -2.0f * (ds0 * dt0 + ds1 * dt1);
    float C = ds0 * ds0 + ds1 * ds1;
    const float F = A * C - 0.25f * B * B;
    if (minorLength == 0.0f || F <= 0.0f) return imgTrilinear(im, s, t, minorLength);
    const float invF = 1.0f / F;
    A *= invF;
    B *= invF;
    C *= invF;
    const float level = im.nLevels - 1 + log2f(minorLength);
    const int iLevel = floorInt(level);
    if (iLevel < 0) return imgBilinear(im, 0, s, t);
    if (iLevel >= im.nLevels - 1) return imgBilinear(im, im.nLevels - 1, s, t);
    const float delta = level - (float)iLevel;
    return (1.0f - delta) * imgEWA(im, iLevel, s, t, A, B, C) + (delta) * imgEWA(im, iLevel + 1, s, t, A, B, C);
}

// Texture<T>::lookup of the texture a program ends in.  prog: [length, node, node, ...] in postfix
// order; node record: [0] value rgb (image: max anisotropy) | type, [1] int bits: filter, mapping, image
// address mode, image first level, [2] uv scale.xy, offset.xy, [3..5] spherical world -> texture rows,
// [6] int bits: image level count, is_float.  Float textures use .x.
// fu, fv, fp: the uv and position the mappings read (the fragment's own, or the bump map's offset probes)
__device__ __noinline__ float3 evalTextureAt(const DeviceScene& sc, unsigned int progOffset, const TexFrag& f, float fu,
    float fv, float3 fp) {
    float3 stack[kTexStack];
    int sp = 0;
    const unsigned int* prog = sc.texProg + progOffset;
    const unsigned int len = __ldg(prog);
    for (unsigned int k = 1; k <= len; ++k) {
        const float4* node = sc.texNodes + kTexNodeVec4 * (size_t)__ldg(prog + k);
        const float4 head = __ldg(node);
        const int type = __float_as_int(head.w);
        if (type == GB_TEX_CONSTANT) {
            stack[sp++] = make3(head.x, head.y, head.z);
        } else if (type == GB_TEX_SCALE) { // mScale->lookup(f) * mTexture->lookup(f): children = texture, scale
            const float scale = stack[--sp].x;
            stack[sp - 1] = stack[sp - 1] * scale;
        } else { // checkerboard (two children on the stack) or image: both start from the texture mapping
            const int4 opt = __ldg(reinterpret_cast<const int4*>(node + 1));
            float s, t, dsdx, dtdx, dsdy, dtdy;
            if (opt.y == GB_MAPPING_SPHERICAL) { // SphericalMapping::map
                const float4 r0 = __ldg(node + 3), r1 = __ldg(node + 4), r2 = __ldg(node + 5);
                pointToST(r0, r1, r2, fp, &s, &t);
                float sdx, tdx, sdy, tdy;
                pointToST(r0, r1, r2, fp + f.dpdx, &sdx, &tdx);
                pointToST(r0, r1, r2, fp + f.dpdy, &sdy, &tdy);
                dsdx = sdx - s;
                if (dsdx > 0.5f) dsdx -= 1.0f; else if (dsdx < -0.5f) dsdx += 1.0f;
                dsdy = sdy - s;
                if (dsdy > 0.5f) dsdy -= 1.0f; else if (dsdy < -0.5f) dsdy += 1.0f;
                dtdx = tdx - t;
                dtdy = tdy - t;
            } else { // UVMapping::map
                const float4 m = __ldg(node + 2);
                s = m.x * fu + m.z;
                t = m.y * fv + m.w;
                dsdx = m.x * f.dudx; dtdx = m.y * f.dvdx;
                dsdy = m.x * f.dudy; dtdy = m.y * f.dvdy;
            }
            if (type == GB_TEX_IMAGE) { // ImageTexture<T>::lookup, GoblinTexture.cpp:447-452
                const int4 img = __ldg(reinterpret_cast<const int4*>(node + 6)); // level count, is_float, -, -
                TexImage im{sc.texLevels + opt.w, sc.imageTexels, img.x, opt.z, img.y != 0};
                stack[sp++] = imgLookup(im, opt.x, head.x, s, t, dsdx, dtdx, dsdy, dtdy);
                continue;
            }
            // CheckboardTexture<T>::lookup, GoblinTexture.cpp:377-416: children = texture1, texture2
            const float3 T2 = stack[--sp];
            const float3 T1 = stack[sp - 1];
            const bool even = (floorInt(s) + floorInt(t)) % 2 == 0;
            float3 r = even ? T1 : T2;
            if (opt.x) {
                const float ds = fmaxf(fabsf(dsdx), fabsf(dsdy));
                const float dt = fmaxf(fabsf(dtdx), fabsf(dtdy));
                const float s0 = s - ds, s1 = s + ds, t0 = t - dt, t1 = t + dt;
                if (!(floorInt(s0) == floorInt(s1) && floorInt(t0) == floorInt(t1))) {
                    const float sRatio = (integrateChecker(s1) - integrateChecker(s0)) / (2.0f * ds);
                    const float tRatio = (integrateChecker(t1) - integrateChecker(t0)) / (2.0f * dt);
                    float tex2Area = sRatio + tRatio - 2.0f * sRatio * tRatio;
                    if (ds > 1.0f || dt > 1.0f) tex2Area = 0.5f;
                    r = (1.0f - tex2Area) * T1 + tex2Area * T2;
                }
            }
            stack[sp - 1] = r;
        }
    }
    return stack[0];
}
__device__ __forceinline__ float3 evalTexture(const DeviceScene& sc, unsigned int progOffset, const TexFrag& f) {
    return evalTextureAt(sc, progOffset, f, f.u, f.v, f.p);
}

// Material::perturb -> BumpShaders::evaluate (GoblinMaterial.cpp:221-281): height-field bump mapping by
// forward differences of the bump texture along dpdu / dpdv, then a tangent-space normal map.  Rewrites the
// fragment's normal and tangents; the lookups see a fresh Fragment (no differentials).
__device__ __noinline__ void perturbFragment(const DeviceScene& sc, unsigned int bumpProg, unsigned int normalProg, TexFrag* tf) {
    if (bumpProg) {
        const float3 p = tf->p, n = tf->n;
        const float bumpD = evalTexture(sc, bumpProg, *tf).x;
        const float du = 0.002f, dv = 0.002f;
        // the two probes: position and uv moved along dpdu / dpdv (Fragment copies in the reference)
        const float bumpDdu = evalTextureAt(sc, bumpProg, *tf, tf->u + du, tf->v + 0.0f, p + du * tf->dpdu).x;
        const float3 bumpDPDU = tf->dpdu + (bumpDdu - bumpD) / du * n;
        const float bumpDdv = evalTextureAt(sc, bumpProg, *tf, tf->u + 0.0f, tf->v + dv, p + dv * tf->dpdv).x;
        const float3 bumpDPDV = tf->dpdv + (bumpDdv - bumpD) / dv * n;
        float3 bumpN = normalize3(cross3(bumpDPDU, bumpDPDV));
        if (dot3(bumpN, n) < 0.0f) bumpN = bumpN * -1.0f;
        tf->n = bumpN;
        tf->dpdu = bumpDPDU;
        tf->dpdv = bumpDPDV;
    }
    if (normalProg) {
        const float3 c = evalTexture(sc, normalProg, *tf);
        const float3 nShade = 2.0f * c - make3(1.0f, 1.0f, 1.0f);
        Frag tmp;
        tmp.n = tf->n;
        tmp.dpdu = tf->dpdu;
        float3 nWorld = normalize3(shadeToWorld(makeFrame(tmp), nShade));
        if (dot3(nWorld, tf->n) < 0.0f) nWorld = nWorld * -1.0f;
        tf->n = nWorld;
    }
}

// Looks up the textured slots of material `material` at this hit and writes them over the constants
// in m; applies the material's bump / normal maps to the fragment first, like Scene::intersect does.
// primary: the hit of a camera ray (the only rays that carry differentials,
// GoblinPathtracer.cpp:77 + RayDifferential(p, wi, epsilon) afterwards).
__device__ __noinline__ void applyTextures(const DeviceScene& sc, int material, int matType, const HitRec& hit, float3 o,
    float3 d, const Frag& fr, bool primary, float imageX, float imageY, float lensU1, float lensU2, DeviceMaterial* m,
    float3* nOut, float3* dpduOut) {
    const int4 slots = __ldg(sc.matTex + 2 * (size_t)material);      // programs: kd, kt, exponent, bump; 0 = none
    const int4 slots2 = __ldg(sc.matTex + 2 * (size_t)material + 1); // normal map, -, -, -
    if ((slots.x | slots.y | slots.z | slots.w | slots2.x) == 0) return;
    TexFrag tf;
    texFragment(sc, hit, o, d, fr, &tf);
    if (slots.w | slots2.x) {
        perturbFragment(sc, (unsigned int)slots.w, (unsigned int)slots2.x, &tf);
        *nOut = tf.n;
        *dpduOut = tf.dpdu;
    }
    if (primary) texDifferentials(sc, imageX, imageY, lensU1, lensU2, &tf);
    if (slots.x) {
        const float3 c = evalTexture(sc, (unsigned int)slots.x, tf);
        m->kdType = make_float4(c.x, c.y, c.z, m->kdType.w);
    }
    if (slots.y) { // transparent Kt
        const float3 c = evalTexture(sc, (unsigned int)slots.y, tf);
        m->ktEta = make_float4(c.x, c.y, c.z, m->ktEta.w);
    }
    if (slots.z && matType == GB_MAT_BLINN) { // blinn packing: ktEta = (k, exponent, fresnel, eta)
        m->ktEta.y = evalTexture(sc, (unsigned int)slots.z, tf).x;
    }
}

// The perturbed frame alone (ambient occlusion only needs the hit frame)
__device__ __noinline__ void perturbOnly(const DeviceScene& sc, const HitRec& hit, float3 o, float3 d, const Frag& fr,
    float3* nOut, float3* dpduOut) {
    const int4 slots = __ldg(sc.matTex + 2 * (size_t)fr.material);
    const int4 slots2 = __ldg(sc.matTex + 2 * (size_t)fr.material + 1);
    if ((slots.w | slots2.x) == 0) return;
    TexFrag tf;
    texFragment(sc, hit, o, d, fr, &tf);
    perturbFragment(sc, (unsigned int)slots.w, (unsigned int)slots2.x, &tf);
    *nOut = tf.n;
    *dpduOut = tf.dpdu;
}

} // namespace gb
