// Device-side scene layout (HBM resident, read-only during rendering).
//
// Everything a traversal step touches is a 16-byte-aligned float4 / int4 array
// read with 128-bit __ldg loads:
//   * BVH nodes stay bit-identical to the reference's 32-byte CompactBVHNode
//     (src/GoblinBVH.h:8-30); the traversal reads records derived from them on the
//     device at upload -- by default a 128-byte 4-wide node per interior node of
//     even depth (four grandchild boxes + references, wide_node.h), for the exact /
//     counting walk a 64-byte "pair node" per interior node (both children) -- and
//     the original nodes only for each level's root and the rare multi-primitive leaf;
//   * instances and triangles are stored in BVH leaf order, so a leaf's
//     primitives are contiguous and no order[] indirection is paid per test;
//   * a triangle test record is p0, e1 = p1 - p0, e2 = p2 - p0 (the values the
//     reference recomputes per test, src/GoblinTriangle.cpp:51-54) plus the
//     face index: 10 words in 48 bytes.  The three vertex normals and uvs of a
//     triangle live in a parallel 64-byte record touched only by shading (one
//     contiguous fetch per accepted hit instead of an index + vertex gather).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "goblin_b200.h"

namespace gb {

constexpr int kMaxLights = 1 << 20;

// Traversal records are fetched with 256-bit loads (LDG.E.256, sm_100+): a lane's gather of one 32-byte sector
// is then ONE pass through the L1 data pipe instead of two 128-bit ones.  The traversal kernels are bound by
// exactly that pipe (ncu: l1tex__data_pipe_lsu_wavefronts at 72 - 79 % of peak on incoherent rays with 128-bit
// loads: every lane of a warp reads a different line, so every lane is a wavefront of its own).  Triangle
// records are padded from 48 to 64 bytes so that both halves are 32-byte aligned.  GB_LD256=0 restores the
// 128-bit loads and the 48-byte records (A/B builds).
#ifndef GB_LD256
#define GB_LD256 1
#endif
constexpr int kTriRecVec4 = GB_LD256 ? 4 : 3; // float4 per triangle test record

struct DeviceLight { // 128 bytes
    float4 colorType;   // rgb, __int_as_float(type)
    float4 posRadius;   // position xyz, radius
    float4 dirCos;      // direction xyz, cosThetaMax | mesh emitter: uint bits (first light triangle, count, has normals, first CDF entry)
    float4 misc;        // cosFalloffStart, area, __int_as_float(geomKind), __int_as_float(instanceSlot)
    float4 toWorld[3];  // light's own Transform (area lights)
    float4 toObject[3];
};

struct DeviceMaterial { // 32 bytes
    float4 kdType;      // kd / kr rgb, __int_as_float(type)
    float4 ktEta;       // kt rgb, eta
};

struct TraceTuning { // warp scheduling of the traversal kernels (traverse.cuh)
    unsigned int refillBelow, leafBatch, levelBatch, moveFloor;
};

struct DeviceScene {
    TraceTuning tune;
    // ---- traversal data
    const float4* topNodes;     // 2 per node
    uint32_t nTopNodes;
    uint32_t nInstances;
    const float4* topPairs;     // 4 per interior node of the top-level BVH (pair nodes, traverse.cuh)
    const float4* topWide;      // 8 per wide root of the top-level BVH (4-wide nodes, wide_node.h)
    uint32_t topRootRef;        // pair index 0, or a leaf reference when the root is a leaf
    const float4* instToObject; // 3 per instance slot (BVH leaf order)
    const int4* instInfo;       // per slot: kind, node base (in modelNodes), tri base (in triRec), __float_as_int(radius)
    const int4* instInfo2;      // per slot: root reference, pair base (in modelPairs), node count (0 = empty mesh), wide base (in modelWide)
    const float4* modelNodes;   // 2 per node, all models concatenated
    const float4* modelPairs;   // 4 per interior node, all models concatenated
    const float4* modelWide;    // 8 per wide root, all models concatenated
    const float4* triRec;       // kTriRecVec4 per triangle slot (BVH leaf order, all models concatenated)
    // ---- shading data
    const float4* instToWorld;  // 3 per instance slot
    const int4* instShade;      // per slot: original instance index, model index, material, area light (-1)
    const int4* modelShade;     // per model: vert base, tri base (face order), has_normal | has_uv << 1, unused
    const float4* triShade;     // 4 per triangle slot (leaf order): the 3 vertex normals and 3 uvs, 64 bytes
    const DeviceMaterial* materials;
    // procedural textures (texture.cuh); all three null when no material slot is textured
    const int4* matTex;         // 2 per material: program offsets in texProg of (kd, kt, exponent, bump map), (normal map, -, -, -); 0 = none
    const unsigned int* texProg; // postfix programs: [length, node index ...]; entry 0 is unused
    const float4* texNodes;     // 7 per texture: value | type, options, uv mapping, 3 rows world -> texture, image info
    const int4* matMask;        // Mask materials (mask.cuh), 2 per material: (is mask, alpha program, colour program, 0),
                                // float bits (alpha, transparent colour rgb); null when the scene has none
    const int4* texLevels;      // image textures: (width, height, first texel in imageTexels, 0) per pyramid level
    const DeviceLight* lights;
    const float* lightPower;    // CDF1D::mFunction
    const float* lightCdf;      // CDF1D::mCDF (nLights + 1)
    const float4* lightTris;    // mesh emitters, 6 per face in FACE order: p0|area, p1, p2, n0, n1, n2
    const float* lightTriCdf;   // mesh emitters: area CDFs
    const float4* imageTexels;  // image based lights: level 0 of the radiance maps, RGBA
    const float* lightDist;     // image based lights: CDF2D tables (layout in goblin_b200.h)
    uint32_t nLights;
    float lightIntegral;        // CDF1D::mIntegral
    uint32_t hasAreaLight;
    uint32_t hasEnvLight;       // some light is image based: missed rays pick up its radiance (k_miss)
    // ---- camera / film
    gb_camera camera;
    int xres, yres;
    int xstart, xcount, ystart, ycount;
    int sx0, sx1, sy0, sy1;
    float invXRes, invYRes;
    float filterWidthX, filterWidthY;
    const float* filterTable;   // 256 floats
};

} // namespace gb
