// Wavefront OBJ reader that reproduces the vertex order, face order and quad
// split of the reference's PolygonMesh::loadObjMesh
// (src/GoblinPolygonMesh.cpp:58-262), so that triangle index i here is
// triangle index i there and the per-mesh BVH comes out bit-identical.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "host_math.h"

namespace gb {

struct MeshData {
    std::vector<float> pos;     // 3 per vertex
    std::vector<float> nrm;     // 3 per vertex (zero when the file has no vn)
    std::vector<float> uv;      // 2 per vertex (zero when the file has no vt)
    std::vector<uint32_t> idx;  // 3 per triangle
    bool hasNormal = false;
    bool hasUv = false;
    BBox bound;                 // over referenced vertices only
    float area = 0.0f;
    size_t numTris() const { return idx.size() / 3; }
    size_t numVerts() const { return pos.size() / 3; }
};

// Returns false (and leaves an empty mesh, as the reference does) when the
// file cannot be opened or has a syntax error; *error says why.
bool loadObjMesh(const std::string& path, MeshData* mesh, std::string* error);

} // namespace gb
