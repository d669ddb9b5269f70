#include "bvh_builder.h"

#include <algorithm>
#include <cstring>

namespace gb {
namespace {

struct BuildItem {
    float center[3];
    uint32_t index;
};

struct CenterLess {
    int dim;
    bool operator()(const BuildItem& a, const BuildItem& b) const { return a.center[dim] < b.center[dim]; }
};

struct Pending {
    uint32_t start, end;
    uint32_t parent; // node whose second child this range is, or ~0u
    int depth;
};

} // namespace

void buildBVH(const std::vector<BBox>& boxes, BuiltBVH* out) {
    out->nodes.clear();
    out->order.clear();
    out->bound = BBox();
    out->maxDepth = 0;
    const uint32_t n = (uint32_t)boxes.size();
    for (const BBox& b : boxes) out->bound.expand(b);
    if (n == 0) return;

    std::vector<BuildItem> items(n);
    for (uint32_t i = 0; i < n; ++i) {
        Vec3 c = 0.5f * (boxes[i].pMin + boxes[i].pMax); // BVHPrimitiveInfo::center
        items[i] = BuildItem{{c.x, c.y, c.z}, i};
    }
    out->nodes.reserve(2 * (size_t)n - 1);

    // Depth-first, left subtree before right subtree, exactly the order the
    // reference's recursion appends nodes in.
    std::vector<Pending> todo;
    todo.push_back(Pending{0, n, ~0u, 0});
    while (!todo.empty()) {
        Pending r = todo.back();
        todo.pop_back();
        const uint32_t nodeIndex = (uint32_t)out->nodes.size();
        if (r.parent != ~0u) out->nodes[r.parent].offset = nodeIndex;
        gb_bvh_node node;
        std::memset(&node, 0, sizeof node); // value-initialised in the reference
        out->nodes.push_back(node);
        out->maxDepth = std::max(out->maxDepth, r.depth);

        BBox bbox;
        for (uint32_t i = r.start; i < r.end; ++i) bbox.expand(boxes[items[i].index]);
        const uint32_t count = r.end - r.start;
        bool leaf = count == 1;
        int dim = 0;
        if (!leaf) {
            BBox centers;
            for (uint32_t i = r.start; i < r.end; ++i) {
                centers.expand(Vec3(items[i].center[0], items[i].center[1], items[i].center[2]));
            }
            dim = centers.longestAxis();
            leaf = centers.pMin[dim] == centers.pMax[dim];
        }
        gb_bvh_node& nd = out->nodes[nodeIndex];
        nd.bmin[0] = bbox.pMin.x; nd.bmin[1] = bbox.pMin.y; nd.bmin[2] = bbox.pMin.z;
        nd.bmax[0] = bbox.pMax.x; nd.bmax[1] = bbox.pMax.y; nd.bmax[2] = bbox.pMax.z;
        if (leaf) {
            // ordered primitives are appended in range order, so the first
            // primitive slot of a leaf is the start of its range
            nd.offset = r.start;
            nd.nprims = (uint8_t)count; // uint8 in the reference: wraps past 255
        } else {
            const uint32_t mid = (r.start + r.end) / 2;
            std::nth_element(items.begin() + r.start, items.begin() + mid, items.begin() + r.end,
                CenterLess{dim});
            nd.axis = (uint8_t)dim;
            nd.nprims = 0;
            todo.push_back(Pending{mid, r.end, nodeIndex, r.depth + 1});
            todo.push_back(Pending{r.start, mid, ~0u, r.depth + 1});
        }
    }
    out->order.resize(n);
    for (uint32_t i = 0; i < n; ++i) out->order[i] = items[i].index;
}

} // namespace gb
