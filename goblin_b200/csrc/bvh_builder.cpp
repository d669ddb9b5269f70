#include "bvh_builder.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <future>
#include <thread>

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

struct CenterBelow { // MidComparator, src/GoblinBVH.cpp:16-23
    int dim;
    float midPoint;
    bool operator()(const BuildItem& a) const { return a.center[dim] < midPoint; }
};

inline float halfArea(const BBox& b) {
    const Vec3 d = b.pMax - b.pMin;
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

// levels a subtree over n primitives needs at least (one primitive per leaf)
inline int minLevels(uint32_t n) {
    int l = 0;
    while ((1ull << l) < n) ++l;
    return l;
}

struct Builder {
    const std::vector<BBox>& boxes;
    std::vector<BuildItem>& items;
    BvhMethod method = BvhMethod::EqualCount;
    int depthLimit = 0; // Sah: deepest level a node may sit on

    // Binned SAH over the three axes of the centroid bounds.  Returns false when no plane
    // separates the centroids into two sides that both still fit under the depth limit.
    bool sahPlane(uint32_t start, uint32_t end, int depth, const BBox& centers, int* dimOut, float* planeOut) {
        constexpr int K = 32;
        const uint32_t n = end - start;
        const int levelsLeft = depthLimit - depth - 1; // levels available to each child's subtree
        const uint64_t sideCap = levelsLeft >= 31 ? 0xffffffffull : (1ull << std::max(levelsLeft, 0));
        float best = INFINITY;
        bool have = false;
        for (int dim = 0; dim < 3; ++dim) {
            const float lo = centers.pMin[dim], hi = centers.pMax[dim];
            if (!(hi > lo)) continue;
            const float scale = (float)K / (hi - lo);
            BBox binBox[K];
            uint32_t binCount[K] = {};
            for (uint32_t i = start; i < end; ++i) {
                int b = (int)((items[i].center[dim] - lo) * scale);
                b = b < 0 ? 0 : (b >= K ? K - 1 : b);
                binBox[b].expand(boxes[items[i].index]);
                ++binCount[b];
            }
            float rightArea[K]; // area of bins [b, K); only read where that range is non-empty
            BBox acc;
            uint32_t nacc = 0;
            for (int b = K - 1; b > 0; --b) {
                if (binCount[b]) acc.expand(binBox[b]);
                nacc += binCount[b];
                rightArea[b] = nacc ? halfArea(acc) : 0.0f;
            }
            BBox left;
            uint32_t nl = 0;
            for (int b = 0; b < K - 1; ++b) {
                if (binCount[b]) left.expand(binBox[b]);
                nl += binCount[b];
                const uint32_t nr = n - nl;
                if (nl == 0 || nr == 0 || nl > sideCap || nr > sideCap) continue;
                const float cost = halfArea(left) * (float)nl + rightArea[b + 1] * (float)nr;
                if (cost < best) {
                    best = cost;
                    have = true;
                    *dimOut = dim;
                    // items with bin index <= b go left: the same float arithmetic decides the
                    // partition below
                    *planeOut = (float)(b + 1);
                }
            }
        }
        return have;
    }

    static void setBox(gb_bvh_node& nd, const BBox& b) {
        nd.bmin[0] = b.pMin.x; nd.bmin[1] = b.pMin.y; nd.bmin[2] = b.pMin.z;
        nd.bmax[0] = b.pMax.x; nd.bmax[1] = b.pMax.y; nd.bmax[2] = b.pMax.z;
    }
    static BBox getBox(const gb_bvh_node& nd) {
        BBox b;
        b.pMin = Vec3(nd.bmin[0], nd.bmin[1], nd.bmin[2]);
        b.pMax = Vec3(nd.bmax[0], nd.bmax[1], nd.bmax[2]);
        return b;
    }

    // Decide leaf / split for [start, end) exactly as buildLinearBVH does
    // (src/GoblinBVH.cpp:93-141): one primitive, or all centres equal along the longest axis of
    // the centre bounds -> leaf; else nth_element at the median of that axis.
    bool split(uint32_t start, uint32_t end, int depth, int* dimOut, uint32_t* midOut) {
        if (end - start == 1) return false;
        BBox centers;
        for (uint32_t i = start; i < end; ++i) {
            centers.expand(Vec3(items[i].center[0], items[i].center[1], items[i].center[2]));
        }
        const int dim = centers.longestAxis();
        if (centers.pMin[dim] == centers.pMax[dim]) return false;
        if (method == BvhMethod::Sah) {
            int sdim = 0;
            float plane = 0.0f;
            if (sahPlane(start, end, depth, centers, &sdim, &plane)) {
                const float lo = centers.pMin[sdim], scale = 32.0f / (centers.pMax[sdim] - lo);
                auto it = std::partition(items.begin() + start, items.begin() + end, [=](const BuildItem& a) {
                    int b = (int)((a.center[sdim] - lo) * scale);
                    b = b < 0 ? 0 : (b >= 32 ? 31 : b);
                    return (float)b < plane;
                });
                const uint32_t m = (uint32_t)(it - items.begin());
                if (m != start && m != end) {
                    *dimOut = sdim;
                    *midOut = m;
                    return true;
                }
            }
        } else if (method == BvhMethod::Middle) {
            const float midPoint = 0.5f * (centers.pMin[dim] + centers.pMax[dim]);
            auto it = std::partition(items.begin() + start, items.begin() + end, CenterBelow{dim, midPoint});
            const uint32_t m = (uint32_t)(it - items.begin());
            if (m != start && m != end) {
                *dimOut = dim;
                *midOut = m;
                return true;
            }
            // "can't split down further with middle method": the equal_count case follows
        }
        const uint32_t mid = (start + end) / 2;
        std::nth_element(items.begin() + start, items.begin() + mid, items.begin() + end, CenterLess{dim});
        *dimOut = dim;
        *midOut = mid;
        return true;
    }

    void makeLeaf(gb_bvh_node& nd, uint32_t start, uint32_t end) {
        BBox bbox;
        for (uint32_t i = start; i < end; ++i) bbox.expand(boxes[items[i].index]);
        setBox(nd, bbox);
        // ordered primitives are appended in range order, so the first primitive slot of a leaf
        // is the start of its range
        nd.offset = start;
        nd.nprims = (uint8_t)(end - start); // uint8 in the reference: wraps past 255
    }

    // Sequential subtree build, pre-order (left subtree before right subtree: the order the
    // reference's recursion appends nodes in), node indices local to `out`.  Boxes are united
    // bottom-up: min / max are exact, so this equals the reference's scan over the range.
    void buildSeq(uint32_t start, uint32_t end, int depth, std::vector<gb_bvh_node>& out, int* maxDepth) {
        struct Frame { uint32_t start, end, node; int depth; int stage; uint32_t mid; };
        std::vector<Frame> st;
        auto open = [&](uint32_t s, uint32_t e, int d) {
            gb_bvh_node nd;
            std::memset(&nd, 0, sizeof nd); // value-initialised in the reference
            out.push_back(nd);
            st.push_back(Frame{s, e, (uint32_t)out.size() - 1, d, 0, 0});
            *maxDepth = std::max(*maxDepth, d);
        };
        open(start, end, depth);
        while (!st.empty()) {
            Frame& f = st.back();
            if (f.stage == 0) {
                int dim = 0;
                uint32_t mid = 0;
                if (!split(f.start, f.end, f.depth, &dim, &mid)) {
                    makeLeaf(out[f.node], f.start, f.end);
                    st.pop_back();
                    continue;
                }
                out[f.node].axis = (uint8_t)dim;
                out[f.node].nprims = 0;
                f.mid = mid;
                f.stage = 1;
                const uint32_t s = f.start;
                const int d = f.depth + 1;
                open(s, mid, d); // left child = node + 1 (invalidates f)
            } else if (f.stage == 1) {
                f.stage = 2;
                out[f.node].offset = (uint32_t)out.size(); // second child
                const uint32_t m = f.mid, e = f.end;
                const int d = f.depth + 1;
                open(m, e, d);
            } else {
                BBox b = getBox(out[f.node + 1]);
                b.expand(getBox(out[out[f.node].offset]));
                setBox(out[f.node], b);
                st.pop_back();
            }
        }
    }

    // Parallel build of the top of the tree: both halves of a large range are independent once
    // nth_element has placed the median (the left subtree only permutes its own half), so they
    // build concurrently into their own vectors and are stitched in pre-order afterwards.
    void buildPar(uint32_t start, uint32_t end, int depth, int spawnLevels, std::vector<gb_bvh_node>& out,
        int* maxDepth) {
        int dim = 0;
        uint32_t mid = 0;
        if (spawnLevels <= 0 || end - start < (1u << 15)) {
            buildSeq(start, end, depth, out, maxDepth);
            return;
        }
        *maxDepth = std::max(*maxDepth, depth);
        gb_bvh_node nd;
        std::memset(&nd, 0, sizeof nd);
        if (!split(start, end, depth, &dim, &mid)) {
            makeLeaf(nd, start, end);
            out.push_back(nd);
            return;
        }
        std::vector<gb_bvh_node> left, right;
        int dl = 0, dr = 0;
        auto fut = std::async(std::launch::async, [&]() { buildPar(mid, end, depth + 1, spawnLevels - 1, right, &dr); });
        buildPar(start, mid, depth + 1, spawnLevels - 1, left, &dl);
        fut.get();
        *maxDepth = std::max(*maxDepth, std::max(dl, dr));
        nd.axis = (uint8_t)dim;
        nd.nprims = 0;
        BBox b = getBox(left[0]);
        b.expand(getBox(right[0]));
        setBox(nd, b);
        const uint32_t base = (uint32_t)out.size();
        nd.offset = base + 1 + (uint32_t)left.size();
        out.reserve(out.size() + 1 + left.size() + right.size());
        out.push_back(nd);
        // interior offsets are node indices local to the sub-vectors; leaf offsets are primitive
        // slots and already global
        const uint32_t lbase = base + 1, rbase = base + 1 + (uint32_t)left.size();
        for (gb_bvh_node& n : left) if (n.nprims == 0) n.offset += lbase;
        for (gb_bvh_node& n : right) if (n.nprims == 0) n.offset += rbase;
        out.insert(out.end(), left.begin(), left.end());
        out.insert(out.end(), right.begin(), right.end());
    }
};

} // namespace

void buildBVH(const std::vector<BBox>& boxes, BuiltBVH* out, BvhMethod method) {
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
    Builder b{boxes, items, method, minLevels(n) + kSahDepthSlack};
    // 2^levels concurrent subtrees: a few more than there are cores
    int levels = 0;
    for (unsigned c = std::max(1u, std::thread::hardware_concurrency()); (1u << levels) < 2 * c && levels < 8; ++levels) {}
    out->nodes.reserve(2 * (size_t)n - 1);
    b.buildPar(0, n, 0, n >= (1u << 16) ? levels : 0, out->nodes, &out->maxDepth);
    out->order.resize(n);
    for (uint32_t i = 0; i < n; ++i) out->order[i] = items[i].index;
}

} // namespace gb
