// The reference's BVH build (src/GoblinBVH.cpp:34-151), restated so that the
// flattened node array and the primitive order are bit-identical:
// pre-order node numbering, implicit left child = node + 1, equal_count split
// with libstdc++'s std::nth_element on the box centre along the longest axis of
// the centre bounds, leaf when one primitive is left or when all centres
// coincide on that axis.  There is no SAH in the reference.
//
// Two more split methods share the same node format, so every kernel walks
// them unchanged:
//   * Middle: the reference's other, unused method (src/GoblinBVH.cpp:124-134):
//     std::partition about the centre of the centroid bounds, falling through
//     to equal_count when that leaves one side empty;
//   * Sah (NOT a parity mode): binned surface-area
//     heuristic over all three axes, one primitive per leaf, tree depth bounded
//     to ceil(log2 n) + kSahDepthSlack by falling back to the median wherever a
//     SAH plane would leave a side too large for the levels that remain -- the
//     traversal stack lives in shared memory and is sized by the tree depth.
#pragma once
#include <cstdint>
#include <vector>

#include "goblin_b200.h"
#include "host_math.h"

namespace gb {

struct BuiltBVH {
    std::vector<gb_bvh_node> nodes;
    std::vector<uint32_t> order; // leaf slot -> input primitive index
    BBox bound;                  // union of the input boxes (BVH::mAABB)
    int maxDepth = 0;            // deepest node level (root = 0)
};

enum class BvhMethod { EqualCount = 0, Middle = 1, Sah = 2 };
constexpr int kSahDepthSlack = 3;

// boxes: one BBox per primitive, in input order.
void buildBVH(const std::vector<BBox>& boxes, BuiltBVH* out, BvhMethod method = BvhMethod::EqualCount);

} // namespace gb
