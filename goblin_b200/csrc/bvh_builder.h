// The reference's BVH build (src/GoblinBVH.cpp:34-151), restated so that the
// flattened node array and the primitive order are bit-identical:
// pre-order node numbering, implicit left child = node + 1, equal_count split
// with libstdc++'s std::nth_element on the box centre along the longest axis of
// the centre bounds, leaf when one primitive is left or when all centres
// coincide on that axis.  There is no SAH in the reference.
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

// boxes: one BBox per primitive, in input order.
void buildBVH(const std::vector<BBox>& boxes, BuiltBVH* out);

} // namespace gb
