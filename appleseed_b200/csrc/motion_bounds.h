//
// motion_bounds.h -- what the assembly-tree build needs from an ANIMATED assembly instance
// (a TransformSequence with two or more keys): the interpolator segments the traversal evaluates
// and the bounding box of the instance over its whole motion.
//
// Reference (src/appleseed/): AssemblyTree::collect_assembly_instances,
// renderer/kernel/intersection/assemblytree.cpp:124-150 -> TransformSequence::prepare,
// renderer/utility/transformsequence.cpp:203-246 -> TransformInterpolator::set_transforms,
// foundation/math/transform.h:641-655 -> Matrix::decompose, foundation/math/matrix.h:1398-1516,
// 2129-2137; and TransformSequence::to_parent(bbox), transformsequence.h:212-236 ->
// compute_motion_segment_bbox, transformsequence.cpp:509-616 (Gruenschloss, "Motion blur", p. 11:
// the extrema of every box corner's trajectory under an interpolated rotation with linearly
// interpolated scaling, found with Newton's method, foundation/math/root.h:136-219).
//
// Same double arithmetic in the same order as the reference (host code on both sides: same libm),
// so the boxes -- and with them the assembly tree -- are the reference's bit for bit
// (tests/test_animated_build.py against oracle/_ref, which links the reference's own
// transformsequence.cpp).
//
#pragma once

#include "../../include/asgpu.h"

namespace asgpu
{

// Segment between two keys: scale, rotation quaternion (s, x, y, z) and translation at both ends,
// as TransformInterpolator keeps them.  Returns false when a rotation is not a unit quaternion
// within 1e-6 (set_transforms' return value: the reference then keeps the box of the first key only).
bool make_transform_segment(const double from_local_to_parent[16], const double to_local_to_parent[16], asgpu_transform_segment& segment);

// TransformSequence::to_parent<float>(bbox) for key_count >= 2 keys (ascending times): bounding box,
// in floats like the reference's GAABB3, of `bbox` (lo[3], hi[3]) moved through all the segments.
void motion_bounds(const double* local_to_parent, const double* parent_to_local, uint32_t key_count,
                   const float bbox_lo[3], const float bbox_hi[3], float out_lo[3], float out_hi[3]);

}   // namespace asgpu
