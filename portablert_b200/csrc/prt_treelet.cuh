// prt_treelet.cuh -- opt-in SAH optimisation of the radix tree by treelet restructuring
// (Karras & Aila, "Fast Parallel Construction of High-Quality Bounding Volume Hierarchies", HPG
// 2013), shared by the sm_100a build kernel (build.cu: k_treelet) and by the host-side logic
// emulator of the CPU-only tests (tests/emu; not a product path).
//
// The reference builds its tree with a binned SAH sweep (include/portableRT/bvh.hpp:58-129); the
// LBVH of build.cu is built in a fraction of a millisecond but knows nothing about surface areas.
// This pass closes part of that quality gap without giving up the parallel build: every internal
// node with at least TREELET_N leaves below it becomes, bottom-up, the root of a treelet of
// TREELET_N "leaves" (subtrees), grown by always expanding the leaf with the largest surface area;
// the topology of the treelet's TREELET_N-1 internal nodes is then replaced by the one that
// minimises the sum of their surface areas -- exact dynamic programming over all 2^7 subsets of
// the leaves -- and written back into the same node slots (the treelet root keeps its index, so
// its parent needs no update).  Every triangle stays a leaf of its own, so with unit costs the
// SAH cost of the tree is (sum of the internal nodes' areas) / (root area) + const: exactly what
// the DP minimises.  Results of nearest_hits do not depend on the tree (prt_traverse.cuh), only
// the number of boxes a ray looks at does.
#pragma once

#include "prt_math.cuh"

namespace prt {

constexpr int TREELET_N = 7;
constexpr int TREELET_SETS = 1 << TREELET_N;

#if defined(__CUDA_ARCH__)
// nodes below the current one were written by other threads: read them through L2 (L1 is not
// coherent), after the acquire fence of the bottom-up protocol
__device__ __forceinline__ Node treelet_load(const Node *nodes, int32_t i) {
	Node nd;
	const float4 *p = reinterpret_cast<const float4 *>(nodes + i);
	float4 *q = reinterpret_cast<float4 *>(&nd);
	q[0] = __ldcg(p);
	q[1] = __ldcg(p + 1);
	q[2] = __ldcg(p + 2);
	q[3] = __ldcg(p + 3);
	return nd;
}
__device__ __forceinline__ void treelet_store(Node *nodes, int32_t i, const Node &nd) {
	const float4 *q = reinterpret_cast<const float4 *>(&nd);
	float4 *p = reinterpret_cast<float4 *>(nodes + i);
	p[0] = q[0];
	p[1] = q[1];
	p[2] = q[2];
	p[3] = q[3];
}
__device__ __forceinline__ int32_t treelet_load_i32(const int32_t *p) { return __ldcg(p); }
#else
inline Node treelet_load(const Node *nodes, int32_t i) { return nodes[i]; }
inline void treelet_store(Node *nodes, int32_t i, const Node &nd) { nodes[i] = nd; }
inline int32_t treelet_load_i32(const int32_t *p) { return *p; }
#endif

// half the surface area; un-fused so that host and device agree bit for bit
PRT_HD float box_half_area(const Box &b) {
	const float dx = fsub(b.hi[0], b.lo[0]), dy = fsub(b.hi[1], b.lo[1]), dz = fsub(b.hi[2], b.lo[2]);
	return fadd(fadd(fmul(dx, dy), fmul(dy, dz)), fmul(dz, dx));
}

PRT_HD Box child_box(const Node &nd, int side) {
	Box b;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = side ? nd.lo1[a] : nd.lo0[a];
		b.hi[a] = side ? nd.hi1[a] : nd.hi0[a];
	}
	return b;
}

PRT_HD void set_child(Node &nd, int side, int32_t ref, const Box &b) {
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		(side ? nd.lo1 : nd.lo0)[a] = b.lo[a];
		(side ? nd.hi1 : nd.hi0)[a] = b.hi[a];
	}
	(side ? nd.child1 : nd.child0) = ref;
}

// Optimises the treelet rooted at internal node x (which must have >= TREELET_N leaves below it).
// depth[i] = height of internal node i's subtree (a node over two triangles has height 1): read
// for the subtrees hanging off the treelet, written for every node slot of the treelet, x included.
// Returns true when the topology was replaced.
PRT_HD bool treelet_optimise(Node *nodes, int32_t x, int32_t *depth) {
	int32_t leaf_ref[TREELET_N];
	Box leaf_box[TREELET_N];
	float leaf_area[TREELET_N];
	int32_t slot[TREELET_N - 1]; // node indices of the treelet's internal nodes, slot[0] = x
	int n = 2, ns = 1;
	float old_cost = 0.0f; // areas of the expanded nodes (the root's is added below)

	const Node root = treelet_load(nodes, x);
	slot[0] = x;
	leaf_ref[0] = root.child0;
	leaf_ref[1] = root.child1;
	leaf_box[0] = child_box(root, 0);
	leaf_box[1] = child_box(root, 1);
	leaf_area[0] = box_half_area(leaf_box[0]);
	leaf_area[1] = box_half_area(leaf_box[1]);
	while (n < TREELET_N) {
		int best = -1;
		float best_area = -1.0f;
		for (int k = 0; k < n; ++k)
			if (leaf_ref[k] >= 0 && (best < 0 || leaf_area[k] > best_area)) {
				best = k;
				best_area = leaf_area[k];
			}
		if (best < 0)
			break; // (cannot happen with >= TREELET_N leaves below x)
		const int32_t e = leaf_ref[best];
		const Node nd = treelet_load(nodes, e);
		slot[ns++] = e;
		old_cost = fadd(old_cost, leaf_area[best]);
		leaf_ref[best] = nd.child0;
		leaf_box[best] = child_box(nd, 0);
		leaf_area[best] = box_half_area(leaf_box[best]);
		leaf_ref[n] = nd.child1;
		leaf_box[n] = child_box(nd, 1);
		leaf_area[n] = box_half_area(leaf_box[n]);
		++n;
	}
	int32_t leaf_depth[TREELET_N];
	for (int k = 0; k < n; ++k)
		leaf_depth[k] = leaf_ref[k] < 0 ? 0 : treelet_load_i32(depth + leaf_ref[k]);

	// ---- dynamic programming over the subsets of the leaves
	const int full = (1 << n) - 1;
	float area[TREELET_SETS], copt[TREELET_SETS];
	uint8_t part[TREELET_SETS];
	for (int s = 1; s <= full; ++s) {
		Box b;
		bool first = true;
		for (int k = 0; k < n; ++k)
			if (s & (1 << k)) {
				b = first ? leaf_box[k] : box_union(b, leaf_box[k]);
				first = false;
			}
		area[s] = box_half_area(b);
	}
	for (int s = 1; s <= full; ++s) {
		if ((s & (s - 1)) == 0) { // a single leaf: nothing to arrange
			copt[s] = 0.0f;
			part[s] = 0;
			continue;
		}
		// all ways to split s in two non-empty halves; the half holding s's lowest bit is the complement
		const int delta = (s - 1) & s;
		int p = (-delta) & s;
		float best = INFINITY;
		int bp = p;
		do {
			const float c = fadd(copt[p], copt[s ^ p]);
			if (c < best) {
				best = c;
				bp = p;
			}
			p = (p - delta) & s;
		} while (p != 0);
		copt[s] = fadd(area[s], best);
		part[s] = (uint8_t)bp;
	}
	old_cost = fadd(old_cost, area[full]);

	if (!(copt[full] < old_cost)) {
		// keep the topology; x's height follows from its two children
		const int32_t d0 = root.child0 < 0 ? 0 : treelet_load_i32(depth + root.child0);
		const int32_t d1 = root.child1 < 0 ? 0 : treelet_load_i32(depth + root.child1);
		depth[x] = 1 + (d0 > d1 ? d0 : d1);
		return false;
	}

	// ---- write the optimal topology back into the same node slots (parents before children)
	uint8_t todo_set[TREELET_N - 1];
	int8_t kid[TREELET_N - 1][2]; // >= 0: slot number, < 0: ~leaf number
	todo_set[0] = (uint8_t)full;
	int used = 1;
	for (int k = 0; k < used; ++k) {
		const int s = todo_set[k];
		const int half[2] = {part[s], s ^ part[s]};
		Node nd;
		nd.pad0 = nd.pad1 = 0;
		for (int side = 0; side < 2; ++side) {
			const int h = half[side];
			Box b;
			bool first = true;
			int only = -1;
			for (int j = 0; j < n; ++j)
				if (h & (1 << j)) {
					b = first ? leaf_box[j] : box_union(b, leaf_box[j]);
					first = false;
					only = j;
				}
			if ((h & (h - 1)) == 0) {
				set_child(nd, side, leaf_ref[only], b);
				kid[k][side] = (int8_t)~only;
			} else {
				todo_set[used] = (uint8_t)h;
				set_child(nd, side, slot[used], b);
				kid[k][side] = (int8_t)used;
				++used;
			}
		}
		treelet_store(nodes, slot[k], nd);
	}
	int32_t slot_depth[TREELET_N - 1];
	for (int k = used - 1; k >= 0; --k) {
		int32_t d[2];
		for (int side = 0; side < 2; ++side)
			d[side] = kid[k][side] < 0 ? leaf_depth[~kid[k][side]] : slot_depth[kid[k][side]];
		slot_depth[k] = 1 + (d[0] > d[1] ? d[0] : d[1]);
		depth[slot[k]] = slot_depth[k];
	}
	return true;
}

} // namespace prt
