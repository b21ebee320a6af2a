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

// A treelet: TREELET_N subtrees ("leaves") hanging off TREELET_N-1 internal node slots.
struct Treelet {
	int32_t leaf_ref[TREELET_N];
	Box leaf_box[TREELET_N];
	float leaf_area[TREELET_N];
	int32_t leaf_depth[TREELET_N];
	int32_t slot[TREELET_N - 1]; // node indices of the treelet's internal nodes, slot[0] = root
	int32_t root_child[2];
	int n;
	float old_cost; // summed half areas of the expanded internal nodes (root excluded)
};

// Grow the treelet below x by always expanding the leaf with the largest surface area.
PRT_HD void treelet_form(const Node *nodes, int32_t x, const int32_t *depth, Treelet &t) {
	const Node root = treelet_load(nodes, x);
	int n = 2, ns = 1;
	float old_cost = 0.0f;
	t.slot[0] = x;
	t.root_child[0] = root.child0;
	t.root_child[1] = root.child1;
	t.leaf_ref[0] = root.child0;
	t.leaf_ref[1] = root.child1;
	t.leaf_box[0] = child_box(root, 0);
	t.leaf_box[1] = child_box(root, 1);
	t.leaf_area[0] = box_half_area(t.leaf_box[0]);
	t.leaf_area[1] = box_half_area(t.leaf_box[1]);
	while (n < TREELET_N) {
		int best = -1;
		float best_area = -1.0f;
		for (int k = 0; k < n; ++k)
			if (t.leaf_ref[k] >= 0 && (best < 0 || t.leaf_area[k] > best_area)) {
				best = k;
				best_area = t.leaf_area[k];
			}
		if (best < 0)
			break; // (cannot happen with >= TREELET_N leaves below x)
		const int32_t e = t.leaf_ref[best];
		const Node nd = treelet_load(nodes, e);
		t.slot[ns++] = e;
		old_cost = fadd(old_cost, t.leaf_area[best]);
		t.leaf_ref[best] = nd.child0;
		t.leaf_box[best] = child_box(nd, 0);
		t.leaf_area[best] = box_half_area(t.leaf_box[best]);
		t.leaf_ref[n] = nd.child1;
		t.leaf_box[n] = child_box(nd, 1);
		t.leaf_area[n] = box_half_area(t.leaf_box[n]);
		++n;
	}
	for (int k = 0; k < n; ++k)
		t.leaf_depth[k] = t.leaf_ref[k] < 0 ? 0 : treelet_load_i32(depth + t.leaf_ref[k]);
	t.n = n;
	t.old_cost = old_cost;
}

// Box of the union of the leaves in subset s, branch-free (the lanes of a warp evaluate different
// subsets in lockstep): starting from the empty box, min/max with +-inf leave a box unchanged.
PRT_HD Box treelet_subset_box(const Treelet &t, int s, Box b, int k0, int k1) {
	for (int k = k0; k < k1; ++k) {
		const bool take = (s >> k) & 1;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			b.lo[a] = smin(b.lo[a], take ? t.leaf_box[k].lo[a] : INFINITY);
			b.hi[a] = smax(b.hi[a], take ? t.leaf_box[k].hi[a] : -INFINITY);
		}
	}
	return b;
}

PRT_HD Box empty_box() {
	Box b;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = INFINITY;
		b.hi[a] = -INFINITY;
	}
	return b;
}

PRT_HD float treelet_subset_area(const Treelet &t, int s) {
	return box_half_area(treelet_subset_box(t, s, empty_box(), 0, t.n));
}

// Best split of subset s (>= 2 leaves) into two non-empty halves, given copt of all its proper
// subsets: minimum cost, and among equal costs the numerically smallest half p (the enumeration
// runs through the non-empty subsets of s-without-its-lowest-bit in increasing order).  Lanes of
// a warp may share one subset: lane `sub` of `nsub` takes every nsub-th candidate.
PRT_HD void treelet_best_split(const float *copt, int s, int sub, int nsub, float &best, int &bp) {
	const int delta = (s - 1) & s;
	int p = (-delta) & s;
	int i = 0;
	best = INFINITY;
	bp = 0xff; // none found (all candidates infinite or NaN): the caller falls back to the first
	do {
		if (nsub == 1 || (i % nsub) == sub) {
			const float c = fadd(copt[p], copt[s ^ p]);
			if (c < best) {
				best = c;
				bp = p;
			}
		}
		++i;
		p = (p - delta) & s;
	} while (p != 0);
}

// The same search restricted to the splits whose two lowest candidate bits equal `sub` (0..3): four
// lanes cover a subset without diverging; each runs through its candidates in increasing order.
PRT_HD void treelet_best_split_quarter(const float *copt, int s, int sub, float &best, int &bp) {
	const int delta = (s - 1) & s;
	const int b0 = delta & -delta, r0 = delta ^ b0;
	const int b1 = r0 & -r0, rest = r0 ^ b1;
	const int base = ((sub & 1) ? b0 : 0) | ((sub & 2) ? b1 : 0);
	best = INFINITY;
	bp = 0xff;
	int q = 0;
	do {
		const int p = base | q;
		if (p != 0) {
			const float c = fadd(copt[p], copt[s ^ p]);
			if (c < best) {
				best = c;
				bp = p;
			}
		}
		q = (q - rest) & rest;
	} while (q != 0);
}

PRT_HD int treelet_first_split(int s) {
	const int delta = (s - 1) & s;
	return (-delta) & s;
}

// Sequential dynamic programme over all subsets (host emulator; the device kernel spreads the same
// recurrence over the lanes of a warp, build.cu).
PRT_HD void treelet_dp(const Treelet &t, float *area, float *copt, uint8_t *part) {
	const int full = (1 << t.n) - 1;
	for (int s = 1; s <= full; ++s)
		area[s] = treelet_subset_area(t, s);
	for (int s = 1; s <= full; ++s) { // every proper subset of s is numerically smaller than s
		if ((s & (s - 1)) == 0) {     // a single leaf: nothing to arrange
			copt[s] = 0.0f;
			part[s] = 0;
			continue;
		}
		float best;
		int bp;
		treelet_best_split(copt, s, 0, 1, best, bp);
		copt[s] = fadd(area[s], best);
		part[s] = (uint8_t)(bp == 0xff ? treelet_first_split(s) : bp);
	}
}

// The optimal topology laid out over the treelet's node slots (parents before children: slot k's
// children take the next free slots).
struct TreeletPlan {
	int used;                          // node slots written
	uint8_t half[TREELET_N - 1][2];    // leaf subset below each child
	int32_t ref[TREELET_N - 1][2];     // child reference to store
	int32_t slot_depth[TREELET_N - 1]; // height of the subtree of each slot
};

// Decide whether the treelet's topology is replaced by the optimal one (only if that lowers the
// summed area) and lay the new topology out.  Sets depth[] of the root when the topology stays.
// `strict`: additionally refuse a topology that makes the subtree of x taller than it is.  By
// induction over the bottom-up order every subtree then stays at most as tall as in the radix tree,
// whose height is bounded by the key length -- the bound the traversal stack is sized for
// (prt_traverse.cuh: STACK_DEPTH).  It costs quality (big triangles want to sit high up, in a
// locally deeper tree), so the build first runs without it, measures the height of the result and
// only falls back to the strict rule if that exceeds the bound (build.cu: optimise_tree).
PRT_HD bool treelet_plan(const Treelet &t, const float *area, const float *copt, const uint8_t *part,
                         int32_t *depth, bool strict, TreeletPlan &pl) {
	const int full = (1 << t.n) - 1;
	const int32_t x = t.slot[0];
	const int32_t c0 = t.root_child[0], c1 = t.root_child[1];
	const int32_t d0 = c0 < 0 ? 0 : treelet_load_i32(depth + c0);
	const int32_t d1 = c1 < 0 ? 0 : treelet_load_i32(depth + c1);
	const int32_t old_depth = 1 + (d0 > d1 ? d0 : d1);
	depth[x] = old_depth;
	if (!(copt[full] < fadd(t.old_cost, area[full])))
		return false;
	int8_t kid[TREELET_N - 1][2]; // >= 0: slot number, < 0: ~leaf number
	pl.half[0][0] = part[full];
	pl.half[0][1] = (uint8_t)(full ^ part[full]);
	int used = 1;
	for (int k = 0; k < used; ++k) {
		for (int side = 0; side < 2; ++side) {
			const int h = pl.half[k][side];
			if ((h & (h - 1)) == 0) {
				int only = 0;
				while (!(h & (1 << only)))
					++only;
				kid[k][side] = (int8_t)~only;
				pl.ref[k][side] = t.leaf_ref[only];
			} else {
				pl.half[used][0] = part[h];
				pl.half[used][1] = (uint8_t)(h ^ part[h]);
				kid[k][side] = (int8_t)used;
				pl.ref[k][side] = t.slot[used];
				++used;
			}
		}
	}
	for (int k = used - 1; k >= 0; --k) {
		int32_t d[2];
		for (int side = 0; side < 2; ++side)
			d[side] = kid[k][side] < 0 ? t.leaf_depth[~kid[k][side]] : pl.slot_depth[kid[k][side]];
		pl.slot_depth[k] = 1 + (d[0] > d[1] ? d[0] : d[1]);
	}
	pl.used = used;
	return !(strict && pl.slot_depth[0] > old_depth);
}

// One child (box + reference) of node slot k of the new topology; side 0 also clears the padding.
PRT_HD void treelet_write_side(Node *nodes, const Treelet &t, const TreeletPlan &pl, int k, int side) {
	const Box b = treelet_subset_box(t, pl.half[k][side], empty_box(), 0, t.n);
	float *f = reinterpret_cast<float *>(nodes + t.slot[k]);
	int32_t *w = reinterpret_cast<int32_t *>(nodes + t.slot[k]);
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		f[6 * side + a] = b.lo[a];
		f[6 * side + 3 + a] = b.hi[a];
	}
	w[12 + side] = pl.ref[k][side];
	if (side == 0)
		w[14] = w[15] = 0;
}

PRT_HD bool treelet_commit(Node *nodes, const Treelet &t, const float *area, const float *copt,
                           const uint8_t *part, int32_t *depth, bool strict) {
	TreeletPlan pl;
	if (!treelet_plan(t, area, copt, part, depth, strict, pl))
		return false;
	for (int k = 0; k < pl.used; ++k) {
		treelet_write_side(nodes, t, pl, k, 0);
		treelet_write_side(nodes, t, pl, k, 1);
		depth[t.slot[k]] = pl.slot_depth[k];
	}
	return true;
}

// Optimises the treelet rooted at internal node x (which must have >= TREELET_N leaves below it).
// depth[i] = height of internal node i's subtree (a node over two triangles has height 1): read
// for the subtrees hanging off the treelet, written for every node slot of the treelet, x included.
PRT_HD bool treelet_optimise(Node *nodes, int32_t x, int32_t *depth, bool strict = false) {
	Treelet t;
	float area[TREELET_SETS], copt[TREELET_SETS];
	uint8_t part[TREELET_SETS];
	treelet_form(nodes, x, depth, t);
	treelet_dp(t, area, copt, part);
	return treelet_commit(nodes, t, area, copt, part, depth, strict);
}

} // namespace prt
