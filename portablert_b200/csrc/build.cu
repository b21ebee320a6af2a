// build.cu -- set_tris on the device: LBVH build (sm_100a).
//
// Replaces the reference's single-threaded binned-SAH BVH2::build (include/portableRT/bvh.hpp:58-193,
// reached from CPUBackend::set_tris, src/intersect_cpu.cpp:12) with a data-parallel build:
//   1 bounds    coalesced float4 tile loads -> per-triangle AABB -> centroid bounds (atomics)
//   2 morton    3*b-bit Morton key of the AABB centre (b = 10/16/21 bits per axis) + identity index
//   3 sort      LSD radix sort (sort.cu): 8-bit digits, (u64 key, u32 index), one kernel per pass,
//               tiles ranked in shared memory with warp-level match/prefix operations,
//               decoupled look-back for the global offsets, coalesced scatter
//   4 hierarchy one bottom-up kernel: gather triangles into Morton order as 64-byte records (v0,
//               edges, own AABB), find each subtree's parent from the neighbouring key deltas
//               (the Karras 2012 radix tree, built bottom-up as in Apetrei 2014), write the final
//               64-byte nodes that carry both children's exact boxes; one atomicExch per node
// HBM-bound integer/byte work: no tensor cores.  Algorithmic bytes per triangle are tallied in
// DESIGN.md (build roofline).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "prt_ctx.h"
#include "prt_treelet.cuh"

namespace prt {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
	uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
	return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

constexpr int TT = 256; // triangles per tile == threads per block for the streaming kernels

// Coalesced load of `cnt` consecutive 36-byte triangle records into shared memory (cnt*9 floats),
// as 16-byte vectors when the tile start is 16-byte aligned, then each thread reads its own
// record with a 9-word stride (odd => bank-conflict free).
__device__ __forceinline__ void load_tri_tile(float *sm, const float *tris9, uint64_t first,
                                              int cnt) {
	const float *src = tris9 + first * 9;
	const int nf = cnt * 9;
	if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
		const int n4 = nf >> 2;
		const float4 *s4 = reinterpret_cast<const float4 *>(src);
		float4 *d4 = reinterpret_cast<float4 *>(sm);
		for (int i = threadIdx.x; i < n4; i += blockDim.x)
			d4[i] = __ldg(s4 + i);
		for (int i = (n4 << 2) + threadIdx.x; i < nf; i += blockDim.x)
			sm[i] = __ldg(src + i);
	} else {
		for (int i = threadIdx.x; i < nf; i += blockDim.x)
			sm[i] = __ldg(src + i);
	}
}

// Streams the triangle array through shared memory tile by tile (grid-stride) and calls
// body(tile in shared memory, first triangle, count) for each.  Full tiles -- 256 triangles =
// 9 216 bytes, a multiple of 16 -- are fetched by the bulk-copy engine (cp.async.bulk: UBLKCP on
// sm_100a, completion on an mbarrier), double-buffered: one thread issues the copy of the block's
// NEXT tile before the block computes on the current one, so no thread spends instructions or
// registers on the load and the copy overlaps the arithmetic.  A ragged last tile or a triangle
// array that is not 16-byte aligned takes the plain loader.
template <class F>
__device__ __forceinline__ void for_each_tri_tile(const float *__restrict__ tris9, uint64_t n,
                                                  float (*sm)[TT * 9], uint64_t *bar, F body) {
	constexpr uint32_t TILE_BYTES = TT * 36;
	const uint64_t ntiles = (n + TT - 1) / TT;
	const bool aligned = (reinterpret_cast<uintptr_t>(tris9) & 15) == 0;
	const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bar);
	const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(&sm[0][0]);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto bulk = [&](uint64_t tile) { return aligned && (tile + 1) * TT <= n; };
	auto issue = [&](uint64_t tile, int buf) {
		const uint32_t bb = bar0 + 8 * buf, dst = sm0 + TILE_BYTES * buf;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(TILE_BYTES)
		             : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             ::"r"(dst), "l"(tris9 + tile * (uint64_t)(TT * 9)), "r"(TILE_BYTES), "r"(bb)
		             : "memory");
	};
	uint64_t tile = blockIdx.x;
	uint32_t phase = 0; // bit b: parity the next completion of buffer b will have
	if (threadIdx.x == 0 && tile < ntiles && bulk(tile))
		issue(tile, 0);
	for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
		const int buf = it & 1;
		const uint64_t next = tile + gridDim.x;
		// (the other buffer was released by the __syncthreads that ended the previous iteration)
		if (threadIdx.x == 0 && next < ntiles && bulk(next))
			issue(next, buf ^ 1);
		const uint64_t first = tile * TT;
		const int cnt = (int)min((uint64_t)TT, n - first);
		if (bulk(tile)) {
			const uint32_t bb = bar0 + 8 * buf, parity = (phase >> buf) & 1u;
			uint32_t done = 0;
			while (!done)
				asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; "
				             "selp.u32 %0, 1, 0, p; }"
				             : "=r"(done)
				             : "r"(bb), "r"(parity)
				             : "memory");
			phase ^= 1u << buf;
		} else {
			load_tri_tile(sm[buf], tris9, first, cnt);
			__syncthreads();
		}
		body(sm[buf], first, cnt);
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// 1. centroid bounds
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TT) k_bounds(const float *__restrict__ tris9, uint64_t n,
                                               uint32_t *__restrict__ bounds_ord) {
	__shared__ __align__(128) float sm[2][TT * 9];
	__shared__ __align__(8) uint64_t bar[2];
	__shared__ float red[6][TT / 32];
	float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	for_each_tri_tile(tris9, n, sm, bar, [&](const float *tile, uint64_t, int cnt) {
		if ((int)threadIdx.x < cnt) {
			Box b = tri_box(tile + threadIdx.x * 9);
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				float c = 0.5f * b.lo[a] + 0.5f * b.hi[a];
				lo[a] = fminf(lo[a], c);
				hi[a] = fmaxf(hi[a], c);
			}
		}
	});
#pragma unroll
	for (int a = 0; a < 3; ++a) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
			hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
		}
	}
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) {
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			red[a][w] = lo[a];
			red[3 + a][w] = hi[a];
		}
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		float v = red[threadIdx.x][0];
		for (int k = 1; k < TT / 32; ++k)
			v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][k]) : fmaxf(v, red[threadIdx.x][k]);
		if (threadIdx.x < 3)
			atomicMin(bounds_ord + threadIdx.x, f2ord(v));
		else
			atomicMax(bounds_ord + threadIdx.x, f2ord(v));
	}
}

__global__ void k_bounds_init(uint32_t *bounds_ord) {
	if (threadIdx.x < 3)
		bounds_ord[threadIdx.x] = 0xffffffffu;
	else if (threadIdx.x < 6)
		bounds_ord[threadIdx.x] = 0u;
}

// ------------------------------------------------------------------------------------------------
// 2. Morton keys
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TT) k_morton(const float *__restrict__ tris9, uint64_t n,
                                               const uint32_t *__restrict__ bounds_ord, int bits,
                                               uint64_t *__restrict__ keys,
                                               uint32_t *__restrict__ vals) {
	__shared__ __align__(128) float sm[2][TT * 9];
	__shared__ __align__(8) uint64_t bar[2];
	float cmin[3], scale[3];
	const float cells = (float)(1u << bits);
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		cmin[a] = ord2f(bounds_ord[a]);
		float ext = ord2f(bounds_ord[3 + a]) - cmin[a];
		scale[a] = (ext > 0.0f && ext < INFINITY) ? cells / ext : 0.0f;
	}
	const uint32_t qmax = (1u << bits) - 1u;
	for_each_tri_tile(tris9, n, sm, bar, [&](const float *tile, uint64_t first, int cnt) {
		if ((int)threadIdx.x < cnt) {
			Box b = tri_box(tile + threadIdx.x * 9);
			uint32_t q[3];
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				float c = 0.5f * b.lo[a] + 0.5f * b.hi[a];
				float x = (c - cmin[a]) * scale[a];
				// NaN / negative -> 0, large -> qmax
				q[a] = (x > 0.0f) ? (uint32_t)fminf(x, (float)qmax) : 0u;
			}
			keys[first + threadIdx.x] = morton3(q[0], q[1], q[2]);
			vals[first + threadIdx.x] = (uint32_t)(first + threadIdx.x);
		}
	});
}

// 3. sort: radix_sort_pairs() in sort.cu (one kernel per 8-bit pass, decoupled look-back)

// ------------------------------------------------------------------------------------------------
// 4. hierarchy + triangle records + boxes in ONE bottom-up pass
// ------------------------------------------------------------------------------------------------
// delta(i) = "distance" between the sorted (key . index) pairs i and i+1; a smaller value means a
// longer common prefix.  Indices break ties between equal Morton keys, so the pairs are distinct and
// the two candidate deltas of a range never tie.
__device__ __forceinline__ bool delta_less(const uint64_t *__restrict__ keys, int a, int b) {
	const uint64_t xa = __ldg(keys + a) ^ __ldg(keys + a + 1);
	const uint64_t xb = __ldg(keys + b) ^ __ldg(keys + b + 1);
	if (xa != xb)
		return xa < xb;
	return (uint32_t)(a ^ (a + 1)) < (uint32_t)(b ^ (b + 1));
}

struct RootInfo {
	int32_t root;
	float lo[3], hi[3];
	int32_t depth; // height of the tree, written by the treelet pass (0 = not measured)
};

// Shuffle phase shared by the warp and block levels: the alive lanes of the calling warp carry
// subtrees that tile a contiguous leaf range in lane order.  Two NEIGHBOURING alive lanes are
// siblings exactly when the lower one wants node r as a left child and the upper one wants node
// l-1 (= that same r) as a right child; such pairs merge through shuffles -- no atomics, no fences.
// Each round merges every mutually agreeing pair; the loop ends when a round merges nothing (the
// remaining subtrees have siblings elsewhere or not complete yet).
__device__ __forceinline__ void merge_neighbours(bool &alive, int &l, int &r, int32_t &ref, Box &b,
                                                 const uint64_t *__restrict__ keys, int n,
                                                 Node *nodes, RootInfo *root_info, unsigned lane) {
	for (;;) {
		const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
		if (alive && l == 0 && r == n - 1) { // the root: publish its index and the scene box
			root_info->root = ref;
			root_info->depth = 0;
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				root_info->lo[a] = b.lo[a];
				root_info->hi[a] = b.hi[a];
			}
			alive = false;
		}
		bool lc = false;
		if (alive)
			lc = (l == 0) || (r != n - 1 && delta_less(keys, r, l - 1));
		const unsigned lc_mask = __ballot_sync(0xffffffffu, alive && lc);
		const unsigned above = lane == 31 ? 0u : (alive_mask & ~((2u << lane) - 1u));
		const unsigned below = alive_mask & ((1u << lane) - 1u);
		const int next_lane = above ? __ffs(above) - 1 : -1;
		const int prev_lane = below ? 31 - __clz(below) : -1;
		const bool as_left = alive && lc && next_lane >= 0 && !((lc_mask >> next_lane) & 1u);
		const bool as_right = alive && !lc && prev_lane >= 0 && ((lc_mask >> prev_lane) & 1u);
		if (!__ballot_sync(0xffffffffu, as_left))
			break;
		const int p = as_left ? r : l - 1;
		if (as_left || as_right) {
			float2 *nd = reinterpret_cast<float2 *>(nodes + p) + (as_left ? 0 : 3);
			nd[0] = make_float2(b.lo[0], b.lo[1]);
			nd[1] = make_float2(b.lo[2], b.hi[0]);
			nd[2] = make_float2(b.hi[1], b.hi[2]);
			reinterpret_cast<int32_t *>(nodes + p)[as_left ? 12 : 13] = ref;
		}
		// the left lane carries the merged subtree on; it pulls the right lane's box and bound
		const int srcl = as_left ? next_lane : (int)lane;
		Box o;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			o.lo[a] = __shfl_sync(0xffffffffu, b.lo[a], srcl);
			o.hi[a] = __shfl_sync(0xffffffffu, b.hi[a], srcl);
		}
		const int r_other = __shfl_sync(0xffffffffu, r, srcl);
		if (as_left) {
			b = box_union(b, o);
			r = r_other;
			ref = p;
		}
		if (as_right)
			alive = false;
	}
}

// One thread per triangle (in Morton order).  It gathers its triangle, writes the 64-byte record
// and then its subtree climbs: a subtree covering the sorted range [l, r] hangs under internal node
// r (as its left child) if delta(r) < delta(l-1), else under node l-1 (as its right child) -- the
// binary radix tree of Karras 2012 found bottom-up (Apetrei 2014), so no separate top-down
// hierarchy pass and no parent pointers are needed.  Three levels of cooperation:
//   warp   32 consecutive leaves: siblings are neighbouring lanes, merged with shuffles;
//   block  the few subtrees each warp is left with are handed to warp 0 through shared memory and
//          merged the same way (256 consecutive leaves);
//   grid   whatever still waits for a sibling outside the block uses global memory: the carrier
//          writes ITS half of the parent's node (child reference + exact box), then swaps its range
//          bound into bound[p] with one acq_rel exchange -- the first arrival finds -1 and
//          retires, the second finds its sibling's bound, reads the sibling's box from the node
//          and carries the union upwards.
// What is written never depends on which thread arrives first, so the build is deterministic.
__global__ void __launch_bounds__(256)
    k_hierarchy(const float *__restrict__ tris9, const uint32_t *__restrict__ sorted_idx,
                const uint64_t *__restrict__ keys, int n, TriRec *__restrict__ recs, Node *nodes,
                int *bound, RootInfo *root_info, bool vertex_form) {
	__shared__ int s_cnt[8];
	__shared__ int s_l[32], s_r[32], s_ref[32];
	__shared__ float s_box[6][32];
	const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	const unsigned lane = threadIdx.x & 31;
	const int w = threadIdx.x >> 5;
	bool alive = j < n;
	Box b;
	if (alive) {
		const uint32_t prim = sorted_idx[j];
		const float *src = tris9 + (uint64_t)prim * 9;
		float t[9];
#pragma unroll
		for (int k = 0; k < 9; ++k)
			t[k] = __ldg(src + k);
		b = tri_box(t);
		float4 *rec = reinterpret_cast<float4 *>(recs + j);
		rec[0] = make_float4(t[0], t[1], t[2], __uint_as_float(prim));
		if (vertex_form) { // watertight mode: the original vertices, shared ones bit-identical
			rec[1] = make_float4(t[3], t[4], t[5], b.lo[0]);
			rec[2] = make_float4(t[6], t[7], t[8], b.lo[1]);
		} else { // edges exactly as core.hpp:33-35 computes them
			rec[1] = make_float4(fsub(t[3], t[0]), fsub(t[4], t[1]), fsub(t[5], t[2]), b.lo[0]);
			rec[2] = make_float4(fsub(t[6], t[0]), fsub(t[7], t[1]), fsub(t[8], t[2]), b.lo[1]);
		}
		rec[3] = make_float4(b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
	} else {
#pragma unroll
		for (int a = 0; a < 3; ++a)
			b.lo[a] = b.hi[a] = 0.f;
	}
	int l = j, r = j;
	int32_t ref = ~j;

	// ---- warp level
	merge_neighbours(alive, l, r, ref, b, keys, n, nodes, root_info, lane);

	// ---- block level: hand the survivors (in leaf order) to warp 0 if they fit its 32 lanes
	const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
	if (lane == 0)
		s_cnt[w] = __popc(alive_mask);
	__syncthreads();
	int offset = 0, total = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		offset += k < w ? s_cnt[k] : 0;
		total += s_cnt[k];
	}
	if (total > 1 && total <= 32) {
		if (alive) {
			const int idx = offset + __popc(alive_mask & ((1u << lane) - 1u));
			s_l[idx] = l;
			s_r[idx] = r;
			s_ref[idx] = ref;
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				s_box[a][idx] = b.lo[a];
				s_box[3 + a][idx] = b.hi[a];
			}
		}
		__syncthreads();
		alive = false; // warp 0 carries the survivors from here on
		if (w == 0) {
			alive = (int)lane < total;
			if (alive) {
				l = s_l[lane];
				r = s_r[lane];
				ref = s_ref[lane];
#pragma unroll
				for (int a = 0; a < 3; ++a) {
					b.lo[a] = s_box[a][lane];
					b.hi[a] = s_box[3 + a][lane];
				}
			}
			merge_neighbours(alive, l, r, ref, b, keys, n, nodes, root_info, lane);
		}
	}
	if (!alive)
		return;

	// ---- grid level
	for (;;) {
		const bool left_child = (l == 0) || (r != n - 1 && delta_less(keys, r, l - 1));
		const int p = left_child ? r : l - 1;
		float2 *nd = reinterpret_cast<float2 *>(nodes + p) + (left_child ? 0 : 3);
		nd[0] = make_float2(b.lo[0], b.lo[1]);
		nd[1] = make_float2(b.lo[2], b.hi[0]);
		nd[2] = make_float2(b.hi[1], b.hi[2]);
		reinterpret_cast<int32_t *>(nodes + p)[left_child ? 12 : 13] = ref;
		// release my half of the node, acquire the sibling's: one acq_rel exchange
		int other;
		asm volatile("atom.acq_rel.gpu.global.exch.b32 %0, [%1], %2;"
		             : "=r"(other)
		             : "l"(bound + p), "r"(left_child ? l : r)
		             : "memory");
		if (other == -1)
			return; // first arrival: the sibling will carry on
		const float2 *sib = reinterpret_cast<const float2 *>(nodes + p) + (left_child ? 3 : 0);
		const float2 s0 = __ldcg(sib), s1 = __ldcg(sib + 1), s2 = __ldcg(sib + 2);
		Box o;
		o.lo[0] = s0.x;
		o.lo[1] = s0.y;
		o.lo[2] = s1.x;
		o.hi[0] = s1.y;
		o.hi[1] = s2.x;
		o.hi[2] = s2.y;
		b = left_child ? box_union(b, o) : box_union(o, b); // always (left, right): bit-reproducible
		if (left_child)
			r = other;
		else
			l = other;
		ref = p;
		if (l == 0 && r == n - 1) { // the root: publish its index and the scene box
			root_info->root = p;
			root_info->depth = 0;
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				root_info->lo[a] = b.lo[a];
				root_info->hi[a] = b.hi[a];
			}
			return;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// 5. leaves: triangle records in Morton order + exact leaf boxes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_leaves(const float *__restrict__ tris9, const uint32_t *__restrict__ sorted_idx, uint64_t n,
             TriRec *__restrict__ recs, float4 *__restrict__ leaf_box, bool vertex_form) {
	const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n)
		return;
	const uint32_t prim = sorted_idx[j];
	const float *src = tris9 + (uint64_t)prim * 9;
	float t[9];
#pragma unroll
	for (int k = 0; k < 9; ++k)
		t[k] = __ldg(src + k);
	const Box b = tri_box(t);
	float4 *rec = reinterpret_cast<float4 *>(recs + j);
	rec[0] = make_float4(t[0], t[1], t[2], __uint_as_float(prim));
	if (vertex_form) {
		rec[1] = make_float4(t[3], t[4], t[5], b.lo[0]);
		rec[2] = make_float4(t[6], t[7], t[8], b.lo[1]);
	} else { // edges exactly as core.hpp:33-35 computes them
		rec[1] = make_float4(fsub(t[3], t[0]), fsub(t[4], t[1]), fsub(t[5], t[2]), b.lo[0]);
		rec[2] = make_float4(fsub(t[6], t[0]), fsub(t[7], t[1]), fsub(t[8], t[2]), b.lo[1]);
	}
	rec[3] = make_float4(b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
	leaf_box[2 * j] = make_float4(b.lo[0], b.lo[1], b.lo[2], 0.0f);
	leaf_box[2 * j + 1] = make_float4(b.hi[0], b.hi[1], b.hi[2], 0.0f);
}

// single-triangle scene (bvh.hpp:165-181): a root whose first child is the triangle and whose
// second child is a zero-area dummy record (det == 0 in intersect_tri, so it can never be hit)
__global__ void k_single(Node *nodes, TriRec *recs, const float4 *leaf_box, RootInfo *root_info) {
	Node nd;
	const float4 lo = leaf_box[0], hi = leaf_box[1];
	nd.lo0[0] = nd.lo1[0] = lo.x;
	nd.lo0[1] = nd.lo1[1] = lo.y;
	nd.lo0[2] = nd.lo1[2] = lo.z;
	nd.hi0[0] = nd.hi1[0] = hi.x;
	nd.hi0[1] = nd.hi1[1] = hi.y;
	nd.hi0[2] = nd.hi1[2] = hi.z;
	nd.child0 = ~0;
	nd.child1 = ~1;
	nd.pad0 = nd.pad1 = 0;
	nodes[0] = nd;
	root_info->root = 0;
	root_info->depth = 0;
	root_info->lo[0] = lo.x;
	root_info->lo[1] = lo.y;
	root_info->lo[2] = lo.z;
	root_info->hi[0] = hi.x;
	root_info->hi[1] = hi.y;
	root_info->hi[2] = hi.z;
	float4 *rec = reinterpret_cast<float4 *>(recs + 1);
	rec[0] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(0xffffffffu));
	rec[1] = make_float4(0.f, 0.f, 0.f, lo.x);
	rec[2] = make_float4(0.f, 0.f, 0.f, lo.y);
	rec[3] = make_float4(lo.z, hi.x, hi.y, hi.z);
}

// ------------------------------------------------------------------------------------------------
// 5b. opt-in SAH optimisation: bottom-up treelet restructuring (prt_treelet.cuh)
// ------------------------------------------------------------------------------------------------
// parent links of the current tree (the hierarchy kernel needs none); rebuilt before every pass
__global__ void __launch_bounds__(256)
    k_parents(const Node *__restrict__ nodes, int n_nodes, int32_t *__restrict__ parent,
              int32_t *__restrict__ leaf_parent) {
	const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= n_nodes)
		return;
	const int4 tail = __ldg(reinterpret_cast<const int4 *>(nodes + i) + 3);
	const int32_t c0 = tail.x, c1 = tail.y;
	if (c0 >= 0)
		parent[c0] = i;
	else
		leaf_parent[~c0] = i;
	if (c1 >= 0)
		parent[c1] = i;
	else
		leaf_parent[~c1] = i;
}

// The dynamic programme of one treelet spread over the 32 lanes of a warp (same recurrence and the
// same tie rule as the sequential treelet_dp of prt_treelet.cuh, so the result is identical):
// subset areas four per lane; subsets of 2..5 leaves one per lane; the 7 subsets of 6 leaves by 4
// lanes each and the full set by all 32, merged with shuffles on the key (cost, half).
struct TreeletShared {
	Treelet t;
	TreeletPlan plan;
	float area[TREELET_SETS], copt[TREELET_SETS];
	uint8_t part[TREELET_SETS];
	int commit;
};

__device__ __forceinline__ void reduce_split(float &best, int &bp, int width) {
	for (int o = 1; o < width; o <<= 1) {
		const float ob = __shfl_xor_sync(0xffffffffu, best, o);
		const int op = __shfl_xor_sync(0xffffffffu, bp, o);
		if (ob < best || (ob == best && op < bp)) {
			best = ob;
			bp = op;
		}
	}
}

__device__ void treelet_optimise_warp(Node *nodes, int32_t x, int32_t *depth, TreeletShared &sh,
                                      const uint8_t *by_size, const uint8_t *size_off,
                                      unsigned lane, bool strict) {
	if (lane == 0)
		treelet_form(nodes, x, depth, sh.t);
	__syncwarp();
	{ // subsets lane, lane+32, lane+64, lane+96 share their five low leaves
		const Box low = treelet_subset_box(sh.t, lane, empty_box(), 0, 5);
#pragma unroll
		for (int hi = 0; hi < 4; ++hi) {
			const int s = (int)lane | (hi << 5);
			sh.area[s] = s ? box_half_area(treelet_subset_box(sh.t, s, low, 5, TREELET_N)) : 0.0f;
			if ((s & (s - 1)) == 0) {
				sh.copt[s] = 0.0f;
				sh.part[s] = 0;
			}
		}
	}
	__syncwarp();
	for (int k = 2; k <= 5; ++k) {
		for (int i = size_off[k] + lane; i < size_off[k + 1]; i += 32) {
			const int s = by_size[i];
			float best;
			int bp;
			treelet_best_split(sh.copt, s, 0, 1, best, bp);
			sh.copt[s] = fadd(sh.area[s], best);
			sh.part[s] = (uint8_t)(bp == 0xff ? treelet_first_split(s) : bp);
		}
		__syncwarp();
	}
	{ // 7 subsets of 6 leaves, 4 lanes each (lanes 28..31 idle)
		const int which = lane >> 2;
		const int s = which < TREELET_N ? by_size[size_off[6] + which] : 0;
		float best = INFINITY;
		int bp = 0xff;
		if (s)
			treelet_best_split_quarter(sh.copt, s, lane & 3, best, bp);
		reduce_split(best, bp, 4);
		if (s && (lane & 3) == 0) {
			sh.copt[s] = fadd(sh.area[s], best);
			sh.part[s] = (uint8_t)(bp == 0xff ? treelet_first_split(s) : bp);
		}
	}
	__syncwarp();
	{ // the full set: its 63 splits are p = 2, 4, .. 126 (leaf 0 stays in the other half)
		const int s = TREELET_SETS - 1;
		float best = INFINITY;
		int bp = 0xff;
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			const int p = ((int)lane + 32 * i + 1) << 1;
			if (p < s) {
				const float c = fadd(sh.copt[p], sh.copt[s ^ p]);
				if (c < best) {
					best = c;
					bp = p;
				}
			}
		}
		reduce_split(best, bp, 32);
		if (lane == 0) {
			sh.copt[s] = fadd(sh.area[s], best);
			sh.part[s] = (uint8_t)(bp == 0xff ? treelet_first_split(s) : bp);
		}
	}
	__syncwarp();
	if (lane == 0)
		sh.commit = treelet_plan(sh.t, sh.area, sh.copt, sh.part, depth, strict, sh.plan) ? 1 : 0;
	__syncwarp();
	if (sh.commit) { // the boxes and halves of the (up to) six nodes, one child per lane
		const int k = lane >> 1, side = lane & 1;
		if (k < sh.plan.used) {
			treelet_write_side(nodes, sh.t, sh.plan, k, side);
			if (side == 0)
				depth[sh.t.slot[k]] = sh.plan.slot_depth[k];
			__threadfence(); // the lane that owns x publishes it to its parent's other subtree next round
		}
	}
	__syncwarp();
}

// One lane per triangle climbs towards the root; at every internal node the first arrival retires
// and the second one -- which therefore sees both finished subtrees -- has the treelet rooted
// there optimised (if at least TREELET_N triangles hang below) and carries on.  The lanes of a
// warp stay together: each round, the treelets its lanes have reached are optimised one after the
// other by the whole warp.  A treelet only rewrites nodes inside the subtree of its root, which no
// other warp touches any more, and what it writes depends only on that subtree: the result does
// not depend on arrival order.
constexpr int TL_THREADS = 128;
__global__ void __launch_bounds__(TL_THREADS)
    k_treelet(Node *nodes, int n_tris, const int32_t *__restrict__ parent,
              const int32_t *__restrict__ leaf_parent, unsigned *flag, int32_t *count, int32_t *depth,
              RootInfo *root_info, bool strict) {
	__shared__ TreeletShared sh[TL_THREADS / 32];
	__shared__ uint8_t by_size[TREELET_SETS], size_off[TREELET_N + 2];
	{ // subsets ordered by their number of leaves (popcount, then value)
		const int s = threadIdx.x;
		if (s < TREELET_SETS) {
			int pos = 0;
			for (int q = 0; q < TREELET_SETS; ++q) {
				const int a = __popc(q), b = __popc(s);
				pos += (a < b || (a == b && q < s)) ? 1 : 0;
			}
			by_size[pos] = (uint8_t)s;
		}
		if (s <= TREELET_N + 1) {
			int off = 0;
			for (int q = 0; q < TREELET_SETS; ++q)
				off += __popc(q) < s ? 1 : 0;
			size_off[s] = (uint8_t)off; // (size_off[8] = 128 fits)
		}
	}
	__syncthreads();
	const unsigned lane = threadIdx.x & 31;
	TreeletShared &mine = sh[threadIdx.x >> 5];
	const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	const int32_t root = root_info->root;
	bool alive = j < n_tris;
	int32_t cur = alive ? leaf_parent[j] : 0;
	while (__any_sync(0xffffffffu, alive)) {
		bool need = false;
		int32_t total = 0;
		if (alive) {
			__threadfence(); // release what this lane's warp wrote below `cur`
			if (atomicAdd(flag + cur, 1u) == 0u) {
				alive = false;
			} else {
				__threadfence(); // acquire the sibling subtree
				const int4 tail = __ldcg(reinterpret_cast<const int4 *>(nodes + cur) + 3);
				const int32_t c0 = tail.x, c1 = tail.y;
				const int32_t n0 = c0 < 0 ? 1 : __ldcg(count + c0), n1 = c1 < 0 ? 1 : __ldcg(count + c1);
				total = n0 + n1;
				need = total >= TREELET_N;
				if (!need) {
					const int32_t d0 = c0 < 0 ? 0 : __ldcg(depth + c0);
					const int32_t d1 = c1 < 0 ? 0 : __ldcg(depth + c1);
					depth[cur] = 1 + max(d0, d1);
				}
			}
		}
		unsigned todo = __ballot_sync(0xffffffffu, alive && need);
		while (todo) {
			const int src = __ffs(todo) - 1;
			todo &= todo - 1;
			const int32_t x = __shfl_sync(0xffffffffu, cur, src);
			treelet_optimise_warp(nodes, x, depth, mine, by_size, size_off, lane, strict);
		}
		if (alive) {
			count[cur] = total;
			if (cur == root) {
				root_info->depth = __ldcg(depth + cur);
				alive = false;
			} else {
				cur = parent[cur];
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// 5c. temporal reuse (opt-in, mode 3): a new frame of a deforming mesh keeps the optimised topology
// ------------------------------------------------------------------------------------------------
// One thread per leaf: rewrite the triangle record from the new vertices (tris9 == nullptr: keep
// the records, just measure), then climb like the hierarchy kernel does -- write this subtree's box
// into its half of the parent, one arrival counter per node, the second arrival carries the union
// upwards -- but along the stored parent links of the existing topology.  Every box is again the
// exact union of what is below it, so results cannot depend on whether a frame was refitted or
// rebuilt; only the quality of the tree can, which is why the summed surface area of the internal
// nodes (the SAH cost, up to constants) is measured on the way: sah[0..63] partial sums, sah[64]
// the root's area.
__global__ void __launch_bounds__(256)
    k_refit(const float *__restrict__ tris9, int n, TriRec *recs, Node *nodes,
            const int32_t *__restrict__ parent, const int32_t *__restrict__ leaf_parent, unsigned *flag,
            RootInfo *root_info, float *sah, bool vertex_form) {
	const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (j >= n)
		return;
	Box b;
	float4 *rec = reinterpret_cast<float4 *>(recs + j);
	if (tris9) {
		const uint32_t prim = __float_as_uint(rec[0].w);
		const float *src = tris9 + (uint64_t)prim * 9;
		float t[9];
#pragma unroll
		for (int k = 0; k < 9; ++k)
			t[k] = __ldg(src + k);
		b = tri_box(t);
		rec[0] = make_float4(t[0], t[1], t[2], __uint_as_float(prim));
		if (vertex_form) {
			rec[1] = make_float4(t[3], t[4], t[5], b.lo[0]);
			rec[2] = make_float4(t[6], t[7], t[8], b.lo[1]);
		} else { // edges exactly as core.hpp:33-35 computes them
			rec[1] = make_float4(fsub(t[3], t[0]), fsub(t[4], t[1]), fsub(t[5], t[2]), b.lo[0]);
			rec[2] = make_float4(fsub(t[6], t[0]), fsub(t[7], t[1]), fsub(t[8], t[2]), b.lo[1]);
		}
		rec[3] = make_float4(b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
	} else {
		const float4 r1 = rec[1], r2 = rec[2], r3 = rec[3];
		b.lo[0] = r1.w;
		b.lo[1] = r2.w;
		b.lo[2] = r3.x;
		b.hi[0] = r3.y;
		b.hi[1] = r3.z;
		b.hi[2] = r3.w;
	}
	const int32_t root = root_info->root;
	int32_t me = ~j, cur = leaf_parent[j];
	float acc = 0.0f;
	for (;;) {
		// (topology is fixed, but the 32-byte sector is written by other threads: no ld.global.nc)
		const int4 tail = __ldcg(reinterpret_cast<const int4 *>(nodes + cur) + 3);
		const int side = tail.y == me ? 1 : 0;
		float2 *nd = reinterpret_cast<float2 *>(nodes + cur) + (side ? 3 : 0);
		nd[0] = make_float2(b.lo[0], b.lo[1]);
		nd[1] = make_float2(b.lo[2], b.hi[0]);
		nd[2] = make_float2(b.hi[1], b.hi[2]);
		__threadfence();
		if (atomicAdd(flag + cur, 1u) == 0u)
			break;
		__threadfence();
		const float2 *sib = reinterpret_cast<const float2 *>(nodes + cur) + (side ? 0 : 3);
		const float2 s0 = __ldcg(sib), s1 = __ldcg(sib + 1), s2 = __ldcg(sib + 2);
		Box o;
		o.lo[0] = s0.x;
		o.lo[1] = s0.y;
		o.lo[2] = s1.x;
		o.hi[0] = s1.y;
		o.hi[1] = s2.x;
		o.hi[2] = s2.y;
		b = side ? box_union(o, b) : box_union(b, o); // always (child0, child1)
		const float area = box_half_area(b);
		acc += area;
		if (cur == root) {
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				root_info->lo[a] = b.lo[a];
				root_info->hi[a] = b.hi[a];
			}
			sah[64] = area;
			break;
		}
		me = cur;
		cur = parent[cur];
	}
	if (acc != 0.0f)
		atomicAdd(sah + (threadIdx.x & 63), acc);
}

// ------------------------------------------------------------------------------------------------
// 6. compressed 4-wide nodes: wide node i = the grandchildren of binary node i, 8-bit boxes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Node load_node(const Node *nodes, int32_t i) {
	Node nd;
	const float4 *p = reinterpret_cast<const float4 *>(nodes + i);
	float4 *q = reinterpret_cast<float4 *>(&nd);
	q[0] = __ldg(p);
	q[1] = __ldg(p + 1);
	q[2] = __ldg(p + 2);
	q[3] = __ldg(p + 3);
	return nd;
}

__global__ void __launch_bounds__(256)
    k_wide(const Node *__restrict__ nodes, int n_nodes, Node4 *__restrict__ out) {
	const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= n_nodes)
		return;
	const Node nd = load_node(nodes, i);
	WideChild ch[4];
	int cnt = 0;
#pragma unroll
	for (int side = 0; side < 2; ++side) {
		const int32_t c = side ? nd.child1 : nd.child0;
		const float *lo = side ? nd.lo1 : nd.lo0, *hi = side ? nd.hi1 : nd.hi0;
		if (c >= 0) { // internal: take its two children
			const Node g = load_node(nodes, c);
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				ch[cnt].lo[a] = g.lo0[a];
				ch[cnt].hi[a] = g.hi0[a];
				ch[cnt + 1].lo[a] = g.lo1[a];
				ch[cnt + 1].hi[a] = g.hi1[a];
			}
			ch[cnt].ref = g.child0;
			ch[cnt + 1].ref = g.child1;
			cnt += 2;
		} else { // leaf: stays a direct child
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				ch[cnt].lo[a] = lo[a];
				ch[cnt].hi[a] = hi[a];
			}
			ch[cnt].ref = c;
			cnt += 1;
		}
	}
	const Node4 w = make_node4(ch, cnt);
	const float4 *src = reinterpret_cast<const float4 *>(&w);
	float4 *dst = reinterpret_cast<float4 *>(out + i);
	dst[0] = src[0];
	dst[1] = src[1];
	dst[2] = src[2];
	dst[3] = src[3];
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static int build_wide(prt_b200 *c, cudaStream_t s) {
	if (!c->wide_built || c->n_nodes == 0)
		return PRT_OK;
	PRT_CUDA(c, c->nodes4.reserve(c->n_nodes * sizeof(Node4)));
	k_wide<<<(int)((c->n_nodes + 255) / 256), 256, 0, s>>>(c->nodes.as<Node>(), (int)c->n_nodes,
	                                                      c->nodes4.as<Node4>());
	c->launches += 1;
	return PRT_OK;
}

static int treelet_passes(prt_b200 *c, cudaStream_t s, int passes, bool strict) {
	const uint64_t n = c->n_tris;
	for (int pass = 0; pass < passes; ++pass) {
		PRT_CUDA(c, cudaMemsetAsync(c->tl_flag.p, 0, (n - 1) * 4, s));
		k_parents<<<(int)((n - 1 + 255) / 256), 256, 0, s>>>(c->nodes.as<Node>(), (int)(n - 1),
		                                                    c->tl_parent.as<int32_t>(),
		                                                    c->tl_leaf_parent.as<int32_t>());
		k_treelet<<<(int)((n + TL_THREADS - 1) / TL_THREADS), TL_THREADS, 0, s>>>(
		    c->nodes.as<Node>(), (int)n, c->tl_parent.as<int32_t>(), c->tl_leaf_parent.as<int32_t>(),
		    c->tl_flag.as<unsigned>(), c->tl_count.as<int32_t>(), c->tl_depth.as<int32_t>(),
		    c->root_info.as<RootInfo>(), strict);
		c->launches += 2;
	}
	PRT_CUDA(c, cudaGetLastError());
	return PRT_OK;
}

// Refit the current topology to new (or, tris9 == nullptr, the current) triangles and measure the
// SAH cost of the result: *sah_out = summed half areas of the internal nodes / the root's.
// Needs the parent links of the current tree (tl_parent / tl_leaf_parent).  Synchronises `s`.
static int refit_tree(prt_b200 *c, const float *d_tris9, cudaStream_t s, double *sah_out) {
	const uint64_t n = c->n_tris;
	PRT_CUDA(c, c->tl_sah.reserve(65 * 4));
	PRT_CUDA(c, cudaMemsetAsync(c->tl_sah.p, 0, 65 * 4, s));
	PRT_CUDA(c, cudaMemsetAsync(c->tl_flag.p, 0, (n - 1) * 4, s));
	k_refit<<<(int)((n + 255) / 256), 256, 0, s>>>(
	    d_tris9, (int)n, c->trirecs.as<TriRec>(), c->nodes.as<Node>(), c->tl_parent.as<int32_t>(),
	    c->tl_leaf_parent.as<int32_t>(), c->tl_flag.as<unsigned>(), c->root_info.as<RootInfo>(),
	    c->tl_sah.as<float>(), c->recs_vertex_form);
	c->launches += 1;
	float h[65];
	PRT_CUDA(c, cudaMemcpyAsync(h, c->tl_sah.p, sizeof h, cudaMemcpyDeviceToHost, s));
	PRT_CUDA(c, cudaStreamSynchronize(s));
	double sum = 0.0;
	for (int k = 0; k < 64; ++k)
		sum += h[k];
	*sah_out = h[64] > 0.0f ? sum / h[64] : INFINITY;
	return PRT_OK;
}

// Mode 3 (the default): set_tris with as many triangles as the current, optimised scene first
// tries the previous topology (a deforming mesh keeps its connectivity and most of its shape); the
// refitted tree is accepted if its SAH cost stays within REFIT_TOLERANCE of the cost it had when it
// was optimised.  Otherwise the caller rebuilds the plain LBVH, the family's ray counter starts
// over and its threshold doubles (callers that alternate unrelated scenes of one size pay for a
// geometrically shrinking number of wasted optimisations).  *reused says which.
int try_reuse_topology(prt_b200 *c, const float *d_tris9, uint64_t n, cudaStream_t s, bool *reused) {
	*reused = false;
	if (c->optimise_mode != 3 || !c->topology_valid || !c->tree_optimised || n != c->n_tris ||
	    c->recs_vertex_form != (c->watertight != 0))
		return PRT_OK;
	double sah = 0.0;
	if (int rc = refit_tree(c, d_tris9, s, &sah))
		return rc;
	if (sah <= c->sah_ref * REFIT_TOLERANCE) {
		if (int rc = build_wide(c, s))
			return rc;
		c->refits++;
		c->last_sah = sah;
		c->lazy_backoff = 1;
		*reused = true;
	} else {
		c->refit_rejects++;
		c->rays_since_build = 0;
		c->lazy_backoff = std::min<uint64_t>(64, c->lazy_backoff * 2);
	}
	return PRT_OK;
}

// SAH optimisation of the current tree, in place (5b).  The restructured tree is measured: should it
// come out taller than the traversal stack is sized for, the saved radix tree is restored and
// optimised again under the height-preserving rule (treelet_commit: strict).  Synchronises `s`.
int optimise_tree(prt_b200 *c, cudaStream_t s) {
	const uint64_t n = c->n_tris;
	if (n < (uint64_t)TREELET_N || c->optimise_passes < 1)
		return PRT_OK;
	PRT_CUDA(c, c->tl_parent.reserve((n - 1) * 4));
	PRT_CUDA(c, c->tl_leaf_parent.reserve(n * 4));
	PRT_CUDA(c, c->tl_flag.reserve((n - 1) * 4));
	PRT_CUDA(c, c->tl_count.reserve((n - 1) * 4));
	PRT_CUDA(c, c->tl_depth.reserve((n - 1) * 4));
	PRT_CUDA(c, c->tl_backup.reserve((n - 1) * sizeof(Node)));
	PRT_CUDA(c, cudaMemcpyAsync(c->tl_backup.p, c->nodes.p, (n - 1) * sizeof(Node),
	                            cudaMemcpyDeviceToDevice, s));
	int32_t depth = 0;
	const char *depth_src = static_cast<const char *>(c->root_info.p) + offsetof(RootInfo, depth);
	if (int rc = treelet_passes(c, s, c->optimise_passes, false))
		return rc;
	PRT_CUDA(c, cudaMemcpyAsync(&depth, depth_src, 4, cudaMemcpyDeviceToHost, s));
	PRT_CUDA(c, cudaStreamSynchronize(s));
	if (depth > c->max_tree_depth) {
		PRT_CUDA(c, cudaMemcpyAsync(c->nodes.p, c->tl_backup.p, (n - 1) * sizeof(Node),
		                            cudaMemcpyDeviceToDevice, s));
		if (int rc = treelet_passes(c, s, c->optimise_passes, true))
			return rc;
		PRT_CUDA(c, cudaMemcpyAsync(&depth, depth_src, 4, cudaMemcpyDeviceToHost, s));
		PRT_CUDA(c, cudaStreamSynchronize(s));
		c->strict_fallbacks++;
	}
	c->tree_depth = depth;
	c->tree_optimised = true;
	if (c->optimise_mode == 3) { // links of the FINAL topology and its cost: the yardstick for refits
		k_parents<<<(int)((n - 1 + 255) / 256), 256, 0, s>>>(c->nodes.as<Node>(), (int)(n - 1),
		                                                    c->tl_parent.as<int32_t>(),
		                                                    c->tl_leaf_parent.as<int32_t>());
		c->launches += 1;
		if (int rc = refit_tree(c, nullptr, s, &c->sah_ref))
			return rc;
		c->last_sah = c->sah_ref;
		c->topology_valid = true;
	}
	return PRT_OK;
}

// Lazy mode (the default): a scene that keeps being traced is optimised once it has been asked for
// max(LAZY_RAYS_PER_TRI rays per triangle, LAZY_MIN_RAYS) rays -- roughly when the time spent
// tracing the plain tree equals what the optimisation costs (measured: 0.5 ms + 1.2 ms per million
// triangles and pass against 0.15-0.4 ns per ray), the break-even rule that is never more than
// twice as expensive as knowing the future.  A per-frame rebuild (config C5: 8 rays per triangle
// and frame) never pays for it; a static scene gets the better tree after a few batches.  Called
// by the trace entry points (api.cu) before they launch anything.
int maybe_optimise_tree(prt_b200 *c, uint64_t n_rays) {
	c->rays_since_build += n_rays;
	if (c->optimise_mode < 2 || c->tree_optimised || c->n_tris < (uint64_t)TREELET_N ||
	    c->rays_since_build <
	        c->lazy_backoff * std::max<uint64_t>(LAZY_RAYS_PER_TRI * c->n_tris, LAZY_MIN_RAYS))
		return PRT_OK;
	cudaStream_t s = c->stream;
	PRT_CUDA(c, cudaEventRecord(c->ev0, s));
	if (int rc = optimise_tree(c, s)) {
		if (rc != PRT_E_OOM)
			return rc;
		// no room for the scratch (it is reserved before anything is touched): the optimisation is
		// optional, the plain tree stays in place and the trace goes on
		cudaGetLastError();
		c->tree_optimised = true;
		c->err.clear();
		return PRT_OK;
	}
	if (int rc = build_wide(c, s))
		return rc;
	PRT_CUDA(c, cudaEventRecord(c->ev1, s));
	PRT_CUDA(c, cudaStreamSynchronize(s));
	PRT_CUDA(c, cudaEventElapsedTime(&c->last_optimise_ms, c->ev0, c->ev1));
	return PRT_OK;
}

static int morton_bits_for(uint64_t n) {
	if (const char *e = std::getenv("PRT_B200_MORTON_BITS")) { // experiments: 1..21 bits per axis
		const int b = std::atoi(e);
		if (b >= 1 && b <= 21)
			return b;
	}
	// Measured (profiles/r01_results.md): on all five configs 10 bits per axis give the same
	// nodes/ray as 16 or 21 (equal keys are split by index, which keeps the subtrees balanced), so
	// the resolution is chosen for robustness against small detailed objects in large scenes, not
	// for these benchmarks; every 8 key bits cost one more sort pass.
	if (n <= (1ull << 16))
		return 10; // 30-bit keys, 4 passes
	if (n <= (1ull << 24))
		return 13; // 39-bit keys, 5 passes (8192 cells per axis; C4's 10 M triangles: same nodes/ray as 16)
	return 16;     // 48-bit keys, 6 passes (PRT_B200_MORTON_BITS=21: 63-bit keys, 8 passes)
}

int build_lbvh(prt_b200 *c, const float *d_tris9, uint64_t n) {
	cudaStream_t s = c->stream;
	// (mode 3 counts the rays of a scene FAMILY: frames of the same size keep the counter)
	if (!(c->optimise_mode == 3 && n == c->n_tris)) {
		c->rays_since_build = 0;
		c->lazy_backoff = 1;
	}
	c->n_tris = n;
	c->n_nodes = n == 0 ? 0 : (n == 1 ? 1 : n - 1);
	c->wide_built = false;
	c->tree_optimised = false;
	c->topology_valid = false;
	c->tree_depth = 0;
	c->last_optimise_ms = 0.f;
	c->recs_vertex_form = c->watertight != 0;
	const bool vf = c->recs_vertex_form;
	if (n == 0)
		return PRT_OK;
	if (n > 0x7ffffffeull)
		return fail(c, PRT_E_LIMIT, "set_tris: more than 2^31-2 triangles");

	PRT_CUDA(c, c->nodes.reserve(c->n_nodes * sizeof(Node)));
	PRT_CUDA(c, c->trirecs.reserve((n + 1) * sizeof(TriRec)));
	PRT_CUDA(c, c->keys[0].reserve(n * 8));
	PRT_CUDA(c, c->keys[1].reserve(n * 8));
	PRT_CUDA(c, c->vals[0].reserve(n * 4));
	PRT_CUDA(c, c->vals[1].reserve(n * 4));
	PRT_CUDA(c, c->bounds.reserve(6 * 4));
	PRT_CUDA(c, c->leaf_box.reserve(64));
	PRT_CUDA(c, c->bound.reserve(n * 4));
	PRT_CUDA(c, c->root_info.reserve(sizeof(RootInfo)));

	const int stream_grid = (int)std::min<uint64_t>((n + TT - 1) / TT, (uint64_t)c->sm_count * 8);
	const int bits = morton_bits_for(n);
	c->morton_bits = bits;
	const int g = (int)((n + 255) / 256);

	// everything from the centroid bounds to the hierarchy: ~11 dependent launches and memsets
	auto enqueue = [&](int *cur_out) -> int {
		k_bounds_init<<<1, 32, 0, s>>>(c->bounds.as<uint32_t>());
		k_bounds<<<stream_grid, TT, 0, s>>>(d_tris9, n, c->bounds.as<uint32_t>());
		k_morton<<<stream_grid, TT, 0, s>>>(d_tris9, n, c->bounds.as<uint32_t>(), bits,
		                                    c->keys[0].as<uint64_t>(), c->vals[0].as<uint32_t>());
		c->launches += 3;
		int cur = 0;
		if (n > 1) {
			uint64_t *const kk[2] = {c->keys[0].as<uint64_t>(), c->keys[1].as<uint64_t>()};
			uint32_t *const vv[2] = {c->vals[0].as<uint32_t>(), c->vals[1].as<uint32_t>()};
			if (int rc = radix_sort_pairs(c, c->sort_scratch, kk, vv, n, 3 * bits, s, &cur))
				return rc;
		}
		if (n == 1) {
			k_leaves<<<1, 256, 0, s>>>(d_tris9, c->vals[cur].as<uint32_t>(), n, c->trirecs.as<TriRec>(),
			                           c->leaf_box.as<float4>(), vf);
			k_single<<<1, 1, 0, s>>>(c->nodes.as<Node>(), c->trirecs.as<TriRec>(), c->leaf_box.as<float4>(),
			                         c->root_info.as<RootInfo>());
			c->launches += 2;
		} else {
			PRT_CUDA(c, cudaMemsetAsync(c->bound.p, 0xff, (n - 1) * 4, s));
			k_hierarchy<<<g, 256, 0, s>>>(d_tris9, c->vals[cur].as<uint32_t>(), c->keys[cur].as<uint64_t>(),
			                              (int)n, c->trirecs.as<TriRec>(), c->nodes.as<Node>(),
			                              c->bound.as<int>(), c->root_info.as<RootInfo>(), vf);
			c->launches += 1;
		}
		*cur_out = cur;
		return PRT_OK;
	};

	// Small scenes are bound by the latency of that launch chain (69 k triangles: 0.13 ms for ~50 us of
	// kernels).  A scene family that is rebuilt from the same buffers -- same triangle pointer and
	// count, scratch not reallocated in between -- therefore gets its chain captured into a CUDA graph
	// at the second build and replayed with one launch from the third on (env PRT_B200_GRAPHS=0: off).
	struct Key {
		const void *p[12];
		uint64_t n;
		int bits, vf;
	} key{};
	const void *ptrs[12] = {d_tris9,         c->nodes.p,    c->trirecs.p,      c->keys[0].p, c->keys[1].p,
	                        c->vals[0].p,    c->vals[1].p,  c->sort_scratch.p, c->bounds.p,  c->bound.p,
	                        c->root_info.p,  c->leaf_box.p};
	for (int k = 0; k < 12; ++k)
		key.p[k] = ptrs[k];
	key.n = n;
	key.bits = bits;
	key.vf = vf ? 1 : 0;
	static_assert(sizeof(Key) <= sizeof(c->build_graph_key), "graph key storage");
	int cur = 0;
	const bool graphable = c->use_graphs && n > 1 && n <= (1ull << 20);
	if (graphable && c->build_graph && std::memcmp(&key, c->build_graph_key, sizeof key) == 0) {
		PRT_CUDA(c, cudaGraphLaunch(c->build_graph, s));
		c->launches += c->build_graph_launches;
		c->graph_replays++;
	} else if (graphable && c->build_seen && std::memcmp(&key, c->build_seen_key, sizeof key) == 0) {
		const uint64_t l0 = c->launches;
		PRT_CUDA(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
		const int rc = enqueue(&cur);
		cudaGraph_t graph = nullptr;
		const cudaError_t e = cudaStreamEndCapture(s, &graph);
		cudaGraphExec_t exec = nullptr;
		if (rc == PRT_OK && e == cudaSuccess && graph &&
		    cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
			if (c->build_graph)
				cudaGraphExecDestroy(c->build_graph);
			c->build_graph = exec;
			std::memcpy(c->build_graph_key, &key, sizeof key);
			c->build_graph_launches = (int)(c->launches - l0);
			PRT_CUDA(c, cudaGraphLaunch(exec, s));
		} else { // capture not possible: run the chain directly
			cudaGetLastError();
			c->launches = l0;
			if (int rc2 = enqueue(&cur))
				return rc2;
		}
		if (graph)
			cudaGraphDestroy(graph);
	} else {
		if (int rc = enqueue(&cur))
			return rc;
		std::memcpy(c->build_seen_key, &key, sizeof key);
		c->build_seen = true;
	}
	if (c->optimise_mode == 1) // inside every set_tris
		if (int rc = optimise_tree(c, s))
			return rc;
	c->wide_built = !vf && (c->wide_mode == 1 || (c->wide_mode == 2 && n >= (1ull << 20)));
	if (int rc = build_wide(c, s))
		return rc;
	PRT_CUDA(c, cudaGetLastError());
	return PRT_OK;
}

} // namespace prt
