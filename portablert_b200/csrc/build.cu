// build.cu -- set_tris on the device: LBVH build (sm_100a).
//
// Replaces the reference's single-threaded binned-SAH BVH2::build (include/portableRT/bvh.hpp:58-193,
// reached from CPUBackend::set_tris, src/intersect_cpu.cpp:12) with a data-parallel build:
//   1 bounds    coalesced float4 tile loads -> per-triangle AABB -> centroid bounds (atomics)
//   2 morton    3*b-bit Morton key of the AABB centre (b = 10/16/21 bits per axis) + identity index
//   3 sort      LSD radix sort (sort.cu): 8-bit digits, (u64 key, u32 index), one kernel per pass,
//               tiles ranked in shared memory with warp-level match/prefix operations,
//               decoupled look-back for the global offsets, coalesced scatter
//   4 karras    one thread per internal node: range + split (Karras 2012), parent links
//   5 leaves    gather triangles into Morton order as 64-byte records (v0, edges, own AABB)
//   6 refit     bottom-up, second-arriver-continues with one atomic counter per node; writes the
//               final 64-byte nodes that carry both children's exact boxes
// HBM-bound integer/byte work: no tensor cores.  Algorithmic bytes per triangle are tallied in
// DESIGN.md (build roofline).
#include <algorithm>

#include "prt_ctx.h"

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

// ------------------------------------------------------------------------------------------------
// 1. centroid bounds
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TT) k_bounds(const float *__restrict__ tris9, uint64_t n,
                                               uint32_t *__restrict__ bounds_ord) {
	__shared__ __align__(16) float sm[TT * 9];
	__shared__ float red[6][TT / 32];
	float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	const uint64_t ntiles = (n + TT - 1) / TT;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const uint64_t first = tile * TT;
		const int cnt = (int)min((uint64_t)TT, n - first);
		__syncthreads();
		load_tri_tile(sm, tris9, first, cnt);
		__syncthreads();
		if ((int)threadIdx.x < cnt) {
			Box b = tri_box(sm + threadIdx.x * 9);
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				float c = 0.5f * b.lo[a] + 0.5f * b.hi[a];
				lo[a] = fminf(lo[a], c);
				hi[a] = fmaxf(hi[a], c);
			}
		}
	}
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
	__shared__ __align__(16) float sm[TT * 9];
	float cmin[3], scale[3];
	const float cells = (float)(1u << bits);
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		cmin[a] = ord2f(bounds_ord[a]);
		float ext = ord2f(bounds_ord[3 + a]) - cmin[a];
		scale[a] = (ext > 0.0f && ext < INFINITY) ? cells / ext : 0.0f;
	}
	const uint32_t qmax = (1u << bits) - 1u;
	const uint64_t ntiles = (n + TT - 1) / TT;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const uint64_t first = tile * TT;
		const int cnt = (int)min((uint64_t)TT, n - first);
		__syncthreads();
		load_tri_tile(sm, tris9, first, cnt);
		__syncthreads();
		if ((int)threadIdx.x < cnt) {
			Box b = tri_box(sm + threadIdx.x * 9);
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
	}
}

// 3. sort: radix_sort_pairs() in sort.cu (one kernel per 8-bit pass, decoupled look-back)

// ------------------------------------------------------------------------------------------------
// 4. Karras hierarchy
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_karras(const uint64_t *__restrict__ keys, int64_t n, Node *__restrict__ nodes,
             int32_t *__restrict__ parent, int32_t *__restrict__ leaf_parent) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1)
		return;
	int32_t l, r;
	karras_node(keys, n, i, l, r);
	nodes[i].child0 = l;
	nodes[i].child1 = r;
	if (l < 0)
		leaf_parent[~l] = (int32_t)i;
	else
		parent[l] = (int32_t)i;
	if (r < 0)
		leaf_parent[~r] = (int32_t)i;
	else
		parent[r] = (int32_t)i;
	if (i == 0)
		parent[0] = -1;
}

// ------------------------------------------------------------------------------------------------
// 5. leaves: triangle records in Morton order + exact leaf boxes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_leaves(const float *__restrict__ tris9, const uint32_t *__restrict__ sorted_idx, uint64_t n,
             TriRec *__restrict__ recs, float4 *__restrict__ leaf_box) {
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
	// edges exactly as core.hpp:33-35 computes them
	rec[0] = make_float4(t[0], t[1], t[2], __uint_as_float(prim));
	rec[1] = make_float4(fsub(t[3], t[0]), fsub(t[4], t[1]), fsub(t[5], t[2]), b.lo[0]);
	rec[2] = make_float4(fsub(t[6], t[0]), fsub(t[7], t[1]), fsub(t[8], t[2]), b.lo[1]);
	rec[3] = make_float4(b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
	leaf_box[2 * j] = make_float4(b.lo[0], b.lo[1], b.lo[2], 0.0f);
	leaf_box[2 * j + 1] = make_float4(b.hi[0], b.hi[1], b.hi[2], 0.0f);
}

// ------------------------------------------------------------------------------------------------
// 6. bottom-up refit
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_box_cg(const float4 *boxes, int64_t idx, Box &b) {
	const float4 lo = __ldcg(boxes + 2 * idx);
	const float4 hi = __ldcg(boxes + 2 * idx + 1);
	b.lo[0] = lo.x;
	b.lo[1] = lo.y;
	b.lo[2] = lo.z;
	b.hi[0] = hi.x;
	b.hi[1] = hi.y;
	b.hi[2] = hi.z;
}

__global__ void __launch_bounds__(256)
    k_refit(Node *__restrict__ nodes, const int32_t *__restrict__ parent,
            const int32_t *__restrict__ leaf_parent, const float4 *__restrict__ leaf_box,
            float4 *__restrict__ node_box, uint32_t *__restrict__ flags, uint64_t n) {
	const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n)
		return;
	int32_t cur = leaf_parent[j];
	while (cur >= 0) {
		// the first thread to arrive stops; the second one owns the node (its sibling subtree is
		// complete and, thanks to the fences, visible)
		if (atomicAdd(&flags[cur], 1u) == 0u)
			return;
		__threadfence();
		const int32_t c0 = nodes[cur].child0, c1 = nodes[cur].child1;
		Box b0, b1;
		if (c0 < 0)
			load_box_cg(leaf_box, ~c0, b0);
		else
			load_box_cg(node_box, c0, b0);
		if (c1 < 0)
			load_box_cg(leaf_box, ~c1, b1);
		else
			load_box_cg(node_box, c1, b1);
		float4 *nd = reinterpret_cast<float4 *>(nodes + cur);
		nd[0] = make_float4(b0.lo[0], b0.lo[1], b0.lo[2], b0.hi[0]);
		nd[1] = make_float4(b0.hi[1], b0.hi[2], b1.lo[0], b1.lo[1]);
		nd[2] = make_float4(b1.lo[2], b1.hi[0], b1.hi[1], b1.hi[2]);
		const Box u = box_union(b0, b1);
		node_box[2 * (int64_t)cur] = make_float4(u.lo[0], u.lo[1], u.lo[2], 0.0f);
		node_box[2 * (int64_t)cur + 1] = make_float4(u.hi[0], u.hi[1], u.hi[2], 0.0f);
		__threadfence();
		cur = parent[cur];
	}
}

// single-triangle scene (bvh.hpp:165-181): a root whose first child is the triangle and whose
// second child is a zero-area dummy record (det == 0 in intersect_tri, so it can never be hit)
__global__ void k_single(Node *nodes, TriRec *recs, const float4 *leaf_box) {
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
	float4 *rec = reinterpret_cast<float4 *>(recs + 1);
	rec[0] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(0xffffffffu));
	rec[1] = make_float4(0.f, 0.f, 0.f, lo.x);
	rec[2] = make_float4(0.f, 0.f, 0.f, lo.y);
	rec[3] = make_float4(lo.z, hi.x, hi.y, hi.z);
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static int morton_bits_for(uint64_t n) {
	if (n <= (1ull << 15))
		return 10; // 30-bit keys, 4 passes
	if (n <= (1ull << 22))
		return 16; // 48-bit keys, 6 passes
	return 21;     // 63-bit keys, 8 passes
}

int build_lbvh(prt_b200 *c, const float *d_tris9, uint64_t n) {
	cudaStream_t s = c->stream;
	c->n_tris = n;
	c->n_nodes = n == 0 ? 0 : (n == 1 ? 1 : n - 1);
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
	PRT_CUDA(c, c->leaf_box.reserve(n * 32));
	PRT_CUDA(c, c->node_box.reserve(n * 32));
	PRT_CUDA(c, c->parent.reserve(n * 4));
	PRT_CUDA(c, c->leaf_parent.reserve(n * 4));
	PRT_CUDA(c, c->flags.reserve(n * 4));

	const int stream_grid = (int)std::min<uint64_t>((n + TT - 1) / TT, (uint64_t)c->sm_count * 8);
	const int bits = morton_bits_for(n);

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

	const int g = (int)((n + 255) / 256);
	k_leaves<<<g, 256, 0, s>>>(d_tris9, c->vals[cur].as<uint32_t>(), n, c->trirecs.as<TriRec>(),
	                           c->leaf_box.as<float4>());
	c->launches += 1;
	if (n == 1) {
		k_single<<<1, 1, 0, s>>>(c->nodes.as<Node>(), c->trirecs.as<TriRec>(), c->leaf_box.as<float4>());
		c->launches += 1;
	} else {
		PRT_CUDA(c, cudaMemsetAsync(c->flags.p, 0, (n - 1) * 4, s));
		k_karras<<<(int)((n - 1 + 255) / 256), 256, 0, s>>>(
		    c->keys[cur].as<uint64_t>(), (int64_t)n, c->nodes.as<Node>(), c->parent.as<int32_t>(),
		    c->leaf_parent.as<int32_t>());
		k_refit<<<g, 256, 0, s>>>(c->nodes.as<Node>(), c->parent.as<int32_t>(),
		                          c->leaf_parent.as<int32_t>(), c->leaf_box.as<float4>(),
		                          c->node_box.as<float4>(), c->flags.as<uint32_t>(), n);
		c->launches += 2;
	}
	PRT_CUDA(c, cudaGetLastError());
	return PRT_OK;
}

} // namespace prt
