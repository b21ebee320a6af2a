// build.cu -- set_tris on the device: LBVH build (sm_100a).
//
// Replaces the reference's single-threaded binned-SAH BVH2::build (include/portableRT/bvh.hpp:58-193,
// reached from CPUBackend::set_tris, src/intersect_cpu.cpp:12) with a data-parallel build:
//   1 bounds    coalesced float4 tile loads -> per-triangle AABB -> centroid bounds (atomics)
//   2 morton    3*b-bit Morton key of the AABB centre (b = 10/16/21 bits per axis) + identity index
//   3 sort      LSD radix sort, 8-bit digits, (u64 key, u32 index), tiles ranked in shared memory
//               with warp-level match/prefix ranking, coalesced scatter
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

// ------------------------------------------------------------------------------------------------
// 3. LSD radix sort: 8-bit digits, u64 keys, u32 values
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RADIX = 256;

// per-tile digit histogram -> counts[digit * n_tiles + tile]; also accumulates totals[digit]
__global__ void __launch_bounds__(RS_THREADS)
    k_sort_hist(const uint64_t *__restrict__ keys, uint64_t n, int shift,
                uint32_t *__restrict__ counts, uint32_t *__restrict__ totals, uint32_t n_tiles) {
	__shared__ uint32_t hist[RADIX];
	hist[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
	for (int k = 0; k < RS_ITEMS; ++k) {
		const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
		const bool ok = i < n;
		const uint32_t d = ok ? (uint32_t)((keys[i] >> shift) & 0xff) : 256u;
		const uint32_t peers = __match_any_sync(0xffffffffu, d);
		if (ok && (threadIdx.x & 31) == (__ffs(peers) - 1))
			atomicAdd(&hist[d], __popc(peers));
	}
	__syncthreads();
	const uint32_t c = hist[threadIdx.x];
	counts[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = c;
	if (c)
		atomicAdd(&totals[threadIdx.x], c);
}

// block-wide exclusive scan of one value per thread (RS_THREADS threads); returns exclusive prefix,
// *total = block sum.
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o)
			inc += t;
	}
	if (lane == 31)
		warp_sums[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
		uint32_t si = s;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
			if (lane >= o)
				si += t;
		}
		warp_sums[lane] = si - s; // exclusive
		if (lane == 31)
			warp_sums[32] = si;
	}
	__syncthreads();
	const uint32_t res = warp_sums[w] + inc - v;
	if (total)
		*total = warp_sums[32];
	__syncthreads();
	return res;
}

// one block per digit: exclusive scan over that digit's per-tile counts, in place
__global__ void __launch_bounds__(RS_THREADS) k_sort_scan(uint32_t *__restrict__ counts,
                                                          uint32_t n_tiles) {
	__shared__ uint32_t ws[33];
	uint32_t *row = counts + (uint64_t)blockIdx.x * n_tiles;
	uint32_t carry = 0;
	for (uint32_t base = 0; base < n_tiles; base += RS_THREADS) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < n_tiles ? row[i] : 0;
		uint32_t tot;
		const uint32_t ex = block_exscan(v, ws, &tot);
		if (i < n_tiles)
			row[i] = carry + ex;
		carry += tot;
	}
}

__global__ void __launch_bounds__(RS_THREADS)
    k_sort_scatter(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                   uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint64_t n,
                   int shift, const uint32_t *__restrict__ counts,
                   const uint32_t *__restrict__ totals, uint32_t n_tiles) {
	__shared__ uint64_t s_keys[RS_TILE];
	__shared__ uint32_t s_vals[RS_TILE];
	__shared__ uint32_t wh[RS_WARPS][RADIX];
	__shared__ uint32_t lbase[RADIX];
	__shared__ int64_t gofs[RADIX];
	__shared__ uint32_t ws[33];

	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint64_t tile_base = (uint64_t)blockIdx.x * RS_TILE;
	const int tile_cnt = (int)min((uint64_t)RS_TILE, n - tile_base);

#pragma unroll
	for (int k = 0; k < RS_WARPS; ++k)
		wh[k][threadIdx.x] = 0;

	// global digit bases: exclusive scan of the 256 digit totals
	const uint32_t dig_total = totals[threadIdx.x];
	const uint32_t dig_base = block_exscan(dig_total, ws, nullptr); // contains __syncthreads

	// warp-striped arrangement: warp w owns tile positions [w*32*ITEMS, (w+1)*32*ITEMS),
	// item k of lane l sits at w*32*ITEMS + k*32 + l -> memory order == (k, l) order per warp
	uint64_t key[RS_ITEMS];
	uint32_t val[RS_ITEMS];
	uint32_t rank[RS_ITEMS];
	const int wbase = w * 32 * RS_ITEMS;
#pragma unroll
	for (int k = 0; k < RS_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		const bool ok = pos < tile_cnt;
		key[k] = ok ? keys_in[tile_base + pos] : ~0ull;
		val[k] = ok ? vals_in[tile_base + pos] : 0u;
	}
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int k = 0; k < RS_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		const bool ok = pos < tile_cnt;
		const uint32_t d = ok ? (uint32_t)((key[k] >> shift) & 0xff) : 256u;
		const uint32_t peers = __match_any_sync(0xffffffffu, d);
		uint32_t before = 0;
		if (ok)
			before = wh[w][d];
		__syncwarp();
		rank[k] = before + __popc(peers & lt);
		if (ok && lane == (__ffs(peers) - 1))
			wh[w][d] = before + __popc(peers);
		__syncwarp();
	}
	__syncthreads();

	// per digit (thread == digit): exclusive scan over warps, block count
	uint32_t run = 0;
#pragma unroll
	for (int k = 0; k < RS_WARPS; ++k) {
		const uint32_t c = wh[k][threadIdx.x];
		wh[k][threadIdx.x] = run;
		run += c;
	}
	const uint32_t lb = block_exscan(run, ws, nullptr);
	lbase[threadIdx.x] = lb;
	gofs[threadIdx.x] = (int64_t)dig_base +
	                    (int64_t)counts[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] - (int64_t)lb;
	__syncthreads();

	// stage the tile in digit order (stable)
#pragma unroll
	for (int k = 0; k < RS_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		if (pos < tile_cnt) {
			const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
			const uint32_t lp = lbase[d] + wh[w][d] + rank[k];
			s_keys[lp] = key[k];
			s_vals[lp] = val[k];
		}
	}
	__syncthreads();

	// coalesced scatter: consecutive threads write consecutive addresses inside each digit run
#pragma unroll
	for (int k = 0; k < RS_ITEMS; ++k) {
		const int i = k * RS_THREADS + threadIdx.x;
		if (i < tile_cnt) {
			const uint64_t kk = s_keys[i];
			const uint32_t d = (uint32_t)((kk >> shift) & 0xff);
			const int64_t g = gofs[d] + i;
			keys_out[g] = kk;
			vals_out[g] = s_vals[i];
		}
	}
}

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

	const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
	PRT_CUDA(c, c->nodes.reserve(c->n_nodes * sizeof(Node)));
	PRT_CUDA(c, c->trirecs.reserve((n + 1) * sizeof(TriRec)));
	PRT_CUDA(c, c->keys[0].reserve(n * 8));
	PRT_CUDA(c, c->keys[1].reserve(n * 8));
	PRT_CUDA(c, c->vals[0].reserve(n * 4));
	PRT_CUDA(c, c->vals[1].reserve(n * 4));
	PRT_CUDA(c, c->counts.reserve((size_t)RADIX * n_tiles * 4));
	PRT_CUDA(c, c->totals.reserve(RADIX * 4 * 8));
	PRT_CUDA(c, c->bounds.reserve(6 * 4));
	PRT_CUDA(c, c->leaf_box.reserve(n * 32));
	PRT_CUDA(c, c->node_box.reserve(n * 32));
	PRT_CUDA(c, c->parent.reserve(n * 4));
	PRT_CUDA(c, c->leaf_parent.reserve(n * 4));
	PRT_CUDA(c, c->flags.reserve(n * 4));

	const int stream_grid = (int)std::min<uint64_t>((n + TT - 1) / TT, (uint64_t)c->sm_count * 8);
	const int bits = morton_bits_for(n);
	const int passes = (3 * bits + 7) / 8;

	k_bounds_init<<<1, 32, 0, s>>>(c->bounds.as<uint32_t>());
	k_bounds<<<stream_grid, TT, 0, s>>>(d_tris9, n, c->bounds.as<uint32_t>());
	k_morton<<<stream_grid, TT, 0, s>>>(d_tris9, n, c->bounds.as<uint32_t>(), bits,
	                                    c->keys[0].as<uint64_t>(), c->vals[0].as<uint32_t>());
	c->launches += 3;

	int cur = 0;
	if (n > 1) {
		PRT_CUDA(c, cudaMemsetAsync(c->totals.p, 0, RADIX * 4 * passes, s));
		for (int p = 0; p < passes; ++p) {
			uint32_t *tot = c->totals.as<uint32_t>() + p * RADIX;
			k_sort_hist<<<n_tiles, RS_THREADS, 0, s>>>(c->keys[cur].as<uint64_t>(), n, 8 * p,
			                                           c->counts.as<uint32_t>(), tot, n_tiles);
			k_sort_scan<<<RADIX, RS_THREADS, 0, s>>>(c->counts.as<uint32_t>(), n_tiles);
			k_sort_scatter<<<n_tiles, RS_THREADS, 0, s>>>(
			    c->keys[cur].as<uint64_t>(), c->vals[cur].as<uint32_t>(),
			    c->keys[cur ^ 1].as<uint64_t>(), c->vals[cur ^ 1].as<uint32_t>(), n, 8 * p,
			    c->counts.as<uint32_t>(), tot, n_tiles);
			c->launches += 3;
			cur ^= 1;
		}
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
