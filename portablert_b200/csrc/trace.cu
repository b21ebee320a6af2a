// trace.cu -- nearest_hits on the device: persistent-thread BVH traversal (sm_100a).
//
// Replaces the reference's per-ray DFS BVH2::nearest_tri (include/portableRT/bvh.hpp:224-265, fanned
// out over std::threads by CPUBackend::nearest_hits, intersect_cpu.hpp:20-43).  Results follow the
// reference's acceptance rule exactly (prt_math.cuh): a triangle counts iff its own AABB passes
// ray_box_intersect and intersect_tri accepts it; the minimum t wins (negative t included, no
// tmax).  What differs is only how much of the tree is visited:
//   * both children's boxes sit in the parent node -> one 64-byte fetch (4 x LDG.128) per step
//   * children are entered nearest-first and a subtree is skipped when its entry distance exceeds
//     the best hit so far (+ slack) -- the reference visits every box the line touches
//   * `valid`-only queries stop at the first accepted triangle
// One kernel per tag mask (31, like the 31 OptiX raygen programs of hitreg.hpp:142-177) x {SoA,
// AoS} output; only the requested fields are tracked and written.  Warps are persistent and pull
// 32-ray packets from a global counter.
#include <algorithm>
#include <cstring>

#include "prt_trace_kernel.cuh"

namespace prt {

// ------------------------------------------------------------------------------------------------
// Ray reordering.  Incoherent batches (bounce rays, random rays) make the 32 lanes of a warp walk
// unrelated parts of the tree: every node fetch touches 32 different lines and lanes finish at
// very different times.  Sorting the batch by a short key -- Morton code of the origin (3 bits per
// axis inside the scene box) above the Morton code of the direction (2 bits per axis) -- puts rays
// that start in the same region and point the same way next to each other.  The traversal then
// processes rays in key order through a permutation; results are written to the rays' own slots,
// so the output order is unchanged.  Measured x1.39 (10 M tris / random rays) and x1.45 (one-bounce
// diffuse rays in the 262 k-tri interior); coherent primary rays gain nothing, which the key
// kernel detects (most neighbouring rays already share their key) so that the sort is skipped.
// Key width (sweep on the B200, traversal kernel ms / whole step ms): what the sort buys is mostly
// lanes that agree on the order of the children; C4 (10^8 random rays) 27.35 / 31.03 with 4 + 4
// bits (3 sort passes), 27.76 / 30.49 with 3 + 2 (2 passes), 27.64 / 29.43 with 1 + 1 (1 pass);
// C3B (8.3 M bounce rays) 0.949 / 1.319, 1.028 / 1.312, 1.145 / 1.336.  3 + 2 is the default: never
// worse than the 24-bit key on either, and it does not bet on uniformly random rays.
__host__ __device__ __forceinline__ uint32_t spread4(uint32_t v) { // 4 bits -> every third bit
	return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

__host__ __device__ __forceinline__ uint32_t ray_key_of(float ox, float oy, float oz, float dx, float dy,
                                                        float dz, float3 lo, float3 inv_ext) {
	// NaN / out-of-box values clamp into the grid; the key only steers the processing order
#if defined(__CUDA_ARCH__)
	const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-37f));
#define PRT_Q16(x) ((uint32_t)fminf(fmaxf((x), 0.f), 15.f))
#else
	const float d2 = dx * dx + dy * dy + dz * dz;
	const float inv = 1.0f / sqrtf(d2 > 1e-37f ? d2 : 1e-37f);
#define PRT_Q16(x) (!((x) > 0.f) ? 0u : ((x) >= 15.f ? 15u : (uint32_t)(x)))
#endif
	const uint32_t qx = PRT_Q16((ox - lo.x) * inv_ext.x * 16.f);
	const uint32_t qy = PRT_Q16((oy - lo.y) * inv_ext.y * 16.f);
	const uint32_t qz = PRT_Q16((oz - lo.z) * inv_ext.z * 16.f);
	const uint32_t ux = PRT_Q16((dx * inv * 0.5f + 0.5f) * 16.f);
	const uint32_t uy = PRT_Q16((dy * inv * 0.5f + 0.5f) * 16.f);
	const uint32_t uz = PRT_Q16((dz * inv * 0.5f + 0.5f) * 16.f);
#undef PRT_Q16
	const uint32_t mo = (spread4(qx) << 2) | (spread4(qy) << 1) | spread4(qz);
	const uint32_t md = (spread4(ux) << 2) | (spread4(uy) << 1) | spread4(uz);
	return (mo << 12) | md;
}

__device__ __forceinline__ uint32_t ray_key(const float *__restrict__ r, float3 lo, float3 inv_ext) {
	return ray_key_of(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3), __ldg(r + 4), __ldg(r + 5),
	                  lo, inv_ext);
}

// Sort key with `ob` bits per axis of the origin above `db` bits per axis of the direction (the
// probe's fixed 4 + 4 key only decides WHETHER to sort).
// Grid-stride over the batch; the digit histograms of the sort passes (digit p = key bits 8p..8p+7)
// are counted on the way -- in shared memory, flushed once per block -- so that the sort does not
// read the keys again for them; the values of the sort are the positions themselves and are not
// written here (the first pass generates them).
__global__ void __launch_bounds__(256)
    k_ray_keys(const float *__restrict__ rays, uint64_t n, float3 lo, float3 inv_ext, int ob, int db,
               uint32_t *__restrict__ keys, int passes, uint32_t *__restrict__ ghist /* [passes][256] */) {
	__shared__ uint32_t sh[4][256];
	for (int p = 0; p < passes; ++p)
		sh[p][threadIdx.x] = 0;
	__syncthreads();
	const float no = (float)(1 << ob), nd = (float)(1 << db);
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const float2 *r = reinterpret_cast<const float2 *>(rays + i * 6); // (6 floats: 8-byte aligned
		float ox, oy, oz, dx, dy, dz;                                     // whenever the batch is)
		if ((reinterpret_cast<uintptr_t>(rays) & 7) == 0) {
			const float2 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
			ox = a.x, oy = a.y, oz = b.x, dx = b.y, dy = c.x, dz = c.y;
		} else {
			const float *f = rays + i * 6;
			ox = __ldg(f), oy = __ldg(f + 1), oz = __ldg(f + 2), dx = __ldg(f + 3), dy = __ldg(f + 4),
			dz = __ldg(f + 5);
		}
		const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-37f));
		// NaN / out-of-box values clamp into the grid; the key only steers the processing order
		const uint32_t qx = (uint32_t)fminf(fmaxf((ox - lo.x) * inv_ext.x * no, 0.f), no - 1.f);
		const uint32_t qy = (uint32_t)fminf(fmaxf((oy - lo.y) * inv_ext.y * no, 0.f), no - 1.f);
		const uint32_t qz = (uint32_t)fminf(fmaxf((oz - lo.z) * inv_ext.z * no, 0.f), no - 1.f);
		const uint32_t ux = (uint32_t)fminf(fmaxf((dx * inv * 0.5f + 0.5f) * nd, 0.f), nd - 1.f);
		const uint32_t uy = (uint32_t)fminf(fmaxf((dy * inv * 0.5f + 0.5f) * nd, 0.f), nd - 1.f);
		const uint32_t uz = (uint32_t)fminf(fmaxf((dz * inv * 0.5f + 0.5f) * nd, 0.f), nd - 1.f);
		const uint32_t key = (uint32_t)((morton3(qx, qy, qz) << (3 * db)) | morton3(ux, uy, uz)); // 3 (ob + db) <= 32 bits
		keys[i] = key;
		const unsigned act = __activemask();
		for (int p = 0; p < passes; ++p) {
			// a digit that is constant across the warp (unused high bits, neighbouring rays of a
			// half-sorted batch) would serialise 32 same-address atomics: aggregate it into one
			const uint32_t d = (key >> (8 * p)) & 0xff;
			int same;
			__match_all_sync(act, d, &same);
			if (same) {
				if ((threadIdx.x & 31) == (__ffs(act) - 1))
					atomicAdd(&sh[p][d], (uint32_t)__popc(act));
			} else {
				atomicAdd(&sh[p][d], 1u);
			}
		}
	}
	__syncthreads();
	for (int p = 0; p < passes; ++p) {
		const uint32_t c = sh[p][threadIdx.x];
		if (c)
			atomicAdd(&ghist[p * 256 + threadIdx.x], c);
	}
}

// Coherence probe: PROBE_WARPS segments of 32 consecutive rays spread over the batch; counts the
// rays that carry the same key as their predecessor.  The last block to finish publishes the total
// through mapped pinned memory, so the host needs one stream synchronisation and no copy.
constexpr int PROBE_BLOCKS = 32, PROBE_WARPS = PROBE_BLOCKS * 8;
__global__ void __launch_bounds__(256)
    k_ray_probe(const float *__restrict__ rays, uint64_t n, float3 lo, float3 inv_ext,
                unsigned long long *__restrict__ acc /* same count */,
                unsigned long long *__restrict__ ticket, volatile unsigned long long *host_flag) {
	const int lane = threadIdx.x & 31;
	const uint64_t warp = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
	const uint64_t n_seg = n / 32;
	const uint64_t seg = n_seg >= PROBE_WARPS ? warp * (n_seg / PROBE_WARPS) : warp;
	const uint64_t i = seg * 32 + lane;
	uint32_t key = 0xffffffffu - lane; // distinct for lanes past the end
	if (seg < n_seg)
		key = ray_key(rays + i * 6, lo, inv_ext);
	const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
	const unsigned m = __ballot_sync(0xffffffffu, lane != 0 && prev == key && seg < n_seg);
	if (lane == 0 && m)
		atomicAdd(acc, (unsigned long long)__popc(m));
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(ticket, 1ull) == PROBE_BLOCKS - 1) {
			*host_flag = atomicAdd(acc, 0ull) + 1; // +1: 0 means "not written yet"
			__threadfence_system();
		}
	}
}

static void k_ray_probe_launch(prt_b200 *c, const float *rays, uint64_t n, float3 lo, float3 ie,
                               unsigned long long *acc, unsigned long long *ticket,
                               unsigned long long *host_flag_dev, cudaStream_t s) {
	k_ray_probe<<<PROBE_BLOCKS, 256, 0, s>>>(rays, n, lo, ie, acc, ticket, host_flag_dev);
	c->launches += 1;
}

constexpr uint64_t SORT_MIN_RAYS = 1u << 16;


KernelFn trace_kernel_fast(uint32_t mask, bool aos) { return kernel_of<false, false, false>(mask, aos); }

static void scene_grid(const prt_b200 *c, float3 &lo, float3 &ie) {
	lo = make_float3(c->scene_lo[0], c->scene_lo[1], c->scene_lo[2]);
	ie.x = c->scene_hi[0] > c->scene_lo[0] ? 1.f / (c->scene_hi[0] - c->scene_lo[0]) : 0.f;
	ie.y = c->scene_hi[1] > c->scene_lo[1] ? 1.f / (c->scene_hi[1] - c->scene_lo[1]) : 0.f;
	ie.z = c->scene_hi[2] > c->scene_lo[2] ? 1.f / (c->scene_hi[2] - c->scene_lo[2]) : 0.f;
}

// The coherence probe of k_ray_probe on a batch that still lives in HOST memory (same sampled
// statistic on fewer segments): the host entry point decides once per call, so that its chunk
// pipeline never waits for a device-side probe.  1 = incoherent (reorder), 0 = coherent.
int host_ray_probe(const prt_b200 *c, const float *rays6, uint64_t n) {
	float3 lo, ie;
	scene_grid(c, lo, ie);
	constexpr uint64_t HOST_SEGS = 64; // a quarter of the device probe's sample: ~20 us of host time
	const uint64_t n_seg = n / 32;
	const uint64_t segs = std::min<uint64_t>(HOST_SEGS, n_seg);
	uint64_t same = 0;
	for (uint64_t w = 0; w < segs; ++w) {
		const uint64_t seg = n_seg >= HOST_SEGS ? w * (n_seg / HOST_SEGS) : w;
		const float *r = rays6 + seg * 32 * 6;
		uint32_t prev = ray_key_of(r[0], r[1], r[2], r[3], r[4], r[5], lo, ie);
		for (int l = 1; l < 32; ++l) {
			r += 6;
			const uint32_t k = ray_key_of(r[0], r[1], r[2], r[3], r[4], r[5], lo, ie);
			same += (k == prev);
			prev = k;
		}
	}
	return same * 2 < segs * 31 ? 1 : 0;
}

// Upper bound of the traversal stack depth for the current tree: one entry per level of the
// current path.  Radix tree over (3*bits-bit key . index): height <= 3*bits + bit length of n;
// an optimised tree has its height measured by the treelet kernel.  Wide nodes push up to three
// entries per two levels.
static int stack_bound(const prt_b200 *c, bool wide) {
	int bitlen = 0;
	for (uint64_t n = c->n_tris; n; n >>= 1)
		++bitlen;
	int depth = std::min(96, 3 * c->morton_bits + bitlen + 2);
	if (c->tree_optimised && c->tree_depth > 0)
		depth = std::min(96, (int)c->tree_depth + 1);
	return wide ? 3 * ((depth + 1) / 2) + 3 : depth;
}

static int launch_kernel(prt_b200 *c, KernelFn fn, int cache_slot, uint32_t mask, TraceParams &P,
                         uint64_t n, bool wide, int slot, cudaStream_t s, bool timed = false,
                         bool coop = false) {
	int bps = cache_slot >= 0 ? c->bps_cache[mask][cache_slot] : 0;
	if (bps == 0) {
		PRT_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, TRACE_THREADS, 0));
		if (bps < 1)
			bps = 1;
		if (cache_slot >= 0)
			c->bps_cache[mask][cache_slot] = bps;
	}
	const uint64_t want = (n + TRACE_THREADS - 1) / TRACE_THREADS;
	const uint64_t grid = std::min<uint64_t>(want, (uint64_t)c->sm_count * bps);
	// stack overflow area: sized for the largest grid this context launches and the current tree
	const uint64_t threads = (uint64_t)c->sm_count * 16 * TRACE_THREADS;
	const int deep = std::max(1, stack_bound(c, wide) - SMEM_STACK + 1);
	PRT_CUDA(c, c->stack_ovf[slot].reserve(threads * (uint64_t)deep * sizeof(uint2)));
	P.stack_ovf = c->stack_ovf[slot].as<uint2>();
	P.ovf_stride = (uint32_t)threads;
	// cooperative tail of the binary fast kernels: a list of handed-over rays (state + stack) and,
	// for the follow-up kernel, a LIFO of 32 entries per tree level for each of its warps
	P.coop = nullptr;
	P.coop_cap = 0;
	P.coop_after = 0;
	const bool use_coop = coop && c->coop_after > 0;
	constexpr uint32_t COOP_CAP = 8192;
	// (8 blocks per SM measured best on a 2 M-ray batch, 4 on the 260 k-ray chunks of the host pipeline,
	// where an almost empty follow-up kernel is pure launch and drain time)
	const int coop_grid = c->sm_count * (n >= (1ull << 20) ? c->coop_blocks : std::min(c->coop_blocks, 4));
	if (use_coop) {
		const uint32_t depth = (uint32_t)stack_bound(c, false) + 1;
		const uint32_t rec = COOP_PARK + 2 * depth;
		const bool fresh = c->coop_buf[slot].p == nullptr;
		PRT_CUDA(c, c->coop_buf[slot].reserve((4 + (uint64_t)COOP_CAP * rec) * 4));
		if (fresh || c->coop_buf[slot].p != c->coop_seen[slot]) { // (re)allocated: arm the counters
			PRT_CUDA(c, cudaMemsetAsync(c->coop_buf[slot].p, 0, 16, s));
			c->coop_seen[slot] = c->coop_buf[slot].p;
		}
		const uint32_t lifo_cap = 32 * (depth + 3);
		PRT_CUDA(c, c->coop_lifo[slot].reserve((uint64_t)coop_grid * (COOP_THREADS / 32) * lifo_cap * 8));
		P.coop = c->coop_buf[slot].as<uint32_t>();
		P.coop_cap = COOP_CAP;
		P.coop_rec = rec;
		P.coop_depth = depth;
		P.coop_lifo = c->coop_lifo[slot].as<uint2>();
		P.coop_lifo_cap = lifo_cap;
		P.coop_after = c->coop_after;
		P.coop_min_sp = c->coop_min_sp;
	}
	if (timed)
		PRT_CUDA(c, cudaEventRecord(c->ev_k0, s));
	fn<<<(unsigned)grid, TRACE_THREADS, 0, s>>>(P);
	c->launches += 1;
	if (use_coop) { // finishes the rays the kernel above handed over (none: its warps leave at once)
		coop_kernel(mask, P.aos != nullptr, c->recs_vertex_form)<<<coop_grid, COOP_THREADS, 0, s>>>(P);
		c->launches += 1;
	}
	if (timed)
		PRT_CUDA(c, cudaEventRecord(c->ev_k1, s));
	PRT_CUDA(c, cudaGetLastError());
	return PRT_OK;
}

// mode of the second (exact) pass over the rays the fast kernel set aside:
//   EXOTIC_INLINE    launched unconditionally right behind the fast kernel (host pipeline: no
//                    synchronisation point between chunks; an empty pass costs a few microseconds)
//   EXOTIC_DEFERRED  the caller synchronises the stream and then calls finish_exotic(), which
//                    launches the pass only if the fast kernel reported set-aside rays
int launch_trace(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t mask, const TraceOut &out,
                 uint32_t *d_counts, cudaStream_t s, int coherence, int exotic_mode) {
	if (mask == 0 || mask > PRT_TAG_ALL)
		return fail(c, PRT_E_ARG, "nearest_hits: tag mask must be in 1..31");
	c->pending_exotic = false;
	if (n == 0)
		return PRT_OK;
	TraceParams P{};
	P.nodes = c->nodes.as<Node>();
	P.nodes4 = c->nodes4.as<Node4>();
	P.tris = c->trirecs.as<TriRec>();
	P.rays = d_rays6;
	P.n_rays = n;
	P.n_tris = c->n_tris;
	P.root = c->root;
	P.uv = reinterpret_cast<float2 *>(out.soa.uv);
	P.t = out.soa.t;
	P.pid = out.soa.pid;
	P.p = out.soa.p;
	P.valid = out.soa.valid;
	P.aos = static_cast<char *>(out.aos);
	P.lay = out.layout;
	P.counts = d_counts;
	P.prune = c->opts.prune;
	P.slack_rel = c->opts.slack_rel;
	P.slack_ulps = c->opts.slack_ulps;
	for (int a = 0; a < 3; ++a)
		P.scene_absmax[a] = c->scene_absmax[a];
	P.refill = c->refill;
	P.leaf_votes = c->leaf_votes;
	P.rays_vec = (reinterpret_cast<uintptr_t>(d_rays6) & 7) == 0;
	// [0] ray counter, [1] coherence-probe count, [2] warps that left, [3] rays set aside: zeroed at
	// create, re-armed by the kernels themselves (32 bytes per launch slot)
	P.counter = c->counter.as<unsigned long long>() + 4 * out.slot;

	// ---- optional ray reordering (see k_ray_keys)
	P.perm = nullptr;
	if (c->sort_rays && !d_counts && n >= SORT_MIN_RAYS && n < (1ull << 32) && c->n_tris > 1) {
		auto &rs = c->rs[out.slot];
		for (int k = 0; k < 2; ++k) {
			PRT_CUDA(c, rs.keys[k].reserve(n * 4));
			PRT_CUDA(c, rs.vals[k].reserve(n * 4));
		}
		float3 lo, ie;
		scene_grid(c, lo, ie);
		bool do_sort = true;
		if (c->sort_rays == 2 && coherence >= 0) { // the caller probed the batch on the host
			do_sort = coherence != 0;
		} else if (c->sort_rays == 2) { // auto: skip batches that are already coherent
			volatile unsigned long long *flag = c->probe_host + out.slot;
			*flag = 0;
			PRT_CUDA(c, cudaMemsetAsync(P.counter + 1, 0, 8, s));
			PRT_CUDA(c, cudaMemsetAsync(c->probe_ticket.as<unsigned long long>() + out.slot, 0, 8, s));
			// acc[0] = same count lives next to the ray counter, acc[1] = the block ticket
			k_ray_probe_launch(c, d_rays6, n, lo, ie, P.counter + 1,
			                   c->probe_ticket.as<unsigned long long>() + out.slot,
			                   c->probe_dev + out.slot, s);
			PRT_CUDA(c, cudaStreamSynchronize(s));
			const unsigned long long same = *flag ? *flag - 1 : 0;
			const uint64_t pairs = (uint64_t)std::min<uint64_t>(PROBE_WARPS, n / 32) * 31;
			do_sort = same * 2 < pairs;
		}
		if (do_sort) {
			const int key_bits = 3 * (c->ray_key_ob + c->ray_key_db);
			uint32_t *ghist = nullptr;
			int passes = 0;
			if (int rc = radix_sort_prepare32(c, rs.scratch, n, key_bits, s, &ghist, &passes))
				return rc;
			const unsigned kgrid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 16);
			k_ray_keys<<<kgrid, 256, 0, s>>>(d_rays6, n, lo, ie, c->ray_key_ob, c->ray_key_db,
			                                 rs.keys[0].as<uint32_t>(), passes, ghist);
			c->launches += 1;
			uint32_t *const kk[2] = {rs.keys[0].as<uint32_t>(), rs.keys[1].as<uint32_t>()};
			uint32_t *const vv[2] = {rs.vals[0].as<uint32_t>(), rs.vals[1].as<uint32_t>()};
			int cur = 0;
			if (int rc = radix_sort_run32_identity(c, rs.scratch, kk, vv, n, key_bits, s, &cur))
				return rc;
			P.perm = vv[cur];
			c->sorted_batches++;
		} else {
			c->unsorted_batches++;
		}
	}

	// compressed 4-wide nodes: always (mode 1) or for batches that were just found incoherent and
	// reordered (mode 2) -- there the traversal is bound by L1 wavefronts of scattered node fetches
	// and halving the fetches per ray pays; coherent batches are issue-bound and prefer the
	// cheaper-to-decode binary nodes
	// (mode 2 builds them only for scenes of >= 2^20 triangles: measured x1.20 on 10 M triangles /
	// reordered random rays, nothing on the 262 k-triangle bounce rays, -20 % on coherent rays; the
	// instrumented count launch follows the choice made for the last traced batch)
	// (the watertight mode keeps to the binary nodes: its records hold vertices, see build.cu)
	const bool wt = c->recs_vertex_form;
	const bool aos = out.aos != nullptr;
	const bool have_wide = c->n_nodes && c->wide_built && !wt && c->fast_boxes;
	const bool wide = have_wide && (c->wide_mode == 1 ||
	                                (c->wide_mode == 2 && (d_counts ? c->last_wide : P.perm != nullptr)));
	if (!d_counts)
		c->last_wide = wide;
	if (wide) {
		c->wide_batches++;
		P.refill = c->refill_wide;
		P.leaf_votes = c->leaf_votes_wide;
	} else if (P.perm || (c->l2_bytes && (c->n_nodes + c->n_tris) * 64 * 2 > c->l2_bytes)) {
		// scattered fetches (a reordered batch, or a tree that does not sit comfortably in L2):
		// refilling less often measured better than keeping the warps fuller
		P.refill = c->refill_scattered;
	}

	if (d_counts) { // instrumented run: set-aside rays (if any) count as zero
		PRT_CUDA(c, cudaMemsetAsync(d_counts, 0, n * 8, s));
		P.slow_cap = 0;
		if (int rc = launch_kernel(c, trace_kernel_count(wide, wt), -1, mask, P, n, wide, out.slot, s,
		                           exotic_mode == EXOTIC_DEFERRED))
			return rc;
		// its set-aside counter is not consumed by an exact pass: re-arm it
		PRT_CUDA(c, cudaMemsetAsync(P.counter + 3, 0, 8, s));
		return PRT_OK;
	}
	if (!c->fast_boxes) // the reference's box arithmetic everywhere
		return launch_kernel(c, wt ? trace_kernel_wt_exact(mask, aos) : trace_kernel_exact(mask, aos),
		                     (wt ? 6 : 4) + (aos ? 1 : 0), mask, P, n, false, out.slot, s,
		                     exotic_mode == EXOTIC_DEFERRED);

	// ---- fast kernel; the rays it cannot take go to a list
	const uint64_t cap = n < (1ull << 32) ? std::min<uint64_t>(n, 1ull << 20) : 0;
	PRT_CUDA(c, c->slow_list[out.slot].reserve(std::max<uint64_t>(cap, 1) * 4));
	P.slow_list = c->slow_list[out.slot].as<uint32_t>();
	P.slow_cap = (uint32_t)cap;
	volatile unsigned long long *flag = c->probe_host + 2 + out.slot;
	if (exotic_mode == EXOTIC_DEFERRED) {
		*flag = 0;
		P.slow_host = c->probe_dev + 2 + out.slot;
	}
	// bps_cache columns: 0/1 fast SoA/AoS, 2/3 wide, 4/5 exact, 6/7 watertight exact; the watertight
	// fast kernels share 2/3 (a watertight scene never uses the wide nodes)
	KernelFn fn = wt ? trace_kernel_wt(mask, aos)
	                 : (wide ? trace_kernel_wide(mask, aos) : trace_kernel_fast(mask, aos));
	if (int rc = launch_kernel(c, fn, (aos ? 1 : 0) + ((wt || wide) ? 2 : 0), mask, P, n, wide,
	                           out.slot, s, exotic_mode == EXOTIC_DEFERRED, !wide))
		return rc;
	// ---- exact pass over the set-aside rays
	P.slow_count = P.counter + 3;
	P.slow_host = nullptr;
	static_assert(sizeof(TraceParams) <= sizeof(c->exotic_blob), "exotic_blob too small");
	std::memcpy(c->exotic_blob, &P, sizeof P);
	const KernelFn exact_fn = wt ? trace_kernel_wt_exact(mask, aos) : trace_kernel_exact(mask, aos);
	c->exotic_fn = reinterpret_cast<const void *>(exact_fn);
	c->exotic_cache_slot = (wt ? 6 : 4) + (aos ? 1 : 0);
	c->exotic_mask = mask;
	c->exotic_slot = out.slot;
	if (exotic_mode == EXOTIC_DEFERRED) {
		c->pending_exotic = true;
		return PRT_OK;
	}
	return launch_kernel(c, exact_fn, c->exotic_cache_slot, mask, P, n, false, out.slot, s);
}

// After the stream of a EXOTIC_DEFERRED launch has been synchronised: trace the set-aside rays, if
// the fast kernel reported any.  *ran says whether a kernel was launched (the caller then
// synchronises again).
int finish_exotic(prt_b200 *c, cudaStream_t s, bool *ran) {
	*ran = false;
	if (!c->pending_exotic)
		return PRT_OK;
	c->pending_exotic = false;
	volatile unsigned long long *flag = c->probe_host + 2 + c->exotic_slot;
	const unsigned long long set_aside = *flag ? *flag - 1 : 0;
	if (set_aside == 0)
		return PRT_OK;
	*ran = true;
	c->exotic_rays += set_aside;
	TraceParams P;
	std::memcpy(&P, c->exotic_blob, sizeof P);
	return launch_kernel(c, reinterpret_cast<KernelFn>(const_cast<void *>(c->exotic_fn)),
	                     c->exotic_cache_slot, c->exotic_mask, P, P.n_rays, false, c->exotic_slot, s);
}

// ------------------------------------------------------------------------------------------------
// read-bandwidth probe (roofline denominators)
__global__ void __launch_bounds__(256) k_read_probe(const uint4 *__restrict__ p, uint64_t n16,
                                                    int iters, uint32_t *sink) {
	uint32_t acc = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (int it = 0; it < iters; ++it) {
		for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
			const uint4 v = __ldcg(p + i); // L2 (not L1) path
			acc += v.x ^ v.y ^ v.z ^ v.w;
		}
	}
	if (acc == 0x12345678u)
		*sink = acc; // never true in practice; keeps the loads alive
}

int launch_read_probe(prt_b200 *c, const void *buf, uint64_t bytes, int iters, float *ms) {
	PRT_CUDA(c, c->counter.reserve(256));
	PRT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
	k_read_probe<<<c->sm_count * 8, 256, 0, c->stream>>>(static_cast<const uint4 *>(buf), bytes / 16,
	                                                     iters, c->counter.as<uint32_t>() + 32);
	c->launches += 1;
	PRT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
	PRT_CUDA(c, cudaStreamSynchronize(c->stream));
	float t = 0.f;
	PRT_CUDA(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
	if (ms)
		*ms = t;
	return PRT_OK;
}

} // namespace prt
