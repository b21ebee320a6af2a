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
#include "prt_ctx.h"
#include "prt_traverse.cuh"

namespace prt {

constexpr int TRACE_THREADS = 128;

struct TraceParams {
	const Node *nodes;
	const TriRec *tris;
	const float *rays;
	uint64_t n_rays;
	uint64_t n_tris;
	// SoA outputs
	float2 *uv;
	float *t;
	uint32_t *pid;
	float *p;
	uint8_t *valid;
	// AoS output
	char *aos;
	prt_hit_layout lay;
	uint32_t *counts;
	unsigned long long *counter;
	int prune;
	float slack_rel, slack_ulps;
};

template <uint32_t MASK, bool AOS, bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(const TraceParams P) {
	constexpr bool ANYHIT = (MASK == PRT_TAG_VALID) && !COUNT;
	constexpr bool WANT_UV = (MASK & PRT_TAG_UV) != 0;
	constexpr bool TRACK_PRIM = (MASK & (PRT_TAG_UV | PRT_TAG_PID)) != 0;

	const int lane = threadIdx.x & 31;
	TraverseOpts opts;
	opts.prune = P.prune;
	opts.slack_rel = P.slack_rel;
	opts.slack_ulps = P.slack_ulps;

	for (;;) {
		unsigned long long base = 0;
		if (lane == 0)
			base = atomicAdd(P.counter, 32ull);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= P.n_rays)
			break;
		const uint64_t i = base + lane;
		if (i < P.n_rays) {
			float r6[6];
#pragma unroll
			for (int k = 0; k < 6; ++k)
				r6[k] = __ldg(P.rays + i * 6 + k);
			const RayC r = make_ray(r6);
			Hit h;
			traverse<ANYHIT, WANT_UV, TRACK_PRIM, COUNT>(P.nodes, P.tris, P.n_tris, r, opts, h);

			// epilogue, bvh.hpp:259-263
			const bool valid = h.t < INFINITY;
			const float px = fadd(r.o[0], fmul(h.t, r.d[0]));
			const float py = fadd(r.o[1], fmul(h.t, r.d[1]));
			const float pz = fadd(r.o[2], fmul(h.t, r.d[2]));
			if (COUNT) {
				P.counts[2 * i] = h.n_nodes;
				P.counts[2 * i + 1] = h.n_tris;
			} else if (AOS) {
				char *rec = P.aos + i * P.lay.stride;
				if (MASK & PRT_TAG_UV) {
					*reinterpret_cast<float *>(rec + P.lay.off_u) = h.u;
					*reinterpret_cast<float *>(rec + P.lay.off_v) = h.v;
				}
				if (MASK & PRT_TAG_T)
					*reinterpret_cast<float *>(rec + P.lay.off_t) = h.t;
				if (MASK & PRT_TAG_PID)
					*reinterpret_cast<uint32_t *>(rec + P.lay.off_pid) = h.prim;
				if (MASK & PRT_TAG_VALID)
					*reinterpret_cast<uint8_t *>(rec + P.lay.off_valid) = valid ? 1 : 0;
				if (MASK & PRT_TAG_P) {
					*reinterpret_cast<float *>(rec + P.lay.off_px) = px;
					*reinterpret_cast<float *>(rec + P.lay.off_py) = py;
					*reinterpret_cast<float *>(rec + P.lay.off_pz) = pz;
				}
			} else {
				if (MASK & PRT_TAG_UV)
					P.uv[i] = make_float2(h.u, h.v);
				if (MASK & PRT_TAG_T)
					P.t[i] = h.t;
				if (MASK & PRT_TAG_PID)
					P.pid[i] = h.prim;
				if (MASK & PRT_TAG_VALID)
					P.valid[i] = valid ? 1 : 0;
				if (MASK & PRT_TAG_P) {
					P.p[3 * i] = px;
					P.p[3 * i + 1] = py;
					P.p[3 * i + 2] = pz;
				}
			}
		}
		__syncwarp();
	}
}

using KernelFn = void (*)(const TraceParams);

template <uint32_t M> struct Table {
	static void fill(KernelFn (*t)[2]) {
		t[M][0] = k_trace<M, false, false>;
		t[M][1] = k_trace<M, true, false>;
		Table<M - 1>::fill(t);
	}
};
template <> struct Table<0> {
	static void fill(KernelFn (*)[2]) {}
};

static KernelFn g_table[32][2];
static int g_blocks_per_sm[32][2];
static bool g_table_ready = false;

int launch_trace(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t mask, const TraceOut &out,
                 uint32_t *d_counts, cudaStream_t s) {
	if (mask == 0 || mask > PRT_TAG_ALL)
		return fail(c, PRT_E_ARG, "nearest_hits: tag mask must be in 1..31");
	if (n == 0)
		return PRT_OK;
	if (!g_table_ready) {
		Table<31>::fill(g_table);
		for (int m = 1; m < 32; ++m)
			for (int a = 0; a < 2; ++a)
				g_blocks_per_sm[m][a] = 0;
		g_table_ready = true;
	}
	TraceParams P{};
	P.nodes = c->nodes.as<Node>();
	P.tris = c->trirecs.as<TriRec>();
	P.rays = d_rays6;
	P.n_rays = n;
	P.n_tris = c->n_tris;
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
	PRT_CUDA(c, c->counter.reserve(256));
	P.counter = c->counter.as<unsigned long long>();
	PRT_CUDA(c, cudaMemsetAsync(P.counter, 0, 8, s));

	const bool aos = out.aos != nullptr;
	KernelFn fn = d_counts ? (KernelFn)k_trace<PRT_TAG_ALL, false, true> : g_table[mask][aos];
	int bps = d_counts ? 0 : g_blocks_per_sm[mask][aos];
	if (bps == 0) {
		PRT_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, TRACE_THREADS, 0));
		if (bps < 1)
			bps = 1;
		if (!d_counts)
			g_blocks_per_sm[mask][aos] = bps;
	}
	uint64_t want = (n + TRACE_THREADS - 1) / TRACE_THREADS;
	uint64_t grid = (uint64_t)c->sm_count * bps;
	if (want < grid)
		grid = want;
	fn<<<(unsigned)grid, TRACE_THREADS, 0, s>>>(P);
	c->launches += 1;
	PRT_CUDA(c, cudaGetLastError());
	return PRT_OK;
}

} // namespace prt
