// prt_trace_kernel.cuh -- the persistent-warp traversal kernel template (sm_100a), instantiated by
// trace.cu (reference arithmetic: 31 tag masks x {SoA, AoS} x {binary, 4-wide nodes}) and by
// trace_wt.cu (opt-in watertight triangle test: 31 tag masks x {SoA, AoS}).
#pragma once

#include "prt_ctx.h"
#include "prt_traverse.cuh"

namespace prt {

constexpr int TRACE_THREADS = 128;

struct TraceParams {
	const Node *nodes;
	const Node4 *nodes4;
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
	float scene_absmax[3];
	int fast; // 1: conservative FFMA test for internal culling where the ray qualifies
	int refill; // re-fetch rays when fewer than this many lanes of a warp are still traversing
	int32_t root;         // index of the root node
	const uint32_t *perm; // ray processing order (reordered batches) or nullptr = identity
};

// Writes one finished ray (epilogue of bvh.hpp:259-263).
template <uint32_t MASK, bool AOS, bool COUNT>
__device__ __forceinline__ void write_hit(const TraceParams &P, uint64_t i, const RayC &r,
                                          const TravState &s) {
	const bool valid = s.t_best < INFINITY;
	const float px = fadd(r.o[0], fmul(s.t_best, r.d[0]));
	const float py = fadd(r.o[1], fmul(s.t_best, r.d[1]));
	const float pz = fadd(r.o[2], fmul(s.t_best, r.d[2]));
	if (COUNT) {
		P.counts[2 * i] = s.n_nodes;
		P.counts[2 * i + 1] = s.n_tris;
	} else if (AOS) {
		char *rec = P.aos + i * P.lay.stride;
		if (MASK & PRT_TAG_UV) {
			*reinterpret_cast<float *>(rec + P.lay.off_u) = s.u_best;
			*reinterpret_cast<float *>(rec + P.lay.off_v) = s.v_best;
		}
		if (MASK & PRT_TAG_T)
			*reinterpret_cast<float *>(rec + P.lay.off_t) = s.t_best;
		if (MASK & PRT_TAG_PID)
			*reinterpret_cast<uint32_t *>(rec + P.lay.off_pid) = s.prim_best;
		if (MASK & PRT_TAG_VALID)
			*reinterpret_cast<uint8_t *>(rec + P.lay.off_valid) = valid ? 1 : 0;
		if (MASK & PRT_TAG_P) {
			*reinterpret_cast<float *>(rec + P.lay.off_px) = px;
			*reinterpret_cast<float *>(rec + P.lay.off_py) = py;
			*reinterpret_cast<float *>(rec + P.lay.off_pz) = pz;
		}
	} else {
		if (MASK & PRT_TAG_UV)
			P.uv[i] = make_float2(s.u_best, s.v_best);
		if (MASK & PRT_TAG_T)
			P.t[i] = s.t_best;
		if (MASK & PRT_TAG_PID)
			P.pid[i] = s.prim_best;
		if (MASK & PRT_TAG_VALID)
			P.valid[i] = valid ? 1 : 0;
		if (MASK & PRT_TAG_P) {
			P.p[3 * i] = px;
			P.p[3 * i + 1] = py;
			P.p[3 * i + 2] = pz;
		}
	}
}

// Persistent warps with dynamic ray fetch: a lane whose ray has finished writes its hit and, as
// soon as fewer than REFILL lanes of the warp are still traversing, all idle lanes pull new rays
// from the global counter with one aggregated atomic.  Rays of very different length (a miss ends
// after a node or two, a hit after dozens) therefore do not leave most of the warp idle.
// (threshold P.refill, env PRT_B200_REFILL; 0 = classic "whole warp finishes, then fetch 32")

template <uint32_t MASK, bool AOS, bool COUNT, bool WIDE, bool WT = false>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(const TraceParams P) {
	constexpr bool ANYHIT = (MASK == PRT_TAG_VALID) && !COUNT;
	constexpr bool WANT_UV = (MASK & PRT_TAG_UV) != 0;
	constexpr bool TRACK_PRIM = (MASK & (PRT_TAG_UV | PRT_TAG_PID)) != 0;

	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	TraverseOpts opts;
	opts.prune = P.prune;
	opts.slack_rel = P.slack_rel;
	opts.slack_ulps = P.slack_ulps;

	StackEntry stack[STACK_DEPTH];
	TravState s;
	RayC r;
	FastRay fr;
	WoopRay wr{};
	uint64_t ray = 0;
	bool has_ray = false;
	bool fast = false;
	bool exhausted = false; // warp-uniform: the global counter ran past the last ray
	s.cur = PRT_DONE;

	for (;;) {
		// ---- fetch rays for the idle lanes (one atomic per warp)
		const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
		if (idle && !exhausted) {
			const int n = __popc(idle);
			const int leader = __ffs(idle) - 1;
			unsigned long long base = 0;
			if ((int)lane == leader)
				base = atomicAdd(P.counter, (unsigned long long)n);
			base = __shfl_sync(0xffffffffu, base, leader);
			if (base + n >= P.n_rays)
				exhausted = true;
			if (!has_ray) {
				uint64_t i = base + __popc(idle & lt);
				if (i < P.n_rays) {
					if (P.perm)
						i = __ldg(P.perm + i);
					float r6[6];
#pragma unroll
					for (int k = 0; k < 6; ++k)
						r6[k] = __ldg(P.rays + i * 6 + k);
					r = make_ray(r6);
					fr = make_fast_ray(r, P.scene_absmax);
					fast = P.fast && fr.ok;
					if (WT)
						wr = make_woop_ray(r);
					trav_init(s, r, opts, P.n_tris, P.root);
					ray = i;
					has_ray = true;
				}
			}
		}
		if (!__any_sync(0xffffffffu, has_ray))
			break;

		// ---- traverse until this lane's ray is finished or too few lanes are still busy
		if (has_ray) {
			while (s.cur != PRT_DONE) {
				if (fast)
					trav_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, true, WIDE, WT>(
					    s, stack, P.nodes, P.tris, r, fr, opts, P.nodes4, &wr);
				else
					trav_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, false, false, WT>(
					    s, stack, P.nodes, P.tris, r, fr, opts, nullptr, &wr);
				if (!exhausted && __popc(__activemask()) < P.refill)
					break;
			}
			if (s.cur == PRT_DONE) {
				write_hit<MASK, AOS, COUNT>(P, ray, r, s);
				has_ray = false;
			}
		}
		__syncwarp();
	}
	// The last warp to leave re-arms the counters for the next launch (no cudaMemset between
	// launches: inside the host pipeline a 16-byte memset queues behind megabytes of DMA).  A warp
	// gets here only after its last fetch, so nobody touches the ray counter any more.
	if (lane == 0) {
		const unsigned long long warps = (unsigned long long)gridDim.x * (TRACE_THREADS / 32);
		if (atomicAdd(P.counter + 2, 1ull) == warps - 1) {
			P.counter[0] = 0;
			P.counter[2] = 0;
		}
	}
}

using KernelFn = void (*)(const TraceParams);

// trace_wt.cu
KernelFn trace_kernel_wt(uint32_t mask, bool aos, bool count);

} // namespace prt
