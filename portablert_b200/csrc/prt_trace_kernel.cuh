// prt_trace_kernel.cuh -- the persistent-warp traversal kernel template (sm_100a), instantiated by
// trace.cu (reference arithmetic: 31 tag masks x {SoA, AoS} x {binary, 4-wide nodes}, plus the
// exact kernel) and by trace_wt.cu (opt-in watertight triangle test).
#pragma once

#include "prt_ctx.h"
#include "prt_traverse.cuh"

namespace prt {

constexpr int TRACE_THREADS = 128;
// entries of the per-thread traversal stack kept in shared memory ([depth][thread]: conflict-free
// whatever depth each lane is at); deeper entries overflow to a global buffer (rare)
#ifndef PRT_SMEM_STACK
#define PRT_SMEM_STACK 12
#endif
constexpr int SMEM_STACK = PRT_SMEM_STACK;
#ifndef PRT_MIN_BLOCKS
#define PRT_MIN_BLOCKS 8
#endif

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
	unsigned long long *counter; // [0] ray counter, [1] coherence probe, [2] warps that left, [3] rays set aside
	int prune;
	float slack_rel, slack_ulps;
	float scene_absmax[3];
	int refill;     // re-fetch rays when fewer than this many lanes of a warp are still traversing
	int leaf_votes; // a leaf phase starts when leaf_votes/32 of the busy lanes wait at a triangle
	int prefetch;   // 1: L2-prefetch nodes as they are pushed (scenes larger than L2)
	int rays_vec;   // the ray buffer is 8-byte aligned: three 8-byte loads per ray
	int32_t root;         // index of the root node
	const uint32_t *perm; // ray processing order (reordered batches) or nullptr = identity
	uint2 *stack_ovf;     // [depth - SMEM_STACK][thread of the grid]
	uint32_t ovf_stride;  // threads of the grid
	// rays the fast box test does not apply to (non-finite / overflowing): set aside by the fast
	// kernel, traced by the exact kernel (which gets the list, or -- list overflown -- every ray)
	uint32_t *slow_list;
	uint32_t slow_cap;
	const unsigned long long *slow_count;      // exact kernel as second pass: how many were set aside
	volatile unsigned long long *slow_host;    // mapped host word: count + 1 once the fast kernel is done
};

// Per-thread stack: SMEM_STACK entries in shared memory, the rest in global memory.
struct DevStack {
	uint2 *sm;  // &block_stack[threadIdx.x], entry k at sm[k * TRACE_THREADS]
	uint2 *ovf; // &overflow[global thread], entry k at ovf[k * stride]
	uint32_t stride;
	const char *pf_base; // nodes are prefetched into L2 as they are pushed (nullptr: off)
	int sp;
	__device__ __forceinline__ void push(uint32_t node, uint32_t tmin_bits) {
		if (sp < SMEM_STACK)
			sm[sp * TRACE_THREADS] = make_uint2(node, tmin_bits);
		else
			ovf[(size_t)(sp - SMEM_STACK) * stride] = make_uint2(node, tmin_bits);
		++sp;
		if (pf_base && (int32_t)node >= 0)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_base + (size_t)node * 64));
	}
	__device__ __forceinline__ void pop(uint32_t &node, uint32_t &tmin_bits) {
		--sp;
		const uint2 e = sp < SMEM_STACK ? sm[sp * TRACE_THREADS] : ovf[(size_t)(sp - SMEM_STACK) * stride];
		node = e.x;
		tmin_bits = e.y;
	}
};

// Writes one finished ray (epilogue of bvh.hpp:259-263); streaming stores, the records are not
// read again on the device.
template <uint32_t MASK, bool AOS, bool COUNT>
__device__ __forceinline__ void write_hit(const TraceParams &P, uint64_t i, const RayC &r,
                                          const TravState &s) {
	const bool valid = s.t_best < INFINITY;
	const float px = fadd(r.o[0], fmul(s.t_best, r.d[0]));
	const float py = fadd(r.o[1], fmul(s.t_best, r.d[1]));
	const float pz = fadd(r.o[2], fmul(s.t_best, r.d[2]));
	if (COUNT) {
		reinterpret_cast<uint2 *>(P.counts)[i] = make_uint2(s.n_nodes, s.n_tris);
	} else if (AOS) {
		char *rec = P.aos + i * P.lay.stride;
		if (MASK & PRT_TAG_UV) {
			__stcs(reinterpret_cast<float *>(rec + P.lay.off_u), s.u_best);
			__stcs(reinterpret_cast<float *>(rec + P.lay.off_v), s.v_best);
		}
		if (MASK & PRT_TAG_T)
			__stcs(reinterpret_cast<float *>(rec + P.lay.off_t), s.t_best);
		if (MASK & PRT_TAG_PID)
			__stcs(reinterpret_cast<uint32_t *>(rec + P.lay.off_pid), s.prim_best);
		if (MASK & PRT_TAG_VALID)
			*reinterpret_cast<uint8_t *>(rec + P.lay.off_valid) = valid ? 1 : 0;
		if (MASK & PRT_TAG_P) {
			__stcs(reinterpret_cast<float *>(rec + P.lay.off_px), px);
			__stcs(reinterpret_cast<float *>(rec + P.lay.off_py), py);
			__stcs(reinterpret_cast<float *>(rec + P.lay.off_pz), pz);
		}
	} else {
		if (MASK & PRT_TAG_UV)
			__stcs(P.uv + i, make_float2(s.u_best, s.v_best));
		if (MASK & PRT_TAG_T)
			__stcs(P.t + i, s.t_best);
		if (MASK & PRT_TAG_PID)
			__stcs(P.pid + i, s.prim_best);
		if (MASK & PRT_TAG_VALID)
			P.valid[i] = valid ? 1 : 0;
		if (MASK & PRT_TAG_P) {
			__stcs(P.p + 3 * i, px);
			__stcs(P.p + 3 * i + 1, py);
			__stcs(P.p + 3 * i + 2, pz);
		}
	}
}

// Persistent warps, one ray per lane, with three measures against SIMT divergence:
//   * dynamic ray fetch: finished lanes write their record and, as soon as fewer than `refill`
//     lanes of the warp are still traversing, all idle lanes pull new rays from the global counter
//     with one aggregated atomic -- rays of very different length (a miss ends after a node or two,
//     a hit after dozens) do not leave most of the warp idle;
//   * node phase / leaf phase: a lane that reaches a triangle WAITS there (its record is
//     prefetched) while the others keep descending; the triangle test runs for all waiting lanes
//     at once as soon as leaf_votes/32 of the busy lanes wait or nobody is left at a node.  With
//     one triangle per leaf and the leaf's own box already tested in its parent, about one step in
//     ten is a triangle: tested as it comes, nearly every iteration of the warp would pay the ~80
//     instructions of Moeller-Trumbore for two or three lanes;
//   * one code path: every ray with finite components takes the conservative FFMA box test (axis-
//     parallel rays included, prt_traverse.cuh); the handful that does not is set aside in a list
//     and traced afterwards by the exact kernel (EXACT = true: the reference's box arithmetic at
//     every node), which is also what PRT_B200_FAST_BOXES=0 runs for every ray.
// The traversal stack lives in shared memory ([depth][thread], bank-conflict-free at any mix of
// depths) with a global overflow, not in local memory.
template <uint32_t MASK, bool AOS, bool COUNT, bool WIDE, bool WT, bool EXACT>
__global__ void __launch_bounds__(TRACE_THREADS, PRT_MIN_BLOCKS) k_trace(const TraceParams P) {
	constexpr bool ANYHIT = (MASK == PRT_TAG_VALID) && !COUNT;
	constexpr bool WANT_UV = (MASK & PRT_TAG_UV) != 0;
	constexpr bool TRACK_PRIM = (MASK & (PRT_TAG_UV | PRT_TAG_PID)) != 0;
	constexpr bool FAST = !EXACT;
	constexpr bool W = WIDE && FAST;
	__shared__ uint2 s_stack[SMEM_STACK * TRACE_THREADS];

	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	TraverseOpts opts;
	opts.prune = P.prune;
	opts.slack_rel = P.slack_rel;
	opts.slack_ulps = P.slack_ulps;

	uint64_t n_rays = P.n_rays;
	const uint32_t *perm = P.perm;
	if (EXACT && P.slow_count) { // second pass: the rays the fast kernel set aside
		const unsigned long long set_aside = *P.slow_count;
		if (set_aside == 0)
			return;
		if (set_aside <= P.slow_cap) {
			n_rays = set_aside;
			perm = P.slow_list;
		} // else: the list overflowed, every ray is traced again (same results)
	}

	DevStack stack;
	stack.sm = s_stack + threadIdx.x;
	stack.ovf = P.stack_ovf + ((size_t)blockIdx.x * TRACE_THREADS + threadIdx.x);
	stack.stride = P.ovf_stride;
	stack.pf_base = P.prefetch ? (W ? reinterpret_cast<const char *>(P.nodes4)
	                                : reinterpret_cast<const char *>(P.nodes))
	                           : nullptr;
	stack.sp = 0;
	TravState s;
	RayC r;
	FastRay fr;
	WoopRay wr{};
	uint64_t ray = 0;
	bool has_ray = false;
	bool exhausted = false; // warp-uniform: the global counter ran past the last ray
	s.cur = PRT_DONE;

	for (;;) {
		// ---- retire finished rays, fetch rays for the idle lanes (one atomic per warp)
		if (has_ray && s.cur == PRT_DONE) {
			write_hit<MASK, AOS, COUNT>(P, ray, r, s);
			has_ray = false;
		}
		const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
		if (idle && !exhausted) {
			const int n = __popc(idle);
			const int leader = __ffs(idle) - 1;
			unsigned long long base = 0;
			if ((int)lane == leader)
				base = atomicAdd(P.counter, (unsigned long long)n);
			base = __shfl_sync(0xffffffffu, base, leader);
			if (base + n >= n_rays)
				exhausted = true;
			if (!has_ray) {
				uint64_t i = base + __popc(idle & lt);
				if (i < n_rays) {
					if (perm)
						i = __ldg(perm + i);
					float r6[6];
					if (P.rays_vec) {
						const float2 *src = reinterpret_cast<const float2 *>(P.rays + i * 6);
						const float2 a = __ldcs(src), b = __ldcs(src + 1), c = __ldcs(src + 2);
						r6[0] = a.x, r6[1] = a.y, r6[2] = b.x, r6[3] = b.y, r6[4] = c.x, r6[5] = c.y;
					} else {
#pragma unroll
						for (int k = 0; k < 6; ++k)
							r6[k] = __ldcs(P.rays + i * 6 + k);
					}
					r = make_ray(r6);
					bool mine = true;
					if (FAST) {
						fr = make_fast_ray(r, P.scene_absmax, W);
						if (!fr.ok) { // set aside for the exact kernel
							const unsigned long long at = atomicAdd(P.counter + 3, 1ull);
							if (at < P.slow_cap)
								P.slow_list[at] = (uint32_t)i;
							mine = false;
						}
					}
					if (mine) {
						if (WT)
							wr = make_woop_ray(r);
						trav_init(s, r, opts, P.n_tris, P.root);
						stack.sp = 0;
						ray = i;
						has_ray = true;
					}
				}
			}
		}
		if (!__any_sync(0xffffffffu, has_ray)) {
			if (exhausted)
				break;
			continue;
		}

		// ---- traverse until too few lanes are still busy
		for (;;) {
			if (has_ray && at_node(s.cur)) {
				trav_node_step<COUNT, FAST, W, WT>(s, stack, P.nodes, P.nodes4, r, fr);
				if (s.cur < 0) // parked at a triangle: have its record on the way
					asm volatile("prefetch.global.L1 [%0];" ::"l"(P.tris + (uint32_t)(~s.cur)));
			}
			const bool leaf = has_ray && s.cur < 0;
			const unsigned leafm = __ballot_sync(0xffffffffu, leaf);
			const unsigned nodem = __ballot_sync(0xffffffffu, has_ray && at_node(s.cur));
			if (leafm && (nodem == 0 ||
			              __popc(leafm) * 32 >= P.leaf_votes * (__popc(leafm) + __popc(nodem)))) {
				if (leaf)
					trav_leaf_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, FAST, WT>(s, stack, P.tris, r,
					                                                             opts, &wr);
			}
			const unsigned busy = __ballot_sync(0xffffffffu, has_ray && s.cur != PRT_DONE);
			if (busy == 0 || (!exhausted && __popc(busy) < P.refill))
				break;
		}
	}
	// The last warp to leave re-arms the counters for the next launch (no cudaMemset between
	// launches: inside the host pipeline a 16-byte memset queues behind megabytes of DMA).  A warp
	// gets here only after its last fetch, so nobody touches the ray counter any more.
	if (lane == 0) {
		const unsigned long long warps = (unsigned long long)gridDim.x * (TRACE_THREADS / 32);
		__threadfence();
		if (atomicAdd(P.counter + 2, 1ull) == warps - 1) {
			__threadfence();
			P.counter[0] = 0;
			P.counter[2] = 0;
			if (FAST) {
				if (P.slow_host) {
					*P.slow_host = atomicAdd(P.counter + 3, 0ull) + 1; // +1: 0 means "not written yet"
					__threadfence_system();
				}
			} else if (P.slow_count) {
				P.counter[3] = 0;
			}
		}
	}
}

using KernelFn = void (*)(const TraceParams);

// One table of 31 masks x {SoA, AoS} per kernel family; every family is instantiated in its own
// translation unit (trace.cu, trace_wide.cu, trace_exact.cu, trace_wt.cu) so that they compile in
// parallel.
template <bool WIDE, bool WT, bool EXACT, uint32_t M> struct KernelTable {
	static void fill(KernelFn (*t)[2]) {
		t[M][0] = k_trace<M, false, false, WIDE, WT, EXACT>;
		t[M][1] = k_trace<M, true, false, WIDE, WT, EXACT>;
		KernelTable<WIDE, WT, EXACT, M - 1>::fill(t);
	}
};
template <bool WIDE, bool WT, bool EXACT> struct KernelTable<WIDE, WT, EXACT, 0> {
	static void fill(KernelFn (*)[2]) {}
};
template <bool WIDE, bool WT, bool EXACT> KernelFn kernel_of(uint32_t mask, bool aos) {
	struct Filled {
		KernelFn t[32][2];
		Filled() { KernelTable<WIDE, WT, EXACT, 31>::fill(t); }
	};
	static const Filled table; // thread-safe initialisation (contexts may trace from several threads)
	return table.t[mask][aos ? 1 : 0];
}

enum KernelFamily { KF_FAST = 0, KF_WIDE = 1, KF_EXACT = 2, KF_WT = 3, KF_WT_EXACT = 4, KF_COUNT = 5 };
KernelFn trace_kernel_fast(uint32_t mask, bool aos);     // trace.cu
KernelFn trace_kernel_wide(uint32_t mask, bool aos);     // trace_wide.cu
KernelFn trace_kernel_exact(uint32_t mask, bool aos);    // trace_exact.cu
KernelFn trace_kernel_wt(uint32_t mask, bool aos);       // trace_wt.cu
KernelFn trace_kernel_wt_exact(uint32_t mask, bool aos); // trace_wt.cu
KernelFn trace_kernel_count(bool wide, bool wt);         // trace_exact.cu (instrumented runs)

} // namespace prt
