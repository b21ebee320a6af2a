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
#ifndef PRT_BURST_UNROLL
#define PRT_BURST_UNROLL 4
#endif
#ifndef PRT_SMEM_STACK
#define PRT_SMEM_STACK 12
#endif
constexpr int SMEM_STACK = PRT_SMEM_STACK;
constexpr int BURST_UNROLL = PRT_BURST_UNROLL;
#ifndef PRT_MIN_BLOCKS
#define PRT_MIN_BLOCKS 8
#endif
// node steps between two votes of the warp (leaf phase? enough lanes busy?)
#ifndef PRT_NODE_BURST
#define PRT_NODE_BURST 4
#endif
#ifndef PRT_LOCAL_STACK
#define PRT_LOCAL_STACK 0
#endif
#ifndef PRT_LEAF_PREFETCH
#define PRT_LEAF_PREFETCH 0
#endif
#ifndef PRT_SIMPLE_LOOP
#define PRT_SIMPLE_LOOP 0
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
	// cooperative tail (k_coop below): rays a warp is still tracing `coop_after` iterations after the
	// end of the batch are handed over -- state and pending stack -- to a follow-up kernel that
	// gives every such ray a whole warp (0 = never)
	uint32_t *coop;          // [0] handed over, [1] taken, then coop_cap records of coop_rec words
	uint32_t coop_cap;       // records available (a full list just means: keep tracing here)
	uint32_t coop_rec;       // words per record: COOP_PARK + 2 * coop_depth
	uint32_t coop_depth;     // stack entries a record can hold
	uint2 *coop_lifo;        // k_coop: LIFO area, coop_lifo_cap entries per warp of its grid
	uint32_t coop_lifo_cap;
	int coop_after;
	int coop_min_sp; // hand over only rays with at least this many stacked subtrees
};

constexpr int COOP_PARK = 16; // words of a parked ray state

// Per-thread stack: SMEM_STACK entries in shared memory, the rest in global memory.
// (An L2 prefetch of every pushed node -- most pushed subtrees of an incoherent ray are visited
// later -- cost six instructions of address arithmetic per entry in divergent code; removing it
// made the C4 kernel 11 % faster.)
struct DevStack {
	uint2 *sm;  // &block_stack[threadIdx.x], entry k at sm[k * TRACE_THREADS]
	uint2 *ovf; // &overflow[global thread], entry k at ovf[k * stride]
	uint32_t stride;
	int sp;
	__device__ __forceinline__ void push(uint32_t node, uint32_t tmin_bits) {
		if (sp < SMEM_STACK)
			sm[sp * TRACE_THREADS] = make_uint2(node, tmin_bits);
		else
			ovf[(size_t)(sp - SMEM_STACK) * stride] = make_uint2(node, tmin_bits);
		++sp;
	}
	__device__ __forceinline__ void pop(uint32_t &node, uint32_t &tmin_bits) {
		--sp;
		const uint2 e = sp < SMEM_STACK ? sm[sp * TRACE_THREADS] : ovf[(size_t)(sp - SMEM_STACK) * stride];
		node = e.x;
		tmin_bits = e.y;
	}
	// The next entry the best hit so far does not rule out, or PRT_DONE.  Entries in the overflow
	// first (rare), then a loop over the shared part alone: no per-iteration "where does it live".
	__device__ __forceinline__ uint32_t pop_live(float limit) {
		while (sp > SMEM_STACK) {
			--sp;
			const uint2 e = ovf[(size_t)(sp - SMEM_STACK) * stride];
			if (!(__uint_as_float(e.y) > limit))
				return e.x;
		}
		const uint2 *p = sm + sp * TRACE_THREADS;
		while (sp > 0) {
			--sp;
			p -= TRACE_THREADS;
			const uint2 e = *p;
			if (!(__uint_as_float(e.y) > limit))
				return e.x;
		}
		return (uint32_t)PRT_DONE;
	}
	__device__ __forceinline__ uint2 peek(int k) const { // entry k from the bottom
		return k < SMEM_STACK ? sm[k * TRACE_THREADS] : ovf[(size_t)(k - SMEM_STACK) * stride];
	}
	__device__ __forceinline__ bool room(int k) const { return sp + k <= SMEM_STACK; }
	__device__ __forceinline__ void put(int at, bool pred, uint32_t node, uint32_t tmin_bits) {
		if (pred) // (a predicated store: no divergent region per entry)
			sm[at * TRACE_THREADS] = make_uint2(node, tmin_bits);
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

// Cooperative tail.  When the batch is exhausted, a kernel's remaining time is the latency of its
// longest rays: a lone lane advances one tree element per ~500 cycles of dependent instructions,
// and a ray through a vertex shared by a fan of 186 triangles needs hundreds of them while the
// other 31 lanes -- and soon the whole GPU -- idle (measured on C2: 0.18 ms of a 0.37 ms launch,
// whatever the batch size).  A warp that is still busy `coop_after` iterations after the end of
// the batch therefore hands its unfinished rays over (state + pending stack, one record each) and
// leaves; k_coop, launched right behind, finishes every such ray WITH ALL 32 LANES OF A WARP: the
// ray's pending subtrees become a LIFO; every round each lane takes one entry off the top, tests
// the two children of a node or the triangle of a leaf, and the surviving children are appended
// with ballot-derived positions; the pruning limit is shared by a warp minimum whenever a lane
// finds a closer hit, and the lanes' best hits are merged at the end.  The result is the minimum
// over all accepted triangles under the order-independent tie rule (lowest primitive id), exactly
// as in the sequential traversal -- the visit order differs, nothing else.  LIFO size: a round
// removes <= 32 entries and adds <= 64 one level deeper, so it holds at most 32 entries per level.
// The last <= 64 entries pushed also sit in a shared-memory window, so a round whose entries were
// all pushed by the round before never waits for L2.
constexpr int COOP_THREADS = 128;

template <uint32_t MASK, bool AOS, bool WT>
__global__ void __launch_bounds__(COOP_THREADS) k_coop(const TraceParams P) {
	constexpr bool ANYHIT = (MASK == PRT_TAG_VALID);
	constexpr bool WANT_UV = (MASK & PRT_TAG_UV) != 0;
	constexpr bool TRACK_PRIM = (MASK & (PRT_TAG_UV | PRT_TAG_PID)) != 0;
	__shared__ uint2 s_xb[64 * (COOP_THREADS / 32)];
	__shared__ float s_idir[3 * COOP_THREADS];
	const unsigned lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	uint2 *xb = s_xb + 64 * (threadIdx.x >> 5);
	const uint32_t handed = min(P.coop[0], P.coop_cap);
	const size_t warp_global = ((size_t)blockIdx.x * COOP_THREADS + threadIdx.x) >> 5;
	uint2 *q = P.coop_lifo + warp_global * P.coop_lifo_cap;
	TraverseOpts opts;
	opts.prune = P.prune;
	opts.slack_rel = P.slack_rel;
	opts.slack_ulps = P.slack_ulps;
	TravState s;
	RayC r;
	FastRay fr;
	WoopRay wr{};
	for (;;) {
		uint32_t w = 0;
		if (lane == 0)
			w = atomicAdd(P.coop + 1, 1u);
		w = __shfl_sync(0xffffffffu, w, 0);
		if (w >= handed)
			break;
		const uint32_t *p = P.coop + 4 + (size_t)w * P.coop_rec;
		const uint64_t ray_b = (uint64_t)__ldcg(p) | ((uint64_t)__ldcg(p + 1) << 32);
		float r6[6];
#pragma unroll
		for (int k = 0; k < 6; ++k)
			r6[k] = u2f(__ldcg(p + 2 + k));
		r = make_ray(r6);
		fr = make_fast_ray(r, P.scene_absmax, false);
		if (WT)
			wr = make_woop_ray(r);
		trav_init(s, r, opts, P.n_tris, P.root);
		s.t_best = u2f(__ldcg(p + 8));
		s.u_best = u2f(__ldcg(p + 9));
		s.v_best = u2f(__ldcg(p + 10));
		s.prim_best = __ldcg(p + 11);
		if (opts.prune && s.t_best < INFINITY)
			s.limit = fadd(s.t_best, fadd(fmul(fabsf(s.t_best), opts.slack_rel), s.slack_abs));
		const int32_t cur_b = (int32_t)__ldcg(p + 12);
		const uint32_t sp_b = __ldcg(p + 13);
#pragma unroll
		for (int a = 0; a < 3; ++a)
			s_idir[a * COOP_THREADS + threadIdx.x] = r.idir[a];
		// seed the LIFO: the ray's stack, bottom first, then its current element on top
		const uint2 *st = reinterpret_cast<const uint2 *>(p + COOP_PARK);
		for (uint32_t k = lane; k < sp_b; k += 32)
			__stcg(q + k, __ldcg(st + k));
		uint32_t qn = sp_b + 1u;
		if (lane == 0) {
			const uint2 top = make_uint2((uint32_t)cur_b, 0xff800000u /* entry distance -inf */);
			__stcg(q + sp_b, top);
			xb[0] = top;
		}
		uint32_t xb_base = sp_b, xb_cnt = 1; // window of the entries pushed last
		__syncwarp();
		bool done = false;
		while (qn > 0 && !done) {
			const uint32_t take = qn < 32u ? qn : 32u;
			int32_t item = PRT_DONE;
			if (lane < take) {
				const uint32_t at = qn - 1 - lane;
				const uint2 e = (at >= xb_base && at < xb_base + xb_cnt) ? xb[at - xb_base] : __ldcg(q + at);
				if (!(u2f(e.y) > s.limit))
					item = (int32_t)e.x;
			}
			qn -= take;
			int32_t c0 = PRT_DONE, c1 = PRT_DONE;
			float tm0 = 0.f, tm1 = 0.f;
			bool found = false;
			if (at_node(item)) {
				const char *np = reinterpret_cast<const char *>(P.nodes + item);
				Vec4 a, bb, c, d;
				ld32(np, a, bb);
				ld32(np + 32, c, d);
				const float lo0[3] = {a.x, a.y, a.z}, hi0[3] = {a.w, bb.x, bb.y};
				const float lo1[3] = {bb.z, bb.w, c.x}, hi1[3] = {c.y, c.z, c.w};
				if (slab_fast(fr, lo0, hi0, s.limit, tm0))
					c0 = (int32_t)f2u(d.x);
				if (slab_fast(fr, lo1, hi1, s.limit, tm1))
					c1 = (int32_t)f2u(d.y);
				if (c0 != PRT_DONE && c1 != PRT_DONE && tm0 < tm1) { // nearer child on top
					const int32_t ti = c0;
					c0 = c1;
					c1 = ti;
					const float tf = tm0;
					tm0 = tm1;
					tm1 = tf;
				}
			} else if (item < 0) {
				const float before = s.limit;
				trav_leaf_test<ANYHIT, WANT_UV, TRACK_PRIM, false, true, WT>(
				    s, item, P.tris, r, opts, &wr, s_idir + threadIdx.x, COOP_THREADS);
				found = s.limit < before;
			}
			if (ANYHIT)
				done = __any_sync(0xffffffffu, s.t_best < INFINITY);
			if (__any_sync(0xffffffffu, found)) { // share the pruning limit
				float l = s.limit;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
				s.limit = l;
			}
			const unsigned m0 = __ballot_sync(0xffffffffu, c0 != PRT_DONE);
			const unsigned m1 = __ballot_sync(0xffffffffu, c1 != PRT_DONE);
			const uint32_t n0 = __popc(m0), n1 = __popc(m1);
			if (qn + n0 + n1 > P.coop_lifo_cap)
				__trap(); // cannot happen (see the bound above); never corrupt memory silently
			if (c0 != PRT_DONE) {
				const uint32_t k = __popc(m0 & lt);
				const uint2 e = make_uint2((uint32_t)c0, f2u(tm0));
				xb[k] = e;
				__stcg(q + qn + k, e);
			}
			if (c1 != PRT_DONE) {
				const uint32_t k = n0 + __popc(m1 & lt);
				const uint2 e = make_uint2((uint32_t)c1, f2u(tm1));
				xb[k] = e;
				__stcg(q + qn + k, e);
			}
			xb_base = qn;
			xb_cnt = n0 + n1;
			qn += n0 + n1;
			__syncwarp();
		}
		// merge the lanes' best hits (same order-independent rule as everywhere else)
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const float ot = __shfl_xor_sync(0xffffffffu, s.t_best, o);
			const uint32_t op = __shfl_xor_sync(0xffffffffu, s.prim_best, o);
			const float ou = __shfl_xor_sync(0xffffffffu, s.u_best, o);
			const float ov = __shfl_xor_sync(0xffffffffu, s.v_best, o);
			const bool take = TRACK_PRIM ? closer(ot, op, s.t_best, s.prim_best) : (ot < s.t_best);
			if (take) {
				s.t_best = ot;
				s.prim_best = op;
				s.u_best = ou;
				s.v_best = ov;
			}
		}
		if (lane == 0)
			write_hit<MASK, AOS, false>(P, ray_b, r, s);
		__syncwarp();
	}
	// the last warp to leave re-arms the list for the next launch
	if (lane == 0) {
		const unsigned warps = gridDim.x * (COOP_THREADS / 32);
		__threadfence();
		if (atomicAdd(P.coop + 2, 1u) == warps - 1) {
			__threadfence();
			P.coop[0] = 0;
			P.coop[1] = 0;
			P.coop[2] = 0;
		}
	}
}

// Persistent warps, one ray per lane, with three measures against SIMT divergence:
//   * dynamic ray fetch: finished lanes write their record and, as soon as fewer than `refill`
//     lanes of the warp are still traversing, all idle lanes pull new rays from the global counter
//     with one aggregated atomic -- rays of very different length (a miss ends after a node or two,
//     a hit after dozens) do not leave most of the warp idle;
//   * node phase / leaf phase: a lane that reaches a triangle WAITS there while the others keep
//     descending; the triangle test runs for all waiting lanes at once as soon as leaf_votes/32 of the busy lanes wait or nobody is left at a node.  With
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
	constexpr bool COOP = FAST && !W && !COUNT; // the cooperative tail serves the binary fast kernels
	__shared__ uint2 s_stack[SMEM_STACK * TRACE_THREADS];
	__shared__ float s_idir[FAST ? 3 * TRACE_THREADS : 1]; // 1.0f / d of the lane's ray (fast kernels)

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

#if PRT_LOCAL_STACK
	ArrayStack stack; // (A/B variant: the round-1 stack in local memory)
#else
	DevStack stack;
	stack.sm = s_stack + threadIdx.x;
	stack.ovf = P.stack_ovf + ((size_t)blockIdx.x * TRACE_THREADS + threadIdx.x);
	stack.stride = P.ovf_stride;
	stack.sp = 0;
#endif
	const float *my_idir = FAST ? s_idir + threadIdx.x : nullptr; // (the exact kernels do not need it)
	constexpr int idir_stride = TRACE_THREADS;
	TravState s;
	RayC r;
	FastRay fr;
	WoopRay wr{};
	uint64_t ray = 0;
	bool has_ray = false;
	bool exhausted = false; // warp-uniform: the global counter ran past the last ray
	int tail_iters = 0;     // warp-uniform: iterations since then
	bool handover = false;  // warp-uniform: the stragglers of this warp go to k_coop
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
						if (FAST) {
#pragma unroll
							for (int a = 0; a < 3; ++a)
								s_idir[a * TRACE_THREADS + threadIdx.x] = r.idir[a];
						}
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

#if PRT_SIMPLE_LOOP
		// (A/B variant: every lane runs its own node-or-triangle loop, no phases, no votes)
		if (has_ray) {
			while (s.cur != PRT_DONE) {
				if (at_node(s.cur))
					trav_node_step<COUNT, FAST, W, WT>(s, stack, P.nodes, P.nodes4, r, fr);
				else
					trav_leaf_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, FAST, WT>(s, stack, P.tris, r, opts, &wr,
					                                                             my_idir, idir_stride);
				if (!exhausted && __popc(__activemask()) < P.refill)
					break;
				if (COOP && exhausted && P.coop_after > 0 && ++tail_iters > 4 * P.coop_after) {
					handover = true;
					break;
				}
			}
		}
		__syncwarp();
		handover = __any_sync(0xffffffffu, handover);
#else
		// ---- traverse until too few lanes are still busy
		for (;;) {
#pragma unroll BURST_UNROLL
			for (int k = 0; k < PRT_NODE_BURST; ++k) {
				if (has_ray && at_node(s.cur)) {
					trav_node_step<COUNT, FAST, W, WT>(s, stack, P.nodes, P.nodes4, r, fr);
#if PRT_LEAF_PREFETCH
					// (A/B variant: a lane parked at a triangle has its record prefetched into L1.  Measured
					// 4-7 % SLOWER on C2/C3/C3B/C5: the address arithmetic runs on every node step.)
					if (!W && s.cur < 0)
						asm volatile("prefetch.global.L1 [%0];" ::"l"(
						    reinterpret_cast<const char *>(P.tris) + ((uint64_t)(uint32_t)(~s.cur) << 6)));
#endif
				}
			}
			const bool leaf = has_ray && s.cur < 0;
			const unsigned leafm = __ballot_sync(0xffffffffu, leaf);
			const unsigned nodem = __ballot_sync(0xffffffffu, has_ray && at_node(s.cur));
			const int busy = __popc(leafm | nodem);
			if (busy == 0 || (!exhausted && busy < P.refill))
				break; // (lanes parked at a triangle stay parked across the refill)
			// rays still being traced long after the end of the batch are handed over to k_coop (below)
			if (COOP && exhausted && ++tail_iters > P.coop_after) {
				handover = P.coop_after > 0;
				if (handover)
					break;
			}
			// (once the batch is exhausted nobody waits: what is left is the tail of the kernel,
			// bound by the latency of its longest rays)
			if (leafm && (nodem == 0 || exhausted || __popc(leafm) * 32 >= P.leaf_votes * busy)) {
				if (leaf)
					trav_leaf_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, FAST, WT>(
					    s, stack, P.tris, r, opts, &wr, my_idir, idir_stride);
			}
		}
#endif
		// Hand the rays this warp is still tracing -- state and pending stack, one record each -- over
		// to k_coop, which finishes every one of them with a whole warp (see k_coop above).
		if (COOP && handover) {
			handover = false;
			tail_iters = 0; // lanes that cannot be handed over (list full) keep tracing and try again
			// (only rays with several pending subtrees: a ray walking down a single path offers the 32
			// lanes nothing to share and finishes sooner where it is)
			if (has_ray && s.cur != PRT_DONE && stack.sp >= P.coop_min_sp && (uint32_t)stack.sp <= P.coop_depth) {
				const uint32_t at = atomicAdd(P.coop, 1u);
				if (at < P.coop_cap) {
					uint32_t *p = P.coop + 4 + (size_t)at * P.coop_rec;
					__stcg(p + 0, (uint32_t)ray);
					__stcg(p + 1, (uint32_t)(ray >> 32));
#pragma unroll
					for (int a = 0; a < 3; ++a) {
						__stcg(p + 2 + a, f2u(r.o[a]));
						__stcg(p + 5 + a, f2u(r.d[a]));
					}
					__stcg(p + 8, f2u(s.t_best));
					__stcg(p + 9, f2u(s.u_best));
					__stcg(p + 10, f2u(s.v_best));
					__stcg(p + 11, s.prim_best);
					__stcg(p + 12, (uint32_t)s.cur);
					__stcg(p + 13, (uint32_t)stack.sp);
					uint2 *st = reinterpret_cast<uint2 *>(p + COOP_PARK);
					for (int k = 0; k < stack.sp; ++k)
						__stcg(st + k, stack.peek(k));
					has_ray = false; // (its record is written by k_coop)
					s.cur = PRT_DONE;
				}
			}
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
KernelFn coop_kernel(uint32_t mask, bool aos, bool wt);  // trace_coop.cu

} // namespace prt
