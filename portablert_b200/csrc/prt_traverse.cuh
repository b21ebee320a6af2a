// prt_traverse.cuh -- the per-ray traversal loop, shared by the sm_100a kernels (trace.cu) and by
// the host-side logic emulator used in CPU-only tests (tests/emu; not a product path).
#pragma once

#include "prt_math.cuh"

namespace prt {

// Radix tree over (64-bit key . 32-bit index): depth <= 96, one push per level; the 4-wide nodes
// push up to three entries per (two-level) step: 3 * 48 = 144
constexpr int STACK_DEPTH = 152;

struct TraverseOpts {
	int prune;
	float slack_rel, slack_ulps;
};

struct Hit {
	float t, u, v;
	uint32_t prim;
	uint32_t n_nodes, n_tris;
};

struct Vec4 {
	float x, y, z, w;
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ Vec4 ld16(const void *p) {
	const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
	return Vec4{v.x, v.y, v.z, v.w};
}
// 32 bytes with ONE instruction: sm_100 has 256-bit global loads (LDG.E.ENL2.256); a 64-byte node
// or triangle record is two L1 lookups instead of four
__device__ __forceinline__ void ld32(const void *p, Vec4 &a, Vec4 &b) {
#if defined(PRT_LD128) // (A/B variant: two 128-bit loads)
	a = ld16(p);
	b = ld16(static_cast<const char *>(p) + 16);
	return;
#endif
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
	             : "l"(p));
}
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
#else
inline Vec4 ld16(const void *p) {
	const float *f = static_cast<const float *>(p);
	return Vec4{f[0], f[1], f[2], f[3]};
}
inline void ld32(const void *p, Vec4 &a, Vec4 &b) {
	a = ld16(p);
	b = ld16(static_cast<const char *>(p) + 16);
}
inline uint32_t f2u(float f) {
	uint32_t u;
	__builtin_memcpy(&u, &f, 4);
	return u;
}
inline float u2f(uint32_t u) {
	float f;
	__builtin_memcpy(&f, &u, 4);
	return f;
}
#endif

#if defined(__CUDA_ARCH__)
#define PRT_WARP_ANY(x) __any_sync(__activemask(), (x))
#else
#define PRT_WARP_ANY(x) (x)
#endif

struct StackEntry {
	uint32_t node;
	uint32_t tmin_bits;
};

// Traversal stack of the scalar driver (instrumented runs, host-side emulator): a plain array.  The
// production kernels use a shared-memory short stack with a global overflow (prt_trace_kernel.cuh);
// the traversal code only sees push / pop / sp.
struct ArrayStack {
	StackEntry e[STACK_DEPTH];
	int sp = 0;
	PRT_HD void push(uint32_t node, uint32_t tmin_bits) {
		e[sp].node = node;
		e[sp].tmin_bits = tmin_bits;
		++sp;
	}
	PRT_HD void pop(uint32_t &node, uint32_t &tmin_bits) {
		--sp;
		node = e[sp].node;
		tmin_bits = e[sp].tmin_bits;
	}
	// the next entry whose entry distance does not exceed `limit`, or PRT_DONE (stack empty)
	PRT_HD uint32_t pop_live(float limit) {
		while (sp > 0) {
			--sp;
			if (!(u2f(e[sp].tmin_bits) > limit))
				return e[sp].node;
		}
		return 0x7fffffffu;
	}
	// several conditional pushes without a branch per entry: room(k) says that k more entries can
	// be placed with put() at absolute positions (the device stack: inside its shared-memory part)
#if defined(__CUDACC__)
	PRT_HD uint2 peek(int k) const { return make_uint2(e[k].node, e[k].tmin_bits); }
#endif
	PRT_HD bool room(int) const { return true; }
	PRT_HD void put(int at, bool pred, uint32_t node, uint32_t tmin_bits) {
		if (pred) {
			e[at].node = node;
			e[at].tmin_bits = tmin_bits;
		}
	}
};

// ---------------------------------------------------------------------------------------------
// Conservative fast box test for INTERNAL culling.
//
// The reference slab arithmetic RN(RN(b - o) * idir) costs 2 ops per plane plus NaN-faithful
// min/max.  Whether an internal box is entered does not have to replay that: it only has to be
// CONSERVATIVE with respect to it (never reject a box the reference arithmetic would pass).  The
// fast test evaluates each plane with one FFMA, t' = fma(b, idir, c), c = -RN(o*idir), picks the
// near / far plane of every axis by the sign of the direction (so no min/max per axis) and widens
// every axis by its OWN margin M_a, folded into two intercepts per axis:
//     near:  tn_a = fma(b_near, idir_a, cn_a),  cn_a = RN(c_a - M_a)
//     far :  tf_a = fma(b_far,  idir_a, cf_a),  cf_a = RN(c_a + M_a)
//     pass iff  !(max_a tn_a > min(min_a tf_a, limit))  and  !(min_a tf_a < 0)
// Error bound, per axis, with T = (b-o)*idir exact, B_a = largest |coordinate| of the scene, |e| <= 2^-24:
//     t_ref = T(1+e1)(1+e2);   t' = (T - o*idir*e3 -+ M_a(1+e5))(1+e6)(1+e4)
//     => t' <= t_ref (near) / t' >= t_ref (far) as soon as
//        M_a (1 - 2^-23) >= 2^-23 |o*idir| + 3.01 * 2^-24 |T|,   |T| <= (|o_a| + B_a) |idir_a|
//        <= 1.26 * 2^-22 (|o_a| + B_a) |idir_a|
// M_a = 2^-20 (|o_a| + B_a) |idir_a| is more than three times that, which also covers the rounding
// of the margin arithmetic itself.  For the compressed 4-wide nodes the margin is doubled (the
// quantiser may be half an ulp short at q = 255 and one more product is rounded).
// An axis the ray is exactly parallel to (d_a == 0: the reference computes with idir = inf, i.e.
// the axis constrains nothing when the origin lies strictly inside the slab and rejects when it
// lies outside) takes a large finite slope K = 2^30 * (largest |idir| of the other axes) with
// the margin M_a = 2^-20 * S * K, S = max_b (|o_b| + B_b): inside or within 2^-20 S of the slab both
// intercepts lie beyond every finite entry/exit distance of the other axes (|t_b| <= S |idir_b| <=
// 2^-10 M_a), so the axis constrains nothing; outside by more than that it rejects, as the
// reference does.  (o_a exactly ON a face plane is the reference's 0 * inf = NaN class -- its
// verdict then depends on its own tree topology, DESIGN.md -- and passes here.)
// NaN planes (NaN / inf coordinates) drop out of FMNMX and the final comparisons are written so
// that NaN passes.  A child that is a LEAF (its box is the triangle's own AABB, i.e. the
// reference's per-leaf ray_box_intersect) is re-tested with the exact reference arithmetic before
// a hit is accepted, so the fast test never decides a result -- it only skips subtrees.
// Rays with non-finite components, a zero direction vector, or magnitudes that would overflow the
// intercepts (`ok == false`) do not qualify; they are traced by the exact kernel.
struct FastRay {
	float idir[3];       // slope per axis (1/d, or +K on an axis the ray is parallel to)
	float cn[3], cf[3];  // near / far intercepts, margin folded in
	bool ok;
};

PRT_HD FastRay make_fast_ray(const RayC &r, const float *scene_absmax, bool wide = false) {
	FastRay f;
	const float mscale = wide ? 1.9073486e-6f /* 2^-19 */ : 9.5367432e-7f /* 2^-20 */;
	bool ok = true;
	float S = 0.0f, imax = 0.0f;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const float ao = fabsf(r.o[a]), ad = fabsf(r.d[a]);
		ok = ok && (ao < 1.0e30f) && (ad < INFINITY) && (r.d[a] == r.d[a]);
		S = fmaxf(S, fadd(ao, scene_absmax[a]));
		if (ad > 0.0f)
			imax = fmaxf(imax, fabsf(r.idir[a]));
	}
	const float K = fmul(1073741824.0f /* 2^30 */, imax);
	ok = ok && (imax > 0.0f) && (imax < 1.0e30f) && (S < 1.0e30f);
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const bool par = !(fabsf(r.d[a]) > 0.0f);
		const float id = par ? K : r.idir[a];
		const float c = -fmul(r.o[a], id);
		const float m = fmul(fmul(mscale, par ? S : fadd(fabsf(r.o[a]), scene_absmax[a])), fabsf(id));
		f.idir[a] = id;
		f.cn[a] = fsub(c, m);
		f.cf[a] = fadd(c, m);
		// no intermediate of the test may overflow: |b * id| <= S |id|, intercepts, margins
		ok = ok && (fabsf(c) < 1.0e36f) && (m < 1.0e36f) && (fmul(S, fabsf(id)) < 1.0e36f);
	}
	f.ok = ok;
	return f;
}

#if defined(__CUDA_ARCH__)
#define PRT_FMA(a, b, c) __fmaf_rn(a, b, c)
#else
#define PRT_FMA(a, b, c) fmaf(a, b, c)
#endif

// conservative pass; tmin_out is a lower bound of the reference's entry distance
PRT_HD bool slab_fast(const FastRay &f, const float *lo, const float *hi, float limit,
                      float &tmin_out) {
	float tn[3], tf[3];
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const bool neg = f.idir[a] < 0.0f;
		tn[a] = PRT_FMA(neg ? hi[a] : lo[a], f.idir[a], f.cn[a]);
		tf[a] = PRT_FMA(neg ? lo[a] : hi[a], f.idir[a], f.cf[a]);
	}
	const float tmin = fmaxf(fmaxf(tn[0], tn[1]), tn[2]);
	const float tmax = fminf(fminf(tf[0], tf[1]), tf[2]);
	tmin_out = tmin;
	// reference: reject iff tmax < 0 or tmin > tmax; pruning: tmin > limit.  Written so NaN passes.
	return !(tmin > fminf(tmax, limit)) && !(tmax < 0.0f);
}

// ---------------------------------------------------------------------------------------------
// Per-ray traversal state machine.
//   internal node : both children's boxes are tested (fast conservative test, or the reference's
//                   arithmetic in the exact kernel), the nearer passing child is entered, the other
//                   is pushed with its entry distance
//   leaf          : intersect_tri, then -- for a hit that would become the best -- the triangle's own
//                   AABB with the reference's ray_box_intersect; best hit and pruning limit updated
//   pop           : stacked subtrees whose entry distance exceeds the best hit (+ slack) are dropped
// The result is the minimum over all accepted triangles under an order-independent tie rule, so
// neither the visit order nor the pruning can change it.
struct TravState {
	float t_best, u_best, v_best;
	uint32_t prim_best;
	float limit; // entry distances (exact ones, or lower bounds of them) are compared against this
	float slack_abs;
	int32_t cur; // >= 0 internal node, < 0 leaf (~index), PRT_DONE finished
	uint32_t n_nodes, n_tris;
};

#define PRT_DONE ((int32_t)0x7fffffff)
PRT_HD bool at_node(int32_t cur) { return (uint32_t)cur < 0x7fffffffu; }

PRT_HD void trav_init(TravState &s, const RayC &r, const TraverseOpts &opt, uint64_t n_tris_scene,
                      int32_t root) {
	s.t_best = INFINITY;
	s.u_best = 0.0f;
	s.v_best = 0.0f;
	s.prim_best = 0xffffffffu;
	s.limit = INFINITY;
	// absolute part of the pruning slack: a few ulps of the origin's magnitude expressed in
	// ray-parameter units (NaN/inf here simply disables pruning for this ray)
	const float omax = fmaxf(fmaxf(fabsf(r.o[0]), fabsf(r.o[1])), fabsf(r.o[2]));
	const float dmax = fmaxf(fmaxf(fabsf(r.d[0]), fabsf(r.d[1])), fabsf(r.d[2]));
	s.slack_abs = fmul(fmul(opt.slack_ulps, 1.1920929e-7f), fdiv(omax, dmax));
	s.cur = n_tris_scene == 0 ? PRT_DONE : root;
	s.n_nodes = 0;
	s.n_tris = 0;
}

// Pop: the next stacked subtree that the best hit so far does not rule out, or PRT_DONE.
template <class STK> PRT_HD void trav_pop(TravState &s, STK &stack) {
	s.cur = (int32_t)stack.pop_live(s.limit);
}

// One step through a compressed 4-wide node (fast rays only): dequantise the four child boxes
// straight into ray parameters, t = fma(q, scale*idir, fma(p, idir, c)), near/far plane by the sign
// of the direction, enter the nearest hit child and push the others farthest first.  Returns true
// when nothing was hit (the caller pops).
template <class STK>
PRT_HD bool wide_node_step(TravState &s, STK &stack, const Node4 *nodes4, const FastRay &fr) {
	const char *np = reinterpret_cast<const char *>(nodes4 + s.cur);
	Vec4 v0, v1, v2, v3;
	ld32(np, v0, v1);
	ld32(np + 32, v2, v3);
	const float scale[3] = {v0.w, v3.z, v3.w};
	const uint32_t qw[6] = {f2u(v1.x), f2u(v1.y), f2u(v1.z), f2u(v1.w), f2u(v2.x), f2u(v2.y)};
	const int32_t ch[4] = {(int32_t)f2u(v2.z), (int32_t)f2u(v2.w), (int32_t)f2u(v3.x),
	                       (int32_t)f2u(v3.y)};
	const float pp[3] = {v0.x, v0.y, v0.z};
	float sc[3], bn[3], bf[3];
	uint32_t qn[3], qf[3];
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const bool neg = fr.idir[a] < 0.0f;
		sc[a] = fmul(scale[a], fr.idir[a]);
		bn[a] = PRT_FMA(pp[a], fr.idir[a], fr.cn[a]);
		bf[a] = PRT_FMA(pp[a], fr.idir[a], fr.cf[a]);
		qn[a] = neg ? qw[3 + a] : qw[a];
		qf[a] = neg ? qw[a] : qw[3 + a];
	}
	float t[4];
	int32_t c[4];
	int nh = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		float tn[3], tf[3];
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			tn[a] = PRT_FMA((float)((qn[a] >> (8 * k)) & 0xffu), sc[a], bn[a]);
			tf[a] = PRT_FMA((float)((qf[a] >> (8 * k)) & 0xffu), sc[a], bf[a]);
		}
		const float tmin = fmaxf(fmaxf(tn[0], tn[1]), tn[2]);
		const float tmax = fminf(fminf(tf[0], tf[1]), tf[2]);
		const bool hit = !(tmin > fminf(tmax, s.limit)) && !(tmax < 0.0f) && (ch[k] != PRT_NO_CHILD);
		t[k] = hit ? tmin : INFINITY;
		c[k] = ch[k];
		nh += hit ? 1 : 0;
	}
	if (nh == 0)
		return true;
	// sort the (up to four) hits by entry distance: 5-comparator network, misses carry +inf
#define PRT_CSWAP(i, j)                                                                            \
	{                                                                                              \
		const bool sw = t[j] < t[i];                                                               \
		const float ti = sw ? t[j] : t[i], tj = sw ? t[i] : t[j];                                  \
		const int32_t ci = sw ? c[j] : c[i], cj = sw ? c[i] : c[j];                                \
		t[i] = ti;                                                                                 \
		t[j] = tj;                                                                                 \
		c[i] = ci;                                                                                 \
		c[j] = cj;                                                                                 \
	}
	PRT_CSWAP(0, 1)
	PRT_CSWAP(2, 3)
	PRT_CSWAP(0, 2)
	PRT_CSWAP(1, 3)
	PRT_CSWAP(1, 2)
#undef PRT_CSWAP
	// push the others farthest first: child k (1 <= k < nh) lands at sp + nh - 1 - k
	if (stack.room(3)) {
#pragma unroll
		for (int k = 3; k >= 1; --k)
			stack.put(stack.sp + nh - 1 - k, k < nh, (uint32_t)c[k], f2u(t[k]));
		stack.sp += nh - 1;
	} else {
#pragma unroll
		for (int k = 3; k >= 1; --k)
			if (k < nh)
				stack.push((uint32_t)c[k], f2u(t[k]));
	}
	s.cur = c[0];
	return false;
}

// One step at an internal node (s.cur >= 0): decide the two children (or the four grandchildren of
// a wide node), descend into the nearest, push the rest; pop when nothing is hit.
// FAST: conservative FFMA test; otherwise the reference's arithmetic (WT: its conservative variant).
template <bool COUNT, bool FAST, bool WIDE, bool WT, class STK>
PRT_HD void trav_node_step(TravState &s, STK &stack, const Node *nodes, const Node4 *nodes4,
                           const RayC &r, const FastRay &fr) {
	if (COUNT)
		++s.n_nodes;
	bool pop;
	if (WIDE) {
		pop = wide_node_step(s, stack, nodes4, fr);
	} else {
		const char *np = reinterpret_cast<const char *>(nodes + s.cur);
		Vec4 a, b, c, d;
		ld32(np, a, b);
		ld32(np + 32, c, d);
		const int32_t c0 = (int32_t)f2u(d.x), c1 = (int32_t)f2u(d.y);
		const float lo0[3] = {a.x, a.y, a.z}, hi0[3] = {a.w, b.x, b.y};
		const float lo1[3] = {b.z, b.w, c.x}, hi1[3] = {c.y, c.z, c.w};
		float tm0, tm1;
		bool h0, h1;
		if (FAST) {
			h0 = slab_fast(fr, lo0, hi0, s.limit, tm0);
			h1 = slab_fast(fr, lo1, hi1, s.limit, tm1);
		} else if (WT) {
			h0 = slab_cons(r, lo0, hi0, tm0) && !(tm0 > s.limit);
			h1 = slab_cons(r, lo1, hi1, tm1) && !(tm1 > s.limit);
		} else {
			h0 = slab_ref(r, lo0, hi0, tm0) && !(tm0 > s.limit);
			h1 = slab_ref(r, lo1, hi1, tm1) && !(tm1 > s.limit);
		}
		pop = !(h0 || h1);
		if (h0 && h1) {
			const bool first0 = !(tm1 < tm0);
			stack.push((uint32_t)(first0 ? c1 : c0), f2u(first0 ? tm1 : tm0));
			s.cur = first0 ? c0 : c1;
		} else if (h0 || h1) {
			s.cur = h0 ? c0 : c1;
		}
	}
	if (pop)
		trav_pop(s, stack);
}

// One step at a leaf (s.cur < 0).  Both of the reference's tests must pass (ray_box_intersect on
// the triangle's own AABB, bvh.hpp:237, then intersect_tri, bvh.hpp:246); they are pure functions,
// so the cheaper rejecter runs first: Moeller-Trumbore, then the exact box verdict only for hits
// that would become the best.  In the exact kernel the parent already applied that verdict to this
// very box.  WT = opt-in watertight mode (prt_math.cuh: woop_watertight): the records carry the
// original vertices (v1, v2 in place of the edges) and the triangle's own box keeps only the
// reference's domain rule, conservatively.
// idir / idir_stride: where the ray's 1.0f / d lives (the fast kernels park it in shared memory: it
// is needed about once per ray, for the exact verdict on a triangle's own box).
// trav_leaf_test: the triangle `leaf` (~record index) against the ray; updates the best hit and the
// pruning limit; returns true when the ray is finished (any-hit queries).
template <bool ANYHIT, bool WANT_UV, bool TRACK_PRIM, bool COUNT, bool FAST, bool WT>
PRT_HD bool trav_leaf_test(TravState &s, int32_t leaf, const TriRec *tris, const RayC &r,
                           const TraverseOpts &opt, const WoopRay *wr, const float *idir,
                           int idir_stride) {
	const char *tp = reinterpret_cast<const char *>(tris + (uint32_t)(~leaf));
	Vec4 q0, q1, q2, q3;
	ld32(tp, q0, q1);
	ld32(tp + 32, q2, q3);
	if (COUNT)
		++s.n_tris;
	const float v0[3] = {q0.x, q0.y, q0.z};
	const float e1[3] = {q1.x, q1.y, q1.z};
	const float e2[3] = {q2.x, q2.y, q2.z};
	const uint32_t prim = f2u(q0.w);
	float t, u, v;
	if (WT ? woop_watertight(r, *wr, v0, e1, e2, t, u, v)
	       : moller_trumbore_ref(r, v0, e1, e2, t, u, v)) {
		const bool better = ANYHIT ? (t < s.t_best)
		                           : (TRACK_PRIM ? closer(t, prim, s.t_best, s.prim_best)
		                                         : (t < s.t_best));
		bool box_ok = true;
		if (FAST && better) {
			const float lo[3] = {q1.w, q2.w, q3.x}, hi[3] = {q3.y, q3.z, q3.w};
			RayC rr;
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				rr.o[a] = r.o[a];
				rr.d[a] = r.d[a];
				rr.idir[a] = idir[a * idir_stride];
			}
			float tm;
			box_ok = WT ? slab_cons(rr, lo, hi, tm) : slab_ref(rr, lo, hi, tm);
		}
		if (better && box_ok) {
			s.t_best = t;
			if (ANYHIT) // `valid` is t_near < inf (bvh.hpp:260): any candidate with t < inf decides it
				return true; // (a NaN or +inf t never updates t_near, bvh.hpp:247)
			s.prim_best = prim;
			if (WANT_UV) {
				s.u_best = u;
				s.v_best = v;
			}
			if (opt.prune)
				s.limit = fadd(t, fadd(fmul(fabsf(t), opt.slack_rel), s.slack_abs));
		}
	}
	return false;
}

template <bool ANYHIT, bool WANT_UV, bool TRACK_PRIM, bool COUNT, bool FAST, bool WT, class STK>
PRT_HD void trav_leaf_step(TravState &s, STK &stack, const TriRec *tris, const RayC &r,
                           const TraverseOpts &opt, const WoopRay *wr, const float *idir,
                           int idir_stride) {
	if (trav_leaf_test<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, FAST, WT>(s, s.cur, tris, r, opt, wr, idir,
	                                                                idir_stride))
		s.cur = PRT_DONE;
	else
		trav_pop(s, stack);
}

// Scalar driver (instrumented kernel and the host-side emulator): one ray start to finish.
template <bool ANYHIT, bool WANT_UV, bool TRACK_PRIM, bool COUNT, bool FAST, bool WIDE = false,
          bool WT = false>
PRT_HD void traverse(const Node *nodes, const TriRec *tris, uint64_t n_tris_scene, int32_t root,
                     const RayC &r, const FastRay &fr, const TraverseOpts &opt, Hit &out,
                     const Node4 *nodes4 = nullptr) {
	TravState s;
	ArrayStack stack;
	trav_init(s, r, opt, n_tris_scene, root);
	WoopRay wr{};
	if (WT)
		wr = make_woop_ray(r);
	while (s.cur != PRT_DONE) {
		if (s.cur >= 0)
			trav_node_step<COUNT, FAST, WIDE, WT>(s, stack, nodes, nodes4, r, fr);
		else
			trav_leaf_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, FAST, WT>(s, stack, tris, r, opt, &wr,
			                                                             r.idir, 1);
	}
	out.t = s.t_best;
	out.u = s.u_best;
	out.v = s.v_best;
	out.prim = s.prim_best;
	out.n_nodes = s.n_nodes;
	out.n_tris = s.n_tris;
}

#undef PRT_FMA

} // namespace prt
