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
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
#else
inline Vec4 ld16(const void *p) {
	const float *f = static_cast<const float *>(p);
	return Vec4{f[0], f[1], f[2], f[3]};
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

// ---------------------------------------------------------------------------------------------
// Conservative fast box test for INTERNAL culling.
//
// The reference slab arithmetic RN(RN(b - o) * idir) costs 2 ops per plane plus NaN-faithful
// min/max (2 ops each).  Whether an internal box is entered does not have to replay that: it only
// has to be CONSERVATIVE with respect to it (never reject a box the reference arithmetic would
// pass).  The fast test uses one FFMA per plane, t' = fma(b, idir, -RN(o*idir)), plain FMNMX
// min/max, and widens the interval by a margin M that bounds |t' - t_ref|:
//     t_ref = T(1+e1)(1+e2),  t' = (T - o*idir*e3)(1+e4),  T = (b-o)*idir exact, |e| <= 2^-24
//     => |t' - t_ref| <= 2^-24 (|o*idir| + 3.01 |T|) <= 2^-22 (|o_a| + B_a) |idir_a|
// with B_a the largest |coordinate| of the scene on axis a.  M is taken as TWICE that bound,
// maximised over the axes, which also covers the rounding of the margin arithmetic itself.
// A child that is a LEAF (its box is the triangle's own AABB, i.e. the reference's per-leaf
// ray_box_intersect) is re-tested with the exact reference arithmetic before the triangle is
// touched, so the fast test never decides a result -- it only skips subtrees.
// Rays with a zero / denormal / very small direction component (|d_a| < 2^-12 max|d|), or
// non-finite intermediates, do not qualify and take the exact path for every box.
struct FastRay {
	float idir[3], c[3]; // t' = fma(b, idir, c), c = -RN(o * idir)
	float M;             // margin (see above)
	bool ok;
};

PRT_HD FastRay make_fast_ray(const RayC &r, const float *scene_absmax) {
	FastRay f;
	float M = 0.0f;
	const float dmax = fmaxf(fmaxf(fabsf(r.d[0]), fabsf(r.d[1])), fabsf(r.d[2]));
	bool ok = dmax < INFINITY;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		f.idir[a] = r.idir[a];
		const float oi = fmul(r.o[a], r.idir[a]);
		f.c[a] = -oi;
		const float m = fmul(fmul(4.76837158e-7f /* 2^-21 */, fadd(fabsf(r.o[a]), scene_absmax[a])),
		                     fabsf(r.idir[a]));
		M = fmaxf(M, m);
		ok = ok && (fabsf(r.d[a]) >= fmul(dmax, 2.44140625e-4f /* 2^-12 */)) &&
		     (fabsf(oi) < 1.0e30f) && (fabsf(r.idir[a]) < 1.0e30f);
	}
	f.M = M;
	f.ok = ok && (M < 1.0e30f) && (M == M);
	return f;
}

// returns conservative pass; tmin_out is a lower bound (minus nothing: compare against limit + M)
PRT_HD bool slab_fast(const FastRay &f, const float *lo, const float *hi, float &tmin_out) {
#if defined(__CUDA_ARCH__)
#define PRT_FMA(a, b, c) __fmaf_rn(a, b, c)
#else
#define PRT_FMA(a, b, c) fmaf(a, b, c)
#endif
	const float x0 = PRT_FMA(lo[0], f.idir[0], f.c[0]), x1 = PRT_FMA(hi[0], f.idir[0], f.c[0]);
	const float y0 = PRT_FMA(lo[1], f.idir[1], f.c[1]), y1 = PRT_FMA(hi[1], f.idir[1], f.c[1]);
	const float z0 = PRT_FMA(lo[2], f.idir[2], f.c[2]), z1 = PRT_FMA(hi[2], f.idir[2], f.c[2]);
#undef PRT_FMA
	const float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
	const float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
	tmin_out = tmin;
	// reference: reject iff tmax < 0 or tmin > tmax; widened by M on both ends
	return (tmax >= -f.M) && (fsub(tmin, tmax) <= fadd(f.M, f.M));
}

// ---------------------------------------------------------------------------------------------
// Per-ray traversal state machine.  One call of trav_step() handles one tree element:
//   internal node : both children's boxes are tested (fast conservative test, or the reference's
//                   arithmetic when the ray does not qualify), the nearer passing child is entered,
//                   the other is pushed with its entry distance
//   leaf          : the triangle's own AABB is tested with the reference's ray_box_intersect, then
//                   intersect_tri; the best hit and the pruning limits are updated
// followed by a pop that discards stacked subtrees whose entry distance exceeds the best hit
// (+ slack).  The result is the minimum over all accepted triangles under an order-independent tie
// rule, so neither the visit order nor the pruning can change it.
struct TravState {
	float t_best, u_best, v_best;
	uint32_t prim_best;
	float limit;  // exact entry distances are compared against this
	float limitM; // fast (lower-bound) entry distances against limit + M
	float slack_abs;
	int sp;
	int32_t cur; // >= 0 internal node, < 0 leaf (~index), PRT_DONE finished
	uint32_t n_nodes, n_tris;
};

#define PRT_DONE ((int32_t)0x7fffffff)

PRT_HD void trav_init(TravState &s, const RayC &r, const TraverseOpts &opt, uint64_t n_tris_scene,
                      int32_t root) {
	s.t_best = INFINITY;
	s.u_best = 0.0f;
	s.v_best = 0.0f;
	s.prim_best = 0xffffffffu;
	s.limit = INFINITY;
	s.limitM = INFINITY;
	// absolute part of the pruning slack: a few ulps of the origin's magnitude expressed in
	// ray-parameter units (NaN/inf here simply disables pruning for this ray)
	const float omax = fmaxf(fmaxf(fabsf(r.o[0]), fabsf(r.o[1])), fabsf(r.o[2]));
	const float dmax = fmaxf(fmaxf(fabsf(r.d[0]), fabsf(r.d[1])), fabsf(r.d[2]));
	s.slack_abs = fmul(fmul(opt.slack_ulps, 1.1920929e-7f), fdiv(omax, dmax));
	s.sp = 0;
	s.cur = n_tris_scene == 0 ? PRT_DONE : root;
	s.n_nodes = 0;
	s.n_tris = 0;
}

// One step through a compressed 4-wide node (FAST rays only): dequantise the four child boxes
// straight into ray parameters, t = fma(q, scale*idir, fma(p, idir, c)), test them with the
// conservative margin (twice the binary node's: the quantiser may be half an ulp short at q = 255
// and one more product is rounded), enter the nearest hit child and push the others farthest
// first.  Returns true when nothing was hit (the caller pops).
PRT_HD bool wide_node_step(TravState &s, StackEntry *stack, const Node4 *nodes4, const FastRay &fr) {
	const char *np = reinterpret_cast<const char *>(nodes4 + s.cur);
	const Vec4 v0 = ld16(np), v1 = ld16(np + 16), v2 = ld16(np + 32), v3 = ld16(np + 48);
	const uint32_t ebits = f2u(v0.w);
	const uint32_t qw[6] = {f2u(v1.x), f2u(v1.y), f2u(v1.z), f2u(v1.w), f2u(v2.x), f2u(v2.y)};
	const int32_t ch[4] = {(int32_t)f2u(v2.z), (int32_t)f2u(v2.w), (int32_t)f2u(v3.x),
	                       (int32_t)f2u(v3.y)};
	const float pp[3] = {v0.x, v0.y, v0.z};
	float sc[3], bb[3];
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		sc[a] = fmul(pow2_from_biased((ebits >> (8 * a)) & 0xffu), fr.idir[a]);
#if defined(__CUDA_ARCH__)
		bb[a] = __fmaf_rn(pp[a], fr.idir[a], fr.c[a]);
#else
		bb[a] = fmaf(pp[a], fr.idir[a], fr.c[a]);
#endif
	}
	const float M4 = fadd(fr.M, fr.M);
	const float lim = fadd(s.limitM, fr.M);
	float t[4];
	int32_t c[4];
	int nh = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const float ql = (float)((qw[a] >> (8 * k)) & 0xffu);
			const float qh = (float)((qw[3 + a] >> (8 * k)) & 0xffu);
#if defined(__CUDA_ARCH__)
			const float tl = __fmaf_rn(ql, sc[a], bb[a]), th = __fmaf_rn(qh, sc[a], bb[a]);
#else
			const float tl = fmaf(ql, sc[a], bb[a]), th = fmaf(qh, sc[a], bb[a]);
#endif
			tmin = fmaxf(tmin, fminf(tl, th));
			tmax = fminf(tmax, fmaxf(tl, th));
		}
		const bool hit = (tmax >= -M4) && (fsub(tmin, tmax) <= fadd(M4, M4)) && !(tmin > lim) &&
		                 (ch[k] != PRT_NO_CHILD);
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
#pragma unroll
	for (int k = 3; k >= 1; --k) {
		if (k < nh) {
			stack[s.sp].node = (uint32_t)c[k];
			stack[s.sp].tmin_bits = f2u(t[k]);
			++s.sp;
		}
	}
	s.cur = c[0];
	return false;
}

// WT = opt-in watertight mode (prt_math.cuh: woop_watertight): triangle records carry the original
// vertices (v1, v2 in place of the edges), every box test is conservative with respect to the exact
// line (fast test with its margin, or slab_cons), and the triangle's own box keeps only the
// reference's domain rule.
template <bool ANYHIT, bool WANT_UV, bool TRACK_PRIM, bool COUNT, bool FAST, bool WIDE = false,
          bool WT = false>
PRT_HD void trav_step(TravState &s, StackEntry *stack, const Node *nodes, const TriRec *tris,
                      const RayC &r, const FastRay &fr, const TraverseOpts &opt,
                      const Node4 *nodes4 = nullptr, const WoopRay *wr = nullptr) {
	bool pop = true;
	if (WIDE && s.cur >= 0) {
		if (COUNT)
			++s.n_nodes;
		pop = wide_node_step(s, stack, nodes4, fr);
	} else if (s.cur >= 0) {
		const char *np = reinterpret_cast<const char *>(nodes + s.cur);
		const Vec4 a = ld16(np), b = ld16(np + 16), c = ld16(np + 32), d = ld16(np + 48);
		const int32_t c0 = (int32_t)f2u(d.x), c1 = (int32_t)f2u(d.y);
		if (COUNT)
			++s.n_nodes;
		const float lo0[3] = {a.x, a.y, a.z}, hi0[3] = {a.w, b.x, b.y};
		const float lo1[3] = {b.z, b.w, c.x}, hi1[3] = {c.y, c.z, c.w};
		float tm0, tm1;
		bool h0, h1;
		if (FAST) {
			h0 = slab_fast(fr, lo0, hi0, tm0) && !(tm0 > s.limitM);
			h1 = slab_fast(fr, lo1, hi1, tm1) && !(tm1 > s.limitM);
		} else if (WT) {
			h0 = slab_cons(r, lo0, hi0, tm0) && !(tm0 > s.limit);
			h1 = slab_cons(r, lo1, hi1, tm1) && !(tm1 > s.limit);
		} else {
			h0 = slab_ref(r, lo0, hi0, tm0) && !(tm0 > s.limit);
			h1 = slab_ref(r, lo1, hi1, tm1) && !(tm1 > s.limit);
		}
		if (h0 && h1) {
			const bool first0 = !(tm1 < tm0);
			stack[s.sp].node = (uint32_t)(first0 ? c1 : c0);
			stack[s.sp].tmin_bits = f2u(first0 ? tm1 : tm0);
			++s.sp;
			s.cur = first0 ? c0 : c1;
			pop = false;
		} else if (h0 || h1) {
			s.cur = h0 ? c0 : c1;
			pop = false;
		}
	} else {
		const char *tp = reinterpret_cast<const char *>(tris + (uint32_t)(~s.cur));
		const Vec4 q0 = ld16(tp), q1 = ld16(tp + 16), q2 = ld16(tp + 32), q3 = ld16(tp + 48);
		if (COUNT)
			++s.n_tris;
		// Both of the reference's tests must pass (ray_box_intersect on the triangle's own AABB,
		// bvh.hpp:237, then intersect_tri, bvh.hpp:246); they are pure functions, so the cheaper
		// rejecter runs first: Moeller-Trumbore, then the exact box verdict only for accepted hits.
		// On the exact path the parent already applied that verdict to this very box.
		const float v0[3] = {q0.x, q0.y, q0.z};
		const float e1[3] = {q1.x, q1.y, q1.z};
		const float e2[3] = {q2.x, q2.y, q2.z};
		const uint32_t prim = f2u(q0.w);
		float t, u, v;
		// (watertight records: e1, e2 hold the vertices v1, v2)
		if (WT ? woop_watertight(r, *wr, v0, e1, e2, t, u, v)
		       : moller_trumbore_ref(r, v0, e1, e2, t, u, v)) {
			const bool better = ANYHIT ? (t < s.t_best)
			                           : (TRACK_PRIM ? closer(t, prim, s.t_best, s.prim_best)
			                                         : (t < s.t_best));
			bool box_ok = true;
			if (FAST && better) {
				const float lo[3] = {q1.w, q2.w, q3.x}, hi[3] = {q3.y, q3.z, q3.w};
				float tm;
				box_ok = WT ? slab_cons(r, lo, hi, tm) : slab_ref(r, lo, hi, tm);
			}
			if (better && box_ok) {
				if (ANYHIT) {
					// `valid` is t_near < inf (bvh.hpp:260): any candidate with t < inf decides it
					// (a NaN or +inf t never updates t_near, bvh.hpp:247)
					s.t_best = t;
					s.cur = PRT_DONE;
					return;
				}
				s.t_best = t;
				s.prim_best = prim;
				if (WANT_UV) {
					s.u_best = u;
					s.v_best = v;
				}
				if (opt.prune) {
					s.limit = fadd(t, fadd(fmul(fabsf(t), opt.slack_rel), s.slack_abs));
					s.limitM = FAST ? fadd(s.limit, fr.M) : s.limit;
				}
			}
		}
	}
	if (pop) {
		s.cur = PRT_DONE;
		while (s.sp > 0) {
			--s.sp;
			// a stacked entry distance is a lower bound within M on the fast path
			const float lim = WIDE ? fadd(s.limitM, fr.M) : (FAST ? s.limitM : s.limit);
			if (!(u2f(stack[s.sp].tmin_bits) > lim)) {
				s.cur = (int32_t)stack[s.sp].node;
				break;
			}
		}
	}
}

// Scalar driver (instrumented kernel and the host-side emulator): one ray start to finish.
template <bool ANYHIT, bool WANT_UV, bool TRACK_PRIM, bool COUNT, bool FAST, bool WIDE = false,
          bool WT = false>
PRT_HD void traverse(const Node *nodes, const TriRec *tris, uint64_t n_tris_scene, int32_t root,
                     const RayC &r, const FastRay &fr, const TraverseOpts &opt, Hit &out,
                     const Node4 *nodes4 = nullptr) {
	TravState s;
	StackEntry stack[STACK_DEPTH];
	trav_init(s, r, opt, n_tris_scene, root);
	WoopRay wr{};
	if (WT)
		wr = make_woop_ray(r);
	while (s.cur != PRT_DONE)
		trav_step<ANYHIT, WANT_UV, TRACK_PRIM, COUNT, FAST, WIDE, WT>(s, stack, nodes, tris, r, fr,
		                                                              opt, nodes4, &wr);
	out.t = s.t_best;
	out.u = s.u_best;
	out.v = s.v_best;
	out.prim = s.prim_best;
	out.n_nodes = s.n_nodes;
	out.n_tris = s.n_tris;
}

} // namespace prt
