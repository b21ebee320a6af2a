// prt_traverse.cuh -- the per-ray traversal loop, shared by the sm_100a kernels (trace.cu) and by
// the host-side logic emulator used in CPU-only tests (tests/emu; not a product path).
#pragma once

#include "prt_math.cuh"

namespace prt {

constexpr int STACK_DEPTH = 96; // Karras tree over (64-bit key . 32-bit index): depth <= 96

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

struct StackEntry {
	uint32_t node;
	uint32_t tmin_bits;
};

// Nearest hit of one ray.  Reproduces the result of BVH2::nearest_tri (bvh.hpp:224-265) over a
// different tree: box tests and the triangle test are the reference's own arithmetic
// (prt_math.cuh), entry order and pruning only change how many boxes are looked at.
template <bool ANYHIT, bool WANT_UV, bool TRACK_PRIM, bool COUNT>
PRT_HD void traverse(const Node *nodes, const TriRec *tris, uint64_t n_tris_scene, const RayC &r,
                     const TraverseOpts &opt, Hit &out) {
	float t_best = INFINITY, u_best = 0.0f, v_best = 0.0f;
	uint32_t prim_best = 0xffffffffu;
	float limit = INFINITY;
	// absolute part of the pruning slack: a few ulps of the origin's magnitude expressed in
	// ray-parameter units (NaN/inf here simply disables pruning for this ray)
	const float omax = fmaxf(fmaxf(fabsf(r.o[0]), fabsf(r.o[1])), fabsf(r.o[2]));
	const float dmax = fmaxf(fmaxf(fabsf(r.d[0]), fabsf(r.d[1])), fabsf(r.d[2]));
	const float slack_abs = fmul(fmul(opt.slack_ulps, 1.1920929e-7f), fdiv(omax, dmax));
	uint32_t n_nodes = 0, n_tris = 0;

	StackEntry stack[STACK_DEPTH];
	int sp = 0;
	int32_t cur = 0;
	bool done = n_tris_scene == 0;
	while (!done) {
		bool pop = true;
		if (cur >= 0) {
			const char *np = reinterpret_cast<const char *>(nodes + cur);
			const Vec4 a = ld16(np), b = ld16(np + 16), c = ld16(np + 32), d = ld16(np + 48);
			const int32_t c0 = (int32_t)f2u(d.x), c1 = (int32_t)f2u(d.y);
			if (COUNT)
				++n_nodes;
			const float lo0[3] = {a.x, a.y, a.z}, hi0[3] = {a.w, b.x, b.y};
			const float lo1[3] = {b.z, b.w, c.x}, hi1[3] = {c.y, c.z, c.w};
			float tm0, tm1;
			bool h0 = slab_ref(r, lo0, hi0, tm0);
			bool h1 = slab_ref(r, lo1, hi1, tm1);
			h1 = h1 && (c1 != PRT_NO_CHILD);
			if (opt.prune) {
				h0 = h0 && !(tm0 > limit);
				h1 = h1 && !(tm1 > limit);
			}
			if (h0 && h1) {
				const bool first0 = !(tm1 < tm0);
				stack[sp].node = (uint32_t)(first0 ? c1 : c0);
				stack[sp].tmin_bits = f2u(first0 ? tm1 : tm0);
				++sp;
				cur = first0 ? c0 : c1;
				pop = false;
			} else if (h0 || h1) {
				cur = h0 ? c0 : c1;
				pop = false;
			}
		} else {
			const char *tp = reinterpret_cast<const char *>(tris + (uint32_t)(~cur));
			const Vec4 q0 = ld16(tp), q1 = ld16(tp + 16), q2 = ld16(tp + 32);
			if (COUNT)
				++n_tris;
			const float v0[3] = {q0.x, q0.y, q0.z};
			const float e1[3] = {q1.x, q1.y, q1.z};
			const float e2[3] = {q2.x, q2.y, q2.z};
			const uint32_t prim = f2u(q0.w);
			float t, u, v;
			if (moller_trumbore_ref(r, v0, e1, e2, t, u, v)) {
				if (ANYHIT) {
					// `valid` is t_near < inf (bvh.hpp:260): any candidate with t < inf decides it
					// (a NaN or +inf t never updates t_near, bvh.hpp:247)
					if (t < t_best) {
						t_best = t;
						break;
					}
				} else {
					const bool better =
					    TRACK_PRIM ? closer(t, prim, t_best, prim_best) : (t < t_best);
					if (better) {
						t_best = t;
						prim_best = prim;
						if (WANT_UV) {
							u_best = u;
							v_best = v;
						}
						limit = fadd(t_best, fadd(fmul(fabsf(t_best), opt.slack_rel), slack_abs));
					}
				}
			}
		}
		if (pop) {
			for (;;) {
				if (sp == 0) {
					done = true;
					break;
				}
				--sp;
				if (opt.prune && u2f(stack[sp].tmin_bits) > limit)
					continue;
				cur = (int32_t)stack[sp].node;
				break;
			}
		}
	}
	out.t = t_best;
	out.u = u_best;
	out.v = v_best;
	out.prim = prim_best;
	out.n_nodes = n_nodes;
	out.n_tris = n_tris;
}

} // namespace prt
