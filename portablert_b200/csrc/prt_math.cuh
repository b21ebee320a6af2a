// prt_math.cuh -- parity-critical arithmetic and BVH record types shared by build and traversal.
//
// Everything that decides `valid`, `primitive_id`, t, u, v replays the reference's float
// operations in the reference's operand order with NO fused multiply-add:
//   * ray_box_intersect  include/portableRT/bvh.hpp:195-222   -> prt::slab_ref
//   * intersect_tri      include/portableRT/core.hpp:27-65    -> prt::moller_trumbore_ref
//   * make_aabb          include/portableRT/bvh.hpp:28-37     -> prt::tri_lo / prt::tri_hi
// On the device the *_rn intrinsics are used (ptxas never contracts them into FFMA); compiled for
// the host (tests/emu only -- NOT a product path) plain operators are used and the build adds
// -ffp-contract=off.
#pragma once

#include <cstdint>
#include <math.h>

#if defined(__CUDACC__)
#define PRT_HD __host__ __device__ __forceinline__
#else
#define PRT_HD inline
#endif

namespace prt {

#if defined(__CUDA_ARCH__)
PRT_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
PRT_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
PRT_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
PRT_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// 1.0f / x: the correctly rounded reciprocal IS RN(1/x), the reference's `1.0f / d` (bvh.hpp:199-201)
PRT_HD float frcp(float x) { return __frcp_rn(x); }
#else
PRT_HD float fmul(float a, float b) { return a * b; }
PRT_HD float fadd(float a, float b) { return a + b; }
PRT_HD float fsub(float a, float b) { return a - b; }
PRT_HD float fdiv(float a, float b) { return a / b; }
PRT_HD float frcp(float x) { return 1.0f / x; }
#endif

// std::min / std::max of the reference: (b<a)?b:a and (a<b)?b:a.  NaN-order-sensitive, unlike
// fminf/fmaxf (which drop NaNs).
PRT_HD float smin(float a, float b) { return (b < a) ? b : a; }
PRT_HD float smax(float a, float b) { return (a < b) ? b : a; }

// ---------------------------------------------------------------------------------------------
// Device BVH records.
//
// Node (64 B, four 16-byte loads): a binary node that carries BOTH children's exact fp32 boxes, so
// one fetch decides two subtrees and a leaf's own AABB test (the reference's per-leaf
// ray_box_intersect) is done by its parent without touching the triangle.
//   child < 0            : leaf, triangle record index = ~child
//   child >= 0           : internal node index
// A single-triangle scene (bvh.hpp:165-181) gets a root whose second child is a zero-area dummy
// triangle record, so the traversal loop needs no "absent child" case.
struct __attribute__((aligned(16))) Node {
	float lo0[3];
	float hi0[3];
	float lo1[3];
	float hi1[3];
	int32_t child0, child1;
	uint32_t pad0, pad1;
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");

#define PRT_NO_CHILD ((int32_t)0x7fffffff)

// Triangle record (64 B, four 16-byte loads), stored in Morton order: v0 and the two edges exactly as
// intersect_tri computes them (core.hpp:33-35: edge = v1 - v0, v2 - v0 in binary32), the caller's
// triangle index (= primitive_id) and the triangle's own AABB exactly as make_aabb computes it
// (bvh.hpp:28-37) -- the box the reference tests right before intersect_tri (bvh.hpp:237-246).
struct __attribute__((aligned(16))) TriRec {
	float v0[3];
	uint32_t prim;
	float e1[3];
	float lox;
	float e2[3];
	float loy;
	float loz, hix, hiy, hiz;
};
static_assert(sizeof(TriRec) == 64, "TriRec must be 64 bytes");

struct Box {
	float lo[3], hi[3];
};

// ---------------------------------------------------------------------------------------------
// Compressed 4-wide node (64 B, four 16-byte loads): the four grandchildren of a binary node with
// their boxes quantised to 8 bits per plane inside the node's own bounds (origin p, one power-of-two
// scale per axis).  One fetch decides four subtrees, i.e. two levels of the binary tree, with the
// same 64 bytes a binary node needs for two.  Quantised boxes are only ever used for CONSERVATIVE
// internal culling (prt_traverse.cuh); a triangle's own box is always taken exact from its
// TriRec, so results cannot change.  Wide node i describes the subtree of binary node i (same
// numbering, no allocation pass); only every other level is ever reached from the root.
//   child >= 0 internal (wide/binary node index), < 0 leaf ~triangle record, PRT_NO_CHILD empty
// (The three scales are stored as whole floats -- the record has the room -- so that the traversal
// spends one multiplication per axis on them instead of unpacking exponent bytes.)
struct __attribute__((aligned(16))) Node4 {
	float p[3];
	float scale_x;     // power of two
	uint8_t qlo[3][4]; // [axis][child]
	uint8_t qhi[3][4];
	int32_t child[4];
	float scale_y, scale_z;
};
static_assert(sizeof(Node4) == 64, "Node4 must be 64 bytes");

PRT_HD float pow2_from_bits(uint32_t bits) {
	float f;
#if defined(__CUDA_ARCH__)
	f = __uint_as_float(bits);
#else
	__builtin_memcpy(&f, &bits, 4);
#endif
	return f;
}
PRT_HD float pow2_from_biased(uint32_t e) { return pow2_from_bits(e << 23); }

struct WideChild {
	float lo[3], hi[3];
	int32_t ref;
};

// Quantise up to four children.  Guarantees for every child and axis, in binary32 arithmetic with
// the canonical dequantisation fma(q, scale, p):  lo' <= lo  and  hi' >= hi, except that hi' may
// fall short of hi by at most half an ulp of the coordinate when q saturates at 255 (covered by
// the traversal's error margin like every other rounding of coordinates).
PRT_HD Node4 make_node4(const WideChild *ch, int count) {
	Node4 nd;
	float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (int k = 0; k < count; ++k)
		for (int a = 0; a < 3; ++a) {
			lo[a] = fminf(lo[a], ch[k].lo[a]);
			hi[a] = fmaxf(hi[a], ch[k].hi[a]);
		}
	float scale[3], inv[3];
	for (int a = 0; a < 3; ++a) {
		nd.p[a] = lo[a];
		const float ext = fsub(hi[a], lo[a]);
		// smallest power of two s with 255 * s >= ext
		int k = 0;
		const float m = frexpf(fdiv(ext, 255.0f), &k); // ext/255 = m * 2^k, m in [0.5, 1)
		(void)m;
		int eb = k + 127;
		if (!(ext > 0.0f))
			eb = 1;
		if (eb < 1)
			eb = 1;
		if (eb > 254 || !(ext < INFINITY))
			eb = 254;
		scale[a] = pow2_from_biased((uint32_t)eb);
		// 1 / scale, exact (2^-127 is a subnormal): x * inv == x / scale bit for bit, without the
		// two dozen divisions per node
		inv[a] = eb == 254 ? pow2_from_bits(0x00400000u) : pow2_from_biased((uint32_t)(254 - eb));
	}
	nd.scale_x = scale[0];
	nd.scale_y = scale[1];
	nd.scale_z = scale[2];
	for (int k = 0; k < 4; ++k) {
		if (k < count) {
			nd.child[k] = ch[k].ref;
			for (int a = 0; a < 3; ++a) {
				float ql = floorf(fmul(fsub(ch[k].lo[a], nd.p[a]), inv[a]));
				ql = fminf(fmaxf(ql, 0.0f), 255.0f);
				if (ql > 0.0f && fmaf(ql, scale[a], nd.p[a]) > ch[k].lo[a])
					ql -= 1.0f;
				float qh = ceilf(fmul(fsub(ch[k].hi[a], nd.p[a]), inv[a]));
				qh = fminf(fmaxf(qh, 0.0f), 255.0f);
				if (qh < 255.0f && fmaf(qh, scale[a], nd.p[a]) < ch[k].hi[a])
					qh += 1.0f;
				if (!(ql == ql))
					ql = 0.0f; // NaN coordinates: widest box
				if (!(qh == qh))
					qh = 255.0f;
				nd.qlo[a][k] = (uint8_t)ql;
				nd.qhi[a][k] = (uint8_t)qh;
			}
		} else {
			nd.child[k] = PRT_NO_CHILD;
			for (int a = 0; a < 3; ++a) {
				nd.qlo[a][k] = 255; // inverted box
				nd.qhi[a][k] = 0;
			}
		}
	}
	return nd;
}

// make_aabb, bvh.hpp:28-37 (same nesting: min(a, min(b, c)))
PRT_HD Box tri_box(const float *t9) {
	Box b;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = smin(t9[a], smin(t9[a + 3], t9[a + 6]));
		b.hi[a] = smax(t9[a], smax(t9[a + 3], t9[a + 6]));
	}
	return b;
}

// extend_aabb, bvh.hpp:39-48
PRT_HD Box box_union(const Box &a, const Box &b) {
	Box r;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		r.lo[k] = smin(a.lo[k], b.lo[k]);
		r.hi[k] = smax(a.hi[k], b.hi[k]);
	}
	return r;
}

// Per-ray constants.  idir = 1.0f / d (bvh.hpp:199-201; the reference recomputes it per box, the
// value is the same).
struct RayC {
	float o[3], d[3], idir[3];
};

PRT_HD RayC make_ray(const float *r6) {
	RayC r;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		r.o[a] = r6[a];
		r.d[a] = r6[a + 3];
		r.idir[a] = frcp(r6[a + 3]);
	}
	return r;
}

// ray_box_intersect, bvh.hpp:195-222.  Returns pass/fail exactly as the reference does and the
// entry parameter tmin (bvh.hpp:210) for ordering and pruning.  Because fp32 subtraction and
// multiplication are monotone, a box that contains another passes whenever the inner one does, so
// applying this same arithmetic to the exact union boxes of internal nodes can never cull a
// triangle the reference would accept.  The one exception is 0*inf = NaN (ray parallel to and
// lying in a face plane of a box): there the reference's own verdict depends on its tree topology
// and cannot be reproduced from any other tree -- see DESIGN.md "NaN corner cases".
PRT_HD bool slab_ref(const RayC &r, const float *lo, const float *hi, float &tmin_out) {
	float t1 = fmul(fsub(lo[0], r.o[0]), r.idir[0]);
	float t2 = fmul(fsub(hi[0], r.o[0]), r.idir[0]);
	float t3 = fmul(fsub(lo[1], r.o[1]), r.idir[1]);
	float t4 = fmul(fsub(hi[1], r.o[1]), r.idir[1]);
	float t5 = fmul(fsub(lo[2], r.o[2]), r.idir[2]);
	float t6 = fmul(fsub(hi[2], r.o[2]), r.idir[2]);
	float tmin = smax(smax(smin(t1, t2), smin(t3, t4)), smin(t5, t6));
	float tmax = smin(smin(smax(t1, t2), smax(t3, t4)), smax(t5, t6));
	tmin_out = tmin;
	if (tmax < 0)
		return false;
	if (tmin > tmax)
		return false;
	return true;
}

// intersect_tri, core.hpp:27-65, on a TriRec (edges precomputed with the same subtraction).
// Sums are left-to-right like the C++ expression a*b + c*d + e*f.
PRT_HD bool moller_trumbore_ref(const RayC &r, const float *v0, const float *e1, const float *e2,
                                float &t, float &u, float &v) {
	float p0 = fsub(fmul(r.d[1], e2[2]), fmul(r.d[2], e2[1]));
	float p1 = fsub(fmul(r.d[2], e2[0]), fmul(r.d[0], e2[2]));
	float p2 = fsub(fmul(r.d[0], e2[1]), fmul(r.d[1], e2[0]));
	float det = fadd(fadd(fmul(e1[0], p0), fmul(e1[1], p1)), fmul(e1[2], p2));
	if (det == 0.0f)
		return false;
	float inv = frcp(det);
	float s0 = fsub(r.o[0], v0[0]);
	float s1 = fsub(r.o[1], v0[1]);
	float s2 = fsub(r.o[2], v0[2]);
	u = fmul(fadd(fadd(fmul(s0, p0), fmul(s1, p1)), fmul(s2, p2)), inv);
	if (u < 0.0f || u > 1.0f)
		return false;
	float q0 = fsub(fmul(s1, e1[2]), fmul(s2, e1[1]));
	float q1 = fsub(fmul(s2, e1[0]), fmul(s0, e1[2]));
	float q2 = fsub(fmul(s0, e1[1]), fmul(s1, e1[0]));
	v = fmul(fadd(fadd(fmul(r.d[0], q0), fmul(r.d[1], q1)), fmul(r.d[2], q2)), inv);
	if (v < 0.0f || fadd(u, v) > 1.0f)
		return false;
	t = fmul(fadd(fadd(fmul(e2[0], q0), fmul(e2[1], q1)), fmul(e2[2], q2)), inv);
	return true;
}

// ---------------------------------------------------------------------------------------------
// Opt-in WATERTIGHT triangle test (Woop, Benthin, Wald: "Watertight Ray/Triangle Intersection",
// JCGT 2013).  Not the reference's arithmetic: intersect_tri (core.hpp:27-65) can miss rays that
// pass exactly through an edge or vertex shared by two triangles, because each triangle rounds
// its own u, v differently.  Here the three edge functions are evaluated in a ray-aligned sheared
// space from the ORIGINAL vertices with un-fused multiplies and subtracts, so the edge function
// of a shared edge is the same number (negated) in both triangles and one of them always
// accepts the ray; an exactly-zero edge function is re-evaluated in binary64, where the products
// of binary32 values are exact and the sign is therefore the true one.  The mode keeps the
// reference's result domain (signed t, minimum wins, no tmax, no back-face culling) and its
// (u, v) = weights of v1, v2, so it differs from the default only on edge/vertex grazing rays and
// by rounding of t, u, v; tests/ and bench.py count those differences against the oracle.
struct WoopRay {
	int kx, ky, kz;
	float Sx, Sy, Sz;
};

PRT_HD float pick3(const float *a, int k) { return k == 0 ? a[0] : (k == 1 ? a[1] : a[2]); }

PRT_HD WoopRay make_woop_ray(const RayC &r) {
	WoopRay w;
	const float ax = fabsf(r.d[0]), ay = fabsf(r.d[1]), az = fabsf(r.d[2]);
	w.kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
	w.kx = w.kz == 2 ? 0 : w.kz + 1;
	w.ky = w.kx == 2 ? 0 : w.kx + 1;
	const float dz = pick3(r.d, w.kz);
	if (dz < 0.0f) { // keep the winding
		const int tmp = w.kx;
		w.kx = w.ky;
		w.ky = tmp;
	}
	w.Sx = fdiv(pick3(r.d, w.kx), dz);
	w.Sy = fdiv(pick3(r.d, w.ky), dz);
	w.Sz = frcp(dz);
	return w;
}

// a, b, c: the triangle's vertices as given (core.hpp:24); u, v: weights of b and c like core.hpp
PRT_HD bool woop_watertight(const RayC &r, const WoopRay &w, const float *a, const float *b,
                            const float *c, float &t, float &u, float &v) {
	const float A[3] = {fsub(a[0], r.o[0]), fsub(a[1], r.o[1]), fsub(a[2], r.o[2])};
	const float B[3] = {fsub(b[0], r.o[0]), fsub(b[1], r.o[1]), fsub(b[2], r.o[2])};
	const float C[3] = {fsub(c[0], r.o[0]), fsub(c[1], r.o[1]), fsub(c[2], r.o[2])};
	const float Akz = pick3(A, w.kz), Bkz = pick3(B, w.kz), Ckz = pick3(C, w.kz);
	const float Ax = fsub(pick3(A, w.kx), fmul(w.Sx, Akz)), Ay = fsub(pick3(A, w.ky), fmul(w.Sy, Akz));
	const float Bx = fsub(pick3(B, w.kx), fmul(w.Sx, Bkz)), By = fsub(pick3(B, w.ky), fmul(w.Sy, Bkz));
	const float Cx = fsub(pick3(C, w.kx), fmul(w.Sx, Ckz)), Cy = fsub(pick3(C, w.ky), fmul(w.Sy, Ckz));
	float U = fsub(fmul(Cx, By), fmul(Cy, Bx));
	float V = fsub(fmul(Ax, Cy), fmul(Ay, Cx));
	float W = fsub(fmul(Bx, Ay), fmul(By, Ax));
	if (U == 0.0f || V == 0.0f || W == 0.0f) {
		// products of two binary32 values are exact in binary64: the sign of the difference is exact
		U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
		V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
		W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
	}
	if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
		return false;
	const float det = fadd(fadd(U, V), W);
	if (det == 0.0f)
		return false;
	const float Az = fmul(w.Sz, Akz), Bz = fmul(w.Sz, Bkz), Cz = fmul(w.Sz, Ckz);
	const float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
	const float inv = frcp(det);
	t = fmul(T, inv);
	u = fmul(V, inv);
	v = fmul(W, inv);
	return true;
}

// Box test of the watertight mode where the fast test does not apply: the reference's slab formula
// made conservative with respect to the exact line.  (a) A NaN plane parameter -- 0 * inf: the ray
// runs parallel to an axis and lies exactly in a face plane of the box -- makes that axis
// unbounded instead of poisoning or, with NaN-dropping min/max, wrongly closing the interval.
// (b) The interval is widened by 2^-21 relative: each bound carries at most three roundings of
// 2^-24.  Keeps the reference's domain rule "tmax >= 0" (bvh.hpp:215).
PRT_HD bool slab_cons(const RayC &r, const float *lo, const float *hi, float &tmin_out) {
	float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const float tl = fmul(fsub(lo[a], r.o[a]), r.idir[a]);
		const float th = fmul(fsub(hi[a], r.o[a]), r.idir[a]);
		const bool nan = !(tl == tl) || !(th == th);
		tmin = fmaxf(tmin, nan ? -INFINITY : fminf(tl, th));
		tmax = fminf(tmax, nan ? INFINITY : fmaxf(tl, th));
	}
	const float eps = 4.76837158e-7f; // 2^-21
	// (the widening term is kept finite so that an infinite bound stays infinite, not NaN)
	tmin = fsub(tmin, fmul(fminf(fabsf(tmin), 3.0e38f), eps));
	tmax = fadd(tmax, fmul(fminf(fabsf(tmax), 3.0e38f), eps));
	tmin_out = tmin;
	return !(tmax < 0.0f) && !(tmin > tmax);
}

// Best-hit update.  The reference keeps the first-visited triangle among equal t (strict <,
// bvh.hpp:247); its visit order is an artefact of its own SAH tree, so this backend defines a
// deterministic, traversal-order-independent rule instead: lowest primitive_id among equal t.
PRT_HD bool closer(float t, uint32_t prim, float t_best, uint32_t prim_best) {
	return (t < t_best) || (t == t_best && prim < prim_best);
}

// ---------------------------------------------------------------------------------------------
// Morton codes: `bits` per axis (<= 21), x in the most significant position of each triple.
PRT_HD uint64_t spread3(uint32_t v) { // 21 bits -> every third bit
	uint64_t x = v & 0x1fffffull;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

PRT_HD uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
	return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}

PRT_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
	return __clzll((long long)x);
#else
	return x ? __builtin_clzll(x) : 64;
#endif
}
PRT_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
	return __clz((int)x);
#else
	return x ? __builtin_clz(x) : 32;
#endif
}

// Karras 2012 delta: length of the common prefix of (key_i . i) and (key_j . j); -1 out of range.
PRT_HD int karras_delta(const uint64_t *keys, int64_t n, int64_t i, int64_t j) {
	if (j < 0 || j >= n)
		return -1;
	uint64_t a = keys[i], b = keys[j];
	if (a != b)
		return clz64(a ^ b);
	return 64 + clz32((uint32_t)i ^ (uint32_t)j);
}

// One internal node of the Karras hierarchy over n sorted keys: children and their ranges.
// left/right are encoded child references (>=0 internal, ~leaf for leaves).
PRT_HD void karras_node(const uint64_t *keys, int64_t n, int64_t i, int32_t &left, int32_t &right) {
	int d = (karras_delta(keys, n, i, i + 1) - karras_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	int dmin = karras_delta(keys, n, i, i - d);
	int64_t lmax = 2;
	while (karras_delta(keys, n, i, i + lmax * d) > dmin)
		lmax *= 2;
	int64_t l = 0;
	for (int64_t t = lmax / 2; t >= 1; t /= 2)
		if (karras_delta(keys, n, i, i + (l + t) * d) > dmin)
			l += t;
	int64_t j = i + l * d;
	int dnode = karras_delta(keys, n, i, j);
	int64_t s = 0;
	int64_t t = l;
	do {
		t = (t + 1) / 2;
		if (karras_delta(keys, n, i, i + (s + t) * d) > dnode)
			s += t;
	} while (t > 1);
	int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
	int64_t lo = i < j ? i : j, hi = i < j ? j : i;
	left = (lo == gamma) ? ~(int32_t)gamma : (int32_t)gamma;
	right = (hi == gamma + 1) ? ~(int32_t)(gamma + 1) : (int32_t)(gamma + 1);
}

} // namespace prt
