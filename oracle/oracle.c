/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the reference's nearest-hit path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libprt_b200.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the reference's known-answer vectors (examples/validation/main.cpp:68-110,
 *     examples/triangle/main.cpp:7-21, README.md:144-165), and
 *   - outputs of the unmodified reference CPU backend (oracle/_ref/libprt_ref.so, built from
 *     /root/reference by oracle/Makefile) and the committed fixtures under tests/golden/.
 *
 * Two restatements live here:
 *   (A) oracle_build / oracle_trace : the reference's own data structure -- binned-SAH BVH2
 *       (include/portableRT/bvh.hpp:58-193) traversed by the right-child-first DFS of
 *       bvh.hpp:224-265.  Same tree topology, same visit order, hence the same tie winners.
 *   (B) oracle_brute : the topology-free acceptance rule the new backend is built on:
 *       triangle k is a candidate iff ray_box_intersect(ray, make_aabb(tri_k)) (bvh.hpp:195-222)
 *       and intersect_tri(tri_k, ray) (core.hpp:27-65); result = candidate of minimum t (strict <),
 *       ties to the lowest primitive index.
 *
 * All arithmetic is IEEE binary32 in the reference's operand order.  Build WITHOUT -march=native /
 * -ffast-math (no FMA contraction), like the reference binary.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORACLE_NO_TRI 0xFFFFFFFFu

typedef struct {
	int is_leaf;      /* bvh.hpp:21 */
	uint32_t tri;     /* bvh.hpp:22, -1 when none */
	uint32_t left;    /* bvh.hpp:23 */
	uint32_t right;   /* bvh.hpp:24 */
	float lo[3], hi[3]; /* bvh.hpp:25 */
} onode;

typedef struct {
	onode *nodes;
	uint64_t n_nodes, cap_nodes;
	float *tris; /* n_tris * 9 */
	uint64_t n_tris;
} oracle_bvh;

/* std::min / std::max semantics: (b<a)?b:a and (a<b)?b:a -- NaN-order-sensitive, unlike fminf. */
static inline float smin(float a, float b) { return (b < a) ? b : a; }
static inline float smax(float a, float b) { return (a < b) ? b : a; }

/* bvh.hpp:13-18 */
static void box_empty(float lo[3], float hi[3]) {
	for (int a = 0; a < 3; ++a) {
		lo[a] = FLT_MAX;
		hi[a] = -FLT_MAX;
	}
}

/* bvh.hpp:28-37 */
static void box_of_tri(const float *t, float lo[3], float hi[3]) {
	for (int a = 0; a < 3; ++a) {
		lo[a] = smin(t[a], smin(t[a + 3], t[a + 6]));
		hi[a] = smax(t[a], smax(t[a + 3], t[a + 6]));
	}
}

/* bvh.hpp:39-48 : result = union(a, b), a is the accumulator */
static void box_extend(float alo[3], float ahi[3], const float blo[3], const float bhi[3]) {
	for (int a = 0; a < 3; ++a) {
		alo[a] = smin(alo[a], blo[a]);
		ahi[a] = smax(ahi[a], bhi[a]);
	}
}

/* bvh.hpp:51-56 */
static float box_harea(const float lo[3], const float hi[3]) {
	const float dx = hi[0] - lo[0];
	const float dy = hi[1] - lo[1];
	const float dz = hi[2] - lo[2];
	return (dx * dy + dx * dz + dy * dz);
}

/* centroid coordinate used by the binning, bvh.hpp:84 and :109-111 */
static inline float centroid(const float *t, int axis) {
	return (t[axis] + t[axis + 3] + t[axis + 6]) / 3.0f;
}

static uint32_t node_new(oracle_bvh *b) {
	if (b->n_nodes == b->cap_nodes) {
		b->cap_nodes = b->cap_nodes ? b->cap_nodes * 2 : 64;
		b->nodes = (onode *)realloc(b->nodes, b->cap_nodes * sizeof(onode));
	}
	onode *n = &b->nodes[b->n_nodes];
	n->is_leaf = 0; /* bvh.hpp:21-25 defaults */
	n->tri = ORACLE_NO_TRI;
	n->left = ORACLE_NO_TRI;
	n->right = ORACLE_NO_TRI;
	box_empty(n->lo, n->hi);
	return (uint32_t)b->n_nodes++;
}

/*
 * divide_sah, bvh.hpp:58-129.  `idx[0..n)` is the node's triangle list in the reference's input
 * order; on return idx is stably partitioned into left = idx[0..*n_left) and right = the rest
 * (the reference push_back's into two fresh vectors, which is a stable partition).
 */
static void split_sah(const float *tris, uint32_t *idx, uint64_t n, uint32_t *scratch,
                      uint64_t *n_left) {
	const int bin_count = 14; /* :60 */
	float min_cost = FLT_MAX; /* :61 */
	int best_axis = 0;        /* :62 */
	int best_bin = bin_count / 2; /* :63 */

	float nlo[3], nhi[3], tlo[3], thi[3];
	box_empty(nlo, nhi);
	for (uint64_t i = 0; i < n; ++i) { /* :65-68 */
		box_of_tri(tris + 9ull * idx[i], tlo, thi);
		box_extend(nlo, nhi, tlo, thi);
	}

	for (int bin = 0; bin < bin_count; ++bin) { /* :70 bin-major */
		for (int axis = 0; axis < 3; ++axis) {  /* :71 axis-minor */
			/* :73-75 -- float*float/size_t : ((hi-lo)*bin)/14 in binary32 */
			float bin_pos = nlo[axis] + (nhi[axis] - nlo[axis]) * (float)bin / (float)bin_count;
			float llo[3], lhi[3], rlo[3], rhi[3];
			box_empty(llo, lhi);
			box_empty(rlo, rhi);
			int lc = 0, rc = 0;
			for (uint64_t i = 0; i < n; ++i) { /* :82-91 */
				const float *t = tris + 9ull * idx[i];
				float c = centroid(t, axis);
				box_of_tri(t, tlo, thi);
				if (c < bin_pos) {
					box_extend(llo, lhi, tlo, thi);
					++lc;
				} else {
					box_extend(rlo, rhi, tlo, thi);
					++rc;
				}
			}
			/* :93-94 ; an empty side gives inf*0 = NaN, and NaN < min_cost is false */
			float cost = box_harea(llo, lhi) * (float)lc + box_harea(rlo, rhi) * (float)rc;
			if (cost < min_cost) { /* :95-99 first strict minimum wins */
				min_cost = cost;
				best_axis = axis;
				best_bin = bin;
			}
		}
	}

	/* :103-105 */
	float bin_pos = nlo[best_axis] +
	                (nhi[best_axis] - nlo[best_axis]) * (float)best_bin / (float)bin_count;
	uint64_t nl = 0, nr = 0;
	for (uint64_t i = 0; i < n; ++i) { /* :107-116 */
		float c = centroid(tris + 9ull * idx[i], best_axis);
		if (c < bin_pos)
			idx[nl++] = idx[i]; /* nl <= i, safe in place */
		else
			scratch[nr++] = idx[i];
	}
	if (nl == 0 || nr == 0) {
		/* :118-128 -- fall back to halving the INPUT order.  When one side is empty the other side
		 * is the whole input in input order, so restore it and cut at n/2. */
		if (nl == 0)
			memcpy(idx, scratch, n * sizeof(uint32_t));
		*n_left = n / 2;
		return;
	}
	memcpy(idx + nl, scratch, nr * sizeof(uint32_t));
	*n_left = nl;
}

/* build_aux, bvh.hpp:131-156 : pre-order emission, left subtree before right */
static uint32_t build_rec(oracle_bvh *b, uint32_t *idx, uint64_t n, uint32_t *scratch) {
	const uint32_t me = node_new(b);
	if (n == 0) {
		b->nodes[me].is_leaf = 1; /* :136-137 */
	} else if (n == 1) {
		b->nodes[me].is_leaf = 1; /* :138-141 */
		b->nodes[me].tri = idx[0];
		box_of_tri(b->tris + 9ull * idx[0], b->nodes[me].lo, b->nodes[me].hi);
	} else {
		float lo[3], hi[3], tlo[3], thi[3];
		box_empty(lo, hi);
		for (uint64_t i = 0; i < n; ++i) { /* :144-147 */
			box_of_tri(b->tris + 9ull * idx[i], tlo, thi);
			box_extend(lo, hi, tlo, thi);
		}
		memcpy(b->nodes[me].lo, lo, sizeof lo);
		memcpy(b->nodes[me].hi, hi, sizeof hi);
		uint64_t nl = 0;
		split_sah(b->tris, idx, n, scratch, &nl); /* :151 */
		uint32_t l = build_rec(b, idx, nl, scratch); /* :152 */
		b->nodes[me].left = l;                       /* (re-index: realloc may move nodes) */
		uint32_t r = build_rec(b, idx + nl, n - nl, scratch + nl); /* :153 */
		b->nodes[me].right = r;
	}
	return me;
}

/* BVH2::build, bvh.hpp:158-193 */
oracle_bvh *oracle_build(const float *tris9, uint64_t n) {
	oracle_bvh *b = (oracle_bvh *)calloc(1, sizeof(oracle_bvh));
	b->n_tris = n;
	b->tris = (float *)malloc((n ? n : 1) * 9 * sizeof(float));
	if (n)
		memcpy(b->tris, tris9, n * 9 * sizeof(float));
	if (n == 1) { /* :165-181 hand-made 3-node tree */
		uint32_t r = node_new(b), l1 = node_new(b), l2 = node_new(b);
		box_of_tri(b->tris, b->nodes[r].lo, b->nodes[r].hi);
		b->nodes[r].left = l1;
		b->nodes[r].right = l2;
		b->nodes[l1].is_leaf = 1;
		b->nodes[l1].tri = 0;
		box_of_tri(b->tris, b->nodes[l1].lo, b->nodes[l1].hi);
		b->nodes[l2].is_leaf = 1;
	} else { /* :183-187 ; n == 0 yields a single empty leaf via build_aux */
		uint32_t *idx = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
		uint32_t *scratch = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
		for (uint64_t i = 0; i < n; ++i)
			idx[i] = (uint32_t)i;
		build_rec(b, idx, n, scratch);
		free(idx);
		free(scratch);
	}
	return b;
}

void oracle_free(oracle_bvh *b) {
	if (!b)
		return;
	free(b->nodes);
	free(b->tris);
	free(b);
}

uint64_t oracle_node_count(const oracle_bvh *b) { return b->n_nodes; }

/* maximum depth of the tree (root = 1); the reference's DFS stack holds at most depth+1 entries */
uint32_t oracle_depth(const oracle_bvh *b) {
	uint32_t best = 0;
	uint64_t cap = 64, sp = 0;
	uint32_t *st = (uint32_t *)malloc(cap * 2 * sizeof(uint32_t));
	st[0] = 0;
	st[1] = 1;
	sp = 1;
	while (sp) {
		--sp;
		uint32_t n = st[2 * sp], d = st[2 * sp + 1];
		if (d > best)
			best = d;
		if (!b->nodes[n].is_leaf) {
			if (sp + 2 > cap) {
				cap *= 2;
				st = (uint32_t *)realloc(st, cap * 2 * sizeof(uint32_t));
			}
			st[2 * sp] = b->nodes[n].left;
			st[2 * sp + 1] = d + 1;
			++sp;
			st[2 * sp] = b->nodes[n].right;
			st[2 * sp + 1] = d + 1;
			++sp;
		}
	}
	free(st);
	return best;
}

/* ray_box_intersect, bvh.hpp:195-222 */
static int slab(const float o[3], const float d[3], const float lo[3], const float hi[3]) {
	float f0 = 1.0f / d[0], f1 = 1.0f / d[1], f2 = 1.0f / d[2]; /* :199-201 */
	float t1 = (lo[0] - o[0]) * f0;
	float t2 = (hi[0] - o[0]) * f0;
	float t3 = (lo[1] - o[1]) * f1;
	float t4 = (hi[1] - o[1]) * f1;
	float t5 = (lo[2] - o[2]) * f2;
	float t6 = (hi[2] - o[2]) * f2;
	float tmin = smax(smax(smin(t1, t2), smin(t3, t4)), smin(t5, t6)); /* :210 */
	float tmax = smin(smin(smax(t1, t2), smax(t3, t4)), smax(t5, t6)); /* :211 */
	if (tmax < 0)
		return 0; /* :213 */
	if (tmin > tmax)
		return 0; /* :217 */
	return 1;
}

/* intersect_tri, core.hpp:27-65 (Moeller-Trumbore, no culling, no epsilon, no t>=0 test) */
static int moller_trumbore(const float *tv, const float o[3], const float d[3], float *t, float *u,
                           float *v) {
	float e1[3] = {tv[3] - tv[0], tv[4] - tv[1], tv[5] - tv[2]};
	float e2[3] = {tv[6] - tv[0], tv[7] - tv[1], tv[8] - tv[2]};
	float pv[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2],
	               d[0] * e2[1] - d[1] * e2[0]};
	float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
	if (det == 0.0f)
		return 0;
	float inv = 1.0f / det;
	float tv3[3] = {o[0] - tv[0], o[1] - tv[1], o[2] - tv[2]};
	*u = (tv3[0] * pv[0] + tv3[1] * pv[1] + tv3[2] * pv[2]) * inv;
	if (*u < 0.0f || *u > 1.0f)
		return 0;
	float qv[3] = {tv3[1] * e1[2] - tv3[2] * e1[1], tv3[2] * e1[0] - tv3[0] * e1[2],
	               tv3[0] * e1[1] - tv3[1] * e1[0]};
	*v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
	if (*v < 0.0f || *u + *v > 1.0f)
		return 0;
	*t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
	return 1;
}

/* SoA hit record, all five tags.  On a miss u, v, pid are indeterminate in the reference
 * (bvh.hpp:232,249-251 never initialise them); the oracle writes 0, 0, 0xFFFFFFFF and the
 * comparator ignores them. */
typedef struct {
	float *t, *u, *v, *px, *py, *pz;
	uint32_t *pid;
	uint8_t *valid;
} oracle_out;

static void finish(const float o[3], const float d[3], float t_near, float u, float v,
                   uint32_t pid, uint64_t i, const oracle_out *out) {
	/* bvh.hpp:259-263 */
	if (out->t)
		out->t[i] = t_near;
	if (out->valid)
		out->valid[i] = t_near < INFINITY;
	if (out->px)
		out->px[i] = o[0] + t_near * d[0];
	if (out->py)
		out->py[i] = o[1] + t_near * d[1];
	if (out->pz)
		out->pz[i] = o[2] + t_near * d[2];
	if (out->u)
		out->u[i] = u;
	if (out->v)
		out->v[i] = v;
	if (out->pid)
		out->pid[i] = pid;
}

/* nearest_tri, bvh.hpp:224-265.  visits (optional) counts node visits / tri tests per ray. */
static int trace_one(const oracle_bvh *b, const float *ray, uint64_t i, const oracle_out *out,
                     uint32_t *stack, uint32_t cap, uint32_t *visits) {
	const float *o = ray, *d = ray + 3;
	uint32_t sp = 0;
	stack[sp++] = 0; /* :229 */
	float t_near = INFINITY, bu = 0.0f, bv = 0.0f;
	uint32_t bpid = ORACLE_NO_TRI, nv = 0, nt = 0;
	while (sp > 0) {
		const onode *n = &b->nodes[stack[--sp]]; /* :235-236 */
		++nv;
		if (!slab(o, d, n->lo, n->hi))
			continue; /* :237-239 */
		if (n->is_leaf) {
			if (n->tri == ORACLE_NO_TRI)
				continue; /* :241-243 */
			float t, u, v;
			++nt;
			if (moller_trumbore(b->tris + 9ull * n->tri, o, d, &t, &u, &v)) {
				if (t < t_near) { /* :247 strict */
					t_near = t;
					bu = u;
					bv = v;
					bpid = n->tri;
				}
			}
		} else {
			if (sp + 2 > cap)
				return -1; /* the reference would overflow its 1024-entry stack here (:226) */
			stack[sp++] = n->left;  /* :255 */
			stack[sp++] = n->right; /* :256 => right is popped first */
		}
	}
	finish(o, d, t_near, bu, bv, bpid, i, out);
	if (visits) {
		visits[2 * i] = nv;
		visits[2 * i + 1] = nt;
	}
	return 0;
}

typedef struct {
	const oracle_bvh *b;
	const float *tris;
	uint64_t n_tris;
	const float *rays;
	uint64_t lo, hi;
	const oracle_out *out;
	uint32_t *visits;
	int brute;
	int rc;
} job;

static void brute_one(const float *tris, uint64_t n_tris, const float *ray, uint64_t i,
                      const oracle_out *out) {
	const float *o = ray, *d = ray + 3;
	float t_near = INFINITY, bu = 0.0f, bv = 0.0f;
	uint32_t bpid = ORACLE_NO_TRI;
	float lo[3], hi[3];
	for (uint64_t k = 0; k < n_tris; ++k) {
		const float *tv = tris + 9 * k;
		box_of_tri(tv, lo, hi);
		if (!slab(o, d, lo, hi))
			continue;
		float t, u, v;
		if (moller_trumbore(tv, o, d, &t, &u, &v) && t < t_near) {
			t_near = t;
			bu = u;
			bv = v;
			bpid = (uint32_t)k;
		}
	}
	finish(o, d, t_near, bu, bv, bpid, i, out);
}

static void *worker(void *p) {
	job *j = (job *)p;
	enum { CAP = 1 << 16 };
	uint32_t *stack = j->brute ? NULL : (uint32_t *)malloc(CAP * sizeof(uint32_t));
	for (uint64_t i = j->lo; i < j->hi; ++i) {
		if (j->brute)
			brute_one(j->tris, j->n_tris, j->rays + 6 * i, i, j->out);
		else if (trace_one(j->b, j->rays + 6 * i, i, j->out, stack, CAP, j->visits))
			j->rc = -1;
	}
	free(stack);
	return NULL;
}

static int run_jobs(job proto, uint64_t n, int n_threads) {
	if (n_threads < 1)
		n_threads = 1;
	if ((uint64_t)n_threads > n)
		n_threads = n ? (int)n : 1;
	job *jobs = (job *)malloc(n_threads * sizeof(job));
	pthread_t *th = (pthread_t *)malloc(n_threads * sizeof(pthread_t));
	for (int k = 0; k < n_threads; ++k) {
		jobs[k] = proto;
		jobs[k].lo = n * k / n_threads;
		jobs[k].hi = n * (k + 1) / n_threads;
		jobs[k].rc = 0;
		if (k)
			pthread_create(&th[k], NULL, worker, &jobs[k]);
	}
	worker(&jobs[0]);
	int rc = jobs[0].rc;
	for (int k = 1; k < n_threads; ++k) {
		pthread_join(th[k], NULL);
		rc |= jobs[k].rc;
	}
	free(jobs);
	free(th);
	return rc;
}

/* (A) the reference's structure and visit order.  visits may be NULL, else 2*n uint32 (nodes, tris).
 * CPUBackend::nearest_hits fans the rays out over hardware threads in contiguous chunks
 * (intersect_cpu.hpp:20-43); the result does not depend on the chunking. */
int oracle_trace(const oracle_bvh *b, const float *rays6, uint64_t n, const oracle_out *out,
                 uint32_t *visits, int n_threads) {
	job j;
	memset(&j, 0, sizeof j);
	j.b = b;
	j.rays = rays6;
	j.out = out;
	j.visits = visits;
	return run_jobs(j, n, n_threads);
}

/* (B) topology-free rule, O(rays * tris) */
int oracle_brute(const float *tris9, uint64_t n_tris, const float *rays6, uint64_t n,
                 const oracle_out *out, int n_threads) {
	job j;
	memset(&j, 0, sizeof j);
	j.tris = tris9;
	j.n_tris = n_tris;
	j.rays = rays6;
	j.out = out;
	j.brute = 1;
	return run_jobs(j, n, n_threads);
}

/* single-primitive helpers for the comparator's exact-tie replay (SURVEY.md section 8c rule 2) */
int oracle_intersect_tri(const float *tri9, const float *ray6, float *t, float *u, float *v) {
	return moller_trumbore(tri9, ray6, ray6 + 3, t, u, v);
}
int oracle_ray_box(const float *ray6, const float *lo, const float *hi) {
	return slab(ray6, ray6 + 3, lo, hi);
}
int oracle_candidate(const float *tri9, const float *ray6, float *t, float *u, float *v) {
	float lo[3], hi[3];
	box_of_tri(tri9, lo, hi);
	if (!slab(ray6, ray6 + 3, lo, hi))
		return 0;
	return moller_trumbore(tri9, ray6, ray6 + 3, t, u, v);
}
