// TEST INFRASTRUCTURE ONLY -- host-side logic emulator for the CPU-only test suite.
//
// The product (libprt_b200.so) runs exclusively on the GPU and has no CPU path.  This file lets the
// `-m "not gpu"` tests exercise, without a GPU, the exact source the kernels are built from:
//   * prt::karras_node / prt::morton3 / prt::tri_box / prt::box_union  (prt_math.cuh)
//   * prt::traverse<...>                                                (prt_traverse.cuh)
// by driving them from sequential host loops (std::sort instead of the device radix sort, a
// post-order walk instead of the atomic refit).  It is compiled by g++ into tests/emu/libprt_emu.so,
// loaded only by tests/, and never by portablert_b200/.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../portablert_b200/csrc/prt_traverse.cuh"
#include "../../portablert_b200/csrc/prt_treelet.cuh"

using namespace prt;

namespace {
struct Emu {
	std::vector<Node> nodes;
	std::vector<Node4> nodes4;
	std::vector<TriRec> tris;
	uint64_t n = 0;
	int32_t root = 0;
	float absmax[3] = {0.f, 0.f, 0.f};
	void widen() { // same gathering as k_wide (build.cu); the quantiser itself is shared source
		nodes4.resize(nodes.size());
		for (size_t i = 0; i < nodes.size(); ++i) {
			const Node &nd = nodes[i];
			WideChild ch[4];
			int cnt = 0;
			for (int side = 0; side < 2; ++side) {
				const int32_t c = side ? nd.child1 : nd.child0;
				const float *lo = side ? nd.lo1 : nd.lo0, *hi = side ? nd.hi1 : nd.hi0;
				if (c >= 0) {
					const Node &g = nodes[c];
					for (int a = 0; a < 3; ++a) {
						ch[cnt].lo[a] = g.lo0[a];
						ch[cnt].hi[a] = g.hi0[a];
						ch[cnt + 1].lo[a] = g.lo1[a];
						ch[cnt + 1].hi[a] = g.hi1[a];
					}
					ch[cnt].ref = g.child0;
					ch[cnt + 1].ref = g.child1;
					cnt += 2;
				} else {
					for (int a = 0; a < 3; ++a) {
						ch[cnt].lo[a] = lo[a];
						ch[cnt].hi[a] = hi[a];
					}
					ch[cnt].ref = c;
					cnt += 1;
				}
			}
			nodes4[i] = make_node4(ch, cnt);
		}
	}
	void bounds() {
		widen();
		for (int a = 0; a < 3; ++a)
			absmax[a] = 0.f;
		if (nodes.empty())
			return;
		const Node &r = nodes[root];
		for (int a = 0; a < 3; ++a) {
			absmax[a] = fmaxf(fabsf(r.lo0[a]), fabsf(r.hi0[a]));
			absmax[a] = fmaxf(absmax[a], fmaxf(fabsf(r.lo1[a]), fabsf(r.hi1[a])));
		}
	}
};

Box refit(Emu &e, const std::vector<Box> &leaf, int32_t ref) {
	if (ref < 0)
		return leaf[~ref];
	Node &nd = e.nodes[ref];
	Box b0 = refit(e, leaf, nd.child0);
	Box b1 = refit(e, leaf, nd.child1);
	for (int a = 0; a < 3; ++a) {
		nd.lo0[a] = b0.lo[a];
		nd.hi0[a] = b0.hi[a];
		nd.lo1[a] = b1.lo[a];
		nd.hi1[a] = b1.hi[a];
	}
	return box_union(b0, b1);
}
} // namespace

// sequential stand-in for k_treelet (build.cu): children before parents, same per-node rule
int32_t treelet_walk(Emu &e, int32_t x, std::vector<int32_t> &depth, int32_t &count, uint64_t &changed,
                     bool strict) {
	const int32_t c0 = e.nodes[x].child0, c1 = e.nodes[x].child1;
	int32_t n0 = 1, n1 = 1;
	const int32_t d0 = c0 < 0 ? 0 : treelet_walk(e, c0, depth, n0, changed, strict);
	const int32_t d1 = c1 < 0 ? 0 : treelet_walk(e, c1, depth, n1, changed, strict);
	count = n0 + n1;
	if (count >= TREELET_N)
		changed += treelet_optimise(e.nodes.data(), x, depth.data(), strict) ? 1 : 0;
	else
		depth[x] = 1 + std::max(d0, d1);
	return depth[x];
}

extern "C" {

// `passes` rounds of treelet restructuring; returns the height of the tree, *changed = treelets
// whose topology was replaced in the last pass
int32_t emu_treelet(void *h, int passes, int strict, uint64_t *changed) {
	Emu *e = static_cast<Emu *>(h);
	if (e->n < (uint64_t)TREELET_N)
		return 0;
	std::vector<int32_t> depth(e->nodes.size(), 0);
	int32_t d = 0;
	for (int p = 0; p < passes; ++p) {
		int32_t cnt = 0;
		uint64_t ch = 0;
		d = treelet_walk(*e, e->root, depth, cnt, ch, strict != 0);
		if (changed)
			*changed = ch;
	}
	e->bounds();
	return d;
}

// The split searches the device kernel spreads over lanes (treelet_best_split_quarter for 4 lanes
// per subset, direct indexing p = 2(i+1) for the full set, merged on the key (cost, split)) must
// select exactly what the sequential recurrence selects.  copt: TREELET_SETS costs (any values,
// ties and infinities included).  Returns the number of subsets on which they disagree.
int emu_check_split_search(const float *copt) {
	int bad = 0;
	for (int s = 1; s < TREELET_SETS; ++s) {
		if ((s & (s - 1)) == 0)
			continue;
		float want;
		int want_p;
		treelet_best_split(copt, s, 0, 1, want, want_p);
		float best = INFINITY;
		int bp = 0xff;
		for (int sub = 0; sub < 4; ++sub) { // lanes merge with: smaller cost, then smaller split
			float b;
			int p;
			treelet_best_split_quarter(copt, s, sub, b, p);
			if (b < best || (b == best && p < bp)) {
				best = b;
				bp = p;
			}
		}
		bad += !(bp == want_p && (best == want || (bp == 0xff)));
		if (s == TREELET_SETS - 1) { // the full set as the kernel indexes it
			float fb = INFINITY;
			int fp = 0xff;
			for (int lane = 0; lane < 32; ++lane) {
				float lb = INFINITY;
				int lp = 0xff;
				for (int i = 0; i < 2; ++i) {
					const int p = (lane + 32 * i + 1) << 1;
					if (p < s) {
						const float c = copt[p] + copt[s ^ p];
						if (c < lb) {
							lb = c;
							lp = p;
						}
					}
				}
				if (lb < fb || (lb == fb && lp < fp)) {
					fb = lb;
					fp = lp;
				}
			}
			bad += !(fp == want_p && (fb == want || fp == 0xff));
		}
	}
	return bad;
}

// sequential stand-in for k_refit (build.cu): same topology, new vertices
void emu_refit(void *h, const float *tris9, int vertex_form) {
	Emu *e = static_cast<Emu *>(h);
	if (e->n < 2)
		return;
	std::vector<Box> leaf(e->n);
	for (uint64_t j = 0; j < e->n; ++j) {
		TriRec &r = e->tris[j];
		const float *t = tris9 + 9ull * r.prim;
		leaf[j] = tri_box(t);
		for (int a = 0; a < 3; ++a) {
			r.v0[a] = t[a];
			r.e1[a] = vertex_form ? t[3 + a] : t[3 + a] - t[a];
			r.e2[a] = vertex_form ? t[6 + a] : t[6 + a] - t[a];
		}
		r.lox = leaf[j].lo[0];
		r.loy = leaf[j].lo[1];
		r.loz = leaf[j].lo[2];
		r.hix = leaf[j].hi[0];
		r.hiy = leaf[j].hi[1];
		r.hiz = leaf[j].hi[2];
	}
	refit(*e, leaf, e->root);
	e->bounds();
}

void *emu_build(const float *tris9, uint64_t n, int bits, int vertex_form) {
	Emu *e = new Emu();
	e->n = n;
	if (n == 0)
		return e;
	float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (uint64_t i = 0; i < n; ++i) {
		Box b = tri_box(tris9 + 9 * i);
		for (int a = 0; a < 3; ++a) {
			float c = 0.5f * b.lo[a] + 0.5f * b.hi[a];
			cmin[a] = fminf(cmin[a], c);
			cmax[a] = fmaxf(cmax[a], c);
		}
	}
	const float cells = (float)(1u << bits);
	const uint32_t qmax = (1u << bits) - 1u;
	std::vector<uint64_t> keys(n);
	std::vector<uint32_t> idx(n);
	for (uint64_t i = 0; i < n; ++i) {
		Box b = tri_box(tris9 + 9 * i);
		uint32_t q[3];
		for (int a = 0; a < 3; ++a) {
			float ext = cmax[a] - cmin[a];
			float scale = (ext > 0.0f && ext < INFINITY) ? cells / ext : 0.0f;
			float c = 0.5f * b.lo[a] + 0.5f * b.hi[a];
			float x = (c - cmin[a]) * scale;
			q[a] = (x > 0.0f) ? (uint32_t)fminf(x, (float)qmax) : 0u;
		}
		keys[i] = morton3(q[0], q[1], q[2]);
		idx[i] = (uint32_t)i;
	}
	std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
	std::vector<uint64_t> skeys(n);
	std::vector<Box> leaf(n);
	e->tris.resize(n);
	for (uint64_t j = 0; j < n; ++j) {
		const float *t = tris9 + 9ull * idx[j];
		skeys[j] = keys[idx[j]];
		leaf[j] = tri_box(t);
		TriRec &r = e->tris[j];
		std::memset(&r, 0, sizeof r);
		for (int a = 0; a < 3; ++a) {
			r.v0[a] = t[a];
			r.e1[a] = vertex_form ? t[3 + a] : t[3 + a] - t[a];
			r.e2[a] = vertex_form ? t[6 + a] : t[6 + a] - t[a];
		}
		r.prim = idx[j];
		r.lox = leaf[j].lo[0];
		r.loy = leaf[j].lo[1];
		r.loz = leaf[j].lo[2];
		r.hix = leaf[j].hi[0];
		r.hiy = leaf[j].hi[1];
		r.hiz = leaf[j].hi[2];
	}
	if (n == 1) {
		e->nodes.resize(1);
		Node &nd = e->nodes[0];
		std::memset(&nd, 0, sizeof nd);
		for (int a = 0; a < 3; ++a) {
			nd.lo0[a] = nd.lo1[a] = leaf[0].lo[a];
			nd.hi0[a] = nd.hi1[a] = leaf[0].hi[a];
		}
		nd.child0 = ~0;
		nd.child1 = ~1;
		TriRec dummy = e->tris[0]; // zero-area: det == 0, never hit
		for (int a = 0; a < 3; ++a) // (vertex form: all three vertices coincide)
			dummy.e1[a] = dummy.e2[a] = vertex_form ? dummy.v0[a] : 0.0f;
		dummy.prim = 0xffffffffu;
		e->tris.push_back(dummy);
		e->bounds();
		return e;
	}
	e->nodes.resize(n - 1);
	std::memset(e->nodes.data(), 0, (n - 1) * sizeof(Node));
	for (int64_t i = 0; i < (int64_t)n - 1; ++i) {
		int32_t l, r;
		karras_node(skeys.data(), (int64_t)n, i, l, r);
		e->nodes[i].child0 = l;
		e->nodes[i].child1 = r;
	}
	refit(*e, leaf, 0);
	e->bounds();
	return e;
}

void emu_free(void *h) { delete static_cast<Emu *>(h); }
uint64_t emu_num_nodes(void *h) { return static_cast<Emu *>(h)->nodes.size(); }
void emu_download_wide(void *h, void *nodes4) {
	Emu *e = static_cast<Emu *>(h);
	if (nodes4 && !e->nodes4.empty())
		std::memcpy(nodes4, e->nodes4.data(), e->nodes4.size() * sizeof(Node4));
}
void emu_download(void *h, void *nodes, void *tris) {
	Emu *e = static_cast<Emu *>(h);
	if (nodes && !e->nodes.empty())
		std::memcpy(nodes, e->nodes.data(), e->nodes.size() * sizeof(Node));
	if (tris && !e->tris.empty())
		std::memcpy(tris, e->tris.data(), e->n * sizeof(TriRec));
}
// load an externally built tree (e.g. one downloaded from the GPU) for host-side checking
void *emu_load(const void *nodes, uint64_t n_nodes, const void *tris, uint64_t n_tris, int32_t root) {
	Emu *e = new Emu();
	e->root = root;
	e->n = n_tris;
	e->nodes.resize(n_nodes);
	e->tris.resize(n_tris + 1);
	std::memset(&e->tris[n_tris], 0, sizeof(TriRec));
	if (n_nodes)
		std::memcpy(e->nodes.data(), nodes, n_nodes * sizeof(Node));
	if (n_tris)
		std::memcpy(e->tris.data(), tris, n_tris * sizeof(TriRec));
	e->bounds();
	return e;
}

// SoA outputs like the device entry point; counts (2 per ray) may be NULL.  anyhit=1 emulates the
// `valid`-only specialisation.
void emu_trace(void *h, const float *rays6, uint64_t n, int prune, float slack_rel, float slack_ulps,
               int anyhit, int fast, int wide, int wt, float *t, float *u, float *v, uint32_t *pid, uint8_t *valid, float *p,
               uint32_t *counts, uint8_t *fastflag) {
	Emu *e = static_cast<Emu *>(h);
	TraverseOpts o{prune, slack_rel, slack_ulps};
	for (uint64_t i = 0; i < n; ++i) {
		RayC r = make_ray(rays6 + 6 * i);
		Hit hit;
		const FastRay fr = make_fast_ray(r, e->absmax, wide != 0 && !wt);
		const bool f = fast && fr.ok; // rays that do not qualify go to the exact kernel
		const Node4 *n4 = e->nodes4.data();
		if (wt && anyhit && f)
			traverse<true, false, false, false, true, false, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else if (wt && anyhit)
			traverse<true, false, false, false, false, false, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else if (wt && f)
			traverse<false, true, true, true, true, false, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else if (wt)
			traverse<false, true, true, true, false, false, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else if (wide && f && anyhit)
			traverse<true, false, false, false, true, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit, n4);
		else if (wide && f)
			traverse<false, true, true, true, true, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit, n4);
		else if (anyhit && f)
			traverse<true, false, false, false, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else if (anyhit)
			traverse<true, false, false, false, false>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else if (f)
			traverse<false, true, true, true, true>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		else
			traverse<false, true, true, true, false>(e->nodes.data(), e->tris.data(), e->n, e->root, r, fr, o, hit);
		if (fastflag)
			fastflag[i] = f;
		if (t)
			t[i] = hit.t;
		if (u)
			u[i] = hit.u;
		if (v)
			v[i] = hit.v;
		if (pid)
			pid[i] = hit.prim;
		if (valid)
			valid[i] = hit.t < INFINITY;
		if (p) {
			p[3 * i] = r.o[0] + hit.t * r.d[0];
			p[3 * i + 1] = r.o[1] + hit.t * r.d[1];
			p[3 * i + 2] = r.o[2] + hit.t * r.d[2];
		}
		if (counts) {
			counts[2 * i] = hit.n_nodes;
			counts[2 * i + 1] = hit.n_tris;
		}
	}
}

} // extern "C"
