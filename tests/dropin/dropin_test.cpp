// Drop-in acceptance program: the UNMODIFIED reference API (portableRT.hpp, patched only by the
// additive USE_CUDA arms of tools/patch_reference.py) with the reference's own CPU backend and
// this repo's CUDA backend compiled into ONE binary.  It does what examples/validation/main.cpp
// does (known-answer triangle, per-backend loop over available_backends() with select_backend)
// and then differentially compares every one of the 31 tag combinations of the CUDA backend with
// the CPU backend on the same rays, through both the free function and the member form.
//
// Built only where /root/reference exists (tests/dropin/Makefile -> oracle/_ref/dropin_test); the
// binary travels to the GPU box.  Without a GPU it checks registration/availability and exits 0.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <portableRT/portableRT.hpp>

using namespace portableRT;

static int g_fail = 0;
#define CHECK(cond, ...)                                                                           \
	do {                                                                                           \
		if (!(cond)) {                                                                             \
			++g_fail;                                                                              \
			std::printf("FAIL %s:%d: ", __FILE__, __LINE__);                                       \
			std::printf(__VA_ARGS__);                                                              \
			std::printf("\n");                                                                     \
		}                                                                                          \
	} while (0)

static Tris make_blob(int nu, int nv, float radius) {
	Tris tris;
	auto P = [&](int i, int j) {
		float th = 3.14159265f * j / nv, ph = 6.2831853f * (i % nu) / nu;
		float r = radius * (1.0f + 0.15f * std::sin(3 * th) * std::cos(4 * ph));
		return std::array<float, 3>{r * std::sin(th) * std::cos(ph), r * std::cos(th),
		                            r * std::sin(th) * std::sin(ph)};
	};
	for (int j = 0; j < nv; ++j)
		for (int i = 0; i < nu; ++i) {
			auto a = P(i, j), b = P(i + 1, j), c = P(i + 1, j + 1), d = P(i, j + 1);
			tris.push_back({a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]});
			tris.push_back({a[0], a[1], a[2], c[0], c[1], c[2], d[0], d[1], d[2]});
		}
	return tris;
}

static std::vector<Ray> make_rays(int w, int h, int n_random) {
	std::vector<Ray> rays;
	for (int y = h - 1; y >= 0; --y) // camera of examples/bunny/main.cpp:101-124
		for (int x = 0; x < w; ++x) {
			float sx = 0.05f * (float(x) / w - 0.5f), sy = 0.05f * (float(y) / h - 0.5f);
			Ray r;
			r.origin = {0.0013f, -0.0007f, -0.3f};
			std::array<float, 3> d = {sx - r.origin[0], sy - r.origin[1], 0.05f};
			float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
			r.direction = {d[0] / len, d[1] / len, d[2] / len};
			rays.push_back(r);
		}
	std::mt19937 g(7);
	std::uniform_real_distribution<float> U(-0.15f, 0.15f);
	std::normal_distribution<float> N;
	for (int i = 0; i < n_random; ++i) { // origins inside/around the mesh: negative-t hits occur
		Ray r;
		r.origin = {U(g), U(g), U(g)};
		r.direction = {N(g), N(g), N(g)};
		rays.push_back(r);
	}
	return rays;
}

struct Stats {
	size_t n = 0, valid = 0, ties = 0;
};

// compare CUDA hits with the slice of the CPU backend's full record (hitreg.hpp:63-85)
template <class... Tags>
static void compare(const char *what, const std::vector<HitReg<Tags...>> &got,
                    const std::vector<FullHitReg> &ref, Stats &st) {
	CHECK(got.size() == ref.size(), "%s: size %zu vs %zu", what, got.size(), ref.size());
	size_t bad = 0;
	for (size_t i = 0; i < got.size() && i < ref.size(); ++i) {
		const auto &g = got[i];
		const FullHitReg &r = ref[i];
		bool ok = true, tie = false;
		if constexpr (has_tag<filter::valid, Tags...>)
			ok &= g.valid == r.valid;
		if constexpr (has_tag<filter::t, Tags...>)
			ok &= (g.t == r.t);
		if constexpr (has_tag<filter::primitive_id, Tags...>)
			if (r.valid && g.primitive_id != r.primitive_id)
				tie = true; // judged below
		if constexpr (has_tag<filter::p, Tags...>)
			ok &= (std::memcmp(&g.px, &r.px, 4) == 0 && std::memcmp(&g.py, &r.py, 4) == 0 &&
			       std::memcmp(&g.pz, &r.pz, 4) == 0) || !(r.px == r.px && r.py == r.py && r.pz == r.pz);
		if constexpr (has_tag<filter::uv, Tags...>)
			if (r.valid && !tie) {
				if constexpr (has_tag<filter::primitive_id, Tags...>)
					ok &= (g.u == r.u && g.v == r.v);
			}
		if (tie) {
			// exact tie: same t from a different triangle (documented exception)
			if constexpr (has_tag<filter::t, Tags...>)
				ok &= (g.t == r.t);
			++st.ties;
		}
		if (!ok)
			++bad;
		++st.n;
		st.valid += r.valid;
	}
	CHECK(bad == 0, "%s: %zu of %zu rays differ from the CPU backend", what, bad, got.size());
}

template <class... Tags>
static void one_combo(const char *name, CUDABackend *cuda, const std::vector<Ray> &rays,
                      const std::vector<FullHitReg> &ref, Stats &st) {
	auto a = nearest_hits<Tags...>(rays);        // free function via std::visit on the variant
	auto b = cuda->nearest_hits<Tags...>(rays);  // member form on the concrete type
	compare<Tags...>(name, a, ref, st);
	CHECK(a.size() == b.size(), "%s: free function and member form disagree in size", name);
	Stats dummy;
	compare<Tags...>(name, b, ref, dummy);
}

int main() {
	Backend *cpu = nullptr, *cuda = nullptr;
	for (auto *b : all_backends()) {
		if (b->name() == "CPU")
			cpu = b;
		if (b->name() == "CUDA")
			cuda = b;
	}
	CHECK(cpu && cuda, "both backends must be compiled in (all_backends has %zu)", all_backends().size());
	bool avail = false;
	for (auto *b : available_backends())
		avail |= (b == cuda);
	std::printf("compiled backends: %zu, available: %zu, CUDA available: %d\n", all_backends().size(),
	            available_backends().size(), (int)avail);
	CHECK(selected_backend == cpu, "CPU registers first and stays the default selection");
	if (!avail) {
		CHECK(!cuda->is_available(), "availability must be consistent");
		std::printf("%s (no compute-capability-10.x GPU here: registration checks only)\n",
		            g_fail ? "FAILED" : "PASS");
		return g_fail ? 1 : 0;
	}

	// ---- examples/validation/main.cpp:66-111 for every available backend
	const float eps = 1e-4f;
	for (auto *b : available_backends()) {
		select_backend(b);
		std::array<float, 9> v = {-1, -1, 0, 1, -1, 0, 0, 1, 0};
		Ray hit{{0.1f, 0, -1}, {0, 0, 1}}, miss{{-2, 0, -1}, {0, 0, 1}};
		selected_backend->set_tris({v});
		auto h1 = nearest_hits({hit});
		auto h2 = nearest_hits({miss});
		CHECK(h1[0].valid && std::fabs(h1[0].t - 1) < eps && std::fabs(h1[0].u - 0.3f) < eps &&
		          std::fabs(h1[0].v - 0.5f) < eps && h1[0].primitive_id == 0 &&
		          std::fabs(h1[0].px - 0.1f) < eps && std::fabs(h1[0].py) < eps && std::fabs(h1[0].pz) < eps,
		      "%s: known-answer hit", b->name().c_str());
		CHECK(!h2[0].valid, "%s: known-answer miss", b->name().c_str());
		CHECK(nearest_hits<filter::valid>({hit})[0].valid && !nearest_hits<filter::valid>({miss})[0].valid,
		      "%s: filtered valid", b->name().c_str());
		std::printf("backend %-5s on '%s': known answers ok\n", b->name().c_str(), b->device_name().c_str());
	}

	// ---- differential: CPU backend vs CUDA backend, all 31 tag combinations
	Tris tris = make_blob(96, 64, 0.1f);
	std::vector<Ray> rays = make_rays(256, 144, 30000);
	select_backend(cpu);
	auto t0 = std::chrono::steady_clock::now();
	cpu->set_tris(tris);
	auto t1 = std::chrono::steady_clock::now();
	std::vector<FullHitReg> ref = nearest_hits(rays);
	auto t2 = std::chrono::steady_clock::now();
	select_backend(cuda);
	select_backend(cuda); // shutdown()+init() on the selected backend must work (validation/main.cpp:242)
	auto t3 = std::chrono::steady_clock::now();
	cuda->set_tris(tris);
	auto t4 = std::chrono::steady_clock::now();
	std::vector<FullHitReg> full = nearest_hits(rays);
	auto t5 = std::chrono::steady_clock::now();
	auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
	std::printf("%zu tris, %zu rays: CPU build %.1f ms trace %.1f ms | CUDA build %.1f ms trace %.1f ms "
	            "(API level, host vectors)\n", tris.size(), rays.size(), ms(t0, t1), ms(t1, t2), ms(t3, t4), ms(t4, t5));
	CHECK(ms(t3, t4) < 1000 && ms(t4, t5) < 1000, "validation perf gate (validation/main.cpp:230-233)");

	Stats st;
	compare<ALL_TAGS>("full (zero-tag overload)", full, ref, st);
	auto *cb = static_cast<CUDABackend *>(cuda);
#define X(...) one_combo<ADD_FILTER(__VA_ARGS__)>(#__VA_ARGS__, cb, rays, ref, st);
	TAG_COMBOS
#undef X
	size_t neg = 0;
	for (auto &h : ref)
		neg += h.valid && h.t < 0;
	CHECK(neg > 100, "the ray set must exercise negative-t hits (got %zu)", neg);

	// extensions of the concrete type: SAH-optimised tree (results must not move) and the opt-in
	// watertight test (may differ from the reference on edge-grazing rays only)
	cb->set_tree_optimisation(1, 3);
	cuda->set_tris(tris);
	compare<ALL_TAGS>("full, tree optimised inside set_tris", nearest_hits(rays), ref, st);
	cb->set_watertight(true);
	cuda->set_tris(tris);
	{
		auto w = cb->nearest_hits<filter::valid, filter::t>(rays);
		size_t differs = 0;
		for (size_t i = 0; i < w.size(); ++i)
			differs += w[i].valid != ref[i].valid;
		CHECK(differs * 1000 < rays.size(), "watertight mode: %zu of %zu rays change validity", differs,
		      rays.size());
		std::printf("watertight mode: valid differs from the reference on %zu of %zu rays\n", differs,
		            rays.size());
	}
	cb->set_watertight(false);
	cb->set_tree_optimisation(3);
	cuda->set_tris(tris);
	compare<ALL_TAGS>("full, defaults restored", nearest_hits(rays), ref, st);

	// empty scene and empty ray list (bvh.hpp:136-137)
	cuda->set_tris({});
	auto e = nearest_hits<filter::valid, filter::t>(rays);
	size_t any = 0;
	for (auto &h : e)
		any += h.valid || !std::isinf(h.t);
	CHECK(any == 0, "empty scene: all rays must miss");
	CHECK(nearest_hits(std::vector<Ray>{}).empty(), "empty ray list");

	std::printf("{\"rays_compared\": %zu, \"valid\": %zu, \"exact_ties\": %zu, \"negative_t\": %zu, "
	            "\"failures\": %d}\n", st.n, st.valid, st.ties, neg, g_fail);
	std::printf("%s\n", g_fail ? "FAILED" : "PASS");
	return g_fail ? 1 : 0;
}
