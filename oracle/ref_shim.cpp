// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI shim around the UNMODIFIED reference CPU backend, compiled from the sources where they
// lie under /root/reference (see oracle/Makefile).  Nothing from the reference is copied into
// this repository: this file only #includes the reference's public umbrella header and calls its
// public API (select_backend / set_tris / nearest_hits<Tags...>), exactly as
// examples/triangle/main.cpp:7-22 does.  The output library lives in oracle/_ref/ (git-ignored).
//
// Used by: tests/ (parity checker), bench.py (--impl reference arm and the cpu_baseline leg).
#include <portableRT/portableRT.hpp>

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>

namespace {

// mask bit order = canonical tag order of the reference (hitreg.hpp:24-26): uv,t,primitive_id,p,valid
template <class H> constexpr uint32_t mask_of() {
	return (H::has_uv::value ? 1u : 0u) | (H::has_t::value ? 2u : 0u) |
	       (H::has_primitive_id::value ? 4u : 0u) | (H::has_p::value ? 8u : 0u) |
	       (H::has_valid::value ? 16u : 0u);
}

struct Layout {
	uint32_t stride;
	int32_t off_u, off_v, off_t, off_pid, off_valid, off_px, off_py, off_pz;
};

#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Winvalid-offsetof"
template <class H> Layout layout_of() {
	Layout l;
	l.stride = sizeof(H);
	l.off_u = H::has_uv::value ? (int32_t)offsetof(H, u) : -1;
	l.off_v = H::has_uv::value ? (int32_t)offsetof(H, v) : -1;
	l.off_t = H::has_t::value ? (int32_t)offsetof(H, t) : -1;
	l.off_pid = H::has_primitive_id::value ? (int32_t)offsetof(H, primitive_id) : -1;
	l.off_valid = H::has_valid::value ? (int32_t)offsetof(H, valid) : -1;
	l.off_px = H::has_p::value ? (int32_t)offsetof(H, px) : -1;
	l.off_py = H::has_p::value ? (int32_t)offsetof(H, py) : -1;
	l.off_pz = H::has_p::value ? (int32_t)offsetof(H, pz) : -1;
	return l;
}
#pragma GCC diagnostic pop

std::vector<portableRT::Ray> to_rays(const float *rays6, uint64_t n) {
	std::vector<portableRT::Ray> rays(n);
	static_assert(sizeof(portableRT::Ray) == 24, "Ray is 6 packed floats (core.hpp:19-22)");
	if (n)
		std::memcpy(rays.data(), rays6, n * sizeof(portableRT::Ray));
	return rays;
}

template <class... Tags> double run(const std::vector<portableRT::Ray> &rays, void *out) {
	auto t0 = std::chrono::steady_clock::now();
	auto hits = portableRT::nearest_hits<Tags...>(rays); // free function, nearest_hits_impl.hpp:28
	auto t1 = std::chrono::steady_clock::now();
	if (out && !hits.empty())
		std::memcpy(out, hits.data(), hits.size() * sizeof(hits[0]));
	return std::chrono::duration<double>(t1 - t0).count();
}

} // namespace

extern "C" {

// number of worker threads CPUBackend::nearest_hits spawns (intersect_cpu.hpp:26)
int ref_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

int ref_device_name(char *buf, size_t cap) {
	std::string s = portableRT::selected_backend ? portableRT::selected_backend->device_name() : "";
	if (cap) {
		std::strncpy(buf, s.c_str(), cap - 1);
		buf[cap - 1] = 0;
	}
	return (int)s.size();
}

// returns seconds spent inside set_tris (single-threaded BVH2::build, bvh.hpp:158-193)
double ref_set_tris(const float *tris9, uint64_t n) {
	portableRT::Tris tris(n);
	if (n)
		std::memcpy(tris.data(), tris9, n * sizeof(portableRT::Tri));
	auto t0 = std::chrono::steady_clock::now();
	portableRT::selected_backend->set_tris(tris);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

int ref_layout(uint32_t mask, Layout *out) {
	switch (mask) {
#define X(...)                                                                                     \
	case mask_of<portableRT::HitReg<ADD_FILTER(__VA_ARGS__)>>():                                   \
		*out = layout_of<portableRT::HitReg<ADD_FILTER(__VA_ARGS__)>>();                           \
		return 0;
		TAG_COMBOS
#undef X
	default:
		return -1;
	}
}

// Runs the reference's nearest_hits<Tags...> for the combo selected by `mask`; copies the AoS
// HitReg<Tags...> records into hits_out (may be NULL for timing only); *seconds = wall time of the
// nearest_hits call alone.  Returns 0, or -1 for an unknown mask.
int ref_nearest_hits(const float *rays6, uint64_t n, uint32_t mask, void *hits_out,
                     double *seconds) {
	auto rays = to_rays(rays6, n);
	double s = 0;
	switch (mask) {
#define X(...)                                                                                     \
	case mask_of<portableRT::HitReg<ADD_FILTER(__VA_ARGS__)>>():                                   \
		s = run<ADD_FILTER(__VA_ARGS__)>(rays, hits_out);                                          \
		break;
		TAG_COMBOS
#undef X
	default:
		return -1;
	}
	if (seconds)
		*seconds = s;
	return 0;
}

// zero-tag overload (backend.hpp:77-79) -> FullHitReg
int ref_nearest_hits_default(const float *rays6, uint64_t n, void *hits_out) {
	auto rays = to_rays(rays6, n);
	auto hits = portableRT::nearest_hits(rays);
	if (hits_out && !hits.empty())
		std::memcpy(hits_out, hits.data(), hits.size() * sizeof(hits[0]));
	return 0;
}

} // extern "C"
