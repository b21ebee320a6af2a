// intersect_cuda.hpp -- the B200-native backend as a portableRT plugin class.
//
// Drop this file next to the reference's other backend headers (include/portableRT/), put
// prt_b200.h on the include path, define USE_CUDA, apply the four additive arms listed in
// INTEGRATION.md (tools/patch_reference.py does it mechanically) and link libprt_b200.so.  Nothing
// else of the reference changes: select_backend(), available_backends(), set_tris(const Tris&) and
// both the free-function and the member form of nearest_hits<filter::...>(rays) keep their exact
// signatures and semantics.
//
// Shape follows the reference's own backends: class + static instance + self-registration like
// CPUBackend (include/portableRT/intersect_cpu.hpp:8-10,55) and OptiXBackend
// (include/portableRT/intersect_optix.hpp:54-56,138); nearest_hits<Tags...> is an inline member
// template like OptiX's (intersect_optix.hpp:64-119), so no per-combination instantiation list is
// needed on the host side -- the 31 device specialisations live behind the tag mask of the C ABI.
// All work happens on the GPU inside libprt_b200.so; there is no CPU fallback in this class.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "backend.hpp"
#include "core.hpp"
#include "prt_b200.h"

namespace portableRT {

namespace cuda_detail {
// std::vector<H>(n) value-initialises: for the result of a 100 M-ray call that is 1.6 GB written
// (and page-faulted) by one thread before the first ray is traced, only to be overwritten.  The
// records are trivial, so the vector's storage is reserved and its size set without touching it
// (libstdc++: vector<T> : protected _Vector_base<T> whose _M_impl holds start/finish/end); the
// library's staging threads are then the first to touch each page, in parallel.  Other standard
// libraries, or -DPRT_B200_NO_VECTOR_HACK, take the plain value-initialising path.
template <class T> struct VectorAccess : std::vector<T> {
	static void set_size(std::vector<T> &v, std::size_t n) {
#if defined(__GLIBCXX__) && !defined(PRT_B200_NO_VECTOR_HACK) && !defined(_GLIBCXX_DEBUG)
		static_assert(std::is_trivially_copyable<T>::value && std::is_trivially_destructible<T>::value,
		              "HitReg records are trivial");
		v.reserve(n);
		auto &a = static_cast<VectorAccess &>(v);
		a._M_impl._M_finish = a._M_impl._M_start + n;
#else
		v.resize(n);
#endif
	}
};
} // namespace cuda_detail

class CUDABackend : public InvokableBackend<CUDABackend> {
  public:
	CUDABackend() : InvokableBackend("CUDA") { static RegisterBackend reg(*this); }
	~CUDABackend() { shutdown(); }

	// Evaluated during static initialisation (backend.hpp:85-95): cheap, never throws, false
	// when there is no driver, no GPU or no compute-capability-10.x device.
	bool is_available() const override { return prt_b200_device_count() > 0; }

	// Called by select_backend (src/backend.cpp:50-56), possibly repeatedly and on an already
	// initialised object (examples/validation/main.cpp:242): idempotent.
	void init() override {
		if (m_ctx)
			return;
		if (prt_b200_create(&m_ctx, -1) != PRT_OK) {
			m_ctx = nullptr;
			// reference backends report and carry on (intersect_optix.hpp:34-50)
			std::fprintf(stderr, "portableRT CUDA backend: %s\n", prt_b200_last_error(nullptr));
			return;
		}
		if (const char *e = std::getenv("PRT_B200_PRUNE")) {
			prt_trace_opts o{std::atoi(e), 1e-4f, 64.0f};
			prt_b200_set_trace_opts(m_ctx, &o);
		}
	}

	void shutdown() override {
		prt_b200_destroy(m_ctx); // NULL-safe
		m_ctx = nullptr;
	}

	// Replaces the previous scene; an empty list is legal (every ray then misses, bvh.hpp:136-137).
	void set_tris(const Tris &tris) override {
		static_assert(sizeof(Tri) == 36, "Tri is 9 packed floats (core.hpp:24)");
		need_ctx();
		check(prt_b200_set_tris(m_ctx, tris.empty() ? nullptr : tris.data()->data(), tris.size()));
	}

	// The move overload the reference leaves as a TODO (backend.hpp:18): the triangles live on the
	// device after the call, so the caller's storage is released.
	void set_tris(Tris &&tris) {
		set_tris(static_cast<const Tris &>(tris));
		Tris().swap(tris);
	}

	std::string device_name() const override {
		char buf[256] = "unavailable";
		if (m_ctx)
			prt_b200_device_name(m_ctx, buf, sizeof buf);
		return buf;
	}

	template <class... Tags>
	std::vector<HitReg<Tags...>> nearest_hits(const std::vector<Ray> &rays) {
		using H = HitReg<Tags...>;
		static_assert(sizeof(Ray) == 24, "Ray is 6 packed floats (core.hpp:19-22)");
		std::vector<H> hits;
		constexpr uint32_t mask = (H::has_uv::value ? PRT_TAG_UV : 0u) | (H::has_t::value ? PRT_TAG_T : 0u) |
		                          (H::has_primitive_id::value ? PRT_TAG_PID : 0u) |
		                          (H::has_p::value ? PRT_TAG_P : 0u) |
		                          (H::has_valid::value ? PRT_TAG_VALID : 0u);
		if (mask == 0 || rays.empty()) {
			hits.resize(rays.size()); // HitReg<> has no fields to fill
			return hits;
		}
		need_ctx();
		cuda_detail::VectorAccess<H>::set_size(hits, rays.size()); // every requested field is written below
		const prt_hit_layout lay = layout<H>();
		check(prt_b200_nearest_hits(m_ctx, rays.data()->origin.data(), rays.size(), mask, &lay,
		                            hits.data()));
		return hits;
	}

	prt_b200 *context() { return m_ctx; } // for benchmarks that use the device-resident entry points

	// Extensions beyond the reference's Backend interface (reached through the concrete type, like
	// the member form of nearest_hits); the same knobs exist as PRT_B200_* environment variables.
	// SAH optimisation of the BVH by treelet restructuring: 0 never, 1 inside every set_tris,
	// 2 lazily for scenes that keep being traced, 3 (default) lazily plus temporal reuse (set_tris
	// with the same triangle count refits the optimised topology).  Never changes a result.
	void set_tree_optimisation(int mode, int passes = 2) {
		need_ctx();
		check(prt_b200_set_tree_optimisation(m_ctx, mode, passes));
	}
	// false (default): the reference's intersect_tri arithmetic (core.hpp:27-65), results identical
	// to the CPU backend; true: watertight test (rays cannot slip between triangles sharing an edge
	// or vertex), results differ from the reference on exactly those rays.  From the next set_tris.
	void set_watertight(bool on) {
		need_ctx();
		check(prt_b200_set_triangle_test(m_ctx, on ? 1 : 0));
	}

  private:
#if defined(__GNUC__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Winvalid-offsetof"
#endif
	// sizeof/offsetof of the reference's own record type (hitreg.hpp:29-44): nothing hard-coded
	template <class H> static prt_hit_layout layout() {
		prt_hit_layout l;
		l.stride = static_cast<uint32_t>(sizeof(H));
		l.off_u = H::has_uv::value ? static_cast<int32_t>(offsetof(H, u)) : -1;
		l.off_v = H::has_uv::value ? static_cast<int32_t>(offsetof(H, v)) : -1;
		l.off_t = H::has_t::value ? static_cast<int32_t>(offsetof(H, t)) : -1;
		l.off_pid = H::has_primitive_id::value ? static_cast<int32_t>(offsetof(H, primitive_id)) : -1;
		l.off_valid = H::has_valid::value ? static_cast<int32_t>(offsetof(H, valid)) : -1;
		l.off_px = H::has_p::value ? static_cast<int32_t>(offsetof(H, px)) : -1;
		l.off_py = H::has_p::value ? static_cast<int32_t>(offsetof(H, py)) : -1;
		l.off_pz = H::has_p::value ? static_cast<int32_t>(offsetof(H, pz)) : -1;
		return l;
	}
#if defined(__GNUC__)
#pragma GCC diagnostic pop
#endif

	void need_ctx() {
		if (!m_ctx)
			init();
		if (!m_ctx) // the only exception type of this API path (nearest_hits_impl.hpp:32)
			throw std::runtime_error(std::string("portableRT CUDA backend unavailable: ") +
			                         prt_b200_last_error(nullptr));
	}
	void check(int rc) {
		if (rc != PRT_OK)
			throw std::runtime_error(std::string("portableRT CUDA backend: ") +
			                         prt_b200_last_error(m_ctx));
	}

	prt_b200 *m_ctx = nullptr;
};

static CUDABackend cuda_backend;

} // namespace portableRT
