// End-to-end timing of the reference's REAL C++ signature with this backend plugged in:
//     select_backend(&cuda_backend); cuda_backend.set_tris(tris);
//     std::vector<HitReg<Tags...>> hits = nearest_hits<Tags...>(std::vector<Ray>)
// (README flow, examples/simple/main.cpp:7-23; free function nearest_hits_impl.hpp:28-36), pageable
// std::vector in, a fresh std::vector out, every call.  bench.py writes the workload's triangles
// and rays to binary files, runs this program and reports its per-call wall time as
// `e2e.cxx_plugin`; the checksums let it verify that the hits are those of its own run.
//
//   bench_cxx TRIS.bin RAYS.bin MASK STEPS WARMUP     MASK: 6 = t,primitive_id  18 = t,valid  31 = all
//
// With PRT_B200_GPUS=N in the environment the same unchanged program spreads over N GPUs.
// Built only where /root/reference exists (tests/dropin/Makefile -> oracle/_ref/bench_cxx).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <portableRT/portableRT.hpp>

using namespace portableRT;

template <class T> static std::vector<T> read_file(const char *path) {
	std::vector<T> v;
	FILE *f = std::fopen(path, "rb");
	if (!f) {
		std::fprintf(stderr, "cannot open %s\n", path);
		std::exit(2);
	}
	std::fseek(f, 0, SEEK_END);
	const long bytes = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	v.resize((size_t)bytes / sizeof(T));
	if (std::fread(v.data(), sizeof(T), v.size(), f) != v.size()) {
		std::fprintf(stderr, "short read on %s\n", path);
		std::exit(2);
	}
	std::fclose(f);
	return v;
}

struct Sums {
	unsigned long long valid = 0, pid_sum = 0;
	uint32_t t_xor = 0;
};

template <class H> static Sums checksum(const std::vector<H> &hits) {
	Sums s;
	for (const H &h : hits) {
		uint32_t bits = 0;
		bool valid = false;
		if constexpr (H::has_t::value) {
			const float t = h.t;
			std::memcpy(&bits, &t, 4);
			valid = t < INFINITY;
		}
		if constexpr (H::has_valid::value)
			valid = h.valid;
		s.t_xor ^= bits;
		s.valid += valid ? 1 : 0;
		if constexpr (H::has_primitive_id::value)
			if (valid)
				s.pid_sum += h.primitive_id;
	}
	return s;
}

template <class... Tags>
static void run(const std::vector<Ray> &rays, int steps, int warmup, double set_tris_ms, size_t n_tris) {
	using clk = std::chrono::steady_clock;
	std::vector<double> ms;
	Sums s;
	for (int k = 0; k < warmup + steps; ++k) {
		const auto t0 = clk::now();
		auto hits = nearest_hits<Tags...>(rays); // the free function: std::visit on the selection
		const double dt = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
		if (k >= warmup)
			ms.push_back(dt);
		if (k == warmup + steps - 1)
			s = checksum(hits);
	}
	double mean = 0, best = 1e300;
	for (double x : ms) {
		mean += x / ms.size();
		best = x < best ? x : best;
	}
	std::printf("{\"api\": \"portableRT::nearest_hits<Tags...>(const std::vector<Ray>&) -> std::vector<HitReg<Tags...>>\", "
	            "\"backend\": \"%s\", \"device\": \"%s\", \"n_rays\": %zu, \"n_tris\": %zu, \"record_bytes\": %zu, "
	            "\"steps\": %d, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"set_tris_ms\": %.3f, "
	            "\"valid\": %llu, \"pid_sum\": %llu, \"t_xor\": %u}\n",
	            selected_backend->name().c_str(), selected_backend->device_name().c_str(), rays.size(),
	            n_tris, sizeof(HitReg<Tags...>), steps, mean, best, set_tris_ms, s.valid, s.pid_sum, s.t_xor);
}

int main(int argc, char **argv) {
	if (argc < 6) {
		std::fprintf(stderr, "usage: %s TRIS.bin RAYS.bin MASK STEPS WARMUP\n", argv[0]);
		return 2;
	}
	static_assert(sizeof(Tri) == 36 && sizeof(Ray) == 24, "packed records (core.hpp:19-25)");
	const Tris tris = read_file<Tri>(argv[1]);
	const std::vector<Ray> rays = read_file<Ray>(argv[2]);
	const int mask = std::atoi(argv[3]), steps = std::atoi(argv[4]), warmup = std::atoi(argv[5]);
#ifdef USE_CUDA
	if (!cuda_backend.is_available()) {
		std::printf("{\"unavailable\": \"no compute-capability-10.x GPU\"}\n");
		return 0;
	}
	select_backend(&cuda_backend);
#else
#error "built with the CUDA backend (tests/dropin/Makefile)"
#endif
	const auto t0 = std::chrono::steady_clock::now();
	selected_backend->set_tris(tris);
	const double st = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	if (mask == 6)
		run<filter::t, filter::primitive_id>(rays, steps, warmup, st, tris.size());
	else if (mask == 18)
		run<filter::t, filter::valid>(rays, steps, warmup, st, tris.size());
	else
		run<filter::uv, filter::t, filter::primitive_id, filter::p, filter::valid>(rays, steps, warmup, st,
		                                                                           tris.size());
	return 0;
}
