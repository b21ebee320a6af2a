// Acceptance program: what the reference's example programs check and print, with this backend
// compiled in next to the reference's own -- examples/backends/main.cpp (listing),
// examples/validation/main.cpp (known-answer triangle, bunny mask, build/traverse time gates,
// results.csv) and examples/bunny/main.cpp (rays/s print) -- without their network-fetched
// dependencies (tinyobjloader, stb): the bunny triangles and the expected 1024x1024 mask come as
// raw binary files (tests/test_dropin.py writes them from tests/golden/).
//
//   acceptance [--bunny TRIS.bin MASK.bin] [--csv results.csv] [--list]
//
// Every backend the build knows takes part: CPU always, CUDA (this repo), and the reference's
// Embree CPU backend when the build found Embree 4 (tests/dropin/Makefile: USE_EMBREE_CPU) -- the
// second oracle `north_star` names; it is reported as not compiled in otherwise.
// results.csv has the reference's schema, header and (argument-order) quirks:
//   Backend,Device,Test,Subtest,Value,Expected,Validation     (validation/main.cpp:249-261)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <type_traits>
#include <portableRT/portableRT.hpp>

using namespace portableRT;

namespace {
struct Row { // one line of results.csv; same value/expected formatting as the reference (std::to_string)
	std::string a, b, c, d, value, expected;
	bool ok;
};
std::vector<Row> rows;
constexpr float EPS = 0.0001f;

template <class T> void range_row(std::string a, std::string b, std::string c, std::string d, T v, T lo, T hi) {
	rows.push_back({a, b, c, d, std::to_string(v), std::to_string(lo) + " - " + std::to_string(hi), v >= lo && v <= hi});
}
template <class T> void row(std::string a, std::string b, std::string c, std::string d, T v, T want) {
	if constexpr (std::is_floating_point_v<T>)
		range_row<T>(a, b, c, d, v, want - EPS, want + EPS);
	else
		range_row<T>(a, b, c, d, v, want, want);
}

template <class T> std::vector<T> read_file(const char *path) {
	std::vector<T> v;
	std::ifstream f(path, std::ios::binary | std::ios::ate);
	if (!f) {
		std::fprintf(stderr, "cannot open %s\n", path);
		std::exit(2);
	}
	v.resize((size_t)f.tellg() / sizeof(T));
	f.seekg(0);
	f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
	return v;
}

// validation/main.cpp:66-111 (the rows pass test names where the constructor expects the backend:
// kept, so that the file is column-for-column what the reference program writes)
void tri_validation(Backend *b) {
	const std::array<float, 9> tri = {-1, -1, 0, 1, -1, 0, 0, 1, 0};
	Ray hit{{0.1f, 0, -1}, {0, 0, 1}}, miss{{-2, 0, -1}, {0, 0, 1}};
	selected_backend->set_tris({tri});
	auto h1 = nearest_hits({hit});
	auto h2 = nearest_hits({miss});
	const std::string n = b->name(), dev = b->device_name();
	row<bool>("FullReg Hit", "valid", n, dev, h1[0].valid, h1[0].valid);
	row<float>("FullReg Hit", "t", n, dev, h1[0].t, 1.0f);
	row<float>("FullReg Hit", "u", n, dev, h1[0].u, 0.3f);
	row<float>("FullReg Hit", "v", n, dev, h1[0].v, 0.5f);
	row<uint32_t>("FullReg Hit", "primitive_id", n, dev, h1[0].primitive_id, 0u);
	row<float>("FullReg Hit", "px", n, dev, h1[0].px, 0.1f);
	row<float>("FullReg Hit", "py", n, dev, h1[0].py, 0.0f);
	row<float>("FullReg Hit", "pz", n, dev, h1[0].pz, 0.0f);
	row<bool>("FullReg Miss", "valid", n, dev, h2[0].valid, false);
	auto f1 = nearest_hits<filter::valid>({hit});
	auto f2 = nearest_hits<filter::valid>({miss});
	row<bool>("Filtered Hit", "valid", n, dev, f1[0].valid, true);
	row<bool>("Filtered Miss", "valid", n, dev, f2[0].valid, false);
}

// the camera of validation/main.cpp:172-198 (directions divided by |sensor_pos|, not normalised)
std::vector<Ray> bunny_rays(int width, int height) {
	std::vector<Ray> rays;
	rays.reserve((size_t)width * height);
	const float camera_dist = 0.5f, sensor_size = 0.05f, sensor_dist = 0.05f;
	for (int y = height - 1; y >= 0; --y)
		for (int x = 0; x < width; ++x) {
			const float sx = sensor_size * (static_cast<float>(x) / width - 0.5);
			const float sy = sensor_size * (static_cast<float>(y) / height - 0.5);
			const std::array<float, 3> cam{0, 0, -camera_dist}, sp{sx, sy, -camera_dist + sensor_dist};
			Ray r;
			r.origin = cam;
			r.direction = {sp[0] - cam[0], sp[1] - cam[1], sp[2] - cam[2]};
			const float len = std::sqrt(sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2]);
			r.direction[0] /= len;
			r.direction[1] /= len;
			r.direction[2] /= len;
			rays.push_back(r);
		}
	return rays;
}

// validation/main.cpp:200-233 + the prints of bunny/main.cpp:95-134
void bunny_validation(Backend *b, const Tris &tris, const std::vector<Ray> &rays,
                      const std::vector<unsigned char> &mask) {
	using clk = std::chrono::high_resolution_clock;
	const auto b0 = clk::now();
	b->set_tris(tris);
	const auto b1 = clk::now();
	const auto t0 = clk::now();
	auto hits = nearest_hits(rays);
	const auto t1 = clk::now();
	const auto build_us = std::chrono::duration_cast<std::chrono::microseconds>(b1 - b0).count();
	const auto trace_us = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
	bool pixels = hits.size() == mask.size();
	size_t differ = 0;
	for (size_t i = 0; i < hits.size() && i < mask.size(); ++i)
		differ += ((hits[i].valid ? 255 : 0) != mask[i]);
	pixels = pixels && differ == 0;
	std::cout << "BVH building time: " << build_us / 1000 << " ms" << std::endl;
	std::cout << "Execution time: " << trace_us / 1000 << " ms" << std::endl;
	std::cout << static_cast<long long>(rays.size() / (std::max<long long>(trace_us, 1) / 1e6)) << "rays/s" << std::endl;
	if (differ)
		std::cout << differ << " of " << hits.size() << " mask pixels differ" << std::endl;
	const std::string n = b->name(), dev = b->device_name();
	row<bool>("Bunny", "Pixel Validation", n, dev, pixels, true);
	range_row<float>("Bunny", "BVH Build time", n, dev, build_us / 1000.0f, 0.0f, 1000.0f);
	range_row<float>("Bunny", "Traverse time", n, dev, trace_us / 1000.0f, 0.0f, 1000.0f);
}
} // namespace

int main(int argc, char **argv) {
	const char *tris_path = nullptr, *mask_path = nullptr, *csv = "results.csv";
	bool list_only = false;
	for (int i = 1; i < argc; ++i) {
		if (!std::strcmp(argv[i], "--bunny") && i + 2 < argc) {
			tris_path = argv[++i];
			mask_path = argv[++i];
		} else if (!std::strcmp(argv[i], "--csv") && i + 1 < argc) {
			csv = argv[++i];
		} else if (!std::strcmp(argv[i], "--list")) {
			list_only = true;
		}
	}
	// ---- examples/backends/main.cpp
	std::cout << "Printing all compiled backends: " << std::endl;
	for (auto b : all_backends())
		std::cout << "\t" << b->name() << std::endl;
	std::cout << std::endl << "Printing all available backends: " << std::endl;
	for (auto b : available_backends()) {
		select_backend(b); // it is necessary to initialise the backend to know the device
		std::cout << "\t" << b->name() << " (" << b->device_name() << ")" << std::endl;
	}
	std::cout << std::endl;
#ifdef USE_EMBREE_CPU
	std::cout << "Embree CPU backend: compiled in (second oracle)" << std::endl;
#else
	std::cout << "Embree CPU backend: not compiled in (Embree 4 not found when this program was built)" << std::endl;
#endif
	if (list_only)
		return 0;
	Tris tris;
	std::vector<Ray> rays;
	std::vector<unsigned char> mask;
	if (tris_path) {
		tris = read_file<Tri>(tris_path);
		mask = read_file<unsigned char>(mask_path);
		rays = bunny_rays(1024, 1024);
	}
	// ---- examples/validation/main.cpp:236-265
	for (auto b : available_backends()) {
		std::cout << "Testing " << b->name() << std::endl;
		select_backend(b);
		tri_validation(selected_backend);
		if (tris_path)
			bunny_validation(selected_backend, tris, rays, mask);
	}
	std::ofstream file(csv);
	file << "Backend,Device,Test,Subtest,Value,Expected,Validation\n";
	int failed = 0;
	for (const Row &r : rows) {
		if (!r.ok) {
			++failed;
			std::cout << "Validation failed in test: " << r.a << "," << r.b << "," << r.c << "," << r.d << ","
			          << r.value << "," << r.expected << "," << r.ok << "\n";
		}
		file << r.a << "," << r.b << "," << r.c << "," << r.d << "," << r.value << "," << r.expected << "," << r.ok
		     << "\n";
	}
	std::cout << rows.size() << " rows written to " << csv << ", " << failed << " failed" << std::endl;
	return failed ? 1 : 0;
}
