// api.cu -- the extern "C" layer (include/prt_b200.h).  Host glue only: context life cycle,
// grow-only device scratch, pinned double-buffered staging for the host-pointer entry points, CUDA
// event timing.  No computation happens on the host and there is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <dlfcn.h>
#include <sys/mman.h>
#include <mutex>
#include <thread>
#include <vector>

#include "prt_ctx.h"
#include "prt_hostpool.h"

using prt::fail;

namespace {
// Hand-offs between the thread that enqueues the pipeline of the host entry point and the two
// threads that stage pageable caller memory through pinned buffers (see prt_b200_nearest_hits).
struct Progress {
	std::mutex m;
	std::condition_variable cv;
	uint64_t produced = 0; // chunks whose rays are in their staging buffer
	uint64_t issued = 0;   // chunks whose copies and kernel have been enqueued (events recorded)
	uint64_t consumed = 0; // chunks whose records have reached the caller's array
	bool stop = false;     // error or early return: everybody leaves
	cudaError_t err = cudaSuccess;
	template <class F> bool wait(F ok) {
		std::unique_lock<std::mutex> l(m);
		cv.wait(l, [&] { return stop || ok(); });
		return !stop;
	}
	template <class F> void post(F f) {
		{
			std::lock_guard<std::mutex> l(m);
			f();
		}
		cv.notify_all();
	}
	void fail(cudaError_t e) {
		post([&] {
			if (err == cudaSuccess)
				err = e;
			stop = true;
		});
	}
};
} // namespace

static thread_local std::string g_create_err;

namespace prt {
// libnccl resolved at run time (dlopen): the library is only needed by multi-GPU contexts, and a
// process that already carries an NCCL (e.g. one that imported torch) shares that copy.  Only the
// five entry points of the triangle broadcast are used; ncclComm_t is an opaque pointer,
// ncclResult_t / ncclDataType_t are ints (ncclChar == 0).
struct NcclApi {
	void *handle = nullptr;
	int (*CommInitAll)(void **, int, const int *) = nullptr;
	int (*CommDestroy)(void *) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	bool load() {
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
			handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (handle)
				break;
		}
		if (!handle)
			return false;
		CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(handle, "ncclCommInitAll"));
		CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
		GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(handle, "ncclGroupStart"));
		GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(handle, "ncclGroupEnd"));
		Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(handle, "ncclBroadcast"));
		GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
		return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast;
	}
};
} // namespace prt

// every device of a (possibly multi-GPU) context: [0] = the context itself
static std::vector<prt_b200 *> all_devices(prt_b200 *c) {
	std::vector<prt_b200 *> v{c};
	v.insert(v.end(), c->peers.begin(), c->peers.end());
	return v;
}

// fn(sub-context, index) on every device concurrently (one host thread per further device, the
// caller's thread drives device 0); returns the first error, whose message is copied to the owner
template <class F> static int on_all_devices(prt_b200 *c, F fn) {
	if (c->peers.empty())
		return fn(c, 0);
	const auto devs = all_devices(c);
	std::vector<int> rc(devs.size(), PRT_OK);
	std::vector<std::thread> th;
	for (size_t i = 1; i < devs.size(); ++i)
		th.emplace_back([&, i] {
			cudaSetDevice(devs[i]->device);
			rc[i] = fn(devs[i], (int)i);
		});
	cudaSetDevice(c->device);
	rc[0] = fn(c, 0);
	for (auto &t : th)
		t.join();
	cudaSetDevice(c->device);
	for (size_t i = 0; i < devs.size(); ++i)
		if (rc[i] != PRT_OK) {
			if (i)
				c->err = "device " + std::to_string(devs[i]->device) + ": " + devs[i]->err;
			return rc[i];
		}
	return PRT_OK;
}

extern "C" {

int prt_b200_abi_version(void) { return PRT_B200_ABI_VERSION; }

int prt_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError(); // clear the sticky "no device" error
		return 0;
	}
	int ok = 0;
	for (int d = 0; d < n; ++d) {
		int major = 0;
		if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess &&
		    major == 10)
			++ok;
	}
	return ok;
}

const char *prt_b200_last_error(const prt_b200 *c) {
	return c ? c->err.c_str() : g_create_err.c_str();
}

// One single-device context (the whole context of a 1-GPU backend, a sub-context otherwise).
static int create_one(prt_b200 **out, int device) {
	prt_b200 *c = new prt_b200();
	c->device = device;
	cudaError_t e = cudaSetDevice(device);
	cudaDeviceProp prop{};
	if (e == cudaSuccess)
		e = cudaGetDeviceProperties(&prop, device);
	if (e == cudaSuccess)
		e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	for (int k = 0; k < 3 && e == cudaSuccess; ++k)
		e = cudaStreamCreateWithFlags(&c->pipe_stream[k], cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaEventCreate(&c->ev0);
	if (e == cudaSuccess)
		e = cudaEventCreate(&c->ev1);
	if (e == cudaSuccess)
		e = cudaEventCreate(&c->ev_k0);
	if (e == cudaSuccess)
		e = cudaEventCreate(&c->ev_k1);
	for (int k = 0; k < prt_b200::PIPE && e == cudaSuccess; ++k)
		for (int j = 0; j < 3 && e == cudaSuccess; ++j)
			e = cudaEventCreateWithFlags(&c->ev_pipe[k][j], cudaEventDisableTiming);
	if (e == cudaSuccess)
		e = cudaHostAlloc(reinterpret_cast<void **>(&c->probe_host), 64, cudaHostAllocMapped);
	if (e == cudaSuccess)
		e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->probe_dev), c->probe_host, 0);
	if (e == cudaSuccess)
		e = c->probe_ticket.reserve(64);
	if (e == cudaSuccess)
		e = c->counter.reserve(256);
	if (e == cudaSuccess)
		e = cudaMemset(c->counter.p, 0, 256); // the traversal kernel re-arms its counters itself
	if (e != cudaSuccess) {
		g_create_err = std::string("create: ") + cudaGetErrorString(e);
		prt_b200_destroy(c);
		return PRT_E_CUDA;
	}
	c->sm_count = prop.multiProcessorCount;
	c->l2_bytes = (uint64_t)prop.l2CacheSize;
	if (const char *e = std::getenv("PRT_B200_FAST_BOXES"))
		c->fast_boxes = std::atoi(e) != 0;
	if (const char *e = std::getenv("PRT_B200_TREELET_MODE"))
		c->optimise_mode = std::max(0, std::min(3, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_MAX_TREE_DEPTH")) // tests: force the strict fallback
		c->max_tree_depth = std::max(1, std::min(96, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_TREELET_PASSES"))
		c->optimise_passes = std::max(1, std::min(8, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_WATERTIGHT"))
		c->watertight = std::atoi(e) != 0;
	if (const char *e = std::getenv("PRT_B200_WIDE"))
		c->wide_mode = std::max(0, std::min(2, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_SORT_RAYS"))
		c->sort_rays = std::max(0, std::min(2, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_RAYKEY")) {
		int ob = 0, db = 0;
		if (std::sscanf(e, "%d,%d", &ob, &db) == 2 && ob >= 1 && ob <= 10 && db >= 0 && db <= 10 &&
		    3 * (ob + db) <= 32) {
			c->ray_key_ob = ob;
			c->ray_key_db = db;
		}
	}
	if (const char *e = std::getenv("PRT_B200_PIPE_TRACE"))
		c->pipe_trace = std::atoi(e) != 0;
	if (const char *e = std::getenv("PRT_B200_CHUNK_LOG2"))
		c->chunk_log2 = std::max(10, std::min(24, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_REFILL"))
		c->refill = c->refill_wide = c->refill_scattered = std::max(0, std::min(32, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_REFILL_WIDE"))
		c->refill_wide = std::max(0, std::min(32, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_LEAF_VOTES"))
		c->leaf_votes = c->leaf_votes_wide = std::max(1, std::min(32, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_LEAF_VOTES_WIDE"))
		c->leaf_votes_wide = std::max(1, std::min(32, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_REFILL_SCATTERED"))
		c->refill_scattered = std::max(0, std::min(32, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_COOP"))
		c->coop_after = std::max(0, std::min(1 << 20, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_GRAPHS"))
		c->use_graphs = std::atoi(e) != 0;
	if (const char *e = std::getenv("PRT_B200_COOP_SP"))
		c->coop_min_sp = std::max(0, std::min(64, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_COOP_BLOCKS"))
		c->coop_blocks = std::max(1, std::min(8, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_COPY_THREADS"))
		c->copy_threads = std::max(1, std::min(32, std::atoi(e)));
	if (const char *e = std::getenv("PRT_B200_BCAST"))
		c->bcast_mode = std::strcmp(e, "p2p") == 0 ? 1 : 0;
	if (const char *e = std::getenv("PRT_B200_PACKED_D2H"))
		c->packed_d2h = std::atoi(e) != 0;
	c->name = prop.name;
	*out = c;
	return PRT_OK;
}

// the CC 10.x devices in enumeration order
static std::vector<int> cc10_devices() {
	std::vector<int> v;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return v;
	}
	for (int d = 0; d < n; ++d) {
		int major = 0;
		if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess &&
		    major == 10)
			v.push_back(d);
	}
	return v;
}

// Single-process NCCL communicators over the context's devices (ncclCommInitAll), created when
// the context is: that takes about a second for 8 GPUs and must not land in the first set_tris.
// Where libnccl cannot be loaded or initialised the broadcast uses peer copies.
static void init_nccl(prt_b200 *c) {
	const auto devs = all_devices(c);
	if (c->bcast_mode != 0 || c->nccl || devs.size() < 2)
		return;
	auto *api = new prt::NcclApi();
	bool ok = api->load();
	if (ok) {
		std::vector<int> ids;
		for (auto *d : devs)
			ids.push_back(d->device);
		c->nccl_comms.assign(devs.size(), nullptr);
		ok = api->CommInitAll(c->nccl_comms.data(), (int)devs.size(), ids.data()) == 0;
		if (!ok)
			c->nccl_comms.clear();
		cudaSetDevice(c->device);
	}
	if (ok)
		c->nccl = api;
	else {
		delete api;
		c->bcast_mode = 1;
	}
}

int prt_b200_create_multi(prt_b200 **out, int n_gpus) {
	if (!out) {
		g_create_err = "create: out == NULL";
		return PRT_E_ARG;
	}
	*out = nullptr;
	const std::vector<int> devs = cc10_devices();
	if (devs.empty()) {
		g_create_err = "no compute-capability-10.x (sm_100a) device visible: kernels are built for "
		               "B200 only and this backend has no CPU fallback";
		return PRT_E_NO_DEVICE;
	}
	if (n_gpus <= 0) {
		const char *e = std::getenv("PRT_B200_GPUS");
		n_gpus = e ? std::atoi(e) : 1;
		if (n_gpus <= 0)
			n_gpus = 1;
	}
	if ((size_t)n_gpus > devs.size()) {
		g_create_err = "create: " + std::to_string(n_gpus) + " GPUs requested, " +
		               std::to_string(devs.size()) + " compute-capability-10.x devices visible";
		return PRT_E_NO_DEVICE;
	}
	// the first device may be chosen with PRT_B200_DEVICE (the others follow in enumeration order)
	size_t first = 0;
	if (const char *e = std::getenv("PRT_B200_DEVICE")) {
		const int want = std::atoi(e);
		for (size_t k = 0; k < devs.size(); ++k)
			if (devs[k] == want)
				first = k;
	}
	prt_b200 *c = nullptr;
	if (int rc = create_one(&c, devs[first]))
		return rc;
	for (int k = 1; k < n_gpus; ++k) {
		prt_b200 *p = nullptr;
		if (int rc = create_one(&p, devs[(first + k) % devs.size()])) {
			prt_b200_destroy(c);
			return rc;
		}
		c->peers.push_back(p);
	}
	if (n_gpus > 1) {
		// peer access lets the triangle broadcast (NCCL or cudaMemcpyPeerAsync) take NVLink
		const auto all = all_devices(c);
		for (auto *a : all) {
			cudaSetDevice(a->device);
			for (auto *b : all)
				if (a != b) {
					int can = 0;
					cudaDeviceCanAccessPeer(&can, a->device, b->device);
					if (can && cudaDeviceEnablePeerAccess(b->device, 0) != cudaSuccess)
						cudaGetLastError(); // already enabled
				}
		}
		c->name += " x" + std::to_string(n_gpus);
		init_nccl(c);
	}
	cudaSetDevice(c->device);
	*out = c;
	return PRT_OK;
}

int prt_b200_create(prt_b200 **out, int device) {
	if (!out) {
		g_create_err = "create: out == NULL";
		return PRT_E_ARG;
	}
	*out = nullptr;
	if (device < 0) // env PRT_B200_DEVICE / PRT_B200_GPUS: a drop-in user needs no source change
		return prt_b200_create_multi(out, 0);
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		cudaGetLastError();
		g_create_err = "no CUDA device visible (this backend has no CPU fallback)";
		return PRT_E_NO_DEVICE;
	}
	int major = 0;
	if (device >= n ||
	    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
	    major != 10) {
		g_create_err = "no compute-capability-10.x (sm_100a) device: kernels are built for B200 only";
		return PRT_E_NO_DEVICE;
	}
	return create_one(out, device);
}

int prt_b200_num_devices(const prt_b200 *c) { return c ? 1 + (int)c->peers.size() : 0; }

void prt_b200_destroy(prt_b200 *c) {
	if (!c)
		return;
	for (prt_b200 *p : c->peers)
		prt_b200_destroy(p);
	c->peers.clear();
	if (c->nccl) {
		for (void *comm : c->nccl_comms)
			if (comm)
				c->nccl->CommDestroy(comm);
		delete c->nccl; // (the library handle stays loaded: other users in the process may share it)
		c->nccl = nullptr;
	}
	if (c->device >= 0)
		cudaSetDevice(c->device);
	for (int k = 0; k < 3; ++k)
		if (c->pipe_stream[k])
			cudaStreamSynchronize(c->pipe_stream[k]);
	if (c->stream)
		cudaStreamSynchronize(c->stream);
	if (c->build_graph)
		cudaGraphExecDestroy(c->build_graph);
	delete c->pool_in;
	delete c->pool_out;
	if (c->probe_host)
		cudaFreeHost(c->probe_host);
	for (int k = 0; k < prt_b200::PIPE; ++k)
		for (int j = 0; j < 3; ++j)
			if (c->ev_pipe[k][j])
				cudaEventDestroy(c->ev_pipe[k][j]);
	for (int k = 0; k < 3; ++k)
		if (c->pipe_stream[k])
			cudaStreamDestroy(c->pipe_stream[k]);
	if (c->ev0)
		cudaEventDestroy(c->ev0);
	if (c->ev1)
		cudaEventDestroy(c->ev1);
	if (c->ev_k0)
		cudaEventDestroy(c->ev_k0);
	if (c->ev_k1)
		cudaEventDestroy(c->ev_k1);
	if (c->stream)
		cudaStreamDestroy(c->stream);
	delete c; // DevBuf / PinnedBuf members release their memory (prt_ctx.h)
}

int prt_b200_device_name(const prt_b200 *c, char *buf, size_t cap) {
	if (!c || !buf || !cap)
		return 0;
	std::strncpy(buf, c->name.c_str(), cap - 1);
	buf[cap - 1] = 0;
	return (int)std::min(c->name.size(), cap - 1);
}

int prt_b200_set_wide_nodes(prt_b200 *c, int mode) {
	if (!c || mode < 0 || mode > 2)
		return fail(c, PRT_E_ARG, "set_wide_nodes: mode must be 0, 1 or 2");
	for (prt_b200 *d : all_devices(c))
		d->wide_mode = mode;
	return PRT_OK;
}

int prt_b200_set_tree_optimisation(prt_b200 *c, int mode, int passes) {
	if (!c || mode < 0 || mode > 3 || passes < 1 || passes > 8)
		return fail(c, PRT_E_ARG, "set_tree_optimisation: mode must be 0..3 and passes 1..8");
	for (prt_b200 *d : all_devices(c)) {
		d->optimise_mode = mode;
		d->optimise_passes = passes;
	}
	return PRT_OK;
}
int32_t prt_b200_tree_depth(const prt_b200 *c) { return c ? c->tree_depth : 0; }
float prt_b200_last_optimise_ms(const prt_b200 *c) { return c ? c->last_optimise_ms : 0.f; }
uint64_t prt_b200_strict_fallbacks(const prt_b200 *c) { return c ? c->strict_fallbacks : 0; }
uint64_t prt_b200_refits(const prt_b200 *c) { return c ? c->refits : 0; }
uint64_t prt_b200_refit_rejects(const prt_b200 *c) { return c ? c->refit_rejects : 0; }

int prt_b200_set_triangle_test(prt_b200 *c, int mode) {
	if (!c || mode < 0 || mode > 1)
		return fail(c, PRT_E_ARG, "set_triangle_test: mode must be 0 or 1");
	for (prt_b200 *d : all_devices(c))
		d->watertight = mode;
	return PRT_OK;
}
int prt_b200_triangle_test(const prt_b200 *c) { return c && c->recs_vertex_form ? 1 : 0; }

int prt_b200_set_ray_sorting(prt_b200 *c, int mode) {
	if (!c || mode < 0 || mode > 2)
		return fail(c, PRT_E_ARG, "set_ray_sorting: mode must be 0, 1 or 2");
	for (prt_b200 *d : all_devices(c))
		d->sort_rays = mode;
	return PRT_OK;
}
uint64_t prt_b200_sorted_batches(const prt_b200 *c) { return c ? c->sorted_batches : 0; }

int prt_b200_set_trace_opts(prt_b200 *c, const prt_trace_opts *o) {
	if (!c)
		return PRT_E_ARG;
	for (prt_b200 *d : all_devices(c))
		d->opts = o ? *o : prt_trace_opts{1, 1e-4f, 64.0f};
	return PRT_OK;
}

// A failed build must not leave a half-committed scene behind (buffers freed or undersized but
// n_tris > 0): fall back to the empty scene, on which every ray misses.
static void reset_scene(prt_b200 *c) {
	c->n_tris = 0;
	c->n_nodes = 0;
	c->root = 0;
	c->wide_built = false;
	c->tree_optimised = false;
	c->topology_valid = false;
	c->tree_depth = 0;
	for (int a = 0; a < 3; ++a)
		c->scene_lo[a] = c->scene_hi[a] = c->scene_absmax[a] = 0.f;
}

static int timed_build_raw(prt_b200 *c, const float *d_tris9, uint64_t n) {
	PRT_CUDA(c, cudaSetDevice(c->device));
	PRT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
	bool reused = false;
	int rc = prt::try_reuse_topology(c, d_tris9, n, c->stream, &reused);
	if (rc)
		return rc;
	if (!reused) {
		rc = prt::build_lbvh(c, d_tris9, n);
		if (rc)
			return rc;
	}
	PRT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
	// root index + scene box (fast box test margin, ray-sort grid), published by the build
	struct {
		int32_t root;
		float lo[3], hi[3];
		int32_t depth;
	} ri{};
	if (c->n_nodes)
		PRT_CUDA(c, cudaMemcpyAsync(&ri, c->root_info.p, sizeof ri, cudaMemcpyDeviceToHost, c->stream));
	PRT_CUDA(c, cudaStreamSynchronize(c->stream));
	PRT_CUDA(c, cudaEventElapsedTime(&c->last_build_ms, c->ev0, c->ev1));
	c->root = c->n_nodes ? ri.root : 0;
	for (int a = 0; a < 3; ++a) {
		c->scene_lo[a] = c->n_nodes ? ri.lo[a] : 0.f;
		c->scene_hi[a] = c->n_nodes ? ri.hi[a] : 0.f;
		c->scene_absmax[a] = std::max(std::fabs(c->scene_lo[a]), std::fabs(c->scene_hi[a]));
	}
	return PRT_OK;
}

static int timed_build(prt_b200 *c, const float *d_tris9, uint64_t n, float *ms) {
	const int rc = timed_build_raw(c, d_tris9, n);
	if (rc) {
		const std::string keep = c->err;
		cudaStreamSynchronize(c->stream);
		if (cudaGetLastError() != cudaSuccess) { /* non-sticky errors are cleared; sticky ones stay */
		}
		reset_scene(c);
		c->err = keep;
		return rc;
	}
	if (ms)
		*ms = c->last_build_ms;
	return PRT_OK;
}

// Multi-GPU: the triangles that device 0 holds at `src` reach every further device's tris_raw over
// NVLink -- one grouped ncclBroadcast on per-device communicators (ncclCommInitAll, single
// process), or cudaMemcpyPeerAsync where libnccl cannot be loaded (PRT_B200_BCAST=p2p forces it).
static int broadcast_tris(prt_b200 *c, const float *src, uint64_t n) {
	const auto devs = all_devices(c);
	const size_t bytes = (size_t)n * 36;
	for (size_t i = 1; i < devs.size(); ++i) {
		PRT_CUDA(c, cudaSetDevice(devs[i]->device));
		PRT_CUDA(c, devs[i]->tris_raw.reserve(bytes));
	}
	PRT_CUDA(c, cudaSetDevice(c->device));
	if (bytes == 0)
		return PRT_OK;
	if (c->nccl) {
		int rc = c->nccl->GroupStart();
		for (size_t i = 0; i < devs.size() && rc == 0; ++i)
			rc = c->nccl->Broadcast(src, i == 0 ? (void *)src : devs[i]->tris_raw.p, bytes, 0 /*ncclChar*/,
			                        0, c->nccl_comms[i], devs[i]->stream);
		const int rc2 = c->nccl->GroupEnd();
		if (rc || rc2) {
			const int bad = rc ? rc : rc2;
			return fail(c, PRT_E_CUDA, c->nccl->GetErrorString ? c->nccl->GetErrorString(bad)
			                                                   : "ncclBroadcast failed");
		}
		c->bcast_used = "nccl";
	} else {
		// the source must be complete before the peers' streams read it
		PRT_CUDA(c, cudaStreamSynchronize(c->stream));
		for (size_t i = 1; i < devs.size(); ++i)
			PRT_CUDA(c, cudaMemcpyPeerAsync(devs[i]->tris_raw.p, devs[i]->device, src, c->device, bytes,
			                                devs[i]->stream));
		c->bcast_used = "p2p";
	}
	for (auto *d : devs) {
		PRT_CUDA(c, cudaSetDevice(d->device));
		PRT_CUDA(c, cudaStreamSynchronize(d->stream));
	}
	PRT_CUDA(c, cudaSetDevice(c->device));
	return PRT_OK;
}

const char *prt_b200_broadcast_path(const prt_b200 *c) { return c ? c->bcast_used.c_str() : ""; }

// build on every device from its own copy of the triangles (deterministic: identical BVHs)
static int build_everywhere(prt_b200 *c, const float *d_tris9_dev0, uint64_t n, float *ms) {
	if (c->peers.empty())
		return timed_build(c, d_tris9_dev0, n, ms);
	if (int rc = broadcast_tris(c, d_tris9_dev0, n))
		return rc;
	const int rc = on_all_devices(c, [&](prt_b200 *d, int i) {
		return timed_build(d, i == 0 ? d_tris9_dev0 : d->tris_raw.as<float>(), n, nullptr);
	});
	if (rc) { // all devices or none
		for (prt_b200 *d : all_devices(c))
			reset_scene(d);
		return rc;
	}
	for (prt_b200 *d : c->peers)
		c->last_build_ms = std::max(c->last_build_ms, d->last_build_ms);
	if (ms)
		*ms = c->last_build_ms;
	return PRT_OK;
}

int prt_b200_set_tris_dev(prt_b200 *c, const float *d_tris9, uint64_t n, float *build_ms) {
	if (!c || (n && !d_tris9))
		return fail(c, PRT_E_ARG, "set_tris_dev: NULL argument");
	return build_everywhere(c, d_tris9, n, build_ms);
}

int prt_b200_set_tris(prt_b200 *c, const float *tris9, uint64_t n) {
	if (!c || (n && !tris9))
		return fail(c, PRT_E_ARG, "set_tris: NULL argument");
	PRT_CUDA(c, cudaSetDevice(c->device));
	if (n) {
		PRT_CUDA(c, c->tris_raw.reserve(n * 36));
		PRT_CUDA(c, cudaMemcpyAsync(c->tris_raw.p, tris9, n * 36, cudaMemcpyHostToDevice, c->stream));
	}
	return build_everywhere(c, c->tris_raw.as<float>(), n, nullptr);
}

static int check_layout(prt_b200 *c, uint32_t mask, const prt_hit_layout *l) {
	if (!l || l->stride == 0)
		return fail(c, PRT_E_ARG, "nearest_hits: NULL/empty hit layout");
	struct {
		uint32_t bit;
		int32_t off;
		int32_t size;
	} f[] = {{PRT_TAG_UV, l->off_u, 4},    {PRT_TAG_UV, l->off_v, 4},  {PRT_TAG_T, l->off_t, 4},
	         {PRT_TAG_PID, l->off_pid, 4}, {PRT_TAG_VALID, l->off_valid, 1},
	         {PRT_TAG_P, l->off_px, 4},    {PRT_TAG_P, l->off_py, 4},  {PRT_TAG_P, l->off_pz, 4}};
	for (auto &x : f) {
		if (!(mask & x.bit))
			continue;
		if (x.off < 0 || (uint32_t)(x.off + x.size) > l->stride || (x.size == 4 && (x.off & 3)))
			return fail(c, PRT_E_ARG, "nearest_hits: hit layout lacks a field the tag mask requests");
	}
	if (l->stride & 3) {
		// 4-byte stores need 4-byte aligned records unless only `valid` is written
		if (mask != PRT_TAG_VALID)
			return fail(c, PRT_E_ARG, "nearest_hits: record stride must be a multiple of 4");
	}
	return PRT_OK;
}

static int timed_trace(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t mask,
                       const prt::TraceOut &out, uint32_t *d_counts, float *ms) {
	PRT_CUDA(c, cudaSetDevice(c->device));
	PRT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
	int rc = prt::launch_trace(c, d_rays6, n, mask, out, d_counts, c->stream, -1, prt::EXOTIC_DEFERRED);
	if (rc)
		return rc;
	PRT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
	PRT_CUDA(c, cudaStreamSynchronize(c->stream));
	PRT_CUDA(c, cudaEventElapsedTime(&c->last_trace_ms, c->ev0, c->ev1));
	c->last_kernel_ms = 0.f;
	if (n)
		PRT_CUDA(c, cudaEventElapsedTime(&c->last_kernel_ms, c->ev_k0, c->ev_k1));
	// rays the fast kernel set aside (non-finite / overflowing components): exact kernel, timed too
	bool ran = false;
	PRT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
	if ((rc = prt::finish_exotic(c, c->stream, &ran)))
		return rc;
	if (ran) {
		float more = 0.f;
		PRT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
		PRT_CUDA(c, cudaStreamSynchronize(c->stream));
		PRT_CUDA(c, cudaEventElapsedTime(&more, c->ev0, c->ev1));
		c->last_trace_ms += more;
	}
	if (ms)
		*ms = c->last_trace_ms;
	return PRT_OK;
}

int prt_b200_trace_dev(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t mask,
                       const prt_soa_out *o, float *trace_ms) {
	if (!c || !o || (n && !d_rays6))
		return fail(c, PRT_E_ARG, "trace_dev: NULL argument");
	if (mask == 0 || mask > PRT_TAG_ALL)
		return fail(c, PRT_E_ARG, "trace_dev: tag mask must be in 1..31");
	if (n && (((mask & PRT_TAG_UV) && !o->uv) || ((mask & PRT_TAG_T) && !o->t) ||
	          ((mask & PRT_TAG_PID) && !o->pid) || ((mask & PRT_TAG_P) && !o->p) ||
	          ((mask & PRT_TAG_VALID) && !o->valid)))
		return fail(c, PRT_E_ARG, "trace_dev: NULL output for a requested tag");
	if (int rc = prt::maybe_optimise_tree(c, n))
		return rc;
	prt::TraceOut out;
	out.soa = *o;
	return timed_trace(c, d_rays6, n, mask, out, nullptr, trace_ms);
}

int prt_b200_trace_dev_aos(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t mask,
                           const prt_hit_layout *layout, void *d_hits, float *trace_ms) {
	if (!c || (n && (!d_rays6 || !d_hits)))
		return fail(c, PRT_E_ARG, "trace_dev_aos: NULL argument");
	if (mask == 0 || mask > PRT_TAG_ALL)
		return fail(c, PRT_E_ARG, "trace_dev_aos: tag mask must be in 1..31");
	if (int rc = check_layout(c, mask, layout))
		return rc;
	if (int rc = prt::maybe_optimise_tree(c, n))
		return rc;
	prt::TraceOut out;
	out.aos = d_hits;
	out.layout = *layout;
	return timed_trace(c, d_rays6, n, mask, out, nullptr, trace_ms);
}

int prt_b200_trace_count_dev(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t *d_counts) {
	if (!c || (n && (!d_rays6 || !d_counts)))
		return fail(c, PRT_E_ARG, "trace_count_dev: NULL argument");
	prt::TraceOut out;
	return timed_trace(c, d_rays6, n, PRT_TAG_ALL, out, d_counts, nullptr);
}

} // extern "C"

// ---- host-pointer entry point -------------------------------------------------------------------
// true iff the WHOLE range is page-locked host memory (a buffer of which only a part was
// registered must be staged like pageable memory, not DMA'd)
static bool is_pinned(const void *p, size_t bytes) {
	auto host = [](const void *q) {
		cudaPointerAttributes a{};
		if (cudaPointerGetAttributes(&a, q) != cudaSuccess) {
			cudaGetLastError();
			return false;
		}
		return a.type == cudaMemoryTypeHost;
	};
	if (!host(p))
		return false;
	return bytes == 0 || host(static_cast<const char *>(p) + bytes - 1);
}

static prt::HostPool *staging_pool(prt_b200 *c, prt::HostPool *&slot, int n_devices) {
	if (!slot) {
		int t = c->copy_threads;
		if (t <= 0) { // two pools per device (in, out): leave the box's cores to all of them
			const unsigned hw = std::thread::hardware_concurrency();
			t = (int)std::max(1u, std::min(8u, (hw ? hw : 4u) / (2u * (unsigned)std::max(1, n_devices))));
		}
		slot = new prt::HostPool(t - 1);
	}
	return slot;
}

// Tightly packed device-side record for results that are staged anyway (pageable caller memory):
// only the requested fields cross PCIe (HitReg<t,primitive_id> is 16 bytes for 8 useful ones,
// HitReg<valid> 8 for 1); the consumer threads scatter them into the caller's records while they
// copy.  Word fields first (u, v, t, pid, px, py, pz), the valid byte last.
struct PackedLayout {
	prt_hit_layout dev{};   // what the kernel writes
	int n_words = 0;        // 4-byte fields
	int32_t src[7], dst[7]; // their offsets in the packed / the caller's record
	int32_t valid_src = -1, valid_dst = -1;
};

static PackedLayout make_packed(uint32_t mask, const prt_hit_layout &user) {
	PackedLayout p;
	p.dev = prt_hit_layout{0, -1, -1, -1, -1, -1, -1, -1, -1};
	int32_t at = 0;
	auto word = [&](int32_t &dev_off, int32_t user_off) {
		dev_off = at;
		p.src[p.n_words] = at;
		p.dst[p.n_words] = user_off;
		++p.n_words;
		at += 4;
	};
	if (mask & PRT_TAG_UV) {
		word(p.dev.off_u, user.off_u);
		word(p.dev.off_v, user.off_v);
	}
	if (mask & PRT_TAG_T)
		word(p.dev.off_t, user.off_t);
	if (mask & PRT_TAG_PID)
		word(p.dev.off_pid, user.off_pid);
	if (mask & PRT_TAG_P) {
		word(p.dev.off_px, user.off_px);
		word(p.dev.off_py, user.off_py);
		word(p.dev.off_pz, user.off_pz);
	}
	if (mask & PRT_TAG_VALID) {
		p.dev.off_valid = at;
		p.valid_src = at;
		p.valid_dst = user.off_valid;
		at += 1;
	}
	p.dev.stride = p.n_words ? (uint32_t)((at + 3) & ~3) : (uint32_t)at;
	return p;
}

template <int NW, bool V>
static void scatter_records(const PackedLayout &pl, uint32_t user_stride, const char *src, char *dst,
                            uint64_t n) {
	const uint32_t ps = pl.dev.stride;
	for (uint64_t i = 0; i < n; ++i, src += ps, dst += user_stride) {
		for (int k = 0; k < NW; ++k) {
			uint32_t w;
			std::memcpy(&w, src + pl.src[k], 4);
			std::memcpy(dst + pl.dst[k], &w, 4);
		}
		if (V)
			dst[pl.valid_dst] = src[pl.valid_src];
	}
}

using ScatterFn = void (*)(const PackedLayout &, uint32_t, const char *, char *, uint64_t);
static ScatterFn scatter_fn(int nw, bool v) {
	static const ScatterFn table[8][2] = {
	    {scatter_records<0, false>, scatter_records<0, true>}, {scatter_records<1, false>, scatter_records<1, true>},
	    {scatter_records<2, false>, scatter_records<2, true>}, {scatter_records<3, false>, scatter_records<3, true>},
	    {scatter_records<4, false>, scatter_records<4, true>}, {scatter_records<5, false>, scatter_records<5, true>},
	    {scatter_records<6, false>, scatter_records<6, true>}, {scatter_records<7, false>, scatter_records<7, true>}};
	return table[nw][v ? 1 : 0];
}

// One device's share of a host batch.  Rays are cut into chunks that flow through three stages on
// three streams -- H2D, traversal kernel writing AoS records, D2H -- over a ring of buffer sets, so
// both copy engines and the SMs work at the same time.  Pinned / registered host memory is DMA'd
// directly; pageable memory (a std::vector, as in the reference API) is staged through pinned
// buffers by a producer and a consumer thread, each with its own pool of copy threads, and its
// results cross PCIe tightly packed.
static int nearest_hits_device(prt_b200 *c, const float *rays6, uint64_t n, uint32_t mask,
                               const prt_hit_layout *layout, void *hits_out, int coherence,
                               int n_devices) {
	if (n == 0)
		return PRT_OK;
	PRT_CUDA(c, cudaSetDevice(c->device));
	const bool in_pinned = is_pinned(rays6, n * 24);
	const bool out_pinned = is_pinned(hits_out, n * (size_t)layout->stride);
	// Chunk schedule: the D2H engine is the bottleneck (records are larger than rays) and starts
	// only when the first chunk has been uploaded and traced, so small batches are cut into ~8
	// chunks; every chunk costs launches, a shorter (less efficient) kernel and some idle time of
	// the copy engines between transfers, so not more.  Chunks are multiples of 2^16 rays (the
	// reordering threshold) and at most 2^19 (2^21 for batches of >= 2^24 rays, where the ramp-up
	// no longer matters and reordering works better on larger chunks).  Measured on C2 (2 073 600 rays,
	// pinned buffers, wall clock per call): 2^16 1.94 ms, 2^17 1.65, 2^18 1.60, n/6 1.69, 2^19 1.72; a small first
	// chunk followed by doubling ones was slower (the uploads fall behind the downloads), and so
	// were kernels that read rays / write records through PCIe themselves (2.07 / 9.7 ms).
	// PRT_B200_CHUNK_LOG2 forces a size.
	const uint64_t CH_MAX = n >= (1ull << 24) ? (1ull << 21) : (1ull << 19);
	uint64_t CH = ((n + 7) / 8 + 65535) / 65536 * 65536;
	if (!in_pinned || !out_pinned)
		CH = std::max<uint64_t>(CH, 1ull << 19); // bound by the host-side staging copies: few large ones
	if (c->chunk_log2 > 0)
		CH = 1ull << c->chunk_log2;
	CH = std::min<uint64_t>(std::min<uint64_t>(CH, c->chunk_log2 > 0 ? (1ull << 24) : CH_MAX), n);
	std::vector<uint64_t> start; // chunk k = rays [start[k], start[k+1])
	for (uint64_t at = 0; at < n; at += CH)
		start.push_back(at);
	start.push_back(n);
	const auto call0 = std::chrono::steady_clock::now();
	constexpr int PIPE = prt_b200::PIPE;
	const PackedLayout packed = make_packed(mask, *layout);
	const bool use_packed = !out_pinned && c->packed_d2h && packed.dev.stride < layout->stride;
	const prt_hit_layout dev_layout = use_packed ? packed.dev : *layout;
	const size_t ray_b = 24, hit_b = layout->stride, dev_hit_b = dev_layout.stride;
	const uint64_t n_chunks = start.size() - 1;
	for (int k = 0; k < PIPE && (uint64_t)k < n_chunks; ++k) {
		PRT_CUDA(c, c->rays_dev[k].reserve(CH * ray_b));
		PRT_CUDA(c, c->hits_dev[k].reserve(CH * dev_hit_b));
		if (!in_pinned)
			PRT_CUDA(c, c->rays_pin[k].reserve(CH * ray_b));
		if (!out_pinned)
			PRT_CUDA(c, c->hits_pin[k].reserve(CH * dev_hit_b));
	}
	prt::HostPool *pool_in = in_pinned ? nullptr : staging_pool(c, c->pool_in, n_devices);
	prt::HostPool *pool_out = out_pinned ? nullptr : staging_pool(c, c->pool_out, n_devices);
	enum { IN = 0, KERN = 1, OUT = 2 };
	cudaStream_t s_in = c->pipe_stream[IN], s_k = c->pipe_stream[KERN], s_out = c->pipe_stream[OUT];
	Progress pg;
	std::thread producer, consumer;
	struct Joiner { // every exit path: tell the helpers to leave, wait for them and for the device
		Progress &pg;
		std::thread &a, &b;
		prt_b200 *c;
		~Joiner() {
			pg.post([&] { pg.stop = true; });
			if (a.joinable())
				a.join();
			if (b.joinable())
				b.join();
			// no copy into the caller's memory may still be in flight once the call has returned
			for (int k = 0; k < 3; ++k)
				cudaStreamSynchronize(c->pipe_stream[k]);
		}
	} joiner{pg, producer, consumer, c};
	if (!in_pinned)
		producer = std::thread([&] {
			cudaSetDevice(c->device);
			for (uint64_t k = 0; k < n_chunks; ++k) {
				const int b = (int)(k % PIPE);
				if (k >= (uint64_t)PIPE) { // the H2D of chunk k-PIPE must have left the buffer
					if (!pg.wait([&] { return pg.issued > k - PIPE; }))
						return;
					const cudaError_t e = cudaEventSynchronize(c->ev_pipe[b][IN]);
					if (e != cudaSuccess)
						return pg.fail(e);
				}
				pool_in->copy(c->rays_pin[b].p, rays6 + start[k] * 6, (start[k + 1] - start[k]) * ray_b);
				pg.post([&] { pg.produced = k + 1; });
			}
		});
	if (!out_pinned)
		consumer = std::thread([&] {
			cudaSetDevice(c->device);
			const ScatterFn scatter = scatter_fn(packed.n_words, packed.valid_src >= 0);
			for (uint64_t k = 0; k < n_chunks; ++k) {
				const int b = (int)(k % PIPE);
				if (!pg.wait([&] { return pg.issued > k; }))
					return;
				const cudaError_t e = cudaEventSynchronize(c->ev_pipe[b][OUT]);
				if (e != cudaSuccess)
					return pg.fail(e);
				char *dst = static_cast<char *>(hits_out) + start[k] * hit_b;
				const uint64_t cnt = start[k + 1] - start[k];
				if (use_packed) {
					const char *src = static_cast<const char *>(c->hits_pin[b].p);
					pool_out->run([&](int part, int parts) {
						const uint64_t per = (cnt + parts - 1) / parts, lo = std::min(cnt, per * part),
						               hi = std::min(cnt, lo + per);
						scatter(packed, layout->stride, src + lo * dev_hit_b, dst + lo * hit_b, hi - lo);
					});
				} else {
					pool_out->copy(dst, c->hits_pin[b].p, cnt * hit_b);
				}
				pg.post([&] { pg.consumed = k + 1; });
			}
		});
	auto helper_failed = [&]() -> int {
		return fail(c, PRT_E_CUDA, "nearest_hits: staging thread", pg.err);
	};
	// Three stages on three streams over a ring of PIPE buffer sets: the H2D engine, the SMs and
	// the D2H engine each run back to back; a stage waits only for the event that frees what it
	// needs (the previous stage of the same chunk, the next stage of the chunk PIPE earlier).
	// The context stream marks the start (the BVH build ran there).
	PRT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
	for (int k = 0; k < 3; ++k)
		PRT_CUDA(c, cudaStreamWaitEvent(c->pipe_stream[k], c->ev0, 0));
	// PRT_B200_PIPE_TRACE=1: device timeline of every stage of every chunk, printed to stderr
	std::vector<cudaEvent_t> tl;
	std::vector<double> host_us;
	const auto wall0 = std::chrono::steady_clock::now();
	auto mark = [&](cudaStream_t st) {
		if (!c->pipe_trace)
			return;
		cudaEvent_t e;
		cudaEventCreate(&e);
		cudaEventRecord(e, st);
		tl.push_back(e);
	};
	for (uint64_t k = 0; k < n_chunks; ++k) {
		const int b = (int)(k % PIPE);
		const uint64_t first = start[k], cnt = start[k + 1] - first;
		// ---- H2D
		const void *src = rays6 + first * 6;
		if (!in_pinned) {
			if (!pg.wait([&] { return pg.produced > k; }))
				return helper_failed();
			src = c->rays_pin[b].p;
		}
		if (k >= (uint64_t)PIPE) // device rays b: the kernel of chunk k-PIPE must have left
			PRT_CUDA(c, cudaStreamWaitEvent(s_in, c->ev_pipe[b][KERN], 0));
		if (c->pipe_trace)
			host_us.push_back(std::chrono::duration<double, std::micro>(
			                      std::chrono::steady_clock::now() - wall0).count());
		mark(s_in);
		PRT_CUDA(c, cudaMemcpyAsync(c->rays_dev[b].p, src, cnt * ray_b, cudaMemcpyHostToDevice, s_in));
		mark(s_in);
		PRT_CUDA(c, cudaEventRecord(c->ev_pipe[b][IN], s_in));
		// ---- traversal (all chunks on one stream: one counter / reordering scratch slot)
		PRT_CUDA(c, cudaStreamWaitEvent(s_k, c->ev_pipe[b][IN], 0));
		if (k >= (uint64_t)PIPE) // device hits b: the D2H of chunk k-PIPE must have left
			PRT_CUDA(c, cudaStreamWaitEvent(s_k, c->ev_pipe[b][OUT], 0));
		prt::TraceOut out;
		out.aos = c->hits_dev[b].p;
		out.layout = dev_layout;
		out.slot = 0;
		mark(s_k);
		int rc = prt::launch_trace(c, c->rays_dev[b].as<float>(), cnt, mask, out, nullptr, s_k,
		                           coherence);
		if (rc)
			return rc;
		mark(s_k);
		PRT_CUDA(c, cudaEventRecord(c->ev_pipe[b][KERN], s_k));
		// ---- D2H
		if (!out_pinned && k >= (uint64_t)PIPE) { // staging buffer b may still hold chunk k-PIPE
			if (!pg.wait([&] { return pg.consumed > k - PIPE; }))
				return helper_failed();
		}
		PRT_CUDA(c, cudaStreamWaitEvent(s_out, c->ev_pipe[b][KERN], 0));
		void *dst = out_pinned ? static_cast<void *>(static_cast<char *>(hits_out) + first * hit_b)
		                       : c->hits_pin[b].p;
		mark(s_out);
		PRT_CUDA(c, cudaMemcpyAsync(dst, c->hits_dev[b].p, cnt * dev_hit_b, cudaMemcpyDeviceToHost, s_out));
		mark(s_out);
		PRT_CUDA(c, cudaEventRecord(c->ev_pipe[b][OUT], s_out));
		pg.post([&] { pg.issued = k + 1; });
	}
	if (consumer.joinable()) { // it leaves after the last chunk has reached the caller's array
		consumer.join();
		if (pg.err != cudaSuccess)
			return helper_failed();
	}
	// the last D2H depends on everything before it: join it into the context stream and time the
	// whole call on the device
	PRT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_pipe[(n_chunks - 1) % PIPE][OUT], 0));
	PRT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
	PRT_CUDA(c, cudaStreamSynchronize(c->stream));
	PRT_CUDA(c, cudaEventElapsedTime(&c->last_trace_ms, c->ev0, c->ev1));
	c->last_h2d_bytes = n * ray_b;
	c->last_d2h_bytes = n * dev_hit_b;
	if (c->pipe_trace) {
		std::fprintf(stderr, "[prt_b200] dev %d nearest_hits %llu rays, %llu chunks of %llu, device %.3f ms, "
		                     "call %.0f us, in %s, out %s%s "
		                     "(chunk: enqueue us | h2d | kernel | d2h, ms since start)\n",
		             c->device, (unsigned long long)n, (unsigned long long)n_chunks, (unsigned long long)CH,
		             c->last_trace_ms,
		             std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - call0)
		                 .count(),
		             in_pinned ? "pinned" : "pageable", out_pinned ? "pinned" : "pageable",
		             use_packed ? " (packed)" : "");
		for (uint64_t k = 0; k < n_chunks; ++k) {
			float t[6];
			for (int j = 0; j < 6; ++j)
				cudaEventElapsedTime(&t[j], c->ev0, tl[k * 6 + j]);
			std::fprintf(stderr, "  %3llu: %7.1f | %.3f-%.3f | %.3f-%.3f | %.3f-%.3f\n",
			             (unsigned long long)k, host_us[k], t[0], t[1], t[2], t[3], t[4], t[5]);
		}
		for (cudaEvent_t e : tl)
			cudaEventDestroy(e);
	}
	return PRT_OK;
}

extern "C" {

int prt_b200_nearest_hits(prt_b200 *c, const float *rays6, uint64_t n, uint32_t mask,
                          const prt_hit_layout *layout, void *hits_out) {
	if (!c || (n && (!rays6 || !hits_out)))
		return fail(c, PRT_E_ARG, "nearest_hits: NULL argument");
	if (mask == 0 || mask > PRT_TAG_ALL)
		return fail(c, PRT_E_ARG, "nearest_hits: tag mask must be in 1..31");
	if (int rc = check_layout(c, mask, layout))
		return rc;
	if (n == 0)
		return PRT_OK;
	PRT_CUDA(c, cudaSetDevice(c->device));
	// coherent or not is decided once, on the host copy of the batch (no device probe, no stream
	// synchronisation inside the pipeline)
	const int coherence = (c->sort_rays == 2 && n >= 65536) ? prt::host_ray_probe(c, rays6, n) : -1;
	// A large pageable result (a fresh std::vector) is about to be touched for the first time by
	// the staging threads: ask for transparent huge pages first (512x fewer page faults where the
	// kernel allows it; harmless otherwise)
	static const bool use_thp = [] {
		const char *e = std::getenv("PRT_B200_THP");
		return !e || std::atoi(e) != 0;
	}();
	if (use_thp && (size_t)n * layout->stride >= (size_t(16) << 20) && !is_pinned(hits_out, 1)) {
		const uintptr_t lo = (reinterpret_cast<uintptr_t>(hits_out) + (2u << 20) - 1) & ~uintptr_t((2u << 20) - 1);
		const uintptr_t hi = (reinterpret_cast<uintptr_t>(hits_out) + (size_t)n * layout->stride) &
		                     ~uintptr_t((2u << 20) - 1);
		if (hi > lo)
			madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);
	}
	const int G = 1 + (int)c->peers.size();
	// Multi-GPU: device i takes the contiguous slice [i*n/G, (i+1)*n/G) and writes its records at
	// the slice's offset in the caller's array -- the hits come back in ray order with no gather
	// step.  Every device counts the WHOLE batch towards its lazy tree optimisation, so that all
	// replicas are optimised before the same call.
	const int rc = on_all_devices(c, [&](prt_b200 *d, int i) -> int {
		if (int r = prt::maybe_optimise_tree(d, n))
			return r;
		const uint64_t lo = n * (uint64_t)i / G, hi = n * (uint64_t)(i + 1) / G;
		return nearest_hits_device(d, rays6 + lo * 6, hi - lo, mask, layout,
		                           static_cast<char *>(hits_out) + lo * (size_t)layout->stride, coherence,
		                           G);
	});
	if (rc)
		return rc;
	for (prt_b200 *d : c->peers) {
		c->last_trace_ms = std::max(c->last_trace_ms, d->last_trace_ms);
		c->last_h2d_bytes += d->last_h2d_bytes;
		c->last_d2h_bytes += d->last_d2h_bytes;
	}
	return PRT_OK;
}

uint64_t prt_b200_last_h2d_bytes(const prt_b200 *c) { return c ? c->last_h2d_bytes : 0; }
uint64_t prt_b200_last_d2h_bytes(const prt_b200 *c) { return c ? c->last_d2h_bytes : 0; }

// Pinned host memory for callers that want the host entry points to DMA directly (no staging).
void *prt_b200_alloc_pinned(size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}
void prt_b200_free_pinned(void *p) {
	if (p)
		cudaFreeHost(p);
}
// Page-lock an existing host range (e.g. a slice of a shared-memory segment that several
// single-GPU processes use as their common ray / hit buffers) so that the host entry points DMA
// directly.  Returns 0 on success.
int prt_b200_host_register(void *p, size_t bytes) {
	const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
	if (e != cudaSuccess)
		cudaGetLastError();
	return e == cudaSuccess ? PRT_OK : PRT_E_CUDA;
}
int prt_b200_host_unregister(void *p) {
	const cudaError_t e = cudaHostUnregister(p);
	if (e != cudaSuccess)
		cudaGetLastError();
	return e == cudaSuccess ? PRT_OK : PRT_E_CUDA;
}

// Bandwidth probe: a persistent grid repeatedly reads `bytes` of device memory with 16-byte
// loads.  With bytes << L2 (126 MB) it measures L2 read bandwidth, with bytes >> L2 HBM read
// bandwidth -- the denominators of the traversal roofline, measured on this box.
int prt_b200_read_bandwidth(prt_b200 *c, uint64_t bytes, int iters, float *gbs) {
	if (!c || !gbs || bytes < 4096 || iters < 1)
		return fail(c, PRT_E_ARG, "read_bandwidth: bad argument");
	PRT_CUDA(c, cudaSetDevice(c->device));
	void *buf = nullptr;
	PRT_CUDA(c, cudaMalloc(&buf, bytes));
	cudaMemsetAsync(buf, 1, bytes, c->stream);
	float ms = 0.f;
	int rc = prt::launch_read_probe(c, buf, bytes, 1, nullptr); // warm
	if (!rc)
		rc = prt::launch_read_probe(c, buf, bytes, iters, &ms);
	cudaFree(buf);
	if (rc)
		return rc;
	*gbs = (float)((double)bytes * iters / (ms * 1e-3) / 1e9);
	return PRT_OK;
}

uint64_t prt_b200_num_tris(const prt_b200 *c) { return c ? c->n_tris : 0; }
uint64_t prt_b200_num_nodes(const prt_b200 *c) { return c ? c->n_nodes : 0; }
int32_t prt_b200_bvh_root(const prt_b200 *c) { return c ? c->root : 0; }
uint64_t prt_b200_bvh_bytes(const prt_b200 *c) {
	return c ? c->n_nodes * sizeof(prt::Node) + c->n_tris * sizeof(prt::TriRec) : 0;
}
int prt_b200_download_wide(const prt_b200 *cc, void *nodes4_out) {
	prt_b200 *c = const_cast<prt_b200 *>(cc);
	if (!c || !nodes4_out)
		return PRT_E_ARG;
	if (!c->wide_built)
		return fail(c, PRT_E_ARG, "download_wide: no wide nodes were built for this scene");
	PRT_CUDA(c, cudaSetDevice(c->device));
	PRT_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->n_nodes)
		PRT_CUDA(c, cudaMemcpy(nodes4_out, c->nodes4.p, c->n_nodes * sizeof(prt::Node4),
		                       cudaMemcpyDeviceToHost));
	return PRT_OK;
}
uint64_t prt_b200_launch_count(const prt_b200 *c) { return c ? c->launches : 0; }
float prt_b200_last_build_ms(const prt_b200 *c) { return c ? c->last_build_ms : 0.f; }
float prt_b200_last_trace_ms(const prt_b200 *c) { return c ? c->last_trace_ms : 0.f; }
float prt_b200_last_kernel_ms(const prt_b200 *c) { return c ? c->last_kernel_ms : 0.f; }
uint64_t prt_b200_exotic_rays(const prt_b200 *c) { return c ? c->exotic_rays : 0; }
uint64_t prt_b200_graph_replays(const prt_b200 *c) { return c ? c->graph_replays : 0; }
uint64_t prt_b200_l2_bytes(const prt_b200 *c) { return c ? c->l2_bytes : 0; }

int prt_b200_download_bvh(const prt_b200 *cc, void *nodes_out, void *tris_out) {
	prt_b200 *c = const_cast<prt_b200 *>(cc);
	if (!c)
		return PRT_E_ARG;
	PRT_CUDA(c, cudaSetDevice(c->device));
	PRT_CUDA(c, cudaStreamSynchronize(c->stream));
	if (nodes_out && c->n_nodes)
		PRT_CUDA(c, cudaMemcpy(nodes_out, c->nodes.p, c->n_nodes * sizeof(prt::Node),
		                       cudaMemcpyDeviceToHost));
	if (tris_out && c->n_tris)
		PRT_CUDA(c, cudaMemcpy(tris_out, c->trirecs.p, c->n_tris * sizeof(prt::TriRec),
		                       cudaMemcpyDeviceToHost));
	return PRT_OK;
}

} // extern "C"
