// prt_ctx.h -- internal context shared by api.cu / build.cu / trace.cu (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/prt_b200.h"
#include "prt_math.cuh"

namespace prt {

// Grow-only device buffer: the reference's GPU backends malloc/free on every call
// (intersect_optix.hpp:73-116); here scratch survives across set_tris / nearest_hits calls so a
// per-frame rebuild (config C5) never touches the allocator.
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap)
			return cudaSuccess;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = bytes + bytes / 8 + 256; // slack so slowly growing inputs do not realloc
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess)
			cap = want;
		return e;
	}
	void release() {
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf() { release(); } // (members of the context: a new buffer cannot be forgotten in destroy)
	template <class T> T *as() const { return static_cast<T *>(p); }
};

struct PinnedBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap)
			return cudaSuccess;
		if (p)
			cudaFreeHost(p);
		p = nullptr;
		cap = 0;
		cudaError_t e = cudaMallocHost(&p, bytes);
		if (e == cudaSuccess)
			cap = bytes;
		return e;
	}
	void release() {
		if (p)
			cudaFreeHost(p);
		p = nullptr;
		cap = 0;
	}
	PinnedBuf() = default;
	PinnedBuf(const PinnedBuf &) = delete;
	PinnedBuf &operator=(const PinnedBuf &) = delete;
	~PinnedBuf() { release(); }
};

class HostPool; // prt_hostpool.h: worker threads that stage pageable caller memory
struct NcclApi; // api.cu: libnccl entry points resolved at run time

} // namespace prt

struct prt_b200 {
	int device = -1;
	int sm_count = 0;
	uint64_t l2_bytes = 0;
	cudaStream_t stream = nullptr;
	// host entry point: one stream per pipeline stage (H2D, kernels, D2H) over a ring of
	// PIPE chunk buffers; ev_pipe[b][stage] = "that stage has left buffer set b"
	static constexpr int PIPE = 4;
	cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr; // around the traversal kernel alone (device entry points)
	cudaEvent_t ev_pipe[PIPE][3] = {};
	std::string name;
	std::string err;

	// multi-GPU (prt_b200_create_multi / env PRT_B200_GPUS): the context the caller holds drives
	// device `device` itself and owns one sub-context per further device; the scene is replicated
	// (triangles broadcast over NVLink, identical deterministic build everywhere) and every host
	// batch is cut into contiguous slices, one per device
	std::vector<prt_b200 *> peers;
	prt::NcclApi *nccl = nullptr;     // resolved lazily at the first broadcast (owner only)
	std::vector<void *> nccl_comms;   // ncclComm_t per device, [0] = this context
	int bcast_mode = 0;               // env PRT_B200_BCAST: 0 NCCL when loadable else peer copies, 1 peer copies
	std::string bcast_used;           // what the last broadcast went through ("nccl", "p2p")
	// host staging workers (pageable caller memory -> pinned ring and back), created on first use
	prt::HostPool *pool_in = nullptr, *pool_out = nullptr;
	int copy_threads = 0;             // env PRT_B200_COPY_THREADS (0 = automatic)
	int bps_cache[32][8] = {};        // resident blocks per SM of each traversal kernel variant

	// scene
	uint64_t n_tris = 0;
	uint64_t n_nodes = 0;
	prt::DevBuf tris_raw;  // staged copy of the caller's 36-byte records (host entry point only)
	prt::DevBuf nodes;     // prt::Node[n_nodes]
	prt::DevBuf trirecs;   // prt::TriRec[n_tris]
	prt::DevBuf nodes4;    // prt::Node4[n_nodes]: compressed 4-wide view of the same tree
	int wide_mode = 2;     // env PRT_B200_WIDE: 0 never built/used, 1 always, 2 incoherent batches only

	// build scratch
	prt::DevBuf keys[2], vals[2], sort_scratch, bounds, leaf_box, bound, root_info;
	// opt-in treelet SAH optimisation (build.cu 5b): parent links, arrival flags, leaf counts, heights
	prt::DevBuf tl_parent, tl_leaf_parent, tl_flag, tl_count, tl_depth, tl_backup, tl_sah;
	int optimise_mode = 3;   // env PRT_B200_TREELET_MODE: 0 never, 1 inside set_tris, 2 lazily, 3 lazily + temporal reuse (default)
	int optimise_passes = 2; // env PRT_B200_TREELET_PASSES
	bool tree_optimised = false;
	bool topology_valid = false;                      // mode 3: parent links match the nodes
	uint64_t lazy_backoff = 1;                        // mode 3: threshold multiplier, doubled by every rejected refit
	double sah_ref = 0.0, last_sah = 0.0;             // mode 3: SAH cost when optimised / after the last refit
	uint64_t refits = 0, refit_rejects = 0;
	uint64_t rays_since_build = 0, strict_fallbacks = 0;
	int max_tree_depth = 96; // what prt_traverse.cuh: STACK_DEPTH is sized for (env PRT_B200_MAX_TREE_DEPTH: tests)
	float last_optimise_ms = 0.f; // device time of the lazy optimisation of the current scene (0 = none yet)
	int32_t tree_depth = 0;       // height of the optimised tree (0 = not optimised)
	int32_t root = 0; // index of the root node (the radix tree numbers nodes by split position)

	// trace scratch
	prt::DevBuf rays_dev[PIPE], hits_dev[PIPE], counter;
	prt::DevBuf stack_ovf[2]; // overflow of the shared-memory traversal stacks, per launch slot
	prt::DevBuf slow_list[2]; // rays the fast kernel set aside for the exact kernel, per launch slot
	prt::DevBuf coop_buf[2], coop_lifo[2]; // cooperative tail (prt_trace_kernel.cuh: k_coop): handed-over rays, LIFOs
	const void *coop_seen[2] = {nullptr, nullptr};
	int coop_min_sp = 0;      // env PRT_B200_COOP_SP: ... for rays with at least this many stacked subtrees
	int coop_blocks = 8;      // env PRT_B200_COOP_BLOCKS: blocks per SM of the follow-up kernel
	int coop_after = 4;       // env PRT_B200_COOP: iterations past the end of the batch before a warp goes cooperative (0 = never)
	// the exact pass of the last EXOTIC_DEFERRED launch (trace.cu: finish_exotic)
	bool pending_exotic = false;
	alignas(16) unsigned char exotic_blob[384] = {}; // its TraceParams
	const void *exotic_fn = nullptr;
	int exotic_cache_slot = 0, exotic_slot = 0;
	uint32_t exotic_mask = 0;
	uint64_t exotic_rays = 0; // rays traced by the exact pass so far
	int morton_bits = 0;      // bits per axis of the current tree's Morton keys
	// CUDA graph of the rebuild chain of a small scene family (build.cu: build_lbvh)
	cudaGraphExec_t build_graph = nullptr;
	alignas(8) unsigned char build_graph_key[128] = {}, build_seen_key[128] = {};
	bool build_seen = false;
	int build_graph_launches = 0;
	int use_graphs = 1;       // env PRT_B200_GRAPHS
	uint64_t graph_replays = 0;
	// ray reordering scratch, one set per launch slot (concurrent launches on different streams)
	struct RaySort {
		prt::DevBuf keys[2], vals[2], scratch;
	} rs[2];
	prt::DevBuf probe_ticket;                    // 2 x u64 block tickets of the coherence probe
	unsigned long long *probe_host = nullptr;    // 2 x u64 mapped pinned flags (host view)
	unsigned long long *probe_dev = nullptr;     // ... and their device view
	int ray_key_ob = 3, ray_key_db = 2; // sort key bits per axis: origin, direction (env PRT_B200_RAYKEY="ob,db"): 15 bits = 2 passes
	int sort_rays = 2; // env PRT_B200_SORT_RAYS: 0 never, 1 always, 2 auto (only incoherent batches)
	float scene_lo[3] = {0.f, 0.f, 0.f}, scene_hi[3] = {0.f, 0.f, 0.f};
	uint64_t sorted_batches = 0, unsorted_batches = 0, wide_batches = 0;
	bool wide_built = false, last_wide = false;
	prt::PinnedBuf rays_pin[PIPE], hits_pin[PIPE];

	prt_trace_opts opts{1, 1e-4f, 64.0f};

	float scene_absmax[3] = {0.f, 0.f, 0.f}; // largest |coordinate| per axis (fast box-test margin)
	int watertight = 0;                      // env PRT_B200_WATERTIGHT / prt_b200_set_triangle_test: takes effect at set_tris
	bool recs_vertex_form = false;           // the current triangle records hold v1, v2 (watertight) instead of the edges
	int fast_boxes = 1;                      // env PRT_B200_FAST_BOXES=0 forces the exact test everywhere
	int refill = 16;                         // env PRT_B200_REFILL: dynamic ray-fetch threshold (lanes), binary-node kernels
	int refill_scattered = 8;                // env PRT_B200_REFILL_SCATTERED: ... for reordered batches and trees beyond half of L2 (C3B +1.6 %, C5 +4.8 %)
	int refill_wide = 24;                    // env PRT_B200_REFILL_WIDE: ... of the wide-node kernels (incoherent rays: measured +3.5 %)
	int leaf_votes_wide = 4;                 // env PRT_B200_LEAF_VOTES_WIDE: ... in the wide-node kernels (C4 +2.3 % over 8)
	int leaf_votes = 8;                      // env PRT_B200_LEAF_VOTES: lanes waiting at a triangle that start a leaf phase
	bool packed_d2h = true;                  // env PRT_B200_PACKED_D2H: pageable results come back tightly packed
	int chunk_log2 = 0; // host entry point: rays per pipeline chunk (0 = automatic)
	bool pipe_trace = false; // env PRT_B200_PIPE_TRACE: print the stage timeline of every host call
	uint64_t launches = 0;
	float last_build_ms = 0.f, last_trace_ms = 0.f;
	float last_kernel_ms = 0.f; // the traversal kernel of the last device-resident trace, without reordering
	uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0; // what the last host call moved over PCIe
};

namespace prt {

inline int fail(prt_b200 *c, int code, const char *what, cudaError_t e = cudaSuccess) {
	if (c) {
		c->err = what;
		if (e != cudaSuccess) {
			c->err += ": ";
			c->err += cudaGetErrorString(e);
		}
	}
	return code;
}

#define PRT_CUDA(ctx, call)                                                                        \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if (e__ != cudaSuccess)                                                                    \
			return prt::fail(ctx, e__ == cudaErrorMemoryAllocation ? PRT_E_OOM : PRT_E_CUDA,       \
			                 #call, e__);                                                          \
	} while (0)

// build.cu
constexpr uint64_t LAZY_RAYS_PER_TRI = 32;  // lazy tree optimisation threshold: max(32 rays per
constexpr uint64_t LAZY_MIN_RAYS = 8u << 20; // triangle, 8 Mi rays) since the last set_tris
int build_lbvh(prt_b200 *c, const float *d_tris9, uint64_t n);
constexpr double REFIT_TOLERANCE = 1.25;     // mode 3: accepted growth of the SAH cost under refitting
int optimise_tree(prt_b200 *c, cudaStream_t s);
int try_reuse_topology(prt_b200 *c, const float *d_tris9, uint64_t n, cudaStream_t s, bool *reused);
int maybe_optimise_tree(prt_b200 *c, uint64_t n_rays);
// trace.cu
struct TraceOut {
	// SoA (aos == nullptr) or AoS (aos != nullptr)
	prt_soa_out soa{};
	void *aos = nullptr;
	prt_hit_layout layout{};
	int slot = 0; // which ray counter to use (concurrent launches on different streams)
};
// coherence: -1 = probe the batch on the device (one stream synchronisation), 0 = known coherent,
// 1 = known incoherent (reorder it); exotic_mode: how the rays the fast kernel cannot take reach
// the exact kernel (trace.cu)
enum { EXOTIC_INLINE = 0, EXOTIC_DEFERRED = 1 };
int launch_trace(prt_b200 *c, const float *d_rays6, uint64_t n, uint32_t mask, const TraceOut &out,
                 uint32_t *d_counts, cudaStream_t stream, int coherence = -1,
                 int exotic_mode = EXOTIC_INLINE);
int finish_exotic(prt_b200 *c, cudaStream_t stream, bool *ran);
int host_ray_probe(const prt_b200 *c, const float *rays6, uint64_t n);

// sort.cu
int radix_sort_pairs(prt_b200 *c, DevBuf &scratch, uint64_t *const keys[2], uint32_t *const vals[2],
                     uint64_t n, int key_bits, cudaStream_t s, int *result_index);
int radix_sort_prepare32(prt_b200 *c, DevBuf &scratch, uint64_t n, int key_bits, cudaStream_t s,
                         uint32_t **ghist, int *passes);
int radix_sort_run32_identity(prt_b200 *c, DevBuf &scratch, uint32_t *const keys[2], uint32_t *const vals[2],
                              uint64_t n, int key_bits, cudaStream_t s, int *result_index);

int launch_read_probe(prt_b200 *c, const void *buf, uint64_t bytes, int iters, float *ms);

} // namespace prt
