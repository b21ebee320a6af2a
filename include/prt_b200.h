/*
 * prt_b200.h -- C ABI of the B200-native (sm_100a) nearest-hit backend for portableRT.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point
 * names the reference interface it stands behind (file:line in the portableRT checkout).  The C++
 * plugin class `portableRT::CUDABackend` (include/portableRT/intersect_cuda.hpp in this repo) and
 * the Python host mirror (portablert_b200/) are thin callers of exactly these functions.
 *
 * Conventions: every int-returning call returns PRT_OK (0) or a PRT_E_* code and never throws;
 * prt_b200_last_error() gives the message.  A context is used from one host thread at a time, like
 * the reference's unsynchronised globals (backend.hpp:39,73).  There is NO CPU fallback: without a
 * compute-capability-10.x device prt_b200_create fails with PRT_E_NO_DEVICE.
 */
#ifndef PRT_B200_H
#define PRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRT_B200_ABI_VERSION 3 /* 2: tree optimisation, triangle test; 3: multi-GPU contexts */

enum {
	PRT_OK = 0,
	PRT_E_NO_DEVICE = 1, /* no CC 10.x GPU visible */
	PRT_E_CUDA = 2,      /* a CUDA runtime call or kernel failed */
	PRT_E_ARG = 3,       /* bad argument (NULL, bad tag mask, bad layout) */
	PRT_E_OOM = 4,       /* device or pinned-host allocation failed */
	PRT_E_LIMIT = 5      /* more than 2^31-2 triangles */
};

/* Tag mask bits, in the reference's canonical tag order uv,t,primitive_id,p,valid
 * (hitreg.hpp:24-26, filter tags hitreg.hpp:48-54).  One of 31 kernel specialisations is picked
 * per non-empty mask -- the analogue of the 31 OptiX raygen programs (hitreg.hpp:142-177). */
enum { PRT_TAG_UV = 1, PRT_TAG_T = 2, PRT_TAG_PID = 4, PRT_TAG_P = 8, PRT_TAG_VALID = 16, PRT_TAG_ALL = 31 };

/* Byte layout of one HitReg<Tags...> record (hitreg.hpp:29-44); -1 = field absent.  The C++ shim
 * fills it with sizeof/offsetof of the reference's own type, so nothing is hard-coded here. */
typedef struct {
	uint32_t stride;
	int32_t off_u, off_v, off_t, off_pid, off_valid, off_px, off_py, off_pz;
} prt_hit_layout;

/* SoA outputs of the device-resident entry point; a NULL pointer for a tag absent from the mask.
 * uv: float2[n]; t: float[n]; pid: uint32[n]; p: float[3n] packed xyz; valid: uint8[n]. */
typedef struct {
	float *uv;
	float *t;
	uint32_t *pid;
	float *p;
	uint8_t *valid;
} prt_soa_out;

/* Traversal knobs (there is no other configuration channel in the reference API; the C++ backend
 * reads the same knobs from PRT_B200_* environment variables in init()). */
/* Pruning is exact in real arithmetic; in binary32 it relies on the slack below covering the
 * rounding of Moeller-Trumbore's t against the slab test of the boxes around the triangle.  That
 * holds on every test of this repository, including needle / sliver triangles (aspect ratio up to
 * 1e7) under grazing and edge-parallel rays (tests/test_emu_parity.py:
 * test_pruning_on_slivers_and_grazing_rays), but it is an empirical bound, not a proof: prune = 0
 * visits every box the line touches, exactly like the reference, for callers who need that. */
typedef struct {
	int prune;        /* 1 (default): cull subtrees whose entry distance exceeds the best hit so far
	                     (+ slack); 0: visit every box the line touches, like bvh.hpp:224-265 */
	float slack_rel;  /* pruning slack, relative to |t_best| (default 1e-4) */
	float slack_ulps; /* extra absolute slack in units of ulp(|origin|/|direction|) (default 64) */
} prt_trace_opts;

typedef struct prt_b200 prt_b200; /* opaque: owns streams, scratch, the BVH, staging buffers */

/* Backend::is_available (backend.hpp:20), evaluated during static initialisation
 * (backend.hpp:85-95): number of visible devices with compute capability 10.x; 0 when there is no
 * driver/GPU.  Never fails, never throws, cheap. */
int prt_b200_device_count(void);

/* Backend::init (backend.hpp:21), called by select_backend (src/backend.cpp:50-56), possibly more
 * than once.  device >= 0: a single-GPU context on that device.  device < 0: what the environment
 * asks for -- PRT_B200_GPUS devices (default 1) starting at PRT_B200_DEVICE (default: the first
 * CC 10.x device) -- so that a program written against the reference API spreads over several
 * GPUs without a source change. */
int prt_b200_create(prt_b200 **out, int device);

/* Multi-GPU context: n_gpus CC 10.x devices driven from this one process (n_gpus <= 0: env
 * PRT_B200_GPUS, default 1).  The reference API has one scene and one ray batch per call
 * (backend.hpp:19,75-83) and rays are independent, so the scene replicates and the batch shards:
 *   set_tris      the triangles go to the first device once and are broadcast to the others over
 *                 NVLink (one grouped ncclBroadcast on single-process communicators; libnccl is
 *                 resolved at run time, cudaMemcpyPeerAsync where it cannot be loaded or
 *                 PRT_B200_BCAST=p2p); every device runs the same deterministic build, so all
 *                 replicas of the BVH are bit-identical
 *   nearest_hits  device i traces the contiguous slice [i*n/G, (i+1)*n/G) of the caller's host
 *                 batch through its own copy pipeline (one host thread per device) and writes its
 *                 records at the slice's offset of the caller's array: hits arrive in ray order,
 *                 there is no gather step and no collective on the data path
 * Results are byte-for-byte those of a single-GPU context.  The device-resident entry points
 * (prt_b200_set_tris_dev takes a pointer on the first device and broadcasts from there;
 * prt_b200_trace_dev*, prt_b200_download_*) act on the first device. */
int prt_b200_create_multi(prt_b200 **out, int n_gpus);
int prt_b200_num_devices(const prt_b200 *ctx);
const char *prt_b200_broadcast_path(const prt_b200 *ctx); /* "nccl" / "p2p" / "" (nothing broadcast yet) */

/* Backend::shutdown (backend.hpp:22); NULL-safe, frees all device and pinned memory. */
void prt_b200_destroy(prt_b200 *ctx);

/* Backend::device_name (backend.hpp:23); returns the length written (excluding NUL). */
int prt_b200_device_name(const prt_b200 *ctx, char *buf, size_t cap);

/* Backend::set_tris(const Tris&) (backend.hpp:19; CPU: src/intersect_cpu.cpp:12 -> BVH2::build,
 * bvh.hpp:158-193).  tris9: HOST pointer, n_tris records of 9 floats (v0,v1,v2 xyz, 36-byte
 * stride, core.hpp:24).  Copies to the device and builds the LBVH there.  n_tris == 0 is legal
 * (every ray then misses, bvh.hpp:136-137). */
int prt_b200_set_tris(prt_b200 *ctx, const float *tris9, uint64_t n_tris);

/* Same, triangles already resident on the context's device (device pointer, same 36-byte
 * records; the buffer is only read).  *build_ms (may be NULL) = device time of the whole build
 * measured with CUDA events on the context's stream. */
int prt_b200_set_tris_dev(prt_b200 *ctx, const float *d_tris9, uint64_t n_tris, float *build_ms);

/* nearest_hits<Tags...>(const std::vector<Ray>&) -- the member template every backend provides
 * (CPU: intersect_cpu.hpp:20-43; OptiX: intersect_optix.hpp:64-119) and that the free function
 * reaches through std::visit (nearest_hits_impl.hpp:28-36).  rays6: HOST pointer, n_rays records
 * of 6 floats (origin, direction; 24-byte stride, core.hpp:19-22).  hits_out: HOST pointer to
 * n_rays records laid out as `layout` describes (the caller's std::vector<HitReg<Tags...>>).
 * Semantics are those of BVH2::nearest_tri (bvh.hpp:224-265): minimum t over all triangles that
 * pass ray_box_intersect on their own AABB (bvh.hpp:195-222) and intersect_tri (core.hpp:27-65),
 * negative t included, no tmax; t = +inf, valid = false, p = o + inf*d on a miss.  Fields the
 * reference leaves indeterminate on a miss are written as u = v = 0, primitive_id = 0xFFFFFFFF. */
int prt_b200_nearest_hits(prt_b200 *ctx, const float *rays6, uint64_t n_rays, uint32_t tag_mask,
                          const prt_hit_layout *layout, void *hits_out);

/* bytes the last prt_b200_nearest_hits moved host->device and device->host (all devices).  Results
 * for pageable caller memory cross PCIe tightly packed (requested fields only) and are scattered
 * into the caller's records by the staging threads; pinned caller memory receives the records by
 * DMA as they are. */
uint64_t prt_b200_last_h2d_bytes(const prt_b200 *ctx);
uint64_t prt_b200_last_d2h_bytes(const prt_b200 *ctx);

/* Page-locked host memory.  prt_b200_set_tris / prt_b200_nearest_hits detect pinned (or
 * cudaHostRegister'ed) buffers and DMA from/to them directly; pageable memory (e.g. a plain
 * std::vector, as in the reference API) is staged through internal pinned buffers. */
void *prt_b200_alloc_pinned(size_t bytes);
void prt_b200_free_pinned(void *p);
/* Page-lock / unlock an existing host range (cudaHostRegister), e.g. this process' slice of a
 * shared-memory segment holding the whole ray batch of a multi-GPU job. */
int prt_b200_host_register(void *p, size_t bytes);
int prt_b200_host_unregister(void *p);

/* Device-resident traversal for device-timed measurement: rays and outputs live on the context's
 * device.  *trace_ms (may be NULL) = device time of the traversal kernel(s), CUDA events on the
 * context's stream.  Fields not in tag_mask are not written (and may be NULL). */
int prt_b200_trace_dev(prt_b200 *ctx, const float *d_rays6, uint64_t n_rays, uint32_t tag_mask,
                       const prt_soa_out *d_out, float *trace_ms);

/* Same traversal, AoS records written on the device at d_hits (layout as for nearest_hits). */
int prt_b200_trace_dev_aos(prt_b200 *ctx, const float *d_rays6, uint64_t n_rays,
                           uint32_t tag_mask, const prt_hit_layout *layout, void *d_hits,
                           float *trace_ms);

/* Traversal knobs; opts == NULL restores the defaults. */
int prt_b200_set_trace_opts(prt_b200 *ctx, const prt_trace_opts *opts);

/* SAH optimisation of the LBVH: `passes` bottom-up rounds of treelet restructuring (Karras & Aila
 * 2013: 7-leaf treelets, exact dynamic programming over their topologies, minimising the summed
 * surface area of the internal nodes -- the quantity the reference's binned SAH sweep,
 * bvh.hpp:58-129, minimises greedily top-down).  Results of nearest_hits do not depend on it, only
 * the number of boxes visited per ray (measured: -6 % on C2, -46 % on the C3 interior).
 *   mode 0  never: set_tris leaves the radix tree as built (fastest build)
 *   mode 1  inside every set_tris
 *   mode 2  lazily: once a scene has been asked for max(32 rays per triangle, 8 Mi rays) -- about
 *           when tracing the plain tree has cost what the optimisation costs -- it is optimised
 *           before the next batch is traced; a static scene pays once
 *   mode 3  (default) mode 2 plus temporal reuse for deforming meshes: rays are counted per scene
 *           FAMILY (consecutive set_tris calls with the same triangle count), and once the family's
 *           tree has been optimised, the next set_tris of that size refits the optimised topology
 *           to the new vertices (one kernel) instead of rebuilding.  The refitted tree is kept only
 *           if its SAH cost stays within 1.25x of the cost it had when optimised; otherwise the
 *           plain LBVH is rebuilt, the family's counter starts over and its threshold doubles, so
 *           that alternating unrelated scenes of one size costs a vanishing number of wasted
 *           optimisations.  Results never depend on any of this: every box is the exact union of
 *           what is below.  prt_b200_refits / prt_b200_refit_rejects count both cases.
 * passes: 1..8 (default 2).  Env PRT_B200_TREELET_MODE / PRT_B200_TREELET_PASSES.
 * prt_b200_tree_depth: height of the optimised tree (0 = the current tree is the plain radix tree);
 * prt_b200_last_optimise_ms: device time the lazy optimisation of the current scene took.
 * The traversal stack is sized for the radix tree's height bound (96); an optimised tree that comes
 * out taller is discarded and the optimisation repeated under a height-preserving rule
 * (prt_b200_strict_fallbacks counts these; none on any benchmark scene). */
int prt_b200_set_tree_optimisation(prt_b200 *ctx, int mode, int passes);
int32_t prt_b200_tree_depth(const prt_b200 *ctx);
float prt_b200_last_optimise_ms(const prt_b200 *ctx);
uint64_t prt_b200_strict_fallbacks(const prt_b200 *ctx);
uint64_t prt_b200_refits(const prt_b200 *ctx);
uint64_t prt_b200_refit_rejects(const prt_b200 *ctx);

/* Triangle test: 0 (default) = the reference's Moeller-Trumbore arithmetic replayed operation for
 * operation (core.hpp:27-65) -- results identical to the reference CPU backend; 1 = opt-in
 * WATERTIGHT test (Woop, Benthin, Wald 2013): rays through an edge or vertex shared by two
 * triangles can no longer slip between them, at the price of results that differ from the
 * reference on exactly those grazing rays and by rounding of t, u, v.  Same result domain as the
 * reference otherwise (signed t, minimum wins, no tmax).  Env PRT_B200_WATERTIGHT; takes effect at
 * the next set_tris (the triangle records then keep the original vertices).
 * prt_b200_triangle_test returns the mode the current scene was built for. */
int prt_b200_set_triangle_test(prt_b200 *ctx, int mode);
int prt_b200_triangle_test(const prt_b200 *ctx);

/* Ray reordering in front of the traversal (results keep the caller's ray order): 0 = never,
 * 1 = always, 2 = automatic (default; env PRT_B200_SORT_RAYS): batches of >= 65 536 rays are
 * sorted by a 24-bit origin/direction key unless most neighbouring rays already share their key.
 * prt_b200_sorted_batches counts the launches that were actually reordered. */
int prt_b200_set_ray_sorting(prt_b200 *ctx, int mode);

/* Compressed 4-wide nodes (64 B per four children, 8-bit quantised boxes, used only for
 * conservative culling): 0 = never build/use them, 1 = every batch, 2 = batches that were found
 * incoherent and reordered (default; env PRT_B200_WIDE).  Takes effect at the next set_tris. */
int prt_b200_set_wide_nodes(prt_b200 *ctx, int mode);
/* copies the wide node array (64 B per binary node index) to the host, for structural tests */
int prt_b200_download_wide(const prt_b200 *ctx, void *nodes4_out);
uint64_t prt_b200_sorted_batches(const prt_b200 *ctx);

/* Instrumented traversal (never used in timed runs): per ray, counts[2i] = internal nodes
 * fetched, counts[2i+1] = triangles tested; feeds the bytes/ray figure of the roofline
 * (SURVEY.md 8d).  d_counts: device pointer to 2*n_rays uint32. */
int prt_b200_trace_count_dev(prt_b200 *ctx, const float *d_rays6, uint64_t n_rays,
                             uint32_t *d_counts);

/* Introspection for tests and the bench. */
uint64_t prt_b200_num_tris(const prt_b200 *ctx);
uint64_t prt_b200_num_nodes(const prt_b200 *ctx);
int32_t prt_b200_bvh_root(const prt_b200 *ctx);       /* index of the root node in the node array */
uint64_t prt_b200_bvh_bytes(const prt_b200 *ctx);      /* node array + triangle array on device */
uint64_t prt_b200_launch_count(const prt_b200 *ctx);   /* kernels launched by this context so far */
float prt_b200_last_build_ms(const prt_b200 *ctx);     /* device time of the last build */
float prt_b200_last_trace_ms(const prt_b200 *ctx);     /* device time of the last traversal (incl. ray reordering) */
float prt_b200_last_kernel_ms(const prt_b200 *ctx);    /* ... of its traversal kernel alone (device entry points) */
uint64_t prt_b200_exotic_rays(const prt_b200 *ctx);    /* rays traced by the exact second pass so far (device entry points) */
uint64_t prt_b200_l2_bytes(const prt_b200 *ctx);       /* L2 cache size of the context's device */
uint64_t prt_b200_graph_replays(const prt_b200 *ctx);  /* builds served by replaying the captured CUDA graph */
/* Copies the built BVH back to the host for structural tests: nodes (64 B each), triangle records
 * (64 B each).  Either pointer may be NULL. */
int prt_b200_download_bvh(const prt_b200 *ctx, void *nodes_out, void *tris_out);

/* Roofline denominators measured on the box: read bandwidth (GB/s) of `bytes` of device memory
 * re-read `iters` times by a persistent grid (bytes << L2 size: L2 bandwidth; >> L2: HBM). */
int prt_b200_read_bandwidth(prt_b200 *ctx, uint64_t bytes, int iters, float *gbs);

const char *prt_b200_last_error(const prt_b200 *ctx); /* ctx may be NULL: last create() error */
int prt_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PRT_B200_H */
