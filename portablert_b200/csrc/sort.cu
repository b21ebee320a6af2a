// sort.cu -- LSD radix sort of (u64 or u32 key, u32 value) pairs, 8-bit digits, one kernel per pass.
//
// Used by the LBVH build (Morton keys -> triangle order) and by the ray reordering in front of the
// traversal.  "Onesweep" organisation: one read of the keys produces the digit histograms of ALL
// passes; each pass is then a single kernel in which a tile (block) ranks its keys in shared
// memory with warp-level match/prefix operations, obtains the global offset of each of its 256
// digit runs by a decoupled look-back over the preceding tiles' published counts (one thread per
// digit), and scatters the tile through shared memory so that consecutive threads write
// consecutive addresses.  HBM traffic per pass: read (8+4) B + write (8+4) B per pair.
// Tiles take their index from an atomic counter, so a tile's predecessors are always already
// running: the look-back cannot deadlock.
#include <algorithm>

#include "prt_ctx.h"

namespace prt {

constexpr int SO_THREADS = 256;
constexpr int SO_ITEMS = 12;
constexpr int SO_TILE = SO_THREADS * SO_ITEMS; // 3072 pairs per tile
constexpr int SO_WARPS = SO_THREADS / 32;
constexpr int SO_RADIX = 256;
constexpr int SO_MAX_PASSES = 8;

constexpr unsigned long long ST_LOCAL = 1ull << 62;  // tile's own count is published
constexpr unsigned long long ST_PREFIX = 2ull << 62; // inclusive prefix over tiles 0..t is published
constexpr unsigned long long ST_MASK = (1ull << 62) - 1;

// ---- all passes' digit histograms in one read of the keys ---------------------------------------
template <class K>
__global__ void __launch_bounds__(SO_THREADS)
    k_sweep_hist(const K *__restrict__ keys, uint64_t n, int passes,
                 uint32_t *__restrict__ ghist /* [passes][256] */) {
	__shared__ uint32_t sh[SO_MAX_PASSES][SO_RADIX];
	for (int p = 0; p < passes; ++p)
		sh[p][threadIdx.x] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * SO_THREADS;
	for (uint64_t i = (uint64_t)blockIdx.x * SO_THREADS + threadIdx.x; i < n; i += stride) {
		const K k = keys[i];
		const unsigned act = __activemask();
		for (int p = 0; p < passes; ++p) {
			// a digit that is constant across the warp (e.g. unused high bits) would serialise
			// 32 same-address atomics: aggregate it into one
			const uint32_t d = (uint32_t)(k >> (8 * p)) & 0xff;
			int same;
			__match_all_sync(act, d, &same);
			if (same) {
				if ((threadIdx.x & 31) == (__ffs(act) - 1))
					atomicAdd(&sh[p][d], (uint32_t)__popc(act));
			} else {
				atomicAdd(&sh[p][d], 1u);
			}
		}
	}
	__syncthreads();
	for (int p = 0; p < passes; ++p) {
		const uint32_t c = sh[p][threadIdx.x];
		if (c)
			atomicAdd(&ghist[p * SO_RADIX + threadIdx.x], c);
	}
}

// block-wide exclusive scan of one value per thread (SO_THREADS threads)
__device__ __forceinline__ uint32_t block_exscan256(uint32_t v, uint32_t *ws) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o)
			inc += t;
	}
	if (lane == 31)
		ws[w] = inc;
	__syncthreads();
	if (w == 0) {
		const uint32_t s = lane < SO_WARPS ? ws[lane] : 0;
		uint32_t si = s;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
			if (lane >= o)
				si += t;
		}
		ws[lane] = si - s;
	}
	__syncthreads();
	const uint32_t res = ws[w] + inc - v;
	__syncthreads();
	return res;
}

// ---- one pass --------------------------------------------------------------------------------
// (u32 keys: capped at 64 registers for 4 resident blocks, 16 bytes of spill: the ray reordering of
// C4 2.47 -> 2.29 ms; the u64 instantiation would spill 96 bytes and keeps its 80 registers)
template <class K>
__global__ void __launch_bounds__(SO_THREADS, (sizeof(K) == 4 ? 4 : 3))
    k_sweep_pass(const K *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                 K *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint64_t n,
                 int shift, const uint32_t *__restrict__ digit_ofs /* [256] pass histogram */,
                 unsigned long long *status /* [n_tiles][256] */, uint32_t *tile_counter) {
	__shared__ K s_keys[SO_TILE];
	__shared__ uint32_t s_vals[SO_TILE];
	__shared__ uint32_t wh[SO_WARPS][SO_RADIX];
	__shared__ uint32_t lbase[SO_RADIX];
	__shared__ long long gofs[SO_RADIX];
	__shared__ uint32_t ws[32];
	__shared__ uint32_t s_tile;

	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		s_tile = atomicAdd(tile_counter, 1u);
#pragma unroll
	for (int k = 0; k < SO_WARPS; ++k)
		wh[k][threadIdx.x] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint64_t tile_base = (uint64_t)tile * SO_TILE;
	const int tile_cnt = (int)min((uint64_t)SO_TILE, n - tile_base);

	// warp-striped arrangement: warp w owns tile positions [w*32*ITEMS, (w+1)*32*ITEMS); item k of
	// lane l sits at w*32*ITEMS + k*32 + l, so memory order == (k, lane) order inside a warp
	K key[SO_ITEMS];
	uint32_t val[SO_ITEMS];
	uint32_t rank[SO_ITEMS];
	const int wbase = w * 32 * SO_ITEMS;
#pragma unroll
	for (int k = 0; k < SO_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		const bool ok = pos < tile_cnt;
		key[k] = ok ? keys_in[tile_base + pos] : (K)~(K)0;
		// (vals_in == nullptr: the first pass of a sort whose values are the positions 0..n-1)
		val[k] = ok ? (vals_in ? vals_in[tile_base + pos] : (uint32_t)(tile_base + pos)) : 0u;
	}
	// ---- early counts: warp-private digit histograms, so that the tile's counts can be published
	// (and the successors' look-back can proceed) before the slower ranking below
#pragma unroll
	for (int k = 0; k < SO_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		if (pos < tile_cnt)
			atomicAdd(&wh[w][(uint32_t)(key[k] >> shift) & 0xff], 1u);
	}
	__syncthreads();
	// per digit (thread == digit): exclusive scan over the warps, tile count
	uint32_t cnt = 0;
#pragma unroll
	for (int k = 0; k < SO_WARPS; ++k) {
		const uint32_t c = wh[k][threadIdx.x];
		wh[k][threadIdx.x] = cnt;
		cnt += c;
	}
	unsigned long long *my = status + (uint64_t)tile * SO_RADIX + threadIdx.x;
	__stcg(my, (tile == 0 ? ST_PREFIX : ST_LOCAL) | cnt);
	__syncthreads();

	// ---- ranking: position of every key among the keys of its digit inside the tile (stable).
	// Peers (lanes holding the same digit in this round) are found with 8 ballots -- MATCH.ANY
	// costs one iteration per distinct value, ~30 for random digits.
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int k = 0; k < SO_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		const bool ok = pos < tile_cnt;
		const uint32_t d = (uint32_t)(key[k] >> shift) & 0xff;
		uint32_t peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
		for (int bit = 0; bit < 8; ++bit) {
			const bool one = (d >> bit) & 1u;
			const uint32_t bal = __ballot_sync(0xffffffffu, one);
			peers &= one ? bal : ~bal;
		}
		uint32_t before = 0;
		if (ok)
			before = wh[w][d]; // running offset of this digit inside the warp's part of the tile
		__syncwarp();
		rank[k] = before + __popc(peers & lt);
		if (ok && lane == (__ffs(peers) - 1))
			wh[w][d] = before + __popc(peers);
		__syncwarp();
	}

	// ---- decoupled look-back: sum of this digit's counts over the preceding tiles
	unsigned long long prefix = 0;
	if (tile > 0) {
		long long t = (long long)tile - 1;
		while (true) {
			const unsigned long long v = __ldcg(status + (uint64_t)t * SO_RADIX + threadIdx.x);
			if ((v >> 62) == 0)
				continue;
			prefix += v & ST_MASK;
			if ((v >> 62) == 2 || t == 0)
				break;
			--t;
		}
		__stcg(my, ST_PREFIX | (prefix + cnt));
	}
	const uint32_t lb = block_exscan256(cnt, ws);
	lbase[threadIdx.x] = lb;
	// global start of each digit = exclusive scan of the pass histogram (256 values: every tile
	// redoes it instead of paying a separate launch)
	const uint32_t dstart = block_exscan256(digit_ofs[threadIdx.x], ws);
	gofs[threadIdx.x] = (long long)dstart + (long long)prefix - (long long)lb;
	__syncthreads();

	// stage the tile in digit order (stable), then write runs of consecutive addresses
#pragma unroll
	for (int k = 0; k < SO_ITEMS; ++k) {
		const int pos = wbase + k * 32 + lane;
		if (pos < tile_cnt) {
			const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
			const uint32_t lp = lbase[d] + rank[k];
			s_keys[lp] = key[k];
			s_vals[lp] = val[k];
		}
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < SO_ITEMS; ++k) {
		const int i = k * SO_THREADS + threadIdx.x;
		if (i < tile_cnt) {
			const K kk = s_keys[i];
			const uint32_t d = (uint32_t)((kk >> shift) & 0xff);
			const long long g = gofs[d] + i;
			if (keys_out) // (nullptr: the last pass of a sort whose caller only wants the values)
				keys_out[g] = kk;
			vals_out[g] = s_vals[i];
		}
	}
}

// Sorts n pairs by the low `key_bits` bits of the key (stable).  keys[0]/vals[0] hold the input;
// the two buffers ping-pong; returns the index (0/1) of the buffer that holds the result.
// Scratch layout: [passes][256] digit histograms, 64 bytes of tile counters, [passes][tiles][256]
// look-back status words.
struct SortScratch {
	uint32_t *ghist, *ctr;
	unsigned long long *status;
	uint64_t n_tiles;
	int passes;
};
template <class K>
static int sort_prepare(prt_b200 *c, DevBuf &scratch, uint64_t n, int key_bits, cudaStream_t s,
                        SortScratch &sc) {
	sc.passes = std::min<int>(sizeof(K), (key_bits + 7) / 8);
	sc.n_tiles = (n + SO_TILE - 1) / SO_TILE;
	const size_t hist_b = (size_t)SO_MAX_PASSES * SO_RADIX * 4, ctr_b = 64;
	const size_t status_b = (size_t)sc.passes * sc.n_tiles * SO_RADIX * 8;
	PRT_CUDA(c, scratch.reserve(hist_b + ctr_b + status_b));
	char *base = scratch.as<char>();
	sc.ghist = reinterpret_cast<uint32_t *>(base);
	sc.ctr = reinterpret_cast<uint32_t *>(base + hist_b);
	sc.status = reinterpret_cast<unsigned long long *>(base + hist_b + ctr_b);
	PRT_CUDA(c, cudaMemsetAsync(base, 0, hist_b + ctr_b + status_b, s));
	return PRT_OK;
}

// Sorts n pairs by the low `key_bits` bits of the key (stable).  keys[0]/vals[0] hold the input;
// the two buffers ping-pong; returns the index (0/1) of the buffer that holds the result.
// prepared != nullptr: the caller has called sort_prepare and accumulated the digit histograms
// itself (the ray-key kernel does, saving one read of the keys).  identity_vals: vals[0] is not
// read, the values are the positions 0..n-1.
template <class K>
static int radix_sort_pairs_t(prt_b200 *c, DevBuf &scratch, K *const keys[2], uint32_t *const vals[2],
                              uint64_t n, int key_bits, cudaStream_t s, int *result_index,
                              const SortScratch *prepared = nullptr, bool identity_vals = false,
                              bool keep_keys = true) {
	*result_index = 0;
	if (n <= 1 || key_bits <= 0)
		return PRT_OK;
	SortScratch sc;
	if (prepared) {
		sc = *prepared;
	} else {
		if (int rc = sort_prepare<K>(c, scratch, n, key_bits, s, sc))
			return rc;
		const int hgrid = (int)std::min<uint64_t>((n + SO_THREADS * 8 - 1) / (SO_THREADS * 8),
		                                          (uint64_t)c->sm_count * 8);
		k_sweep_hist<K><<<hgrid, SO_THREADS, 0, s>>>(keys[0], n, sc.passes, sc.ghist);
		c->launches += 1;
	}
	int cur = 0;
	for (int p = 0; p < sc.passes; ++p) {
		k_sweep_pass<K><<<(unsigned)sc.n_tiles, SO_THREADS, 0, s>>>(
		    keys[cur], (p == 0 && identity_vals) ? nullptr : vals[cur],
		    (p == sc.passes - 1 && !keep_keys) ? nullptr : keys[cur ^ 1], vals[cur ^ 1], n,
		    8 * p, sc.ghist + p * SO_RADIX, sc.status + (size_t)p * sc.n_tiles * SO_RADIX, sc.ctr + p);
		c->launches += 1;
		cur ^= 1;
	}
	PRT_CUDA(c, cudaGetLastError());
	*result_index = cur;
	return PRT_OK;
}

int radix_sort_pairs(prt_b200 *c, DevBuf &scratch, uint64_t *const keys[2], uint32_t *const vals[2],
                     uint64_t n, int key_bits, cudaStream_t s, int *result_index) {
	return radix_sort_pairs_t<uint64_t>(c, scratch, keys, vals, n, key_bits, s, result_index);
}
// 32-bit keys (the ray reordering: <= 32 key bits; a pass moves 8 + 8 bytes per pair instead of
// 12 + 12), in two steps for a caller that produces the keys on the device and can count their
// digits while it does (trace.cu: k_ray_keys): prepare hands out the zeroed histogram
// ([passes][256], digit p = bits 8p..8p+7), run sorts with values = positions.
int radix_sort_prepare32(prt_b200 *c, DevBuf &scratch, uint64_t n, int key_bits, cudaStream_t s,
                         uint32_t **ghist, int *passes) {
	SortScratch sc;
	if (int rc = sort_prepare<uint32_t>(c, scratch, n, key_bits, s, sc))
		return rc;
	*ghist = sc.ghist;
	*passes = sc.passes;
	return PRT_OK;
}
int radix_sort_run32_identity(prt_b200 *c, DevBuf &scratch, uint32_t *const keys[2], uint32_t *const vals[2],
                              uint64_t n, int key_bits, cudaStream_t s, int *result_index) {
	SortScratch sc; // same layout as prepare computed (scratch was reserved and zeroed there)
	sc.passes = std::min<int>(sizeof(uint32_t), (key_bits + 7) / 8);
	sc.n_tiles = (n + SO_TILE - 1) / SO_TILE;
	const size_t hist_b = (size_t)SO_MAX_PASSES * SO_RADIX * 4, ctr_b = 64;
	char *base = scratch.as<char>();
	sc.ghist = reinterpret_cast<uint32_t *>(base);
	sc.ctr = reinterpret_cast<uint32_t *>(base + hist_b);
	sc.status = reinterpret_cast<unsigned long long *>(base + hist_b + ctr_b);
	return radix_sort_pairs_t<uint32_t>(c, scratch, keys, vals, n, key_bits, s, result_index, &sc, true,
	                                    false); // (the traversal reads the permutation only)
}

} // namespace prt
