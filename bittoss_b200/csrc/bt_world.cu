/*
 * bt_world.cu — host side of the C ABI (include/bittoss_b200.h): device memory, the
 * per-tick kernel schedule, upload/download between the caller's plain arrays and the
 * device SoA.  One bt_world = one CUDA stream; every kernel of a world runs on it.
 *
 * Tick schedule (reference order main.cpp:1072-1097, ent.cpp:842-911):
 *   k_pre                       flushCtrls
 *   flush1: k_flush_scan, k_flush_serial, k_flush_misc
 *   k_begin_fused               doGravity + flush2 + integrate (fast path, no pending deltas)
 *     or k_gravity, k_flush2_apply, k_integrate
 *   every `rebuild_every` ticks: k_bounds, k_morton, radix sort, k_lbvh   (LBVH build)
 *   loop: [k_refit, k_pairs when the pair cache is stale] k_narrow -> (host reads "any proposal")
 *         -> k_decide, k_push (+ tie-cycle detection), k_push_serial, k_post
 *   flush1 again, (flush2 if deltas), k_hash on request
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <unordered_map>
#include <algorithm>

#include "bt_kernels.cuh"
#include "bt_small.h"
#include "bt_raycast.cuh"

static thread_local char g_err[512] = "";
static void set_err(const char *fmt, const char *a, const char *b, int line) {
	snprintf(g_err, sizeof g_err, fmt, a, b, line);
}
#define CK(call)                                                                      \
	do {                                                                              \
		cudaError_t _e = (call);                                                      \
		if (_e != cudaSuccess) {                                                      \
			set_err("%s failed: %s (bt_world.cu:%d)", #call, cudaGetErrorString(_e), __LINE__); \
			return BT_ERR_CUDA;                                                       \
		}                                                                             \
	} while (0)

enum {
	PK_PRE = 0, PK_FLUSH, PK_BEGIN, PK_BUILD_KEYS, PK_SORT, PK_LBVH, PK_REFIT, PK_PAIRS, PK_NARROW, PK_COLLIDE, PK_DECIDE, PK_PUSH,
	PK_POST, PK_HASH, PK_RAY_BUILD, PK_RAYCAST, PK_EXCHANGE, PK_TICK, PK_FUSED, PK_NUM
};
static const char *PK_NAMES[PK_NUM] = {"pre",    "flush1", "begin_integrate", "morton_keys", "radix_sort",
                                       "lbvh_build", "refit", "pairs", "narrow", "collide",       "decide",      "push",
                                       "post",   "hash",   "ray_tree_build",  "raycast", "host_exchange", "tick_handlers", "world_tick_small"};

typedef long long i64;

/* second LBVH (all live entities, static boxes) + live-rank maps for bt_raycast_* */
struct RayTree {
	u64 *keys, *keys_alt;
	i32 *leafEnt, *leafEnt_alt, *leafParent, *nodeParent;
	u32 *nodeCount;
	BvhNode *nodes;
	Box6 *nodeBox, *leafBox;
	i32 *rank2slot, *slot2rank;
	i32 *leafPos; /* entity -> position in this tree's sorted leaf array (ray ordering) */
	i32 *wroot;
	i32 n_leaves;
	i32 key_shift[3];
};

struct bt_world {
	int device;
	cudaStream_t stream;
	i32 cap_ents, cap_sliders, cap_wires, cap_worlds;
	DevWorld W;
	std::vector<void *> allocs;
	std::vector<size_t> alloc_bytes;
	u32 *h_flag; /* pinned + mapped, CT_WORDS words (+ 4 for a single world's hash sum and count): kernels write it (k_publish) */
	u32 *d_hflag; /* the same memory as the device sees it */
	u64 *d_hsum; u32 *d_hcnt;
	u32 *d_sorthist; int sorthist_len;
	bool any_rules; /* some world's gamerules contain a scan rule (crush / lava / bounce) */
	bool tree_valid, deltas_pending;
	std::unordered_map<u64, i32> uid_map; /* (world << 32 | uid) -> slot, built on the first bt_world_apply_requests */
	bool uid_map_valid;
	bool small_worlds; /* option "small_worlds": brute-force pair search when every world is small (default on) */
	bool pdl;       /* programmatic dependent launches in effect for the short kernels ... */
	bool pdl_long;  /* ... and for k_narrow / k_push (+ k_push_serial) */
	int narrow_threads;
	int narrow_blocks; /* grid of the grid-stride narrow phase (option "narrow_blocks") */
	i32 margin_opt; /* option "cache_margin": -1 auto */
	int pdl_opt;    /* option "pdl": 0 off, 1 on, -1 auto = on unless one big world (decided at upload) */
	bool ctr_fresh; /* h_flag mirrors the device counters (nothing that can change CT_DEAD_N ran since the last fetch) */
	bool any_ticks, any_rand; /* some entity has a tick handler / a TICK_RAND */
	int bad_tick;             /* first tick handler id without a device implementation, 0 = none */
	int device_ticks;         /* option: bt_world_tick runs the tick handlers itself when it can */
	BtReq *d_reqs; i32 *d_req_starts; i32 reqs_cap;
	int scan_flushes_left; /* of those, the ones that must look for pending requests themselves (k_flush_scan) */
	int full_flush_left; /* flush1 calls that must run in full (after an upload); afterwards CT_FLUSH_HINT decides */
	bool ctrls_fused;    /* flushCtrls of this step rides in k_begin_fused */
	i32 ticks_since_build;
	i64 rebuild_every;
	i64 max_iterations;
	i64 pair_cache; /* 1: cached candidate pairs (k_pairs + k_narrow); 0: traverse every iteration (k_collide) */
	i64 pair_rebuilds, pair_overflows;
	u32 seen_pair_builds, seen_pair_patches, pair_builds_at_tree; /* the last one counts builds + patches */
	int predicted_iters; /* iterations (incl. the final empty one) the last tick needed: run-ahead depth */
	i64 run_ahead;       /* option: 1 = enqueue the predicted iterations without waiting for the host */
	bt_tick_stats stats;
	bool profiling;
	cudaEvent_t ev0, ev1;
	i64 prof_launches[PK_NUM];
	double prof_ms[PK_NUM];
	std::vector<i32> up_root_order; /* per slot, from the last upload */
	/* raycast tree (all live entities, static boxes) */
	RayTree ray;
	bool ray_valid;
	i32 *d_stage_a, *d_stage_b; /* 3*cap ints each, staging for the per-tick host exchange */
	/* the pipelined exchange (bt_world_apply_host_deltas_async / bt_world_read_state_async): copy streams, double buffers */
	cudaStream_t s_h2d, s_d2h;
	i32 *px_in[2][2], *px_out[2][2];                  /* [buffer][centre / velocity] */
	cudaEvent_t px_arrived[2], px_added[2], px_packed[2], px_read[2];
	int px_in_k, px_out_k, px_commit_k;
	bool px_ready, px_sent[2]; /* sent[k]: buffer k was sent and not yet committed (FIFO: commit_k is the older one) */
	i32 px_sent_n[2];
	bool ranks_valid;           /* ray.slot2rank reflects the current live set */
	/* incremental host binding */
	i32 root_phase_ctr;         /* last value handed to a launch / the host as W.root_phase */
	i32 sliders_used, wires_used; /* high-water marks of the slider / wire arenas */
	bt_poke *d_pokes; i32 pokes_cap;
	/* ray ordering scratch (grows with the largest batch seen) */
	u64 *d_raykeys, *d_raykeys_alt; i32 *d_rayidx, *d_rayidx_alt; u32 *d_rayhist; i32 rays_cap;
	i64 ray_blocks;  /* its grid (option "ray_blocks") */
	i64 ray_dynamic; /* option "ray_dynamic": warps pull rays from a shared counter as their lanes fall idle (k_raycast_dyn) */
	i64 sort_rays; /* option "sort_rays": order the rays of a batch by the leaf position of their origin (default 0:
	                * measured on 4 M rays over the 1 M-entity world, 6.87 ms either way -- the sort itself costs 0.08 ms,
	                * the traversal does not get faster: it is bound by lanes idling behind the longest ray of their warp
	                * (9.2 of 32 threads active per instruction), not by where the rays start) */
	bool poked;                 /* the host wrote into the world since the counters were last fetched */
	cudaError_t launch_err;     /* first failed kernel launch since the last check (reported at the next sync point) */
	i64 small_warps;            /* option "small_warps": 0 = pick the k_tick_small variant by the number of worlds, 16 / 32 = force it */
	i64 fused_small;            /* option "fused_small": whole ticks of small worlds in one block per world (k_tick_small) */
	i32 small_biggest;          /* slots of the largest world of the last upload */
	bool fused_dirty;           /* fused ticks ran since the host last looked: counters, live count, cached pairs are stale */
	u64 seen_eticks, seen_iters; /* SK_ETICKS / SK_ITERS already added to `stats` (the device sums keep running) */
	bool wq_host_pending;       /* the host queued uWire / uUnwire requests in the world-wide pending-wire pool (bt_world_poke): the
	                             * next flush1 must be the generic one, k_tick_small keeps a pool per world in shared memory */
	DevArrays arrays;
};
template <typename T> static int dalloc(bt_world *w, T **p, size_t count) {
	void *q = NULL;
	cudaError_t e = cudaMalloc(&q, sizeof(T) * (count ? count : 1));
	if (e != cudaSuccess) {
		set_err("%s failed: %s (bt_world.cu:%d)", "cudaMalloc", cudaGetErrorString(e), __LINE__);
		return BT_ERR_CUDA;
	}
	cudaMemset(q, 0, sizeof(T) * (count ? count : 1));
	w->allocs.push_back(q);
	w->alloc_bytes.push_back(sizeof(T) * (count ? count : 1));
	*p = (T *)q;
	return 0;
}

/* Inflation of the boxes candidate pairs are cached with (worlds with a tree).  Measured on the 1 M-cube worlds, ms per
 * tick settled / first 12 ticks of the push-all world: 32: 0.584 / 1.32, 48: 0.583 / 1.25, 56: 0.585 / -, 64: - / 1.15
 * (but +14 % settled in round 1: neighbouring columns start to pair up), 96: - / 1.57.  Results never depend on it. */
#define BT_TREE_MARGIN 48

/* "pdl" auto: measured on single box-stack worlds, dependent launches gain 34 % at 20 k entities, 15 % at
 * 200 k, 8 % at 400 k, 2 % at 700 k and lose 3 % at 1 M */
#define BT_PDL_MAX_ENTS (3 << 18)
#define BT_PDL_BIG_WORLD_MODE 1 /* beyond that only the short kernels: +0.1-0.9 % at 1 M, all kernels would lose 3 % */

/* every kernel goes through here; pdl: programmatic dependent launch (see PDL_ENTRY in bt_kernels.cuh) */
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args &&...args) {
	if (!pdl) { kern<<<grid, block, smem, s>>>(KArgs(args)...); return cudaPeekAtLastError(); }
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid);
	cfg.blockDim = dim3((unsigned)block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
	cfg.attrs = at;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
/* a launch that failed (bad grid, sticky fault of an earlier kernel) is remembered with its own name and
 * line and turns the next sync point (fetch_ctr, the end of a step) into BT_ERR_CUDA */
static void note_launch_error(struct bt_world *w, cudaError_t e, const char *kern, int line);

/* the two long kernels (classes narrow and push; k_push_serial rides along) follow pdl_long */
#define pdl_for(w, cls) (((cls) == PK_NARROW || (cls) == PK_PUSH) ? (w)->pdl_long : (w)->pdl)

/* launch helper: counts launches and, while profiling, brackets each one with events */
#define LAUNCH(w, cls, kern, grid, block, ...) LAUNCH_S(w, cls, kern, grid, block, 0, __VA_ARGS__)
/* ... with `smem` bytes of dynamic shared memory */
#define LAUNCH_S(w, cls, kern, grid, block, smem, ...)                                 \
	do {                                                                               \
		if ((grid) > 0) {                                                              \
			if ((w)->profiling) cudaEventRecord((w)->ev0, (w)->stream);                \
			cudaError_t _le = launch_k(pdl_for((w), cls), kern, (grid), (block), (size_t)(smem), (w)->stream, __VA_ARGS__); \
			if (_le != cudaSuccess) note_launch_error((w), _le, #kern, __LINE__);      \
			(w)->stats.kernel_launches++;                                              \
			(w)->prof_launches[cls]++;                                                 \
			if ((w)->profiling) {                                                      \
				cudaEventRecord((w)->ev1, (w)->stream);                                \
				cudaEventSynchronize((w)->ev1);                                        \
				float ms = 0;                                                          \
				cudaEventElapsedTime(&ms, (w)->ev0, (w)->ev1);                         \
				(w)->prof_ms[cls] += ms;                                               \
			}                                                                          \
		}                                                                              \
	} while (0)

static inline int nblk(i64 n, int b = BT_BLOCK) { return (int)((n + b - 1) / b); }
/* fixed one-wave grid for the grid-stride kernels: 148 SMs x 8 blocks of 256 threads */
static inline int gsblk(i64 n) { return (int)std::min<i64>((n + GS_BLOCK - 1) / GS_BLOCK, 148 * 8); }
/* the two full-N scans of every iteration (decide, post): grid-stride over 148 SMs x 16 blocks of 128 threads.  One
 * thread per entity (7 863 blocks at 1 M) spends its time in block scheduling: measured 75 -> 63 us (decide) and
 * 74 -> 57 us (post) per tick on the 1 M-cube world, state hash unchanged (profiles/r2_ab_experiments.md) */
static inline int scanblk(i64 n) { return (int)std::min<i64>((n + BT_BLOCK - 1) / BT_BLOCK, 148 * 16); }
/* one full wave of k_pairs: SM count x resident blocks per SM (occupancy API, cached) */
static int pairs_wave() {
	static int wave = 0;
	if (!wave) {
		int dev = 0, sms = 148, per = 8;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_pairs, BT_BLOCK, 0) != cudaSuccess || per < 1) per = 8;
		wave = sms * per;
	}
	return wave;
}

static void set_pdl(struct bt_world *w, int n_worlds, long long n_ents);
static void note_launch_error(bt_world *w, cudaError_t e, const char *kern, int line) {
	if (w->launch_err != cudaSuccess) return;
	w->launch_err = e;
	snprintf(g_err, sizeof g_err, "launch of %s failed: %s (bt_world.cu:%d)", kern, cudaGetErrorString(e), line);
}
#define CHECK_LAUNCHES(w) do { if ((w)->launch_err != cudaSuccess) { (w)->launch_err = cudaSuccess; cudaGetLastError(); return BT_ERR_CUDA; } } while (0)
extern "C" {

const char *bt_version(void) { return "bittoss_b200 0.1 (sm_100a)"; }
const char *bt_last_error(void) { return g_err; }

int bt_device_count(void) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		set_err("%s: %s (bt_world.cu:%d)", "no usable CUDA device", cudaGetErrorString(e), __LINE__);
		return BT_ERR_NO_DEVICE;
	}
	return n;
}

void bt_world_destroy(bt_world *w) {
	if (!w) return;
	cudaSetDevice(w->device);
	for (void *p : w->allocs) cudaFree(p);
	if (w->h_flag) cudaFreeHost(w->h_flag);
	if (w->d_reqs) cudaFree(w->d_reqs);
	if (w->d_req_starts) cudaFree(w->d_req_starts);
	if (w->d_pokes) cudaFree(w->d_pokes);
	if (w->d_raykeys) cudaFree(w->d_raykeys);
	if (w->d_raykeys_alt) cudaFree(w->d_raykeys_alt);
	if (w->d_rayidx) cudaFree(w->d_rayidx);
	if (w->d_rayidx_alt) cudaFree(w->d_rayidx_alt);
	if (w->d_rayhist) cudaFree(w->d_rayhist);
	if (w->ev0) cudaEventDestroy(w->ev0);
	if (w->ev1) cudaEventDestroy(w->ev1);
	if (w->px_ready) {
		for (int k = 0; k < 2; k++) {
			cudaEventDestroy(w->px_arrived[k]); cudaEventDestroy(w->px_added[k]); cudaEventDestroy(w->px_packed[k]); cudaEventDestroy(w->px_read[k]);
			for (int j = 0; j < 2; j++) { cudaFree(w->px_in[k][j]); cudaFree(w->px_out[k][j]); }
		}
		cudaStreamDestroy(w->s_h2d); cudaStreamDestroy(w->s_d2h);
	}
	if (w->stream) cudaStreamDestroy(w->stream);
	delete w;
}

int bt_world_create(bt_world **out, int device, int32_t max_ents, int32_t max_sliders, int32_t max_wires,
                    int32_t max_worlds) {
	if (!out || max_ents < 0 || max_sliders < 0 || max_wires < 0 || max_worlds < 1) return BT_ERR_ARG;
	if (max_ents >= BT_MAX_ENTS || max_worlds > 65535) return BT_ERR_CAPACITY;
	int nd = bt_device_count();
	if (nd < 0) return nd;
	if (device < 0 || device >= nd) return BT_ERR_ARG;
	CK(cudaSetDevice(device));
	bt_world *w = new bt_world();
	memset(&w->W, 0, sizeof(w->W));
	memset(&w->stats, 0, sizeof(w->stats));
	memset(w->prof_launches, 0, sizeof(w->prof_launches));
	memset(w->prof_ms, 0, sizeof(w->prof_ms));
	memset(&w->ray, 0, sizeof(w->ray));
	w->device = device;
	w->cap_ents = max_ents; w->cap_sliders = max_sliders; w->cap_wires = max_wires; w->cap_worlds = max_worlds;
	w->tree_valid = false; w->deltas_pending = true; w->ticks_since_build = 0; w->full_flush_left = 2; w->scan_flushes_left = 2; w->ctrls_fused = false;
	w->rebuild_every = 16; w->max_iterations = 100000; w->pair_cache = 1; w->pair_rebuilds = 0; w->pair_overflows = 0; w->seen_pair_builds = 0; w->seen_pair_patches = 0; w->pair_builds_at_tree = 0; w->profiling = false; w->ray_valid = false; w->uid_map_valid = false; w->ctr_fresh = false; w->margin_opt = -1; w->narrow_blocks = 148 * 8; w->narrow_threads = 256; w->pdl_opt = getenv("BT_NO_PDL") ? 0 : -1; w->pdl = false; w->pdl_long = false; /* decided at upload, together with W.pdl */ w->small_worlds = true; w->d_reqs = NULL; w->d_req_starts = NULL; w->reqs_cap = 0; w->any_ticks = w->any_rand = false; w->bad_tick = 0; w->device_ticks = 1;
	w->h_flag = NULL; w->ev0 = w->ev1 = NULL; w->stream = NULL;
	w->px_ready = false; w->s_h2d = w->s_d2h = NULL; w->px_in_k = w->px_out_k = w->px_commit_k = 0; w->px_sent[0] = w->px_sent[1] = false;
	w->root_phase_ctr = 0; w->sliders_used = 0; w->wires_used = 0; w->d_pokes = NULL; w->pokes_cap = 0; w->poked = false;
	w->launch_err = cudaSuccess; w->fused_small = 1; w->small_warps = getenv("BT_SMALL_WARPS") ? atoi(getenv("BT_SMALL_WARPS")) : 0; w->fused_dirty = false; w->seen_eticks = w->seen_iters = 0; w->wq_host_pending = false; w->small_biggest = 0;
	w->d_raykeys = w->d_raykeys_alt = NULL; w->d_rayidx = w->d_rayidx_alt = NULL; w->d_rayhist = NULL; w->rays_cap = 0; w->sort_rays = 0; w->ray_dynamic = 1; w->ray_blocks = 148 * 12; /* measured on 4 M rays: 4.28 ms at 148 x 8 blocks, 4.06 at x 12 and beyond (13 fit an SM) */
	*out = w;
#define TRY(x) do { int _r = (x); if (_r) { bt_world_destroy(w); *out = NULL; return _r; } } while (0)
#define TRYCU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_err("%s failed: %s (bt_world.cu:%d)", #x, cudaGetErrorString(_e), __LINE__); bt_world_destroy(w); *out = NULL; return BT_ERR_CUDA; } } while (0)
	TRYCU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
	TRYCU(cudaEventCreate(&w->ev0));
	TRYCU(cudaEventCreate(&w->ev1));
	TRYCU(cudaHostAlloc((void **)&w->h_flag, sizeof(u32) * (CT_WORDS + 4), cudaHostAllocMapped));
	TRYCU(cudaHostGetDevicePointer((void **)&w->d_hflag, w->h_flag, 0));
	DevWorld &W = w->W;
	size_t n = (size_t)max_ents, nw = (size_t)max_worlds;
	TRY(dalloc(w, &W.wbegin, nw)); TRY(dalloc(w, &W.wend, nw)); TRY(dalloc(w, &W.flip, nw));
	TRY(dalloc(w, &W.gamerules, nw)); TRY(dalloc(w, &W.rnd, nw)); TRY(dalloc(w, &W.wflush, nw)); TRY(dalloc(w, &W.wunsup, nw)); TRY(dalloc(w, &W.wties, nw));
	TRY(dalloc(w, &W.world, n)); TRY(dalloc(w, &W.uid, n));
	TRY(dalloc(w, &W.center, n)); TRY(dalloc(w, &W.old, n)); TRY(dalloc(w, &W.vel, n));
	TRY(dalloc(w, &W.radius, n)); TRY(dalloc(w, &W.d_center, n)); TRY(dalloc(w, &W.d_vel, n));
	TRY(dalloc(w, &W.type, n)); TRY(dalloc(w, &W.collide, n)); TRY(dalloc(w, &W.ntype, n));
	TRY(dalloc(w, &W.ncollide, n));
	TRY(dalloc(w, &W.friction, n)); TRY(dalloc(w, &W.holder, n)); TRY(dalloc(w, &W.holdFlags, n));
	TRY(dalloc(w, &W.holdRoot, n)); TRY(dalloc(w, &W.holdee, n)); TRY(dalloc(w, &W.LLn, n));
	TRY(dalloc(w, &W.LLp, n)); TRY(dalloc(w, &W.newHolder, n)); TRY(dalloc(w, &W.newHoldFlags, n));
	TRY(dalloc(w, &W.hphys, n)); TRY(dalloc(w, &W.hpair, n)); TRY(dalloc(w, &W.htick, n));
	TRY(dalloc(w, &W.btn_v, n)); TRY(dalloc(w, &W.btn_v2, n));
	TRY(dalloc(w, &W.newDrop, n)); TRY(dalloc(w, &W.dead_max0, n)); TRY(dalloc(w, &W.dead_max1, n));
	TRY(dalloc(w, &W.fullFlush, n)); TRY(dalloc(w, &W.dead, n)); TRY(dalloc(w, &W.inList, n));
	TRY(dalloc(w, &W.needsCollision, n)); TRY(dalloc(w, &W.hasBox, n)); TRY(dalloc(w, &W.ready, n));
	TRY(dalloc(w, &W.forced, n)); TRY(dalloc(w, &W.forcedHoldees, n)); TRY(dalloc(w, &W.pend, n));
	TRY(dalloc(w, &W.slider_off, n + 1)); TRY(dalloc(w, &W.sv, (size_t)max_sliders));
	TRY(dalloc(w, &W.smin, (size_t)max_sliders)); TRY(dalloc(w, &W.smax, (size_t)max_sliders));
	TRY(dalloc(w, &W.wire_off, n + 1)); TRY(dalloc(w, &W.wire_cnt, n)); TRY(dalloc(w, &W.wire_to, (size_t)max_wires + n)); /* + one spare slot per (seat) entity at worst */
	TRY(dalloc(w, &W.wireAdd, n)); TRY(dalloc(w, &W.wireRm, n)); TRY(dalloc(w, &W.wire_cap, n));
	TRY(dalloc(w, &W.rootKey, n)); TRY(dalloc(w, &W.wrootseq, nw));
	W.wq_cap = (i32)std::min<size_t>((size_t)1 << 24, 2 * n + 1024);
	TRY(dalloc(w, &W.wq_next, (size_t)W.wq_cap)); TRY(dalloc(w, &W.wq_to, (size_t)W.wq_cap));
	TRY(dalloc(w, &W.keys, n)); TRY(dalloc(w, &W.keys_alt, n));
	TRY(dalloc(w, &W.leafEnt, n)); TRY(dalloc(w, &W.leafEnt_alt, n));
	TRY(dalloc(w, &W.leafParent, n)); TRY(dalloc(w, &W.nodeParent, n)); TRY(dalloc(w, &W.nodeCount, n));
	TRY(dalloc(w, &W.nodes, n)); TRY(dalloc(w, &W.nodeBox, n)); TRY(dalloc(w, &W.leafBox, n));
	TRY(dalloc(w, &W.nodeHi, n)); TRY(dalloc(w, &W.ibox, n)); TRY(dalloc(w, &W.wroot, nw + 1)); TRY(dalloc(w, &w->ray.wroot, nw + 1));
	TRY(dalloc(w, &W.iboxOld, n)); TRY(dalloc(w, &W.leafOf, n)); TRY(dalloc(w, &W.dirtyList, n)); TRY(dalloc(w, &W.dirtyStamp, n));
	W.patch = 1; W.patch_shift = 4;
	W.pair_cap = (i32)std::min<size_t>((size_t)1 << 30, 12 * n + 4096);
	W.margin = BT_TREE_MARGIN;
	TRY(dalloc(w, &W.pairs, (size_t)W.pair_cap));
	W.serial_cap = 1 << 16; W.dead_cap = max_ents > 0 ? max_ents : 1; W.event_cap = 1 << 16;
	TRY(dalloc(w, &W.events, (size_t)W.event_cap));
	TRY(dalloc(w, &W.serialList, (size_t)W.serial_cap)); TRY(dalloc(w, &W.deadList, (size_t)W.dead_cap));
	TRY(dalloc(w, &W.hash, nw)); TRY(dalloc(w, &W.ctr, (size_t)CT_WORDS)); TRY(dalloc(w, &W.stat, (size_t)BT_STAT_ROWS * SK_NUM));
	TRY(dalloc(w, &w->d_hsum, nw)); TRY(dalloc(w, &w->d_hcnt, nw));
	w->sorthist_len = 256 * (int)((n + SORT_TILE - 1) / SORT_TILE + 1);
	TRY(dalloc(w, &w->d_sorthist, (size_t)w->sorthist_len));
	TRY(dalloc(w, &w->ray.keys, n)); TRY(dalloc(w, &w->ray.keys_alt, n));
	TRY(dalloc(w, &w->ray.leafEnt, n)); TRY(dalloc(w, &w->ray.leafEnt_alt, n));
	TRY(dalloc(w, &w->ray.leafParent, n)); TRY(dalloc(w, &w->ray.nodeParent, n)); TRY(dalloc(w, &w->ray.nodeCount, n));
	TRY(dalloc(w, &w->ray.nodes, n)); TRY(dalloc(w, &w->ray.nodeBox, n)); TRY(dalloc(w, &w->ray.leafBox, n));
	TRY(dalloc(w, &w->ray.rank2slot, n)); TRY(dalloc(w, &w->ray.slot2rank, n)); TRY(dalloc(w, &w->ray.leafPos, n));
	TRY(dalloc(w, &w->d_stage_a, 3 * n)); TRY(dalloc(w, &w->d_stage_b, 3 * n));
	w->ranks_valid = false;
#undef TRY
#undef TRYCU
	{
		void **b = w->arrays.base;
		b[BT_A_CENTER] = W.center; b[BT_A_VEL] = W.vel; b[BT_A_OLD] = W.old; b[BT_A_DCENTER] = W.d_center; b[BT_A_DVEL] = W.d_vel;
		b[BT_A_RADIUS] = W.radius; b[BT_A_TYPE] = W.type; b[BT_A_COLLIDE] = W.collide; b[BT_A_NTYPE] = W.ntype;
		b[BT_A_NCOLLIDE] = W.ncollide; b[BT_A_FRICTION] = W.friction; b[BT_A_HOLDER] = W.holder; b[BT_A_HOLDFLAGS] = W.holdFlags;
		b[BT_A_HOLDROOT] = W.holdRoot; b[BT_A_HOLDEE] = W.holdee; b[BT_A_LLN] = W.LLn; b[BT_A_LLP] = W.LLp;
		b[BT_A_NEWHOLDER] = W.newHolder; b[BT_A_NEWHOLDFLAGS] = W.newHoldFlags; b[BT_A_UID] = W.uid; b[BT_A_WORLD] = W.world;
		b[BT_A_HPHYS] = W.hphys; b[BT_A_HPAIR] = W.hpair; b[BT_A_HTICK] = W.htick; b[BT_A_BTN_V] = W.btn_v; b[BT_A_BTN_V2] = W.btn_v2;
		b[BT_A_NEWDROP] = W.newDrop; b[BT_A_DEADMAX0] = W.dead_max0; b[BT_A_DEADMAX1] = W.dead_max1; b[BT_A_FULLFLUSH] = W.fullFlush;
		b[BT_A_DEAD] = W.dead; b[BT_A_INLIST] = W.inList; b[BT_A_ROOTKEY] = W.rootKey; b[BT_A_SLIDER_OFF] = W.slider_off;
		b[BT_A_WIRE_OFF] = W.wire_off; b[BT_A_WIRE_CAP] = W.wire_cap; b[BT_A_WIRE_CNT] = W.wire_cnt; b[BT_A_SV] = W.sv;
		b[BT_A_SMIN] = W.smin; b[BT_A_SMAX] = W.smax; b[BT_A_WIRE_TO] = W.wire_to; b[BT_A_W_FLIP] = W.flip;
		b[BT_A_W_GAMERULES] = W.gamerules; b[BT_A_W_RAND] = W.rnd; b[BT_A_DEADLIST] = W.deadList;
	}
	return BT_OK;
}

void *bt_world_stream(bt_world *w) { return w ? (void *)w->stream : NULL; }

int bt_world_set_option(bt_world *w, const char *name, int64_t value) {
	if (!w || !name) return BT_ERR_ARG;
	if (!strcmp(name, "rebuild_every")) { w->rebuild_every = value < 1 ? 1 : value; return BT_OK; }
	if (!strcmp(name, "max_iterations")) { w->max_iterations = value; return BT_OK; }
	if (!strcmp(name, "allow_unsupported")) { w->W.lenient = value != 0; return BT_OK; }
	if (!strcmp(name, "pair_cache")) { w->pair_cache = value != 0; return BT_OK; }
	if (!strcmp(name, "cache_horizon")) { w->W.horizon = value < 0 ? 0 : (value > 64 ? 64 : (i32)value); return BT_OK; }
	if (!strcmp(name, "narrow_threads")) { w->narrow_threads = value <= 64 ? 64 : (value <= 128 ? 128 : 256); return BT_OK; }
	if (!strcmp(name, "narrow_blocks")) { w->narrow_blocks = value < 1 ? 1 : (int)value; return BT_OK; }
	if (!strcmp(name, "pdl")) {
		w->pdl_opt = value < 0 ? -1 : (int)(value > 3 ? 3 : value);
		set_pdl(w, w->W.n_worlds, w->W.n);
		return BT_OK;
	}
	if (!strcmp(name, "small_worlds")) { /* takes effect at the next upload */
		w->small_worlds = value != 0;
		if (!w->small_worlds && w->W.small_mode) {
			w->W.small_mode = 0; w->tree_valid = false;
			if (w->margin_opt < 0) w->W.margin = BT_TREE_MARGIN;
		}
		return BT_OK;
	}
	if (!strcmp(name, "device_ticks")) { w->device_ticks = value != 0; return BT_OK; }
	if (!strcmp(name, "sort_rays")) { w->sort_rays = value != 0; return BT_OK; }
	if (!strcmp(name, "ray_dynamic")) { w->ray_dynamic = value != 0; return BT_OK; }
	if (!strcmp(name, "ray_blocks")) { w->ray_blocks = value < 1 ? 1 : value; return BT_OK; }
	if (!strcmp(name, "fused_small")) { w->fused_small = value != 0; return BT_OK; }
	if (!strcmp(name, "small_warps")) { w->small_warps = value; return BT_OK; }
	if (!strcmp(name, "pair_patch")) { w->W.patch = value != 0; return BT_OK; }
	if (!strcmp(name, "patch_shift")) { w->W.patch_shift = value < 0 ? 0 : (value > 20 ? 20 : (i32)value); return BT_OK; }
	if (!strcmp(name, "run_ahead")) { w->run_ahead = value != 0; return BT_OK; }
	if (!strcmp(name, "cache_margin")) { /* < 0: auto (32; 128 in small-world mode, where a rebuild is cheap but every one is global) */
		w->margin_opt = value < 0 ? -1 : (i32)value;
		w->W.margin = value < 0 ? (w->W.small_mode ? 128 : BT_TREE_MARGIN) : (i32)value;
		return BT_OK;
	}
	return BT_ERR_ARG;
}

/* ---------------------------------------------------------------- upload */

} /* extern "C" */
template <typename T> static int h2d(bt_world *w, T *dst, const std::vector<T> &src) {
	if (src.empty()) return 0;
	CK(cudaMemcpyAsync(dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, w->stream));
	return 0;
}

static bool supported(uint32_t mask, int id) { return id < 32 && ((mask >> id) & 1u); }
/* option "pdl": 0 off, 1 short kernels only, 3 all kernels, -1 auto (all up to BT_PDL_MAX_ENTS entities or when
 * worlds are batched, short kernels only beyond).  The launch attribute and the instructions in the
 * kernels always switch together. */
static void set_pdl(bt_world *w, int n_worlds, long long n_ents) {
	int mode = w->pdl_opt < 0 ? ((n_worlds > 1 || n_ents <= BT_PDL_MAX_ENTS) ? 3 : BT_PDL_BIG_WORLD_MODE) : w->pdl_opt;
	w->pdl = (mode & 1) != 0;
	w->pdl_long = (mode & 2) != 0;
	w->W.pdl = mode & 3;
}

extern "C" {

static int raise_cache_invalid(bt_world *w);
int bt_world_upload(bt_world *w, const bt_world_desc *d) {
	if (!w || !d || d->n_ents < 0 || d->n_worlds < 1) return BT_ERR_ARG;
	if (d->n_ents > w->cap_ents || d->n_sliders > w->cap_sliders || d->n_wires > w->cap_wires ||
	    d->n_worlds > w->cap_worlds)
		return BT_ERR_CAPACITY;
	CK(cudaSetDevice(w->device));
	const int n = d->n_ents, nw = d->n_worlds;
	/* handler IDs must have device implementations (no silent divergence) */
	for (int i = 0; i < n; i++) {
		const uint8_t *h = d->handlers + (size_t)BT_H_NUM * i;
		const char *bad = NULL;
		if (!supported(BT_SUPPORTED_WHOMOVES, h[BT_H_WHOMOVES])) bad = "whoMoves";
		else if (!supported(BT_SUPPORTED_PUSH, h[BT_H_PUSH])) bad = "push";
		else if (!supported(BT_SUPPORTED_PUSHED, h[BT_H_PUSHED])) bad = "pushed";
		else if (!supported(BT_SUPPORTED_CRUSH, h[BT_H_CRUSH])) bad = "crush";
		else
			for (int k = BT_H_ONFUMBLE; k <= BT_H_ONPICKEDUP; k++)
				if (!supported(BT_SUPPORTED_PAIR, h[k])) bad = "entPair";
		if (bad && !w->W.lenient) {
			snprintf(g_err, sizeof g_err, "entity %d: %s handler id has no device-side implementation", i, bad);
			return BT_ERR_UNSUPPORTED;
		}
	}
	std::vector<i32> world(n), wbegin(nw, n), wend(nw, 0);
#define BAD_DESC(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return BT_ERR_ARG; } while (0)
	for (int i = 0; i < n; i++) {
		int g = d->world_id ? d->world_id[i] : 0;
		if (g < 0 || g >= nw) BAD_DESC("entity %d: world id %d outside [0, %d)", i, g, nw);
		if (i > 0 && g < world[i - 1]) BAD_DESC("entity %d: world ids must be non-decreasing (the entities of one world are contiguous)", i);
		world[i] = g;
		wbegin[g] = std::min(wbegin[g], i);
		wend[g] = std::max(wend[g], i + 1);
	}
	/* every reference stays inside [-1, n) and inside its own world; the CSR offsets are monotonic and match the totals */
	if (d->slider_off[0] != 0 || d->wire_off[0] != 0 || d->slider_off[n] != d->n_sliders || d->wire_off[n] != d->n_wires)
		BAD_DESC("slider_off / wire_off do not start at 0 or do not end at n_sliders / n_wires");
	for (int i = 0; i < n; i++) {
		if (d->slider_off[i + 1] < d->slider_off[i] || d->wire_off[i + 1] < d->wire_off[i]) BAD_DESC("entity %d: CSR offsets decrease", i);
		const i32 refs[4] = {d->holder[i], d->holdee_head[i], d->holder[i] >= 0 ? d->holdee_next[i] : -1, d->new_holder[i]};
		for (int k = 0; k < 4; k++)
			if (refs[k] < -1 || refs[k] >= n || (refs[k] >= 0 && world[refs[k]] != world[i]))
				BAD_DESC("entity %d: %s %d is outside its world", i, k == 0 ? "holder" : k == 1 ? "holdee_head" : k == 2 ? "holdee_next" : "new_holder", refs[k]);
		for (i32 k = d->wire_off[i]; k < d->wire_off[i + 1]; k++)
			if (d->wire_to[k] < 0 || d->wire_to[k] >= n || world[d->wire_to[k]] != world[i]) BAD_DESC("entity %d: wire target %d is outside its world", i, d->wire_to[k]);
	}
#undef BAD_DESC
	for (int g = 0; g < nw; g++)
		if (wbegin[g] > wend[g]) wbegin[g] = wend[g] = 0;
	auto v3 = [&](const int32_t *src) {
		std::vector<int4> o(n);
		for (int i = 0; i < n; i++) o[i] = make_int4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0);
		return o;
	};
	bool deltas = false;
	for (int i = 0; i < 3 * n && !deltas; i++) deltas = d->d_center[i] || d->d_vel[i];
	/* a pending pickup with HOLD_FREEZE leaves a d_vel behind when prepPhysics applies it (ent.cpp:176-181):
	 * the flush2 at the start of doPhysics has to look (the fused begin kernel never reads d_vel) */
	for (int i = 0; i < n && !deltas; i++) deltas = d->new_holder[i] >= 0 && (d->new_hold_flags[i] & HOLD_FREEZE);
	std::vector<i32> holdRoot(n), LLn(n, -1), LLp(n, -1);
	for (int i = 0; i < n; i++) {
		i32 h = d->holder[i];
		if (h < -1 || h >= n) return BT_ERR_ARG;
		if (h >= 0) {
			LLn[i] = d->holdee_next[i];
			if (LLn[i] >= n) return BT_ERR_ARG;
			if (LLn[i] >= 0) LLp[LLn[i]] = i;
		}
	}
	for (int i = 0; i < n; i++) {
		i32 r = i;
		int guard = 0;
		while (d->holder[r] >= 0 && guard++ < n) r = d->holder[r];
		holdRoot[i] = r;
	}
	std::vector<u32> hphys(n), hpair(n), htick(n), bv(n), bv2(n), forced(n, 0);
	std::vector<u8> newDrop(n), dm0(n), dm1(n), ff(n), dead(n, 0), inList(n, 1), needsC(n, 0), hasBox(n), ready(n, 0);
	std::vector<int4> fh(n, make_int4(-1, -1, -1, 0));
	std::vector<u64> pend(n, 0);
	std::vector<i32> wire_cnt(n), wire_cap(n);
	std::vector<u64> rootKey(n, 0);
	bool any_ticks = false, any_rand = false;
	int bad_tick = 0;
	for (int i = 0; i < n; i++) {
		const uint8_t *h = d->handlers + (size_t)BT_H_NUM * i;
		hphys[i] = h[BT_H_WHOMOVES] | (h[BT_H_PUSH] << 8) | (h[BT_H_PUSHED] << 16) | ((u32)h[BT_H_CRUSH] << 24);
		hpair[i] = h[BT_H_ONFUMBLE] | (h[BT_H_ONFUMBLED] << 8) | (h[BT_H_ONPICKUP] << 16) | ((u32)h[BT_H_ONPICKEDUP] << 24);
		htick[i] = h[BT_H_TICK] | (h[BT_H_TICKHELD] << 8);
		for (int k = 0; k < 2; k++) {
			int id = h[k ? BT_H_TICKHELD : BT_H_TICK];
			if (id) any_ticks = true;
			if (id >= 64 || !((BT_SUPPORTED_TICK >> id) & 1ull)) { if (!bad_tick) bad_tick = id; }
			if (id < 64 && ((BT_SERIAL_TICK >> id) & 1ull)) any_rand = true; /* the per-world serial pass has work */
		}
		const uint8_t *b = d->btn + 8 * (size_t)i;
		bv[i] = b[0] | (b[1] << 8) | (b[2] << 16) | ((u32)b[3] << 24);
		bv2[i] = b[4] | (b[5] << 8) | (b[6] << 16) | ((u32)b[7] << 24);
		const uint8_t *m = d->misc + (size_t)BT_M_NUM * i;
		newDrop[i] = m[BT_M_NEWDROP]; dm0[i] = m[BT_M_DEADMAX0]; dm1[i] = m[BT_M_DEADMAX1]; ff[i] = m[BT_M_FULLFLUSH];
		hasBox[i] = (u8)(d->collide_mask[i] || (d->type_mask[i] & T_COLLIDABLE));
		wire_cnt[i] = d->wire_off[i + 1] - d->wire_off[i];
	}
	/* device wire CSR: the caller's lists plus one spare slot behind every seat */
	std::vector<i32> wire_off(n + 1), wire_to;
	{
		i32 o = 0;
		for (int i = 0; i < n; i++) {
			wire_off[i] = o;
			wire_cap[i] = wire_cnt[i] + (d->handlers[(size_t)BT_H_NUM * i + BT_H_ONPICKUP] == 1 ? 1 : 0);
			o += wire_cap[i];
			/* gs->rootEnts order (0 = head) as a key that sorts the head highest; phase 0 */
			if (d->holder[i] < 0) rootKey[i] = (u64)(n - (d->root_order[i] >= 0 ? d->root_order[i] : n - 1 - i));
		}
		wire_off[n] = o;
		w->wires_used = o;
		if (o > w->cap_wires + w->cap_ents) {
			snprintf(g_err, sizeof g_err, "world needs %d wire slots (incl. one spare per seat), capacity %d", o, w->cap_wires);
			return BT_ERR_CAPACITY;
		}
		wire_to.assign((size_t)o, -1);
		for (int i = 0; i < n; i++)
			for (i32 k = 0; k < wire_cnt[i]; k++) wire_to[wire_off[i] + k] = d->wire_to[d->wire_off[i] + k];
	}
	std::vector<i32> minus1(n, -1);
	DevWorld &W = w->W;
	W.n = n; W.n_worlds = nw; W.n_leaves = 0; W.iter_id = 0;
#define UP(dst, vec) do { int _r = h2d(w, dst, vec); if (_r) return _r; } while (0)
#define UPRAW(dst, src, cnt, T) do { if ((cnt) > 0) CK(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)(cnt), cudaMemcpyHostToDevice, w->stream)); } while (0)
	UP(W.world, world); UP(W.wbegin, wbegin); UP(W.wend, wend);
	UPRAW(W.flip, d->w_flip_flop, nw, i32); UPRAW(W.gamerules, d->w_gamerules, nw, i32); UPRAW(W.rnd, d->w_rand, nw, i32);
	UPRAW(W.uid, d->uid, n, i32);
	{
		std::vector<int4> a = v3(d->center), b = v3(d->old), c = v3(d->vel), r = v3(d->radius), dc = v3(d->d_center),
		                  dv = v3(d->d_vel);
		UP(W.center, a); UP(W.old, b); UP(W.vel, c); UP(W.radius, r); UP(W.d_center, dc); UP(W.d_vel, dv);
		CK(cudaStreamSynchronize(w->stream)); /* the staging vectors die here */
	}
	UPRAW(W.type, d->type_mask, n, u32); UPRAW(W.collide, d->collide_mask, n, u32);
	UPRAW(W.ntype, d->new_type_mask, n, u32); UPRAW(W.ncollide, d->new_collide_mask, n, u32);
	UPRAW(W.friction, d->friction, n, i32); UPRAW(W.holder, d->holder, n, i32);
	UPRAW(W.holdFlags, d->hold_flags, n, i32); UPRAW(W.holdee, d->holdee_head, n, i32);
	UPRAW(W.newHolder, d->new_holder, n, i32); UPRAW(W.newHoldFlags, d->new_hold_flags, n, i32);
	UP(W.holdRoot, holdRoot); UP(W.LLn, LLn); UP(W.LLp, LLp);
	UP(W.hphys, hphys); UP(W.hpair, hpair); UP(W.htick, htick); UP(W.btn_v, bv); UP(W.btn_v2, bv2);
	UP(W.newDrop, newDrop); UP(W.dead_max0, dm0); UP(W.dead_max1, dm1); UP(W.fullFlush, ff); UP(W.dead, dead);
	UP(W.inList, inList); UP(W.needsCollision, needsC); UP(W.hasBox, hasBox); UP(W.ready, ready);
	UP(W.forced, forced); UP(W.forcedHoldees, fh); UP(W.pend, pend);
	UPRAW(W.slider_off, d->slider_off, n + 1, i32); UPRAW(W.sv, d->slider_v, d->n_sliders, i32);
	UPRAW(W.smin, d->slider_min, d->n_sliders, i32); UPRAW(W.smax, d->slider_max, d->n_sliders, i32);
	UP(W.wire_off, wire_off); UP(W.wire_to, wire_to); UP(W.wire_cap, wire_cap); UP(W.rootKey, rootKey);
	UP(W.wire_cnt, wire_cnt); UP(W.wireAdd, minus1); UP(W.wireRm, minus1);
	w->sliders_used = d->n_sliders; w->root_phase_ctr = 0; w->poked = false; w->fused_dirty = false;
	CK(cudaMemsetAsync(W.ctr, 0, sizeof(u32) * CT_WORDS, w->stream));
	CK(cudaMemsetAsync(W.stat, 0, sizeof(u64) * BT_STAT_ROWS * SK_NUM, w->stream));
	w->seen_eticks = w->seen_iters = 0;
	{ int rr = raise_cache_invalid(w); if (rr) return rr; }
	CK(cudaMemsetAsync(W.dirtyStamp, 0, sizeof(u32) * (size_t)w->cap_ents, w->stream));
	w->seen_pair_builds = 0; w->seen_pair_patches = 0; w->pair_builds_at_tree = 0;
	CK(cudaMemsetAsync(W.nodeCount, 0, sizeof(u32) * (size_t)w->cap_ents, w->stream));
	CK(cudaMemsetAsync(W.wflush, 0, (size_t)nw, w->stream));
	CK(cudaStreamSynchronize(w->stream));
#undef UP
#undef UPRAW
	w->up_root_order.assign(d->root_order, d->root_order + n);
	w->tree_valid = false; w->ray_valid = false; w->ranks_valid = false;
	w->any_rules = false;
	for (int g = 0; g < nw; g++) w->any_rules = w->any_rules || (d->w_gamerules[g] & (1 | 2 | 16 | 32));
	w->deltas_pending = deltas;
	w->full_flush_left = 2; /* both dead_max slots of the upload get looked at */
	w->scan_flushes_left = 2;
	w->uid_map_valid = false; w->ctr_fresh = false;
	{
		i32 biggest = 0;
		for (int g = 0; g < nw; g++) biggest = std::max(biggest, wend[g] - wbegin[g]);
		W.small_mode = (w->small_worlds && biggest <= BT_SMALL_WORLD) ? 1 : 0;
		w->small_biggest = biggest;
		if (w->margin_opt < 0) W.margin = W.small_mode ? 128 : BT_TREE_MARGIN;
		/* Dependent launches pay when a tick is many short kernels (batched or small worlds: +5-7 %).
		 * One big world spends its tick in two long kernels; there the resident waiting blocks cost
		 * 0.1-0.4 % instead (measured A/B on the 1 M-entity world), so "auto" leaves them off. */
		set_pdl(w, nw, n);
	}
	w->any_ticks = any_ticks; w->any_rand = any_rand; w->bad_tick = bad_tick;
	w->ticks_since_build = 0;
	w->stats.live_ents = n;
	return BT_OK;
}

/* ---------------------------------------------------------------- tick */

/* stable LSD radix sort of (key, payload) pairs, 8 bits per pass; the sorted data ends up in *keys / *vals (the
 * pointers are swapped with their alternates after every pass).  hist: 256 x tiles counters. */
static int radix_sort_buf(bt_world *w, int cls, u64 **keys, i32 **vals, u64 **keys_alt, i32 **vals_alt, i32 n, int key_bits, u32 *hist) {
	const int nb = (n + SORT_TILE - 1) / SORT_TILE;
	for (int shift = 0; shift < key_bits; shift += 8) {
		LAUNCH(w, cls, k_sort_hist, nb, SORT_THREADS, (const u64 *)*keys, n, shift, hist, nb);
		LAUNCH(w, cls, k_sort_scan, 8, 1024, hist, nb); /* 256 warps: one per digit */
		LAUNCH(w, cls, k_sort_scatter, nb, SORT_THREADS, (const u64 *)*keys, (const i32 *)*vals, *keys_alt, *vals_alt, n, shift, (const u32 *)hist, nb);
		std::swap(*keys, *keys_alt);
		std::swap(*vals, *vals_alt);
	}
	return 0;
}
static int radix_sort(bt_world *w, int key_bits) {
	DevWorld &W = w->W;
	return radix_sort_buf(w, PK_SORT, &W.keys, &W.leafEnt, &W.keys_alt, &W.leafEnt_alt, W.n, key_bits, w->d_sorthist);
}

static int bits_for(u32 x) { int b = 0; while (x) { b++; x >>= 1; } return b; }


/* Morton sort + Karras hierarchy over every live boxed entity */
static int build_tree(bt_world *w) {
	DevWorld &W = w->W;
	const i32 n = W.n;
	i32 init[8] = {0, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN, 0};
	CK(cudaMemcpyAsync(&W.ctr[CT_N_LEAVES], init, sizeof(i32) * 7, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(w, PK_BUILD_KEYS, k_bounds, gsblk(n), GS_BLOCK, W);
	LAUNCH(w, PK_BUILD_KEYS, k_publish, 1, 64, (const u32 *)W.ctr, w->d_hflag, (int)CT_WORDS);
	CK(cudaStreamSynchronize(w->stream));
	W.n_leaves = (i32)w->h_flag[CT_N_LEAVES];
	const i32 *b = (const i32 *)&w->h_flag[CT_BOUNDS];
	for (int k = 0; k < 3; k++) {
		u32 ext = W.n_leaves ? (u32)(b[3 + k] - b[k]) : 0;
		int bits = bits_for(ext);
		W.key_shift[k] = bits > 9 ? bits - 9 : 0; /* 9 bits per axis */
	}
	W.world_bits = bits_for((u32)W.n_worlds); /* ids 0..n_worlds, the last one = "not a leaf" */
	LAUNCH(w, PK_BUILD_KEYS, k_morton, nblk(n), BT_BLOCK, W);
	int r = radix_sort(w, 30 + W.world_bits);
	if (r) return r;
	CK(cudaMemsetAsync(W.wroot, 0x80, sizeof(i32) * (size_t)(w->cap_worlds + 1), w->stream)); /* "no subtree" */
	if (W.n_leaves >= 2) LAUNCH(w, PK_LBVH, k_lbvh, nblk(W.n_leaves - 1), BT_BLOCK, W);
	/* new topology: node boxes are stale until the next refit, counters restart even */
	CK(cudaMemsetAsync(W.nodeCount, 0, sizeof(u32) * (size_t)n, w->stream));
	if (!W.ray_mode) { int rr = raise_cache_invalid(w); if (rr) return rr; }
	w->tree_valid = true;
	w->ticks_since_build = 0;
	w->pair_builds_at_tree = w->seen_pair_builds + w->seen_pair_patches + 1; /* + the pair search this build itself forces */
	w->stats.tree_builds++;
	return 0;
}

/* cond: only if CT_CACHE_INVALID is raised (decided on the device, no host round trip) */
static int refit(bt_world *w, bool cond) {
	DevWorld &W = w->W;
	if (W.n_leaves < 1) return 0;
	LAUNCH(w, PK_REFIT, k_refit, gsblk(W.n_leaves), GS_BLOCK, W, cond ? 1 : 0);
	w->stats.tree_refits++;
	return 0;
}

/* the next pair build must be a full one */
static int raise_cache_invalid(bt_world *w) {
	static const u32 one = 1;
	CK(cudaMemcpyAsync(&w->W.ctr[CT_CACHE_FULL], &one, sizeof(u32), cudaMemcpyHostToDevice, w->stream));
	CK(cudaMemcpyAsync(&w->W.ctr[CT_CACHE_INVALID], &one, sizeof(u32), cudaMemcpyHostToDevice, w->stream));
	return 0;
}

/* flush1 (ent.cpp:794-804).  `hint` = the host's latest copy of CT_FLUSH_HINT: when
 * nothing asked for a flush (no slider/mask request, no death, nothing uploaded) the three
 * full passes would find nothing, so only the death flip-flop is toggled. */
static int flush1(bt_world *w, bool hint) {
	DevWorld &W = w->W;
	if (w->full_flush_left <= 0 && !hint) {
		LAUNCH(w, PK_FLUSH, k_flip_only, nblk(W.n_worlds), BT_BLOCK, W);
		return 0;
	}
	const bool scan = w->scan_flushes_left > 0; /* pending requests an upload brought along */
	if (w->scan_flushes_left > 0) w->scan_flushes_left--;
	if (w->full_flush_left > 0) w->full_flush_left--;
	w->ctr_fresh = false;
	CK(cudaMemsetAsync(&W.ctr[CT_FLUSH_HINT], 0, sizeof(u32), w->stream));
	if (scan) LAUNCH(w, PK_FLUSH, k_flush_scan, nblk(W.n), BT_BLOCK, W);
	W.root_phase = ++w->root_phase_ctr; /* roots this flush makes go ahead of every older one in gs->rootEnts */
	LAUNCH(w, PK_FLUSH, k_flush_serial, nblk((i64)W.n_worlds * 32), BT_BLOCK, W); /* one warp per world */
	LAUNCH(w, PK_FLUSH, k_flush_misc, nblk(W.n), BT_BLOCK, W);
	CK(cudaMemsetAsync(&W.ctr[CT_WQ_N], 0, sizeof(u32), w->stream)); /* every pending-wire list has been consumed */
	w->wq_host_pending = false;
	return 0;
}

static int fetch_ctr(bt_world *w) {
	CHECK_LAUNCHES(w);
	LAUNCH(w, PK_EXCHANGE, k_publish, 1, 64, (const u32 *)w->W.ctr, w->d_hflag, (int)CT_WORDS); /* not a copy: see k_publish */
	CK(cudaStreamSynchronize(w->stream));
	w->poked = false;
	if (w->h_flag[CT_ERROR]) {
		int code = -(int)w->h_flag[CT_ERROR];
		snprintf(g_err, sizeof g_err, "device raised error %d (overflow of a work list or unsupported flush)", code);
		return code;
	}
	return 0;
}

/* doPhysics up to and including the resolution loop (ent.cpp:843-905) */
static int do_physics_iter(bt_world *w, bool gravity) {
	DevWorld &W = w->W;
	const i32 n = W.n;
	if (!w->deltas_pending) {
		LAUNCH(w, PK_BEGIN, k_begin_fused, nblk(n), BT_BLOCK, W, gravity ? 1 : 0, w->ctrls_fused ? 1 : 0);
	} else {
		if (gravity) LAUNCH(w, PK_BEGIN, k_gravity, nblk(n), BT_BLOCK, W);
		LAUNCH(w, PK_BEGIN, k_flush2_apply, nblk(n), BT_BLOCK, W);
		LAUNCH(w, PK_BEGIN, k_integrate, nblk(n), BT_BLOCK, W);
		w->deltas_pending = false;
	}
	/* leaf set changes (an entity gained a box) force a rebuild; deaths do not */
	/* ... and an old tree is only rebuilt if pair lists are being rebuilt at all (a world at
	 * rest keeps its cached pairs for as long as nobody leaves its inflated box) */
	/* small-world mode finds its pairs without a tree: none is built unless the fallback needs it */
	const bool need_tree = !(W.small_mode && w->pair_cache);
	if (need_tree && (!w->tree_valid || (w->ticks_since_build >= w->rebuild_every &&
	                       (!w->pair_cache || w->seen_pair_builds + w->seen_pair_patches > w->pair_builds_at_tree)))) {
		int r = build_tree(w);
		if (r) return r;
	}
	w->ticks_since_build++;
	W.iter_base = W.iter_id;
	i64 it = 0;
	bool use_cache = w->pair_cache != 0;
	/* 0: the host does not know whether the cache is valid (first iteration: integrate may have
	 * invalidated it) -> conditional kernels decide on the device; 1: known stale; 2: known good */
	int cache_state = 0;
	int iters_this_tick = 0;
	CK(cudaMemsetAsync(&W.ctr[CT_STOP], 0, sizeof(u32), w->stream));
	auto pair_build = [&](int cond) -> int {
		if (W.small_mode) { /* one thread resets the list (k_refit, cond), then one block per world */
			LAUNCH(w, PK_REFIT, k_refit, 1, GS_BLOCK, W, 1);
			LAUNCH(w, PK_PAIRS, k_pairs_small, W.n_worlds, BT_SMALL_WORLD, W);
			return 0;
		}
		int r = refit(w, cond != 0);
		if (r) return r;
		LAUNCH(w, PK_PAIRS, k_pairs, std::min(nblk(W.n_leaves), pairs_wave()), BT_BLOCK, W);
		return 0;
	};
	auto resolve = [&]() {
		/* speculative behind the narrow phase (see k_decide) */
		LAUNCH(w, PK_DECIDE, k_decide, scanblk(n), BT_BLOCK, W);
		W.root_phase = ++w->root_phase_ctr; /* both push lanes of one iteration share it (canonical push order inside) */
		LAUNCH(w, PK_PUSH, k_push, nblk(n), BT_BLOCK, W);
		LAUNCH(w, PK_PUSH, k_push_serial, 1, 32, W);
		LAUNCH(w, PK_POST, k_post, scanblk(n), BT_BLOCK, W);
	};
	while (1) {
		int r;
		W.iter_id++;
		const i32 first_id = W.iter_id;
		int enq = 1; /* iterations enqueued in this batch */
		if (use_cache) {
			/* refit + pair search run only if CT_CACHE_INVALID is up (upload, rebuilt tree, or an
			 * entity left the inflated box its pairs were found with); state 1: the flag is up (the
			 * host has just read it), the same conditional kernels run */
			if (cache_state != 2) { r = pair_build(1); if (r) return r; }
			/* pair-parallel in every iteration.  (Measured alternative: walking per-entity neighbour
			 * lists for the re-armed entities only -- 2.6x slower per launch at 1 M entities: one
			 * thread per scanner serialises its 2-5 neighbours' gather chains, one thread per pair
			 * keeps them all in flight.) */
			/* bit 1: the first doIteration of the tick, right behind the integrate kernel (every listed entity scans) */
			LAUNCH(w, PK_NARROW, k_narrow, w->narrow_blocks, w->narrow_threads, W, iters_this_tick == 0 ? 2 : 0);
			resolve();
			if (w->run_ahead) {
				/* run-ahead: as many further iterations as the previous tick needed, each of which
				 * checks on the device that its predecessor proposed something and that the pair
				 * cache is still good, and otherwise stands down together with everything behind it */
				int more = (iters_this_tick == 0 ? w->predicted_iters - 1 : 2);
				for (int k = 0; k < more; k++) {
					W.iter_id++;
					LAUNCH(w, PK_NARROW, k_narrow, w->narrow_blocks, w->narrow_threads, W, 1);
					resolve();
					enq++;
				}
			}
		} else {
			if (!w->tree_valid) { r = build_tree(w); if (r) return r; } /* small-world mode had none */
			r = refit(w, false);
			if (r) return r;
			LAUNCH(w, PK_COLLIDE, k_collide, nblk(W.n_leaves), BT_BLOCK, W);
			resolve();
		}
		r = fetch_ctr(w);
		if (r) return r;
		if (use_cache && w->h_flag[CT_PAIR_BUILDS] != w->seen_pair_builds) {
			w->stats.pair_builds += (i64)(w->h_flag[CT_PAIR_BUILDS] - w->seen_pair_builds);
			w->stats.pairs_cached += (i64)w->h_flag[CT_PAIR_N];
			w->seen_pair_builds = w->h_flag[CT_PAIR_BUILDS];
		}
		if (use_cache && w->h_flag[CT_PAIR_PATCHES] != w->seen_pair_patches) {
			w->stats.pair_patches += (i64)(w->h_flag[CT_PAIR_PATCHES] - w->seen_pair_patches);
			w->seen_pair_patches = w->h_flag[CT_PAIR_PATCHES];
		}
		if (use_cache && w->h_flag[CT_PAIR_OVERFLOW]) {
			/* more touching pairs than the buffer holds: every kernel of this batch stood down;
			 * redo it, and the rest of the tick, with per-iteration traversal */
			use_cache = false;
			w->pair_overflows++;
			w->stats.pair_overflows++;
			CK(cudaMemsetAsync(&W.ctr[CT_PAIR_OVERFLOW], 0, 2 * sizeof(u32), w->stream)); /* + CT_STOP */
			continue;
		}
		/* which of the enqueued iterations actually ran */
		const u32 stop = w->h_flag[CT_STOP];
		const i32 last_run = stop ? (i32)stop - 1 : first_id + enq - 1;
		const bool last_proposed = w->h_flag[ANY_SLOT(last_run)] == (u32)last_run;
		const int ran = last_run - first_id + 1;
		w->stats.iterations += ran - (last_proposed ? 0 : 1);
		iters_this_tick += ran;
		it += ran;
		if (!last_proposed) break; /* doIteration returned 0 (ent.cpp:866) */
		if (stop) CK(cudaMemsetAsync(&W.ctr[CT_STOP], 0, sizeof(u32), w->stream));
		cache_state = w->h_flag[CT_CACHE_INVALID] ? 1 : 2;
		if (it > w->max_iterations) {
			snprintf(g_err, sizeof g_err, "doPhysics did not converge in %lld iterations", (long long)w->max_iterations);
			return BT_ERR_OVERFLOW;
		}
	}
	w->predicted_iters = iters_this_tick > 0 ? iters_this_tick : 1;
	if (w->pair_cache && !use_cache) { /* overflowed: the partial pair list must not be reused */
		int rr = raise_cache_invalid(w);
		if (rr) return rr;
	}
	return 0;
}

/* invokeOnCrush (ent.cpp:907) lies between the two halves: crush handlers are host-side events (dead list).
 * This is the rest of doPhysics: flush1; flush2 (ent.cpp:909-910). */
static int do_physics_end(bt_world *w) {
	DevWorld &W = w->W;
	const i32 n = W.n;
	/* h_flag holds the counters fetched after the last k_narrow, i.e. after every push -- unless the host
	 * has written into the world since (crush handlers of a host binding): look again */
	if (w->poked) { int rr = fetch_ctr(w); if (rr) return rr; }
	W.flush_stage = 1; /* deaths flushed from here on are past invokeOnCrush (ent.cpp:907-909) */
	const bool full_flush = w->full_flush_left > 0 || w->h_flag[CT_FLUSH_HINT] != 0;
	int r = flush1(w, w->h_flag[CT_FLUSH_HINT] != 0);
	W.flush_stage = 0;
	if (r) return r;
	/* a flush1 that runs k_flush_serial can itself kill (dead_max) and so leave wires to clean
	 * up: keep the next flush1 full in that case.  A flush1 that only toggled the flip-flop
	 * changed no counter: the copy fetched after the last iteration is still exact, and the host
	 * round trip (the GPU idles through it) is skipped -- the steady state of a world at rest. */
	if (full_flush) {
		r = fetch_ctr(w);
		if (r) return r;
	}
	if (w->h_flag[CT_FLUSH_HINT] && w->full_flush_left < 1) w->full_flush_left = 1;
	if (w->h_flag[CT_ANY_DELTA]) {
		/* flush2 (ent.cpp:910): a pickup with HOLD_FREEZE left a d_vel behind (ent.cpp:176-181) */
		LAUNCH(w, PK_FLUSH, k_flush2_apply, nblk(n), BT_BLOCK, W);
		LAUNCH(w, PK_FLUSH, k_zero_deltas, nblk(n), BT_BLOCK, W);
		CK(cudaMemsetAsync(&W.ctr[CT_ANY_DELTA], 0, sizeof(u32), w->stream));
	}
	if (w->h_flag[CT_NEED_REBUILD]) {
		w->tree_valid = false; /* build_tree raises CT_CACHE_INVALID */
		if (W.small_mode) { int rr = raise_cache_invalid(w); if (rr) return rr; } /* no build_tree to do it */
		CK(cudaMemsetAsync(&W.ctr[CT_NEED_REBUILD], 0, sizeof(u32), w->stream));
	}
	w->ray_valid = false;
	w->ctr_fresh = true;
	return 0;
}

} /* extern "C" */

/* the statistics: BT_STAT_ROWS partial sums per counter (bt_types.h), added up here */
static int read_stats(bt_world *w, u64 *s) {
	std::vector<u64> rows((size_t)BT_STAT_ROWS * SK_NUM);
	CK(cudaMemcpyAsync(rows.data(), w->W.stat, sizeof(u64) * rows.size(), cudaMemcpyDeviceToHost, w->stream));
	CK(cudaStreamSynchronize(w->stream));
	for (int k = 0; k < SK_NUM; k++) s[k] = 0;
	for (int r = 0; r < BT_STAT_ROWS; r++)
		for (int k = 0; k < SK_NUM; k++) s[k] += rows[(size_t)r * SK_NUM + k];
	return 0;
}

/* After ticks of the one-block-per-world kernel the host has not looked at the device: the dead entities of those
 * ticks are still on the dead list, the live count and entity-ticks are in device counters, and the generic schedule's
 * cached pairs / flush hints know nothing of what happened.  Called before anything that needs them. */
static int fused_settle(bt_world *w, bool leaving) {
	if (!w->fused_dirty) return 0;
	DevWorld &W = w->W;
	int r = fetch_ctr(w);
	if (r) return r;
	u64 s[SK_NUM];
	r = read_stats(w, s);
	if (r) return r;
	w->stats.entity_ticks += (i64)(s[SK_ETICKS] - w->seen_eticks);
	const u64 di = s[SK_ITERS] - w->seen_iters;
	w->stats.iterations += (i64)((di + (u64)W.n_worlds / 2) / (u64)(W.n_worlds ? W.n_worlds : 1)); /* mean over the worlds */
	w->seen_eticks = s[SK_ETICKS]; w->seen_iters = s[SK_ITERS];
	if (w->h_flag[CT_DEAD_N]) {
		w->stats.live_ents -= (i64)w->h_flag[CT_DEAD_N];
		w->ranks_valid = false;
		CK(cudaMemsetAsync(&W.ctr[CT_DEAD_N], 0, sizeof(u32), w->stream));
		w->h_flag[CT_DEAD_N] = 0;
	}
	if (w->h_flag[CT_NEED_REBUILD]) CK(cudaMemsetAsync(&W.ctr[CT_NEED_REBUILD], 0, sizeof(u32), w->stream));
	if (leaving) { /* the generic schedule takes over: it must rebuild its pair list and look at every flush flag */
		r = raise_cache_invalid(w);
		if (r) return r;
		w->tree_valid = false;
		w->deltas_pending = true;
		if (w->full_flush_left < 2) w->full_flush_left = 2;
		if (w->scan_flushes_left < 2) w->scan_flushes_left = 2;
		w->fused_dirty = false;
		w->ctr_fresh = false;
	}
	return 0;
}

static int step_small_fused(bt_world *w, uint32_t phases) {
	DevWorld &W = w->W;
	if ((phases & BT_PHASE_TICKS) && w->bad_tick && !W.lenient) {
		snprintf(g_err, sizeof g_err, "BT_PHASE_TICKS: tick handler id %d has no device implementation", w->bad_tick);
		return BT_ERR_UNSUPPORTED;
	}
	if ((phases & BT_PHASE_TICKS) && !w->any_ticks) phases &= ~BT_PHASE_TICKS;
	if (w->root_phase_ctr > (1 << 27) - 2 * SW_PHASES) { snprintf(g_err, sizeof g_err, "root-order phases exhausted: download and upload the world again"); return BT_ERR_OVERFLOW; }
	W.root_phase = w->root_phase_ctr + 1;
	w->root_phase_ctr += SW_PHASES;
	const u32 kp = phases | (w->deltas_pending ? SWF_HOST_DELTAS : 0u);
	if (W.n_worlds > 0) { /* LAUNCH() for a kernel that lives in another translation unit (bt_small.cu) */
		if (w->profiling) cudaEventRecord(w->ev0, w->stream);
		cudaError_t le = bt_launch_tick_small(W, kp, w->small_biggest, (int)w->small_warps, w->pdl, w->stream);
		if (le != cudaSuccess) note_launch_error(w, le, "k_tick_small", __LINE__);
		w->stats.kernel_launches++;
		w->prof_launches[PK_FUSED]++;
		if (w->profiling) {
			cudaEventRecord(w->ev1, w->stream);
			cudaEventSynchronize(w->ev1);
			float ms = 0;
			cudaEventElapsedTime(&ms, w->ev0, w->ev1);
			w->prof_ms[PK_FUSED] += ms;
		}
	}
	w->fused_dirty = true;
	w->ctr_fresh = false; w->ray_valid = false;
	if (phases & BT_PHASE_PHYSICS) { w->deltas_pending = false; w->stats.ticks++; }
	else if (phases & BT_PHASE_GRAVITY) w->deltas_pending = true;
	if (phases & BT_PHASE_PREP) { w->full_flush_left = 0; w->scan_flushes_left = 0; }
	CHECK_LAUNCHES(w);
	return BT_OK;
}

extern "C" {

int bt_world_step(bt_world *w, uint32_t phases) {
	if (!w) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	DevWorld &W = w->W;
	const i32 n = W.n;
	/* small worlds: the whole step in one launch, one block per world, nothing comes back to the host */
	/* (doPhysics and clearDeads together or not at all: the dead of a fused tick are counted off when the host settles) */
	if (W.small_mode && w->fused_small && w->pair_cache && !w->wq_host_pending &&
	    !(phases & ~(BT_PHASE_ALL | BT_PHASE_WORLD_RULES | BT_PHASE_TICKS)) &&
	    !(phases & BT_PHASE_PHYSICS) == !(phases & BT_PHASE_FINISH))
		return step_small_fused(w, phases);
	{ int rr = fused_settle(w, true); if (rr) return rr; }
	if (phases & BT_PHASE_WORLD_RULES) {
		LAUNCH(w, PK_PRE, k_world_rules, nblk(n), BT_BLOCK, W);
		/* a death requested here is flushed by this step's prepPhysics: make sure it looks */
		if (w->full_flush_left < 1) w->full_flush_left = 1;
	}
	/* flushCtrls rides in k_begin_fused when this step also integrates on the fast path */
	if (phases & ~BT_PHASE_FINISH) w->ctr_fresh = false; /* until doPhysics refreshes it */
	if ((phases & BT_PHASE_TICKS) && w->bad_tick && !W.lenient) {
		snprintf(g_err, sizeof g_err, "BT_PHASE_TICKS: tick handler id %d has no device implementation", w->bad_tick);
		return BT_ERR_UNSUPPORTED;
	}
	w->ctrls_fused = (phases & BT_PHASE_CTRLS) && (phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_ITER)) && !w->deltas_pending && !(phases & BT_PHASE_TICKS);
	if ((phases & BT_PHASE_CTRLS) && !w->ctrls_fused) LAUNCH(w, PK_PRE, k_pre, nblk(n), BT_BLOCK, W, phases);
	if ((phases & BT_PHASE_TICKS) && w->any_ticks) {
		LAUNCH(w, PK_TICK, k_tick, nblk(n), BT_BLOCK, W);
		if (w->any_rand) LAUNCH(w, PK_TICK, k_tick_serial, nblk(W.n_worlds), BT_BLOCK, W);
		w->deltas_pending = true;                           /* door_tick's uVel */
		if (w->full_flush_left < 1) w->full_flush_left = 1; /* sliders, deaths, drops the handlers asked for */
	}
	if (phases & BT_PHASE_PREP) {
		int r = flush1(w, false);
		if (r) return r;
	}
	if ((phases & BT_PHASE_GRAVITY) && !(phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_ITER))) {
		LAUNCH(w, PK_BEGIN, k_gravity, nblk(n), BT_BLOCK, W);
		w->deltas_pending = true;
	}
	if (phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_ITER)) {
		int r = do_physics_iter(w, (phases & BT_PHASE_GRAVITY) != 0);
		if (r) return r;
	}
	if (phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_END)) {
		int r = do_physics_end(w);
		if (r) return r;
		w->stats.ticks++;
	}
	if (phases & BT_PHASE_FINISH) {
		/* clearDeads (ent.cpp:471-481): slots of dead entities stay retired */
		if (!w->ctr_fresh) { /* doPhysics of this same step leaves an exact host copy behind: no second round trip */
			int r = fetch_ctr(w);
			if (r) return r;
		}
		w->stats.live_ents -= (i64)w->h_flag[CT_DEAD_N];
		if (w->h_flag[CT_DEAD_N]) {
			w->ranks_valid = false;
			CK(cudaMemsetAsync(&W.ctr[CT_DEAD_N], 0, sizeof(u32), w->stream));
			w->h_flag[CT_DEAD_N] = 0;
		}
		w->stats.entity_ticks += w->stats.live_ents;
	}
	CHECK_LAUNCHES(w);
	return BT_OK;
}

/* one doWholeStep (main.cpp:1019-1097) as far as the device can run it */
static uint32_t tick_phases(const bt_world *w) {
	return BT_PHASE_ALL | (w->any_rules ? BT_PHASE_WORLD_RULES : 0u) |
	       (w->device_ticks && w->any_ticks && !w->bad_tick ? BT_PHASE_TICKS : 0u);
}

static inline u64 h_mix64(u64 x) {
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
	x ^= x >> 27; x *= 0x94d049bb133111ebULL;
	x ^= x >> 31;
	return x;
}
static int hash_into(bt_world *w, uint64_t *out, int full = 0) {
	DevWorld &W = w->W;
	CK(cudaMemsetAsync(w->d_hsum, 0, sizeof(u64) * (size_t)W.n_worlds, w->stream));
	CK(cudaMemsetAsync(w->d_hcnt, 0, sizeof(u32) * (size_t)W.n_worlds, w->stream));
	if (W.n_worlds == 1) LAUNCH(w, PK_HASH, k_hash_one, gsblk(W.n), GS_BLOCK, W, w->d_hsum, w->d_hcnt, full); /* one atomic per block */
	else LAUNCH(w, PK_HASH, k_hash, nblk(W.n), BT_BLOCK, W, w->d_hsum, w->d_hcnt, full);
	std::vector<u64> s(W.n_worlds);
	std::vector<u32> c(W.n_worlds);
	if (W.n_worlds == 1) { /* 12 bytes: through the mapped words behind the counters, not through the copy engine */
		LAUNCH(w, PK_HASH, k_publish_hash, 1, 32, (const u64 *)w->d_hsum, (const u32 *)w->d_hcnt, w->d_hflag + CT_WORDS);
		CK(cudaStreamSynchronize(w->stream));
		memcpy(&s[0], &w->h_flag[CT_WORDS], sizeof(u64));
		c[0] = w->h_flag[CT_WORDS + 2];
	} else {
		CK(cudaMemcpyAsync(s.data(), w->d_hsum, sizeof(u64) * s.size(), cudaMemcpyDeviceToHost, w->stream));
		CK(cudaMemcpyAsync(c.data(), w->d_hcnt, sizeof(u32) * c.size(), cudaMemcpyDeviceToHost, w->stream));
		CK(cudaStreamSynchronize(w->stream));
	}
	std::vector<i32> flip, rules, rnd;
	if (full) {
		flip.resize(W.n_worlds); rules.resize(W.n_worlds); rnd.resize(W.n_worlds);
		CK(cudaMemcpy(flip.data(), W.flip, sizeof(i32) * flip.size(), cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(rules.data(), W.gamerules, sizeof(i32) * rules.size(), cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(rnd.data(), W.rnd, sizeof(i32) * rnd.size(), cudaMemcpyDeviceToHost));
	}
	for (int g = 0; g < W.n_worlds; g++) {
		out[g] = s[g] + h_mix64((u64)c[g]);
		if (full) { /* the gamestate scalars: flipFlop_death, rand, gamerules (ent.h:187-189) */
			u64 h = h_mix64((u64)(u32)flip[g] + 0x9E3779B97F4A7C15ULL);
			h = h_mix64(h + (u64)(u32)rnd[g] + 0x9E3779B97F4A7C15ULL);
			h = h_mix64(h + (u64)(u32)rules[g] + 0x9E3779B97F4A7C15ULL);
			out[g] += h;
		}
	}
	return 0;
}

int bt_world_hash(bt_world *w, uint64_t *out) {
	if (!w || !out) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	return hash_into(w, out);
}

int bt_world_hash_full(bt_world *w, uint64_t *out) {
	if (!w || !out) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	return hash_into(w, out, 1);
}

int bt_world_tick(bt_world *w, int32_t n_ticks, uint64_t *hashes) {
	if (!w || n_ticks < 0) return BT_ERR_ARG;
	for (int t = 0; t < n_ticks; t++) {
		int r = bt_world_step(w, tick_phases(w));
		if (r) return r;
		if (hashes) {
			r = hash_into(w, hashes + (size_t)t * w->W.n_worlds);
			if (r) return r;
		}
	}
	CK(cudaStreamSynchronize(w->stream));
	return BT_OK;
}

int bt_world_tick_async(bt_world *w, int32_t n_ticks) { return bt_world_tick(w, n_ticks, NULL); }

int bt_world_tick_timed(bt_world *w, int32_t n_ticks, float *ms) {
	if (!w || n_ticks < 0 || !ms) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a));
	CK(cudaEventCreate(&b));
	CK(cudaStreamSynchronize(w->stream));
	CK(cudaEventRecord(a, w->stream));
	int r = BT_OK;
	for (int t = 0; t < n_ticks && r == BT_OK; t++)
		r = bt_world_step(w, tick_phases(w));
	cudaEventRecord(b, w->stream);
	cudaEventSynchronize(b);
	cudaEventElapsedTime(ms, a, b);
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	return r;
}

/* live-rank maps (shared with the raycast path) */
static int live_ranks(bt_world *w, int cls) {
	const int tiles = (w->W.n + RANK_TILE - 1) / RANK_TILE;
	if (tiles > w->sorthist_len) return BT_ERR_CAPACITY; /* the radix sort's histogram table doubles as scratch */
	LAUNCH(w, cls, k_rank_count, tiles, RANK_TILE, w->W, w->d_sorthist);
	LAUNCH(w, cls, k_rank_scan, 1, 1024, w->d_sorthist, tiles);
	LAUNCH(w, cls, k_rank_apply, tiles, RANK_TILE, w->W, (const u32 *)w->d_sorthist, w->ray.rank2slot, w->ray.slot2rank);
	w->ranks_valid = true;
	return 0;
}
static int ensure_ranks(bt_world *w) {
	if (w->ranks_valid) return 0;
	return live_ranks(w, PK_EXCHANGE);
}

int bt_world_apply_host_deltas(bt_world *w, const int32_t *d_center, const int32_t *d_vel, int32_t n_ents) {
	if (!w || !d_center || !d_vel || n_ents < 0) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	if ((i64)n_ents != w->stats.live_ents) return BT_ERR_ARG;
	CK(cudaMemcpyAsync(w->d_stage_a, d_center, sizeof(i32) * 3 * (size_t)n_ents, cudaMemcpyHostToDevice, w->stream));
	CK(cudaMemcpyAsync(w->d_stage_b, d_vel, sizeof(i32) * 3 * (size_t)n_ents, cudaMemcpyHostToDevice, w->stream));
	bool dense = (i64)w->W.n == w->stats.live_ents;
	if (!dense) { int r = ensure_ranks(w); if (r) return r; }
	LAUNCH(w, PK_EXCHANGE, k_add_deltas, nblk(w->W.n), BT_BLOCK, w->W, w->d_stage_a, w->d_stage_b,
	       dense ? (const i32 *)NULL : (const i32 *)w->ray.slot2rank);
	w->deltas_pending = true;
	return BT_OK;
}

int bt_world_apply_requests(bt_world *w, const bt_request *reqs, int32_t n) {
	if (!w || n < 0 || (n > 0 && !reqs)) return BT_ERR_ARG;
	if (n == 0) return BT_OK;
	CK(cudaSetDevice(w->device));
	DevWorld &W = w->W;
	if (!w->uid_map_valid) {
		std::vector<i32> uid(W.n), world(W.n);
		CK(cudaStreamSynchronize(w->stream));
		if (W.n) {
			CK(cudaMemcpy(uid.data(), W.uid, sizeof(i32) * (size_t)W.n, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(world.data(), W.world, sizeof(i32) * (size_t)W.n, cudaMemcpyDeviceToHost));
		}
		w->uid_map.clear();
		w->uid_map.reserve((size_t)W.n * 2);
		for (i32 i = 0; i < W.n; i++) w->uid_map[((u64)(u32)world[i] << 32) | (u32)uid[i]] = i;
		w->uid_map_valid = true;
	}
	/* group by world, keeping the caller's order inside a world (the updaters of one gamestate are
	 * order-sensitive, e.g. uPickup is first come first served; different worlds are independent) */
	std::vector<i32> order((size_t)n);
	for (i32 i = 0; i < n; i++) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](i32 a, i32 b) { return reqs[a].world < reqs[b].world; });
	std::vector<i32> starts;
	std::vector<BtReq> q((size_t)n);
	bool deltas = false;
	for (i32 i = 0; i < n; i++) {
		const bt_request &r = reqs[order[i]];
		if (i == 0 || r.world != reqs[order[i - 1]].world) starts.push_back(i);
		if (r.kind < 0 || r.kind >= BT_REQ_NUM) { snprintf(g_err, sizeof g_err, "request %d: unknown kind %d", order[i], r.kind); return BT_ERR_ARG; }
		auto it = w->uid_map.find(((u64)(u32)r.world << 32) | (u32)r.me_uid);
		if (it == w->uid_map.end()) { snprintf(g_err, sizeof g_err, "request %d: no entity uid %d in world %d", order[i], r.me_uid, r.world); return BT_ERR_ARG; }
		q[i].kind = r.kind; q[i].me = it->second; q[i].other = -1;
		if (r.kind == BT_REQ_PICKUP) {
			auto jt = w->uid_map.find(((u64)(u32)r.world << 32) | (u32)r.other_uid);
			if (jt == w->uid_map.end()) { snprintf(g_err, sizeof g_err, "request %d: no entity uid %d in world %d", order[i], r.other_uid, r.world); return BT_ERR_ARG; }
			q[i].other = jt->second;
		}
		q[i].a0 = r.a[0]; q[i].a1 = r.a[1]; q[i].a2 = r.a[2];
		/* a HOLD_FREEZE pickup leaves a d_vel behind when prepPhysics applies it (ent.cpp:176-181) */
		deltas |= r.kind == BT_REQ_VEL || r.kind == BT_REQ_CENTER || (r.kind == BT_REQ_PICKUP && (r.a[0] & HOLD_FREEZE));
	}
	starts.push_back(n);
	const i32 n_groups = (i32)starts.size() - 1;
	if (n > w->reqs_cap) {
		if (w->d_reqs) cudaFree(w->d_reqs);
		if (w->d_req_starts) cudaFree(w->d_req_starts);
		w->d_reqs = NULL; w->d_req_starts = NULL; w->reqs_cap = 0;
		CK(cudaMalloc(&w->d_reqs, sizeof(BtReq) * (size_t)n));
		CK(cudaMalloc(&w->d_req_starts, sizeof(i32) * ((size_t)n + 1)));
		w->reqs_cap = n;
	}
	CK(cudaMemcpyAsync(w->d_reqs, q.data(), sizeof(BtReq) * (size_t)n, cudaMemcpyHostToDevice, w->stream));
	CK(cudaMemcpyAsync(w->d_req_starts, starts.data(), sizeof(i32) * starts.size(), cudaMemcpyHostToDevice, w->stream));
	LAUNCH(w, PK_EXCHANGE, k_requests, nblk(n_groups), BT_BLOCK, W, (const BtReq *)w->d_reqs, (const i32 *)w->d_req_starts, n_groups);
	CK(cudaStreamSynchronize(w->stream)); /* q goes out of scope */
	if (deltas) w->deltas_pending = true;
	if (w->full_flush_left < 1) w->full_flush_left = 1; /* the next flush1 has requests to look at */
	return BT_OK;
}

int bt_world_read_state(bt_world *w, int32_t *center, int32_t *vel, int32_t n_ents) {
	if (!w || !center || !vel || n_ents < 0) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	if ((i64)n_ents != w->stats.live_ents) return BT_ERR_ARG;
	bool dense = (i64)w->W.n == w->stats.live_ents;
	if (!dense) { int r = ensure_ranks(w); if (r) return r; }
	LAUNCH(w, PK_EXCHANGE, k_pack_state, nblk(w->W.n), BT_BLOCK, w->W, w->d_stage_a, w->d_stage_b,
	       dense ? (const i32 *)NULL : (const i32 *)w->ray.slot2rank);
	CK(cudaMemcpyAsync(center, w->d_stage_a, sizeof(i32) * 3 * (size_t)n_ents, cudaMemcpyDeviceToHost, w->stream));
	CK(cudaMemcpyAsync(vel, w->d_stage_b, sizeof(i32) * 3 * (size_t)n_ents, cudaMemcpyDeviceToHost, w->stream));
	CK(cudaStreamSynchronize(w->stream));
	return BT_OK;
}

/* ---- the pipelined exchange ---- */
static int px_init(bt_world *w) {
	if (w->px_ready) return 0;
	CK(cudaStreamCreateWithFlags(&w->s_h2d, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&w->s_d2h, cudaStreamNonBlocking));
	const size_t bytes = sizeof(i32) * 3 * (size_t)(w->cap_ents > 0 ? w->cap_ents : 1);
	for (int k = 0; k < 2; k++) {
		for (int j = 0; j < 2; j++) { CK(cudaMalloc(&w->px_in[k][j], bytes)); CK(cudaMalloc(&w->px_out[k][j], bytes)); }
		CK(cudaEventCreateWithFlags(&w->px_arrived[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&w->px_added[k], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&w->px_packed[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&w->px_read[k], cudaEventDisableTiming));
	}
	w->px_ready = true;
	return 0;
}

int bt_world_send_host_deltas(bt_world *w, const int32_t *d_center, const int32_t *d_vel, int32_t n_ents) {
	if (!w || !d_center || !d_vel || n_ents < 0 || n_ents > w->cap_ents) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = px_init(w); if (rr) return rr; }
	if (w->px_sent[0] && w->px_sent[1]) { snprintf(g_err, sizeof g_err, "bt_world_send_host_deltas: two batches are already on their way, commit one first"); return BT_ERR_ARG; }
	const int k = w->px_in_k;
	w->px_in_k ^= 1;
	const size_t bytes = sizeof(i32) * 3 * (size_t)n_ents;
	CK(cudaStreamWaitEvent(w->s_h2d, w->px_added[k], 0)); /* the kernel that last consumed this buffer is done */
	CK(cudaMemcpyAsync(w->px_in[k][0], d_center, bytes, cudaMemcpyHostToDevice, w->s_h2d));
	CK(cudaMemcpyAsync(w->px_in[k][1], d_vel, bytes, cudaMemcpyHostToDevice, w->s_h2d));
	CK(cudaEventRecord(w->px_arrived[k], w->s_h2d));
	w->px_sent[k] = true;
	w->px_sent_n[k] = n_ents;
	return BT_OK;
}

int bt_world_commit_host_deltas(bt_world *w) {
	if (!w || !w->px_ready) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	const int k = w->px_commit_k;
	if (!w->px_sent[k]) { snprintf(g_err, sizeof g_err, "bt_world_commit_host_deltas: nothing was sent"); return BT_ERR_ARG; }
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	w->px_sent[k] = false;
	w->px_commit_k ^= 1;
	if ((i64)w->px_sent_n[k] != w->stats.live_ents) { /* an entity died since the batch was sent: its rows no longer line up */
		snprintf(g_err, sizeof g_err, "bt_world_commit_host_deltas: %d rows were sent, the world holds %lld entities now", w->px_sent_n[k], (long long)w->stats.live_ents);
		return BT_ERR_ARG;
	}
	bool dense = (i64)w->W.n == w->stats.live_ents;
	if (!dense) { int r = ensure_ranks(w); if (r) return r; }
	CK(cudaStreamWaitEvent(w->stream, w->px_arrived[k], 0));
	LAUNCH(w, PK_EXCHANGE, k_add_deltas, nblk(w->W.n), BT_BLOCK, w->W, (const i32 *)w->px_in[k][0], (const i32 *)w->px_in[k][1],
	       dense ? (const i32 *)NULL : (const i32 *)w->ray.slot2rank);
	CK(cudaEventRecord(w->px_added[k], w->stream));
	w->deltas_pending = true;
	return BT_OK;
}

int bt_world_apply_host_deltas_async(bt_world *w, const int32_t *d_center, const int32_t *d_vel, int32_t n_ents) {
	int r = bt_world_send_host_deltas(w, d_center, d_vel, n_ents);
	return r ? r : bt_world_commit_host_deltas(w);
}

int bt_world_read_state_async(bt_world *w, int32_t *center, int32_t *vel, int32_t n_ents, int32_t *ticket) {
	if (!w || !center || !vel || n_ents < 0 || !ticket) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	if ((i64)n_ents != w->stats.live_ents) return BT_ERR_ARG;
	{ int rr = px_init(w); if (rr) return rr; }
	const int k = w->px_out_k;
	w->px_out_k ^= 1;
	const size_t bytes = sizeof(i32) * 3 * (size_t)n_ents;
	bool dense = (i64)w->W.n == w->stats.live_ents;
	if (!dense) { int r = ensure_ranks(w); if (r) return r; }
	CK(cudaStreamWaitEvent(w->stream, w->px_read[k], 0)); /* the copy that last read this buffer is done */
	LAUNCH(w, PK_EXCHANGE, k_pack_state, nblk(w->W.n), BT_BLOCK, w->W, w->px_out[k][0], w->px_out[k][1],
	       dense ? (const i32 *)NULL : (const i32 *)w->ray.slot2rank);
	CK(cudaEventRecord(w->px_packed[k], w->stream));
	CK(cudaStreamWaitEvent(w->s_d2h, w->px_packed[k], 0));
	CK(cudaMemcpyAsync(center, w->px_out[k][0], bytes, cudaMemcpyDeviceToHost, w->s_d2h));
	CK(cudaMemcpyAsync(vel, w->px_out[k][1], bytes, cudaMemcpyDeviceToHost, w->s_d2h));
	CK(cudaEventRecord(w->px_read[k], w->s_d2h));
	*ticket = k;
	CHECK_LAUNCHES(w);
	return BT_OK;
}

int bt_world_read_state_wait(bt_world *w, int32_t ticket) {
	if (!w || ticket < 0 || ticket > 1 || !w->px_ready) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaEventSynchronize(w->px_read[ticket]));
	return BT_OK;
}

int bt_world_sync(bt_world *w) {
	if (!w) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaStreamSynchronize(w->stream));
	if (w->px_ready) { CK(cudaStreamSynchronize(w->s_h2d)); CK(cudaStreamSynchronize(w->s_d2h)); }
	return BT_OK;
}

int64_t bt_world_live_ents(bt_world *w) {
	if (!w) return (int64_t)BT_ERR_ARG;
	if (w->fused_dirty) { /* one-block-per-world ticks keep their dead on the device until somebody asks */
		if (cudaSetDevice(w->device) != cudaSuccess) return (int64_t)BT_ERR_CUDA;
		int r = fused_settle(w, false);
		if (r) return (int64_t)r;
	}
	return (int64_t)w->stats.live_ents;
}

int bt_world_stats(bt_world *w, bt_tick_stats *out, int reset) {
	if (!w || !out) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	u64 s[SK_NUM];
	{ int rr = read_stats(w, s); if (rr) return rr; }
	w->stats.candidate_pairs = (i64)s[SK_CAND]; w->stats.positives = (i64)s[SK_POS];
	w->stats.pushes = (i64)s[SK_PUSH]; w->stats.serial_pushes = (i64)s[SK_SERIAL];
	w->stats.crushes = (i64)s[SK_CRUSH]; w->stats.fumbles = (i64)s[SK_FUMBLE];
	{
		u32 un = 0;
		CK(cudaMemcpy(&un, &w->W.ctr[CT_UNSUPPORTED], sizeof(u32), cudaMemcpyDeviceToHost));
		w->stats.unsupported_events = (i64)un;
		CK(cudaMemcpy(&un, &w->W.ctr[CT_ORDER_TIES], sizeof(u32), cudaMemcpyDeviceToHost));
		w->stats.order_ties = (i64)un;
	}
	*out = w->stats;
	if (reset) {
		i64 live = w->stats.live_ents;
		memset(&w->stats, 0, sizeof(w->stats));
		w->stats.live_ents = live;
		CK(cudaMemsetAsync(w->W.stat, 0, sizeof(u64) * BT_STAT_ROWS * SK_NUM, w->stream));
		w->seen_eticks = w->seen_iters = 0;
		CK(cudaMemsetAsync(&w->W.ctr[CT_UNSUPPORTED], 0, sizeof(u32), w->stream));
		CK(cudaMemsetAsync(w->W.wunsup, 0, sizeof(u32) * (size_t)w->cap_worlds, w->stream));
		CK(cudaMemsetAsync(&w->W.ctr[CT_ORDER_TIES], 0, sizeof(u32), w->stream));
		CK(cudaMemsetAsync(w->W.wties, 0, sizeof(u32) * (size_t)w->cap_worlds, w->stream));
	}
	return BT_OK;
}

int bt_world_drain_events(bt_world *w, bt_event *out, int32_t max, int32_t *n) {
	if (!w || !n || max < 0 || (max > 0 && !out)) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaStreamSynchronize(w->stream));
	u32 cnt = 0;
	CK(cudaMemcpy(&cnt, &w->W.ctr[CT_EVENT_N], sizeof(u32), cudaMemcpyDeviceToHost));
	bool overflow = cnt > (u32)w->W.event_cap;
	if (overflow) cnt = (u32)w->W.event_cap;
	if (cnt > (u32)max) { /* nothing is consumed: the caller comes back with room for all of them */
		*n = (int32_t)cnt;
		snprintf(g_err, sizeof g_err, "bt_world_drain_events: %u events queued, room for %d", cnt, max);
		return BT_ERR_CAPACITY;
	}
	u32 take = cnt;
	std::vector<int4> ev(take);
	std::vector<i32> uid(w->W.n), world(w->W.n);
	if (take) {
		CK(cudaMemcpy(ev.data(), w->W.events, sizeof(int4) * take, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(uid.data(), w->W.uid, sizeof(i32) * (size_t)w->W.n, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(world.data(), w->W.world, sizeof(i32) * (size_t)w->W.n, cudaMemcpyDeviceToHost));
	}
	/* push events keep their (canonical) order; crush events follow them, most recent death first,
	 * the order invokeOnCrush walks gs->deadTail in (ent.cpp:464-469) */
	std::stable_sort(ev.begin(), ev.end(), [](const int4 &a, const int4 &b) {
		int ka = a.x >> 8, kb = b.x >> 8;
		if (ka != kb) return ka < kb;
		return ka == 3 && (u32)a.w > (u32)b.w;
	});
	for (u32 i = 0; i < take; i++) {
		out[i].kind = ev[i].x >> 8; out[i].handler = ev[i].x & 0xff; out[i].world = world[ev[i].y];
		out[i].me_uid = uid[ev[i].y]; out[i].him_uid = uid[ev[i].z]; out[i].arg = ev[i].w;
	}
	*n = (int32_t)take;
	CK(cudaMemsetAsync(&w->W.ctr[CT_EVENT_N], 0, sizeof(u32), w->stream));
	return overflow ? BT_ERR_OVERFLOW : BT_OK;
}

int bt_world_unsupported(bt_world *w, uint32_t *per_world) {
	if (!w || !per_world) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaStreamSynchronize(w->stream));
	CK(cudaMemcpy(per_world, w->W.wunsup, sizeof(u32) * (size_t)w->W.n_worlds, cudaMemcpyDeviceToHost));
	return BT_OK;
}

int bt_world_order_ties(bt_world *w, uint32_t *per_world) {
	if (!w || !per_world) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaStreamSynchronize(w->stream));
	CK(cudaMemcpy(per_world, w->W.wties, sizeof(u32) * (size_t)w->W.n_worlds, cudaMemcpyDeviceToHost));
	return BT_OK;
}

/* ---------------------------------------------------------------- dup (ent.cpp:1135-1200) */
} /* extern "C" */

/* every device pointer inside a POD block is re-based from src's allocation to dst's */
static void rebase(void *block, size_t bytes, const bt_world *src, const bt_world *dst) {
	uintptr_t *wd = (uintptr_t *)block;
	for (size_t i = 0; i + sizeof(uintptr_t) <= bytes; i += sizeof(uintptr_t), wd++)
		for (size_t a = 0; a < src->allocs.size(); a++)
			if (*wd == (uintptr_t)src->allocs[a]) { *wd = (uintptr_t)dst->allocs[a]; break; }
}

extern "C" {

/* dst := src.  Both worlds must have been created with the same capacities on the same
 * device (then their allocation lists line up); the copy is device-to-device on dst's stream. */
int bt_world_assign(bt_world *dst, bt_world *src) {
	if (!dst || !src || dst == src) return BT_ERR_ARG;
	if (dst->device != src->device || dst->cap_ents != src->cap_ents || dst->cap_sliders != src->cap_sliders ||
	    dst->cap_wires != src->cap_wires || dst->cap_worlds != src->cap_worlds || dst->allocs.size() != src->allocs.size())
		return BT_ERR_CAPACITY;
	CK(cudaSetDevice(src->device));
	CK(cudaStreamSynchronize(src->stream));
	for (size_t a = 0; a < src->allocs.size(); a++)
		CK(cudaMemcpyAsync(dst->allocs[a], src->allocs[a], src->alloc_bytes[a], cudaMemcpyDeviceToDevice, dst->stream));
	dst->W = src->W;
	rebase(&dst->W, sizeof(dst->W), src, dst);
	dst->ray = src->ray;
	rebase(&dst->ray, sizeof(dst->ray), src, dst);
	dst->arrays = src->arrays;
	rebase(&dst->arrays, sizeof(dst->arrays), src, dst);
	dst->fused_small = src->fused_small; dst->small_warps = src->small_warps; dst->fused_dirty = src->fused_dirty; dst->small_biggest = src->small_biggest; dst->wq_host_pending = src->wq_host_pending; dst->seen_eticks = src->seen_eticks; dst->seen_iters = src->seen_iters;
	dst->root_phase_ctr = src->root_phase_ctr; dst->sliders_used = src->sliders_used; dst->wires_used = src->wires_used; dst->poked = src->poked;
	dst->tree_valid = src->tree_valid; dst->deltas_pending = src->deltas_pending; dst->any_rules = src->any_rules;
	dst->full_flush_left = src->full_flush_left; dst->scan_flushes_left = src->scan_flushes_left; dst->ctrls_fused = src->ctrls_fused;
	dst->uid_map_valid = false; dst->ctr_fresh = false; dst->pdl = src->pdl; dst->pdl_long = src->pdl_long; dst->pdl_opt = src->pdl_opt; dst->margin_opt = src->margin_opt; dst->narrow_blocks = src->narrow_blocks; dst->ray_dynamic = src->ray_dynamic; dst->ray_blocks = src->ray_blocks; dst->narrow_threads = src->narrow_threads; dst->small_worlds = src->small_worlds;
	dst->any_ticks = src->any_ticks; dst->any_rand = src->any_rand; dst->bad_tick = src->bad_tick; dst->device_ticks = src->device_ticks;
	dst->ticks_since_build = src->ticks_since_build; dst->rebuild_every = src->rebuild_every;
	dst->max_iterations = src->max_iterations; dst->pair_cache = src->pair_cache;
	dst->seen_pair_builds = src->seen_pair_builds; dst->seen_pair_patches = src->seen_pair_patches; dst->pair_builds_at_tree = src->pair_builds_at_tree;
	dst->predicted_iters = src->predicted_iters; dst->run_ahead = src->run_ahead;
	dst->stats = src->stats; dst->up_root_order = src->up_root_order;
	dst->ray_valid = src->ray_valid; dst->ranks_valid = src->ranks_valid;
	CK(cudaStreamSynchronize(dst->stream));
	return BT_OK;
}

/* dup(gs) ent.cpp:1135-1200: a second, independent world in the same state (cached pairs and
 * tree included, like the reference clones velbox's cached intersects). */
int bt_world_dup(bt_world *src, bt_world **out) {
	if (!src || !out) return BT_ERR_ARG;
	int r = bt_world_create(out, src->device, src->cap_ents, src->cap_sliders, src->cap_wires, src->cap_worlds);
	if (r) return r;
	r = bt_world_assign(*out, src);
	if (r) { bt_world_destroy(*out); *out = NULL; }
	return r;
}

int bt_world_profile_enable(bt_world *w, int on) {
	if (!w) return BT_ERR_ARG;
	w->profiling = on != 0;
	return BT_OK;
}

int bt_world_profile_read(bt_world *w, bt_profile *out, int reset) {
	if (!w || !out) return BT_ERR_ARG;
	out->n = PK_NUM;
	for (int i = 0; i < PK_NUM; i++) {
		out->name[i] = PK_NAMES[i];
		out->launches[i] = w->prof_launches[i];
		out->ms[i] = w->prof_ms[i];
	}
	if (reset) {
		memset(w->prof_launches, 0, sizeof(w->prof_launches));
		memset(w->prof_ms, 0, sizeof(w->prof_ms));
	}
	return BT_OK;
}

/* ---------------------------------------------------------------- download */

int bt_world_count(bt_world *w, int32_t *n_ents, int32_t *n_sliders, int32_t *n_wires) {
	if (!w) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaStreamSynchronize(w->stream)); /* the copies below do not run on the world's (non-blocking) stream */
	const int n = w->W.n;
	std::vector<u8> inList(n);
	std::vector<i32> so(n + 1), wc(n);
	if (n) {
		CK(cudaMemcpy(inList.data(), w->W.inList, n, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(so.data(), w->W.slider_off, sizeof(i32) * (n + 1), cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(wc.data(), w->W.wire_cnt, sizeof(i32) * n, cudaMemcpyDeviceToHost));
	}
	i32 a = 0, b = 0, c = 0;
	for (int i = 0; i < n; i++)
		if (inList[i]) { a++; b += so[i + 1] - so[i]; c += wc[i]; }
	if (n_ents) *n_ents = a;
	if (n_sliders) *n_sliders = b;
	if (n_wires) *n_wires = c;
	return BT_OK;
}

} /* extern "C" */
template <typename T> static int d2h(std::vector<T> &dst, const T *src, size_t cnt) {
	dst.resize(cnt);
	if (cnt) CK(cudaMemcpy(dst.data(), src, sizeof(T) * cnt, cudaMemcpyDeviceToHost));
	return 0;
}
extern "C" {

int bt_world_download(bt_world *w, bt_world_desc *d) {
	if (!w || !d) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	CK(cudaStreamSynchronize(w->stream));
	DevWorld &W = w->W;
	const int n = W.n, nw = W.n_worlds;
#define DN(vec, src, cnt) do { int _r = d2h(vec, src, (size_t)(cnt)); if (_r) return _r; } while (0)
	std::vector<u8> inList, newDrop, dm0, dm1, ff;
	std::vector<i32> world, uid, friction, holder, holdFlags, holdee, LLn, newHolder, newHoldFlags, so, sv, smin, smax, wo, wc, wt;
	std::vector<int4> center, old, vel, radius, dc, dv;
	std::vector<u32> type, collide, ntype, ncollide, hphys, hpair, htick, bv, bv2;
	DN(inList, W.inList, n); DN(newDrop, W.newDrop, n); DN(dm0, W.dead_max0, n); DN(dm1, W.dead_max1, n);
	DN(ff, W.fullFlush, n); DN(world, W.world, n); DN(uid, W.uid, n); DN(friction, W.friction, n);
	DN(holder, W.holder, n); DN(holdFlags, W.holdFlags, n); DN(holdee, W.holdee, n); DN(LLn, W.LLn, n);
	DN(newHolder, W.newHolder, n); DN(newHoldFlags, W.newHoldFlags, n);
	DN(center, W.center, n); DN(old, W.old, n); DN(vel, W.vel, n); DN(radius, W.radius, n);
	DN(dc, W.d_center, n); DN(dv, W.d_vel, n);
	DN(type, W.type, n); DN(collide, W.collide, n); DN(ntype, W.ntype, n); DN(ncollide, W.ncollide, n);
	DN(hphys, W.hphys, n); DN(hpair, W.hpair, n); DN(htick, W.htick, n); DN(bv, W.btn_v, n); DN(bv2, W.btn_v2, n);
	DN(so, W.slider_off, n + 1); DN(wo, W.wire_off, n + 1); DN(wc, W.wire_cnt, n);
	DN(sv, W.sv, so.empty() ? 0 : so[n]); DN(smin, W.smin, so.empty() ? 0 : so[n]); DN(smax, W.smax, so.empty() ? 0 : so[n]);
	DN(wt, W.wire_to, wo.empty() ? 0 : wo[n]);
#undef DN
	std::vector<i32> map(n + 1, -1);
	i32 k = 0;
	for (int i = 0; i < n; i++)
		if (inList[i]) map[i] = k++;
	if (k > d->n_ents) return BT_ERR_CAPACITY;
	auto ref = [&](i32 x) { return x >= 0 ? map[x] : -1; };
	std::vector<i32> flip(nw), rules(nw), rnd(nw);
	CK(cudaMemcpy(flip.data(), W.flip, sizeof(i32) * nw, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(rules.data(), W.gamerules, sizeof(i32) * nw, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(rnd.data(), W.rnd, sizeof(i32) * nw, cudaMemcpyDeviceToHost));
	for (int g = 0; g < nw; g++) { d->w_flip_flop[g] = flip[g]; d->w_gamerules[g] = rules[g]; d->w_rand[g] = rnd[g]; }
	/* gs->rootEnts order (0 = head) = descending rootKey inside a world */
	{
		std::vector<u64> rk;
		int rr = d2h(rk, (const u64 *)W.rootKey, (size_t)n);
		if (rr) return rr;
		std::vector<std::vector<std::pair<u64, i32>>> per(nw);
		for (int i = 0; i < n; i++) {
			if (!inList[i] || holder[i] >= 0) continue;
			per[world[i]].push_back(std::make_pair(~rk[i], (i32)i));
		}
		for (int g = 0; g < nw; g++) {
			std::sort(per[g].begin(), per[g].end());
			for (size_t p = 0; p < per[g].size(); p++) d->root_order[map[per[g][p].second]] = (i32)p;
		}
	}
	i32 sOut = 0, wOut = 0;
	for (int i = 0; i < n; i++) {
		if (!inList[i]) continue;
		i32 o = map[i];
		if (d->world_id) d->world_id[o] = world[i];
		d->uid[o] = uid[i];
		const int4 *src[6] = {&center[i], &vel[i], &old[i], &dc[i], &dv[i], &radius[i]};
		int32_t *dst[6] = {d->center, d->vel, d->old, d->d_center, d->d_vel, d->radius};
		for (int f = 0; f < 6; f++) { dst[f][3 * o] = src[f]->x; dst[f][3 * o + 1] = src[f]->y; dst[f][3 * o + 2] = src[f]->z; }
		d->type_mask[o] = type[i]; d->collide_mask[o] = collide[i];
		d->new_type_mask[o] = ntype[i]; d->new_collide_mask[o] = ncollide[i];
		d->friction[o] = friction[i]; d->holder[o] = ref(holder[i]); d->hold_flags[o] = holdFlags[i];
		d->holdee_head[o] = ref(holdee[i]); d->holdee_next[o] = holder[i] >= 0 ? ref(LLn[i]) : -1;
		if (holder[i] >= 0) d->root_order[o] = -1;
		d->new_holder[o] = ref(newHolder[i]); d->new_hold_flags[o] = newHoldFlags[i];
		uint8_t *h = d->handlers + (size_t)BT_H_NUM * o;
		h[BT_H_WHOMOVES] = hphys[i] & 0xff; h[BT_H_PUSH] = (hphys[i] >> 8) & 0xff;
		h[BT_H_PUSHED] = (hphys[i] >> 16) & 0xff; h[BT_H_CRUSH] = (hphys[i] >> 24) & 0xff;
		h[BT_H_ONFUMBLE] = hpair[i] & 0xff; h[BT_H_ONFUMBLED] = (hpair[i] >> 8) & 0xff;
		h[BT_H_ONPICKUP] = (hpair[i] >> 16) & 0xff; h[BT_H_ONPICKEDUP] = (hpair[i] >> 24) & 0xff;
		h[BT_H_TICK] = htick[i] & 0xff; h[BT_H_TICKHELD] = (htick[i] >> 8) & 0xff;
		uint8_t *b = d->btn + 8 * (size_t)o;
		for (int q = 0; q < 4; q++) { b[q] = (bv[i] >> (8 * q)) & 0xff; b[4 + q] = (bv2[i] >> (8 * q)) & 0xff; }
		uint8_t *m = d->misc + (size_t)BT_M_NUM * o;
		m[BT_M_NEWDROP] = newDrop[i]; m[BT_M_DEADMAX0] = dm0[i]; m[BT_M_DEADMAX1] = dm1[i]; m[BT_M_FULLFLUSH] = ff[i];
		d->slider_off[o] = sOut;
		for (int s = so[i]; s < so[i + 1]; s++, sOut++) {
			if (sOut >= d->n_sliders) return BT_ERR_CAPACITY;
			d->slider_v[sOut] = sv[s]; d->slider_min[sOut] = smin[s]; d->slider_max[sOut] = smax[s];
		}
		d->wire_off[o] = wOut;
		for (int q = 0; q < wc[i]; q++, wOut++) {
			if (wOut >= d->n_wires) return BT_ERR_CAPACITY;
			d->wire_to[wOut] = ref(wt[wo[i] + q]);
		}
	}
	d->slider_off[k] = sOut; d->wire_off[k] = wOut;
	d->n_ents = k; d->n_sliders = sOut; d->n_wires = wOut; d->n_worlds = nw;
	return BT_OK;
}

/* ---------------------------------------------------------------- incremental host binding */

/* bytes per element of a BT_A_* array */
static int array_elem_bytes(int a) {
	if (a >= BT_A_CENTER && a <= BT_A_RADIUS) return 16;
	if (a >= BT_A_NEWDROP && a <= BT_A_INLIST) return 1;
	if (a == BT_A_ROOTKEY || a == BT_A_DEADLIST) return 8;
	return 4;
}
static int array_extent(bt_world *w, int a, i32 *count) {
	if (a < 0 || a >= BT_A_NUM) return BT_ERR_ARG;
	if (a >= BT_A_SV && a <= BT_A_SMAX) *count = w->sliders_used;
	else if (a == BT_A_WIRE_TO) *count = w->wires_used;
	else if (a >= BT_A_W_FLIP && a <= BT_A_W_RAND) *count = w->W.n_worlds;
	else if (a == BT_A_SLIDER_OFF) *count = w->W.n + 1;
	else if (a == BT_A_DEADLIST) {
		if (!w->ctr_fresh || w->poked) { int r = fetch_ctr(w); if (r) return r; }
		*count = (i32)std::min<u32>(w->h_flag[CT_DEAD_N], (u32)w->W.dead_cap);
	} else *count = w->W.n;
	return BT_OK;
}

int bt_world_extent(bt_world *w, int32_t array, int32_t *count) {
	if (!w || !count) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	return array_extent(w, array, count);
}

int bt_world_pull(bt_world *w, int32_t array, int32_t first, int32_t count, void *dst) {
	if (!w || first < 0 || count < 0 || (count > 0 && !dst)) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	i32 ext = 0;
	int r = array_extent(w, array, &ext);
	if (r) return r;
	if ((i64)first + count > ext) { snprintf(g_err, sizeof g_err, "bt_world_pull: [%d, %d) exceeds the %d elements of array %d", first, first + count, ext, array); return BT_ERR_ARG; }
	if (!count) return BT_OK;
	const size_t eb = (size_t)array_elem_bytes(array);
	CHECK_LAUNCHES(w);
	CK(cudaMemcpyAsync(dst, (const char *)w->arrays.base[array] + eb * (size_t)first, eb * (size_t)count, cudaMemcpyDeviceToHost, w->stream));
	CK(cudaStreamSynchronize(w->stream));
	return BT_OK;
}

int64_t bt_world_next_root_phase(bt_world *w) { return w ? (int64_t)++w->root_phase_ctr : (int64_t)BT_ERR_ARG; }

int bt_world_tree_edits(bt_world *w, uint32_t *count) {
	if (!w || !count) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	if (!w->ctr_fresh || w->poked) { int r = fetch_ctr(w); if (r) return r; }
	*count = w->h_flag[CT_TREE_EDITS];
	return BT_OK;
}

/* after anything the host wrote: the next flush1 looks at everything (both dead_max slots), flush2 runs in
 * full, the device counters are stale */
static void host_wrote(bt_world *w) {
	w->poked = true;
	w->ctr_fresh = false;
	w->deltas_pending = true;
	if (w->full_flush_left < 2) w->full_flush_left = 2;
	if (w->scan_flushes_left < 2) w->scan_flushes_left = 2;
	w->uid_map_valid = false;
	w->ray_valid = false;
}

int bt_world_grow(bt_world *w, int32_t n_new, const int32_t *n_sliders, const int32_t *wire_cap, int32_t world, int32_t *first_slot) {
	if (!w || n_new < 0 || !first_slot || (n_new > 0 && (!n_sliders || !wire_cap))) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	DevWorld &W = w->W;
	if (world < 0 || world >= W.n_worlds) return BT_ERR_ARG;
	if (W.n_worlds > 1) { snprintf(g_err, sizeof g_err, "bt_world_grow: batched device worlds keep the entities of one world contiguous and cannot grow"); return BT_ERR_UNSUPPORTED; }
	*first_slot = W.n;
	if (!n_new) return BT_OK;
	i64 ns = 0, nwc = 0;
	for (i32 i = 0; i < n_new; i++) {
		if (n_sliders[i] < 0 || wire_cap[i] < 0) return BT_ERR_ARG;
		ns += n_sliders[i]; nwc += wire_cap[i];
	}
	if ((i64)W.n + n_new > w->cap_ents || w->sliders_used + ns > w->cap_sliders || w->wires_used + nwc > (i64)w->cap_wires + w->cap_ents) {
		snprintf(g_err, sizeof g_err, "bt_world_grow: %d more entities / %lld sliders / %lld wire slots exceed the capacities given to bt_world_create", n_new, (long long)ns, (long long)nwc);
		return BT_ERR_CAPACITY;
	}
	std::vector<i32> so((size_t)n_new + 1), wo((size_t)n_new), wc((size_t)n_new);
	std::vector<i32> sinit((size_t)ns * 3);
	i32 s = w->sliders_used, q = w->wires_used;
	for (i32 i = 0; i < n_new; i++) {
		so[i] = s; wo[i] = q; wc[i] = wire_cap[i];
		s += n_sliders[i]; q += wire_cap[i];
	}
	so[n_new] = s;
	const i32 first = W.n;
	LAUNCH(w, PK_EXCHANGE, k_grow_init, nblk(n_new), BT_BLOCK, W, first, n_new, world);
	/* slider_off[first] already holds sliders_used (the CSR end of the slots before) */
	CK(cudaMemcpyAsync(W.slider_off + first, so.data(), sizeof(i32) * ((size_t)n_new + 1), cudaMemcpyHostToDevice, w->stream));
	CK(cudaMemcpyAsync(W.wire_off + first, wo.data(), sizeof(i32) * (size_t)n_new, cudaMemcpyHostToDevice, w->stream));
	CK(cudaMemcpyAsync(W.wire_cap + first, wc.data(), sizeof(i32) * (size_t)n_new, cudaMemcpyHostToDevice, w->stream));
	if (ns) { /* prepNewSliders entFuncs.c:13-18: v 0, max INT32_MIN, min INT32_MAX */
		std::vector<i32> z((size_t)ns, 0), mn((size_t)ns, INT32_MAX), mx((size_t)ns, INT32_MIN);
		CK(cudaMemcpyAsync(W.sv + w->sliders_used, z.data(), sizeof(i32) * (size_t)ns, cudaMemcpyHostToDevice, w->stream));
		CK(cudaMemcpyAsync(W.smin + w->sliders_used, mn.data(), sizeof(i32) * (size_t)ns, cudaMemcpyHostToDevice, w->stream));
		CK(cudaMemcpyAsync(W.smax + w->sliders_used, mx.data(), sizeof(i32) * (size_t)ns, cudaMemcpyHostToDevice, w->stream));
		CK(cudaStreamSynchronize(w->stream));
	}
	const i32 wend = first + n_new;
	CK(cudaMemcpyAsync(W.wend + world, &wend, sizeof(i32), cudaMemcpyHostToDevice, w->stream));
	CK(cudaStreamSynchronize(w->stream));
	W.n = first + n_new;
	w->small_biggest = W.n; /* one world: its slot count */
	w->sliders_used = s; w->wires_used = q;
	if (W.small_mode && W.n > BT_SMALL_WORLD) { /* the brute-force pair search holds one world per block */
		W.small_mode = 0;
		if (w->margin_opt < 0) W.margin = BT_TREE_MARGIN;
	}
	w->tree_valid = false; /* the leaf set is about to change */
	{ int rr = raise_cache_invalid(w); if (rr) return rr; }
	host_wrote(w);
	CHECK_LAUNCHES(w);
	return BT_OK;
}

int bt_world_poke(bt_world *w, const bt_poke *pokes, int32_t n) {
	if (!w || n < 0 || (n > 0 && !pokes)) return BT_ERR_ARG;
	if (!n) return BT_OK;
	CK(cudaSetDevice(w->device));
	DevWorld &W = w->W;
	std::vector<bt_poke> raw, ops;
	raw.reserve((size_t)n);
	bool leafset = false;
	i32 last_rebox = -1;
	for (i32 i = 0; i < n; i++) {
		const bt_poke &p = pokes[i];
		if (p.array >= 0 && p.array < BT_A_NUM) {
			i32 ext = 0;
			if (p.array == BT_A_DEADLIST || p.array == BT_A_SLIDER_OFF || p.array == BT_A_WIRE_OFF || p.array == BT_A_WIRE_CAP) { snprintf(g_err, sizeof g_err, "poke %d: array %d is read-only", i, p.array); return BT_ERR_ARG; }
			int r = array_extent(w, p.array, &ext);
			if (r) return r;
			if (p.index < 0 || p.index >= ext) { snprintf(g_err, sizeof g_err, "poke %d: element %d outside the %d of array %d", i, p.index, ext, p.array); return BT_ERR_ARG; }
			raw.push_back(p);
			if ((p.array == BT_A_TYPE || p.array == BT_A_COLLIDE) && last_rebox != p.index) {
				bt_poke o = p; o.array = BT_OP_REBOX; ops.push_back(o); last_rebox = p.index; leafset = true;
			}
			if (p.array == BT_A_INLIST || p.array == BT_A_DEAD) leafset = true;
		} else if (p.array == BT_OP_SPAWN || p.array == BT_OP_UWIRE || p.array == BT_OP_UUNWIRE) {
			if (p.index < 0 || p.index >= W.n || (p.array != BT_OP_SPAWN && (p.v[0] < 0 || p.v[0] >= W.n))) { snprintf(g_err, sizeof g_err, "poke %d: slot out of range", i); return BT_ERR_ARG; }
			ops.push_back(p);
			if (p.array == BT_OP_SPAWN) leafset = true;
			else w->wq_host_pending = true;
		} else { snprintf(g_err, sizeof g_err, "poke %d: unknown array / operation %d", i, p.array); return BT_ERR_ARG; }
	}
	const i32 total = (i32)(raw.size() + ops.size());
	if (total > w->pokes_cap) {
		if (w->d_pokes) cudaFree(w->d_pokes);
		w->d_pokes = NULL; w->pokes_cap = 0;
		const i32 cap = std::max(total * 2, 1024);
		CK(cudaMalloc(&w->d_pokes, sizeof(bt_poke) * (size_t)cap));
		w->pokes_cap = cap;
	}
	if (!raw.empty()) CK(cudaMemcpyAsync(w->d_pokes, raw.data(), sizeof(bt_poke) * raw.size(), cudaMemcpyHostToDevice, w->stream));
	if (!ops.empty()) CK(cudaMemcpyAsync(w->d_pokes + raw.size(), ops.data(), sizeof(bt_poke) * ops.size(), cudaMemcpyHostToDevice, w->stream));
	if (!raw.empty()) LAUNCH(w, PK_EXCHANGE, k_poke_raw, nblk((i64)raw.size()), BT_BLOCK, W, w->arrays, (const bt_poke *)w->d_pokes, (i32)raw.size());
	if (!ops.empty()) LAUNCH(w, PK_EXCHANGE, k_poke_ops, 1, 32, W, (const bt_poke *)(w->d_pokes + raw.size()), (i32)ops.size());
	CK(cudaStreamSynchronize(w->stream)); /* raw / ops go out of scope */
	if (leafset) {
		w->tree_valid = false; w->ranks_valid = false;
		int rr = raise_cache_invalid(w);
		if (rr) return rr;
	}
	host_wrote(w);
	CHECK_LAUNCHES(w);
	return BT_OK;
}

int bt_world_set_wires(bt_world *w, int32_t slot, const int32_t *targets, int32_t n, int32_t reserve) {
	if (!w || slot < 0 || slot >= w->W.n || n < 0 || (n > 0 && !targets) || reserve < 0) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	DevWorld &W = w->W;
	for (i32 k = 0; k < n; k++)
		if (targets[k] < 0 || targets[k] >= W.n) return BT_ERR_ARG;
	i32 hdr[2] = {0, 0}; /* off, cap */
	CK(cudaStreamSynchronize(w->stream));
	CK(cudaMemcpy(&hdr[0], W.wire_off + slot, sizeof(i32), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(&hdr[1], W.wire_cap + slot, sizeof(i32), cudaMemcpyDeviceToHost));
	if (n + reserve > hdr[1]) { /* a fresh block at the end of the arena; the old one is abandoned */
		const i32 cap = n + reserve + 2;
		if ((i64)w->wires_used + cap > (i64)w->cap_wires + w->cap_ents) { snprintf(g_err, sizeof g_err, "bt_world_set_wires: wire arena exhausted"); return BT_ERR_CAPACITY; }
		hdr[0] = w->wires_used; hdr[1] = cap;
		w->wires_used += cap;
		CK(cudaMemcpy(W.wire_off + slot, &hdr[0], sizeof(i32), cudaMemcpyHostToDevice));
		CK(cudaMemcpy(W.wire_cap + slot, &hdr[1], sizeof(i32), cudaMemcpyHostToDevice));
	}
	if (n) CK(cudaMemcpy(W.wire_to + hdr[0], targets, sizeof(i32) * (size_t)n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(W.wire_cnt + slot, &n, sizeof(i32), cudaMemcpyHostToDevice));
	host_wrote(w);
	return BT_OK;
}

/* ---------------------------------------------------------------- raycast */
} /* extern "C" */

/* the ray tree reuses the collision-tree kernels on a DevWorld whose tree pointers are
 * swapped for the RayTree arrays and whose ray_mode is set */
static void ray_swap(bt_world *w) {
	DevWorld &W = w->W;
	RayTree &R = w->ray;
	std::swap(W.keys, R.keys); std::swap(W.keys_alt, R.keys_alt);
	std::swap(W.leafEnt, R.leafEnt); std::swap(W.leafEnt_alt, R.leafEnt_alt);
	std::swap(W.leafParent, R.leafParent); std::swap(W.nodeParent, R.nodeParent); std::swap(W.wroot, R.wroot);
	std::swap(W.nodeCount, R.nodeCount); std::swap(W.nodes, R.nodes);
	std::swap(W.nodeBox, R.nodeBox); std::swap(W.leafBox, R.leafBox);
	std::swap(W.n_leaves, R.n_leaves);
	for (int k = 0; k < 3; k++) std::swap(W.key_shift[k], R.key_shift[k]);
	W.ray_mode ^= 1;
}

static int ray_build(bt_world *w) {
	bool tv = w->tree_valid;
	i32 tsb = w->ticks_since_build;
	const u32 pbat = w->pair_builds_at_tree; /* build_tree bumps it: the collision tree's rebuild schedule must not notice */
	i64 builds = w->stats.tree_builds, refits = w->stats.tree_refits;
	ray_swap(w);
	int r = build_tree(w);
	if (!r) r = refit(w, false);
	if (!r) r = live_ranks(w, PK_RAY_BUILD);
	if (!r && w->W.n_leaves > 0) LAUNCH(w, PK_RAY_BUILD, k_ray_leafpos, gsblk(w->W.n_leaves), GS_BLOCK, w->W, w->ray.leafPos);
	ray_swap(w);
	w->tree_valid = tv; w->ticks_since_build = tsb; w->pair_builds_at_tree = pbat;
	w->stats.tree_builds = builds; w->stats.tree_refits = refits;
	if (r) return r;
	w->ray_valid = true;
	return 0;
}

static int ray_cast(bt_world *w, const i32 *d_origin, const i32 *d_look, i32 n_rays, i32 *d_winner) {
	const i32 *order = NULL;
	if (w->sort_rays && n_rays >= 4096 && w->ray.n_leaves > 1) {
		/* (leaf position of the origin, ray index) sorted by position: see k_ray_keys */
		if (n_rays > w->rays_cap) {
			cudaFree(w->d_raykeys); cudaFree(w->d_raykeys_alt); cudaFree(w->d_rayidx); cudaFree(w->d_rayidx_alt); cudaFree(w->d_rayhist);
			w->d_raykeys = w->d_raykeys_alt = NULL; w->d_rayidx = w->d_rayidx_alt = NULL; w->d_rayhist = NULL; w->rays_cap = 0;
			const size_t nr = (size_t)n_rays;
			CK(cudaMalloc(&w->d_raykeys, sizeof(u64) * nr)); CK(cudaMalloc(&w->d_raykeys_alt, sizeof(u64) * nr));
			CK(cudaMalloc(&w->d_rayidx, sizeof(i32) * nr)); CK(cudaMalloc(&w->d_rayidx_alt, sizeof(i32) * nr));
			CK(cudaMalloc(&w->d_rayhist, sizeof(u32) * 256 * ((nr + SORT_TILE - 1) / SORT_TILE + 1)));
			w->rays_cap = n_rays;
		}
		LAUNCH(w, PK_RAYCAST, k_ray_keys, gsblk(n_rays), GS_BLOCK, d_origin, n_rays, (const i32 *)w->ray.rank2slot, (const i32 *)w->ray.leafPos,
		       (i32)w->stats.live_ents, w->d_raykeys, w->d_rayidx);
		int r = radix_sort_buf(w, PK_RAYCAST, &w->d_raykeys, &w->d_rayidx, &w->d_raykeys_alt, &w->d_rayidx_alt, n_rays,
		                       bits_for((u32)w->ray.n_leaves), w->d_rayhist);
		if (r) return r;
		order = w->d_rayidx;
	}
	const bool dyn = w->ray_dynamic && n_rays >= 4096;
	if (dyn) CK(cudaMemsetAsync(&w->W.ctr[CT_RAY_NEXT], 0, sizeof(u32), w->stream));
	ray_swap(w);
	if (dyn) LAUNCH(w, PK_RAYCAST, k_raycast_dyn, std::min(nblk(n_rays, 128), (int)w->ray_blocks), 128, w->W, d_origin, d_look, n_rays, d_winner,
	                (const i32 *)w->ray.rank2slot, (const i32 *)w->ray.slot2rank, (i32)w->stats.live_ents, order, &w->W.ctr[CT_RAY_NEXT]);
	else LAUNCH(w, PK_RAYCAST, k_raycast, nblk(n_rays, 128), 128, w->W, d_origin, d_look, n_rays, d_winner,
	            (const i32 *)w->ray.rank2slot, (const i32 *)w->ray.slot2rank, (i32)w->stats.live_ents, order);
	ray_swap(w);
	return 0;
}

extern "C" {

int bt_raycast_build(bt_world *w) {
	if (!w) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	return ray_build(w);
}

int bt_raycast_batch_device(bt_world *w, const int32_t *d_origin, const int32_t *d_look, int32_t n_rays,
                            int32_t *d_winner) {
	if (!w || n_rays < 0) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	{ int rr = fused_settle(w, false); if (rr) return rr; }
	if (!w->ray_valid) {
		int r = ray_build(w);
		if (r) return r;
	}
	return ray_cast(w, d_origin, d_look, n_rays, d_winner);
}

int bt_raycast_batch(bt_world *w, const int32_t *origin, const int32_t *look, int32_t n_rays, int32_t *winner) {
	if (!w || !origin || !look || !winner || n_rays < 0) return BT_ERR_ARG;
	CK(cudaSetDevice(w->device));
	i32 *d_o = NULL, *d_l = NULL, *d_w = NULL;
	size_t nr = (size_t)(n_rays ? n_rays : 1);
	CK(cudaMalloc(&d_o, sizeof(i32) * nr));
	CK(cudaMalloc(&d_l, sizeof(i32) * 3 * nr));
	CK(cudaMalloc(&d_w, sizeof(i32) * nr));
	int r = BT_OK;
	cudaError_t e = cudaMemcpyAsync(d_o, origin, sizeof(i32) * n_rays, cudaMemcpyHostToDevice, w->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_l, look, sizeof(i32) * 3 * n_rays, cudaMemcpyHostToDevice, w->stream);
	if (e == cudaSuccess) r = bt_raycast_batch_device(w, d_o, d_l, n_rays, d_w);
	if (e == cudaSuccess && r == BT_OK)
		e = cudaMemcpyAsync(winner, d_w, sizeof(i32) * n_rays, cudaMemcpyDeviceToHost, w->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
	cudaFree(d_o); cudaFree(d_l); cudaFree(d_w);
	if (e != cudaSuccess) { set_err("%s failed: %s (bt_world.cu:%d)", "raycast copy", cudaGetErrorString(e), __LINE__); return BT_ERR_CUDA; }
	return r;
}

} /* extern "C" */
