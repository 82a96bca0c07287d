/*
 * bt_kernels.cuh — __global__ wrappers around the per-thread bodies of bt_phys.cuh,
 * plus the hand-written LSD radix sort used by the LBVH build.  sm_100a only.
 *
 * Grid sizing: every kernel is one thread per entity / leaf / node, 128 threads per
 * block (4 warps), so 1 M entities = 7 813 blocks = 52.8 blocks per SM on 148 SMs,
 * i.e. > 3 full waves at 16 resident blocks/SM.  All of these kernels are bound by
 * HBM/L2 latency, not math; nothing here is a dense contraction, so no tensor cores.
 */
#ifndef BT_KERNELS_CUH
#define BT_KERNELS_CUH

#include "bt_phys.cuh"

#define BT_BLOCK 128
#ifndef BT_NARROW_BATCH
#define BT_NARROW_BATCH 2 /* pairs whose cheap filter one k_narrow thread runs at once */
#endif

static __device__ __forceinline__ u32 warp_sum(u32 v) {
	return __reduce_add_sync(0xffffffffu, v);
}

/* add a per-thread count into a 64-bit stat with one atomic per warp, into this block's row (bt_types.h: BT_STAT_ROWS) */
static __device__ __forceinline__ void stat_add(const DevWorld &W, int which, u32 v) {
	u32 s = warp_sum(v);
	if ((threadIdx.x & 31) == 0 && s) atomicAdd(&W.stat[(blockIdx.x & (BT_STAT_ROWS - 1)) * SK_NUM + which], (u64)s);
}
/* "this iteration proposed something": a warp stores only while the word is not stamped yet (read at L2: every warp of
 * the first iteration has something to say, and they all hit the same word).  No block-wide vote: warps that are done
 * leave instead of holding their slots until the block's slowest is (k_narrow -6.6 %, tick 0.586 -> 0.567 ms at 1 M). */
static __device__ __forceinline__ void stamp_any(const DevWorld &W, u32 any, int slot) {
	if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0 && __ldcg(&W.ctr[slot]) != (u32)W.iter_id) W.ctr[slot] = (u32)W.iter_id;
}
static __device__ __forceinline__ void stats_flush(const DevWorld &W, const LocalStats &st) {
	stat_add(W, SK_CAND, st.cand);
	stat_add(W, SK_POS, st.pos);
	stat_add(W, SK_PUSH, st.push);
	stat_add(W, SK_CRUSH, st.crush);
	stat_add(W, SK_FUMBLE, st.fumble);
}

/* Programmatic dependent launch: every kernel of the tick is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization.  First thing, a kernel lets ITS successor be
 * scheduled (its blocks fill SMs as ours drain, hiding launch latency and part of the tail), then
 * waits until its own predecessor has completed and its memory is visible.  Without the launch
 * attribute both instructions are no-ops. */
/* The instructions sit behind a launch-uniform flag: executed unconditionally they cost the 1 M-entity
 * world 2.3 % even without the launch attribute (the asm is a compiler barrier at every kernel entry). */
#define PDL_ENTRY() do { if (W.pdl & 1) asm volatile("griddepcontrol.launch_dependents;\ngriddepcontrol.wait;" ::: "memory"); } while (0)
/* the two long multi-wave kernels (k_narrow, k_push) only wait: their successor is released when they
 * complete, so that no waiting block holds registers while their last waves run */
#define PDL_WAIT() do { if (W.pdl & 2) asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
/* kernels without a DevWorld parameter (radix sort passes of the tree build) */
#define PDL_ENTRY_ALWAYS() asm volatile("griddepcontrol.launch_dependents;\ngriddepcontrol.wait;" ::: "memory")
#define TID ((i32)(blockIdx.x * blockDim.x + threadIdx.x))
/* the proposal stamp of iteration `id` lives in one of two slots, by parity */
#define ANY_SLOT(id) (((id) & 1) ? CT_ANY_ODD : CT_ANY_PROPOSAL)
/* grid-stride loop for the two kernels that usually return at once (k_refit, k_pairs are
 * conditional on CT_CACHE_INVALID): a one-wave grid makes the no-op launch ~3 us instead of
 * ~10 us at 1 M entities.  Measured: the always-busy per-iteration kernels (decide, push,
 * post) gain nothing from it and keep one thread per entity. */
#define GS_FOR(i, n) for (i32 i = TID; i < (n); i += (i32)(gridDim.x * blockDim.x))
#define GS_BLOCK 256

/* flush1's ordered part, one WARP per world.  t_flush_serial (the one-thread restatement the emulator
 * runs) walks the world's entities three times, newest first; almost every visit finds nothing.  Here
 * the 32 lanes look at 32 entities at once and only the hits are handled, one at a time, in the
 * same newest-first order by the lane that found them.  The candidate flags of a pass are not
 * changed by the pass itself (a crush kills only its own entity, a fumble sets no drop flag, a
 * pickup's handlers request no pickup), every other condition is re-evaluated when the hit is
 * handled, so the result equals the sequential walk. */
/* the body: all 32 lanes of ONE warp, world g */
static __device__ __forceinline__ void bt_flush_serial_warp(const DevWorld &W, const i32 g, LocalStats &st) {
	const int lane = threadIdx.x & 31;
	st.rk_ordered = 1; /* fumbles of this launch take their place in gs->rootEnts from the world's counter */
	{
		int work = 0;
		if (lane == 0) {
			W.flip[g] ^= 1;
			work = W.wflush[g];
			W.wflush[g] = 0;
			W.wrootseq[g] = 0;
		}
		work = __shfl_sync(0xffffffffu, work, 0);
		if (work) {
			const int slot = 1 ^ __shfl_sync(0xffffffffu, lane == 0 ? W.flip[g] : 0, 0);
			const i32 b0 = W.wbegin[g], b1 = W.wend[g];
			for (int pass = 0; pass < 2; pass++) {
				for (i32 top = b1 - 1; top >= b0; top -= 32) {
					const i32 i = top - lane; /* lane 0 = newest */
					int hit = 0;
					if (i >= b0 && W.inList[i]) {
						if (pass == 0) hit = (slot ? W.dead_max1[i] : W.dead_max0[i]) != 0;
						else hit = W.newDrop[i] != 0;
					}
					u32 m = __ballot_sync(0xffffffffu, hit);
					while (m) {
						const int l = __ffs(m) - 1;
						m &= m - 1;
						if (lane == l && W.inList[i]) {
							if (pass == 0) { /* flushDeaths ent.cpp:747-761 */
								u8 d = slot ? W.dead_max1[i] : W.dead_max0[i];
								bt_crush(W, i, 0xfffffu - ((u32)i & 0xfffffu), st, d);
								W.dead[i] = d;
							} else { /* flushDrops :765-773 */
								W.newDrop[i] = 0;
								bt_fumble(W, i, st);
							}
						}
						__syncwarp();
					}
				}
				__syncwarp();
			}
			/* flushPickups ent.cpp:775-792 walks gs->rootEnts head first = descending rootKey: the lanes
			 * look for the requesting root with the largest key, lane 0 serves it, until none is left
			 * (normally the first look finds none) */
			while (1) {
				u64 bk = 0;
				i32 bi = -1;
				for (i32 i = b0 + lane; i < b1; i += 32) {
					if (!W.inList[i] || W.holder[i] >= 0 || W.newHolder[i] < 0) continue;
					const u64 k = W.rootKey[i];
					if (bi < 0 || k > bk) { bk = k; bi = i; }
				}
#pragma unroll
				for (int off = 16; off > 0; off >>= 1) {
					const u64 ok = __shfl_xor_sync(0xffffffffu, bk, off);
					const i32 oi = __shfl_xor_sync(0xffffffffu, bi, off);
					if (oi >= 0 && (bi < 0 || ok > bk || (ok == bk && oi > bi))) { bk = ok; bi = oi; }
				}
				if (bi < 0) break;
				if (lane == 0) bt_flush_pickup_one(W, bi);
				__syncwarp();
			}
		}
	}
	st.rk_ordered = 0;
}
/* Block-level re-packing: every thread offers up to `per` items; survivors are appended to a
 * shared-memory queue with one shared atomic per warp (ballot + popc), so that the expensive
 * second step runs on full warps instead of on the 1-in-4 lanes that had work. */
static __device__ __forceinline__ void queue_push(u32 *q, u32 *count, int pred, u32 item) {
	u32 m = __ballot_sync(0xffffffffu, pred);
	if (!m) return;
	int lane = threadIdx.x & 31;
	u32 base = 0;
	if (lane == 0) base = atomicAdd(count, (u32)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (pred) q[base + __popc(m & ((1u << lane) - 1))] = item;
}

#ifndef BT_KERNEL_HELPERS_ONLY /* bt_small.cu takes the helpers above and defines its own kernel */

__global__ void __launch_bounds__(BT_BLOCK) k_pre(DevWorld W, u32 phases) { PDL_ENTRY(); t_pre(W, TID, phases); }
__global__ void __launch_bounds__(BT_BLOCK) k_flush_scan(DevWorld W) { PDL_ENTRY(); t_flush_scan(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_flush_serial(DevWorld W) { PDL_ENTRY();
	LocalStats st = {0, 0, 0, 0, 0};
	const i32 g = TID >> 5;
	if (g < W.n_worlds) bt_flush_serial_warp(W, g, st);
	stats_flush(W, st);
}
__global__ void __launch_bounds__(BT_BLOCK) k_flush_misc(DevWorld W) { PDL_ENTRY(); t_flush_misc(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_tick(DevWorld W) { PDL_ENTRY(); t_tick(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_tick_serial(DevWorld W) { PDL_ENTRY(); t_tick_serial(W, TID); }
/* requests arrive grouped by world (stable), one thread per world applies its run in order */
__global__ void __launch_bounds__(BT_BLOCK) k_requests(DevWorld W, const BtReq *q, const i32 *starts, i32 n_groups) { PDL_ENTRY();
	i32 g = TID;
	if (g >= n_groups) return;
	for (i32 i = starts[g]; i < starts[g + 1]; i++) t_request(W, q[i]);
}
__global__ void __launch_bounds__(BT_BLOCK) k_gravity(DevWorld W) { PDL_ENTRY(); t_gravity(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_world_rules(DevWorld W) { PDL_ENTRY(); t_world_rules(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_flush2_apply(DevWorld W) { PDL_ENTRY(); t_flush2_apply(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_zero_deltas(DevWorld W) { PDL_ENTRY(); t_zero_deltas(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_integrate(DevWorld W) { PDL_ENTRY(); t_integrate(W, TID); }

__global__ void __launch_bounds__(BT_BLOCK) k_begin_fused(DevWorld W, int gravity, int ctrls) { PDL_ENTRY(); t_begin_fused(W, TID, gravity, ctrls); }
__global__ void __launch_bounds__(BT_BLOCK) k_flip_only(DevWorld W) { PDL_ENTRY(); t_flip_only(W, TID); }

/* bounds of the leaf centres + leaf count: per-thread partials over a grid-stride loop, warp shuffles, then ONE set
 * of atomics per block (one per warp of a one-thread-per-entity grid made 31 000 warps queue on the same seven
 * words: 145 us at 1 M entities; this is a 4 MB read) */
__global__ void __launch_bounds__(GS_BLOCK) k_bounds(DevWorld W) { PDL_ENTRY();
	__shared__ i32 part[7][GS_BLOCK / 32];
	i32 mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
	u32 cnt = 0;
	GS_FOR(e, W.n) {
		if (!bt_is_leaf_ent(W, e)) continue;
		const int4 c = W.center[e];
		mn[0] = min(mn[0], c.x); mn[1] = min(mn[1], c.y); mn[2] = min(mn[2], c.z);
		mx[0] = max(mx[0], c.x); mx[1] = max(mx[1], c.y); mx[2] = max(mx[2], c.z);
		cnt++;
	}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
		mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
	}
	cnt = warp_sum(cnt);
	if (lane == 0) {
		for (int k = 0; k < 3; k++) { part[k][warp] = mn[k]; part[3 + k][warp] = mx[k]; }
		part[6][warp] = (i32)cnt;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		u32 total = 0;
		for (int w2 = 0; w2 < GS_BLOCK / 32; w2++) {
			for (int k = 0; k < 3; k++) { mn[k] = min(mn[k], part[k][w2]); mx[k] = max(mx[k], part[3 + k][w2]); }
			total += (u32)part[6][w2];
		}
		if (total) {
			i32 *b = (i32 *)&W.ctr[CT_BOUNDS];
			atomicMin(&b[0], mn[0]); atomicMin(&b[1], mn[1]); atomicMin(&b[2], mn[2]);
			atomicMax(&b[3], mx[0]); atomicMax(&b[4], mx[1]); atomicMax(&b[5], mx[2]);
			atomicAdd(&W.ctr[CT_N_LEAVES], total);
		}
	}
}
__global__ void __launch_bounds__(BT_BLOCK) k_morton(DevWorld W) { PDL_ENTRY(); t_morton(W, TID); }
__global__ void __launch_bounds__(BT_BLOCK) k_lbvh(DevWorld W) { PDL_ENTRY(); t_lbvh_node(W, TID); }
__global__ void __launch_bounds__(GS_BLOCK) k_refit(DevWorld W, int cond) { PDL_ENTRY();
	if (cond && !W.ctr[CT_CACHE_INVALID]) return;
	if (cond && W.small_mode) { /* no tree to refit: only the pair list is reset, k_pairs_small does the rest */
		if (TID == 0) { W.ctr[CT_PAIR_N] = 0; atomicAdd(&W.ctr[CT_PAIR_BUILDS], 1u); }
		return;
	}
	if (cond && bt_patch_mode(W)) {
		GS_FOR(t, (i32)W.ctr[CT_DIRTY_N]) t_refit_patch(W, t);
		return;
	}
	GS_FOR(t, W.n_leaves) t_refit(W, t, cond);
}

/* traversal stacks live in shared memory, interleaved so that lane i's slot k is at
 * word k*128+i: the 32 lanes of a warp hit 32 different banks (no conflicts). */
__global__ void __launch_bounds__(BT_BLOCK) k_collide(DevWorld W) { PDL_ENTRY();
	__shared__ i32 stacks[BT_STACK * BT_BLOCK];
	LocalStats st = {0, 0, 0, 0, 0};
	u32 any = 0;
	t_collide(W, TID, stacks + threadIdx.x, BT_BLOCK, st, any);
	stats_flush(W, st);
	stamp_any(W, any, ANY_SLOT(W.iter_id));
}
/* Pair emitter: hits are buffered per thread and written with ONE atomic per warp flush
 * (warp-ballot/shuffle compaction) instead of one atomic per pair. */
#define PAIR_BUF 8
struct PairEmit {
	u64 *pairs;
	u32 *counter;
	u32 cap;
	u64 *buf; /* this thread's slots in shared memory, stride BT_BLOCK */
	int n;
	__device__ __forceinline__ PairEmit(u64 *p, u32 *c, i32 cap_, u64 *sbuf)
	    : pairs(p), counter(c), cap((u32)cap_), buf(sbuf), n(0) {}
	__device__ __forceinline__ void store(u32 pos, u64 pr) {
		if (pos < cap) pairs[pos] = pr;
	}
	/* called from divergent code when this thread's buffer is full */
	__device__ __forceinline__ void operator()(i32 a, i32 b) {
		buf[(n++) * BT_BLOCK] = (u64)(u32)a | ((u64)(u32)b << 32);
		if (n == PAIR_BUF) {
			u32 m = __activemask();
			int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
			u32 rank = __popc(m & ((1u << lane) - 1));
			u32 base = 0;
			if (lane == leader) base = atomicAdd(counter, (u32)(PAIR_BUF * __popc(m)));
			base = __shfl_sync(m, base, leader) + rank * PAIR_BUF;
#pragma unroll
			for (int k = 0; k < PAIR_BUF; k++) store(base + k, buf[k * BT_BLOCK]);
			n = 0;
		}
	}
	/* called by every thread of the BLOCK, converged: what the threads hold goes out in thread order behind ONE atomic */
	__device__ __forceinline__ void finish_block(u32 *wsum, u32 *base_sh) {
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		u32 x = (u32)n;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, x, off);
			if (lane >= off) x += y;
		}
		if (lane == 31) wsum[warp] = x;
		__syncthreads();
		u32 before = 0, tot = 0;
		for (int k = 0; k < (int)(blockDim.x >> 5); k++) { if (k < warp) before += wsum[k]; tot += wsum[k]; }
		if (threadIdx.x == 0 && tot) *base_sh = atomicAdd(counter, tot);
		__syncthreads();
		if (tot) {
			const u32 base = *base_sh + before + x - (u32)n;
			for (int k = 0; k < n; k++) store(base + k, buf[k * BT_BLOCK]);
		}
		n = 0;
		__syncthreads(); /* wsum / base_sh are reused by the next round */
	}
	/* called once by every thread of the warp, converged */
	__device__ __forceinline__ void finish() {
		int lane = threadIdx.x & 31;
		u32 x = (u32)n;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, x, off);
			if (lane >= off) x += y;
		}
		u32 tot = __shfl_sync(0xffffffffu, x, 31);
		u32 base = 0;
		if (lane == 31 && tot) base = atomicAdd(counter, tot);
		base = __shfl_sync(0xffffffffu, base, 31) + x - (u32)n;
		for (int k = 0; k < n; k++) store(base + k, buf[k * BT_BLOCK]);
	}
};

__global__ void __launch_bounds__(BT_BLOCK) k_pairs(DevWorld W) { PDL_ENTRY();
	__shared__ i32 stacks[BT_STACK * BT_BLOCK];
	__shared__ u64 pbuf[PAIR_BUF * BT_BLOCK];
	if (!W.ctr[CT_CACHE_INVALID]) return; /* uniform: the cached pairs are still a superset */
	PairEmit emit(W.pairs, &W.ctr[CT_PAIR_N], W.pair_cap, pbuf + threadIdx.x);
	if (bt_patch_mode(W)) { /* a few moved entities: spread them over the whole grid */
		GS_FOR(t, (i32)W.ctr[CT_DIRTY_N]) t_pairs_patch(W, t, stacks + threadIdx.x, BT_BLOCK, emit);
		emit.finish();
		return;
	}
	/* blocked distribution: each block walks one contiguous run of leaves, 128 consecutive leaves per round, and the
	 * pairs of a round go out together in leaf order (one atomic per block and round): the list is sorted by the
	 * position of the emitting leaf down to the grain of a round, and the 32 pairs a warp of k_narrow reads come from
	 * ~32 neighbouring leaves.  (With every thread flushing its own buffer at the end of its whole run, 8 consecutive
	 * list entries came from one thread, i.e. from leaves 128 apart: k_narrow, whose gathers live on that locality, went
	 * 306 -> 270 us per tick on the settled 1 M world and 459 -> 357 in the dynamic phase with the ordered list; handing
	 * the leaves out dynamically like the rays of k_raycast_dyn made the walk itself no faster, 410 -> 434 us, and the
	 * list worse.  profiles/r2_ab_experiments.md §10) */
	__shared__ u32 wsum[BT_BLOCK / 32];
	__shared__ u32 rbase;
	const i32 chunk = (W.n_leaves + (i32)gridDim.x - 1) / (i32)gridDim.x;
	const i32 t0 = (i32)blockIdx.x * chunk, t1 = t0 + chunk < W.n_leaves ? t0 + chunk : W.n_leaves;
	for (i32 r0 = t0; r0 < t1; r0 += BT_BLOCK) {
		const i32 t = r0 + (i32)threadIdx.x;
		if (t < t1) t_pairs(W, t, stacks + threadIdx.x, BT_BLOCK, emit);
		emit.finish_block(wsum, &rbase);
	}
}

/* small-world mode: one block per world, boxes in shared memory, all pairs (see t_pairs_small_world) */
__global__ void __launch_bounds__(BT_SMALL_WORLD) k_pairs_small(DevWorld W) { PDL_ENTRY();
	__shared__ int4 blo[BT_SMALL_WORLD], bhi[BT_SMALL_WORLD];
	__shared__ u64 pbuf[PAIR_BUF * BT_SMALL_WORLD];
	if (!W.ctr[CT_CACHE_INVALID]) return;
	const int g = (int)blockIdx.x;
	const i32 b0 = W.wbegin[g], cnt = W.wend[g] - b0;
	const int i = (int)threadIdx.x;
	int leaf = 0;
	Box6 q;
	q.lo = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, 0); /* empty: touches nothing */
	q.hi = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, 0);
	if (i < cnt && bt_is_leaf_ent(W, b0 + i)) {
		leaf = 1;
		q = bt_inflated_box(W, b0 + i);
		W.ibox[b0 + i] = q;
	}
	blo[i] = q.lo; bhi[i] = q.hi;
	__syncthreads();
	PairEmit emit(W.pairs, &W.ctr[CT_PAIR_N], W.pair_cap, pbuf + threadIdx.x);
	if (leaf) {
		/* balanced triangle: entity i looks at the next (cnt-1)/2 entities cyclically, and at the
		 * one opposite when cnt is even and i is in the lower half: every unordered pair once */
		const int half = (cnt - 1) / 2;
		for (int k = 1; k <= half + ((cnt & 1) == 0 && i < cnt / 2 ? 1 : 0); k++) {
			int j = i + k;
			if (j >= cnt) j -= cnt;
			Box6 o;
			o.lo = blo[j]; o.hi = bhi[j];
			if (bt_box_touch(q, o)) emit(b0 + (i < j ? i : j), b0 + (i < j ? j : i));
		}
	}
	emit.finish();
}

/* one thread per cached pair, grid-stride so the launch does not need the pair count.
 * (A two-step variant that re-packs the survivors of the cheap tests into full warps was
 * measured 20 % slower here: the step is bound by the latency of its gathers, which the
 * re-packing doubles, not by divergent ALU work.  Collecting the survivors of a whole chunk of
 * trips -- 4 / 8 / 16 -- in a shared-memory queue first and running the narrow phase over the
 * queue on full warps was measured too: +1 % tick time at 1 M entities, so the cost follows the
 * number of scanning pairs, not the number of warps that carry them.  k_push below does gain
 * from re-packing.) */
/* spec = 1: an iteration enqueued ahead of the host's knowledge (run-ahead).  It proceeds only
 * if the previous iteration proposed something and no push has invalidated the pair cache;
 * otherwise it records its id in CT_STOP and it and everything enqueued behind it stand down. */
#ifndef BT_NARROW_MINBLOCKS
#define BT_NARROW_MINBLOCKS 3 /* resident 256-thread blocks per SM the register allocation is held to: 80 registers,
                               * 20 bytes of spills, 24 warps per SM instead of 16 (101 registers uncapped): k_narrow -7 %,
                               * 4 blocks (64 registers, 324 bytes of spills) +5 % (profiles/r2_ab_experiments.md) */
#endif
__global__ void __launch_bounds__(256, BT_NARROW_MINBLOCKS) k_narrow(DevWorld W, int spec) { PDL_WAIT();
	LocalStats st = {0, 0, 0, 0, 0};
	u32 any = 0;
	if (spec & 1) {
		/* none of these words is written by this kernel while `go` holds (the stamp of THIS
		 * iteration goes to the other parity slot), so early and late blocks agree */
		int go = W.ctr[CT_STOP] == 0 && W.ctr[ANY_SLOT(W.iter_id - 1)] == (u32)(W.iter_id - 1) && !W.ctr[CT_CACHE_INVALID] &&
		         !W.ctr[CT_PAIR_OVERFLOW];
		if (!go) {
			if (blockIdx.x == 0 && threadIdx.x == 0 && W.ctr[CT_STOP] == 0) W.ctr[CT_STOP] = (u32)W.iter_id;
			return;
		}
	}
	u32 np = W.ctr[CT_PAIR_N];
	if (np > (u32)W.pair_cap) { /* incomplete pair list: nobody may act on it */
		if (blockIdx.x == 0 && threadIdx.x == 0) W.ctr[CT_PAIR_OVERFLOW] = 1;
		return;
	}
	/* k_refit / k_pairs of this iteration have consumed the flag; pushes may raise it again
	 * (a run-ahead iteration only gets here with the flag down, and other blocks are reading it) */
	if (!(spec & 1) && blockIdx.x == 0 && threadIdx.x == 0) bt_consume_cache_flags(W, np);
	const u32 stride = gridDim.x * blockDim.x;
	u32 p = blockIdx.x * blockDim.x + threadIdx.x;
#ifndef BT_NO_NARROW_FIRST
	if (spec & 2) { /* first iteration of the tick: everybody scans, no flag gathers (see t_narrow_t) */
		for (; p < np; p += stride) t_narrow_t<1>(W, W.pairs[p], st, any);
		stats_flush(W, st);
		stamp_any(W, any, ANY_SLOT(W.iter_id));
		return;
	}
#endif
#if BT_NARROW_BATCH > 1
	/* The cheap filter (pair word -> needsCollision bytes of both ends: two dependent memory hops
	 * that most pairs end with after the first iteration) runs on BT_NARROW_BATCH pairs at once, so
	 * the hops of different pairs overlap; the survivors then go through the narrow phase proper, one
	 * inlined copy picked by a bit mask.  Proposals are CAS-maxima and the statistics are sums, so
	 * the order in which a thread visits its pairs is not observable.  Measured on the 1 M world
	 * (same box, A/B): batch 2 = -2.5 % tick time (k_narrow -5 %), 4 = -1.5 %, 8 = 0 (the heavy
	 * path of a thread's survivors is serial, so wide batches lengthen its chain). */
	for (; p + (BT_NARROW_BATCH - 1) * stride < np && p + (BT_NARROW_BATCH - 1) * stride >= p; p += BT_NARROW_BATCH * stride) {
		u64 q[BT_NARROW_BATCH];
		u32 m = 0;
#pragma unroll
		for (int k = 0; k < BT_NARROW_BATCH; k++) q[k] = W.pairs[p + k * stride];
#pragma unroll
		for (int k = 0; k < BT_NARROW_BATCH; k++)
			m |= (u32)((W.needsCollision[(u32)q[k]] | W.needsCollision[(u32)(q[k] >> 32)]) != 0) << k;
		while (m) {
			int k = __ffs(m) - 1;
			m &= m - 1;
			u64 qq = q[0];
#pragma unroll
			for (int j = 1; j < BT_NARROW_BATCH; j++) qq = k == j ? q[j] : qq;
			t_narrow(W, qq, st, any);
		}
	}
#endif
	for (; p < np; p += stride) t_narrow(W, W.pairs[p], st, any);
	stats_flush(W, st);
	stamp_any(W, any, ANY_SLOT(W.iter_id));
}

/* The resolve kernels of an iteration are enqueued right behind its narrow phase, before the
 * host knows whether anything was proposed (one host sync per iteration instead of two).  With
 * no proposal they are no-ops by construction; if the pair buffer overflowed they stand down
 * (CT_PAIR_OVERFLOW) and the host redoes the iteration through k_collide. */
__global__ void __launch_bounds__(BT_BLOCK) k_decide(DevWorld W) { PDL_ENTRY();
	if (W.ctr[CT_PAIR_OVERFLOW] | W.ctr[CT_STOP] | W.ctr[CT_ERROR]) return; /* a raised error (traversal stack, work list) stops the iteration: nobody acts on a partial candidate set */
	GS_FOR(e, W.n) t_decide(W, e); /* 148 x 16 blocks, grid-stride: 53 blocks of 128 threads per SM cost more to schedule than to run */
}
/* ready roots (typically 1 entity in 4) are re-packed into full warps before the push.  (Letting a block look at 256 /
 * 512 / 1024 entities so that every one of its warps has roots to push -- at one entity per thread one warp in four
 * does -- was measured: k_push 228 -> 232 / 237 / 229 us per tick at 1 M entities.  The kernel moves ~150 MB of 32-byte
 * sectors per launch for ~50 MB it uses, 3.3 TB/s: more chains in flight do not help it.) */
__global__ void __launch_bounds__(BT_BLOCK) k_push(DevWorld W) { PDL_WAIT();
	__shared__ u32 queue[BT_BLOCK];
	__shared__ u32 qn;
	LocalStats st = {0, 0, 0, 0, 0};
	if (W.ctr[CT_PAIR_OVERFLOW] | W.ctr[CT_STOP] | W.ctr[CT_ERROR]) return; /* a raised error (traversal stack, work list) stops the iteration: nobody acts on a partial candidate set */
	if (threadIdx.x == 0) qn = 0;
	__syncthreads();
	i32 e = TID;
	queue_push(queue, &qn, e < W.n && W.ready[e] != 0, (u32)e);
	__syncthreads();
	if (threadIdx.x < qn) t_push(W, (i32)queue[threadIdx.x], st);
	stats_flush(W, st);
}
__global__ void k_push_serial(DevWorld W) { PDL_ENTRY(); PDL_WAIT(); /* launched with k_push's attribute (class push): with option pdl = 2 only the second applies */
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (W.ctr[CT_SERIAL_N] == 0 || W.ctr[CT_PAIR_OVERFLOW] || W.ctr[CT_STOP] || W.ctr[CT_ERROR]) return;
	LocalStats st = {0, 0, 0, 0, 0};
	t_push_serial(W, st);
	u64 *s = W.stat; /* row 0, shared with other kernels' blocks */
	atomicAdd(&s[SK_PUSH], (u64)st.push); atomicAdd(&s[SK_SERIAL], (u64)st.push); atomicAdd(&s[SK_CRUSH], (u64)st.crush); atomicAdd(&s[SK_FUMBLE], (u64)st.fumble);
}
__global__ void __launch_bounds__(BT_BLOCK) k_post(DevWorld W) { PDL_ENTRY();
	if (W.ctr[CT_PAIR_OVERFLOW] | W.ctr[CT_STOP] | W.ctr[CT_ERROR]) return; /* a raised error (traversal stack, work list) stops the iteration: nobody acts on a partial candidate set */
	GS_FOR(e, W.n) t_post(W, e);
}

/* host exchange: staging buffers hold [3*n_live] ints in live oldest-first order; while no
 * entity has died, live rank == slot (the common case, checked by the host) */
__global__ void __launch_bounds__(BT_BLOCK) k_add_deltas(DevWorld W, const i32 *dc, const i32 *dv, const i32 *slot2rank) { PDL_ENTRY();
	i32 e = TID;
	if (e >= W.n || !W.inList[e]) return;
	i32 r = slot2rank ? slot2rank[e] : e;
	int4 a = W.d_center[e], b = W.d_vel[e];
	a.x += dc[3 * r]; a.y += dc[3 * r + 1]; a.z += dc[3 * r + 2];
	b.x += dv[3 * r]; b.y += dv[3 * r + 1]; b.z += dv[3 * r + 2];
	W.d_center[e] = a;
	W.d_vel[e] = b;
}
__global__ void __launch_bounds__(BT_BLOCK) k_pack_state(DevWorld W, i32 *oc, i32 *ov, const i32 *slot2rank) { PDL_ENTRY();
	i32 e = TID;
	if (e >= W.n || !W.inList[e]) return;
	i32 r = slot2rank ? slot2rank[e] : e;
	int4 c = W.center[e], v = W.vel[e];
	oc[3 * r] = c.x; oc[3 * r + 1] = c.y; oc[3 * r + 2] = c.z;
	ov[3 * r] = v.x; ov[3 * r + 1] = v.y; ov[3 * r + 2] = v.z;
}

__global__ void __launch_bounds__(BT_BLOCK) k_poke_raw(DevWorld W, DevArrays A, const bt_poke *pk, i32 n) { PDL_ENTRY(); t_poke_raw(W, A, pk, TID, n); }
__global__ void k_poke_ops(DevWorld W, const bt_poke *ops, i32 n) { PDL_ENTRY();
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	for (i32 k = 0; k < n; k++) t_poke_op(W, ops[k]);
}
/* bt_world_grow: blank slots [first, first + cnt) */
__global__ void __launch_bounds__(BT_BLOCK) k_grow_init(DevWorld W, i32 first, i32 cnt, i32 world) { PDL_ENTRY();
	if (TID >= cnt) return;
	t_grow_init(W, first + TID, world);
}

/* per-world commutative hash: warp-level pre-reduction when the whole warp is one world */
__global__ void __launch_bounds__(BT_BLOCK) k_hash(DevWorld W, u64 *sum, u32 *cnt, int full) { PDL_ENTRY();
	i32 e = TID;
	int live = e < W.n && W.inList[e];
	u64 h = live ? (full ? bt_ent_hash_full(W, e) : bt_ent_hash(W, e)) : 0;
	i32 g = live ? W.world[e] : -1;
	i32 g0 = __shfl_sync(0xffffffffu, g, 0);
	int uniform = __all_sync(0xffffffffu, g == g0 || g < 0);
	if (uniform) {
		i32 gm = __reduce_max_sync(0xffffffffu, g);
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) h += __shfl_down_sync(0xffffffffu, h, off);
		u32 c = warp_sum((u32)live);
		if ((threadIdx.x & 31) == 0 && gm >= 0) {
			atomicAdd(&sum[gm], h);
			atomicAdd(&cnt[gm], c);
		}
	} else if (live) {
		atomicAdd(&sum[g], h);
		atomicAdd(&cnt[g], 1u);
	}
}

/* the same sum for a device world that holds ONE world: per-thread partial sums over a grid-stride loop, then one pair of
 * atomics per block (k_hash's one per warp all land on the same two words: 31 000 of them at 1 M entities) */
__global__ void __launch_bounds__(GS_BLOCK) k_hash_one(DevWorld W, u64 *sum, u32 *cnt, int full) { PDL_ENTRY();
	__shared__ u64 ph[GS_BLOCK / 32];
	__shared__ u32 pc[GS_BLOCK / 32];
	u64 h = 0;
	u32 c = 0;
	GS_FOR(e, W.n) {
		if (!W.inList[e]) continue;
		h += full ? bt_ent_hash_full(W, e) : bt_ent_hash(W, e);
		c++;
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) h += __shfl_down_sync(0xffffffffu, h, off);
	c = warp_sum(c);
	if ((threadIdx.x & 31) == 0) { ph[threadIdx.x >> 5] = h; pc[threadIdx.x >> 5] = c; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int k = 1; k < GS_BLOCK / 32; k++) { h += ph[k]; c += pc[k]; }
		if (c) { atomicAdd(&sum[0], h); atomicAdd(&cnt[0], c); }
	}
}

/* A few words for the host, written straight into mapped pinned host memory (zero copy).  The counters the host reads
 * once per tick used to come back with cudaMemcpyAsync: that is a job for the device-to-host COPY ENGINE, which a host
 * streaming a 24 MB state read-back on another stream (bt_world_read_state_async) keeps busy for 0.5 ms at a time: the
 * 256 bytes queued behind it and the tick with them. */
__global__ void __launch_bounds__(64) k_publish(const u32 *src, u32 *host_mapped, int n) { PDL_ENTRY_ALWAYS();
	for (int i = threadIdx.x; i < n; i += 64) host_mapped[i] = src[i];
	/* no fence: the host reads after cudaStreamSynchronize, i.e. after the kernel has completed and its writes with it */
}

/* a single world's hash sum and entity count, the same way */
__global__ void k_publish_hash(const u64 *sum, const u32 *cnt, u32 *host_mapped) { PDL_ENTRY_ALWAYS();
	if (threadIdx.x == 0) {
		const u64 s = sum[0];
		host_mapped[0] = (u32)s; host_mapped[1] = (u32)(s >> 32); host_mapped[2] = cnt[0];
	}
}

/* --------------------------------------------------------------------------
 * LSD radix sort, 8 bits per pass, keys u64 + payload i32, stable.
 * pass = histogram -> scan -> scatter.  Tile = 256 threads x 16 keys.
 * -------------------------------------------------------------------------- */
#define SORT_THREADS 256
#define SORT_ITEMS 16
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const u64 *keys, i32 n, int shift, u32 *hist, int nb) { PDL_ENTRY_ALWAYS();
	__shared__ u32 h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	i32 base = blockIdx.x * SORT_TILE;
#pragma unroll 4
	for (int k = 0; k < SORT_ITEMS; k++) {
		i32 i = base + k * SORT_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
	}
	__syncthreads();
	hist[threadIdx.x * nb + blockIdx.x] = h[threadIdx.x]; /* digit-major */
}

/* The counters (256 digits x nb tiles, digit-major) become positions in two steps.  Here one WARP per digit scans its
 * row of nb tile counts in place (exclusive, inside the digit) and leaves the row total in hist[256 * nb + d]; the
 * scatter kernel turns the 256 totals into digit bases itself (a 256-element scan per block).  (One block of 1024
 * threads scanning all 256 * nb counters took 54 us per radix pass at 1 M keys, 4 passes per tree build.) */
__global__ void __launch_bounds__(1024) k_sort_scan(u32 *hist, int nb) { PDL_ENTRY_ALWAYS();
	const int d = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (d >= 256) return;
	u32 *row = hist + (size_t)d * nb;
	u32 carry = 0;
	for (int b0 = 0; b0 < nb; b0 += 32) {
		const int i = b0 + lane;
		const u32 v = i < nb ? row[i] : 0u;
		u32 x = v;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, x, off);
			if (lane >= off) x += y;
		}
		if (i < nb) row[i] = carry + x - v;
		carry += __shfl_sync(0xffffffffu, x, 31);
	}
	if (lane == 0) hist[(size_t)256 * nb + d] = carry;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const u64 *keys, const i32 *vals, u64 *okeys, i32 *ovals,
                                                             i32 n, int shift, const u32 *hist, int nb) { PDL_ENTRY_ALWAYS();
	__shared__ u32 warpOff[SORT_THREADS / 32][256];
	__shared__ u32 dbase[256], wtot[SORT_THREADS / 32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	{ /* digit bases = exclusive scan of the 256 row totals k_sort_scan left behind the table (SORT_THREADS == 256) */
		const u32 v = hist[(size_t)256 * nb + threadIdx.x];
		u32 x = v;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, x, off);
			if (lane >= off) x += y;
		}
		if (lane == 31) wtot[warp] = x;
		__syncthreads();
		u32 before = 0;
		for (int k = 0; k < warp; k++) before += wtot[k];
		dbase[threadIdx.x] = before + x - v;
	}
	for (int d = lane; d < 256; d += 32) warpOff[warp][d] = 0;
	__syncwarp();
	/* each warp owns a contiguous run of 32*SORT_ITEMS keys -> stability */
	i32 wbase = blockIdx.x * SORT_TILE + warp * (32 * SORT_ITEMS);
	u64 kreg[SORT_ITEMS];
#pragma unroll
	for (int k = 0; k < SORT_ITEMS; k++) {
		i32 i = wbase + k * 32 + lane;
		kreg[k] = i < n ? keys[i] : 0;
		u32 d = i < n ? (u32)((kreg[k] >> shift) & 0xff) : 0xffffffffu;
		u32 peers = __match_any_sync(0xffffffffu, d);
		if (d != 0xffffffffu && (peers & ((1u << lane) - 1)) == 0) warpOff[warp][d] += __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	{ /* thread d turns the per-warp counts into running offsets */
		int d = threadIdx.x;
		u32 run = dbase[d] + hist[d * nb + blockIdx.x];
#pragma unroll
		for (int w = 0; w < SORT_THREADS / 32; w++) {
			u32 c = warpOff[w][d];
			warpOff[w][d] = run;
			run += c;
		}
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < SORT_ITEMS; k++) {
		i32 i = wbase + k * 32 + lane;
		u32 d = i < n ? (u32)((kreg[k] >> shift) & 0xff) : 0xffffffffu;
		u32 peers = __match_any_sync(0xffffffffu, d);
		u32 rank = __popc(peers & ((1u << lane) - 1));
		if (d != 0xffffffffu) {
			u32 pos = warpOff[warp][d] + rank;
			okeys[pos] = kreg[k];
			ovals[pos] = vals[i];
		}
		__syncwarp();
		if (d != 0xffffffffu && rank == 0) warpOff[warp][d] += __popc(peers);
		__syncwarp();
	}
}

#endif /* BT_KERNEL_HELPERS_ONLY */

#endif
