/*
 * bt_small.cu — whole steps of SMALL worlds in one launch: one thread block per world.
 *
 * The batched case (BASELINE.json configs[3]: 8192 x arena.sav, <= 128 entities each) spends the generic schedule's
 * tick in ~40 launches over all worlds plus one host round trip per resolution batch, although nothing ever crosses
 * a world boundary.  Here block g runs doWholeStep (main.cpp:1019-1097) for world g from the first flush to
 * clearDeads with __syncthreads() where the generic schedule has a kernel boundary:
 *
 *   thread i          = entity wbegin[g] + i (the per-thread bodies of bt_phys.cuh, unchanged)
 *   loop control      = the CT_* words in SHARED memory (W.ctr is pointed at them), one set per world: the
 *                       iteration count, cache invalidation, serial lane and error state of a world are its own
 *   candidate pairs   = all pairs of the world's inflated boxes, from shared memory, rebuilt when the tick begins and
 *                       whenever a push leaves its box (bt_check_cover); the list lives in shared memory too
 *   serial lane, dead list, events, pending-wire pool = shared-memory lists; deaths and events are appended to the
 *                       world-wide lists (one atomic per block that has any) when the step ends
 *   entity state      = the arrays the resolution loop lives on (SW_STAGED below: geometry, masks, hold tree, pending
 *                       collision, flags; 165 B per entity) are copied into shared memory when the step begins and
 *                       back when it ends; in between W.center etc. point INTO shared memory, offset by -wbegin[g] so
 *                       that the bodies keep indexing them by slot: a world's gathers cost a shared-memory access
 *                       instead of an L2 / HBM round trip per hop
 *
 * Results are those of the generic schedule bit for bit (the proposals are CAS-maxima, pushes of one iteration own
 * their trees, the serial lane runs in canonical order: none of it depends on which worlds share a launch).
 * The host sees none of it until it asks (fused_settle in bt_world.cu).
 *
 * Its own translation unit: every per-thread body is inlined into this one kernel (W is edited in registers, so
 * nothing may take its address), which makes it by far the longest compile of the library.  It is compiled four times,
 * side by side (bittoss_b200/capi.py: build): block size SW_BS = 64 or 128 threads (the largest world of the upload)
 * x SW_MIN_WARPS = 16 or 32 resident warps per SM the registers are held to.  Measured on 8192 / 1024 arena.sav worlds
 * (64 entities, same box, ms per tick): 16 warps (128 registers) 0.266 / 0.078, 24 warps (80 registers, 216 B of
 * spills) 0.250 / 0.083, 32 warps (64 registers, 368 B) 0.241 / 0.101, 40 warps (48 registers) 0.288 / 0.110: when
 * there are enough worlds to fill the SMs the step is bound by latency and more warps pay for their spills, when
 * there are not only the spills are left.  bt_launch_tick_small picks by the number of worlds.
 */
#define BT_KERNEL_HELPERS_ONLY
#define BT_NO_PREFETCH
#include "bt_kernels.cuh"
#include "bt_small.h"

#define SW_PAIRS_PER_ENT 12 /* shared pair list: entries per thread; beyond it the narrow phase walks the box triangle itself */
#define SW_EVENTS 64        /* host events (TEAMSELECT / LOADBLOCK pushes, crush handlers) one world can raise per step */

#if !defined(SW_BS) || !defined(SW_MIN_WARPS)
#error "compile with -DSW_BS=64|128 -DSW_MIN_WARPS=16|32 (bittoss_b200/capi.py: build)"
#endif
#define BS SW_BS
/* Two regimes (measured on arena.sav worlds, ms per tick once the kicks have died down, unstaged at 32 warps / staged
 * at 16): 8192 worlds 0.234 / 0.286, 1024 worlds 0.081 / 0.062.  With enough worlds to fill every SM the other
 * blocks hide a world's memory latency and what counts is how many fit (the staged arrays take 10 KB per 64 entities,
 * the copies in and out triple the traffic); with fewer, a world's chain of dependent gathers is the step time and
 * shared memory cuts every hop.  So the 16-warp variant stages, the 32-warp variant does not. */
#define SW_STAGE (SW_MIN_WARPS == 16)
#define SW_CAT3(a, b, c) a##b##_##c
#define SW_NAME(prefix, bs, w) SW_CAT3(prefix, bs, w)
#define k_tick_small SW_NAME(k_tick_small_, SW_BS, SW_MIN_WARPS)
__global__ void __launch_bounds__(BS, SW_MIN_WARPS * 32 / BS) k_tick_small(DevWorld W, u32 phases) {
	PDL_ENTRY();
	__shared__ int4 blo[BS], bhi[BS];
	__shared__ u64 spairs[SW_PAIRS_PER_ENT * BS];
	__shared__ u64 sdead[BS];
	__shared__ int4 sevents[SW_EVENTS];
	__shared__ u32 sctr[CT_WORDS];
	__shared__ i32 sserial[BS];
	__shared__ i32 swq_next[2 * BS], swq_to[2 * BS];
	__shared__ u32 sbase_[2];

	const int g = (int)blockIdx.x;
	const int i = (int)threadIdx.x;
	/* the world-wide lists and counters, before W is pointed at this world's own */
	u32 *const gctr = W.ctr;
	u64 *const gdead = W.deadList;
	int4 *const gevents = W.events;
	const i32 gdead_cap = W.dead_cap, gevent_cap = W.event_cap;
	for (int k = i; k < CT_WORDS; k += BS) sctr[k] = 0;
	W.ctr = sctr;
	W.pairs = spairs; W.pair_cap = SW_PAIRS_PER_ENT * BS;
	W.serialList = sserial; W.serial_cap = BS;
	W.deadList = sdead; W.dead_cap = BS;
	W.events = sevents; W.event_cap = SW_EVENTS;
	W.wq_next = swq_next; W.wq_to = swq_to; W.wq_cap = 2 * BS;
	const i32 b0 = W.wbegin[g], cnt = W.wend[g] - b0;
	const i32 e = b0 + i;
	const bool mine = i < cnt; /* threads past the world's last slot only take part in the barriers */
	/* the staged arrays: in, and W pointed at the copies (slot e of the world = element e - b0) */
#define SW_STAGED(X)                                                                                        \
	X(int4, center) X(int4, old) X(int4, vel) X(int4, radius) X(int4, forcedHoldees) X(Box6, ibox) X(u64, pend) \
	X(u32, type) X(u32, collide) X(u32, hphys) X(u32, forced)                                               \
	X(i32, holdRoot) X(i32, holder) X(i32, holdee) X(i32, LLn) X(i32, LLp) X(i32, holdFlags) X(i32, friction) \
	X(u8, inList) X(u8, needsCollision) X(u8, hasBox) X(u8, ready) X(u8, dead)
#if SW_STAGE
#define SW_DECL(T, name) __shared__ T s_##name[BS]; T *const g_##name = W.name;
#define SW_IN(T, name) if (mine) s_##name[i] = g_##name[e]; W.name = s_##name - b0;
#define SW_OUT(T, name) if (mine) g_##name[e] = s_##name[i];
#else
#define SW_DECL(T, name)
#define SW_IN(T, name)
#define SW_OUT(T, name)
#endif
	/* whose tick handler runs (read up here so that the loads are under way while the state is staged; nothing before
	 * the handlers themselves changes the word, the hold tree or the list) */
	u32 tick_ht = 0, tick_id = 0;
	if ((phases & BT_PHASE_TICKS) && mine && W.inList[e]) {
		tick_ht = W.htick[e];
		tick_id = W.holder[e] >= 0 ? (tick_ht >> 8) & 0xffu : tick_ht & 0xffu;
	}
	SW_STAGED(SW_DECL)
	SW_STAGED(SW_IN)
	/* the world's own words (death flip-flop, "flush1 has ordered work", roots keyed by the ordered lane): read once, up
	 * here; the serial warp of every flush1 looks at them right behind a barrier, with the whole block waiting */
	__shared__ i32 s_flip;
	__shared__ u32 s_wrootseq;
	__shared__ u8 s_wflush;
	i32 *const g_flip = W.flip;
	u32 *const g_wrootseq = W.wrootseq;
	u8 *const g_wflush = W.wflush;
	if (i == 0) { s_flip = g_flip[g]; s_wrootseq = g_wrootseq[g]; s_wflush = g_wflush[g]; }
	W.flip = &s_flip - g; W.wrootseq = &s_wrootseq - g; W.wflush = &s_wflush - g;
	const i32 rp0 = W.root_phase;
	/* row i of the cyclic balanced triangle of k_pairs_small: j = i + 1 .. i + kmax (mod cnt) sees every unordered pair once */
	const int kmax = (cnt - 1) / 2 + (((cnt & 1) == 0 && i < cnt / 2) ? 1 : 0);
	LocalStats st = {0, 0, 0, 0, 0};
	u32 n_serial = 0, iters = 0;
	__syncthreads();

	if ((phases & BT_PHASE_WORLD_RULES) && mine) t_world_rules(W, e);
	if ((phases & BT_PHASE_CTRLS) && mine) t_pre(W, e, phases);
	if (phases & BT_PHASE_TICKS) { /* doUpdates ent.cpp:826-832: handlers read what flushCtrls just published */
		__syncthreads();
		int rnd = 0;
		if (tick_id) {
			rnd = tick_id < 64 && ((BT_SERIAL_TICK >> tick_id) & 1ull); /* the serial pass has work */
			t_tick_id(W, e, tick_ht, (int)tick_id);
		}
		if (__syncthreads_or(rnd) && i == 0) t_tick_serial(W, g); /* gs->rand draws in list order */
	}
	__syncthreads();

	/* stage 0: prepPhysics, then doPhysics up to invokeOnCrush; stage 1: the flush1 that ends doPhysics (one copy of
	 * the flush code for both) */
#pragma unroll 1
	for (int stage = 0; stage < 2; stage++) {
		if (phases & (stage ? BT_PHASE_PHYSICS : BT_PHASE_PREP)) {
			/* flush1 ent.cpp:794-804: requests are looked for by every thread (t_flush_scan), the ordered part is one
			 * warp (bt_flush_serial_warp), flushMisc one thread per entity */
			W.flush_stage = stage; /* deaths flushed in stage 1 are past invokeOnCrush (ent.cpp:907-909) */
			W.root_phase = stage ? rp0 + SW_PHASES - 1 : rp0;
			if (mine) t_flush_scan(W, e);
			__syncthreads();
			if (i < 32) bt_flush_serial_warp(W, g, st);
			__syncthreads();
			if (mine) t_flush_misc(W, e);
			__syncthreads();
			if (i == 0) sctr[CT_WQ_N] = 0; /* every pending-wire list has been consumed */
			W.flush_stage = 0;
			__syncthreads();
		}
		if (stage || !(phases & BT_PHASE_PHYSICS)) continue;

		const int gravity = (phases & BT_PHASE_GRAVITY) != 0;
		/* ---- doGravity, flush2, integrate (ent.cpp:842-858) ---- */
		if (!(phases & SWF_HOST_DELTAS) && !sctr[CT_ANY_DELTA]) {
			if (mine) t_begin_fused(W, e, gravity, 0);
		} else {
			if (gravity && mine) t_gravity(W, e);
			__syncthreads();
			if (mine) t_flush2_apply(W, e);
			__syncthreads();
			if (mine) t_integrate(W, e);
		}
		__syncthreads();
		if (i == 0) { sctr[CT_ANY_DELTA] = 0; sctr[CT_CACHE_INVALID] = 1; } /* candidate pairs are found afresh every tick */
		__syncthreads();
		/* ---- the resolution loop (ent.cpp:860-905) ---- */
		W.iter_base = W.iter_id;
		u32 np = 0;
		int leaf = 0;
		Box6 q;
		q.lo = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, 0); /* empty: touches nothing */
		q.hi = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, 0);
#pragma unroll 1
		while (1) {
			W.iter_id++;
			if (sctr[CT_CACHE_INVALID]) { /* block-uniform: last written before the barrier that ended the previous round */
				leaf = mine && bt_is_leaf_ent(W, e);
				q.lo = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, 0);
				q.hi = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, 0);
				if (leaf) {
					q = bt_inflated_box(W, e);
					W.ibox[e] = q;
				}
				blo[i] = q.lo; bhi[i] = q.hi;
				if (i == 0) sctr[CT_PAIR_N] = 0;
				__syncthreads();
				if (i == 0) sctr[CT_CACHE_INVALID] = 0;
				if (leaf) {
					for (int k = 1; k <= kmax; k++) {
						int j = i + k;
						if (j >= cnt) j -= cnt;
						Box6 o;
						o.lo = blo[j]; o.hi = bhi[j];
						if (bt_box_touch(q, o)) {
							const u32 pos = atomicAdd(&sctr[CT_PAIR_N], 1u);
							if (pos < (u32)(SW_PAIRS_PER_ENT * BS))
								spairs[pos] = (u64)(u32)(b0 + (i < j ? i : j)) | ((u64)(u32)(b0 + (i < j ? j : i)) << 32);
						}
					}
				}
				__syncthreads();
				np = sctr[CT_PAIR_N];
			}
			/* doIteration ent.cpp:437-462: the cached pairs, or -- more touching pairs than the list holds -- straight
			 * from the boxes of the last build; one copy of the narrow phase either way */
			u32 any = 0;
			{
				const bool listed = np <= (u32)(SW_PAIRS_PER_ENT * BS);
				u32 p = (u32)i;
				int k = 1;
#pragma unroll 1
				while (1) {
					u64 pr;
					if (listed) {
						if (p >= np) break;
						pr = spairs[p];
						p += BS;
					} else {
						if (!leaf || k > kmax) break;
						int j = i + k++;
						if (j >= cnt) j -= cnt;
						Box6 o;
						o.lo = blo[j]; o.hi = bhi[j];
						if (!bt_box_touch(q, o)) continue;
						pr = (u64)(u32)(b0 + (i < j ? i : j)) | ((u64)(u32)(b0 + (i < j ? j : i)) << 32);
					}
					t_narrow(W, pr, st, any);
				}
			}
			any = (u32)__syncthreads_or((int)any);
			if (mine) t_decide(W, e); /* clears needsCollision also in the round that proposes nothing (ent.cpp:459) */
			if (!any) break;          /* doIteration returned 0 (ent.cpp:866) */
			if (iters >= (u32)(SW_PHASES - 2)) { /* out of root-order phases for this launch */
				if (i == 0) sctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
				break;
			}
			iters++;
			W.root_phase = rp0 + (i32)iters; /* both push lanes of one iteration share it */
			__syncthreads();
			/* push loop ent.cpp:879-882, both lanes through ONE copy of the push code: round 0 = every ready root by its own
			 * thread (roots whose handlers edit other trees, and tied cycles, queue up instead: bt_push_route), round 1 =
			 * the queue, canonical order, thread 0 */
			const int now = mine && bt_push_route(W, e);
#pragma unroll 1
			for (int round = 0; round < 2; round++) {
				i32 n = now;
				if (round) {
					__syncthreads();
					n = i == 0 && sctr[CT_SERIAL_N] ? bt_serial_sort(W) : 0;
					if (n) { n_serial += (u32)n; sctr[CT_SERIAL_N] = 0; }
				}
#pragma unroll 1
				for (i32 k = 0; k < n; k++) bt_push_root(W, round ? sserial[k] : e, st);
			}
			__syncthreads();
			if (mine) t_post(W, e);
			__syncthreads();
			if (sctr[CT_ERROR]) break; /* block-uniform */
		}
		__syncthreads();
		/* invokeOnCrush (ent.cpp:907): crush handlers are host events */
	}
	if (phases & BT_PHASE_PHYSICS) {
		/* flush2 (ent.cpp:910): a pickup with HOLD_FREEZE left a d_vel behind (ent.cpp:176-181) */
		if (sctr[CT_ANY_DELTA]) {
			if (mine) t_flush2_apply(W, e);
			__syncthreads();
			if (mine) t_zero_deltas(W, e);
		}
	} else if ((phases & BT_PHASE_GRAVITY) && mine) {
		t_gravity(W, e);
	}
	__syncthreads();

	/* ---- what the rest of the library has to see ---- */
	SW_STAGED(SW_OUT)
	u32 live = 0;
	if (phases & BT_PHASE_FINISH) live = (u32)__syncthreads_count(mine && W.inList[e]); /* clearDeads ent.cpp:471-481 */
	if (i == 0) {
		u32 nd = sctr[CT_DEAD_N], ne = sctr[CT_EVENT_N];
		if (ne > (u32)SW_EVENTS) { sctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW); ne = SW_EVENTS; }
		sbase_[0] = nd ? atomicAdd(&gctr[CT_DEAD_N], nd) : 0u;
		sbase_[1] = ne ? atomicAdd(&gctr[CT_EVENT_N], ne) : 0u;
		if (sctr[CT_ERROR]) gctr[CT_ERROR] = sctr[CT_ERROR];
		if (sctr[CT_UNSUPPORTED]) { atomicAdd(&gctr[CT_UNSUPPORTED], sctr[CT_UNSUPPORTED]); gctr[CT_UNSUP_LAST] = sctr[CT_UNSUP_LAST]; }
		if (sctr[CT_ORDER_TIES]) atomicAdd(&gctr[CT_ORDER_TIES], sctr[CT_ORDER_TIES]);
		if (sctr[CT_TREE_EDITS]) atomicAdd(&gctr[CT_TREE_EDITS], sctr[CT_TREE_EDITS]);
		if (sctr[CT_NEED_REBUILD]) gctr[CT_NEED_REBUILD] = 1;
		g_flip[g] = s_flip; g_wrootseq[g] = s_wrootseq; g_wflush[g] = s_wflush;
		u64 *s = W.stat + (blockIdx.x & (BT_STAT_ROWS - 1)) * SK_NUM;
		if (iters) atomicAdd(&s[SK_ITERS], (u64)iters);
		if (live) atomicAdd(&s[SK_ETICKS], (u64)live);
		if (n_serial) atomicAdd(&s[SK_SERIAL], (u64)n_serial);
	}
	__syncthreads();
	{
		const u32 nd = min(sctr[CT_DEAD_N], (u32)BS), ne = min(sctr[CT_EVENT_N], (u32)SW_EVENTS);
		for (u32 k = (u32)i; k < nd; k += BS)
			if (sbase_[0] + k < (u32)gdead_cap) gdead[sbase_[0] + k] = sdead[k];
		for (u32 k = (u32)i; k < ne; k += BS) /* in the order the serial lane raised them */
			if (sbase_[1] + k < (u32)gevent_cap) gevents[sbase_[1] + k] = sevents[k];
	}
	stats_flush(W, st);
}

cudaError_t SW_NAME(bt_launch_tick_small_, SW_BS, SW_MIN_WARPS)(const DevWorld &W, unsigned phases, bool pdl, cudaStream_t stream) {
	if (W.n_worlds <= 0) return cudaSuccess;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)W.n_worlds);
	cfg.blockDim = dim3((unsigned)BS);
	cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, k_tick_small, W, (u32)phases);
}
