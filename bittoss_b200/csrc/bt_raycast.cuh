/*
 * bt_raycast.cuh — batched selectCast (raycast.cpp:16-57) over an LBVH of ALL live
 * entities (the reference tests every ent, boxed or not, raycast.cpp:33).
 *
 * The reference is a brute-force scan in gs->ents order (newest first) that keeps a hit
 * only if it is STRICTLY nearer than the best so far, so the winner is the hit with the
 * smallest entry parameter `lower`, ties going to the entity met first, i.e. the larger
 * index.  That is an order-independent reduction, which is what lets a tree traversal
 * in arbitrary order return exactly the same entity.  All parameters are exact
 * rationals n/d compared by 64-bit cross-multiplication like the reference's lt().
 */
#ifndef BT_RAYCAST_CUH
#define BT_RAYCAST_CUH

#include "bt_phys.cuh"

struct Rat {
	i32 n, d;
};
/* lt() raycast.cpp:14 */
static BT_HD int rat_lt(const Rat &a, const Rat &b) { return (long long)a.n * b.d < (long long)b.n * a.d; }

/* The reference's per-entity slab test (raycast.cpp:36-50) with `upper` starting at the
 * largest rational instead of the running best; returns 1 and *out = lower on a hit. */
static BT_HD int ray_hit_ent(const i32 pos[3], const i32 look[3], const i32 flip[3], const int4 &c, const int4 &r,
                             Rat *out) {
	const i32 cc[3] = {c.x, c.y, c.z}, rr[3] = {r.x, r.y, r.z};
	Rat lower = {0, 1};
	Rat upper = {INT32_MAX, 1};
#pragma unroll
	for (int d = 0; d < 3; d++) {
		i32 x = flip[d] * (cc[d] - pos[d]);
		Rat f1 = {x - rr[d], look[d]}, f2 = {x + rr[d], look[d]};
		if (rat_lt(lower, f1)) lower = f1;
		if (rat_lt(f2, upper)) upper = f2;
		if (!rat_lt(lower, upper)) return 0;
	}
	if (!lower.n) return 0; /* only things a positive distance in front of us (raycast.cpp:51) */
	*out = lower;
	return 1;
}

/* conservative slab test of a node box against the ray, pruned by the best hit so far:
 * a subtree is skipped only if every entity in it would enter strictly after `best`. */
static BT_HD int ray_hit_box(const i32 pos[3], const i32 look[3], const i32 flip[3], const i32 lo[3], const i32 hi[3],
                             const Rat &best, Rat *entry) {
	Rat lower = {0, 1};
	Rat upper = {INT32_MAX, 1};
#pragma unroll
	for (int d = 0; d < 3; d++) {
		/* in sign-folded coordinates the box spans [a, b] */
		long long a64 = flip[d] > 0 ? (long long)lo[d] - pos[d] : (long long)pos[d] - hi[d];
		long long b64 = flip[d] > 0 ? (long long)hi[d] - pos[d] : (long long)pos[d] - lo[d];
		/* clamp to int32: boxes of huge entities may exceed it; clamping only widens */
		i32 a = a64 < INT32_MIN ? INT32_MIN : (a64 > INT32_MAX ? INT32_MAX : (i32)a64);
		i32 b = b64 < INT32_MIN ? INT32_MIN : (b64 > INT32_MAX ? INT32_MAX : (i32)b64);
		Rat f1 = {a, look[d]}, f2 = {b, look[d]};
		if (look[d] == 0) {
			if (a > 0 || b < 0) return 0; /* origin outside the slab on an axis we do not move along */
			continue;
		}
		if (rat_lt(lower, f1)) lower = f1;
		if (rat_lt(f2, upper)) upper = f2;
		if (rat_lt(upper, lower)) return 0; /* empty even when touching is allowed */
	}
	if (rat_lt(best, lower)) return 0; /* enters strictly after the best hit */
	*entry = lower; /* where the ray enters the box: orders the two children front to back */
	return 1;
}

/* one ray: origin entity `p` (its hold tree is skipped), direction look[3] */
static BT_HD i32 t_raycast(const DevWorld &W, i32 p, const i32 look_in[3], i32 *stack, int stride) {
	int L = W.n_leaves;
	if (L < 1) return -1;
	int4 pc = W.center[p];
	const i32 pos[3] = {pc.x, pc.y, pc.z};
	i32 look[3], flip[3];
	for (int i = 0; i < 3; i++) { /* raycast.cpp:20-28 (on a private copy) */
		if (look_in[i] >= 0) { flip[i] = 1; look[i] = look_in[i]; }
		else { flip[i] = -1; look[i] = -look_in[i]; }
	}
	i32 root = W.holdRoot[p];
	u32 wp = (u32)W.world[p];
	Rat best = {INT32_MAX, 1};
	i32 winner = -1;
#define RAY_LEAF(t)                                                                              \
	do {                                                                                         \
		i32 e_ = W.leafEnt[t];                                                                   \
		if (W.inList[e_] && W.holdRoot[e_] != root && (u32)W.world[e_] == wp) {                  \
			Rat lw_;                                                                             \
			if (ray_hit_ent(pos, look, flip, W.center[e_], W.radius[e_], &lw_)) {                \
				if (rat_lt(lw_, best) || (winner >= 0 && !rat_lt(best, lw_) && e_ > winner)) { best = lw_; winner = e_; } \
			}                                                                                    \
		}                                                                                        \
	} while (0)
	if (L == 1) {
		RAY_LEAF(0);
		return winner;
	}
	int sp = 0, node = W.wroot[wp];
	if (node < 0) { /* a single leaf (~node), or none */
		if (~node < L) RAY_LEAF(~node);
		return winner;
	}
	while (1) {
		const NodeRegs N = bt_load_node(W, node);
		i32 li = N.d.x, ri = N.d.y;
		u32 wl = (u32)N.d.z, wr = (u32)N.d.w;
		const i32 llo[3] = {N.a.x, N.a.y, N.a.z}, lhi[3] = {N.a.w, N.b.x, N.b.y};
		const i32 rlo[3] = {N.b.z, N.b.w, N.c.x}, rhi[3] = {N.c.y, N.c.z, N.c.w};
		Rat eL = {0, 1}, eR = {0, 1};
		int ovL = (wl & 0xffff) <= wp && wp <= (wl >> 16) && ray_hit_box(pos, look, flip, llo, lhi, best, &eL);
		int ovR = (wr & 0xffff) <= wp && wp <= (wr >> 16) && ray_hit_box(pos, look, flip, rlo, rhi, best, &eR);
		if (ovL && li < 0) RAY_LEAF(~li);
		if (ovR && ri < 0) RAY_LEAF(~ri);
		int goL = ovL && li >= 0, goR = ovR && ri >= 0;
		if (goL && goR) {
			/* nearer child first: its hits shrink `best` and prune the farther one when it is
			 * popped (its children are tested against the new best).  Order only affects speed. */
			int rightFirst = rat_lt(eR, eL);
			node = rightFirst ? ri : li;
			if (sp < BT_STACK) stack[(sp++) * stride] = rightFirst ? li : ri;
			else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
		} else if (goL) {
			node = li;
		} else if (goR) {
			node = ri;
		} else {
			if (sp == 0) break;
			node = stack[(--sp) * stride];
		}
	}
#undef RAY_LEAF
	return winner;
}

#ifndef BT_EMU
/* Ray order.  A batch arrives in the caller's order (origins hashed all over the world), so the
 * 32 rays of a warp would start in 32 different corners of the tree and share nothing.  Rays are therefore traversed
 * in the order of their origin's position in the ray tree's Morton-sorted leaf array: neighbours in a warp start in
 * the same leaves and walk the same upper nodes.  k_ray_leafpos inverts leafEnt once per tree build, k_ray_keys makes
 * (leaf position, ray index) pairs for the radix sort; the winners are written back by ray index, so the caller sees
 * no reordering. */
__global__ void __launch_bounds__(256) k_ray_leafpos(DevWorld W, i32 *leafPos) { PDL_ENTRY();
	GS_FOR(t, W.n_leaves) leafPos[W.leafEnt[t]] = t;
}
__global__ void __launch_bounds__(256) k_ray_keys(const i32 *origin, i32 n_rays, const i32 *rank2slot, const i32 *leafPos,
                                                   i32 n_live, u64 *keys, i32 *idx) { PDL_ENTRY_ALWAYS();
	for (i32 i = TID; i < n_rays; i += (i32)(gridDim.x * blockDim.x)) {
		const i32 og = origin[i];
		keys[i] = (og < 0 || og >= n_live) ? 0ull : (u64)(u32)leafPos[rank2slot[og]];
		idx[i] = i;
	}
}

/* origin/winner are LIVE RANKS (oldest-first order of bt_world_download); rank2slot /
 * slot2rank translate to device slots.  order: NULL, or the ray index each thread takes. */
__global__ void __launch_bounds__(128) k_raycast(DevWorld W, const i32 *origin, const i32 *look, i32 n_rays,
                                                  i32 *winner, const i32 *rank2slot, const i32 *slot2rank, i32 n_live,
                                                  const i32 *order) { PDL_ENTRY();
	__shared__ i32 stacks[BT_STACK * 128];
	i32 i = (i32)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= n_rays) return;
	if (order) i = order[i];
	i32 lk[3] = {look[3 * i], look[3 * i + 1], look[3 * i + 2]};
	i32 og = origin[i];
	if (og < 0 || og >= n_live) { winner[i] = -1; return; } /* not a live entity: no caster, no hit */
	i32 win = t_raycast(W, rank2slot[og], lk, stacks + threadIdx.x, 128);
	winner[i] = win >= 0 ? slot2rank[win] : -1;
}

/* The same sweep with the rays handed out dynamically (option ray_dynamic, the default for batches of >= 4096 rays).
 * In k_raycast a lane that has found its winner idles until the longest ray of its warp is done: rays that leave the
 * world traverse an order of magnitude more nodes than rays that hit the next column, 9.2 of 32 threads were active per
 * issued instruction and the kernel is bound by instruction issue (63 % of the issue slots).  Here a warp keeps pulling
 * rays: chunks of RAY_CHUNK rays from one global counter, and whenever RAY_REFILL or more of its lanes are idle each of
 * them takes the next ray of the chunk, so every traversal step runs on (nearly) full warps.  The step is the loop body
 * of t_raycast, the result the same order-independent minimum. */
#define RAY_CHUNK 256
#ifndef RAY_REFILL
#define RAY_REFILL 8
#endif
__global__ void __launch_bounds__(128) k_raycast_dyn(DevWorld W, const i32 *origin, const i32 *look_in, i32 n_rays,
                                                      i32 *winner_out, const i32 *rank2slot, const i32 *slot2rank, i32 n_live,
                                                      const i32 *order, u32 *next) { PDL_ENTRY();
	__shared__ i32 stacks[BT_STACK * 128];
	i32 *const stack = stacks + threadIdx.x;
	const int lane = threadIdx.x & 31;
	const int L = W.n_leaves;
	u32 pb = 0, pe = 0; /* this warp's chunk of rays (uniform over the warp) */
	int exhausted = 0;  /* the global counter has run past the batch */
	int active = 0;
	i32 ray = -1, root = -1, winner = -1, node = 0;
	u32 wp = 0;
	int sp = 0;
	i32 pos[3] = {0, 0, 0}, look[3] = {0, 0, 1}, flip[3] = {1, 1, 1};
	Rat best = {INT32_MAX, 1};
#define RAY_LEAF(t)                                                                              \
	do {                                                                                         \
		i32 e_ = W.leafEnt[t];                                                                   \
		if (W.inList[e_] && W.holdRoot[e_] != root && (u32)W.world[e_] == wp) {                  \
			Rat lw_;                                                                             \
			if (ray_hit_ent(pos, look, flip, W.center[e_], W.radius[e_], &lw_)) {                \
				if (rat_lt(lw_, best) || (winner >= 0 && !rat_lt(best, lw_) && e_ > winner)) { best = lw_; winner = e_; } \
			}                                                                                    \
		}                                                                                        \
	} while (0)
	while (1) {
		const u32 idle = __ballot_sync(0xffffffffu, !active);
		if (idle && (__popc(idle) >= RAY_REFILL || idle == 0xffffffffu || exhausted)) {
			if (pb == pe && !exhausted) {
				u32 b = 0;
				if (lane == 0) b = atomicAdd(next, (u32)RAY_CHUNK);
				b = __shfl_sync(0xffffffffu, b, 0);
				if (b >= (u32)n_rays) { exhausted = 1; pb = pe = 0; }
				else { pb = b; pe = b + RAY_CHUNK < (u32)n_rays ? b + RAY_CHUNK : (u32)n_rays; }
			}
			const u32 avail = pe - pb;
			if (!avail && idle == 0xffffffffu) break; /* nothing left anywhere and nobody is travelling */
			const u32 rank = (u32)__popc(idle & ((1u << lane) - 1));
			if (!active && rank < avail) { /* take ray pb + rank */
				ray = (i32)(pb + rank);
				if (order) ray = order[ray];
				const i32 og = origin[ray];
				winner = -1;
				if (og < 0 || og >= n_live || L < 1) {
					winner_out[ray] = -1; /* not a live entity: no caster, no hit */
				} else {
					const i32 p = rank2slot[og];
					const int4 pc = W.center[p];
					pos[0] = pc.x; pos[1] = pc.y; pos[2] = pc.z;
#pragma unroll
					for (int i = 0; i < 3; i++) { /* raycast.cpp:20-28 (on a private copy) */
						const i32 l = look_in[3 * ray + i];
						if (l >= 0) { flip[i] = 1; look[i] = l; }
						else { flip[i] = -1; look[i] = -l; }
					}
					root = W.holdRoot[p];
					wp = (u32)W.world[p];
					best.n = INT32_MAX; best.d = 1;
					sp = 0;
					node = L == 1 ? ~0 : W.wroot[wp];
					if (node < 0) { /* a single leaf (~node), or none */
						if (~node < L) RAY_LEAF(~node);
						winner_out[ray] = winner >= 0 ? slot2rank[winner] : -1;
					} else {
						active = 1;
					}
				}
			}
			pb += __popc(idle) < (int)avail ? (u32)__popc(idle) : avail;
		}
		if (active) { /* one node */
			const NodeRegs N = bt_load_node(W, node);
			i32 li = N.d.x, ri = N.d.y;
			u32 wl = (u32)N.d.z, wr = (u32)N.d.w;
			const i32 llo[3] = {N.a.x, N.a.y, N.a.z}, lhi[3] = {N.a.w, N.b.x, N.b.y};
			const i32 rlo[3] = {N.b.z, N.b.w, N.c.x}, rhi[3] = {N.c.y, N.c.z, N.c.w};
			Rat eL = {0, 1}, eR = {0, 1};
			int ovL = (wl & 0xffff) <= wp && wp <= (wl >> 16) && ray_hit_box(pos, look, flip, llo, lhi, best, &eL);
			int ovR = (wr & 0xffff) <= wp && wp <= (wr >> 16) && ray_hit_box(pos, look, flip, rlo, rhi, best, &eR);
			/* leaf children: ONE copy of the entity test for both sides, so that lanes with a left leaf and lanes with a
			 * right one go through it together (a few lanes of a warp have one at any time, and the test is three
			 * dependent gathers) */
			i32 lf0 = -1, lf1 = -1;
			if (ovL && li < 0) lf0 = ~li;
			if (ovR && ri < 0) { if (lf0 < 0) lf0 = ~ri; else lf1 = ~ri; }
#pragma unroll 1
			while (lf0 >= 0) {
				RAY_LEAF(lf0);
				lf0 = lf1;
				lf1 = -1;
			}
			int goL = ovL && li >= 0, goR = ovR && ri >= 0;
			if (goL && goR) {
				int rightFirst = rat_lt(eR, eL);
				node = rightFirst ? ri : li;
				if (sp < BT_STACK) stack[(sp++) * 128] = rightFirst ? li : ri;
				else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
			} else if (goL) {
				node = li;
			} else if (goR) {
				node = ri;
			} else if (sp == 0) {
				winner_out[ray] = winner >= 0 ? slot2rank[winner] : -1;
				active = 0;
			} else {
				node = stack[(--sp) * 128];
			}
		}
	}
#undef RAY_LEAF
}

/* inList -> live rank per slot (rank2slot / slot2rank), three small launches over tiles of RANK_TILE slots:
 * count the live slots of every tile, scan the tile counts (one block), rank inside each tile from its base.
 * (One block looping over the whole world took 0.89 ms at 1 M slots.) */
#define RANK_TILE 1024
__global__ void __launch_bounds__(RANK_TILE) k_rank_count(DevWorld W, u32 *tileCount) { PDL_ENTRY();
	const i32 i = (i32)(blockIdx.x * RANK_TILE + threadIdx.x);
	const int live = i < W.n && W.inList[i];
	const int c = __syncthreads_count(live);
	if (threadIdx.x == 0) tileCount[blockIdx.x] = (u32)c;
}
__global__ void __launch_bounds__(1024) k_rank_scan(u32 *tileCount, int tiles) { PDL_ENTRY_ALWAYS();
	__shared__ u32 warpTot[32];
	const int chunk = (tiles + 1023) / 1024;
	const int b0 = min(tiles, (int)threadIdx.x * chunk), b1 = min(tiles, b0 + chunk);
	u32 sum = 0;
	for (int i = b0; i < b1; i++) sum += tileCount[i];
	u32 x = sum;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		u32 y = __shfl_up_sync(0xffffffffu, x, off);
		if ((threadIdx.x & 31) >= off) x += y;
	}
	if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
	__syncthreads();
	if (threadIdx.x < 32) {
		u32 w = warpTot[threadIdx.x], t = w;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, t, off);
			if (threadIdx.x >= off) t += y;
		}
		warpTot[threadIdx.x] = t - w;
	}
	__syncthreads();
	u32 run = warpTot[threadIdx.x >> 5] + x - sum;
	for (int i = b0; i < b1; i++) {
		const u32 v = tileCount[i];
		tileCount[i] = run;
		run += v;
	}
}
__global__ void __launch_bounds__(RANK_TILE) k_rank_apply(DevWorld W, const u32 *tileBase, i32 *rank2slot, i32 *slot2rank) { PDL_ENTRY();
	__shared__ u32 warpTot[RANK_TILE / 32];
	const i32 i = (i32)(blockIdx.x * RANK_TILE + threadIdx.x);
	const u32 v = (i < W.n && W.inList[i]) ? 1u : 0u;
	const u32 m = __ballot_sync(0xffffffffu, v);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) warpTot[warp] = (u32)__popc(m);
	__syncthreads();
	if (warp == 0) {
		u32 w = warpTot[lane], t = w;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 y = __shfl_up_sync(0xffffffffu, t, off);
			if (lane >= off) t += y;
		}
		warpTot[lane] = t - w;
	}
	__syncthreads();
	if (i < W.n) {
		const u32 r = tileBase[blockIdx.x] + warpTot[warp] + (u32)__popc(m & ((1u << lane) - 1));
		slot2rank[i] = v ? (i32)r : -1;
		if (v) rank2slot[r] = i;
	}
}
#endif

#endif
