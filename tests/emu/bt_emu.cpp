/*
 * bt_emu.cpp — TEST TOOLING ONLY (built by tests/conftest.py into tests/emu/libbtemu.so).
 *
 * Single-steps the per-thread kernel bodies of bittoss_b200/csrc/bt_phys.cuh and
 * bt_raycast.cuh on the HOST, one logical thread after another, in the same kernel order
 * as bittoss_b200/csrc/bt_world.cu.  It exists because the development container has no
 * GPU: it lets the `-m "not gpu"` tests check the kernel logic (LBVH build, refit,
 * traversal, proposal folding, push lanes, flushes, hash, raycast) against the oracle
 * before the same source is run as CUDA on a B200.  It is NOT part of the product, is
 * not linked into libbittoss_b200.so, and is never used by bench.py: the shipped library
 * has no host execution path.
 */
#define BT_EMU 1
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../bittoss_b200/csrc/bt_phys.cuh"
#include "../../bittoss_b200/csrc/bt_raycast.cuh"

struct Emu {
	DevWorld W;
	std::vector<void *> allocs;
	bool tree_valid, deltas_pending, ctrls_fused;
	bool small_ok; /* every world <= BT_SMALL_WORLD entities */
	int full_flush_left, scan_flushes_left;
	int ticks_since_build, rebuild_every;
	int64_t iterations, ticks;
	LocalStats st;
	std::vector<i32> up_root_order;
	int pair_cache;   /* 1: k_pairs + k_narrow, 0: k_collide every iteration */
	u32 pair_builds_at_tree;
	int64_t pair_rebuilds, pair_patches;
	int shuffle_seed; /* != 0: run the threads of order-sensitive kernels in a scrambled order */
};

template <typename T> static void A(Emu *E, T **p, size_t n) {
	*p = (T *)calloc(n ? n : 1, sizeof(T));
	E->allocs.push_back(*p);
}

/* thread order of a "kernel": identity, or a seeded shuffle to mimic the GPU's freedom */
static std::vector<i32> order(Emu *E, i32 n) {
	std::vector<i32> o(n);
	for (i32 i = 0; i < n; i++) o[i] = i;
	if (E->shuffle_seed) {
		uint32_t s = (uint32_t)E->shuffle_seed * 2654435761u + (uint32_t)E->W.iter_id;
		for (i32 i = n - 1; i > 0; i--) {
			s = s * 1664525u + 1013904223u;
			i32 j = (i32)((s >> 8) % (uint32_t)(i + 1));
			std::swap(o[i], o[j]);
		}
	}
	return o;
}

extern "C" {

void bte_destroy(Emu *E) {
	if (!E) return;
	for (void *p : E->allocs) free(p);
	delete E;
}

Emu *bte_create(const bt_world_desc *d, int rebuild_every, int shuffle_seed) {
	Emu *E = new Emu();
	memset(&E->W, 0, sizeof(E->W));
	memset(&E->st, 0, sizeof(E->st));
	E->tree_valid = false; E->ticks_since_build = 0; E->ctrls_fused = false; E->full_flush_left = 2; E->scan_flushes_left = 2; E->rebuild_every = rebuild_every > 0 ? rebuild_every : 1;
	E->iterations = 0; E->ticks = 0; E->shuffle_seed = shuffle_seed; E->pair_cache = 1; E->pair_rebuilds = 0; E->pair_patches = 0;
	DevWorld &W = E->W;
	const int n = d->n_ents, nw = d->n_worlds;
	W.n = n; W.n_worlds = nw;
	A(E, &W.wbegin, nw); A(E, &W.wend, nw); A(E, &W.flip, nw); A(E, &W.gamerules, nw); A(E, &W.rnd, nw); A(E, &W.wflush, nw); A(E, &W.wunsup, nw); A(E, &W.wties, nw);
	A(E, &W.world, n); A(E, &W.uid, n); A(E, &W.center, n); A(E, &W.old, n); A(E, &W.vel, n); A(E, &W.radius, n);
	A(E, &W.d_center, n); A(E, &W.d_vel, n); A(E, &W.type, n); A(E, &W.collide, n); A(E, &W.ntype, n); A(E, &W.ncollide, n);
	A(E, &W.friction, n); A(E, &W.holder, n); A(E, &W.holdFlags, n); A(E, &W.holdRoot, n); A(E, &W.holdee, n);
	A(E, &W.LLn, n); A(E, &W.LLp, n); A(E, &W.newHolder, n); A(E, &W.newHoldFlags, n);
	A(E, &W.hphys, n); A(E, &W.hpair, n); A(E, &W.htick, n); A(E, &W.btn_v, n); A(E, &W.btn_v2, n);
	A(E, &W.newDrop, n); A(E, &W.dead_max0, n); A(E, &W.dead_max1, n); A(E, &W.fullFlush, n); A(E, &W.dead, n);
	A(E, &W.inList, n); A(E, &W.needsCollision, n); A(E, &W.hasBox, n); A(E, &W.ready, n);
	A(E, &W.forced, n); A(E, &W.forcedHoldees, n); A(E, &W.pend, n);
	A(E, &W.slider_off, n + 1); A(E, &W.sv, d->n_sliders); A(E, &W.smin, d->n_sliders); A(E, &W.smax, d->n_sliders);
	A(E, &W.wire_off, n + 1); A(E, &W.wire_cnt, n); A(E, &W.wire_to, d->n_wires + n);
	A(E, &W.wireAdd, n); A(E, &W.wireRm, n);
	A(E, &W.keys, n); A(E, &W.keys_alt, n); A(E, &W.leafEnt, n); A(E, &W.leafEnt_alt, n);
	A(E, &W.leafParent, n); A(E, &W.nodeParent, n); A(E, &W.nodeCount, n); A(E, &W.nodes, n);
	A(E, &W.nodeBox, n); A(E, &W.leafBox, n); A(E, &W.nodeHi, n); A(E, &W.ibox, n);
	A(E, &W.wroot, nw + 1);
	A(E, &W.iboxOld, n); A(E, &W.leafOf, n); A(E, &W.dirtyList, n); A(E, &W.dirtyStamp, n);
	W.patch = 1; W.patch_shift = 4;
	W.pair_cap = 12 * n + 4096; W.margin = 32;
	A(E, &W.pairs, W.pair_cap);
	W.serial_cap = 1 << 16; W.dead_cap = n ? n : 1; W.event_cap = 1 << 16;
	A(E, &W.events, W.event_cap);
	A(E, &W.serialList, W.serial_cap); A(E, &W.deadList, W.dead_cap); A(E, &W.hash, nw); A(E, &W.ctr, CT_WORDS);
	for (int g = 0; g < nw; g++) {
		W.wbegin[g] = n; W.wend[g] = 0;
		W.flip[g] = d->w_flip_flop[g]; W.gamerules[g] = d->w_gamerules[g]; W.rnd[g] = d->w_rand[g];
	}
	bool deltas = false;
	for (int i = 0; i < n; i++) {
		int g = d->world_id ? d->world_id[i] : 0;
		W.world[i] = g;
		W.wbegin[g] = std::min(W.wbegin[g], i); W.wend[g] = std::max(W.wend[g], i + 1);
		W.uid[i] = d->uid[i];
#define V3(dst, src) dst[i] = make_int4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0)
		V3(W.center, d->center); V3(W.old, d->old); V3(W.vel, d->vel); V3(W.radius, d->radius);
		V3(W.d_center, d->d_center); V3(W.d_vel, d->d_vel);
#undef V3
		for (int k = 0; k < 3; k++) deltas = deltas || d->d_center[3 * i + k] || d->d_vel[3 * i + k];
		W.type[i] = d->type_mask[i]; W.collide[i] = d->collide_mask[i];
		W.ntype[i] = d->new_type_mask[i]; W.ncollide[i] = d->new_collide_mask[i];
		W.friction[i] = d->friction[i]; W.holder[i] = d->holder[i]; W.holdFlags[i] = d->hold_flags[i];
		W.holdee[i] = d->holdee_head[i]; W.newHolder[i] = d->new_holder[i]; W.newHoldFlags[i] = d->new_hold_flags[i];
		const uint8_t *h = d->handlers + (size_t)BT_H_NUM * i;
		W.hphys[i] = h[BT_H_WHOMOVES] | (h[BT_H_PUSH] << 8) | (h[BT_H_PUSHED] << 16) | ((u32)h[BT_H_CRUSH] << 24);
		W.hpair[i] = h[BT_H_ONFUMBLE] | (h[BT_H_ONFUMBLED] << 8) | (h[BT_H_ONPICKUP] << 16) | ((u32)h[BT_H_ONPICKEDUP] << 24);
		W.htick[i] = h[BT_H_TICK] | (h[BT_H_TICKHELD] << 8);
		const uint8_t *b = d->btn + 8 * (size_t)i;
		W.btn_v[i] = b[0] | (b[1] << 8) | (b[2] << 16) | ((u32)b[3] << 24);
		W.btn_v2[i] = b[4] | (b[5] << 8) | (b[6] << 16) | ((u32)b[7] << 24);
		const uint8_t *m = d->misc + (size_t)BT_M_NUM * i;
		W.newDrop[i] = m[BT_M_NEWDROP]; W.dead_max0[i] = m[BT_M_DEADMAX0]; W.dead_max1[i] = m[BT_M_DEADMAX1];
		W.fullFlush[i] = m[BT_M_FULLFLUSH];
		W.inList[i] = 1;
		W.hasBox[i] = (u8)bt_needs_box(W.type[i], W.collide[i]);
		W.forcedHoldees[i] = make_int4(-1, -1, -1, 0);
		W.slider_off[i] = d->slider_off[i];
		W.wire_cnt[i] = d->wire_off[i + 1] - d->wire_off[i];
		W.wireAdd[i] = -1; W.wireRm[i] = -1;
		W.LLn[i] = -1; W.LLp[i] = -1;
	}
	W.slider_off[n] = d->slider_off[n];
	{ /* one spare wire slot behind every seat, like bt_world_upload */
		i32 o = 0;
		for (int i = 0; i < n; i++) {
			W.wire_off[i] = o;
			o += W.wire_cnt[i] + (d->handlers[(size_t)BT_H_NUM * i + BT_H_ONPICKUP] == 1 ? 1 : 0);
		}
		W.wire_off[n] = o;
	}
	for (int s = 0; s < d->n_sliders; s++) { W.sv[s] = d->slider_v[s]; W.smin[s] = d->slider_min[s]; W.smax[s] = d->slider_max[s]; }
	for (int i = 0; i < n; i++)
		for (i32 k = 0; k < W.wire_cnt[i]; k++) W.wire_to[W.wire_off[i] + k] = d->wire_to[d->wire_off[i] + k];
	for (int i = 0; i < n; i++)
		if (W.holder[i] >= 0) {
			W.LLn[i] = d->holdee_next[i];
			if (W.LLn[i] >= 0) W.LLp[W.LLn[i]] = i;
		}
	for (int i = 0; i < n; i++) {
		i32 r = i;
		while (W.holder[r] >= 0) r = W.holder[r];
		W.holdRoot[i] = r;
	}
	for (int g = 0; g < nw; g++)
		if (W.wbegin[g] > W.wend[g]) W.wbegin[g] = W.wend[g] = 0;
	W.ctr[CT_CACHE_INVALID] = 1; W.ctr[CT_CACHE_FULL] = 1;
	E->pair_builds_at_tree = 0;
	E->deltas_pending = deltas;
	E->up_root_order.assign(d->root_order, d->root_order + n);
	{
		i32 biggest = 0;
		for (int g = 0; g < nw; g++) biggest = std::max(biggest, W.wend[g] - W.wbegin[g]);
		E->small_ok = biggest <= BT_SMALL_WORLD;
		W.small_mode = E->small_ok ? 1 : 0;
	}
	return E;
}

static int bits_for(u32 x) { int b = 0; while (x) { b++; x >>= 1; } return b; }

static void build_tree(Emu *E) {
	DevWorld &W = E->W;
	i32 *b = (i32 *)&W.ctr[CT_BOUNDS];
	b[0] = b[1] = b[2] = INT32_MAX; b[3] = b[4] = b[5] = INT32_MIN;
	i32 L = 0;
	for (i32 e = 0; e < W.n; e++) {
		t_bounds(W, e);
		if (bt_is_leaf_ent(W, e)) L++;
	}
	W.n_leaves = L;
	for (int k = 0; k < 3; k++) {
		u32 ext = L ? (u32)(b[3 + k] - b[k]) : 0;
		int bits = bits_for(ext);
		W.key_shift[k] = bits > 9 ? bits - 9 : 0; /* 9 bits per axis */
	}
	for (i32 e = 0; e < W.n; e++) t_morton(W, e);
	/* stand-in for the GPU radix sort: any stable sort by key gives the same result */
	std::vector<i32> idx(W.n);
	for (i32 i = 0; i < W.n; i++) idx[i] = i;
	std::stable_sort(idx.begin(), idx.end(), [&](i32 x, i32 y) { return W.keys[x] < W.keys[y]; });
	std::vector<u64> k2(W.n);
	std::vector<i32> v2(W.n);
	for (i32 i = 0; i < W.n; i++) { k2[i] = W.keys[idx[i]]; v2[i] = W.leafEnt[idx[i]]; }
	for (i32 i = 0; i < W.n; i++) { W.keys[i] = k2[i]; W.leafEnt[i] = v2[i]; }
	for (i32 g = 0; g <= W.n_worlds; g++) W.wroot[g] = INT32_MIN;
	for (i32 i = 0; i < L - 1; i++) t_lbvh_node(W, i);
	for (i32 i = 0; i < W.n; i++) W.nodeCount[i] = 0;
	if (!W.ray_mode) { W.ctr[CT_CACHE_INVALID] = 1; W.ctr[CT_CACHE_FULL] = 1; }
	E->tree_valid = true;
	E->ticks_since_build = 0;
	E->pair_builds_at_tree = W.ctr[CT_PAIR_BUILDS] + W.ctr[CT_PAIR_PATCHES] + 1; /* + the pair search this build itself forces */
}

static void refit(Emu *E, int cond) {
	DevWorld &W = E->W;
	/* thread 0 may reset counters: run it first like any other thread would see it (the flag is
	 * only read, never written, by this kernel) */
	std::vector<i32> o = order(E, W.n_leaves);
	for (i32 t : o) t_refit(W, t, cond);
}

static void flush1(Emu *E, bool hint) {
	DevWorld &W = E->W;
	if (E->full_flush_left <= 0 && !hint) {
		for (i32 g = 0; g < W.n_worlds; g++) t_flip_only(W, g);
		return;
	}
	if (E->full_flush_left > 0) E->full_flush_left--;
	W.ctr[CT_FLUSH_HINT] = 0;
	if (E->scan_flushes_left > 0) { /* only what the upload brought along needs looking for */
		E->scan_flushes_left--;
		for (i32 e = 0; e < W.n; e++) t_flush_scan(W, e);
	}
	for (i32 g = 0; g < W.n_worlds; g++) t_flush_serial(W, g, E->st);
	for (i32 e = 0; e < W.n; e++) t_flush_misc(W, e);
}

static int do_physics(Emu *E, bool gravity) {
	DevWorld &W = E->W;
	const i32 n = W.n;
	if (!E->deltas_pending) {
		for (i32 e = 0; e < n; e++) t_begin_fused(W, e, gravity ? 1 : 0, E->ctrls_fused ? 1 : 0);
	} else {
		if (gravity) for (i32 e = 0; e < n; e++) t_gravity(W, e);
		for (i32 e = 0; e < n; e++) t_flush2_apply(W, e);
		for (i32 e = 0; e < n; e++) t_integrate(W, e);
		E->deltas_pending = false;
	}
	const bool need_tree = !(W.small_mode && E->pair_cache);
	if (need_tree && (!E->tree_valid || (E->ticks_since_build >= E->rebuild_every &&
	                       (!E->pair_cache || W.ctr[CT_PAIR_BUILDS] + W.ctr[CT_PAIR_PATCHES] > E->pair_builds_at_tree))))
		build_tree(E);
	E->ticks_since_build++;
	W.iter_base = W.iter_id;
	std::vector<i32> stack(BT_STACK);
	int guard = 0;
	struct HostEmit {
		const DevWorld &W;
		void operator()(i32 a, i32 b) {
			u32 pos = bt_add32(&W.ctr[CT_PAIR_N], 1u);
			if ((i32)pos < W.pair_cap) W.pairs[pos] = (u64)(u32)a | ((u64)(u32)b << 32);
		}
	};
	int cache_state = 0, it = 0;
	while (1) {
		W.iter_id++;
		u32 any = 0;
		if (E->pair_cache) {
			if (cache_state != 2) {
				u32 before = W.ctr[CT_PAIR_BUILDS], pbefore = W.ctr[CT_PAIR_PATCHES];
				if (W.small_mode) { /* k_refit(cond) in small-world mode: reset only */
					if (W.ctr[CT_CACHE_INVALID]) { W.ctr[CT_PAIR_N] = 0; W.ctr[CT_PAIR_BUILDS]++; }
				} else {
					refit(E, 1);
				}
				{
					HostEmit emit = {W};
					if (W.small_mode) {
						std::vector<i32> o = order(E, W.n_worlds);
						for (i32 g : o) t_pairs_small_world(W, g, emit);
					} else if (W.ctr[CT_CACHE_INVALID] && bt_patch_mode(W)) {
						std::vector<i32> o = order(E, (i32)W.ctr[CT_DIRTY_N]);
						for (i32 t : o) t_pairs_patch(W, t, stack.data(), 1, emit);
					} else {
						std::vector<i32> o = order(E, W.n_leaves);
						for (i32 t : o) t_pairs(W, t, stack.data(), 1, emit);
					}
				}
				if (W.ctr[CT_PAIR_PATCHES] != pbefore) E->pair_patches++;
				if ((i32)W.ctr[CT_PAIR_N] > W.pair_cap) return BT_ERR_OVERFLOW;
				if (W.ctr[CT_PAIR_BUILDS] != before) E->pair_rebuilds++;
			}
			bt_consume_cache_flags(W, W.ctr[CT_PAIR_N]); /* narrow kernels, block 0 thread 0 */
			{
				std::vector<i32> o = order(E, (i32)W.ctr[CT_PAIR_N]);
				for (i32 p : o) t_narrow(W, W.pairs[p], E->st, any);
			}
		} else {
			if (!E->tree_valid) build_tree(E);
			refit(E, 0);
			std::vector<i32> o = order(E, W.n_leaves);
			for (i32 t : o) t_collide(W, t, stack.data(), 1, E->st, any);
		}
		if (W.ctr[CT_ERROR]) return -(int)W.ctr[CT_ERROR];
		/* the resolve kernels run speculatively behind the narrow phase (no-ops without proposals) */
		for (i32 e = 0; e < n; e++) t_decide(W, e);
		{
			std::vector<i32> o = order(E, n);
			for (i32 e : o) t_push(W, e, E->st);
		}
		if (W.ctr[CT_SERIAL_N]) t_push_serial(W, E->st);
		for (i32 e = 0; e < n; e++) t_post(W, e);
		if (!any) break;
		E->iterations++;
		cache_state = W.ctr[CT_CACHE_INVALID] ? 1 : 2;
		it++;
		if (++guard > 100000) return BT_ERR_OVERFLOW;
	}
	W.flush_stage = 1;
	flush1(E, W.ctr[CT_FLUSH_HINT] != 0);
	W.flush_stage = 0;
	if (W.ctr[CT_FLUSH_HINT] && E->full_flush_left < 1) E->full_flush_left = 1;
	if (W.ctr[CT_ANY_DELTA]) { /* flush2 ent.cpp:910 */
		for (i32 e = 0; e < n; e++) t_flush2_apply(W, e);
		for (i32 e = 0; e < n; e++) t_zero_deltas(W, e);
		W.ctr[CT_ANY_DELTA] = 0;
	}
	if (W.ctr[CT_ERROR]) return -(int)W.ctr[CT_ERROR];
	if (W.ctr[CT_NEED_REBUILD]) {
		E->tree_valid = false; W.ctr[CT_NEED_REBUILD] = 0;
		if (W.small_mode) { W.ctr[CT_CACHE_INVALID] = 1; W.ctr[CT_CACHE_FULL] = 1; }
	}
	return 0;
}

int bte_step(Emu *E, uint32_t phases) {
	DevWorld &W = E->W;
	if (phases & BT_PHASE_WORLD_RULES) {
		for (i32 e = 0; e < W.n; e++) t_world_rules(W, e);
		if (E->full_flush_left < 1) E->full_flush_left = 1;
	}
	E->ctrls_fused = (phases & BT_PHASE_CTRLS) && (phases & BT_PHASE_PHYSICS) && !E->deltas_pending && !(phases & BT_PHASE_TICKS);
	if ((phases & BT_PHASE_CTRLS) && !E->ctrls_fused) for (i32 e = 0; e < W.n; e++) t_pre(W, e, phases);
	if (phases & BT_PHASE_TICKS) {
		std::vector<i32> o = order(E, W.n);
		for (i32 e : o) t_tick(W, e);
		for (i32 g = 0; g < W.n_worlds; g++) t_tick_serial(W, g);
		E->deltas_pending = true;
		if (E->full_flush_left < 1) E->full_flush_left = 1;
	}
	if (phases & BT_PHASE_PREP) flush1(E, false);
	if ((phases & BT_PHASE_GRAVITY) && !(phases & BT_PHASE_PHYSICS)) {
		for (i32 e = 0; e < W.n; e++) t_gravity(W, e);
		E->deltas_pending = true;
	}
	if (phases & BT_PHASE_PHYSICS) {
		int r = do_physics(E, (phases & BT_PHASE_GRAVITY) != 0);
		if (r) return r;
		E->ticks++;
	}
	if (phases & BT_PHASE_FINISH) W.ctr[CT_DEAD_N] = 0;
	if (W.ctr[CT_ERROR]) return -(int)W.ctr[CT_ERROR];
	return 0;
}

void bte_hash(Emu *E, uint64_t *out) {
	DevWorld &W = E->W;
	for (int g = 0; g < W.n_worlds; g++) out[g] = 0;
	std::vector<u64> cnt(W.n_worlds, 0);
	for (i32 e = 0; e < W.n; e++)
		if (W.inList[e]) { out[W.world[e]] += bt_ent_hash(W, e); cnt[W.world[e]]++; }
	for (int g = 0; g < W.n_worlds; g++) out[g] += bt_mix64(cnt[g]);
}

void bte_set_option(Emu *E, const char *name, int64_t v) {
	if (!strcmp(name, "pair_cache")) E->pair_cache = v != 0;
	if (!strcmp(name, "allow_unsupported")) E->W.lenient = v != 0;
	if (!strcmp(name, "cache_margin")) E->W.margin = (i32)v;
	if (!strcmp(name, "cache_horizon")) E->W.horizon = (i32)v;
	if (!strcmp(name, "small_worlds")) { E->W.small_mode = v != 0 && E->small_ok; if (!v) E->tree_valid = false; }
	if (!strcmp(name, "pair_patch")) E->W.patch = v != 0;
	if (!strcmp(name, "patch_shift")) E->W.patch_shift = (i32)v;
}
int64_t bte_pair_rebuilds(Emu *E) { return E->pair_rebuilds; }
int64_t bte_pair_patches(Emu *E) { return E->pair_patches; }
int64_t bte_events(Emu *E) { return (int64_t)E->W.ctr[CT_EVENT_N]; }

/* events since the last call as {kind, handler, me uid, him uid, arg} rows, ordered like
 * bt_world_drain_events orders them */
int bte_drain_events(Emu *E, int32_t *out, int max) {
	DevWorld &W = E->W;
	int n = (int)W.ctr[CT_EVENT_N];
	std::vector<int4> ev(W.events, W.events + n);
	std::stable_sort(ev.begin(), ev.end(), [](const int4 &a, const int4 &b) {
		int ka = a.x >> 8, kb = b.x >> 8;
		if (ka != kb) return ka < kb;
		return ka == 3 && (u32)a.w > (u32)b.w;
	});
	for (int i = 0; i < n && i < max; i++) {
		out[5 * i] = ev[i].x >> 8; out[5 * i + 1] = ev[i].x & 0xff; out[5 * i + 2] = W.uid[ev[i].y];
		out[5 * i + 3] = W.uid[ev[i].z]; out[5 * i + 4] = ev[i].w;
	}
	W.ctr[CT_EVENT_N] = 0;
	return n;
}

/* bt_world_apply_requests for the emulator; me / other are uids (single world) */
int bte_request(Emu *E, int kind, int32_t me_uid, int32_t other_uid, int32_t a0, int32_t a1, int32_t a2) {
	DevWorld &W = E->W;
	BtReq q = {kind, -1, -1, a0, a1, a2};
	for (i32 e = 0; e < W.n; e++) {
		if (W.uid[e] == me_uid) q.me = e;
		if (W.uid[e] == other_uid) q.other = e;
	}
	if (q.me < 0) return BT_ERR_ARG;
	t_request(W, q);
	if (kind == BT_REQ_VEL || kind == BT_REQ_CENTER) E->deltas_pending = true;
	if (E->full_flush_left < 1) E->full_flush_left = 1;
	return W.ctr[CT_ERROR] ? -(int)W.ctr[CT_ERROR] : 0;
}

/* sliders, wires, holders of every LIVE entity in slot order (compact CSR), for full-state diffs */
int bte_export_misc(Emu *E, int32_t *uid, int32_t *holder_uid, int32_t *n_sliders, int32_t *sliders, int32_t *n_wires,
                    int32_t *wires_uid, uint32_t *btn, int max_ents, int max_items) {
	DevWorld &W = E->W;
	int k = 0, so = 0, wo = 0;
	for (i32 e = 0; e < W.n; e++) {
		if (!W.inList[e]) continue;
		if (k >= max_ents) return -1;
		uid[k] = W.uid[e];
		holder_uid[k] = W.holder[e] >= 0 ? W.uid[W.holder[e]] : -1;
		n_sliders[k] = W.slider_off[e + 1] - W.slider_off[e];
		for (i32 s = W.slider_off[e]; s < W.slider_off[e + 1]; s++) { if (so >= max_items) return -1; sliders[so++] = W.sv[s]; }
		n_wires[k] = W.wire_cnt[e];
		btn[2 * k] = W.btn_v[e]; btn[2 * k + 1] = W.btn_v2[e];
		for (i32 q = 0; q < W.wire_cnt[e]; q++) { if (wo >= max_items) return -1; wires_uid[wo++] = W.uid[W.wire_to[W.wire_off[e] + q]]; }
		k++;
	}
	return k;
}
int64_t bte_unsupported(Emu *E) { return (int64_t)E->W.ctr[CT_UNSUPPORTED]; }
int64_t bte_order_ties(Emu *E) { return (int64_t)E->W.ctr[CT_ORDER_TIES]; }
int64_t bte_unsupported_last(Emu *E) { return (int64_t)E->W.ctr[CT_UNSUP_LAST]; }

void bte_stats(Emu *E, int64_t *out) {
	out[0] = E->ticks; out[1] = E->iterations; out[2] = E->st.cand; out[3] = E->st.pos; out[4] = E->st.push;
	out[5] = 0; out[6] = E->st.crush; out[7] = E->st.fumble;
}

/* centre / vel / liveness of every slot, enough for field-level diffs in tests */
void bte_peek(Emu *E, int32_t *center, int32_t *vel, uint8_t *alive) {
	DevWorld &W = E->W;
	for (i32 e = 0; e < W.n; e++) {
		center[3 * e] = W.center[e].x; center[3 * e + 1] = W.center[e].y; center[3 * e + 2] = W.center[e].z;
		vel[3 * e] = W.vel[e].x; vel[3 * e + 1] = W.vel[e].y; vel[3 * e + 2] = W.vel[e].z;
		alive[e] = W.inList[e];
	}
}

/* raycast over the all-entity tree; origin / winner are live ranks */
int bte_raycast(Emu *E, const int32_t *origin, const int32_t *look, int32_t n_rays, int32_t *winner) {
	DevWorld W = E->W; /* private copy with its own tree arrays */
	std::vector<u64> keys(W.n);
	std::vector<i32> leafEnt(W.n), leafParent(W.n), nodeParent(W.n);
	std::vector<u32> nodeCount(W.n);
	std::vector<i32> wroot(W.n_worlds + 1);
	std::vector<BvhNode> nodes(W.n ? W.n : 1);
	std::vector<Box6> nodeBox(W.n ? W.n : 1), leafBox(W.n ? W.n : 1);
	u32 ctr[CT_WORDS];
	memset(ctr, 0, sizeof ctr);
	W.keys = keys.data(); W.leafEnt = leafEnt.data(); W.leafParent = leafParent.data(); W.nodeParent = nodeParent.data();
	W.nodeCount = nodeCount.data(); W.nodes = nodes.data(); W.nodeBox = nodeBox.data(); W.leafBox = leafBox.data();
	W.ctr = ctr; W.ray_mode = 1; W.wroot = wroot.data();
	Emu tmp = *E;
	tmp.W = W; tmp.shuffle_seed = 0;
	build_tree(&tmp);
	refit(&tmp, 0);
	W = tmp.W;
	std::vector<i32> rank2slot, slot2rank(W.n, -1);
	for (i32 e = 0; e < W.n; e++)
		if (W.inList[e]) { slot2rank[e] = (i32)rank2slot.size(); rank2slot.push_back(e); }
	std::vector<i32> stack(BT_STACK);
	for (i32 i = 0; i < n_rays; i++) {
		i32 lk[3] = {look[3 * i], look[3 * i + 1], look[3 * i + 2]};
		if (origin[i] < 0 || origin[i] >= (i32)rank2slot.size()) { winner[i] = -1; continue; }
		i32 win = t_raycast(W, rank2slot[origin[i]], lk, stack.data(), 1);
		winner[i] = win >= 0 ? slot2rank[win] : -1;
	}
	return ctr[CT_ERROR] ? -(int)ctr[CT_ERROR] : 0;
}

} /* extern "C" */
