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
	/* incremental host binding (mirrors bt_world.cu) */
	i32 cap_ents, cap_sliders, cap_wires;
	i32 root_phase_ctr, sliders_used, wires_used;
	DevArrays arrays;
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

Emu *bte_create_cap(const bt_world_desc *d, int rebuild_every, int shuffle_seed, int cap_ents, int cap_sliders, int cap_wires);
Emu *bte_create(const bt_world_desc *d, int rebuild_every, int shuffle_seed) {
	return bte_create_cap(d, rebuild_every, shuffle_seed, d->n_ents, d->n_sliders, d->n_wires);
}
/* capacities beyond the uploaded world leave room for bt_world_grow */
Emu *bte_create_cap(const bt_world_desc *d, int rebuild_every, int shuffle_seed, int cap_ents, int cap_sliders, int cap_wires) {
	if (cap_ents < d->n_ents || cap_sliders < d->n_sliders || cap_wires < d->n_wires) return NULL;
	Emu *E = new Emu();
	E->cap_ents = cap_ents; E->cap_sliders = cap_sliders; E->cap_wires = cap_wires;
	E->root_phase_ctr = 0;
	memset(&E->W, 0, sizeof(E->W));
	memset(&E->st, 0, sizeof(E->st));
	E->tree_valid = false; E->ticks_since_build = 0; E->ctrls_fused = false; E->full_flush_left = 2; E->scan_flushes_left = 2; E->rebuild_every = rebuild_every > 0 ? rebuild_every : 1;
	E->iterations = 0; E->ticks = 0; E->shuffle_seed = shuffle_seed; E->pair_cache = 1; E->pair_rebuilds = 0; E->pair_patches = 0;
	DevWorld &W = E->W;
	const int nw = d->n_worlds;
	{
	const int n = cap_ents; /* every per-slot array is allocated at capacity */
	A(E, &W.wbegin, nw); A(E, &W.wend, nw); A(E, &W.flip, nw); A(E, &W.gamerules, nw); A(E, &W.rnd, nw); A(E, &W.wflush, nw); A(E, &W.wunsup, nw); A(E, &W.wties, nw);
	A(E, &W.world, n); A(E, &W.uid, n); A(E, &W.center, n); A(E, &W.old, n); A(E, &W.vel, n); A(E, &W.radius, n);
	A(E, &W.d_center, n); A(E, &W.d_vel, n); A(E, &W.type, n); A(E, &W.collide, n); A(E, &W.ntype, n); A(E, &W.ncollide, n);
	A(E, &W.friction, n); A(E, &W.holder, n); A(E, &W.holdFlags, n); A(E, &W.holdRoot, n); A(E, &W.holdee, n);
	A(E, &W.LLn, n); A(E, &W.LLp, n); A(E, &W.newHolder, n); A(E, &W.newHoldFlags, n);
	A(E, &W.hphys, n); A(E, &W.hpair, n); A(E, &W.htick, n); A(E, &W.btn_v, n); A(E, &W.btn_v2, n);
	A(E, &W.newDrop, n); A(E, &W.dead_max0, n); A(E, &W.dead_max1, n); A(E, &W.fullFlush, n); A(E, &W.dead, n);
	A(E, &W.inList, n); A(E, &W.needsCollision, n); A(E, &W.hasBox, n); A(E, &W.ready, n);
	A(E, &W.forced, n); A(E, &W.forcedHoldees, n); A(E, &W.pend, n);
	A(E, &W.slider_off, n + 1); A(E, &W.sv, cap_sliders); A(E, &W.smin, cap_sliders); A(E, &W.smax, cap_sliders);
	A(E, &W.wire_off, n + 1); A(E, &W.wire_cnt, n); A(E, &W.wire_to, cap_wires + n);
	A(E, &W.wireAdd, n); A(E, &W.wireRm, n); A(E, &W.wire_cap, n); A(E, &W.rootKey, n); A(E, &W.wrootseq, nw);
	W.wq_cap = 2 * n + 1024;
	A(E, &W.wq_next, W.wq_cap); A(E, &W.wq_to, W.wq_cap);
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
	}
	const int n = d->n_ents;
	W.n = n; W.n_worlds = nw;
	{
		void **b = E->arrays.base;
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
		deltas = deltas || (d->new_holder[i] >= 0 && (d->new_hold_flags[i] & HOLD_FREEZE)); /* ent.cpp:176-181 */
		if (d->holder[i] < 0) W.rootKey[i] = (u64)(n - (d->root_order[i] >= 0 ? d->root_order[i] : n - 1 - i));
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
			W.wire_cap[i] = W.wire_cnt[i] + (d->handlers[(size_t)BT_H_NUM * i + BT_H_ONPICKUP] == 1 ? 1 : 0);
			o += W.wire_cap[i];
		}
		W.wire_off[n] = o;
		E->wires_used = o; E->sliders_used = d->n_sliders;
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
	W.root_phase = ++E->root_phase_ctr;
	for (i32 g = 0; g < W.n_worlds; g++) t_flush_serial(W, g, E->st);
	E->st.rk_ordered = 0;
	for (i32 e = 0; e < W.n; e++) t_flush_misc(W, e);
	W.ctr[CT_WQ_N] = 0;
}

static int do_physics_iter(Emu *E, bool gravity) {
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
		W.root_phase = ++E->root_phase_ctr;
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
	return 0;
}

/* what follows invokeOnCrush: flush1; flush2 (ent.cpp:909-910) */
static int do_physics_end(Emu *E) {
	DevWorld &W = E->W;
	const i32 n = W.n;
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
	E->ctrls_fused = (phases & BT_PHASE_CTRLS) && (phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_ITER)) && !E->deltas_pending && !(phases & BT_PHASE_TICKS);
	if ((phases & BT_PHASE_CTRLS) && !E->ctrls_fused) for (i32 e = 0; e < W.n; e++) t_pre(W, e, phases);
	if (phases & BT_PHASE_TICKS) {
		std::vector<i32> o = order(E, W.n);
		for (i32 e : o) t_tick(W, e);
		for (i32 g = 0; g < W.n_worlds; g++) t_tick_serial(W, g);
		E->deltas_pending = true;
		if (E->full_flush_left < 1) E->full_flush_left = 1;
	}
	if (phases & BT_PHASE_PREP) flush1(E, false);
	if ((phases & BT_PHASE_GRAVITY) && !(phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_ITER))) {
		for (i32 e = 0; e < W.n; e++) t_gravity(W, e);
		E->deltas_pending = true;
	}
	if (phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_ITER)) {
		int r = do_physics_iter(E, (phases & BT_PHASE_GRAVITY) != 0);
		if (r) return r;
	}
	if (phases & (BT_PHASE_PHYSICS | BT_PHASE_PHYSICS_END)) {
		int r = do_physics_end(E);
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

/* ---------------------------------------------------------------------------------------------
 * The part of the C ABI (include/bittoss_b200.h) the host binding host/bt_glue.cpp is built on, served by
 * the emulator: lets the `-m "not gpu"` tests run the reference's own host code (tick handlers, players,
 * spawners) against the kernel bodies on a machine without a GPU.  Same entry-point names as
 * libbittoss_b200.so; the glue is bound to one library or the other at run time (bt_glue_bind).
 * --------------------------------------------------------------------------------------------- */
struct bt_world {
	Emu *E;
	int cap_ents, cap_sliders, cap_wires, cap_worlds;
	bool poked;
};
static char emu_err[256] = "";
typedef long long i64;

extern "C" {

const char *bt_version(void) { return "bittoss_b200 kernel-body emulator (tests only)"; }
const char *bt_last_error(void) { return emu_err; }

int bt_world_create(bt_world **out, int device, int32_t max_ents, int32_t max_sliders, int32_t max_wires, int32_t max_worlds) {
	(void)device;
	if (!out || max_ents < 0 || max_worlds < 1) return BT_ERR_ARG;
	bt_world *w = new bt_world();
	w->E = NULL; w->cap_ents = max_ents; w->cap_sliders = max_sliders; w->cap_wires = max_wires; w->cap_worlds = max_worlds; w->poked = false;
	*out = w;
	return BT_OK;
}
void bt_world_destroy(bt_world *w) {
	if (!w) return;
	bte_destroy(w->E);
	delete w;
}
int bt_world_upload(bt_world *w, const bt_world_desc *d) {
	if (!w || !d) return BT_ERR_ARG;
	if (d->n_ents > w->cap_ents || d->n_sliders > w->cap_sliders || d->n_wires > w->cap_wires || d->n_worlds > w->cap_worlds) return BT_ERR_CAPACITY;
	bte_destroy(w->E);
	w->E = bte_create_cap(d, 16, 0, w->cap_ents, w->cap_sliders, w->cap_wires);
	return w->E ? BT_OK : BT_ERR_ARG;
}
int bt_world_set_option(bt_world *w, const char *name, int64_t value) {
	if (!w || !w->E) return BT_ERR_ARG;
	if (!strcmp(name, "shuffle_seed")) { w->E->shuffle_seed = (int)value; return BT_OK; }
	bte_set_option(w->E, name, value);
	return BT_OK;
}
int bt_world_step(bt_world *w, uint32_t phases) {
	if (!w || !w->E) return BT_ERR_ARG;
	int r = bte_step(w->E, phases);
	if (r) snprintf(emu_err, sizeof emu_err, "emulated step failed: %d", r);
	return r;
}
int bt_world_hash(bt_world *w, uint64_t *out) { if (!w || !w->E) return BT_ERR_ARG; bte_hash(w->E, out); return BT_OK; }
int bt_world_hash_full(bt_world *w, uint64_t *out) {
	if (!w || !w->E) return BT_ERR_ARG;
	DevWorld &W = w->E->W;
	std::vector<u64> cnt(W.n_worlds, 0);
	for (int g = 0; g < W.n_worlds; g++) out[g] = 0;
	for (i32 e = 0; e < W.n; e++)
		if (W.inList[e]) { out[W.world[e]] += bt_ent_hash_full(W, e); cnt[W.world[e]]++; }
	for (int g = 0; g < W.n_worlds; g++) {
		u64 h = bt_mix64((u64)(u32)W.flip[g] + 0x9E3779B97F4A7C15ULL);
		h = bt_mix64(h + (u64)(u32)W.rnd[g] + 0x9E3779B97F4A7C15ULL);
		h = bt_mix64(h + (u64)(u32)W.gamerules[g] + 0x9E3779B97F4A7C15ULL);
		out[g] += bt_mix64(cnt[g]) + h;
	}
	return BT_OK;
}
int bt_world_drain_events(bt_world *w, bt_event *out, int32_t max, int32_t *n) {
	if (!w || !w->E || !n) return BT_ERR_ARG;
	DevWorld &W = w->E->W;
	int cnt = (int)W.ctr[CT_EVENT_N];
	if (cnt > max) { *n = cnt; return BT_ERR_CAPACITY; }
	std::vector<int32_t> rows((size_t)5 * (cnt ? cnt : 1));
	std::vector<int4> ev(W.events, W.events + cnt);
	std::stable_sort(ev.begin(), ev.end(), [](const int4 &a, const int4 &b) {
		int ka = a.x >> 8, kb = b.x >> 8;
		if (ka != kb) return ka < kb;
		return ka == 3 && (u32)a.w > (u32)b.w;
	});
	for (int i = 0; i < cnt; i++) {
		out[i].kind = ev[i].x >> 8; out[i].handler = ev[i].x & 0xff; out[i].world = W.world[ev[i].y];
		out[i].me_uid = W.uid[ev[i].y]; out[i].him_uid = W.uid[ev[i].z]; out[i].arg = ev[i].w;
	}
	W.ctr[CT_EVENT_N] = 0;
	*n = cnt;
	return BT_OK;
}

static int emu_elem_bytes(int a) {
	if (a >= BT_A_CENTER && a <= BT_A_RADIUS) return 16;
	if (a >= BT_A_NEWDROP && a <= BT_A_INLIST) return 1;
	if (a == BT_A_ROOTKEY || a == BT_A_DEADLIST) return 8;
	return 4;
}
int bt_world_extent(bt_world *w, int32_t a, int32_t *count) {
	if (!w || !w->E || !count || a < 0 || a >= BT_A_NUM) return BT_ERR_ARG;
	Emu *E = w->E;
	if (a >= BT_A_SV && a <= BT_A_SMAX) *count = E->sliders_used;
	else if (a == BT_A_WIRE_TO) *count = E->wires_used;
	else if (a >= BT_A_W_FLIP && a <= BT_A_W_RAND) *count = E->W.n_worlds;
	else if (a == BT_A_SLIDER_OFF) *count = E->W.n + 1;
	else if (a == BT_A_DEADLIST) *count = (i32)E->W.ctr[CT_DEAD_N];
	else *count = E->W.n;
	return BT_OK;
}
int bt_world_pull(bt_world *w, int32_t a, int32_t first, int32_t count, void *dst) {
	i32 ext = 0;
	int r = bt_world_extent(w, a, &ext);
	if (r) return r;
	if (first < 0 || count < 0 || first + count > ext) return BT_ERR_ARG;
	const size_t eb = (size_t)emu_elem_bytes(a);
	memcpy(dst, (const char *)w->E->arrays.base[a] + eb * (size_t)first, eb * (size_t)count);
	return BT_OK;
}
int64_t bt_world_next_root_phase(bt_world *w) { return w && w->E ? (int64_t)++w->E->root_phase_ctr : (int64_t)BT_ERR_ARG; }
int bt_world_tree_edits(bt_world *w, uint32_t *count) {
	if (!w || !w->E || !count) return BT_ERR_ARG;
	*count = w->E->W.ctr[CT_TREE_EDITS];
	return BT_OK;
}
static void emu_host_wrote(Emu *E) {
	E->deltas_pending = true;
	if (E->full_flush_left < 2) E->full_flush_left = 2;
	if (E->scan_flushes_left < 2) E->scan_flushes_left = 2;
}
static void emu_leafset_changed(Emu *E) {
	E->tree_valid = false;
	E->W.ctr[CT_CACHE_INVALID] = 1; E->W.ctr[CT_CACHE_FULL] = 1;
}
int bt_world_grow(bt_world *w, int32_t n_new, const int32_t *n_sliders, const int32_t *wire_cap, int32_t world, int32_t *first_slot) {
	if (!w || !w->E || n_new < 0 || !first_slot) return BT_ERR_ARG;
	Emu *E = w->E;
	DevWorld &W = E->W;
	if (W.n_worlds > 1) return BT_ERR_UNSUPPORTED;
	*first_slot = W.n;
	if (!n_new) return BT_OK;
	i64 ns = 0, nwc = 0;
	for (i32 i = 0; i < n_new; i++) { ns += n_sliders[i]; nwc += wire_cap[i]; }
	if ((i64)W.n + n_new > E->cap_ents || E->sliders_used + ns > E->cap_sliders || E->wires_used + nwc > (i64)E->cap_wires + E->cap_ents) return BT_ERR_CAPACITY;
	i32 s = E->sliders_used, q = E->wires_used;
	const i32 first = W.n;
	for (i32 i = 0; i < n_new; i++) {
		const i32 e = first + i;
		t_grow_init(W, e, world);
		W.slider_off[e] = s; W.wire_off[e] = q; W.wire_cap[e] = wire_cap[i];
		for (i32 k = 0; k < n_sliders[i]; k++) { W.sv[s + k] = 0; W.smin[s + k] = INT32_MAX; W.smax[s + k] = INT32_MIN; }
		s += n_sliders[i]; q += wire_cap[i];
	}
	W.slider_off[first + n_new] = s;
	W.n = first + n_new;
	W.wend[world] = W.n;
	E->sliders_used = s; E->wires_used = q;
	if (W.n > BT_SMALL_WORLD) { E->small_ok = false; W.small_mode = 0; }
	emu_leafset_changed(E);
	emu_host_wrote(E);
	return BT_OK;
}
int bt_world_poke(bt_world *w, const bt_poke *pokes, int32_t n) {
	if (!w || !w->E || n < 0) return BT_ERR_ARG;
	Emu *E = w->E;
	DevWorld &W = E->W;
	std::vector<bt_poke> raw, ops;
	bool leafset = false;
	i32 last_rebox = -1;
	for (i32 i = 0; i < n; i++) {
		const bt_poke &p = pokes[i];
		if (p.array >= 0 && p.array < BT_A_NUM) {
			i32 ext = 0;
			if (p.array == BT_A_DEADLIST || p.array == BT_A_SLIDER_OFF || p.array == BT_A_WIRE_OFF || p.array == BT_A_WIRE_CAP) return BT_ERR_ARG;
			bt_world_extent(w, p.array, &ext);
			if (p.index < 0 || p.index >= ext) { snprintf(emu_err, sizeof emu_err, "poke %d: element %d outside the %d of array %d", i, p.index, ext, p.array); return BT_ERR_ARG; }
			raw.push_back(p);
			if ((p.array == BT_A_TYPE || p.array == BT_A_COLLIDE) && last_rebox != p.index) {
				bt_poke o = p; o.array = BT_OP_REBOX; ops.push_back(o); last_rebox = p.index; leafset = true;
			}
			if (p.array == BT_A_INLIST || p.array == BT_A_DEAD) leafset = true;
		} else if (p.array == BT_OP_SPAWN || p.array == BT_OP_UWIRE || p.array == BT_OP_UUNWIRE) {
			if (p.index < 0 || p.index >= W.n) return BT_ERR_ARG;
			ops.push_back(p);
			if (p.array == BT_OP_SPAWN) leafset = true;
		} else return BT_ERR_ARG;
	}
	for (i32 t = 0; t < (i32)raw.size(); t++) t_poke_raw(W, E->arrays, raw.data(), t, (i32)raw.size());
	for (const bt_poke &p : ops) t_poke_op(W, p);
	if (leafset) emu_leafset_changed(E);
	emu_host_wrote(E);
	return W.ctr[CT_ERROR] ? -(int)W.ctr[CT_ERROR] : BT_OK;
}
int bt_world_set_wires(bt_world *w, int32_t slot, const int32_t *targets, int32_t n, int32_t reserve) {
	if (!w || !w->E || slot < 0 || slot >= w->E->W.n || n < 0) return BT_ERR_ARG;
	Emu *E = w->E;
	DevWorld &W = E->W;
	if (n + reserve > W.wire_cap[slot]) {
		const i32 cap = n + reserve + 2;
		if ((i64)E->wires_used + cap > (i64)E->cap_wires + E->cap_ents) return BT_ERR_CAPACITY;
		W.wire_off[slot] = E->wires_used; W.wire_cap[slot] = cap;
		E->wires_used += cap;
	}
	for (i32 k = 0; k < n; k++) W.wire_to[W.wire_off[slot] + k] = targets[k];
	W.wire_cnt[slot] = n;
	emu_host_wrote(E);
	return BT_OK;
}

} /* extern "C" */
