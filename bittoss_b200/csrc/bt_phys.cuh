/*
 * bt_phys.cuh — per-thread bodies of the tick kernels.
 *
 * Each t_*() function is what ONE CUDA thread does; bt_world.cu wraps them in
 * __global__ kernels (grid = ceil(n/128) x 128, sm_100a).  Citations are to the
 * sboerwinkle/bittoss reference the behaviour is matched against.
 *
 * How the sequential reference loop maps onto parallel kernels:
 *   doIteration (ent.cpp:437-462)  -> t_collide: one thread per active leaf walks the
 *       LBVH; every narrow-phase hit PROPOSES a collision to the hold root it would
 *       move, with a 64-bit CAS loop that keeps the maximum under collisionBetter +
 *       canonical tie-break.  A maximum does not depend on arrival order, so the
 *       result is deterministic although the order of atomics is not.
 *   shouldResolveCollision (ent.cpp:631-672) -> t_decide: read-only walk of pend[].
 *   push loop (ent.cpp:879-882)   -> t_push: one thread per ready root.  A ready root
 *       owns its hold tree exclusively for the iteration (its buddy's tree is either
 *       static this iteration or its mutual partner), except for tied cycles, which
 *       t_conflict routes to the serial lane (t_push_serial, canonical order).
 */
#ifndef BT_PHYS_CUH
#define BT_PHYS_CUH

#include "bt_types.h"
#include <math.h>

#if defined(BT_EMU)
static inline u64 bt_cas64(u64 *p, u64 cmp, u64 val) { u64 o = *p; if (o == cmp) *p = val; return o; }
static inline u32 bt_add32(u32 *p, u32 v) { u32 o = *p; *p += v; return o; }
static inline void bt_add64(u64 *p, u64 v) { *p += v; }
static inline void bt_min32(i32 *p, i32 v) { if (v < *p) *p = v; }
static inline void bt_max32(i32 *p, i32 v) { if (v > *p) *p = v; }
static inline void bt_or32(u32 *p, u32 v) { *p |= v; }
static inline u32 bt_xchg32(u32 *p, u32 v) { u32 o = *p; *p = v; return o; }
static inline i32 bt_ldv32(const i32 *p) { return *p; }
static inline u64 bt_ld64(const u64 *p) { return *p; }
static inline int4 bt_ldcg4(const int4 *p) { return *p; }
static inline void bt_fence() {}
static inline int bt_clz32(u32 x) { return x ? __builtin_clz(x) : 32; }
static inline int bt_clz64(u64 x) { return x ? __builtin_clzll(x) : 64; }
static inline int bt_abs(int x) { return x < 0 ? -x : x; }
static inline void bt_prefetch(const void *) {}
#else
static BT_HD u64 bt_cas64(u64 *p, u64 cmp, u64 val) { return atomicCAS(p, cmp, val); }
static BT_HD u32 bt_add32(u32 *p, u32 v) { return atomicAdd(p, v); }
static BT_HD void bt_add64(u64 *p, u64 v) { atomicAdd(p, v); }
static BT_HD void bt_min32(i32 *p, i32 v) { atomicMin(p, v); }
static BT_HD void bt_max32(i32 *p, i32 v) { atomicMax(p, v); }
static BT_HD void bt_or32(u32 *p, u32 v) { atomicOr(p, v); }
static BT_HD u32 bt_xchg32(u32 *p, u32 v) { return atomicExch(p, v); }
static BT_HD i32 bt_ldv32(const i32 *p) { return *(const volatile i32 *)p; }
static BT_HD u64 bt_ld64(const u64 *p) { return *(const volatile u64 *)p; }
static BT_HD int4 bt_ldcg4(const int4 *p) { return __ldcg(p); } /* L2 load: another SM wrote it in this kernel */
static BT_HD void bt_fence() { __threadfence(); }
static BT_HD int bt_clz32(u32 x) { return __clz((int)x); }
static BT_HD int bt_clz64(u64 x) { return __clzll((long long)x); }
static BT_HD int bt_abs(int x) { return x < 0 ? -x : x; }
/* start a line on its way to L1 without waiting for it */
#ifdef BT_NO_PREFETCH /* bt_small.cu: the arrays it would touch are staged in shared memory there */
static BT_HD void bt_prefetch(const void *) {}
#else
static BT_HD void bt_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif
#endif

struct LocalStats {
	u32 cand, pos, push, crush, fumble;
	/* where the next fumble of this thread puts the new root in gs->rootEnts order (see bt_fumble):
	 * rk_ordered = 0: rk_base | rk_local++ (a push: base = phase, pushing root); 1: phase | the world's counter */
	u32 rk_ordered, rk_local;
	u64 rk_base;
};
#define BT_ROOTKEY_SHIFT 37
static BT_HD u64 bt_push_rootkey(const DevWorld &W, i32 root) {
	return ((u64)(u32)W.root_phase << BT_ROOTKEY_SHIFT) | ((u64)(u32)(BT_MAX_ENTS - 1 - root) << 8);
}

static BT_HD void bt_check_cover(const DevWorld &W, i32 e, const int4 &o, const int4 &c);
static BT_HD void bt_upickup(const DevWorld &W, i32 p, i32 e, i32 holdFlags);

/* ------------------------------------------------------------------------- */
/* helpers                                                                   */

/* needsVelbox ent.cpp:75-77 */
static BT_HD int bt_needs_box(u32 type, u32 collide) { return collide || (type & T_COLLIDABLE); }

/* Something the device cannot answer like the reference happened for entity e: a handler without
 * a device implementation would have run (lenient uploads only; kinds 0 push, 1 pushed, 2 entPair,
 * 3 crush, 5 wire request overflow, 6 tick), or the reference itself leaves defined behaviour
 * (kind 0 id 6: teleport while pad / destination are being pushed; kind 7: pickup by a dead entity).
 * Counted, never silent. */
static BT_HD void bt_unsup(const DevWorld &W, i32 e, int kind, int id) {
	W.ctr[CT_UNSUP_LAST] = (u32)(kind << 8 | id);
	bt_add32(&W.ctr[CT_UNSUPPORTED], 1u);
	bt_add32(&W.wunsup[W.world[e]], 1u);
}

/* boundVec ent.cpp:20-30 (int64 multiply, truncating division) */
static BT_HD void bt_bound_vec(i32 *v, i32 bound, int len) {
	i32 mx = bound;
	for (int i = 0; i < len; i++) {
		i32 x = bt_abs(v[i]);
		if (x > mx) mx = x;
	}
	if (mx == bound) return;
	for (int i = 0; i < len; i++) v[i] = (i32)((long long)bound * v[i] / mx);
}

/* pre-order walk of the hold subtree of `r` without a stack (holdee / LL.n / holder links) */
#define BT_FOR_SUBTREE(W, r, it, ...)                                    \
	do {                                                                 \
		i32 it = (r);                                                    \
		while (1) {                                                      \
			__VA_ARGS__;                                                     \
			if ((W).holdee[it] >= 0) { it = (W).holdee[it]; continue; }  \
			while (it != (r) && (W).LLn[it] < 0) it = (W).holder[it];    \
			if (it == (r)) break;                                        \
			it = (W).LLn[it];                                            \
		}                                                                \
	} while (0)

/* uSlider entUpdaters.c:58-68 */
static BT_HD void bt_uslider(const DevWorld &W, i32 e, int ix, i32 value) {
	int ns = W.slider_off[e + 1] - W.slider_off[e];
	if (ix < 0 || ix >= ns) return;
	W.fullFlush[e] = 1;
	W.ctr[CT_FLUSH_HINT] = 1;
	int s = W.slider_off[e] + ix;
	if (value > W.smax[s]) W.smax[s] = value;
	if (value < W.smin[s]) W.smin[s] = value;
}
/* getSlider entGetters.c:42-48 */
static BT_HD i32 bt_getslider(const DevWorld &W, i32 e, int ix) {
	int ns = W.slider_off[e + 1] - W.slider_off[e];
	if (ix < 0 || ix >= ns) return 0;
	return W.sv[W.slider_off[e] + ix];
}

/* ------------------------------------------------------------------------- */
/* handlers, selected by the catalog IDs of handlerRegistrar.h:79-187        */

/* whoMoves_t: move_me common.c:48; explosion.cpp:8; player.c:84; stackem.c:9-22;
 * material.c:5-24.  0 = unset -> MOVE_ME (ent.cpp:220-225), 3 aliases 1 (platform.c:56). */
static BT_HD int bt_who_moves(int id, u32 ta, u32 tb, int axis, int dir) {
	switch (id) {
	case 2: return (tb & T_TERRAIN) ? MOVE_ME : MOVE_HIM;
	case 4: return (tb & T_TERRAIN) ? MOVE_ME : MOVE_BOTH;
	case 5: {
		if (tb & T_TERRAIN) return MOVE_ME;
		int goodTeam = (TEAM_MASK & tb & ta) != 0;
		if (axis == 2) {
			if (dir == -1 && ((tb & T_HEAVY) || goodTeam)) return MOVE_ME;
			return MOVE_HIM;
		}
		return goodTeam ? MOVE_BOTH : MOVE_HIM;
	}
	case 6:
		if (tb & T_TERRAIN) return MOVE_ME;
		if (axis == 2) return dir == -1 ? MOVE_ME : MOVE_HIM;
		return MOVE_BOTH;
	case 7: return (tb & T_HEAVY) ? MOVE_ME : MOVE_HIM;
	case 8:
		if (dir == 1 && axis == 2 && !(tb & (T_TERRAIN | T_HEAVY))) return MOVE_BOTH;
		return MOVE_ME;
	default: return MOVE_ME;
	}
}

/* which handler IDs have a device implementation (bt_world_upload checks these) */
#define BT_SUPPORTED_WHOMOVES 0x1ffu /* 0..8 */
#define BT_SUPPORTED_PUSH 0x3ffu     /* the whole catalog (handlerRegistrar.h:163-175); TEAMSELECT / LOADBLOCK as events */
#define BT_SERIAL_PUSH 0x3f2u        /* the ones that run in the serial lane (order-sensitive side effects): all but NIL, LOGIC, SEAT */
#define TICK_BOTTLE 37
#define TICK_PLAYER 7                /* handlerRegistrar.h:93-101 */
#define BT_SUPPORTED_PUSHED 0x12fu   /* NIL, PUDDLEJUMP, FRAGILE, FLAG, PLAYER, BULLET */
#define BT_SUPPORTED_PAIR 0x03fu     /* NIL, PICKUP_SEAT, FUMBLE_SEAT_OLD, FUMBLED_DECOR, FUMBLE_SEAT, PICKUP_PLAYER */
#define BT_SUPPORTED_CRUSH 0x007u    /* NIL; FLAG and PLATFORM create entities (explode): delivered as host events */

/* player_push player.c:108-165 (sliders s_equip_processed=10, s_editmode=11).  Returns 1 =
 * "I pick him up".  Only ever runs in the serial lane (see t_push): it reads and edits the
 * player's own holdee list and flags. */
static BT_HD int bt_player_push(const DevWorld &W, i32 me, i32 him) {
	u32 th = W.type[him];
	if (!(th & EQUIP_MASK)) return 0;
	if ((W.btn_v[me] >> 8) & 0xff) return 0; /* getButton(me, 1): shift inhibits pickups */
	if (bt_getslider(W, me, 10) & 1) return 0;
	if (bt_getslider(W, me, 11)) return 0;
	if (W.holder[him] >= 0 || W.holder[me] >= 0) return 0;
	u32 tEquip = th & T_EQUIP, tEquipSm = th & T_EQUIP_SM;
	i32 existing = -1;
	for (i32 h = W.holdee[me]; h >= 0; h = W.LLn[h]) {
		u32 heldEquip = W.type[h] & T_EQUIP, heldEquipSm = W.type[h] & T_EQUIP_SM;
		if (heldEquip) {
			if (heldEquipSm || tEquip || existing >= 0) return 0;
			existing = h;
		} else if (heldEquipSm) {
			if (existing >= 0) return 0;
			if (tEquip && tEquipSm) return 0;
			existing = h;
		}
	}
	if (!(tEquip && tEquipSm)) {
		int hand = existing >= 0 && !bt_getslider(W, existing, 0);
		bt_uslider(W, him, 0, hand);
	}
	W.sv[W.slider_off[me] + 10] |= 1; /* direct write, player.c:157 */
	return 1;
}

/* uWire / uUnwire entUpdaters.c:66-73: one pending request of each kind per entity and flush
 * (a seat wires / unwires its single rider); a second, different one cannot be represented */
/* The pending lists e->wiresAdd / e->wiresRm (ent.h:121-122): almost always empty or one entry, kept in
 * the word itself; a second, different target moves the list into the pool (request order kept; a repeated
 * target is dropped, flushMisc would skip it anyway: `if (!e->wires.has(w))`, ent.cpp:695-697). */
static BT_HD void bt_wq_append(const DevWorld &W, i32 e, i32 *slot, i32 w, int kind) {
	i32 v = *slot;
	if (v == -1) { *slot = w; return; }
	if (v >= 0) {
		if (v == w) return;
		u32 a = bt_add32(&W.ctr[CT_WQ_N], 2u);
		if ((i32)a + 2 > W.wq_cap) { bt_unsup(W, e, 5, kind); return; }
		W.wq_to[a] = v; W.wq_next[a] = (i32)a + 1;
		W.wq_to[a + 1] = w; W.wq_next[a + 1] = -1;
		*slot = -2 - (i32)a;
		return;
	}
	i32 h = -2 - v;
	while (1) {
		if (W.wq_to[h] == w) return;
		if (W.wq_next[h] < 0) break;
		h = W.wq_next[h];
	}
	u32 a = bt_add32(&W.ctr[CT_WQ_N], 1u);
	if ((i32)a + 1 > W.wq_cap) { bt_unsup(W, e, 5, kind); return; }
	W.wq_to[a] = w; W.wq_next[a] = -1;
	W.wq_next[h] = (i32)a;
}
static BT_HD void bt_uwire(const DevWorld &W, i32 e, i32 w) {
	W.fullFlush[e] = 1;
	W.ctr[CT_FLUSH_HINT] = 1;
	bt_wq_append(W, e, &W.wireAdd[e], w, 0);
}
static BT_HD void bt_uunwire(const DevWorld &W, i32 e, i32 w) {
	W.ctr[CT_FLUSH_HINT] = 1;
	bt_wq_append(W, e, &W.wireRm[e], w, 1);
}
/* walk a pending list in request order: BT_WQ_FOR(W, word, t) { ... t ... } */
static BT_HD i32 bt_wq_first(const DevWorld &W, i32 word, i32 &h) {
	h = -1;
	if (word >= -1) return word;
	h = -2 - word;
	return W.wq_to[h];
}
static BT_HD i32 bt_wq_next(const DevWorld &W, i32 &h) {
	if (h < 0) return -1;
	h = W.wq_next[h];
	return h >= 0 ? W.wq_to[h] : -1;
}
#define BT_WQ_FOR(W, word, t) for (i32 _h, t = bt_wq_first(W, word, _h); t >= 0; t = bt_wq_next(W, _h))

/* teleport_push teleport.c:12-37.  Serial lane only.  It reads the pad's and the destination's
 * position while other pushes of the same iteration may move them; the reference interleaves
 * those pushes in root order, the serial lane runs after them, so that case is refused. */
static BT_HD int bt_teleport_push(const DevWorld &W, i32 me, i32 him) {
	if (W.type[him] & (T_TERRAIN | T_OBSTACLE | T_DEBRIS)) bt_or32(&W.btn_v2[me], 1u);
	i32 dest = -1;
	for (i32 k = 0; k < W.wire_cnt[me]; k++) {
		i32 w = W.wire_to[W.wire_off[me] + k];
		if ((W.btn_v[w] >> 16) & 0xffu) {
			if (dest >= 0) return 0; /* "Behavior with multiple teleport destinations active is undefined!" */
			dest = w;
		}
	}
	if (dest < 0) return 0;
	if (W.ready[W.holdRoot[dest]] || W.ready[W.holdRoot[me]]) bt_unsup(W, me, 0, 6);
	i32 r = W.holdRoot[him];
	int4 dc = W.center[dest], mc = W.center[me], dv = W.vel[dest], mv = W.vel[me];
	int4 a = W.d_center[r], b = W.d_vel[r];
	a.x += dc.x - mc.x; a.y += dc.y - mc.y; a.z += dc.z - mc.z; /* uCenter(him->holdRoot, ...) */
	b.x += dv.x - mv.x; b.y += dv.y - mv.y; b.z += dv.z - mv.z; /* uVel */
	W.d_center[r] = a; W.d_vel[r] = b;
	W.ctr[CT_ANY_DELTA] = 1;
	return 0;
}

/* bottle_push bottle.cpp:92-116 (id 9) and pj_push puddlejumper.cpp:67-78 (id 8): both pour a
 * bottle (`him->tick == bottle_tick`) into `me` with direct slider writes, then bottleUpdate
 * (bottle.cpp:118-130).  Serial lane only.  Sliders: bottle hand 0 type 1 amt 2 max 3; puddlejumper
 * fuel 0 fuel_max 1. */
static BT_HD int bt_bottle_push(const DevWorld &W, int id, i32 me, i32 him) {
	if ((W.htick[him] & 0xffu) != TICK_BOTTLE) return 0;
	i32 *ms = &W.sv[W.slider_off[me]], *hs = &W.sv[W.slider_off[him]];
	int mn = W.slider_off[me + 1] - W.slider_off[me], hn = W.slider_off[him + 1] - W.slider_off[him];
	if (hn < 3 || mn < (id == 9 ? 4 : 2)) { bt_unsup(W, me, 0, id); return 0; } /* the reference would read out of bounds */
	int poured = 0;
	if (id == 8) {
		if (hs[1] != 1) return 0; /* BOTTLE_FUEL */
		i32 capacity = ms[1] - ms[0], avail = hs[2];
		if (avail < capacity) capacity = avail;
		hs[2] -= capacity;
		ms[0] += capacity;
		poured = 1;
	} else {
		i32 myType = ms[1], hisType = hs[1];
		if (myType == -1 || myType == hisType) {
			i32 capacity = ms[3] - ms[2], hisAmt = hs[2];
			if (hisAmt < capacity) capacity = hisAmt;
			if (capacity <= 0) return 0;
			hs[2] -= capacity;
			ms[2] += capacity;
			poured = 1;
		}
		if (myType == -1) ms[1] = hisType;
	}
	if (poured && hs[2] == 0) { /* bottleUpdate(him) */
		if (((W.hphys[him] >> 8) & 0xffu) == 9) hs[1] = -1;
		else { /* uDead entUpdaters.c:20-22 */
			int g = W.world[him];
			if (W.flip[g]) W.dead_max1[him] |= 1; else W.dead_max0[him] |= 1;
			W.wflush[g] = 1;
			W.ctr[CT_FLUSH_HINT] = 1;
		}
	}
	return 0;
}

/* push_t: 0 onPushDefault handlerRegistrar.cpp:16; 1 player_push; 2 logic_push logic.c:44-49.
 * Writing v2=1 is idempotent, so concurrent pushers of one logic block do not race. */
static BT_HD int bt_push_handler(const DevWorld &W, i32 me, i32 him) {
	int id = (W.hphys[me] >> 8) & 0xff;
	if (id == 1) return bt_player_push(W, me, him);
	if (id == 4 || id == 7) {
		/* teamselect_push teamselect.cpp:36-45 / loadblock_push loadblock.cpp:17-31: both act on
		 * gs->players, i.e. on the host; "him is some player's entity" is "him runs TICK_PLAYER" */
		if (id == 7 && bt_getslider(W, me, 1) > 0) return 0;
		if ((W.htick[him] & 0xff) != TICK_PLAYER) return 0;
		u32 slot = bt_add32(&W.ctr[CT_EVENT_N], 1u);
		if ((i32)slot < W.event_cap) W.events[slot] = make_int4(id, me, him, id == 7 ? bt_getslider(W, me, 0) : 0);
		if (id == 7 && W.slider_off[me + 1] - W.slider_off[me] > 1) W.sv[W.slider_off[me] + 1] = 15; /* direct write */
		return 0;
	}
	if (id == 5) { /* scoreboard_push scoreboard.cpp:19-36 */
		if ((W.htick[me] & 0xff) != (W.htick[him] & 0xff)) return 0;
		bt_uslider(W, me, 0, 0);
		bt_uslider(W, me, 1, 0);
		bt_uslider(W, me, 2, bt_getslider(W, him, 2));
		bt_uslider(W, me, 3, bt_getslider(W, him, 3));
		for (i32 k = 0; k < W.wire_cnt[me]; k++) bt_or32(&W.btn_v2[W.wire_to[W.wire_off[me] + k]], 1u);
		return 0;
	}
	if (id == 3) { /* seat_push seat.cpp:10-16 */
		if (W.wire_cnt[me]) return 0;
		if ((W.type[him] & T_INPUTS) && !(W.btn_v[him] & 0xffu)) bt_upickup(W, me, him, HOLD_DROP + HOLD_FREEZE + HOLD_SINGLE);
		return 0;
	}
	if (id == 6) return bt_teleport_push(W, me, him);
	if (id == 8 || id == 9) return bt_bottle_push(W, id, me, him);
	if (id == 2) {
		if (W.type[him] & (T_TERRAIN | T_OBSTACLE | T_DEBRIS)) bt_or32(&W.btn_v2[me], 1u); /* pushBtn(me,0) */
	} else if (id != 0) {
		bt_unsup(W, me, 0, id); /* only reachable after a lenient upload */
	}
	return 0;
}

/* uPickup entUpdaters.c:29-52: a request, applied by the next flushPickups */
static BT_HD void bt_upickup(const DevWorld &W, i32 p, i32 e, i32 holdFlags) {
	if (W.holder[e] >= 0) return;
	if (W.dead[e] || W.dead[p]) return;
	i32 nh = W.newHolder[e];
	if (nh >= 0) {
		if (nh != p || W.newHoldFlags[e] != holdFlags) W.newHolder[e] = e;
		return;
	}
	W.newHolder[e] = p;
	W.newHoldFlags[e] = holdFlags;
	W.wflush[W.world[e]] = 1; /* the world has ordered flush1 work (k_flush_scan only runs after an upload) */
	W.ctr[CT_FLUSH_HINT] = 1;
}

/* pushed_t: 2 fragile_pushed common.c:56; 3 flag_pushed flag.c:13; 5 player_pushed
 * player.c:88-106; 1 pj_pushed puddlejumper.cpp:44-65.  NULL slot (0) is skipped (ent.cpp:556). */
static BT_HD int bt_pushed_handler(const DevWorld &W, i32 me, i32 him, int axis, int dir, i32 dx, i32 dv) {
	int id = (W.hphys[me] >> 16) & 0xff;
	switch (id) {
	case 2: return 1;
	case 3: return (W.type[him] & T_FLAG) ? 1 : 0;
	case 8: { /* bullet_pushed gun.c:90-114: stick to whatever was hit */
		if (bt_getslider(W, me, 0) == 0) return 0;
		bt_upickup(W, him, me, HOLD_DROP | HOLD_FREEZE);
		W.sv[W.slider_off[me]] = 0; /* immediate slider write */
		if (dv) {
			i32 scale = 7000;
			i32 ratio = dx * scale / dv;
			int4 c = W.center[me], vm = W.vel[me], vh = W.vel[him];
			const i32 rel[3] = {vm.x - vh.x, vm.y - vh.y, vm.z - vh.z};
			for (int i = 0; i < 2; i++) {
				int dim = (axis + 1 + i) % 3;
				bt_set3(c, dim, bt_get3(c, dim) - rel[dim] * ratio / scale);
			}
			W.center[me] = c;
			if (W.hasBox[me]) bt_check_cover(W, me, W.old[me], c);
		}
		W.type[me] = 0; /* immediate, without touching new*Mask or the box */
		W.collide[me] = 0;
		return 0;
	}
	case 1:
	case 5: {
		int s_vel = id == 5 ? 5 : 2, s_grounded = id == 5 ? 7 : 4;
		int4 vm = W.vel[me], vh = W.vel[him];
		i32 vel[3] = {vm.x - vh.x, vm.y - vh.y, vm.z - vh.z};
		if (id == 1) bt_bound_vec(vel, 200, 2);
		if (axis == 2) {
			if (dir < 0) bt_uslider(W, me, s_grounded, 1);
		} else {
			i32 currentSlide = bt_getslider(W, me, s_vel + axis);
			vel[axis] = currentSlide + dir * dv;
			if (dir * vel[axis] > 0) vel[axis] = 0;
		}
		bt_uslider(W, me, s_vel, vel[0]);
		bt_uslider(W, me, s_vel + 1, vel[1]);
		return 0;
	}
	default:
		bt_unsup(W, me, 1, id);
		return 0;
	}
}

/* entPair_t: 3 decor_fumbled common.c:61-67.  slot: 0 onFumble 1 onFumbled 2 onPickUp 3 onPickedUp */
static BT_HD void bt_pair_handler(const DevWorld &W, int slot, i32 me, i32 him) {
	int id = (W.hpair[me] >> (8 * slot)) & 0xff;
	if (id == 5) { /* player_pickup player.c:167-178 -> updateTrajectory entUpdaters.c:102-114 */
		if (W.type[him] & EQUIP_MASK) {
			int4 mo = W.old[me], mc = W.center[me], mv = W.vel[me];
			int4 ho = W.old[him], hc = W.center[him], hv = W.vel[him];
			BT_FOR_SUBTREE(W, him, it, {
				int4 o = W.old[it], c = W.center[it], v = W.vel[it];
				o.x += mo.x - ho.x; o.y += mo.y - ho.y; o.z += mo.z - ho.z;
				c.x += mc.x - hc.x; c.y += mc.y - hc.y; c.z += mc.z - hc.z;
				v.x += mv.x - hv.x; v.y += mv.y - hv.y; v.z += mv.z - hv.z;
				W.old[it] = o; W.center[it] = c; W.vel[it] = v;
			});
			W.ctr[CT_CACHE_FULL] = 1; /* it jumped: the cached pairs no longer cover it */
			W.ctr[CT_CACHE_INVALID] = 1;
			W.holdFlags[him] |= HOLD_DROP;
		}
		return;
	}
	if (id == 1) { /* seat_pickUp seat.cpp:32-36 */
		if (!W.wire_cnt[me] && (W.type[him] & T_INPUTS)) bt_uwire(W, me, him);
		return;
	}
	if (id == 2 || id == 4) { /* seat_fumble_old / seat_fumble seat.cpp:18-30 */
		bt_uunwire(W, me, him);
		if (id == 4) {
			int ns = W.slider_off[me + 1] - W.slider_off[me];
			for (int i = 0; i < 5; i++) {
				if (i < ns) bt_uslider(W, me, i, 0);
			}
		}
		return;
	}
	if (id == 3) {
		W.fullFlush[me] = 1;
		W.ctr[CT_FLUSH_HINT] = 1;
		W.ntype[me] = W.type[me] | T_DEBRIS;
		W.ncollide[me] = W.collide[me] | T_TERRAIN | T_OBSTACLE | T_DEBRIS;
		if (slot == 1) W.hpair[me] &= ~(0xffu << 8);
	} else if (id != 0) {
		bt_unsup(W, me, 2, id);
	}
}

/* ------------------------------------------------------------------------- */
/* tree edits (only ever called by the thread that owns the tree this iteration) */

/* fumble ent.cpp:136-158; addRootEnt :60-66 has no list here (roots = holder < 0) */
static BT_HD void bt_fumble(const DevWorld &W, i32 h, LocalStats &st) {
	i32 a = W.holder[h];
	if (a < 0) return;
	st.fumble++;
	W.needsCollision[h] = 1;
	bt_pair_handler(W, 0, a, h);
	bt_pair_handler(W, 1, h, a);
	for (int i = 0; i < 3; i++) {
		i32 x = h, y;
		while ((y = W.holder[x]) >= 0 && bt_forced_get(W.forced[y], i) && bt_get3(W.forcedHoldees[y], i) == x) {
			W.forced[y] = bt_forced_set(W.forced[y], i, 0);
			x = y;
		}
	}
	i32 n = W.LLn[h], p = W.LLp[h];
	if (n >= 0) W.LLp[n] = p;
	if (p >= 0) W.LLn[p] = n;
	else W.holdee[a] = n;
	W.holder[h] = -1;
	W.LLn[h] = -1;
	W.LLp[h] = -1;
	BT_FOR_SUBTREE(W, h, it, W.holdRoot[it] = h); /* recursiveHoldRoot ent.cpp:39-45 */
	/* addRootEnt ent.cpp:60-66 prepends to gs->rootEnts: the newest root gets the largest key.  Inside one
	 * launch: the ordered lane (flush1) counts per world; pushes order by canonical push order (descending
	 * root), then by the fumbles of that one push. */
	if (st.rk_ordered) {
		W.rootKey[h] = ((u64)(u32)W.root_phase << BT_ROOTKEY_SHIFT) | (u64)bt_add32(&W.wrootseq[W.world[h]], 1u);
	} else {
		if (st.rk_local > 255u) bt_unsup(W, h, 8, 0); /* > 256 fumbles by one push: their mutual order is lost */
		W.rootKey[h] = st.rk_base | (u64)(st.rk_local > 255u ? 255u : st.rk_local);
		st.rk_local++;
	}
	bt_add32(&W.ctr[CT_TREE_EDITS], 1u);
}

/* killEntNoHandlers ent.cpp:190-208 + crushEnt :211-217.  The dead list replaces
 * gs->deadTail; `key` orders it canonically (iteration, pusher) for the host. */
static BT_HD void bt_crush(const DevWorld &W, i32 e, u32 key, LocalStats &st, int d = 1) {
	if (W.dead[e]) return;
	st.crush++;
	while (W.holdee[e] >= 0) bt_fumble(W, W.holdee[e], st);
	bt_fumble(W, e, st);
	W.inList[e] = 0;
	W.hasBox[e] = 0;
	W.dead[e] = 1;
	bt_add32(&W.ctr[CT_TREE_EDITS], 1u);
	W.ctr[CT_FLUSH_HINT] = 1; /* wires that point at it must go (flushMisc ent.cpp:706-712) */
	int cid = (int)((W.hphys[e] >> 24) & 0xff);
	if (cid > 2) bt_unsup(W, e, 3, cid);
	else if (cid && d == 1 && !W.flush_stage) {
		/* invokeOnCrush ent.cpp:464-469 runs flag_crush / platform crush on the host (they create
		 * entities): one event per death that the reference would call the handler for */
		u32 ev = bt_add32(&W.ctr[CT_EVENT_N], 1u);
		if ((i32)ev < W.event_cap) W.events[ev] = make_int4(0x300 | cid, e, e, (i32)key);
	}
	u32 slot = bt_add32(&W.ctr[CT_DEAD_N], 1u);
	if ((i32)slot < W.dead_cap) W.deadList[slot] = ((u64)key << 32) | (u32)e;
}

static BT_HD void bt_swept_box(const int4 &o, const int4 &c, const int4 &r, i32 lo[3], i32 hi[3]) {
	lo[0] = (o.x < c.x ? o.x : c.x) - r.x; hi[0] = (o.x > c.x ? o.x : c.x) + r.x;
	lo[1] = (o.y < c.y ? o.y : c.y) - r.y; hi[1] = (o.y > c.y ? o.y : c.y) + r.y;
	lo[2] = (o.z < c.z ? o.z : c.z) - r.z; hi[2] = (o.z > c.z ? o.z : c.z) + r.z;
}

/* does the entity still lie inside the inflated box its cached pairs were found with? */
static BT_HD void bt_check_cover(const DevWorld &W, i32 e, const int4 &o, const int4 &c) {
	int4 r = W.radius[e];
	Box6 ib = W.ibox[e];
	i32 lo[3], hi[3];
	bt_swept_box(o, c, r, lo, hi);
	if (lo[0] < ib.lo.x || lo[1] < ib.lo.y || lo[2] < ib.lo.z || hi[0] > ib.hi.x || hi[1] > ib.hi.y || hi[2] > ib.hi.z) {
		if (W.small_mode) { W.ctr[CT_CACHE_INVALID] = 1; return; } /* brute-force pair search: always a full build, no dirty list */
		/* once per entity and epoch onto the list the next patch works from */
		u32 stamp = W.ctr[CT_PATCH_EPOCH] + 1u;
		if (bt_xchg32(&W.dirtyStamp[e], stamp) != stamp) {
#ifdef BT_EMU
			W.dirtyList[bt_add32(&W.ctr[CT_DIRTY_N], 1u)] = e;
#else
			/* one atomic per warp for the lanes that are here together (a kicked world appends a quarter of its
			 * entities in one integrate kernel, all to the same counter) */
			const u32 m = __activemask();
			const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
			u32 base = 0;
			if (lane == leader) base = atomicAdd(&W.ctr[CT_DIRTY_N], (u32)__popc(m));
			base = __shfl_sync(m, base, leader);
			W.dirtyList[base + __popc(m & ((1u << lane) - 1))] = e;
#endif
		}
		W.ctr[CT_CACHE_INVALID] = 1;
	}
}

/* moveRecursive ent.cpp:115-126 */
static BT_HD void bt_move_recursive(const DevWorld &W, i32 who, const i32 *dc, const i32 *dv) {
	BT_FOR_SUBTREE(W, who, it, {
		int4 c = W.center[it], v = W.vel[it];
		c.x += dc[0]; c.y += dc[1]; c.z += dc[2];
		v.x += dv[0]; v.y += dv[1]; v.z += dv[2];
		W.center[it] = c;
		W.vel[it] = v;
		if (W.hasBox[it]) bt_check_cover(W, it, W.old[it], c);
	});
}

/* ------------------------------------------------------------------------- */
/* narrow phase                                                              */

struct EntGeo {
	int4 old, center, radius;
};

/* getAxisAndDir ent.cpp:238-313.  Returns +-(axis+1) or 0. */
static BT_HD int bt_axis_and_dir(const EntGeo &a, const EntGeo &b) {
	int d1[3], d2[3], wd[3];
	const int ao[3] = {a.old.x, a.old.y, a.old.z}, bo[3] = {b.old.x, b.old.y, b.old.z};
	const int ac[3] = {a.center.x, a.center.y, a.center.z}, bc[3] = {b.center.x, b.center.y, b.center.z};
	const int ar[3] = {a.radius.x, a.radius.y, a.radius.z}, br[3] = {b.radius.x, b.radius.y, b.radius.z};
#pragma unroll
	for (int i = 0; i < 3; i++) {
		d1[i] = bo[i] - ao[i];
		d2[i] = bc[i] - ac[i];
		wd[i] = ar[i] + br[i];
		if (d1[i] > 0) {
			if (d1[i] >= wd[i] && d2[i] >= wd[i]) return 0;
		} else {
			if (d1[i] <= -wd[i] && d2[i] <= -wd[i]) return 0;
		}
		d2[i] = d1[i] - d2[i];
	}
	int dir_by_axis[3];
#pragma unroll
	for (int i = 0; i < 3; i++) {
		if (d2[i] >= 0) {
			dir_by_axis[i] = -(i + 1);
		} else {
			dir_by_axis[i] = (i + 1);
			d1[i] = -d1[i];
			d2[i] = -d2[i];
		}
	}
#pragma unroll
	for (int i = 0; i < 3; i++) d1[i] -= wd[i];
	int best;
	{
		long long cmp = ((long long)d1[0]) * d2[1] - ((long long)d1[1]) * d2[0];
		if (cmp > 0) best = 0;
		else if (cmp < 0) best = 1;
		else best = (wd[0] > wd[1]) ? 1 : 0;
	}
	{
		int A = best;
		long long cmp = ((long long)d1[A]) * d2[2] - ((long long)d1[2]) * d2[A];
		if (cmp > 0) best = A;
		else if (cmp < 0) best = 2;
		else best = (wd[A] > wd[2]) ? 2 : A;
	}
	if (d1[best] < 0) {
		int depth = INT32_MAX;
		int ret = 0;
#pragma unroll
		for (int i = 0; i < 3; i++) {
			int d = d1[i] + wd[i];
			int new_depth, new_ret;
			if (d >= 0) {
				new_depth = wd[i] - d;
				new_ret = dir_by_axis[i];
			} else {
				new_depth = wd[i] + d;
				new_ret = d2[i] ? 0 : -dir_by_axis[i];
			}
			if (new_depth < depth) {
				depth = new_depth;
				ret = new_ret;
			}
		}
		return ret;
	}
#pragma unroll
	for (int i = 0; i < 3; i++) {
		int far_ = d1[i] + 2 * wd[i];
		long long cmp = ((long long)d1[best]) * d2[i] - ((long long)far_) * d2[best];
		if (cmp >= 0) return 0;
	}
	return dir_by_axis[best];
}

/* the three geometry arrays, passed by value into the non-inlined proposal code so that
 * the whole DevWorld does not have to be spilled to the stack for the call */
struct GeoView {
	const int4 *old, *center, *radius;
	u64 *pend;
	const i32 *holdRoot;
	const i32 *world;
	u32 *wties, *ctr;
};
static BT_HD GeoView bt_view(const DevWorld &W) {
	GeoView v;
	v.old = W.old; v.center = W.center; v.radius = W.radius; v.pend = W.pend; v.holdRoot = W.holdRoot;
	v.world = W.world; v.wties = W.wties; v.ctr = W.ctr;
	return v;
}
static BT_HD EntGeo bt_geo(const GeoView &V, i32 e) {
	EntGeo g;
	g.old = V.old[e];
	g.center = V.center[e];
	g.radius = V.radius[e];
	return g;
}
static BT_HD EntGeo bt_geo(const DevWorld &W, i32 e) {
	EntGeo g;
	g.old = W.old[e];
	g.center = W.center[e];
	g.radius = W.radius[e];
	return g;
}

/* collisionBetter ent.cpp:317-382 as a three-way compare of two explicit records:
 * +1 c0 strictly better, -1 strictly worse, 0 full tie.  32-bit products wrap
 * exactly like the reference's `delta * speed` (ent.cpp:343).
 * Every term of the comparison belongs to ONE of the two collisions, so each is first boiled down to a key of
 * four numbers from its own leaf and buddy (two entities in registers at a time, not four):
 *   delta  distance still to go along the axis at `old`          (ent.cpp:338-339)
 *   speed  closing speed along the axis, clamped to >= 1          (:340-342); the OTHER record's speed scales delta
 *   lat    smaller of the two lateral overlaps at `old`           (:352-368)
 *   depth  displacement the push would need at `center`           (:370-378) */
struct CKey {
	i32 delta, speed, lat, depth;
	int mutual;
};
static BT_HD CKey bt_ckey(const EntGeo &l, const EntGeo &f, int a, int dir, int mutual) {
	CKey k;
	k.delta = (i32)((u32)(bt_get3(l.old, a) - bt_get3(f.old, a)) * (u32)dir) - bt_get3(f.radius, a) - bt_get3(l.radius, a);
	i32 sp = (i32)((u32)(bt_get3(f.center, a) - bt_get3(f.old, a) + bt_get3(l.old, a) - bt_get3(l.center, a)) * (u32)dir);
	k.speed = sp < 1 ? 1 : sp;
	int x1 = (a + 1) % 3, x2 = (a + 2) % 3;
	i32 p = bt_get3(f.radius, x1) + bt_get3(l.radius, x1) - bt_abs(bt_get3(f.old, x1) - bt_get3(l.old, x1));
	i32 q = bt_get3(f.radius, x2) + bt_get3(l.radius, x2) - bt_abs(bt_get3(f.old, x2) - bt_get3(l.old, x2));
	k.lat = p < q ? p : q;
	k.depth = bt_get3(f.radius, a) + bt_get3(l.radius, a) + (bt_get3(f.center, a) - bt_get3(l.center, a)) * dir;
	k.mutual = mutual;
	return k;
}
static BT_HD int bt_ckey_cmp(const CKey &k0, const CKey &k1) {
	long long v0 = (i32)((u32)k0.delta * (u32)k1.speed);
	long long v1 = (i32)((u32)k1.delta * (u32)k0.speed);
	if (v0 != v1) return v0 < v1 ? 1 : -1;
	int m0 = !k0.mutual, m1 = !k1.mutual;
	if (m0 != m1) return m0 > m1 ? 1 : -1;
	if (k0.lat != k1.lat) return k0.lat > k1.lat ? 1 : -1;
	if (k0.depth != k1.depth) return k0.depth > k1.depth ? 1 : -1;
	return 0;
}
static BT_HD CKey bt_ckey_of(const GeoView &W, const Coll &c) {
	const EntGeo l = bt_geo(W, c.leaf), f = bt_geo(W, c.buddy);
	return bt_ckey(l, f, c.axis, c.dir, c.mutual);
}
static BT_HD int bt_collision_cmp(const GeoView &W, const Coll &c0, const Coll &c1) {
	const CKey k0 = bt_ckey_of(W, c0);
	const CKey k1 = bt_ckey_of(W, c1);
	return bt_ckey_cmp(k0, k1);
}

/* tryCollide ent.cpp:384-396 as an order-independent atomic maximum on pend[root].  gl / gf = the proposal's own
 * leaf and buddy, which the caller holds in registers: its key costs no load, only an incumbent's does. */
static BT_HD void bt_propose_g(const GeoView &W, i32 leaf, i32 buddy, int axis, int dir, int mutual, const EntGeo &gl, const EntGeo &gf) {
	i32 root = W.holdRoot[leaf];
	Coll c;
	c.leaf = leaf; c.buddy = buddy; c.axis = axis; c.dir = dir; c.mutual = mutual;
	u64 mine = bt_pack_coll(c);
	/* a root usually gets ONE collision per iteration (seen from both ends of its pair at most): try the empty slot
	 * straight away -- one round trip to L2 instead of a load and then the CAS */
	u64 cur = bt_cas64(&W.pend[root], 0ull, mine);
	if (cur == 0 || cur == mine) return;
	const CKey mk = bt_ckey(gl, gf, axis, dir, mutual);
	while (1) {
		if (cur == mine) return; /* the same collision seen from the other end of the pair */
		if (cur != 0) {
			int r = bt_ckey_cmp(mk, bt_ckey_of(W, bt_unpack_coll(cur)));
			if (r < 0) return;
			if (r == 0) {
				/* A full tie between two different collisions: the reference keeps whichever its list
				 * order met first (tryCollide only replaces on "better", ent.cpp:367-396), here the
				 * larger key wins.  Almost always the same outcome; counted, because when it is
				 * not (two equal supporters, one of which is hit in the same tick) the reference's
				 * answer depends on its visiting order. */
				if ((mine >> 2) != (cur >> 2)) { bt_add32(&W.wties[W.world[root]], 1u); bt_add32(&W.ctr[CT_ORDER_TIES], 1u); }
				if ((mine >> 2) <= (cur >> 2)) return;
			}
		}
		u64 prev = bt_cas64(&W.pend[root], cur, mine);
		if (prev == cur) return;
		cur = prev;
	}
}
static BT_HD void bt_propose_v(const GeoView &W, i32 leaf, i32 buddy, int axis, int dir, int mutual) {
	bt_propose_g(W, leaf, buddy, axis, dir, mutual, bt_geo(W, leaf), bt_geo(W, buddy));
}

static BT_HD void bt_propose(const DevWorld &W, i32 leaf, i32 buddy, int axis, int dir, int mutual) {
	bt_propose_v(bt_view(W), leaf, buddy, axis, dir, mutual);
}

/* checkCollision ent.cpp:398-435 for scanning entity a against b, both already in registers.
 * bScans = b is scanned itself in this doIteration (coordinator rule, DESIGN.md "Canonical order") */
static BT_HD void bt_check_loaded(const DevWorld &W, i32 a, const EntGeo &ga, u32 ta, u32 ca, i32 rootA, u32 ha, i32 b,
                                  const EntGeo &gb, u32 tb, u32 cb, i32 rootB, u32 hbp, int bScans, LocalStats &st,
                                  u32 &any) {
	int AonB = (tb & ca) != 0, BonA = (ta & cb) != 0;
	if (!AonB && !BonA) return;
	if (rootA == rootB) return;
	int dir = bt_axis_and_dir(ga, gb);
	if (dir == 0) return;
	st.pos++;
	int axis;
	if (dir < 0) { axis = -dir - 1; dir = -1; }
	else { axis = dir - 1; dir = 1; }
	int hb = (int)(hbp & 0xff);
	int movee;
	if (AonB) {
		if (BonA) { /* arbitrateMovement ent.cpp:229-236 */
			int A = bt_who_moves((int)(ha & 0xff), ta, tb, axis, dir);
			int B = bt_who_moves(hb, tb, ta, axis, -dir);
			if (B == MOVE_ME) B = MOVE_HIM;
			else if (B == MOVE_HIM) B = MOVE_ME;
			A = A & B;
			movee = A ? A : MOVE_BOTH;
		} else {
			movee = bt_who_moves((int)(ha & 0xff), ta, tb, axis, dir);
		}
	} else {
		movee = 3 ^ bt_who_moves(hb, tb, ta, axis, -dir);
	}
	if (movee & 8) return;
	any = 1;
	/* one inlined copy of the proposal code, run once or (MOVE_BOTH) twice */
	int both = (movee & 4) != 0;
	/* a scans; it coordinates unless b scans too and comes first in the reference's newest-first order */
	int coordA = bScans ? (a > b) : 1;
	int first_is_a = both || movee == MOVE_ME;
	GeoView V = bt_view(W);
#pragma unroll 1
	for (int k = 0; k < 1 + both; k++) {
		int forA = k == 0 ? first_is_a : 0;
		int mutual = both ? ((forA ? coordA : !coordA) ? 1 : 2) : 0;
		if (forA) bt_propose_g(V, a, b, axis, dir, mutual, ga, gb);
		else bt_propose_g(V, b, a, axis, -dir, mutual, gb, ga);
	}
}

static BT_HD void bt_check_collision(const DevWorld &W, i32 a, const EntGeo &ga, u32 ta, u32 ca, i32 rootA, u32 ha,
                                     i32 b, LocalStats &st, u32 &any) {
	u32 tb = W.type[b], cb = W.collide[b];
	if (!((tb & ca) | (ta & cb))) return;
	int bScans = W.needsCollision[b] && W.hasBox[b] && W.inList[b];
	bt_check_loaded(W, a, ga, ta, ca, rootA, ha, b, bt_geo(W, b), tb, cb, W.holdRoot[b], W.hphys[b], bScans, st, any);
}

/* ------------------------------------------------------------------------- */
/* broadphase: Morton keys, Karras LBVH, refit, traversal                    */

static BT_HD int bt_is_leaf_ent(const DevWorld &W, i32 e) { return W.inList[e] && (W.ray_mode || W.hasBox[e]); }

/* min/max of leaf centres -> ctr[CT_BOUNDS..+5] */
static BT_HD void t_bounds(const DevWorld &W, i32 e) {
	if (e >= W.n || !bt_is_leaf_ent(W, e)) return;
	int4 c = W.center[e];
	i32 *b = (i32 *)&W.ctr[CT_BOUNDS];
	bt_min32(&b[0], c.x); bt_min32(&b[1], c.y); bt_min32(&b[2], c.z);
	bt_max32(&b[3], c.x); bt_max32(&b[4], c.y); bt_max32(&b[5], c.z);
}

static BT_HD u32 bt_expand10(u32 v) { /* 10 bits -> every third bit */
	v &= 0x3ff;
	v = (v | (v << 16)) & 0x030000FF;
	v = (v | (v << 8)) & 0x0300F00F;
	v = (v | (v << 4)) & 0x030C30C3;
	v = (v | (v << 2)) & 0x09249249;
	return v;
}

static BT_HD int bt_bits_for(u32 x) { return 32 - bt_clz32(x); }

/* key = world id << 30 | size class << 27 | 27-bit Morton code of the centre (9 bits/axis).
 * The size class (x4 radius steps, like velbox's SCALE=4 levels, velbox/box.h:24-30) sits
 * ABOVE the Morton bits, so inside a world the Karras hierarchy first splits by class: huge
 * boxes ("whales": ground tiles, walls) get their own tight subtree instead of bloating
 * every ancestor of the small boxes around them, and because larger classes sort later, the
 * "earlier leaf emits the pair" rule of t_pairs makes the many small boxes look up the few
 * large ones and never the other way round.  Non-leaves get world id n_worlds and sort last
 * (radix passes cover 30 + bits(n_worlds) bits). */
static BT_HD void t_morton(const DevWorld &W, i32 e) {
	if (e >= W.n) return;
	u64 key = (u64)(u32)W.n_worlds << 30;
	if (bt_is_leaf_ent(W, e)) {
		int4 c = W.center[e], r = W.radius[e];
		const i32 *b = (const i32 *)&W.ctr[CT_BOUNDS];
		u32 qx = ((u32)(c.x - b[0])) >> W.key_shift[0];
		u32 qy = ((u32)(c.y - b[1])) >> W.key_shift[1];
		u32 qz = ((u32)(c.z - b[2])) >> W.key_shift[2];
		u32 m = (bt_expand10(qx) << 2) | (bt_expand10(qy) << 1) | bt_expand10(qz);
		i32 rmax = r.x > r.y ? r.x : r.y;
		if (r.z > rmax) rmax = r.z;
		int cls = (bt_bits_for((u32)(rmax < 0 ? 0 : rmax)) - 7) >> 1;
		cls = cls < 0 ? 0 : (cls > 7 ? 7 : cls);
		key = ((u64)(u32)W.world[e] << 30) | ((u64)cls << 27) | (m & 0x7ffffffu);
	}
	W.keys[e] = key;
	W.leafEnt[e] = e;
}

static BT_HD int bt_delta(const DevWorld &W, int i, int j, int L) {
	if (j < 0 || j >= L) return -1;
	u64 a = W.keys[i], b = W.keys[j];
	if (a == b) return 64 + bt_clz32((u32)i ^ (u32)j);
	return bt_clz64(a ^ b);
}

/* Karras 2012 "Maximizing parallelism in the construction of BVHs": internal node i */
static BT_HD void t_lbvh_node(const DevWorld &W, int i) {
	int L = W.n_leaves;
	if (i >= L - 1) return;
	int d = (bt_delta(W, i, i + 1, L) - bt_delta(W, i, i - 1, L)) >= 0 ? 1 : -1;
	int dmin = bt_delta(W, i, i - d, L);
	int lmax = 2;
	while (bt_delta(W, i, i + lmax * d, L) > dmin) lmax *= 2;
	int l = 0;
	for (int t = lmax / 2; t >= 1; t /= 2)
		if (bt_delta(W, i, i + (l + t) * d, L) > dmin) l += t;
	int j = i + l * d;
	int dnode = bt_delta(W, i, j, L);
	int s = 0;
	int t = l;
	do {
		t = (t + 1) / 2;
		if (bt_delta(W, i, i + (s + t) * d, L) > dnode) s += t;
	} while (t > 1);
	int gamma = i + s * d + (d < 0 ? -1 : 0);
	int lo = i < j ? i : j, hi = i < j ? j : i;
	BvhNode &N = W.nodes[i];
	if (lo == gamma) { N.left = ~gamma; W.leafParent[gamma] = i; }
	else { N.left = gamma; W.nodeParent[gamma] = i; }
	if (hi == gamma + 1) { N.right = ~(gamma + 1); W.leafParent[gamma + 1] = i; }
	else { N.right = gamma + 1; W.nodeParent[gamma + 1] = i; }
	W.nodeHi[i] = hi;
	if (i == 0) W.nodeParent[0] = -1;
	/* World roots.  Keys carry the world id in their top bits, so every world is one subtree;
	 * the nodes above (they span several worlds, whose boxes all overlap) would only cost every
	 * query and every refit ~log2(worlds) extra levels.  A child of a multi-world node that is
	 * all one world IS that world (a Karras node holds every key with its prefix): cut it loose. */
	u32 wlo = (u32)(W.keys[lo] >> 30), whi = (u32)(W.keys[hi] >> 30);
	if (wlo != whi) {
		if ((u32)(W.keys[gamma] >> 30) == wlo) {
			if (lo == gamma) { W.leafParent[gamma] = -1; W.wroot[wlo] = ~gamma; }
			else { W.nodeParent[gamma] = -1; W.wroot[wlo] = gamma; }
		}
		if ((u32)(W.keys[gamma + 1] >> 30) == whi) {
			if (hi == gamma + 1) { W.leafParent[gamma + 1] = -1; W.wroot[whi] = ~(gamma + 1); }
			else { W.nodeParent[gamma + 1] = -1; W.wroot[whi] = gamma + 1; }
		}
	} else if (i == 0) {
		W.wroot[wlo] = 0;
	}
}

static BT_HD int bt_min(int a, int b) { return a < b ? a : b; }
static BT_HD int bt_max(int a, int b) { return a > b ? a : b; }

/* refit: leaf position t computes its swept box (what velbox_step / velbox_update keep in
 * p1,p2,r: box.cpp:442-459, ent.cpp:892-904) and climbs; the second thread to reach a node
 * (atomic counter) builds it from both children. */
/* Patch or full build?  Uniform over a launch and the same for k_refit, k_pairs and the
 * k_narrow that consumes the flags: none of the three writes a word this reads before the
 * last of them has read it. */
static BT_HD int bt_patch_mode(const DevWorld &W) {
	return W.patch && !W.small_mode && !W.ray_mode && !W.ctr[CT_CACHE_FULL] && ((u64)W.ctr[CT_DIRTY_N] << W.patch_shift) <= (u64)(u32)W.n_leaves;
}

static BT_HD Box6 bt_inflated_box(const DevWorld &W, i32 e) {
	i32 lo[3], hi[3];
	int4 cc = W.center[e];
	bt_swept_box(W.ray_mode ? cc : W.old[e], cc, W.radius[e], lo, hi);
	int m = W.ray_mode ? 0 : W.margin;
	if (!W.ray_mode && W.horizon > 0) {
		/* where it will be for the next `horizon` ticks if nothing hits it; gravity adds 32 per
		 * tick to +z (effects.cpp:36).  A heuristic only: bt_check_cover stays exact. */
		int4 v = W.vel[e];
		i32 H = W.horizon;
		long long ex = (long long)v.x * H, ey = (long long)v.y * H, ez = (long long)v.z * H + 16 * H * (H + 1);
		const long long cap = 1 << 20; /* no stretching by absurd speeds */
		ex = ex < -cap ? -cap : (ex > cap ? cap : ex); ey = ey < -cap ? -cap : (ey > cap ? cap : ey);
		ez = ez < -cap ? -cap : (ez > cap ? cap : ez);
		if (ex < 0) lo[0] += (i32)ex; else hi[0] += (i32)ex;
		if (ey < 0) lo[1] += (i32)ey; else hi[1] += (i32)ey;
		if (ez < 0) lo[2] += (i32)ez; else hi[2] += (i32)ez;
	}
	Box6 box;
	box.lo = make_int4(lo[0] - m, lo[1] - m, lo[2] - m, W.world[e]);
	box.hi = make_int4(hi[0] + m, hi[1] + m, hi[2] + m, W.world[e]);
	return box;
}

/* grow p[k] to include [lo,hi]; returns whether anything changed */
static BT_HD int bt_grow3(i32 *plo, i32 *phi, const int4 &lo, const int4 &hi) {
	int ch = 0;
	if (lo.x < bt_ldv32(&plo[0])) { bt_min32(&plo[0], lo.x); ch = 1; }
	if (lo.y < bt_ldv32(&plo[1])) { bt_min32(&plo[1], lo.y); ch = 1; }
	if (lo.z < bt_ldv32(&plo[2])) { bt_min32(&plo[2], lo.z); ch = 1; }
	if (hi.x > bt_ldv32(&phi[0])) { bt_max32(&phi[0], hi.x); ch = 1; }
	if (hi.y > bt_ldv32(&phi[1])) { bt_max32(&phi[1], hi.y); ch = 1; }
	if (hi.z > bt_ldv32(&phi[2])) { bt_max32(&phi[2], hi.z); ch = 1; }
	return ch;
}

/* Incremental refit, thread t = t-th entry of dirtyList: what velbox_update does for one box
 * that outran its validity (velbox/box.cpp:442-459, ent.cpp:892-904).  The entity gets a
 * fresh inflated box; its parent takes it as is, and every higher ancestor only GROWS to
 * contain it (atomic min/max), so the tree stays a superset until the next full refit. */
static BT_HD void t_refit_patch(const DevWorld &W, int t) {
	if ((u32)t >= W.ctr[CT_DIRTY_N]) return;
	if (t == 0) bt_add32(&W.ctr[CT_PAIR_PATCHES], 1u);
	i32 e = W.dirtyList[t];
	W.iboxOld[e] = W.ibox[e];
	if (!bt_is_leaf_ent(W, e)) return; /* died meanwhile: its cached pairs are filtered by the narrow phase */
	Box6 box = bt_inflated_box(W, e);
	W.ibox[e] = box;
	int pos = W.leafOf[e];
	W.leafBox[pos] = box;
	if (W.n_leaves < 2) return;
	int node = W.leafParent[pos];
	{
		BvhNode &N = W.nodes[node];
		if (N.left == ~pos) {
			N.llo[0] = box.lo.x; N.llo[1] = box.lo.y; N.llo[2] = box.lo.z;
			N.lhi[0] = box.hi.x; N.lhi[1] = box.hi.y; N.lhi[2] = box.hi.z;
		} else {
			N.rlo[0] = box.lo.x; N.rlo[1] = box.lo.y; N.rlo[2] = box.lo.z;
			N.rhi[0] = box.hi.x; N.rhi[1] = box.hi.y; N.rhi[2] = box.hi.z;
		}
	}
	int child = node;
	node = W.nodeParent[node];
	while (node >= 0) {
		BvhNode &N = W.nodes[node];
		int ch = N.left == child ? bt_grow3(N.llo, N.lhi, box.lo, box.hi) : bt_grow3(N.rlo, N.rhi, box.lo, box.hi);
		/* already inside: every higher ancestor contains this node's box, or will by the end of
		 * the launch (whoever grew it is still climbing with a box whose union covers ours) */
		if (!ch) break;
		child = node;
		node = W.nodeParent[node];
	}
}

/* Bookkeeping of the non-speculative k_narrow (block 0, thread 0) once k_refit / k_pairs have
 * consumed the invalidation: np = CT_PAIR_N as this narrow phase sees it. */
static BT_HD void bt_consume_cache_flags(const DevWorld &W, u32 np) {
	if (!W.ctr[CT_CACHE_INVALID]) return;
	if (!bt_patch_mode(W)) {
		W.ctr[CT_PAIR_N_FULL] = np;
		W.ctr[CT_CACHE_FULL] = 0;
	} else {
		/* patches only append: once the list has grown by half, the next build is a full one */
		u32 f = W.ctr[CT_PAIR_N_FULL];
		if (np > f + f / 2 + 4096u) W.ctr[CT_CACHE_FULL] = 1;
	}
	W.ctr[CT_DIRTY_N] = 0;
	W.ctr[CT_PATCH_EPOCH]++;
	W.ctr[CT_CACHE_INVALID] = 0;
}

static BT_HD void t_refit(const DevWorld &W, int t, int cond) {
	int L = W.n_leaves;
	if (t >= L) return;
	if (cond && !W.ctr[CT_CACHE_INVALID]) return; /* the cached pairs still cover everybody: nothing to do */
	if (cond && bt_patch_mode(W)) { t_refit_patch(W, t); return; }
	if (cond && t == 0) { W.ctr[CT_PAIR_N] = 0; bt_add32(&W.ctr[CT_PAIR_BUILDS], 1u); }
	i32 e = W.leafEnt[t];
	Box6 box;
	if (bt_is_leaf_ent(W, e)) {
		box = bt_inflated_box(W, e);
		if (!W.ray_mode) { W.ibox[e] = box; W.leafOf[e] = t; }
	} else { /* died since the tree was built: empty box, never overlaps */
		box.lo = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, 0xffff);
		box.hi = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, 0);
	}
	W.leafBox[t] = box;
	if (L < 2) return;
	int node = W.leafParent[t];
	while (node >= 0) {
		bt_fence();
		/* two arrivals per node and refit, so the counter's parity tells first from second
		 * and never needs resetting */
		u32 prev = bt_add32(&W.nodeCount[node], 1u);
		if (!(prev & 1u)) return; /* first arrival: the sibling subtree is not done yet */
		bt_fence();
		BvhNode &N = W.nodes[node];
		i32 li = N.left, ri = N.right;
		const Box6 *pa = li < 0 ? &W.leafBox[~li] : &W.nodeBox[li];
		const Box6 *pb = ri < 0 ? &W.leafBox[~ri] : &W.nodeBox[ri];
		int4 alo = bt_ldcg4(&pa->lo), ahi = bt_ldcg4(&pa->hi), blo = bt_ldcg4(&pb->lo), bhi = bt_ldcg4(&pb->hi);
		N.llo[0] = alo.x; N.llo[1] = alo.y; N.llo[2] = alo.z;
		N.lhi[0] = ahi.x; N.lhi[1] = ahi.y; N.lhi[2] = ahi.z;
		N.rlo[0] = blo.x; N.rlo[1] = blo.y; N.rlo[2] = blo.z;
		N.rhi[0] = bhi.x; N.rhi[1] = bhi.y; N.rhi[2] = bhi.z;
		N.wl = (u32)alo.w | ((u32)ahi.w << 16);
		N.wr = (u32)blo.w | ((u32)bhi.w << 16);
		Box6 u;
		u.lo = make_int4(bt_min(alo.x, blo.x), bt_min(alo.y, blo.y), bt_min(alo.z, blo.z), bt_min(alo.w, blo.w));
		u.hi = make_int4(bt_max(ahi.x, bhi.x), bt_max(ahi.y, bhi.y), bt_max(ahi.z, bhi.z), bt_max(ahi.w, bhi.w));
		W.nodeBox[node] = u;
		node = W.nodeParent[node];
	}
}

/* One LBVH node = exactly four 128-bit loads, issued together (no short-circuit chains of
 * dependent scalar loads), then branch-free box tests on registers. */
struct NodeRegs {
	int4 a, b, c, d; /* a={llo.xyz,lhi.x} b={lhi.yz,rlo.xy} c={rlo.z,rhi.xyz} d={left,right,wl,wr} */
};
static BT_HD NodeRegs bt_load_node(const DevWorld &W, int node) {
	const int4 *p = (const int4 *)&W.nodes[node];
	NodeRegs r;
	r.a = p[0]; r.b = p[1]; r.c = p[2]; r.d = p[3];
	return r;
}
static BT_HD int bt_touch_left(const NodeRegs &n, const i32 *qlo, const i32 *qhi, u32 w) {
	u32 wl = (u32)n.d.z;
	return (qlo[0] <= n.a.w) & (n.a.x <= qhi[0]) & (qlo[1] <= n.b.x) & (n.a.y <= qhi[1]) & (qlo[2] <= n.b.y) &
	       (n.a.z <= qhi[2]) & ((wl & 0xffff) <= w) & (w <= (wl >> 16));
}
static BT_HD int bt_touch_right(const NodeRegs &n, const i32 *qlo, const i32 *qhi, u32 w) {
	u32 wr = (u32)n.d.w;
	return (qlo[0] <= n.c.y) & (n.b.z <= qhi[0]) & (qlo[1] <= n.c.z) & (n.b.w <= qhi[1]) & (qlo[2] <= n.c.w) &
	       (n.c.x <= qhi[2]) & ((wr & 0xffff) <= w) & (w <= (wr >> 16));
}

#ifndef BT_STACK
#define BT_STACK 32 /* deferred right children; observed <= 6 on the benchmark worlds; overflow raises BT_ERR_OVERFLOW */
#endif

/* doIteration ent.cpp:437-462 for the entity at sorted leaf position t.
 * `stack` is this thread's traversal stack (shared memory on the device). */
static BT_HD void t_collide(const DevWorld &W, int t, i32 *stack, int stride, LocalStats &st, u32 &any) {
	int L = W.n_leaves;
	if (t >= L) return;
	i32 a = W.leafEnt[t];
	if (!(W.needsCollision[a] && bt_is_leaf_ent(W, a))) return;
	if (L < 2) return;
	EntGeo ga = bt_geo(W, a);
	i32 qlo[3], qhi[3];
	bt_swept_box(ga.old, ga.center, ga.radius, qlo, qhi);
	u32 ta = W.type[a], ca = W.collide[a], ha = W.hphys[a], wa = (u32)W.world[a];
	i32 rootA = W.holdRoot[a];
	int sp = 0;
	int node = W.wroot[wa]; /* this world's subtree */
	if (node < 0) return;   /* it is the world's only leaf */
	while (1) {
		const NodeRegs N = bt_load_node(W, node);
		i32 li = N.d.x, ri = N.d.y;
		/* touching counts as a candidate, the narrow phase decides */
		int ovL = bt_touch_left(N, qlo, qhi, wa);
		int ovR = bt_touch_right(N, qlo, qhi, wa);
		if (ovL && li < 0) {
			i32 b = W.leafEnt[~li];
			if (b != a) { st.cand++; bt_check_collision(W, a, ga, ta, ca, rootA, ha, b, st, any); }
		}
		if (ovR && ri < 0) {
			i32 b = W.leafEnt[~ri];
			if (b != a) { st.cand++; bt_check_collision(W, a, ga, ta, ca, rootA, ha, b, st, any); }
		}
		int goL = ovL && li >= 0, goR = ovR && ri >= 0;
		if (goL) {
			node = li;
			if (goR) {
				if (sp < BT_STACK) stack[(sp++) * stride] = ri;
				else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
			}
		} else if (goR) {
			node = ri;
		} else {
			if (sp == 0) break;
			node = stack[(--sp) * stride];
		}
	}
}

/* ---- cached candidate pairs --------------------------------------------------------
 * k_pairs walks the tree once per tick (again only if an entity leaves its inflated box)
 * and writes every unordered pair of leaves whose INFLATED boxes touch, once, from the
 * leaf with the earlier position: subtrees that end at or before the query's own leaf
 * position are skipped, which halves the traversal.  k_narrow then runs the narrow phase
 * for a pair from both ends, like the reference does when both ends are scanned
 * (ent.cpp:446-449).  This plays the role of velbox's cached `intersects` lists
 * (velbox/box.cpp:95-136) within a tick. */
static BT_HD int bt_box_touch(const Box6 &a, const Box6 &b) {
	return (a.lo.x <= b.hi.x) & (b.lo.x <= a.hi.x) & (a.lo.y <= b.hi.y) & (b.lo.y <= a.hi.y) & (a.lo.z <= b.hi.z) &
	       (b.lo.z <= a.hi.z);
}

/* Patch of the pair list, thread t = t-th entry of dirtyList (velbox_single_refresh,
 * velbox/box.cpp:95-136, for the boxes that moved): the entity's new box is checked against the
 * whole tree and every neighbour it did not already touch with its OLD box is appended.
 * Invariant kept: the list holds every pair whose current inflated boxes touch.  Pairs that
 * stopped touching stay until the next full build (the narrow phase rejects them). */
template <class Emit> static BT_HD void t_pairs_patch(const DevWorld &W, int t, i32 *stack, int stride, Emit &emit) {
	if ((u32)t >= W.ctr[CT_DIRTY_N] || W.n_leaves < 2) return;
	i32 a = W.dirtyList[t];
	if (!bt_is_leaf_ent(W, a)) return;
	const u32 stamp = W.ctr[CT_PATCH_EPOCH] + 1u;
	const Box6 q = W.ibox[a], qo = W.iboxOld[a];
	const u32 wa = (u32)q.lo.w;
	const i32 qlo[3] = {q.lo.x, q.lo.y, q.lo.z}, qhi[3] = {q.hi.x, q.hi.y, q.hi.z};
	int sp = 0, node = W.wroot[wa];
	if (node < 0) return;
	while (1) {
		const NodeRegs N = bt_load_node(W, node);
		i32 li = N.d.x, ri = N.d.y;
		int ovL = bt_touch_left(N, qlo, qhi, wa);
		int ovR = bt_touch_right(N, qlo, qhi, wa);
#pragma unroll 1
		for (int side = 0; side < 2; side++) {
			i32 ci = side ? ri : li;
			if (!(side ? ovR : ovL) || ci >= 0) continue;
			i32 b = W.leafEnt[~ci];
			if (b == a) continue;
			int bdirty = W.dirtyStamp[b] == stamp;
			const Box6 bo = bdirty ? W.iboxOld[b] : W.ibox[b];
			if (bt_box_touch(qo, bo)) continue; /* they touched before this patch: already cached */
			if (bdirty && b < a) continue;       /* both moved: the lower index emits */
			emit(a, b);
		}
		int goL = ovL && li >= 0, goR = ovR && ri >= 0;
		if (goL) {
			node = li;
			if (goR) {
				if (sp < BT_STACK) stack[(sp++) * stride] = ri;
				else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
			}
		} else if (goR) {
			node = ri;
		} else {
			if (sp == 0) break;
			node = stack[(--sp) * stride];
		}
	}
}

/* Small-world mode (every world <= BT_SMALL_WORLD entities, the batched-worlds case): the candidate
 * pairs of a world are all pairs of its leaf entities whose inflated boxes touch -- the same set the
 * tree walk yields -- found by brute force over the world's contiguous slot range, no sort, no
 * tree, no refit.  This is the one-thread restatement (emulator); k_pairs_small does it with one
 * block per world out of shared memory. */
template <class Emit> static BT_HD void t_pairs_small_world(const DevWorld &W, int g, Emit &emit) {
	if (g >= W.n_worlds || !W.ctr[CT_CACHE_INVALID]) return;
	i32 b0 = W.wbegin[g], b1 = W.wend[g];
	for (i32 e = b0; e < b1; e++)
		if (bt_is_leaf_ent(W, e)) W.ibox[e] = bt_inflated_box(W, e);
	for (i32 a = b0; a < b1; a++) {
		if (!bt_is_leaf_ent(W, a)) continue;
		const Box6 qa = W.ibox[a];
		for (i32 b = a + 1; b < b1; b++)
			if (bt_is_leaf_ent(W, b) && bt_box_touch(qa, W.ibox[b])) emit(a, b);
	}
}

template <class Emit> static BT_HD void t_pairs(const DevWorld &W, int t, i32 *stack, int stride, Emit &emit) {
	int L = W.n_leaves;
	if (t >= L || L < 2) return;
	if (!W.ctr[CT_CACHE_INVALID]) return; /* pairs of an earlier iteration / tick are still a superset */
	i32 a = W.leafEnt[t];
	if (!bt_is_leaf_ent(W, a)) return;
	Box6 q = W.leafBox[t];
	u32 wa = (u32)q.lo.w;
	const i32 qlo[3] = {q.lo.x, q.lo.y, q.lo.z}, qhi[3] = {q.hi.x, q.hi.y, q.hi.z};
	int sp = 0, node = W.wroot[wa];
	if (node < 0) return;
	int cur_hi = W.nodeHi[node];
	while (1) {
		const NodeRegs N = bt_load_node(W, node);
		i32 li = N.d.x, ri = N.d.y;
		int gamma = li < 0 ? ~li : li; /* Karras numbering: the split position is the left child's id */
		int ovL = (gamma > t) & bt_touch_left(N, qlo, qhi, wa);
		int ovR = (cur_hi > t) & bt_touch_right(N, qlo, qhi, wa);
		if (ovL && li < 0) emit(a, W.leafEnt[~li]);
		if (ovR && ri < 0 && ~ri > t) emit(a, W.leafEnt[~ri]);
		int goL = ovL && li >= 0, goR = ovR && ri >= 0;
		if (goL) {
			node = li;
			if (goR) {
				if (sp < BT_STACK) stack[(sp++) * stride] = ri;
				else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
			}
			cur_hi = gamma;
		} else if (goR) {
			node = ri;
		} else {
			if (sp == 0) break;
			node = stack[(--sp) * stride];
			cur_hi = W.nodeHi[node];
		}
	}
}

/* checkCollision ent.cpp:398-435 for one cached pair, from whichever ends are scanning: all
 * loads of both ends first (they are independent), then one copy of the narrow phase run from
 * each scanning end */
/* FIRST = the first doIteration of a tick: the integrate kernel has just set needsCollision on every entity that is
 * in gs->ents (ent.cpp:856), so the two byte gathers -- one dependent hop of the chain pair -> flags -> fields ->
 * proposal -- are skipped; entities without a box or out of the list fall to the inList / hasBox test below */
/* both ends of one candidate pair in registers */
struct NarrowEnd {
	i32 e, root;
	u32 type, collide, hphys;
	int needs; /* it scans in this doIteration */
	EntGeo g;
};
/* the narrow phase proper for a pair whose ends are live, boxed and loaded: one copy of checkCollision, run from each
 * scanning end (ent.cpp:446-449) */
static BT_HD void bt_narrow_eval(const DevWorld &W, NarrowEnd A, NarrowEnd B, LocalStats &st, u32 &any) {
	if (!((B.type & A.collide) | (A.type & B.collide))) { st.cand += (u32)(A.needs + B.needs); return; }
#pragma unroll 1
	for (int k = 0; k < 2; k++) {
		if (A.needs) {
			st.cand++;
			bt_check_loaded(W, A.e, A.g, A.type, A.collide, A.root, A.hphys, B.e, B.g, B.type, B.collide, B.root, B.hphys, B.needs, st, any);
		}
		{ NarrowEnd t_ = A; A = B; B = t_; }
	}
}
template <int FIRST> static BT_HD void t_narrow_t(const DevWorld &W, u64 pr, LocalStats &st, u32 &any) {
	NarrowEnd A, B;
	A.e = (i32)(u32)pr; B.e = (i32)(pr >> 32);
	A.needs = FIRST ? 1 : W.needsCollision[A.e]; B.needs = FIRST ? 1 : W.needsCollision[B.e];
	if (!(A.needs | B.needs)) return; /* the common case after the first iteration: two byte loads */
	/* everything else in ONE batch of independent loads (no branch in between, or the compiler
	 * must wait for each test before it may issue the next gather) */
	int ia = W.inList[A.e], ba = W.hasBox[A.e], ib = W.inList[B.e], bb = W.hasBox[B.e];
	A.type = W.type[A.e]; A.collide = W.collide[A.e]; B.type = W.type[B.e]; B.collide = W.collide[B.e];
	A.root = W.holdRoot[A.e]; B.root = W.holdRoot[B.e];
	A.hphys = W.hphys[A.e]; B.hphys = W.hphys[B.e];
	A.g = bt_geo(W, A.e); B.g = bt_geo(W, B.e);
	if (!(ia & ba & ib & bb)) return; /* `o->dead` ent.cpp:455 */
	bt_narrow_eval(W, A, B, st, any);
}

static BT_HD void t_narrow(const DevWorld &W, u64 pr, LocalStats &st, u32 &any) { t_narrow_t<0>(W, pr, st, any); }

/* ------------------------------------------------------------------------- */
/* resolution                                                                */

/* shouldResolveCollision ent.cpp:631-672.  Returns 0 no / 1 yes / 2 yes, decided by
 * walking a cycle of pending collisions back to e. */
static BT_HD int bt_should_resolve(const DevWorld &W, i32 e, u64 pe) {
	Coll mine = bt_unpack_coll(pe);
	if (mine.mutual == 2) return 0;
	i32 otherRoot = W.holdRoot[mine.buddy];
	u64 po = W.pend[otherRoot];
	if (po == 0) return 1;
	i32 loopDetector = otherRoot;
	int flip = 0;
	while (1) {
		Coll other = bt_unpack_coll(po);
		if (bt_collision_cmp(bt_view(W), other, mine) > 0) return 0;
		otherRoot = W.holdRoot[other.buddy];
		if (otherRoot == e) return 2;
		po = W.pend[otherRoot];
		if (po == 0 || otherRoot == loopDetector) return 0;
		if (flip) loopDetector = W.holdRoot[bt_unpack_coll(W.pend[loopDetector]).buddy];
		flip = !flip;
	}
}

/* every entity: doIteration clears needsCollision (ent.cpp:459); roots with a pending
 * collision decide whether they resolve now (ent.cpp:870-878) */
static BT_HD void t_decide(const DevWorld &W, i32 e) {
	if (e >= W.n) return;
	W.needsCollision[e] = 0;
	u64 p = W.pend[e];
	W.ready[e] = p ? (u8)bt_should_resolve(W, e, p) : 0;
}

/* pinch ent.cpp:483-507 */
static BT_HD i32 bt_pinch(const DevWorld &W, i32 e, int axis, i32 o, u32 key, LocalStats &st) {
	i32 leaf1 = e;
	while (bt_get3(W.forcedHoldees[leaf1], axis) >= 0) leaf1 = bt_get3(W.forcedHoldees[leaf1], axis);
	i32 leaf2;
	if (o >= 0) {
		leaf2 = o;
		while (bt_get3(W.forcedHoldees[leaf2], axis) >= 0) leaf2 = bt_get3(W.forcedHoldees[leaf2], axis);
	} else {
		leaf2 = e;
	}
	if (leaf1 != e) {
		bt_fumble(W, leaf1, st);
		if (leaf2 != e) {
			bt_fumble(W, leaf2, st);
			return leaf2;
		}
		return e;
	}
	if (leaf2 != e) {
		bt_fumble(W, leaf2, st);
		return leaf2;
	}
	bt_crush(W, e, key, st);
	return -1;
}

/* push_helper ent.cpp:509-525 */
static BT_HD int bt_push_helper(const DevWorld &W, i32 e, i32 prev, int axis, int dir, u32 key, LocalStats &st) {
	if (bt_forced_get(W.forced[e], axis) * dir < 0) {
		i32 ret = bt_pinch(W, e, axis, prev, key, st);
		if (ret < 0) return 1;
		if (e != ret) return 0;
	}
	int f = bt_forced_get(W.forced[e], axis) + dir;
	W.forced[e] = bt_forced_set(W.forced[e], axis, f);
	if (f == dir * 15) { /* "Safety crush!" */
		bt_crush(W, e, key, st);
		return 1;
	}
	int4 fh = W.forcedHoldees[e];
	bt_set3(fh, axis, prev);
	W.forcedHoldees[e] = fh;
	return 0;
}

/* getPushDv ent.cpp:527-540 */
static BT_HD void bt_push_dv(const DevWorld &W, i32 dv[3], i32 e, i32 o, int axis, int dir, i32 accel) {
	int4 ve = W.vel[e], vo = W.vel[o];
	i32 d[3] = {vo.x - ve.x, vo.y - ve.y, vo.z - ve.z};
	dv[0] = d[0]; dv[1] = d[1]; dv[2] = d[2];
	dv[axis] = 0;
	bt_bound_vec(dv, W.friction[e], 3);
	dv[axis] = accel * dir;
}

/* doPushRegular ent.cpp:551-577 */
static BT_HD void bt_push_regular(const DevWorld &W, i32 e, i32 o, int axis, int dir, i32 displacement,
                                  const i32 dv[3], u32 key, LocalStats &st) {
	i32 prev = -1;
	while (1) {
		if (bt_push_helper(W, e, prev, axis, dir, key, st)) break;
		if ((W.hphys[e] >> 16) & 0xff) {
			i32 accel = dv[axis] * dir;
			if (bt_pushed_handler(W, e, o, axis, dir, displacement, accel)) {
				bt_crush(W, e, key, st);
				break;
			}
		}
		prev = e;
		e = W.holder[e];
		if (e >= 0) {
			i32 hf = W.holdFlags[prev];
			if (hf == HOLD_PASS) continue;
			if (hf == HOLD_DROP) bt_fumble(W, prev, st);
		}
		i32 dc[3] = {0, 0, 0};
		dc[axis] = displacement * dir;
		bt_move_recursive(W, prev, dc, dv);
		break;
	}
}

/* pickupNoHandlers ent.cpp:166-182 + pickup :184-188 (the root list has no counterpart here) */
static BT_HD void bt_pickup(const DevWorld &W, i32 x, i32 y, i32 holdFlags) {
	i32 first = W.holdee[x];
	W.LLn[y] = first;
	W.LLp[y] = -1;
	if (first >= 0) W.LLp[first] = y;
	W.holdee[x] = y;
	W.holder[y] = x;
	i32 r = W.holdRoot[x];
	BT_FOR_SUBTREE(W, y, it, W.holdRoot[it] = r);
	W.holdFlags[y] = HOLD_MASK & holdFlags;
	bt_add32(&W.ctr[CT_TREE_EDITS], 1u);
	if (holdFlags & HOLD_FREEZE) {
		int4 vx = W.vel[x], vy = W.vel[y];
		W.d_vel[y] = make_int4(vx.x - vy.x, vx.y - vy.y, vx.z - vy.z, 0);
		W.ctr[CT_ANY_DELTA] = 1; /* the flush2 that follows has something to apply */
	}
	bt_pair_handler(W, 2, x, y);
	bt_pair_handler(W, 3, y, x);
}

/* doPushPickup ent.cpp:542-549 */
static BT_HD void bt_push_pickup(const DevWorld &W, i32 e, i32 holder, u32 key, LocalStats &st) {
	if (W.holder[e] >= 0 || W.holdRoot[holder] == e) {
		bt_crush(W, e, key, st);
		return;
	}
	/* The reference runs every push it decided on (ent.cpp:879-882) even if an earlier push of
	 * the same iteration killed the would-be holder: e->holder then dangles once clearDeads frees
	 * it (use-after-free in the reference; observed with 2048 arena worlds, one case in ~10^5
	 * world-ticks).  Done the same way here, the entity ends up attached to a retired slot,
	 * and counted, because from this point on the reference's own behaviour is undefined. */
	if (W.dead[holder]) bt_unsup(W, e, 7, 0);
	bt_pickup(W, holder, e, 0);
}

/* push ent.cpp:579-629 */
static BT_HD void bt_push(const DevWorld &W, const Coll &c, u32 key, LocalStats &st) {
	st.push++;
	i32 e = c.leaf, o = c.buddy;
	int axis = c.axis, dir = c.dir;
	/* the push is a chain of dependent gathers through handler code the compiler cannot
	 * reorder; start everything it is going to touch on its way to L1 now */
	bt_prefetch(&W.forced[e]); bt_prefetch(&W.friction[e]); bt_prefetch(&W.holder[e]); bt_prefetch(&W.hphys[e]);
	bt_prefetch(&W.hphys[o]); bt_prefetch(&W.ibox[e]); bt_prefetch(&W.old[e]);
	bt_prefetch(&W.forcedHoldees[e]); bt_prefetch(&W.holdee[e]); bt_prefetch(&W.hasBox[e]);
	/* (not type[o] and LLn[e]: a plain push -- handler NIL, nothing held -- never reads them, and the kernel is bound by
	 * the sectors it moves) */
	i32 displacement = (bt_get3(W.center[o], axis) - bt_get3(W.center[e], axis)) * dir + bt_get3(W.radius[e], axis) +
	                   bt_get3(W.radius[o], axis);
	i32 accel = (bt_get3(W.vel[o], axis) - bt_get3(W.vel[e], axis)) * dir;
	if (accel < 0) accel = 0;
	if (c.mutual) {
		int p1 = bt_push_handler(W, o, e);
		int p2 = bt_push_handler(W, e, o);
		if (p1 != p2) {
			if (p1) { i32 t_ = e; e = o; o = t_; } /* the branchless swap of ent.cpp:601-603 */
			dir *= (p2 - p1);
			i32 dv[3];
			bt_push_dv(W, dv, e, o, axis, dir, accel);
			bt_push_regular(W, e, o, axis, dir, displacement, dv, key, st);
			bt_push_pickup(W, o, e, key, st);
		} else {
			i32 dv1[3], dv2[3];
			bt_push_dv(W, dv1, e, o, axis, dir, accel);
			bt_push_dv(W, dv2, o, e, axis, -dir, accel);
			bt_push_regular(W, e, o, axis, dir, displacement, dv1, key, st);
			bt_push_regular(W, o, e, axis, -dir, displacement, dv2, key, st);
		}
	} else {
		int p1 = bt_push_handler(W, o, e);
		if (p1) {
			bt_push_pickup(W, e, o, key, st);
		} else {
			i32 dv[3];
			bt_push_dv(W, dv, e, o, axis, dir, accel);
			bt_push_regular(W, e, o, axis, dir, displacement, dv, key, st);
		}
	}
}

/* event-order key of a push: iteration counter, then descending root index */
static BT_HD u32 bt_push_key(const DevWorld &W, i32 root) {
	return ((u32)(W.iter_id - W.iter_base) << 20) ^ (u32)(0xfffff - (root & 0xfffff));
}

/* parallel lane: one thread per ready root (ent.cpp:879-882).  A root that resolves by
 * closing a cycle (ready == 2) first checks the other members of its cycle: if one of them
 * resolves too (a tie) both go to the serial lane.  ready[] of other roots is only ever
 * rewritten 2 -> 1 by their own thread here, and the test is "non-zero", so the concurrent
 * reads are safe. */
/* 1: root e pushes now, in the parallel lane; 0: it is not ready, or it was handed to the serial lane */
static BT_HD int bt_push_route(const DevWorld &W, i32 e) {
	if (e >= W.n) return 0;
	u8 rd = W.ready[e];
	if (!rd) return 0;
	Coll pc = bt_unpack_coll(W.pend[e]);
	/* push handlers that edit somebody else's tree (player_push: pickup into the player's
	 * holdee list, first come first served) run in the serial lane, canonical order */
	int serial_handler = ((BT_SERIAL_PUSH >> ((W.hphys[pc.buddy] >> 8) & 0x1f)) & 1) ||
	                     (pc.mutual && ((BT_SERIAL_PUSH >> ((W.hphys[pc.leaf] >> 8) & 0x1f)) & 1));
	if (serial_handler) {
		u32 slot = bt_add32(&W.ctr[CT_SERIAL_N], 1u);
		if ((i32)slot < W.serial_cap) W.serialList[slot] = e;
		else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
		return 0;
	}
	if (rd == 2) {
		i32 r = W.holdRoot[bt_unpack_coll(W.pend[e]).buddy];
		int conflict = 0, guard = 0;
		while (r != e && guard++ < 1000000) {
			if (W.ready[r]) conflict = 1;
			r = W.holdRoot[bt_unpack_coll(W.pend[r]).buddy];
		}
		if (conflict) {
			u32 slot = bt_add32(&W.ctr[CT_SERIAL_N], 1u);
			if ((i32)slot < W.serial_cap) W.serialList[slot] = e;
			else W.ctr[CT_ERROR] = (u32)(-BT_ERR_OVERFLOW);
			return 0;
		}
	}
	return 1;
}
/* the push of ready root e (either lane) */
static BT_HD void bt_push_root(const DevWorld &W, i32 e, LocalStats &st) {
	st.rk_ordered = 0; st.rk_local = 0; st.rk_base = bt_push_rootkey(W, e);
	bt_push(W, bt_unpack_coll(W.pend[e]), bt_push_key(W, e), st);
}
static BT_HD void t_push(const DevWorld &W, i32 e, LocalStats &st) {
	if (bt_push_route(W, e)) bt_push_root(W, e, st);
}

/* serial lane: one thread, descending root index (DESIGN.md "Canonical order" B3) */
static BT_HD i32 bt_serial_sort(const DevWorld &W) {
	i32 n = (i32)W.ctr[CT_SERIAL_N];
	if (n > W.serial_cap) n = W.serial_cap;
	for (i32 i = 1; i < n; i++) { /* insertion sort, descending */
		i32 v = W.serialList[i];
		i32 j = i - 1;
		while (j >= 0 && W.serialList[j] < v) { W.serialList[j + 1] = W.serialList[j]; j--; }
		W.serialList[j + 1] = v;
	}
	return n;
}
static BT_HD void t_push_serial(const DevWorld &W, LocalStats &st) {
	const i32 n = bt_serial_sort(W);
	for (i32 i = 0; i < n; i++) bt_push_root(W, W.serialList[i], st);
	W.ctr[CT_SERIAL_N] = 0;
}

/* ent.cpp:883-887: every root that had a pending collision forgets it and re-arms
 * collision checks for its (post-push) tree */
static BT_HD void t_post(const DevWorld &W, i32 e) {
	if (e >= W.n || W.pend[e] == 0) return;
	W.pend[e] = 0;
	W.ready[e] = 0;
	BT_FOR_SUBTREE(W, e, it, W.needsCollision[it] = 1); /* requireCollisionRecursive ent.cpp:128-134 */
}

/* ------------------------------------------------------------------------- */
/* tick handlers (doUpdates ent.cpp:826-832), SURVEY.md §8(f)-1: the ones that only read flushed
 * state (buttons .v, sliders .v, centres) and write commutative requests (pushBtn = OR, uSlider =
 * min/max, uVel = add, uDead = OR, uDrop), so every entity can run its handler at the same time.
 * TICK_RAND draws from gs->rand and therefore runs per world, in list order (t_tick_serial). */
enum {
	TICK_EXPLOSION = 1, TICK_PLATFORM = 3, TICK_THRUST = 10, TICK_BULLET = 25, TICK_HELD_BULLET = 26, TICK_DUST = 32,
	TICK_INCINERATE = 40, TICK_BOTTLE_HELD = 41, TICK_LEGG_OLD = 16, TICK_LEGG = 22, TICK_FACTORY = 20, TICK_FACTORY_WIRES = 21, TICK_FLAG_SPAWNER = 2,
	TICK_RAND = 6, TICK_EYE = 8, TICK_HELD_EYE = 9, TICK_HELD_VEH_EYE = 18, TICK_LOGIC = 11, TICK_LOGIC_DEBUG = 12, TICK_DOOR = 13, TICK_CURSED = 14, TICK_RESPAWN = 15,
	TICK_SEAT_OLD = 17, TICK_ADJUST = 19, TICK_SEAT = 23, TICK_HELD_GUN = 24, TICK_SCOREBOARD = 27, TICK_TIMER = 28,
	TICK_DEMOLISH = 29, TICK_HELD_BLINK = 30, TICK_HELD_JUMPER = 31, TICK_SPIDER = 33, TICK_SIGN = 34, TICK_LOADBLOCK = 35, TICK_PUDDLEJUMP = 36
};
/* NIL + the ids above */
#define BT_SUPPORTED_TICK ((1ull << 0) | (1ull << TICK_RAND) | (1ull << TICK_LOGIC) | (1ull << TICK_LOGIC_DEBUG) | \
	(1ull << TICK_DOOR) | (1ull << TICK_CURSED) | (1ull << TICK_RESPAWN) | (1ull << TICK_SEAT_OLD) |                  \
	(1ull << TICK_ADJUST) | (1ull << TICK_SEAT) | (1ull << TICK_TIMER) | (1ull << TICK_DEMOLISH) | (1ull << TICK_SIGN) | \
	(1ull << TICK_HELD_GUN) | (1ull << TICK_SCOREBOARD) | (1ull << TICK_HELD_BLINK) | (1ull << TICK_HELD_JUMPER) |        \
	(1ull << TICK_LOADBLOCK) | (1ull << TICK_PUDDLEJUMP) | (1ull << TICK_SPIDER) | (1ull << TICK_EYE) |                  \
	(1ull << TICK_HELD_EYE) | (1ull << TICK_HELD_VEH_EYE) | (1ull << TICK_EXPLOSION) | (1ull << TICK_PLATFORM) |           \
	(1ull << TICK_THRUST) | (1ull << TICK_BULLET) | (1ull << TICK_HELD_BULLET) | (1ull << TICK_DUST) |                   \
	(1ull << TICK_BOTTLE) | (1ull << TICK_INCINERATE) | (1ull << TICK_BOTTLE_HELD) | (1ull << TICK_LEGG_OLD) |          \
	(1ull << TICK_LEGG) | (1ull << TICK_FACTORY) | (1ull << TICK_FACTORY_WIRES) | (1ull << TICK_FLAG_SPAWNER) |          \
	(1ull << TICK_PLAYER))
/* the ones that run per world, in list order, after everybody else (t_tick_serial): they draw from gs->rand or mark many
 * other entities at once.  (TICK_DEPLETE, logic.c:110-121, writes a slider of its HOLDER directly, which the holder's own
 * handler may read in the same tick before or after, by list order: not reproducible from a parallel pass, left to the host.) */
#define BT_SERIAL_TICK ((1ull << TICK_RAND) | (1ull << TICK_SCOREBOARD) | (1ull << TICK_INCINERATE) | (1ull << TICK_FACTORY) | \
	(1ull << TICK_FACTORY_WIRES))

static BT_HD int bt_button(const DevWorld &W, i32 e, int i) { return (int)(int8_t)((W.btn_v[e] >> (8 * i)) & 0xffu); }
static BT_HD void bt_push_btn(const DevWorld &W, i32 e, int i) { bt_or32(&W.btn_v2[e], 1u << (8 * i)); }
/* uSlider on an entity other threads may be writing too */
static BT_HD void bt_uslider_shared(const DevWorld &W, i32 e, int ix, i32 value) {
	int ns = W.slider_off[e + 1] - W.slider_off[e];
	if (ix < 0 || ix >= ns) return;
	W.fullFlush[e] = 1;
	W.ctr[CT_FLUSH_HINT] = 1;
	int s = W.slider_off[e] + ix;
	bt_max32(&W.smax[s], value);
	bt_min32(&W.smin[s], value);
}
static BT_HD void bt_udead(const DevWorld &W, i32 e) {
	if (W.flip[W.world[e]]) W.dead_max1[e] |= 1; else W.dead_max0[e] |= 1;
	W.wflush[W.world[e]] = 1;
	W.ctr[CT_FLUSH_HINT] = 1;
}
/* logic_common_input logic.c:51-61 / logic_common_output :63-80 */
static BT_HD int bt_logic_input(const DevWorld &W, i32 me, i32 mode) {
	int a = bt_button(W, me, 0), b = bt_button(W, me, 1);
	if ((mode | 1) == 1) return (a ^ b ^ mode) != 0;
	return a && b;
}
static BT_HD void bt_logic_output(const DevWorld &W, i32 me, int sl) {
	i32 outputs = bt_getslider(W, me, sl);
	for (i32 k = 0; k < W.wire_cnt[me]; k++) {
		i32 w = W.wire_to[W.wire_off[me] + k];
		if (outputs & 1) bt_push_btn(W, w, 0);
		if (outputs & 2) bt_push_btn(W, w, 1);
	}
	bt_push_btn(W, me, 2);
}
/* getApproachSpeed / getStoppingDist modules/common.c:6-37 */
static BT_HD i32 bt_approach_speed(const DevWorld &W, i32 e, i32 d, i32 a) {
	if (a <= 0 || d <= 0) return 0;
	i32 q = d / a;
	if (q >= (1 << 28)) { bt_unsup(W, e, 6, TICK_DOOR); return 0; } /* 1 + 8q overflows: sqrt of a negative in the reference */
	i32 n = ((i32)sqrt((double)(1 + 8 * q)) - 1) / 2;
	i32 r = (d - n * (n + 1) / 2 * a) / (n + 1);
	return n * a + r;
}
static BT_HD i32 bt_stopping_dist(i32 s, i32 a) {
	if (s <= 0 || a <= 0) return 0;
	i32 n = s / a, r = s % a;
	return n * (n + 1) / 2 * a + r * (n + 1);
}

/* door_tick door.c:14-62 */
static BT_HD void bt_door_tick(const DevWorld &W, i32 me) {
	i32 nw = W.wire_cnt[me];
	if (!nw) return;
	i32 pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0};
	int4 mc = W.center[me], mv = W.vel[me];
	int numActive = 0;
	for (i32 k = 0; k < nw; k++) {
		i32 w = W.wire_to[W.wire_off[me] + k];
		int4 wc = W.center[w], wv = W.vel[w];
		if (bt_button(W, w, 2)) {
			if (!numActive) vel[0] = vel[1] = vel[2] = 0;
			numActive++;
			pos[0] += wc.x - mc.x; pos[1] += wc.y - mc.y; pos[2] += wc.z - mc.z;
			vel[0] += wv.x - mv.x; vel[1] += wv.y - mv.y; vel[2] += wv.z - mv.z;
		} else if (!numActive) {
			vel[0] += wv.x - mv.x; vel[1] += wv.y - mv.y; vel[2] += wv.z - mv.z;
		}
	}
	if (!numActive) numActive = nw;
	for (int i = 0; i < 3; i++) { pos[i] /= numActive; vel[i] /= numActive; }
	i32 speed = bt_getslider(W, me, 0), accel = bt_getslider(W, me, 1);
	i32 dist = 0;
	for (int i = 0; i < 3; i++) { i32 x = bt_abs(pos[i]); if (x > dist) dist = x; }
	if (dist && bt_stopping_dist(speed, accel) > dist) speed = bt_approach_speed(W, me, dist, accel);
	bt_bound_vec(pos, speed, 3);
	for (int i = 0; i < 3; i++) vel[i] = pos[i] + vel[i];
	bt_bound_vec(vel, accel, 3);
	int4 dv = W.d_vel[me]; /* uVel(me, vel) */
	dv.x += vel[0]; dv.y += vel[1]; dv.z += vel[2];
	W.d_vel[me] = dv;
	W.ctr[CT_ANY_DELTA] = 1;
}

/* scoreboard_tick scoreboard.cpp:37-90 with gs->players.num == 0 (like TICK_RESPAWN): nobody is alive.  Runs in the
 * per-world serial pass: at the end of a round it erases every other entity of the world (uErase sets dead_max to 3,
 * uDead of a handler running at the same time ORs 1 into the same byte: either order gives 3 in the reference). */
static BT_HD void bt_scoreboard_tick(const DevWorld &W, i32 me) {
	const i32 alive = 0;
	i32 timer = bt_getslider(W, me, 1);
	if (bt_getslider(W, me, 0) != alive) { bt_uslider_shared(W, me, 0, alive); timer = 75; }
	if (!timer) return;
	timer--;
	bt_uslider_shared(W, me, 1, timer);
	if (timer) return;
	/* the round is over, a tie: everything but the board is erased (uErase entUpdaters.c:25-27) ... */
	const int g = W.world[me];
	for (i32 e = W.wbegin[g]; e < W.wend[g]; e++) {
		if (e == me || !W.inList[e]) continue;
		if (W.flip[g]) W.dead_max1[e] = 3; else W.dead_max0[e] = 3;
	}
	W.wflush[g] = 1;
	bt_uslider_shared(W, me, 0, 0);
	/* ... requestReload + showMessage belong to the host: an event (kind 6 = tick handler, arg = the alive mask) */
	const u32 slot = bt_add32(&W.ctr[CT_EVENT_N], 1u);
	if ((i32)slot < W.event_cap) W.events[slot] = make_int4(0x600 | TICK_SCOREBOARD, me, me, alive);
	W.fullFlush[me] = 1; /* uMyCollideMask(me, T_TERRAIN) entUpdaters.c:80 */
	W.ncollide[me] = T_TERRAIN;
	W.ctr[CT_FLUSH_HINT] = 1;
}

/* legg_tick / legg_tick_old legg.cpp:101-164 (a leg: tickHeld).  o = 2 / 1: where the gait sliders start.  The "fancy"
 * max / min of legg.cpp:30-39 are kept as written (they wrap like the reference's on huge values). */
static BT_HD i32 bt_legg_max(i32 a, i32 b) { return (a + b + bt_abs(a - b)) / 2; }
static BT_HD i32 bt_legg_min(i32 a, i32 b) { return (a + b - bt_abs(a - b)) / 2; }
static BT_HD i32 bt_bound1(i32 x, i32 b) { return x > b ? b : (x < -b ? -b : x); } /* bound entGetters.c:7-11 */
static BT_HD void bt_legg_tick(const DevWorld &W, i32 me, int old_style) {
	const i32 h = W.holder[me];
	if (h < 0) return;
	i32 v1[3] = {0, 0, 0};
	if (old_style) { /* getInputOld :42-52: the wired entity's offset from ITS holder */
		for (i32 k = 0; k < W.wire_cnt[me]; k++) {
			const i32 w = W.wire_to[W.wire_off[me] + k], wh = W.holder[w];
			if (wh < 0) continue;
			const int4 a = W.center[wh], b = W.center[w];
			v1[0] += b.x - a.x; v1[1] += b.y - a.y;
		}
	} else { /* getInput :55-63 */
		const i32 width = bt_getslider(W, me, 1);
		for (i32 k = 0; k < W.wire_cnt[me]; k++) {
			const i32 w = W.wire_to[W.wire_off[me] + k];
			v1[0] += bt_getslider(W, w, 0) * width / 64; v1[1] += bt_getslider(W, w, 1) * width / 64;
		}
	}
	int o = old_style ? 1 : 2;
	i32 counter = bt_getslider(W, me, 0);
	i32 height = bt_getslider(W, me, o), stepHeight = bt_getslider(W, me, o + 1), tread = bt_getslider(W, me, o + 2);
	i32 s = bt_getslider(W, me, o + 3);
	const i32 a = bt_getslider(W, me, o + 4), time = bt_getslider(W, me, o + 5), v_time = bt_getslider(W, me, o + 6);
	const int hasInput = v1[0] || v1[1];
	const i32 half = time + v_time;
	if (hasInput && half) {
		counter++;
		while (counter >= 2 * half) counter -= 2 * half;
		bt_uslider_shared(W, me, 0, counter);
		if (counter >= time) {
			tread = 0;
			if (counter < half + time) height -= stepHeight;
		}
		if (counter < half) { v1[0] *= -1; v1[1] *= -1; }
	} else {
		tread = 0;
	}
	v1[2] = height;
	/* move :65-99 */
	const int4 mc = W.center[me], hc = W.center[h], mv = W.vel[me], hv = W.vel[h];
	i32 vel[3] = {hc.x - mc.x, hc.y - mc.y, hc.z - mc.z}, acc[3];
	for (int i = 0; i < 3; i++) v1[i] += vel[i];
	vel[0] = mv.x - hv.x; vel[1] = mv.y - hv.y; vel[2] = mv.z - hv.z;
	i32 v_s = s;
	const i32 dist = bt_legg_max(bt_abs(v1[0]), bt_abs(v1[1]));
	if (dist) s = bt_legg_min(s, bt_approach_speed(W, me, dist, a));
	bt_bound_vec(v1, s, 2);
	for (int i = 0; i < 2; i++) acc[i] = v1[i] - vel[i];
	bt_bound_vec(acc, a, 2);
	const i32 new_speed = bt_legg_max(bt_abs(vel[0] + acc[0]), bt_abs(vel[1] + acc[1]));
	if (s) v1[2] += tread * new_speed / s;
	if (v1[2]) v_s = bt_legg_min(v_s, bt_approach_speed(W, me, bt_abs(v1[2]), a));
	v1[2] = bt_bound1(v1[2], v_s);
	acc[2] = bt_bound1(v1[2] - vel[2], a);
	int4 dv = W.d_vel[me]; /* uVel(me, acc) */
	dv.x += acc[0]; dv.y += acc[1]; dv.z += acc[2];
	W.d_vel[me] = dv;
	W.ctr[CT_ANY_DELTA] = 1;
}

/* player_tick player.c:229-443.  The controls are sliders and buttons the host's doInputs writes (axis 0 / 1, look 2..4,
 * the extra buttons in slider 12; jump = button 0, shift = 1, triggers = 2 / 3), so walking, jumping, dropping and
 * swapping equipment, charge and cooldown all run here.  What creates entities does not: a stackem block or a platform
 * fired (mkStackem / mkPlatform, the second draws from gs->rand) and the edit tool's baubles are counted as unsupported,
 * with the sliders advanced as the reference would have. */
static BT_HD int bt_equip_hands(const DevWorld &W, i32 e) { /* getEquipHands :217-227 */
	const u32 f = W.type[e] & EQUIP_MASK;
	if (!f) return 0;
	if (f == EQUIP_MASK) return 3;
	return 1 + (bt_getslider(W, e, 0) != 0);
}
static BT_HD void bt_uvel_other(const DevWorld &W, i32 e, const i32 *v, int center_too) {
	i32 *dv = (i32 *)&W.d_vel[e], *dc = (i32 *)&W.d_center[e];
	for (int i = 0; i < 3; i++) {
		bt_add32((u32 *)&dv[i], (u32)v[i]);
		if (center_too) bt_add32((u32 *)&dc[i], (u32)v[i]);
	}
	W.ctr[CT_ANY_DELTA] = 1;
}
static BT_HD void bt_player_tick(const DevWorld &W, i32 me, int id) {
	const int grounded = bt_getslider(W, me, 7) > 0;
	bt_uslider_shared(W, me, 7, 0);
	{
		const i32 axis[2] = {bt_getslider(W, me, 0), bt_getslider(W, me, 1)};
		i32 sl[2] = {bt_getslider(W, me, 5), bt_getslider(W, me, 6)};
		i32 speed;
		if (!grounded) { bt_bound_vec(sl, 256, 2); speed = 256; } /* PLAYER_SPD */
		else {
			speed = bt_abs(sl[0]) > bt_abs(sl[1]) ? bt_abs(sl[0]) : bt_abs(sl[1]); /* getMagnitude common.c:39-46 */
			if (speed < 256) speed = 256;
		}
		i32 vel[3];
		vel[2] = bt_button(W, me, 0) && grounded ? -384 : 0;
		for (int i = 0; i < 2; i++) vel[i] = speed * axis[i] / 64 - sl[i];
		bt_bound_vec(vel, grounded ? 56 : 16, 2);
		for (int i = 0; i < 2; i++) bt_uslider_shared(W, me, 5 + i, sl[i] + vel[i]);
		int4 dv = W.d_vel[me]; /* uVel(me, vel) */
		dv.x += vel[0]; dv.y += vel[1]; dv.z += vel[2];
		W.d_vel[me] = dv;
		W.ctr[CT_ANY_DELTA] = 1;
	}
	i32 charge = bt_getslider(W, me, 8), cooldown = bt_getslider(W, me, 9);
	const int fire = bt_button(W, me, 2), fire2 = bt_button(W, me, 3), utility = bt_button(W, me, 1);
	const i32 other = bt_getslider(W, me, 12);
	const int altfire = (other & 1) != 0, altfire2 = (other & 2) != 0;
	if (bt_getslider(W, me, 11)) { /* edit mode */
		for (int mode = 0; mode < 2; mode++) { /* player_clearBaubles :198-204 */
			if (!(mode ? altfire2 : altfire)) continue;
			for (i32 h = W.holdee[me]; h >= 0; h = W.LLn[h])
				if ((W.type[h] & 64u) && W.wire_cnt[h] == 1 && mode == bt_getslider(W, h, 0)) { W.newDrop[h] = 1; W.wflush[W.world[h]] = 1; W.ctr[CT_FLUSH_HINT] = 1; }
		}
		if (cooldown >= 3) {
			if (utility && bt_button(W, me, 0)) { /* player_flipBaubles :206-215 */
				for (i32 h = W.holdee[me]; h >= 0; h = W.LLn[h])
					if ((W.type[h] & 64u) && W.wire_cnt[h] == 1) bt_push_btn(W, h, 2);
				cooldown = 0;
			} else if (fire || fire2) { /* edittool_select makes baubles */
				bt_unsup(W, me, 6, id);
				cooldown = 0;
			}
		}
	} else {
		int hands = 0, drops = 0;
		for (i32 h = W.holdee[me]; h >= 0; h = W.LLn[h]) {
			const int eh = bt_equip_hands(W, h);
			drops |= hands & eh;
			hands |= eh;
		}
		i32 ep = bt_getslider(W, me, 10);
		ep &= ~1;
		if (altfire) { if (!(ep & 2)) { ep |= 2; drops |= 1; } } else ep &= ~2;
		if (altfire2) { if (!(ep & 4)) { ep |= 4; drops |= 2; } } else ep &= ~4;
		if (fire) { if (hands & 1) ep |= 8; } else ep &= ~8;
		if (fire2) { if (hands & 2) ep |= 16; } else ep &= ~16;
		bt_uslider_shared(W, me, 10, ep);
		const i32 look[3] = {bt_getslider(W, me, 2), bt_getslider(W, me, 3), bt_getslider(W, me, 4)};
		if (drops) {
			if (drops & hands) {
				const i32 v[3] = {200 * look[0] / 64, 200 * look[1] / 64, 200 * look[2] / 64};
				for (i32 h = W.holdee[me]; h >= 0; h = W.LLn[h])
					if (drops & bt_equip_hands(W, h)) {
						bt_uvel_other(W, h, v, 1);
						W.newDrop[h] = 1; W.wflush[W.world[h]] = 1; W.ctr[CT_FLUSH_HINT] = 1; /* uDrop */
					}
			} else { /* nothing in that hand: move the off-hand item over */
				for (i32 h = W.holdee[me]; h >= 0; h = W.LLn[h])
					if (W.type[h] & EQUIP_MASK) bt_uslider_shared(W, h, 0, drops - 1);
			}
		}
		if (hands == 3) return;
		if (cooldown >= 5 && charge >= 60) {
			const i32 rules = W.gamerules[W.world[me]];
			if (fire && !(hands & 1) && !(ep & 8) && !(rules & (1 << 7))) { /* mkStackem */
				bt_unsup(W, me, 6, id);
				charge -= 60; cooldown = 0;
			} else if (charge >= 180 && fire2 && !(hands & 2) && !(ep & 16) && !(rules & (1 << 8))) { /* mkPlatform, random() */
				bt_unsup(W, me, 6, id);
				charge -= 180; cooldown = 0;
			}
		}
	}
	if (charge < 180) bt_uslider_shared(W, me, 8, charge + 2);
	if (cooldown < 5) bt_uslider_shared(W, me, 9, cooldown + 1);
}

/* one entity's tick / tickHeld handler.  Returns without effect for TICK_RAND and TICK_SCOREBOARD (serial pass). */
static BT_HD void t_tick_id(const DevWorld &W, i32 me, u32 ht, int id);
static BT_HD void t_tick(const DevWorld &W, i32 me) {
	if (me >= W.n || !W.inList[me]) return;
	u32 ht = W.htick[me];
	int id = (int)(W.holder[me] >= 0 ? (ht >> 8) & 0xffu : ht & 0xffu);
	if (id) t_tick_id(W, me, ht, id);
}
/* ... with the handler word and the id (tick or tickHeld, by whether it is held) already read by the caller */
static BT_HD void t_tick_id(const DevWorld &W, i32 me, u32 ht, int id) {
	switch (id) {
	case 0: case TICK_SIGN: case TICK_RAND: return;
	case TICK_LOGIC: case TICK_LOGIC_DEBUG: /* logic_tick logic.c:131-149 (the debug colour is not simulation state) */
		if (bt_logic_input(W, me, bt_getslider(W, me, 0))) bt_logic_output(W, me, 1);
		return;
	case TICK_TIMER: { /* timer_tick logic.c:82-99 */
		if (bt_button(W, me, 0) ^ bt_button(W, me, 1)) { bt_uslider_shared(W, me, 2, bt_getslider(W, me, 0)); return; }
		i32 timer = bt_getslider(W, me, 2);
		if (timer) { bt_uslider_shared(W, me, 2, timer - 1); return; }
		bt_logic_output(W, me, 1);
		return;
	}
	case TICK_ADJUST: /* adjust_tick logic.c:101-109 */
		if (bt_button(W, me, 0) ^ bt_button(W, me, 1)) {
			i32 sl = bt_getslider(W, me, 0), value = bt_getslider(W, me, 1);
			for (i32 k = 0; k < W.wire_cnt[me]; k++) bt_uslider_shared(W, W.wire_to[W.wire_off[me] + k], sl, value);
		}
		return;
	case TICK_DEMOLISH: if (bt_button(W, me, 0)) bt_udead(W, me); return; /* logic.c:170-172 */
	case TICK_CURSED: bt_udead(W, me); return;                             /* common.c:52-54 */
	case TICK_DOOR: bt_door_tick(W, me); return;
	case TICK_RESPAWN: { /* respawn_tick respawn.cpp:24-66; a device world has no player table, so
	                      * nobody is ever waiting to be revived (gs->players.num == 0) */
		i32 cooldown = bt_getslider(W, me, 4);
		if (cooldown > 0) bt_uslider_shared(W, me, 4, cooldown - 1);
		return;
	}
	case TICK_SEAT_OLD: case TICK_SEAT: /* seat_tick(_old) seat.cpp:38-60 */
		for (i32 k = 0; k < W.wire_cnt[me]; k++) {
			i32 w = W.wire_to[W.wire_off[me] + k];
			if (bt_button(W, w, 0)) { W.newDrop[w] = 1; W.wflush[W.world[w]] = 1; W.ctr[CT_FLUSH_HINT] = 1; continue; } /* uDrop */
			if (id == TICK_SEAT_OLD) continue;
			for (int i = 0; i < 5; i++) bt_uslider_shared(W, me, i, bt_getslider(W, w, i));
			for (int i = 0; i < 4; i++)
				if (bt_button(W, w, i)) bt_push_btn(W, me, i);
		}
		return;
	case TICK_EYE: /* eye_tick eyes.cpp:14-22: the first tick an eye is NOT held it becomes debris and retires its handler
	                * ("it's fine to set this directly in a tick handler": a direct write, visible to this tick's hash) */
		W.fullFlush[me] = 1;
		W.ntype[me] = W.type[me] | T_DEBRIS;
		W.ncollide[me] = W.collide[me] | T_TERRAIN | T_OBSTACLE | T_DEBRIS;
		W.htick[me] = ht & ~0xffu;
		W.ctr[CT_FLUSH_HINT] = 1;
		return;
	case TICK_HELD_EYE: { /* eye_tick_held eyes.cpp:24-47: sits on the holder's surface where the holder looks (getLook =
	                       * sliders 2..4, player.c:53-57) */
		const i32 p = W.holder[me];
		if (p < 0) return;
		i32 vec[3] = {bt_getslider(W, p, 2), bt_getslider(W, p, 3), bt_getslider(W, p, 4)};
		const int4 pr = W.radius[p], pc = W.center[p], pv = W.vel[p], mc = W.center[me], mv = W.vel[me];
		const i32 rad[3] = {pr.x, pr.y, pr.z};
		i32 component = bt_abs(vec[2]), r = rad[2] - 50; /* eyeInset */
		for (int i = 0; i < 2; i++) {
			const i32 i_comp = bt_abs(vec[i]), i_r = rad[i] - 50;
			if (i_comp * r > component * i_r) { r = i_r; component = i_comp; }
		}
		if (!component) return;
		const i32 pos[3] = {mc.x - pc.x, mc.y - pc.y, mc.z - pc.z}, vel[3] = {mv.x - pv.x, mv.y - pv.y, mv.z - pv.z};
		int4 dv = W.d_vel[me]; /* uVel(me, vec) */
		dv.x += r * vec[0] / component - pos[0] - vel[0];
		dv.y += r * vec[1] / component - pos[1] - vel[1];
		dv.z += r * vec[2] / component - pos[2] - vel[2];
		W.d_vel[me] = dv;
		W.ctr[CT_ANY_DELTA] = 1;
		return;
	}
	case TICK_HELD_VEH_EYE: { /* vehicle_eye_tick_held eyes.cpp:49-66: wired to a seat that is wired to (at most) a pilot */
		const i32 p = W.holder[me];
		if (p < 0) return;
		i32 vec[3] = {0, 0, 0};
		for (i32 k = 0; k < W.wire_cnt[me]; k++) {
			const i32 w = W.wire_to[W.wire_off[me] + k];
			for (i32 j = 0; j < W.wire_cnt[w]; j++) {
				const i32 e2 = W.wire_to[W.wire_off[w] + j];
				vec[0] += bt_getslider(W, e2, 0); vec[1] += bt_getslider(W, e2, 1); /* getAxis */
			}
		}
		const int4 pr = W.radius[p], pc = W.center[p], mc = W.center[me];
		int4 dc = W.d_center[me]; /* uCenter(me, vec) */
		dc.x += vec[0] * pr.x / 64 - (mc.x - pc.x);
		dc.y += vec[1] * pr.y / 64 - (mc.y - pc.y);
		dc.z += vec[2] * pr.z / 64 - (mc.z - pc.z);
		W.d_center[me] = dc;
		W.ctr[CT_ANY_DELTA] = 1;
		return;
	}
	case TICK_FLAG_SPAWNER: { /* flagSpawner_tick flag.c:36-45: idle while its flag (its one wire) exists; making a new one is
	                           * entity creation: host side, counted */
		const i32 cooldown = bt_getslider(W, me, 0);
		if (cooldown == 0) {
			if (!W.wire_cnt[me]) bt_unsup(W, me, 6, id);
			return;
		}
		bt_uslider_shared(W, me, 0, cooldown - 1);
		return;
	}
	case TICK_PLAYER: bt_player_tick(W, me, id); return;
	case TICK_LEGG: bt_legg_tick(W, me, 0); return;
	case TICK_LEGG_OLD: bt_legg_tick(W, me, 1); return;
	case TICK_SPIDER: { /* spider_tick spider.cpp:58-122 (a leg: tickHeld): a spring towards where the gait wants the foot,
	                     * the reaction on the body; several legs add to one body, so uVel is an atomic add here */
		const i32 h = W.holder[me];
		if (h < 0) return;
		const i32 height_top = bt_getslider(W, me, 0), height_bottom = bt_getslider(W, me, 1), width = bt_getslider(W, me, 2);
		const i32 t_top = bt_getslider(W, me, 3), t_bottom = bt_getslider(W, me, 4);
		i32 counter = bt_getslider(W, me, 5), spring = bt_getslider(W, me, 6), ratio_legs = bt_getslider(W, me, 9);
		const i32 drag = bt_getslider(W, me, 7), a_max = bt_getslider(W, me, 8);
		if (spring <= 0) spring = 1;
		if (ratio_legs <= 0) ratio_legs = 1;
		i32 v1[3] = {0, 0, 0}, vel[3] = {0, 0, 0};
		for (i32 k = 0; k < W.wire_cnt[me]; k++) { /* getInput :30-39 */
			const i32 w = W.wire_to[W.wire_off[me] + k];
			v1[0] += bt_getslider(W, w, 0) * width / 64;
			v1[1] += bt_getslider(W, w, 1) * width / 64;
		}
		const int hasInput = v1[0] || v1[1];
		v1[2] = height_bottom;
		const i32 t_cycle = t_top + t_bottom;
		if (hasInput && t_cycle) {
			counter++;
			while (counter >= t_cycle) counter -= t_cycle;
			bt_uslider_shared(W, me, 5, counter);
			i32 numer, denom;
			if (counter >= t_bottom) { v1[2] = height_top; denom = t_top; numer = counter - t_bottom; }
			else { v1[0] *= -1; v1[1] *= -1; denom = t_bottom; numer = counter; }
			for (int i = 0; i < 2; i++) vel[i] = v1[i] / denom;
			for (int i = 0; i < 2; i++) v1[i] = (2 * numer - denom) * v1[i] / denom;
		}
		{ /* computeAccel :41-56 */
			const int4 mc = W.center[me], hc = W.center[h], mv = W.vel[me], hv = W.vel[h];
			const i32 cp[3] = {mc.x - hc.x, mc.y - hc.y, mc.z - hc.z}, cv[3] = {mv.x - hv.x, mv.y - hv.y, mv.z - hv.z};
			for (int i = 0; i < 3; i++) v1[i] = (v1[i] - cp[i]) * 64 / spring;
			bt_bound_vec(v1, a_max, 3);
			for (int i = 0; i < 3; i++) vel[i] -= cv[i];
			bt_bound_vec(vel, drag, 3);
		}
		i32 a[3];
		for (int i = 0; i < 3; i++) a[i] = (v1[i] + vel[i]) / -ratio_legs;
		const i32 r = W.holdRoot[me];
		i32 *dr = (i32 *)&W.d_vel[r], *dm = (i32 *)&W.d_vel[me];
		for (int i = 0; i < 3; i++) bt_add32((u32 *)&dr[i], (u32)a[i]);              /* uVel(me->holdRoot, a) */
		for (int i = 0; i < 3; i++) bt_add32((u32 *)&dm[i], (u32)(a[i] * -ratio_legs)); /* uVel(me, a) */
		W.ctr[CT_ANY_DELTA] = 1;
		return;
	}
	case TICK_LOADBLOCK: { /* loadblock_tick loadblock.cpp:33-36 */
		i32 cooldown = bt_getslider(W, me, 1);
		if (cooldown > 0) bt_uslider_shared(W, me, 1, cooldown - 1);
		return;
	}
	case TICK_PUDDLEJUMP: { /* pj_tick puddlejumper.cpp:94-162.  The pilot's controls arrive through the wired seat as
	                         * sliders (getAxis = sliders 0 / 1, player.c:48-51) and buttons (getTrigger(e, i) = button i + 2,
	                         * entGetters.c:27-29).  Vertical thrust makes dust (mkDust creates entities): host side */
		i32 dvel[3] = {0, 0, 16};
		const int flying = !bt_getslider(W, me, 4);
		bt_uslider_shared(W, me, 4, 0);
		const i32 fuel = bt_getslider(W, me, 0);
		if (fuel && flying) bt_uslider_shared(W, me, 0, fuel - 1);
		if (W.wire_cnt[me] == 1) {
			const i32 input = W.wire_to[W.wire_off[me]];
			if (fuel && bt_button(W, input, 2)) { bt_unsup(W, me, 6, id); dvel[2] = -8; }
			const i32 axis[2] = {bt_getslider(W, input, 0), bt_getslider(W, input, 1)};
			if (axis[0] || axis[1]) {
				const i32 sl[2] = {bt_getslider(W, me, 2), bt_getslider(W, me, 3)};
				const int signX = 2 * (axis[0] >= 0) - 1, signY = 2 * (axis[1] >= 0) - 1;
				i32 denomX = signX * axis[0], numerX = 200 - signX * sl[0]; /* PJ_SPD */
				i32 denomY = signY * axis[1], numerY = 200 - signY * sl[1];
				if (numerX > 20) numerX = 20; else if (numerX < 0) numerX = 0; /* PJ_ACCEL */
				if (numerY > 20) numerY = 20; else if (numerY < 0) numerY = 0;
				if (denomY && numerY * denomX <= numerX * denomY) { numerX = numerY; denomX = denomY; }
				for (int k = 0; k < 2; k++) {
					dvel[k] = numerX * axis[k] / denomX;
					bt_uslider_shared(W, me, 2 + k, sl[k] + dvel[k]);
					dvel[k] += axis[k] / 16;
				}
				bt_uslider_shared(W, me, 5, numerX ? 2 : 1);
			} else {
				bt_uslider_shared(W, me, 5, 0);
			}
		}
		int4 dv = W.d_vel[me]; /* uVel(me, dvel) */
		dv.x += dvel[0]; dv.y += dvel[1]; dv.z += dvel[2];
		W.d_vel[me] = dv;
		W.ctr[CT_ANY_DELTA] = 1;
		return;
	}
	case TICK_HELD_GUN: case TICK_HELD_BLINK: case TICK_HELD_JUMPER: {
		/* gun_tick_held gun.c:41-88, blink_tick_held blink.c:33-60, jumper_tick_held jumper.c:39-70: "doesn't do anything
		 * unless held by a person (or similar)".  Held by one they count their cooldown down and reload; pulling the
		 * holder's trigger (getTrigger entGetters.c:27-29 = button 2 + hand) creates a bullet / a dust cloud: the host's */
		const i32 h = W.holder[me];
		if (h < 0 || !(W.type[h] & T_INPUTS) || W.holder[h] >= 0) return;
		const int trigger = bt_button(W, h, 2 + (bt_getslider(W, me, 0) != 0));
		if (id == TICK_HELD_GUN) { /* sliders: hand, ammo, cooldown, reload, ammo max, cooldown max, reload max, speed */
			const i32 cooldown = bt_getslider(W, me, 2), ammo = bt_getslider(W, me, 1);
			if (cooldown > 1) bt_uslider(W, me, 2, cooldown - 1);
			else if (trigger && ammo) { bt_unsup(W, me, 6, id); return; }
			const i32 ammo_max = bt_getslider(W, me, 4);
			if (ammo != ammo_max) {
				const i32 reload = bt_getslider(W, me, 3) + 1;
				if (reload >= bt_getslider(W, me, 6)) { bt_uslider(W, me, 3, 0); bt_uslider(W, me, 1, ammo_max); }
				else bt_uslider(W, me, 3, reload);
			}
		} else if (id == TICK_HELD_BLINK) { /* hand, cooldown, cooldown max, range */
			const i32 cooldown = bt_getslider(W, me, 1);
			if (cooldown > 0) bt_uslider(W, me, 1, cooldown - 1);
			else if (trigger) bt_unsup(W, me, 6, id);
		} else { /* jumper: hand, ammo, reload, cooldown, ammo max, reload max, cooldown max, speed */
			i32 ammo = bt_getslider(W, me, 1);
			if (ammo < bt_getslider(W, me, 4)) {
				i32 reload = 1 + bt_getslider(W, me, 2);
				if (reload >= bt_getslider(W, me, 5)) { reload = 0; ammo++; }
				bt_uslider(W, me, 2, reload);
			}
			const i32 cooldown = bt_getslider(W, me, 3);
			if (cooldown > 0) bt_uslider(W, me, 3, cooldown - 1);
			else if (ammo && trigger) { bt_unsup(W, me, 6, id); return; }
			bt_uslider(W, me, 1, ammo);
		}
		return;
	}
	case TICK_SCOREBOARD: case TICK_INCINERATE: case TICK_FACTORY: case TICK_FACTORY_WIRES: return; /* per world, in list order (t_tick_serial) */
	case TICK_BOTTLE: return; /* bottle_tick bottle.cpp:42-43 is empty */
	case TICK_EXPLOSION: { /* explosion_tick explosion.cpp:12-19 */
		const i32 timer = bt_getslider(W, me, 0);
		if (timer < 3) bt_uslider_shared(W, me, 0, timer + 1);
		else bt_udead(W, me);
		return;
	}
	case TICK_PLATFORM: /* platform_tick platform.c:15-23: surviving its first tick clears the special behaviour (direct writes) */
		W.htick[me] = ht & ~0xffu;
		W.hphys[me] &= ~(0xffu << 24);
		for (i32 k = 0; k < W.wire_cnt[me]; k++) bt_uunwire(W, me, W.wire_to[W.wire_off[me] + k]);
		return;
	case TICK_THRUST: { /* thrust_tick thrust.c:19-37: pushed away from its lit wires; several thrusters share a body */
		if (!W.wire_cnt[me] || bt_button(W, me, 0)) return;
		const i32 divisor = bt_getslider(W, me, 0);
		if (!divisor) return;
		i32 off[3] = {0, 0, 0};
		const int4 mc = W.center[me];
		for (i32 k = 0; k < W.wire_cnt[me]; k++) {
			const i32 w = W.wire_to[W.wire_off[me] + k];
			if (!bt_button(W, w, 2)) continue;
			const int4 wc = W.center[w];
			off[0] += mc.x - wc.x; off[1] += mc.y - wc.y; off[2] += mc.z - wc.z;
		}
		i32 *dr = (i32 *)&W.d_vel[W.holdRoot[me]];
		for (int i = 0; i < 3; i++) bt_add32((u32 *)&dr[i], (u32)(off[i] * 64 / divisor)); /* uVel(me->holdRoot, offset) */
		W.ctr[CT_ANY_DELTA] = 1;
		return;
	}
	case TICK_BULLET: { /* bullet_tick gun.c:116-123 */
		const i32 ttl = bt_getslider(W, me, 0);
		if (ttl <= 0) bt_udead(W, me);
		else bt_uslider_shared(W, me, 0, ttl - 1);
		return;
	}
	case TICK_HELD_BULLET: { /* bullet_tick_held gun.c:125-152: stuck in something; enough of them kill it */
		const i32 ttl = bt_getslider(W, me, 0);
		if (ttl <= -15 || ttl > 0) {
			const i32 h = W.holder[me];
			const u32 t = W.type[h];
			int hits = (t & T_DEBRIS) ? 1 : ((t & T_TERRAIN) ? ((t & T_HEAVY) ? 15 : 5) : 3);
			for (i32 c = W.holdee[h]; c >= 0; c = W.LLn[c]) {
				if (((W.htick[c] >> 8) & 0xffu) != TICK_HELD_BULLET) continue;
				if (!--hits) { bt_udead(W, h); break; }
			}
			bt_udead(W, me);
		} else {
			bt_uslider_shared(W, me, 0, ttl - 1); /* (+ a colour: not simulation state) */
		}
		return;
	}
	case TICK_DUST: { /* dust_tick dust.c:9-21 */
		const i32 time = bt_getslider(W, me, 3);
		if (time <= -8) { bt_udead(W, me); return; }
		if (time > 0) {
			int4 dv = W.d_vel[me]; /* uVel(me, sliders 0..2) */
			dv.x += bt_getslider(W, me, 0); dv.y += bt_getslider(W, me, 1); dv.z += bt_getslider(W, me, 2);
			W.d_vel[me] = dv;
			W.ctr[CT_ANY_DELTA] = 1;
		}
		bt_uslider_shared(W, me, 3, time - 1);
		return;
	}
	case TICK_BOTTLE_HELD: { /* bottle_tick_held bottle.cpp:47-90: "doesn't do anything unless held by a person (or similar)" */
		const i32 h = W.holder[me];
		if (h < 0 || !(W.type[h] & T_INPUTS)) return;
		bt_unsup(W, me, 6, id);
		return;
	}
	default: bt_unsup(W, me, 6, id); return;
	}
}

/* stepRand random.cpp:7-9 */
static BT_HD i32 bt_random(const DevWorld &W, int g) {
	i32 x = (i32)(((long long)W.rnd[g] * 48271) % 2147483647);
	W.rnd[g] = x;
	return x;
}
/* randomazzo_tick logic.c:151-168 for every TICK_RAND entity of world g, newest first (gs->ents order), and the
 * scoreboards */
static BT_HD void t_tick_serial(const DevWorld &W, int g) {
	if (g >= W.n_worlds) return;
	for (i32 me = W.wend[g] - 1; me >= W.wbegin[g]; me--) {
		if (!W.inList[me]) continue;
		u32 ht = W.htick[me];
		const u32 id = W.holder[me] >= 0 ? (ht >> 8) & 0xffu : ht & 0xffu;
		if (id == TICK_SCOREBOARD) { bt_scoreboard_tick(W, me); continue; }
		if (id == TICK_FACTORY) { /* factory_tick factory.cpp:63-73.  spawn() creates an entity: host side, counted.  What
		                           * follows a spawn -- finalize (:30-61) of the wired product -- is all deferred updaters and
		                           * writes to the product's own handler slots */
			if (bt_button(W, me, 0) || bt_button(W, me, 1)) { bt_unsup(W, me, 6, (int)id); continue; }
			if (!bt_button(W, me, 2) || !W.wire_cnt[me]) continue;
			if (W.wire_cnt[me] == 1) {
				const i32 e = W.wire_to[W.wire_off[me]];
				/* a product is newer than its factory (spawn() prepends it to gs->ents), so in the reference's newest-first loop
				 * it has had this tick's turn with its old handler when the factory rewrites the slot -- what the serial pass
				 * behind the parallel one amounts to.  An OLDER entity wired in by hand would tick again with the new one */
				if (e < me) bt_unsup(W, me, 6, (int)id);
				int4 dv = W.d_vel[e]; /* uVel(e, sliders FACT_VEL..) */
				dv.x += bt_getslider(W, me, 1); dv.y += bt_getslider(W, me, 2); dv.z += bt_getslider(W, me, 3);
				W.d_vel[e] = dv;
				W.ctr[CT_ANY_DELTA] = 1;
				const u32 misc = (u32)bt_getslider(W, me, 7);
				if (misc & 0x80000000u) { /* FACT_MISC_HOLD: whoever the factory's holder is wired to picks the product up */
					const i32 h = W.holder[me];
					if (h >= 0 && W.wire_cnt[h] == 1) bt_upickup(W, W.wire_to[W.wire_off[h]], e, W.holdFlags[me]);
				}
				for (int i = 0; i < 4; i++)
					if (misc & (1u << i)) bt_push_btn(W, e, i);
				W.fullFlush[e] = 1; /* uMyTypeMask / uMyCollideMask */
				W.ntype[e] = (u32)bt_getslider(W, me, 4);
				W.ncollide[e] = (u32)bt_getslider(W, me, 5);
				W.ctr[CT_FLUSH_HINT] = 1;
				W.htick[e] = ((u32)bt_getslider(W, me, 9) & 0xffu) | (((u32)bt_getslider(W, me, 10) & 0xffu) << 8);
				W.hphys[e] = (W.hphys[e] & 0xffu) | (((u32)bt_getslider(W, me, 12) & 0xffu) << 8) | (((u32)bt_getslider(W, me, 13) & 0xffu) << 16) |
				             (((u32)bt_getslider(W, me, 11) & 0xffu) << 24);
				W.hpair[e] = ((u32)bt_getslider(W, me, 14) & 0xffu) | (((u32)bt_getslider(W, me, 15) & 0xffu) << 8) |
				             (((u32)bt_getslider(W, me, 16) & 0xffu) << 16) | (((u32)bt_getslider(W, me, 17) & 0xffu) << 24);
			}
			for (i32 k = 0; k < W.wire_cnt[me]; k++) bt_uunwire(W, me, W.wire_to[W.wire_off[me] + k]);
			continue;
		}
		if (id == TICK_FACTORY_WIRES) { /* factory_wires_tick factory.cpp:75-91 */
			if (bt_button(W, me, 0) || bt_button(W, me, 1)) bt_push_btn(W, me, 2);
			if (!bt_button(W, me, 2)) continue;
			const i32 h = W.holder[me];
			if (h < 0 || W.wire_cnt[h] != 1) continue;
			const i32 e = W.wire_to[W.wire_off[h]];
			for (i32 k = 0; k < W.wire_cnt[me]; k++) {
				const i32 w = W.wire_to[W.wire_off[me] + k];
				for (i32 j = 0; j < W.wire_cnt[w]; j++) bt_uwire(W, e, W.wire_to[W.wire_off[w] + j]);
			}
			continue;
		}
		if (id == TICK_INCINERATE) { /* incin_tick incinerate.cpp:15-32: everything fully beyond `range` dies */
			if (bt_button(W, me, 0) ^ bt_button(W, me, 1)) continue;
			const i32 range = bt_getslider(W, me, 0);
			if (range <= 0) continue;
			const int4 mc = W.center[me];
			for (i32 e = W.wbegin[g]; e < W.wend[g]; e++) {
				if (!W.inList[e]) continue;
				const int4 c = W.center[e], r = W.radius[e];
				if (bt_abs(c.x - mc.x) - r.x >= range || bt_abs(c.y - mc.y) - r.y >= range || bt_abs(c.z - mc.z) - r.z >= range) bt_udead(W, e);
			}
			continue;
		}
		if (id != TICK_RAND) continue;
		if (!bt_logic_input(W, me, bt_getslider(W, me, 0))) continue;
		int options = 0;
		for (i32 k = 0; k < W.wire_cnt[me]; k++) options += !bt_button(W, W.wire_to[W.wire_off[me] + k], 0);
		if (!options) continue;
		int roll = bt_random(W, g) % options;
		for (i32 k = 0; k < W.wire_cnt[me]; k++) {
			i32 w = W.wire_to[W.wire_off[me] + k];
			if (bt_button(W, w, 0)) continue;
			if (roll--) continue;
			i32 outputs = bt_getslider(W, me, 1);
			if (outputs & 1) bt_push_btn(W, w, 0);
			if (outputs & 2) bt_push_btn(W, w, 1);
			bt_push_btn(W, me, 2);
			break;
		}
	}
}

/* The deferred updaters of entUpdaters.c called by HOST-side code (tick handlers, doInputs,
 * edit commands) between two steps.  Applied by one thread in the order given, like the host
 * would have called them. */
static BT_HD void t_request(const DevWorld &W, const BtReq &q) {
	i32 e = q.me;
	if (e < 0 || e >= W.n || !W.inList[e]) return;
	int g = W.world[e];
	switch (q.kind) {
	case BT_REQ_DROP: W.newDrop[e] = 1; break;                                  /* uDrop :28 */
	case BT_REQ_DEAD: if (W.flip[g]) W.dead_max1[e] |= 1; else W.dead_max0[e] |= 1; break; /* uDead :20 */
	case BT_REQ_ERASE: if (W.flip[g]) W.dead_max1[e] = 3; else W.dead_max0[e] = 3; break;  /* uErase :24 */
	case BT_REQ_PICKUP:                                                         /* uPickup :32 */
		if (q.other >= 0 && q.other < W.n && W.inList[q.other]) bt_upickup(W, e, q.other, q.a0);
		break;
	case BT_REQ_BUTTON: if (q.a0 >= 0 && q.a0 < 4) W.btn_v2[e] |= 1u << (8 * q.a0); break; /* pushBtn :6 */
	case BT_REQ_VEL: { int4 v = W.d_vel[e]; v.x += q.a0; v.y += q.a1; v.z += q.a2; W.d_vel[e] = v; W.ctr[CT_ANY_DELTA] = 1; break; }       /* uVel :14 */
	case BT_REQ_CENTER: { int4 v = W.d_center[e]; v.x += q.a0; v.y += q.a1; v.z += q.a2; W.d_center[e] = v; W.ctr[CT_ANY_DELTA] = 1; break; } /* uCenter :8 */
	case BT_REQ_SLIDER: bt_uslider(W, e, q.a0, q.a1); break;                    /* uSlider :54 */
	case BT_REQ_TYPEMASK: W.fullFlush[e] = 1; W.ntype[e] = (u32)q.a0; break;    /* uMyTypeMask :75 */
	case BT_REQ_COLLIDEMASK: W.fullFlush[e] = 1; W.ncollide[e] = (u32)q.a0; break; /* uMyCollideMask :80 */
	default: W.ctr[CT_ERROR] = (u32)(-BT_ERR_ARG); break;
	}
	if (q.kind == BT_REQ_DROP || q.kind == BT_REQ_DEAD || q.kind == BT_REQ_ERASE) W.wflush[g] = 1;
	W.ctr[CT_FLUSH_HINT] = 1;
}

/* ------------------------------------------------------------------------- */
/* flushes, gravity, integration                                             */

/* flushCtrls ent.cpp:32-37 + detection of work for the serial part of flush1 */
static BT_HD void t_pre(const DevWorld &W, i32 e, u32 phases) {
	if (e >= W.n || !W.inList[e]) return;
	if (phases & BT_PHASE_CTRLS) {
		W.btn_v[e] = W.btn_v2[e];
		W.btn_v2[e] = 0;
	}
}

/* marks worlds whose flush1 has deaths / drops / pickups to apply.  Only needed for what an upload
 * brought along (the first two flush1 after it: one per dead_max slot); everything requested on
 * the device marks its world itself (bt_udead, bt_upickup, uDrop, world rules, host requests). */
static BT_HD void t_flush_scan(const DevWorld &W, i32 e) {
	if (e >= W.n || !W.inList[e]) return;
	if (W.dead_max0[e] | W.dead_max1[e] | W.newDrop[e] | (W.newHolder[e] >= 0)) {
		W.wflush[W.world[e]] = 1;
		W.ctr[CT_FLUSH_WORK] = 1;
	}
}

static BT_HD void bt_flush_pickup_one(const DevWorld &W, i32 i) {
	i32 x = W.newHolder[i];
	if (W.holdRoot[x] != i && !W.dead[x]) bt_pickup(W, x, i, W.newHoldFlags[i]);
	W.newHolder[i] = -1;
}

/* flush1 ent.cpp:794-802, serial part, one thread per world: flushDeaths :747-761,
 * flushDrops :765-773, flushPickups :775-792.  Worlds with nothing pending (wflush) return at once. */
static BT_HD void t_flush_serial(const DevWorld &W, i32 g, LocalStats &st) {
	if (g >= W.n_worlds) return;
	W.flip[g] ^= 1;
	if (!W.wflush[g]) return;
	W.wflush[g] = 0;
	W.wrootseq[g] = 0;
	st.rk_ordered = 1;
	int slot = 1 ^ W.flip[g];
	for (i32 i = W.wend[g] - 1; i >= W.wbegin[g]; i--) {
		if (!W.inList[i]) continue;
		u8 d = slot ? W.dead_max1[i] : W.dead_max0[i];
		if (d) {
			bt_crush(W, i, 0xfffffu - ((u32)i & 0xfffffu), st, d); /* before every push of the tick, newest first */
			W.dead[i] = d;
		}
	}
	for (i32 i = W.wend[g] - 1; i >= W.wbegin[g]; i--) {
		if (!W.inList[i]) continue;
		if (W.newDrop[i]) {
			W.newDrop[i] = 0;
			bt_fumble(W, i, st);
		}
	}
	/* flushPickups ent.cpp:775-792 walks gs->rootEnts, head first = descending rootKey.  Each root with a
	 * request is visited once (its request is cleared), nothing visited adds a root or a request. */
	while (1) {
		i32 best = -1;
		u64 bk = 0;
		for (i32 i = W.wbegin[g]; i < W.wend[g]; i++) {
			if (!W.inList[i] || W.holder[i] >= 0 || W.newHolder[i] < 0) continue;
			if (best < 0 || W.rootKey[i] > bk) { best = i; bk = W.rootKey[i]; }
		}
		if (best < 0) break;
		bt_flush_pickup_one(W, best);
	}
}

/* flush1 with nothing to do still toggles the death flip-flop (ent.cpp:795) */
static BT_HD void t_flip_only(const DevWorld &W, i32 g) {
	if (g < W.n_worlds) W.flip[g] ^= 1;
}

/* flushMisc ent.cpp:674-714 */
static BT_HD void t_flush_misc(const DevWorld &W, i32 e) {
	if (e >= W.n || !W.inList[e]) return;
	if (W.fullFlush[e]) {
		W.fullFlush[e] = 0;
		for (int s = W.slider_off[e]; s < W.slider_off[e + 1]; s++) {
			if (W.smax[s] < W.smin[s]) continue;
			W.sv[s] = (i32)(((long long)W.smax[s] + W.smin[s]) / 2);
			W.smax[s] = INT32_MIN;
			W.smin[s] = INT32_MAX;
		}
		if (W.ntype[e] != W.type[e] || W.collide[e] != W.ncollide[e]) {
			W.type[e] = W.ntype[e];
			W.collide[e] = W.ncollide[e];
			u8 nb = (u8)bt_needs_box(W.type[e], W.collide[e]); /* fixHasVelbox ent.cpp:79-89 */
			if (nb && !W.hasBox[e]) W.ctr[CT_NEED_REBUILD] = 1;
			W.hasBox[e] = nb;
		}
		const i32 addw = W.wireAdd[e];
		if (addw != -1) { /* wiresAdd ent.cpp:694-698, in request order */
			W.wireAdd[e] = -1;
			const i32 b0 = W.wire_off[e];
			BT_WQ_FOR(W, addw, add) {
				i32 c0 = W.wire_cnt[e];
				int has = 0;
				for (i32 i = 0; i < c0; i++) has |= W.wire_to[b0 + i] == add;
				if (has) continue;
				if (c0 < W.wire_cap[e]) { W.wire_to[b0 + c0] = add; W.wire_cnt[e] = c0 + 1; }
				else W.ctr[CT_ERROR] = (u32)(-BT_ERR_CAPACITY);
			}
		}
	}
	const i32 rmw = W.wireRm[e];
	if (rmw != -1) W.wireRm[e] = -1;
	const i32 cnt0 = W.wire_cnt[e];
	if (!cnt0) return; /* most entities have no wires: nothing to clean, nothing to write */
	i32 base = W.wire_off[e];
	i32 cnt = cnt0;
	for (i32 i = 0; i < cnt; i++) {
		i32 x = W.wire_to[base + i];
		int gone = W.dead[x];
		if (!gone && rmw != -1) BT_WQ_FOR(W, rmw, rm) gone |= rm == x;
		if (gone) { /* quickRmAt: pull the last wire into this place (list.h) */
			W.wire_to[base + i] = W.wire_to[base + (--cnt)];
			i--;
		}
	}
	if (cnt != cnt0) W.wire_cnt[e] = cnt;
}

/* world rules that are per-entity scans: crush() effects.cpp:48-60 (doCrushtainer 64000,
 * doBigCrushtainer 2560000), doLava :43-46, bounce() :71-92.  uDead = dead_max[flip] |= 1
 * (entUpdaters.c:21-23), picked up by the next flush1. */
#define EFFECT_CRUSH 1
#define EFFECT_CRUSH_BIG 2
#define EFFECT_LAVA 16
#define EFFECT_BOUNCE 32
static BT_HD int bt_outside(const int4 &c, i32 r) { return bt_abs(c.x) > r * 2 || bt_abs(c.y) > r || bt_abs(c.z) > r; }
static BT_HD void t_world_rules(const DevWorld &W, i32 e) {
	if (e >= W.n || !W.inList[e]) return;
	int g = W.world[e];
	i32 rules = W.gamerules[g];
	if (!(rules & (EFFECT_CRUSH | EFFECT_CRUSH_BIG | EFFECT_LAVA | EFFECT_BOUNCE))) return;
	int4 c = W.center[e];
	int die = 0;
	if ((rules & EFFECT_CRUSH) && bt_outside(c, 64000)) die = 1;
	if ((rules & EFFECT_CRUSH_BIG) && bt_outside(c, 2560000)) die = 1;
	if ((rules & EFFECT_LAVA) && c.z > 64000) die = 1;
	if ((rules & EFFECT_BOUNCE) && bt_outside(c, 64000)) {
		if (W.type[e] & T_TERRAIN) {
			int4 v = W.vel[e];
			if ((c.x > 128000 && v.x > 0) || (c.x < -128000 && v.x < 0)) v.x = -v.x;
			if ((c.y > 64000 && v.y > 0) || (c.y < -64000 && v.y < 0)) v.y = -v.y;
			if ((c.z > 64000 && v.z > 0) || (c.z < -64000 && v.z < 0)) v.z = -v.z;
			W.vel[e] = v;
		} else {
			die = 1;
		}
	}
	if (die) {
		if (W.flip[g]) W.dead_max1[e] |= 1;
		else W.dead_max0[e] |= 1;
		W.wflush[g] = 1;
		W.ctr[CT_FLUSH_HINT] = 1;
	}
}

/* doGravity effects.cpp:35-41 */
static BT_HD void t_gravity(const DevWorld &W, i32 e) {
	if (e >= W.n || !W.inList[e] || W.holder[e] >= 0) return;
	if (!(W.gamerules[W.world[e]] & EFFECT_GRAV)) return;
	if (W.type[e] & T_WEIGHTLESS) return;
	W.d_vel[e].z += 32; /* consumed by the flush2 at the start of doPhysics (host-tracked) */
}

/* flush2 / flushPosVel ent.cpp:716-744, 806-811 as a gather: every entity sums the pending
 * deltas of itself and its holders (they are only zeroed by the following kernel) */
static BT_HD void t_flush2_apply(const DevWorld &W, i32 e) {
	if (e >= W.n || !W.inList[e]) return;
	i32 dc[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
	for (i32 x = e; x >= 0; x = W.holder[x]) {
		int4 a = W.d_center[x], b = W.d_vel[x];
		dc[0] += a.x; dc[1] += a.y; dc[2] += a.z;
		dv[0] += b.x; dv[1] += b.y; dv[2] += b.z;
	}
	if ((dc[0] | dc[1] | dc[2]) != 0) {
		int4 c = W.center[e], o = W.old[e];
		c.x += dc[0]; c.y += dc[1]; c.z += dc[2];
		o.x += dc[0]; o.y += dc[1]; o.z += dc[2];
		W.center[e] = c;
		W.old[e] = o;
	}
	if ((dv[0] | dv[1] | dv[2]) != 0) {
		int4 v = W.vel[e];
		v.x += dv[0]; v.y += dv[1]; v.z += dv[2];
		W.vel[e] = v;
	}
}

static BT_HD void t_zero_deltas(const DevWorld &W, i32 e) {
	if (e >= W.n) return;
	W.d_center[e] = make_int4(0, 0, 0, 0);
	W.d_vel[e] = make_int4(0, 0, 0, 0);
}

/* integrate loop ent.cpp:845-858 (+ zeroing of the deltas flush2 just applied) */
static BT_HD void t_integrate(const DevWorld &W, i32 e) {
	if (e >= W.n) return;
	W.d_center[e] = make_int4(0, 0, 0, 0);
	W.d_vel[e] = make_int4(0, 0, 0, 0);
	if (!W.inList[e]) return;
	int4 c = W.center[e], v = W.vel[e];
	W.old[e] = c;
	int4 o = c;
	c.x += v.x; c.y += v.y; c.z += v.z;
	W.center[e] = c;
	W.forced[e] = 0;
	W.needsCollision[e] = 1;
	if (W.hasBox[e]) bt_check_cover(W, e, o, c);
}

/* Fast path of flush2 + doGravity + integrate when no delta is pending anywhere except the
 * gravity this very kernel stands for: nothing is read from or written to d_center / d_vel.
 * Gravity reaches every member of a tree whose root is not weightless, exactly what
 * uVel(root) + flushPosVel do.  (ent.cpp:806-811, effects.cpp:35-41, ent.cpp:845-858) */
static BT_HD void t_begin_fused(const DevWorld &W, i32 e, int gravity, int ctrls) {
	if (e >= W.n || !W.inList[e]) return;
	if (ctrls) { /* flushCtrls ent.cpp:32-37; nothing on the device reads buttons before this point */
		u32 v2 = W.btn_v2[e];
		if (W.btn_v[e] | v2) { W.btn_v[e] = v2; W.btn_v2[e] = 0; }
	}
	int4 c = W.center[e], v = W.vel[e];
	if (gravity) {
		i32 r = W.holdRoot[e];
		if ((W.gamerules[W.world[e]] & EFFECT_GRAV) && !(W.type[r] & T_WEIGHTLESS)) {
			v.z += 32;
			W.vel[e] = v;
		}
	}
	W.old[e] = c;
	int4 o = c;
	c.x += v.x; c.y += v.y; c.z += v.z;
	W.center[e] = c;
	W.forced[e] = 0;
	W.needsCollision[e] = 1;
	/* pairs cached in an earlier tick stay valid while nobody leaves its inflated box */
	if (W.hasBox[e]) bt_check_cover(W, e, o, c);
}

/* state hash (DESIGN.md "State hash"), summed per world with 64-bit atomics */
static BT_HD u64 bt_ent_hash(const DevWorld &W, i32 e) {
	int4 c = W.center[e], v = W.vel[e];
	i32 hd = W.holder[e];
	/* A holder that died this tick after its holdees were fumbled can still pick something up in a
	 * later push of the same iteration (the reference runs every decided push, ent.cpp:879-882, and
	 * leaves a dangling e->holder behind once clearDeads frees it): hashed as "no holder", which is
	 * also what an export / download reports. */
	i32 wd[10] = {c.x, c.y, c.z, v.x, v.y, v.z, (i32)W.type[e], (i32)W.collide[e],
	              hd >= 0 && W.inList[hd] ? W.uid[hd] : -1, W.holdFlags[e]};
	u64 h = bt_mix64((u64)(u32)W.uid[e] + 0x9E3779B97F4A7C15ULL);
	for (int k = 0; k < 10; k++) h = bt_mix64(h + (u64)(u32)wd[k] + 0x9E3779B97F4A7C15ULL);
	return h;
}


/* ------------------------------------------------------------------------- */
/* incremental host binding (bt_world_poke): sparse writes of what HOST code changed on its own
 * `ent`s since the last exchange. */
struct DevArrays { /* BT_A_* -> base pointer, bytes per element, ints written by a poke */
	void *base[BT_A_NUM];
};
#define BT_OP_REBOX 99 /* internal: a direct typeMask / collideMask write landed on this slot, the box follows (ent.cpp:75-89) */
/* step 1, one thread per poke: the raw array writes (no two pokes name the same element) */
static BT_HD void t_poke_raw(const DevWorld &W, const DevArrays &A, const bt_poke *pk, i32 t, i32 n) {
	if (t >= n) return;
	const bt_poke p = pk[t];
	const i32 e = p.index;
	if (p.array < 0 || p.array >= BT_A_NUM) return; /* an operation: step 2 */
	switch (p.array) {
	case BT_A_CENTER: case BT_A_VEL: case BT_A_OLD: case BT_A_DCENTER: case BT_A_DVEL: case BT_A_RADIUS:
		((int4 *)A.base[p.array])[e] = make_int4(p.v[0], p.v[1], p.v[2], 0);
		if (p.array == BT_A_DCENTER || p.array == BT_A_DVEL) W.ctr[CT_ANY_DELTA] = 1;
		break;
	case BT_A_NEWDROP: case BT_A_DEADMAX0: case BT_A_DEADMAX1: case BT_A_FULLFLUSH: case BT_A_DEAD: case BT_A_INLIST:
		((u8 *)A.base[p.array])[e] = (u8)p.v[0];
		W.ctr[CT_FLUSH_HINT] = 1;
		break;
	case BT_A_ROOTKEY:
		W.rootKey[e] = (u64)(u32)p.v[0] | ((u64)(u32)p.v[1] << 32);
		break;
	case BT_A_DEADLIST:
		W.ctr[CT_ERROR] = (u32)(-BT_ERR_ARG);
		break;
	default: /* every other array has 4-byte elements */
		((i32 *)A.base[p.array])[e] = p.v[0];
		if (p.array == BT_A_SMIN || p.array == BT_A_SMAX || p.array == BT_A_NEWHOLDER) W.ctr[CT_FLUSH_HINT] = 1;
		break;
	}
}
/* step 2, one thread, in the order given: the operations */
static BT_HD void t_poke_op(const DevWorld &W, const bt_poke &p) {
	const i32 e = p.index;
	switch (p.array) {
	case BT_OP_SPAWN: /* addEnt ent.cpp:107-113 */
		W.inList[e] = 1;
		W.dead[e] = 0;
		W.hasBox[e] = (u8)bt_needs_box(W.type[e], W.collide[e]);
		if (W.hasBox[e]) W.ctr[CT_NEED_REBUILD] = 1;
		W.needsCollision[e] = 0;
		W.pend[e] = 0; W.ready[e] = 0; W.forced[e] = 0;
		W.forcedHoldees[e] = make_int4(-1, -1, -1, 0);
		bt_add32(&W.ctr[CT_TREE_EDITS], 1u);
		break;
	case BT_OP_REBOX:
		if (W.inList[e]) {
			u8 nb = (u8)bt_needs_box(W.type[e], W.collide[e]);
			if (nb && !W.hasBox[e]) W.ctr[CT_NEED_REBUILD] = 1;
			W.hasBox[e] = nb;
		}
		break;
	case BT_OP_UWIRE: bt_uwire(W, e, p.v[0]); break;
	case BT_OP_UUNWIRE: bt_uunwire(W, e, p.v[0]); break;
	default: W.ctr[CT_ERROR] = (u32)(-BT_ERR_ARG); break;
	}
}

/* bt_world_grow: one blank slot (out of the world until BT_OP_SPAWN) */
static BT_HD void t_grow_init(const DevWorld &W, i32 e, i32 world) {
	const int4 z = make_int4(0, 0, 0, 0);
	W.world[e] = world; W.uid[e] = -1;
	W.center[e] = z; W.old[e] = z; W.vel[e] = z; W.radius[e] = z; W.d_center[e] = z; W.d_vel[e] = z;
	W.type[e] = 0; W.collide[e] = 0; W.ntype[e] = 0; W.ncollide[e] = 0;
	W.friction[e] = 0; W.holder[e] = -1; W.holdFlags[e] = 0; W.holdRoot[e] = e; W.holdee[e] = -1;
	W.LLn[e] = -1; W.LLp[e] = -1; W.newHolder[e] = -1; W.newHoldFlags[e] = 0;
	W.hphys[e] = 0; W.hpair[e] = 0; W.htick[e] = 0; W.btn_v[e] = 0; W.btn_v2[e] = 0;
	W.newDrop[e] = 0; W.dead_max0[e] = 0; W.dead_max1[e] = 0; W.fullFlush[e] = 0; W.dead[e] = 0; W.inList[e] = 0;
	W.needsCollision[e] = 0; W.hasBox[e] = 0; W.ready[e] = 0; W.forced[e] = 0;
	W.forcedHoldees[e] = make_int4(-1, -1, -1, 0); W.pend[e] = 0;
	W.wire_cnt[e] = 0; W.wireAdd[e] = -1; W.wireRm[e] = -1; W.rootKey[e] = 0;
	W.dirtyStamp[e] = 0; W.nodeCount[e] = 0; W.leafOf[e] = 0;
	Box6 b;
	b.lo = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, 0); b.hi = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, 0);
	W.ibox[e] = b; W.iboxOld[e] = b;
}

/* bt_world_hash_full (include/bittoss_b200.h): everything serialize() writes (serialize.cpp:119-164) plus
 * the flush state a tick leaves behind.  Restated by oracle/ref_harness.cpp (btref_hash_full) and the host
 * binding (host/bt_glue.cpp) over `ent`s. */
static BT_HD u64 bt_hash_word(u64 h, i32 w) { return bt_mix64(h + (u64)(u32)w + 0x9E3779B97F4A7C15ULL); }
static BT_HD i32 bt_uid_or_none(const DevWorld &W, i32 e) { return e >= 0 && W.inList[e] ? W.uid[e] : -1; }
static BT_HD u64 bt_ent_hash_full(const DevWorld &W, i32 e) {
	u64 h = bt_ent_hash(W, e);
	const int4 r = W.radius[e], o = W.old[e], fh = W.forcedHoldees[e];
	(void)fh;
	h = bt_hash_word(h, r.x); h = bt_hash_word(h, r.y); h = bt_hash_word(h, r.z);
	h = bt_hash_word(h, o.x); h = bt_hash_word(h, o.y); h = bt_hash_word(h, o.z);
	h = bt_hash_word(h, W.friction[e]);
	h = bt_hash_word(h, (i32)W.ntype[e]); h = bt_hash_word(h, (i32)W.ncollide[e]);
	h = bt_hash_word(h, (i32)W.hphys[e]); h = bt_hash_word(h, (i32)W.hpair[e]); h = bt_hash_word(h, (i32)W.htick[e]);
	h = bt_hash_word(h, (i32)W.btn_v[e]); h = bt_hash_word(h, (i32)W.btn_v2[e]);
	{
		i32 nh = W.newHolder[e];
		h = bt_hash_word(h, nh == e ? -2 : bt_uid_or_none(W, nh)); /* newHolder == itself: contested request (entUpdaters.c:51) */
		h = bt_hash_word(h, nh >= 0 ? W.newHoldFlags[e] : 0);
	}
	h = bt_hash_word(h, (i32)W.newDrop[e] | ((i32)W.dead_max0[e] << 8) | ((i32)W.dead_max1[e] << 16) | ((i32)W.fullFlush[e] << 24));
	const i32 s0 = W.slider_off[e], s1 = W.slider_off[e + 1];
	h = bt_hash_word(h, s1 - s0);
	for (i32 s = s0; s < s1; s++) { h = bt_hash_word(h, W.sv[s]); h = bt_hash_word(h, W.smin[s]); h = bt_hash_word(h, W.smax[s]); }
	const i32 wc = W.wire_cnt[e], w0 = W.wire_off[e];
	h = bt_hash_word(h, wc);
	for (i32 k = 0; k < wc; k++) h = bt_hash_word(h, bt_uid_or_none(W, W.wire_to[w0 + k]));
	return h;
}

#endif
