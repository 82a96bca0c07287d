/*
 * bt_oracle.c — TEST INFRASTRUCTURE ONLY.  Plain-C, single-threaded restatement of
 * the sboerwinkle/bittoss per-tick physics hot path, used to CHECK the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or
 * load it; the product (bittoss_b200/) never does.
 *
 * Parity pin: the reference has no golden vectors of its own, so this
 * file is pinned against the reference itself: tests/test_oracle_vs_ref.py steps
 * oracle/_ref/libbtref.so (unmodified reference sources) and this restatement from
 * the same exported world and requires identical per-tick hashes and full-state
 * equality; the resulting hash streams are committed under tests/golden/.
 *
 * Every function cites the reference file:line it follows.  Pointers of the
 * reference are entity indices here (-1 = NULL).  Deliberate, documented
 * differences from the reference (DESIGN.md "Canonical order"):
 *   B1. Broadphase: the reference's velbox tree (velbox/box.cpp) only has to hand
 *       the narrow phase a SUPERSET of the colliding pairs; here
 *       the candidate set is "every live boxed pair of the same world whose swept
 *       AABBs touch", found with a hash grid.  Non-colliding candidates have no
 *       effect (checkCollision returns 0).
 *   B2. The reference folds candidate collisions per hold-root in velbox visiting
 *       order and keeps the incumbent on a full tie (ent.cpp:380).  That order is a
 *       function of velbox's mutation history and cannot be reproduced; here (and in
 *       the CUDA path) the winner is the maximum under collisionBetter with full
 *       ties broken by the larger (buddy, leaf, axis, dir) tuple, which is
 *       independent of visiting order.  For a MOVE_BOTH pair the reference makes
 *       whichever end is scanned first the coordinator (mutual=1, ent.cpp:426-427) and
 *       the later re-evaluation from the other end ties and is dropped; scanning is
 *       newest-first, so here the coordinator is the larger index when both ends need
 *       collision at the start of the iteration, else the one end that does.
 *   B3. Ready collisions of one iteration are pushed in descending hold-root index
 *       (newest first, the reference's gs->ents scan direction).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <time.h>

#include "../include/bittoss_b200.h"

/* ent.h:243-246, 249-256, 266-284 */
#define MOVE_ME 1
#define MOVE_HIM 2
#define MOVE_BOTH 7
#define MOVE_NONE 15
#define HOLD_PASS 0
#define HOLD_DROP 1
#define HOLD_MOVE 2
#define HOLD_MASK 3
#define HOLD_FREEZE 4
#define T_HEAVY 1
#define T_TERRAIN 2
#define T_OBSTACLE 4
#define T_DEBRIS 8
#define T_WEIGHTLESS 16
#define T_FLAG 32
#define TEAM_BIT 256
#define TEAM_MASK (7 * TEAM_BIT)
#define T_COLLIDABLE (T_HEAVY + T_TERRAIN + T_OBSTACLE + T_DEBRIS + T_FLAG)
#define T_EQUIP (1 << 12)
#define T_EQUIP_SM (1 << 13)
#define EQUIP_MASK (T_EQUIP + T_EQUIP_SM)
#define EFFECT_GRAV (1 << 3)
#define T_INPUTS (1 << 11)
#define HOLD_SINGLE 8
#define TICK_BOTTLE 37

typedef int32_t i32;
typedef int32_t v3[3];

typedef struct {
	i32 *items;
	int num, max;
} ilist;

static void il_add(ilist *l, i32 v) {
	if (l->num == l->max) {
		l->max = l->max ? l->max * 2 : 4;
		l->items = (i32 *)realloc(l->items, sizeof(i32) * l->max);
	}
	l->items[l->num++] = v;
}
static int il_has(const ilist *l, i32 v) {
	for (int i = 0; i < l->num; i++)
		if (l->items[i] == v) return 1;
	return 0;
}

typedef struct {
	int n, n_worlds;
	i32 *wbegin, *wend; /* entity range of each world */
	i32 *flip, *gamerules, *rnd, *rootHead, *deadTail;
	i32 *world, *uid;
	v3 *center, *old, *vel, *d_center, *d_vel, *radius;
	uint32_t *type, *collide, *ntype, *ncollide;
	i32 *friction, *holder, *holdFlags, *holdRoot, *holdee, *LLn, *LLp, *newHolder, *newHoldFlags;
	i32 *deadLink; /* e->ll.p reused as the deadTail stack link (ent.cpp:205) */
	uint8_t (*h)[BT_H_NUM];
	uint8_t (*btn)[8];
	uint8_t *newDrop, (*dead_max)[2], *fullFlush, *dead, *inList;
	uint8_t *needsCollision, *needsPhysUpdate, *hasBox;
	uint8_t *needs0; /* needsCollision as of the start of the current doIteration */
	v3 *forced, *forcedHoldees;
	i32 *cLeaf, *cBuddy, *cAxis, *cDir, *cMutual;
	i32 *slider_off, *sv, *smin, *smax;
	ilist *wires, *wiresAdd, *wiresRm;
	/* scratch */
	ilist entList;
	i32 *stamp;
	i32 stampCtr;
	int64_t stats[8];
	int64_t n_events; /* host-side handler invocations reported instead of executed */
	int64_t order_ties; /* full ties resolved by the canonical key */
	ilist crush_log;  /* uids whose crush handler invokeOnCrush would have run, in call order */
} World;

enum { ST_TICKS, ST_ITERS, ST_CAND, ST_POS, ST_PUSH, ST_SERIAL, ST_CRUSH, ST_FUMBLE };

#define ALLOC(p, cnt) p = calloc((size_t)((cnt) > 0 ? (cnt) : 1), sizeof(*(p)))

/* ------------------------------------------------------------------ */
/* small helpers                                                        */

/* needsVelbox ent.cpp:75-77 */
static int needsVelbox(const World *w, i32 e) { return w->collide[e] || (w->type[e] & T_COLLIDABLE); }

/* boundVec ent.cpp:20-30 */
static void boundVec(i32 *values, i32 bound, i32 len) {
	i32 max = bound;
	for (int i = 0; i < len; i++) {
		i32 x = abs(values[i]);
		if (x > max) max = x;
	}
	if (max == bound) return;
	for (int i = 0; i < len; i++) values[i] = (i32)((int64_t)bound * values[i] / max);
}

/* recursiveHoldRoot ent.cpp:39-45 */
static void recursiveHoldRoot(World *w, i32 who, i32 r) {
	w->holdRoot[who] = r;
	for (i32 e = w->holdee[who]; e >= 0; e = w->LLn[e]) recursiveHoldRoot(w, e, r);
}

/* addRootEnt ent.cpp:60-66 */
static void addRootEnt(World *w, i32 e) {
	int g = w->world[e];
	if (w->rootHead[g] >= 0) w->LLp[w->rootHead[g]] = e;
	w->LLn[e] = w->rootHead[g];
	w->LLp[e] = -1;
	w->rootHead[g] = e;
	recursiveHoldRoot(w, e, e);
}

/* rmRootEnt ent.cpp:160-164 */
static void rmRootEnt(World *w, i32 e) {
	int g = w->world[e];
	if (w->LLp[e] >= 0) w->LLn[w->LLp[e]] = w->LLn[e];
	else w->rootHead[g] = w->LLn[e];
	if (w->LLn[e] >= 0) w->LLp[w->LLn[e]] = w->LLp[e];
}

/* entUpdaters.c:58-68 */
static void uSlider(World *w, i32 e, int ix, i32 value) {
	int ns = w->slider_off[e + 1] - w->slider_off[e];
	if (ix < 0 || ix >= ns) return;
	w->fullFlush[e] = 1;
	int s = w->slider_off[e] + ix;
	if (value > w->smax[s]) w->smax[s] = value;
	if (value < w->smin[s]) w->smin[s] = value;
}
/* entGetters.c:42-48 */
static i32 getSlider(const World *w, i32 e, int ix) {
	int ns = w->slider_off[e + 1] - w->slider_off[e];
	if (ix < 0 || ix >= ns) return 0;
	return w->sv[w->slider_off[e] + ix];
}

/* ------------------------------------------------------------------ */
/* handlers, selected by the catalog IDs of handlerRegistrar.h:79-187   */

/* modules/common.c:48, material.c:5-24, stackem.c:9-22, explosion.cpp:8, player.c:84;
 * ID 0 = unset -> MOVE_ME (ent.cpp:220-225); ID 3 aliases 1 (platform.c:56) */
static int whoMovesH(const World *w, i32 a, i32 b, int axis, int dir) {
	uint32_t tb = w->type[b];
	switch (w->h[a][BT_H_WHOMOVES]) {
	case 2: return (tb & T_TERRAIN) ? MOVE_ME : MOVE_HIM;
	case 4: return (tb & T_TERRAIN) ? MOVE_ME : MOVE_BOTH;
	case 5: {
		if (tb & T_TERRAIN) return MOVE_ME;
		int goodTeam = !!(TEAM_MASK & tb & w->type[a]);
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
	default: return MOVE_ME; /* 0 (unset), 1 (move_me), 3 (platform alias) */
	}
}

/* player_push modules/player.c:108-165 (sliders s_equip_processed=10, s_editmode=11) */
static int player_push(World *w, i32 me, i32 him) {
	uint32_t th = w->type[him];
	if (!((th & EQUIP_MASK) && !w->btn[me][1] && !(1 & getSlider(w, me, 10)) && !getSlider(w, me, 11) &&
	      w->holder[him] < 0 && w->holder[me] < 0))
		return 0;
	uint32_t tEquip = th & T_EQUIP, tEquipSm = th & T_EQUIP_SM;
	i32 existing = -1;
	for (i32 h = w->holdee[me]; h >= 0; h = w->LLn[h]) {
		uint32_t heldEquip = w->type[h] & T_EQUIP, heldEquipSm = w->type[h] & T_EQUIP_SM;
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
		int hand = existing >= 0 && !getSlider(w, existing, 0);
		uSlider(w, him, 0, hand);
	}
	w->sv[w->slider_off[me] + 10] |= 1; /* direct write, player.c:157 */
	return 1;
}

static void uPickup(World *w, i32 p, i32 e, i32 holdFlags);
/* push_t handlers: 0 onPushDefault handlerRegistrar.cpp:16; 1 player_push; 2 logic_push modules/logic.c:44-49 */
static int pushH(World *w, i32 me, i32 him, int axis, int dir, i32 displacement, i32 dv) {
	(void)axis; (void)dir; (void)displacement; (void)dv;
	switch (w->h[me][BT_H_PUSH]) {
	case 1: return player_push(w, me, him);
	case 4: /* teamselect_push teamselect.cpp:36-45: acts on gs->players (host); reported as an event */
	case 7: { /* loadblock_push loadblock.cpp:17-31 */
		int id = w->h[me][BT_H_PUSH];
		if (id == 7 && getSlider(w, me, 1) > 0) return 0;
		if (w->h[him][BT_H_TICK] != 7) return 0; /* "him is a player's entity" = runs TICK_PLAYER */
		w->n_events++;
		if (id == 7 && w->slider_off[me + 1] - w->slider_off[me] > 1) w->sv[w->slider_off[me] + 1] = 15;
		return 0;
	}
	case 5: /* scoreboard_push scoreboard.cpp:19-36 */
		if (w->h[me][BT_H_TICK] != w->h[him][BT_H_TICK]) return 0;
		uSlider(w, me, 0, 0);
		uSlider(w, me, 1, 0);
		uSlider(w, me, 2, getSlider(w, him, 2));
		uSlider(w, me, 3, getSlider(w, him, 3));
		for (int k = 0; k < w->wires[me].num; k++) w->btn[w->wires[me].items[k]][4 + 0] = 1;
		return 0;
	case 2:
		if (w->type[him] & (T_TERRAIN | T_OBSTACLE | T_DEBRIS)) w->btn[me][4 + 0] = 1; /* pushBtn(me,0) */
		return 0;
	case 3: /* seat_push seat.cpp:10-16 */
		if (w->wires[me].num) return 0;
		if ((w->type[him] & T_INPUTS) && !w->btn[him][0]) uPickup(w, me, him, HOLD_DROP + HOLD_FREEZE + HOLD_SINGLE);
		return 0;
	case 6: { /* teleport_push teleport.c:12-37 */
		if (w->type[him] & (T_TERRAIN | T_OBSTACLE | T_DEBRIS)) w->btn[me][4 + 0] = 1;
		i32 dest = -1;
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k];
			if (w->btn[x][2]) {
				if (dest >= 0) return 0; /* several lit destinations: undefined, nothing happens */
				dest = x;
			}
		}
		if (dest < 0) return 0;
		i32 r = w->holdRoot[him];
		for (int i = 0; i < 3; i++) {
			w->d_center[r][i] += w->center[dest][i] - w->center[me][i]; /* uCenter(him->holdRoot) */
			w->d_vel[r][i] += w->vel[dest][i] - w->vel[me][i];          /* uVel */
		}
		return 0;
	}
	case 8:   /* pj_push puddlejumper.cpp:67-78: sliders fuel 0, fuel_max 1 */
	case 9: { /* bottle_push bottle.cpp:92-116: sliders hand 0, type 1, amt 2, max 3 */
		int id = w->h[me][BT_H_PUSH];
		if (w->h[him][BT_H_TICK] != TICK_BOTTLE) return 0;
		i32 *ms = &w->sv[w->slider_off[me]], *hs = &w->sv[w->slider_off[him]];
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
		if (poured && hs[2] == 0) { /* bottleUpdate bottle.cpp:118-130 */
			if (w->h[him][BT_H_PUSH] == 9) hs[1] = -1;
			else w->dead_max[him][w->flip[w->world[him]]] |= 1; /* uDead */
		}
		return 0;
	}
	default: return 0;
	}
}

/* pushed_t handlers: 2 fragile_pushed common.c:56; 3 flag_pushed flag.c:13;
 * 5 player_pushed player.c:88-106; 1 pj_pushed puddlejumper.cpp:44-65.
 * returns -1 when the slot is NULL (ent.cpp:556) */
/* uPickup entUpdaters.c:29-52 */
static void uPickup(World *w, i32 p, i32 e, i32 holdFlags) {
	if (w->holder[e] >= 0) return;
	if (w->dead[e] || w->dead[p]) return;
	if (w->newHolder[e] >= 0) {
		if (w->newHolder[e] != p || w->newHoldFlags[e] != holdFlags) w->newHolder[e] = e;
		return;
	}
	w->newHolder[e] = p;
	w->newHoldFlags[e] = holdFlags;
}

static int pushedH(World *w, i32 me, i32 him, int axis, int dir, i32 dx, i32 dv) {
	int id = w->h[me][BT_H_PUSHED];
	switch (id) {
	case 0: return -1;
	case 2: return 1;
	case 3: return (w->type[him] & T_FLAG) ? 1 : 0;
	case 8: { /* bullet_pushed gun.c:90-114 */
		if (getSlider(w, me, 0) == 0) return 0;
		uPickup(w, him, me, HOLD_DROP | HOLD_FREEZE);
		w->sv[w->slider_off[me]] = 0;
		if (dv) {
			i32 scale = 7000;
			i32 ratio = dx * scale / dv;
			for (int i = 0; i < 2; i++) {
				int dim = (axis + 1 + i) % 3;
				w->center[me][dim] -= (w->vel[me][dim] - w->vel[him][dim]) * ratio / scale;
			}
		}
		w->type[me] = 0;
		w->collide[me] = 0;
		return 0;
	}
	case 1:
	case 5: {
		int s_vel = id == 5 ? 5 : 2, s_grounded = id == 5 ? 7 : 4;
		i32 vel[3];
		for (int i = 0; i < 3; i++) vel[i] = w->vel[me][i] - w->vel[him][i]; /* getVel(vel, him, me) */
		if (id == 1) boundVec(vel, 200, 2);
		if (axis == 2) {
			if (dir < 0) uSlider(w, me, s_grounded, 1);
		} else {
			i32 currentSlide = getSlider(w, me, s_vel + axis);
			vel[axis] = currentSlide + dir * dv;
			if (dir * vel[axis] > 0) vel[axis] = 0;
		}
		uSlider(w, me, s_vel, vel[0]);
		uSlider(w, me, s_vel + 1, vel[1]);
		return 0;
	}
	default: return 0;
	}
}

/* updateTrajectory entUpdaters.c:102-114 */
static void updateTrajectory(World *w, i32 who, const i32 *o, const i32 *c, const i32 *v) {
	w->needsPhysUpdate[who] = 1;
	for (int i = 0; i < 3; i++) {
		w->old[who][i] += o[i];
		w->center[who][i] += c[i];
		w->vel[who][i] += v[i];
	}
	for (i32 e = w->holdee[who]; e >= 0; e = w->LLn[e]) updateTrajectory(w, e, o, c, v);
}

/* entPair_t handlers: 3 decor_fumbled common.c:61-67; 5 player_pickup player.c:167-178 */
static void entPairH(World *w, int slot, i32 me, i32 him) {
	if (w->h[me][slot] == 5) {
		if (w->type[him] & EQUIP_MASK) {
			i32 o[3], c[3], v[3];
			for (int i = 0; i < 3; i++) {
				o[i] = w->old[me][i] - w->old[him][i];
				c[i] = w->center[me][i] - w->center[him][i];
				v[i] = w->vel[me][i] - w->vel[him][i];
			}
			updateTrajectory(w, him, o, c, v);
			w->holdFlags[him] |= HOLD_DROP;
		}
		return;
	}
	if (w->h[me][slot] == 1) { /* seat_pickUp seat.cpp:32-36 */
		if (!w->wires[me].num && (w->type[him] & T_INPUTS)) { w->fullFlush[me] = 1; il_add(&w->wiresAdd[me], him); } /* uWire */
		return;
	}
	if (w->h[me][slot] == 2 || w->h[me][slot] == 4) { /* seat_fumble_old / seat_fumble seat.cpp:18-30 */
		il_add(&w->wiresRm[me], him); /* uUnwire */
		if (w->h[me][slot] == 4)
			for (int i = 0; i < 5; i++) uSlider(w, me, i, 0);
		return;
	}
	if (w->h[me][slot] == 3) {
		w->fullFlush[me] = 1;
		w->ntype[me] = w->type[me] | T_DEBRIS; /* uMyTypeMask */
		w->ncollide[me] = w->collide[me] | T_TERRAIN | T_OBSTACLE | T_DEBRIS; /* uMyCollideMask */
		if (slot == BT_H_ONFUMBLED) w->h[me][slot] = 0;
	}
}

/* ------------------------------------------------------------------ */
/* tree edits                                                           */

/* fumble ent.cpp:136-158 */
static void fumble(World *w, i32 h) {
	i32 a = w->holder[h];
	if (a < 0) return;
	w->stats[ST_FUMBLE]++;
	w->needsCollision[h] = 1;
	entPairH(w, BT_H_ONFUMBLE, a, h);
	entPairH(w, BT_H_ONFUMBLED, h, a);
	for (int i = 0; i < 3; i++) {
		i32 x = h, y;
		while ((y = w->holder[x]) >= 0 && w->forced[y][i] && w->forcedHoldees[y][i] == x) {
			w->forced[y][i] = 0;
			x = y;
		}
	}
	if (w->LLn[h] >= 0) w->LLp[w->LLn[h]] = w->LLp[h];
	if (w->LLp[h] >= 0) w->LLn[w->LLp[h]] = w->LLn[h];
	else w->holdee[a] = w->LLn[h];
	w->holder[h] = -1;
	addRootEnt(w, h);
}

/* pickupNoHandlers ent.cpp:166-182 */
static void pickupNoHandlers(World *w, i32 x, i32 y, i32 holdFlags) {
	rmRootEnt(w, y);
	w->LLn[y] = w->holdee[x];
	w->LLp[y] = -1;
	if (w->holdee[x] >= 0) w->LLp[w->holdee[x]] = y;
	w->holdee[x] = y;
	w->holder[y] = x;
	recursiveHoldRoot(w, y, w->holdRoot[x]);
	w->holdFlags[y] = HOLD_MASK & holdFlags;
	if (holdFlags & HOLD_FREEZE)
		for (int i = 0; i < 3; i++) w->d_vel[y][i] = w->vel[x][i] - w->vel[y][i];
}

/* pickup ent.cpp:184-188 */
static void pickup(World *w, i32 x, i32 y, i32 holdFlags) {
	pickupNoHandlers(w, x, y, holdFlags);
	entPairH(w, BT_H_ONPICKUP, x, y);
	entPairH(w, BT_H_ONPICKEDUP, y, x);
}

/* killEntNoHandlers ent.cpp:190-208 */
static void killEntNoHandlers(World *w, i32 e) {
	while (w->holdee[e] >= 0) fumble(w, w->holdee[e]);
	fumble(w, e);
	w->inList[e] = 0; /* entsLlRemove */
	rmRootEnt(w, e);
	w->hasBox[e] = 0; /* velbox_remove */
	int g = w->world[e];
	w->deadLink[e] = w->deadTail[g];
	w->deadTail[g] = e;
	w->dead[e] = 1;
}

/* crushEnt ent.cpp:211-217 */
static void crushEnt(World *w, i32 e) {
	if (w->dead[e]) return;
	w->stats[ST_CRUSH]++;
	killEntNoHandlers(w, e);
}

/* moveRecursive ent.cpp:115-126 */
static void moveRecursive(World *w, i32 who, const i32 *dc, const i32 *dv) {
	w->needsPhysUpdate[who] = 1;
	for (int i = 0; i < 3; i++) {
		w->center[who][i] += dc[i];
		w->vel[who][i] += dv[i];
	}
	for (i32 e = w->holdee[who]; e >= 0; e = w->LLn[e]) moveRecursive(w, e, dc, dv);
}

/* requireCollisionRecursive ent.cpp:128-134 */
static void requireCollisionRecursive(World *w, i32 who) {
	w->needsCollision[who] = 1;
	for (i32 e = w->holdee[who]; e >= 0; e = w->LLn[e]) requireCollisionRecursive(w, e);
}

/* ------------------------------------------------------------------ */
/* narrow phase                                                         */

/* arbitrateMovement ent.cpp:229-236 */
static int arbitrateMovement(const World *w, i32 a, i32 b, int axis, int dir) {
	int A = whoMovesH(w, a, b, axis, dir);
	int B = whoMovesH(w, b, a, axis, -dir);
	if (B == MOVE_ME) B = MOVE_HIM;
	else if (B == MOVE_HIM) B = MOVE_ME;
	A = A & B;
	return A ? A : MOVE_BOTH;
}

/* getAxisAndDir ent.cpp:238-313.  int arithmetic wraps (what g++ -O2 emits). */
static int getAxisAndDir(const World *w, i32 a, i32 b) {
	int d1[3], d2[3], wd[3];
	int i;
	for (i = 0; i < 3; i++) {
		d1[i] = w->old[b][i] - w->old[a][i];
		d2[i] = w->center[b][i] - w->center[a][i];
		wd[i] = w->radius[a][i] + w->radius[b][i];
		if (d1[i] > 0) {
			if (d1[i] >= wd[i] && d2[i] >= wd[i]) return 0;
		} else {
			if (d1[i] <= -wd[i] && d2[i] <= -wd[i]) return 0;
		}
		d2[i] = d1[i] - d2[i];
	}
	int dir_by_axis[3];
	for (i = 0; i < 3; i++) {
		if (d2[i] >= 0) {
			dir_by_axis[i] = -(i + 1);
		} else {
			dir_by_axis[i] = (i + 1);
			d1[i] *= -1;
			d2[i] *= -1;
		}
	}
	for (i = 0; i < 3; i++) d1[i] -= wd[i];
	int best;
#define COMPARE(target, A, B)                                                        \
	do {                                                                             \
		int64_t cmp = ((int64_t)d1[A]) * d2[B] - ((int64_t)d1[B]) * d2[A];           \
		if (cmp > 0) target = A;                                                     \
		else if (cmp < 0) target = B;                                                \
		else {                                                                       \
			if (wd[A] > wd[B]) target = B;                                           \
			else target = A;                                                         \
		}                                                                            \
	} while (0)
	COMPARE(best, 0, 1);
	{
		int bb = best;
		COMPARE(best, bb, 2);
	}
#undef COMPARE
	if (d1[best] < 0) {
		int depth = INT32_MAX;
		int ret = 0;
		for (i = 0; i < 3; i++) {
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
	for (i = 0; i < 3; i++) {
		int far = d1[i] + 2 * wd[i];
		int64_t cmp = ((int64_t)d1[best]) * d2[i] - ((int64_t)far) * d2[best];
		if (cmp >= 0) return 0;
	}
	return dir_by_axis[best];
}

typedef struct {
	i32 leaf, buddy;
	int axis, dir, mutual;
} Coll;

static inline int imin(int a, int b) { return a < b ? a : b; }

/* collisionBetter ent.cpp:317-382 on explicit collision records: returns +1 if c0 is
 * strictly better than c1, -1 if strictly worse, 0 on a full tie.
 * (the reference returns `better ? 1 : 0`; worse and tie both keep the incumbent) */
static int collisionCmp(const World *w, const Coll *c0, const Coll *c1) {
	const Coll *c[2] = {c0, c1};
	int64_t v64s[2];
	for (int i = 0; i < 2; i++) {
		int o = 1 - i;
		int a_i = c[i]->axis, a_o = c[o]->axis;
		i32 li = c[i]->leaf, fi = c[i]->buddy, lo = c[o]->leaf, fo = c[o]->buddy;
		/* 32-bit wrapping arithmetic, as in the reference (ent.cpp:338-343) */
		int32_t delta = (int32_t)((uint32_t)(w->old[li][a_i] - w->old[fi][a_i]) * (uint32_t)c[i]->dir) -
		                w->radius[fi][a_i] - w->radius[li][a_i];
		int32_t speed = (int32_t)((uint32_t)(w->center[fo][a_o] - w->old[fo][a_o] + w->old[lo][a_o] -
		                                     w->center[lo][a_o]) * (uint32_t)c[o]->dir);
		if (speed < 1) speed = 1;
		v64s[i] = (int32_t)((uint32_t)delta * (uint32_t)speed);
	}
	if (v64s[0] != v64s[1]) return v64s[0] < v64s[1] ? 1 : -1;
	i32 vals[2];
	for (int i = 0; i < 2; i++) vals[i] = !c[i]->mutual;
	if (vals[0] != vals[1]) return vals[0] > vals[1] ? 1 : -1;
	for (int i = 0; i < 2; i++) {
		i32 l = c[i]->leaf, f = c[i]->buddy;
		int x1 = (c[i]->axis + 1) % 3, x2 = (c[i]->axis + 2) % 3;
		i32 k1 = w->radius[f][x1] + w->radius[l][x1] - abs(w->old[f][x1] - w->old[l][x1]);
		i32 k2 = w->radius[f][x2] + w->radius[l][x2] - abs(w->old[f][x2] - w->old[l][x2]);
		vals[i] = imin(k1, k2);
	}
	if (vals[0] != vals[1]) return vals[0] > vals[1] ? 1 : -1;
	for (int i = 0; i < 2; i++) {
		i32 l = c[i]->leaf, f = c[i]->buddy;
		int ax = c[i]->axis;
		vals[i] = w->radius[f][ax] + w->radius[l][ax] + (w->center[f][ax] - w->center[l][ax]) * c[i]->dir;
	}
	if (vals[0] != vals[1]) return vals[0] > vals[1] ? 1 : -1;
	return 0;
}

static Coll pendingOf(const World *w, i32 root) {
	Coll c = {w->cLeaf[root], w->cBuddy[root], w->cAxis[root], w->cDir[root], w->cMutual[root]};
	return c;
}

/* tryCollide ent.cpp:384-396 with the order-independent tie rule B2 */
static int tryCollide(World *w, i32 leaf, i32 buddy, int axis, int dir, int mutual) {
	i32 root = w->holdRoot[leaf];
	Coll c = {leaf, buddy, axis, dir, mutual};
	if (w->cBuddy[root] >= 0) {
		Coll inc = pendingOf(w, root);
		int r = collisionCmp(w, &c, &inc);
		if (r < 0) return 0;
		if (r == 0) {
			/* full tie: larger (buddy, leaf, axis, dir) wins */
			int64_t kc[4] = {c.buddy, c.leaf, c.axis, c.dir};
			int64_t ki[4] = {inc.buddy, inc.leaf, inc.axis, inc.dir};
			int k = 0;
			while (k < 4 && kc[k] == ki[k]) k++;
			if (k < 4) w->order_ties++; /* two different collisions, equal under collisionBetter: the reference keeps the first it met */
			if (k == 4 || kc[k] < ki[k]) return 0;
		}
	} else {
		il_add(&w->entList, root);
	}
	w->cLeaf[root] = leaf;
	w->cBuddy[root] = buddy;
	w->cAxis[root] = axis;
	w->cDir[root] = dir;
	w->cMutual[root] = mutual;
	return 1;
}

/* checkCollision ent.cpp:398-435 */
static int checkCollision(World *w, i32 a, i32 b) {
	int AonB = 0 != (w->type[b] & w->collide[a]);
	int BonA = 0 != (w->type[a] & w->collide[b]);
	if (!AonB && !BonA) return 0;
	if (w->holdRoot[a] == w->holdRoot[b]) return 0;
	int dir = getAxisAndDir(w, a, b);
	if (dir == 0) return 0;
	w->stats[ST_POS]++;
	int axis;
	if (dir < 0) {
		axis = -dir - 1;
		dir = -1;
	} else {
		axis = dir - 1;
		dir = 1;
	}
	int movee = (AonB ? (BonA ? arbitrateMovement(w, a, b, axis, dir) : whoMovesH(w, a, b, axis, dir))
	                  : 3 ^ whoMovesH(w, b, a, axis, -dir));
	if (movee & 8) return 0;
	if (movee & 4) {
		/* a is being scanned; it coordinates unless b is scanned too and comes first */
		int coordA = w->needs0[b] ? (a > b) : 1;
		int ret = tryCollide(w, a, b, axis, dir, coordA ? 1 : 2);
		ret |= tryCollide(w, b, a, axis, -dir, coordA ? 2 : 1);
		return ret;
	}
	if (movee == MOVE_ME) return tryCollide(w, a, b, axis, dir, 0);
	return tryCollide(w, b, a, axis, -dir, 0);
}

/* ------------------------------------------------------------------ */
/* broadphase stand-in (B1): two-level hash grid over swept AABBs       */

typedef struct {
	i32 ent, next;
} GEntry;
typedef struct {
	int shift;
	uint32_t mask; /* table size - 1 */
	i32 *head;
	uint64_t *key;
	GEntry *entries;
	int n_entries, cap_entries;
} Grid;

static void swept(const World *w, i32 e, i32 lo[3], i32 hi[3]) {
	for (int k = 0; k < 3; k++) {
		i32 a = w->old[e][k], b = w->center[e][k], r = w->radius[e][k];
		lo[k] = (a < b ? a : b) - r;
		hi[k] = (a > b ? a : b) + r;
	}
}
static uint64_t cellKey(i32 x, i32 y, i32 z) {
	return ((uint64_t)(uint32_t)(x & 0x1fffff) << 42) | ((uint64_t)(uint32_t)(y & 0x1fffff) << 21) |
	       (uint64_t)(uint32_t)(z & 0x1fffff);
}
static uint32_t cellHash(uint64_t k) {
	k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 29;
	return (uint32_t)k;
}
/* expect = total number of (entity, cell) entries, an upper bound of distinct cells */
static void gridInit(Grid *g, int shift, int64_t expect) {
	uint32_t sz = 1024;
	while ((int64_t)sz < expect * 2) sz <<= 1;
	g->shift = shift;
	g->mask = sz - 1;
	g->head = (i32 *)malloc(sizeof(i32) * sz);
	g->key = (uint64_t *)malloc(sizeof(uint64_t) * sz);
	memset(g->head, 0xff, sizeof(i32) * sz);
	g->cap_entries = (int)expect + 16;
	g->entries = (GEntry *)malloc(sizeof(GEntry) * g->cap_entries);
	g->n_entries = 0;
}
static void gridFree(Grid *g) {
	free(g->head); free(g->key); free(g->entries);
}
static i32 *gridSlot(Grid *g, uint64_t key, int create) {
	uint32_t s = cellHash(key) & g->mask;
	while (1) {
		if (g->head[s] == -1) {
			if (!create) return NULL;
			g->key[s] = key;
			g->head[s] = -2; /* empty list marker */
			return &g->head[s];
		}
		if (g->key[s] == key) return &g->head[s];
		s = (s + 1) & g->mask;
	}
}
static void gridInsert(Grid *g, i32 e, const i32 lo[3], const i32 hi[3]) {
	for (i32 x = lo[0] >> g->shift; x <= hi[0] >> g->shift; x++)
		for (i32 y = lo[1] >> g->shift; y <= hi[1] >> g->shift; y++)
			for (i32 z = lo[2] >> g->shift; z <= hi[2] >> g->shift; z++) {
				i32 *slot = gridSlot(g, cellKey(x, y, z), 1);
				if (g->n_entries == g->cap_entries) {
					g->cap_entries *= 2;
					g->entries = (GEntry *)realloc(g->entries, sizeof(GEntry) * g->cap_entries);
				}
				g->entries[g->n_entries].ent = e;
				g->entries[g->n_entries].next = (*slot == -2) ? -1 : *slot;
				*slot = g->n_entries++;
			}
}
static int64_t cellCount(const i32 lo[3], const i32 hi[3], int shift) {
	int64_t c = 1;
	for (int k = 0; k < 3; k++) c *= (int64_t)((hi[k] >> shift) - (lo[k] >> shift) + 1);
	return c;
}

#define FINE_SHIFT 10
#define COARSE_SHIFT 14

/* doIteration ent.cpp:437-462 for one world */
static int doIteration(World *w, int g) {
	int ret = 0;
	w->entList.num = 0;
	i32 b0 = w->wbegin[g], b1 = w->wend[g];
	/* is anybody asking? */
	int any = 0, nbox = 0;
	for (i32 e = b0; e < b1; e++) w->needs0[e] = (uint8_t)(w->inList[e] && w->needsCollision[e] && w->hasBox[e]);
	for (i32 e = b0; e < b1; e++)
		if (w->inList[e]) {
			if (w->hasBox[e]) nbox++;
			if (w->needsCollision[e] && w->hasBox[e]) any = 1;
		}
	if (any) {
		Grid fine, coarse;
		ilist big = {0, 0, 0}, fineM = {0, 0, 0}, coarseM = {0, 0, 0};
		int64_t nfine = 0, ncoarse = 0;
		for (i32 e = b0; e < b1; e++) {
			if (!w->inList[e] || !w->hasBox[e]) continue;
			i32 lo[3], hi[3];
			swept(w, e, lo, hi);
			int64_t cf = cellCount(lo, hi, FINE_SHIFT), cc = cellCount(lo, hi, COARSE_SHIFT);
			if (cf <= 27) nfine += cf;
			else if (cc <= 27) ncoarse += cc;
		}
		(void)nbox;
		gridInit(&fine, FINE_SHIFT, nfine);
		gridInit(&coarse, COARSE_SHIFT, ncoarse);
		for (i32 e = b0; e < b1; e++) {
			if (!w->inList[e] || !w->hasBox[e]) continue;
			i32 lo[3], hi[3];
			swept(w, e, lo, hi);
			if (cellCount(lo, hi, FINE_SHIFT) <= 27) { gridInsert(&fine, e, lo, hi); il_add(&fineM, e); }
			else if (cellCount(lo, hi, COARSE_SHIFT) <= 27) { gridInsert(&coarse, e, lo, hi); il_add(&coarseM, e); }
			else il_add(&big, e);
		}
		/* scan order: gs->ents head to tail = newest first = descending index */
		for (i32 e = b1 - 1; e >= b0; e--) {
			if (!w->inList[e] || !w->needsCollision[e]) continue;
			if (w->hasBox[e]) {
				i32 lo[3], hi[3];
				swept(w, e, lo, hi);
				i32 st = ++w->stampCtr;
				w->stamp[e] = st;
				Grid *gs[2] = {&fine, &coarse};
				ilist *members[2] = {&fineM, &coarseM};
				for (int gi = 0; gi < 2; gi++) {
					Grid *G = gs[gi];
					if (cellCount(lo, hi, G->shift) > 64) {
						/* e spans too many cells of this level: test every member instead */
						for (int k = 0; k < members[gi]->num; k++) {
							i32 o = members[gi]->items[k];
							if (w->stamp[o] == st) continue;
							w->stamp[o] = st;
							w->stats[ST_CAND]++;
							ret |= checkCollision(w, e, o);
						}
						continue;
					}
					for (i32 x = lo[0] >> G->shift; x <= hi[0] >> G->shift; x++)
						for (i32 y = lo[1] >> G->shift; y <= hi[1] >> G->shift; y++)
							for (i32 z = lo[2] >> G->shift; z <= hi[2] >> G->shift; z++) {
								i32 *slot = gridSlot(G, cellKey(x, y, z), 0);
								if (!slot || *slot < 0) continue;
								for (i32 it = *slot; it >= 0; it = G->entries[it].next) {
									i32 o = G->entries[it].ent;
									if (w->stamp[o] == st) continue;
									w->stamp[o] = st;
									w->stats[ST_CAND]++;
									ret |= checkCollision(w, e, o);
								}
							}
				}
				for (int k = 0; k < big.num; k++) {
					i32 o = big.items[k];
					if (w->stamp[o] == st) continue;
					w->stamp[o] = st;
					w->stats[ST_CAND]++;
					ret |= checkCollision(w, e, o);
				}
			}
			w->needsCollision[e] = 0;
		}
		gridFree(&fine);
		gridFree(&coarse);
		free(big.items); free(fineM.items); free(coarseM.items);
	} else {
		for (i32 e = b0; e < b1; e++)
			if (w->inList[e]) w->needsCollision[e] = 0;
	}
	return ret;
}

/* ------------------------------------------------------------------ */
/* resolution                                                           */

/* pinch ent.cpp:483-507 */
static i32 pinch(World *w, i32 e, int axis, i32 o) {
	i32 leaf1 = e;
	while (w->forcedHoldees[leaf1][axis] >= 0) leaf1 = w->forcedHoldees[leaf1][axis];
	i32 leaf2;
	if (o >= 0) {
		leaf2 = o;
		while (w->forcedHoldees[leaf2][axis] >= 0) leaf2 = w->forcedHoldees[leaf2][axis];
	} else {
		leaf2 = e;
	}
	if (leaf1 != e) {
		fumble(w, leaf1);
		if (leaf2 != e) {
			fumble(w, leaf2);
			return leaf2;
		}
		return e;
	}
	if (leaf2 != e) {
		fumble(w, leaf2);
		return leaf2;
	}
	crushEnt(w, e);
	return -1;
}

/* push_helper ent.cpp:509-525 */
static int push_helper(World *w, i32 e, i32 prev, int axis, int dir) {
	if (w->forced[e][axis] * dir < 0) {
		i32 ret = pinch(w, e, axis, prev);
		if (ret < 0) return 1;
		if (e != ret) return 0;
	}
	w->forced[e][axis] += dir;
	if (w->forced[e][axis] == dir * 15) {
		crushEnt(w, e);
		return 1;
	}
	w->forcedHoldees[e][axis] = prev;
	return 0;
}

/* getPushDv ent.cpp:527-540 */
static void getPushDv(const World *w, i32 dv[3], i32 e, i32 o, int axis, int dir, i32 accel) {
	dv[axis] = 0;
	int i = (axis + 1) % 3;
	dv[i] = w->vel[o][i] - w->vel[e][i];
	i = (i + 1) % 3;
	dv[i] = w->vel[o][i] - w->vel[e][i];
	boundVec(dv, w->friction[e], 3);
	dv[axis] = accel * dir;
}

/* doPushPickup ent.cpp:542-549 */
static void doPushPickup(World *w, i32 e, i32 holder) {
	if (w->holder[e] >= 0 || w->holdRoot[holder] == e) {
		crushEnt(w, e);
		return;
	}
	pickup(w, holder, e, 0);
}

/* doPushRegular ent.cpp:551-577 */
static void doPushRegular(World *w, i32 e, i32 o, int axis, int dir, i32 displacement, const i32 dv[3]) {
	i32 prev = -1;
	while (1) {
		if (push_helper(w, e, prev, axis, dir)) break;
		{
			i32 accel = dv[axis] * dir;
			int r = pushedH(w, e, o, axis, dir, displacement, accel);
			if (r > 0) {
				crushEnt(w, e);
				break;
			}
		}
		prev = e;
		e = w->holder[e];
		if (e >= 0) {
			i32 holdFlags = w->holdFlags[prev];
			if (holdFlags == HOLD_PASS) continue;
			if (holdFlags == HOLD_DROP) fumble(w, prev);
		}
		i32 dc[3];
		dc[(axis + 1) % 3] = dc[(axis + 2) % 3] = 0;
		dc[axis] = displacement * dir;
		moveRecursive(w, prev, dc, dv);
		break;
	}
}

/* push ent.cpp:579-629 */
static void push(World *w, i32 e, i32 o, int axis, int dir, int mutual) {
	w->stats[ST_PUSH]++;
	i32 displacement = (w->center[o][axis] - w->center[e][axis]) * dir + w->radius[e][axis] + w->radius[o][axis];
	i32 accel = (w->vel[o][axis] - w->vel[e][axis]) * dir;
	if (accel < 0) accel = 0;
	if (mutual) {
		int p1 = pushH(w, o, e, axis, dir, displacement, accel);
		int p2 = pushH(w, e, o, axis, -dir, displacement, accel);
		if (p1 != p2) {
			if (p1) { i32 t = e; e = o; o = t; } /* the branchless swap of ent.cpp:601-603 */
			dir *= (p2 - p1);
			i32 dv[3];
			getPushDv(w, dv, e, o, axis, dir, accel);
			doPushRegular(w, e, o, axis, dir, displacement, dv);
			doPushPickup(w, o, e);
		} else {
			i32 dv1[3], dv2[3];
			getPushDv(w, dv1, e, o, axis, dir, accel);
			getPushDv(w, dv2, o, e, axis, -dir, accel);
			doPushRegular(w, e, o, axis, dir, displacement, dv1);
			doPushRegular(w, o, e, axis, -dir, displacement, dv2);
		}
	} else {
		int p1 = pushH(w, o, e, axis, dir, displacement, accel);
		if (p1) {
			doPushPickup(w, e, o);
		} else {
			i32 dv[3];
			getPushDv(w, dv, e, o, axis, dir, accel);
			doPushRegular(w, e, o, axis, dir, displacement, dv);
		}
	}
}

/* shouldResolveCollision ent.cpp:631-672.  *cyc is set when the decision came
 * from walking a cycle back to e (used only for the serial-lane statistic). */
static int shouldResolveCollision(const World *w, i32 e, int *cyc) {
	*cyc = 0;
	if (w->cMutual[e] == 2) return 0;
	i32 otherRoot = w->holdRoot[w->cBuddy[e]];
	if (w->cBuddy[otherRoot] < 0) return 1;
	i32 loopDetector = otherRoot;
	int flip = 0;
	Coll mine = pendingOf(w, e);
	while (1) {
		Coll other = pendingOf(w, otherRoot);
		/* collisionBetter(e, other...) : is `other` strictly better than mine? */
		if (collisionCmp(w, &other, &mine) > 0) return 0;
		otherRoot = w->holdRoot[w->cBuddy[otherRoot]];
		if (otherRoot == e) {
			*cyc = 1;
			return 1;
		}
		if (w->cBuddy[otherRoot] < 0 || otherRoot == loopDetector) return 0;
		if (flip) loopDetector = w->holdRoot[w->cBuddy[loopDetector]];
		flip = !flip;
	}
}

/* ------------------------------------------------------------------ */
/* flushes                                                              */

/* flushMisc ent.cpp:674-714 */
static void flushMisc(World *w, i32 e) {
	if (w->fullFlush[e]) {
		w->fullFlush[e] = 0;
		for (int s = w->slider_off[e]; s < w->slider_off[e + 1]; s++) {
			if (w->smax[s] < w->smin[s]) continue;
			w->sv[s] = (i32)(((int64_t)w->smax[s] + w->smin[s]) / 2);
			w->smax[s] = INT32_MIN;
			w->smin[s] = INT32_MAX;
		}
		if (w->ntype[e] != w->type[e] || w->collide[e] != w->ncollide[e]) {
			w->type[e] = w->ntype[e];
			w->collide[e] = w->ncollide[e];
			w->hasBox[e] = (uint8_t)needsVelbox(w, e); /* fixHasVelbox ent.cpp:79-89 */
		}
		for (int i = 0; i < w->wiresAdd[e].num; i++) {
			i32 x = w->wiresAdd[e].items[i];
			if (!il_has(&w->wires[e], x)) il_add(&w->wires[e], x);
		}
		w->wiresAdd[e].num = 0;
	}
	ilist *wl = &w->wires[e];
	for (int i = 0; i < wl->num; i++) {
		i32 x = wl->items[i];
		if (w->dead[x] || il_has(&w->wiresRm[e], x)) {
			wl->items[i] = wl->items[--wl->num]; /* quickRmAt */
			i--;
		}
	}
	w->wiresRm[e].num = 0;
}

/* flushPosVel ent.cpp:716-744 */
static void flushPosVel(World *w, i32 e, const i32 *pdc, const i32 *pdv) {
	for (int i = 0; i < 3; i++) {
		w->d_center[e][i] += pdc[i];
		w->center[e][i] += w->d_center[e][i];
		w->old[e][i] += w->d_center[e][i];
		w->d_vel[e][i] += pdv[i];
		w->vel[e][i] += w->d_vel[e][i];
	}
	for (i32 c = w->holdee[e]; c >= 0; c = w->LLn[c]) flushPosVel(w, c, w->d_center[e], w->d_vel[e]);
	for (int i = 0; i < 3; i++) {
		w->d_center[e][i] = 0;
		w->d_vel[e][i] = 0;
	}
}

/* flush1 ent.cpp:794-804 (flushDeaths :747-761, flushDrops :765-773, flushPickups :775-792) */
static void flush1(World *w, int g) {
	i32 b0 = w->wbegin[g], b1 = w->wend[g];
	w->flip[g] ^= 1;
	for (i32 i = b1 - 1; i >= b0; i--) {
		if (!w->inList[i]) continue;
		uint8_t d = w->dead_max[i][1 ^ w->flip[g]];
		if (d) {
			crushEnt(w, i);
			w->dead[i] = d;
		}
	}
	for (i32 i = b1 - 1; i >= b0; i--) {
		if (!w->inList[i]) continue;
		if (w->newDrop[i]) {
			w->newDrop[i] = 0;
			fumble(w, i);
		}
	}
	for (i32 i = w->rootHead[g], next; i >= 0; i = next) {
		next = w->LLn[i];
		i32 x = w->newHolder[i];
		if (x < 0) continue;
		if (w->holdRoot[x] != i && !w->dead[x]) pickup(w, x, i, w->newHoldFlags[i]);
		w->newHolder[i] = -1;
	}
	for (i32 i = b1 - 1; i >= b0; i--)
		if (w->inList[i]) flushMisc(w, i);
}

/* flush2 ent.cpp:806-811 */
static void flush2(World *w, int g) {
	static const i32 zero[3] = {0, 0, 0};
	for (i32 i = w->rootHead[g]; i >= 0; i = w->LLn[i]) flushPosVel(w, i, zero, zero);
}

/* doGravity effects.cpp:35-41 (uVel entUpdaters.c:14-18) */
static void doGravity(World *w, int g) {
	for (i32 i = w->rootHead[g]; i >= 0; i = w->LLn[i])
		if (!(w->type[i] & T_WEIGHTLESS)) w->d_vel[i][2] += 32;
}

static int cmp_desc(const void *a, const void *b) {
	i32 x = *(const i32 *)a, y = *(const i32 *)b;
	return x < y ? 1 : (x > y ? -1 : 0);
}

/* doPhysics ent.cpp:842-911 for one world */
static void doPhysics(World *w, int g) {
	i32 b0 = w->wbegin[g], b1 = w->wend[g];
	flush2(w, g);
	for (i32 i = b1 - 1; i >= b0; i--) {
		if (!w->inList[i]) continue;
		for (int j = 0; j < 3; j++) {
			w->old[i][j] = w->center[i][j];
			w->center[i][j] += w->vel[i][j];
			w->forced[i][j] = 0;
		}
		w->needsCollision[i] = 1;
		w->needsPhysUpdate[i] = 0;
	}
	int dbg_it = 0;
	while (doIteration(w, g)) {
		if (++dbg_it > 300) { fprintf(stderr, "DBG: >300 iterations, entList %d\n", w->entList.num); 
			for (int j = 0; j < w->entList.num && j < 5; j++) { i32 e = w->entList.items[j]; fprintf(stderr, " root %d leaf %d buddy %d axis %d dir %d mut %d forced %d %d %d c %d %d %d bc %d %d %d\n", e, w->cLeaf[e], w->cBuddy[e], w->cAxis[e], w->cDir[e], w->cMutual[e], w->forced[e][0], w->forced[e][1], w->forced[e][2], w->center[e][0], w->center[e][1], w->center[e][2], w->center[w->cBuddy[e]][0], w->center[w->cBuddy[e]][1], w->center[w->cBuddy[e]][2]); }
			abort(); }
		w->stats[ST_ITERS]++;
		int realNum = w->entList.num;
		/* B3: canonical processing order = descending hold-root index */
		qsort(w->entList.items, realNum, sizeof(i32), cmp_desc);
		i32 *all = (i32 *)malloc(sizeof(i32) * realNum);
		memcpy(all, w->entList.items, sizeof(i32) * realNum);
		uint8_t *ready = (uint8_t *)malloc(realNum);
		for (int j = 0; j < realNum; j++) {
			int cyc;
			ready[j] = (uint8_t)shouldResolveCollision(w, all[j], &cyc);
		}
		for (int j = 0; j < realNum; j++) {
			if (!ready[j]) continue;
			i32 e = all[j];
			push(w, w->cLeaf[e], w->cBuddy[e], w->cAxis[e], w->cDir[e], w->cMutual[e]);
		}
		for (int j = 0; j < realNum; j++) {
			i32 e = all[j];
			w->cBuddy[e] = -1;
			requireCollisionRecursive(w, e);
		}
		free(all);
		free(ready);
		/* ent.cpp:892-904: the velbox repair of moved ents has no counterpart here (B1) */
		for (i32 i = b0; i < b1; i++)
			if (w->inList[i]) w->needsPhysUpdate[i] = 0;
	}
	/* invokeOnCrush ent.cpp:464-469: the crush handlers themselves create entities (explode) and
	 * stay on the host; what is restated is WHICH entities they run for and in what order */
	for (i32 e = w->deadTail[g]; e >= 0; e = w->deadLink[e])
		if (w->h[e][BT_H_CRUSH] && w->dead[e] == 1) il_add(&w->crush_log, w->uid[e]);
	flush1(w, g);
	flush2(w, g);
}

/* finishStep ent.cpp:913-915 -> clearDeads :471-481 */
static void finishStep(World *w, int g) { w->deadTail[g] = -1; }

/* ------------------------------------------------------------------ */
/* C API (see oracle/pybind.py)                                         */

void bto_destroy(World *w);

World *bto_create(const bt_world_desc *d) {
	World *w = (World *)calloc(1, sizeof(World));
	int n = d->n_ents, nw = d->n_worlds > 0 ? d->n_worlds : 1;
	w->n = n;
	w->n_worlds = nw;
	ALLOC(w->wbegin, nw); ALLOC(w->wend, nw); ALLOC(w->flip, nw); ALLOC(w->gamerules, nw);
	ALLOC(w->rnd, nw); ALLOC(w->rootHead, nw); ALLOC(w->deadTail, nw);
	ALLOC(w->world, n); ALLOC(w->uid, n);
	ALLOC(w->center, n); ALLOC(w->old, n); ALLOC(w->vel, n); ALLOC(w->d_center, n); ALLOC(w->d_vel, n);
	ALLOC(w->radius, n);
	ALLOC(w->type, n); ALLOC(w->collide, n); ALLOC(w->ntype, n); ALLOC(w->ncollide, n);
	ALLOC(w->friction, n); ALLOC(w->holder, n); ALLOC(w->holdFlags, n); ALLOC(w->holdRoot, n);
	ALLOC(w->holdee, n); ALLOC(w->LLn, n); ALLOC(w->LLp, n); ALLOC(w->newHolder, n);
	ALLOC(w->newHoldFlags, n); ALLOC(w->deadLink, n);
	ALLOC(w->h, n); ALLOC(w->btn, n);
	ALLOC(w->newDrop, n); ALLOC(w->dead_max, n); ALLOC(w->fullFlush, n); ALLOC(w->dead, n);
	ALLOC(w->inList, n); ALLOC(w->needsCollision, n); ALLOC(w->needsPhysUpdate, n); ALLOC(w->hasBox, n);
	ALLOC(w->needs0, n);
	ALLOC(w->forced, n); ALLOC(w->forcedHoldees, n);
	ALLOC(w->cLeaf, n); ALLOC(w->cBuddy, n); ALLOC(w->cAxis, n); ALLOC(w->cDir, n); ALLOC(w->cMutual, n);
	ALLOC(w->slider_off, n + 1); ALLOC(w->sv, d->n_sliders); ALLOC(w->smin, d->n_sliders);
	ALLOC(w->smax, d->n_sliders);
	ALLOC(w->wires, n); ALLOC(w->wiresAdd, n); ALLOC(w->wiresRm, n);
	ALLOC(w->stamp, n);
	for (int g = 0; g < nw; g++) {
		w->wbegin[g] = n; w->wend[g] = 0;
		w->flip[g] = d->w_flip_flop[g];
		w->gamerules[g] = d->w_gamerules[g];
		w->rnd[g] = d->w_rand[g];
		w->rootHead[g] = -1;
		w->deadTail[g] = -1;
	}
	for (int i = 0; i < n; i++) {
		int g = d->world_id ? d->world_id[i] : 0;
		if (g < 0 || g >= nw) { bto_destroy(w); return NULL; }
		w->world[i] = g;
		if (i < w->wbegin[g]) w->wbegin[g] = i;
		if (i + 1 > w->wend[g]) w->wend[g] = i + 1;
		w->uid[i] = d->uid[i];
		for (int k = 0; k < 3; k++) {
			w->center[i][k] = d->center[3 * i + k];
			w->old[i][k] = d->old[3 * i + k];
			w->vel[i][k] = d->vel[3 * i + k];
			w->d_center[i][k] = d->d_center[3 * i + k];
			w->d_vel[i][k] = d->d_vel[3 * i + k];
			w->radius[i][k] = d->radius[3 * i + k];
			w->forcedHoldees[i][k] = -1;
		}
		w->type[i] = d->type_mask[i]; w->collide[i] = d->collide_mask[i];
		w->ntype[i] = d->new_type_mask[i]; w->ncollide[i] = d->new_collide_mask[i];
		w->friction[i] = d->friction[i];
		w->holder[i] = d->holder[i]; w->holdFlags[i] = d->hold_flags[i];
		w->holdee[i] = d->holdee_head[i];
		w->newHolder[i] = d->new_holder[i]; w->newHoldFlags[i] = d->new_hold_flags[i];
		memcpy(w->h[i], d->handlers + (size_t)BT_H_NUM * i, BT_H_NUM);
		memcpy(w->btn[i], d->btn + 8 * (size_t)i, 8);
		const uint8_t *m = d->misc + (size_t)BT_M_NUM * i;
		w->newDrop[i] = m[BT_M_NEWDROP];
		w->dead_max[i][0] = m[BT_M_DEADMAX0];
		w->dead_max[i][1] = m[BT_M_DEADMAX1];
		w->fullFlush[i] = m[BT_M_FULLFLUSH];
		w->inList[i] = 1;
		w->hasBox[i] = (uint8_t)needsVelbox(w, i);
		w->cBuddy[i] = -1; w->cLeaf[i] = -1;
		w->slider_off[i] = d->slider_off[i];
		for (int k = d->wire_off[i]; k < d->wire_off[i + 1]; k++) il_add(&w->wires[i], d->wire_to[k]);
	}
	w->slider_off[n] = d->slider_off[n];
	for (int s = 0; s < d->n_sliders; s++) {
		w->sv[s] = d->slider_v[s]; w->smin[s] = d->slider_min[s]; w->smax[s] = d->slider_max[s];
	}
	for (int g = 0; g < nw; g++)
		if (w->wbegin[g] > w->wend[g]) w->wbegin[g] = w->wend[g] = 0;
	/* sibling links of held ents; root list in rootEnts order */
	for (int i = 0; i < n; i++) { w->LLn[i] = -1; w->LLp[i] = -1; }
	for (int i = 0; i < n; i++)
		if (w->holder[i] >= 0) {
			w->LLn[i] = d->holdee_next[i];
			if (w->LLn[i] >= 0) w->LLp[w->LLn[i]] = i;
		}
	for (int g = 0; g < nw; g++) {
		/* bucket roots by root_order (a permutation of 0..nroots-1 inside the world) */
		int nroots = 0;
		for (i32 i = w->wbegin[g]; i < w->wend[g]; i++)
			if (w->holder[i] < 0) nroots++;
		i32 *byOrder = (i32 *)malloc(sizeof(i32) * (nroots + 1));
		for (int k = 0; k < nroots; k++) byOrder[k] = -1;
		int ok = 1;
		for (i32 i = w->wbegin[g]; i < w->wend[g]; i++)
			if (w->holder[i] < 0) {
				i32 r = d->root_order[i];
				if (r < 0 || r >= nroots || byOrder[r] >= 0) ok = 0;
				else byOrder[r] = i;
			}
		if (!ok) { /* fall back: newest first, the order a freshly built world has */
			int k = 0;
			for (i32 i = w->wend[g] - 1; i >= w->wbegin[g]; i--)
				if (w->holder[i] < 0) byOrder[k++] = i;
		}
		i32 prev = -1;
		for (int k = 0; k < nroots; k++) {
			i32 i = byOrder[k];
			w->LLp[i] = prev;
			if (prev >= 0) w->LLn[prev] = i;
			else w->rootHead[g] = i;
			prev = i;
		}
		if (prev >= 0) w->LLn[prev] = -1;
		free(byOrder);
		for (i32 i = w->rootHead[g]; i >= 0; i = w->LLn[i]) recursiveHoldRoot(w, i, i);
	}
	return w;
}

void bto_destroy(World *w) {
	if (!w) return;
	for (int i = 0; i < w->n; i++) {
		if (w->wires) free(w->wires[i].items);
		if (w->wiresAdd) free(w->wiresAdd[i].items);
		if (w->wiresRm) free(w->wiresRm[i].items);
	}
	free(w->wbegin); free(w->wend); free(w->flip); free(w->gamerules); free(w->rnd); free(w->rootHead);
	free(w->deadTail); free(w->world); free(w->uid); free(w->center); free(w->old); free(w->vel);
	free(w->d_center); free(w->d_vel); free(w->radius); free(w->type); free(w->collide); free(w->ntype);
	free(w->ncollide); free(w->friction); free(w->holder); free(w->holdFlags); free(w->holdRoot);
	free(w->holdee); free(w->LLn); free(w->LLp); free(w->newHolder); free(w->newHoldFlags);
	free(w->deadLink); free(w->h); free(w->btn); free(w->newDrop); free(w->dead_max); free(w->fullFlush);
	free(w->dead); free(w->inList); free(w->needsCollision); free(w->needsPhysUpdate); free(w->hasBox); free(w->needs0);
	free(w->forced); free(w->forcedHoldees); free(w->cLeaf); free(w->cBuddy); free(w->cAxis); free(w->cDir);
	free(w->cMutual); free(w->slider_off); free(w->sv); free(w->smin); free(w->smax); free(w->wires);
	free(w->wiresAdd); free(w->wiresRm); free(w->stamp); free(w->entList.items); free(w->crush_log.items);
	free(w);
}

/* world rules that are plain scans (main.cpp:1030-1034): crush() effects.cpp:48-60, doLava
 * :43-46, bounce() :71-92; uDead entUpdaters.c:21-23 */
static int outside(const i32 *c, i32 r) { return abs(c[0]) > r * 2 || abs(c[1]) > r || abs(c[2]) > r; }
static void worldRules(World *w, int g) {
	i32 rules = w->gamerules[g];
	for (i32 e = w->wend[g] - 1; e >= w->wbegin[g]; e--) {
		if (!w->inList[e]) continue;
		if ((rules & 1) && outside(w->center[e], 64000)) w->dead_max[e][w->flip[g]] |= 1;
		if ((rules & 2) && outside(w->center[e], 2560000)) w->dead_max[e][w->flip[g]] |= 1;
		if ((rules & 16) && w->center[e][2] > 64000) w->dead_max[e][w->flip[g]] |= 1;
		if ((rules & 32) && outside(w->center[e], 64000)) {
			if (w->type[e] & T_TERRAIN) {
				for (int i = 0; i < 3; i++) {
					i32 bound = i ? 64000 : 128000;
					if ((w->center[e][i] > bound && w->vel[e][i] > 0) || (w->center[e][i] < -bound && w->vel[e][i] < 0))
						w->vel[e][i] *= -1;
				}
			} else {
				w->dead_max[e][w->flip[g]] |= 1;
			}
		}
	}
}

/* one step of the phases given, in doWholeStep order (main.cpp:1072-1097) */
/* ---- tick handlers, doUpdates ent.cpp:823-832: the logic family of
 * modules/logic.c, door.c, seat.cpp, cursed (common.c:52), respawn.cpp without players, sign.c.
 * Run one entity after the other in gs->ents order (newest first), like the reference. ---- */
enum {
	TICK_RAND = 6, TICK_LOGIC = 11, TICK_LOGIC_DEBUG = 12, TICK_DOOR = 13, TICK_CURSED = 14, TICK_RESPAWN = 15,
	TICK_SEAT_OLD = 17, TICK_ADJUST = 19, TICK_SEAT = 23, TICK_SCOREBOARD = 27, TICK_TIMER = 28, TICK_DEMOLISH = 29, TICK_SIGN = 34,
	TICK_HELD_GUN = 24, TICK_HELD_BLINK = 30, TICK_HELD_JUMPER = 31, TICK_LOADBLOCK = 35, TICK_PUDDLEJUMP = 36, TICK_SPIDER = 33, TICK_EYE = 8, TICK_HELD_EYE = 9, TICK_HELD_VEH_EYE = 18,
	TICK_EXPLOSION = 1, TICK_PLATFORM = 3, TICK_THRUST = 10, TICK_BULLET = 25, TICK_HELD_BULLET = 26, TICK_DUST = 32, TICK_INCINERATE = 40, TICK_LEGG_OLD = 16, TICK_LEGG = 22, TICK_FACTORY = 20, TICK_FACTORY_WIRES = 21, TICK_FLAG_SPAWNER = 2, TICK_PLAYER = 7
};
static int getButton(const World *w, i32 e, int i) { return (int)(signed char)w->btn[e][i]; }
static void pushBtn(World *w, i32 e, int i) { w->btn[e][4 + i] = 1; }
static void uDead(World *w, i32 e) { w->dead_max[e][w->flip[w->world[e]]] |= 1; }

static int logic_common_input(const World *w, i32 me, i32 mode) { /* logic.c:51-61 */
	int a = getButton(w, me, 0), b = getButton(w, me, 1);
	if ((mode | 1) == 1) return (a ^ b ^ mode) != 0;
	return a && b;
}
static void logic_common_output(World *w, i32 me, int sl) { /* logic.c:63-80 */
	i32 outputs = getSlider(w, me, sl);
	if (outputs & 1)
		for (int k = 0; k < w->wires[me].num; k++) pushBtn(w, w->wires[me].items[k], 0);
	if (outputs & 2)
		for (int k = 0; k < w->wires[me].num; k++) pushBtn(w, w->wires[me].items[k], 1);
	pushBtn(w, me, 2);
}
static i32 getApproachSpeed(i32 d, i32 a) { /* modules/common.c:6-29 */
	if (a <= 0 || d <= 0) return 0;
	i32 q = d / a;
	i32 n = (((i32)sqrt((double)(1 + 8 * q))) - 1) / 2;
	i32 r = (d - n * (n + 1) / 2 * a) / (n + 1);
	return n * a + r;
}
static i32 getStoppingDist(i32 s, i32 a) { /* modules/common.c:31-37 */
	if (s <= 0 || a <= 0) return 0;
	i32 n = s / a, r = s % a;
	return n * (n + 1) / 2 * a + r * (n + 1);
}
static void door_tick(World *w, i32 me) { /* door.c:14-62 */
	int nw = w->wires[me].num;
	if (!nw) return;
	i32 pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0};
	int numActive = 0;
	for (int k = 0; k < nw; k++) {
		i32 x = w->wires[me].items[k];
		if (getButton(w, x, 2)) {
			if (!numActive) vel[0] = vel[1] = vel[2] = 0;
			numActive++;
			for (int i = 0; i < 3; i++) { pos[i] += w->center[x][i] - w->center[me][i]; vel[i] += w->vel[x][i] - w->vel[me][i]; }
		} else if (!numActive) {
			for (int i = 0; i < 3; i++) vel[i] += w->vel[x][i] - w->vel[me][i];
		}
	}
	if (!numActive) numActive = nw;
	for (int i = 0; i < 3; i++) { pos[i] /= numActive; vel[i] /= numActive; }
	i32 speed = getSlider(w, me, 0), accel = getSlider(w, me, 1);
	i32 dist = 0;
	for (int i = 0; i < 3; i++) if (abs(pos[i]) > dist) dist = abs(pos[i]);
	if (dist && getStoppingDist(speed, accel) > dist) speed = getApproachSpeed(dist, accel);
	boundVec(pos, speed, 3);
	for (int i = 0; i < 3; i++) vel[i] = pos[i] + vel[i];
	boundVec(vel, accel, 3);
	for (int i = 0; i < 3; i++) w->d_vel[me][i] += vel[i]; /* uVel */
}
/* legg.cpp:30-164 */
static i32 leggMax(i32 a, i32 b) { return (a + b + abs(a - b)) / 2; }
static i32 leggMin(i32 a, i32 b) { return (a + b - abs(a - b)) / 2; }
static i32 bound1(i32 x, i32 b) { return x > b ? b : (x < -b ? -b : x); }
static void legg_tick(World *w, i32 me, int old_style) {
	i32 h = w->holder[me];
	if (h < 0) return;
	i32 v1[3] = {0, 0, 0};
	if (old_style) {
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k], xh = w->holder[x];
			if (xh < 0) continue;
			for (int i = 0; i < 2; i++) v1[i] += w->center[x][i] - w->center[xh][i];
		}
	} else {
		i32 width = getSlider(w, me, 1);
		for (int k = 0; k < w->wires[me].num; k++)
			for (int i = 0; i < 2; i++) v1[i] += getSlider(w, w->wires[me].items[k], i) * width / 64;
	}
	int o = old_style ? 1 : 2;
	i32 counter = getSlider(w, me, 0);
	i32 height = getSlider(w, me, o), stepHeight = getSlider(w, me, o + 1), tread = getSlider(w, me, o + 2);
	i32 s = getSlider(w, me, o + 3), a = getSlider(w, me, o + 4), time = getSlider(w, me, o + 5), v_time = getSlider(w, me, o + 6);
	int hasInput = v1[0] || v1[1];
	i32 half = time + v_time;
	if (hasInput && half) {
		counter++;
		while (counter >= 2 * half) counter -= 2 * half;
		uSlider(w, me, 0, counter);
		if (counter >= time) {
			tread = 0;
			if (counter < half + time) height -= stepHeight;
		}
		if (counter < half) { v1[0] *= -1; v1[1] *= -1; }
	} else {
		tread = 0;
	}
	v1[2] = height;
	i32 vel[3], acc[3];
	for (int i = 0; i < 3; i++) v1[i] += w->center[h][i] - w->center[me][i];
	for (int i = 0; i < 3; i++) vel[i] = w->vel[me][i] - w->vel[h][i];
	i32 v_s = s;
	i32 dist = leggMax(abs(v1[0]), abs(v1[1]));
	if (dist) s = leggMin(s, getApproachSpeed(dist, a));
	boundVec(v1, s, 2);
	for (int i = 0; i < 2; i++) acc[i] = v1[i] - vel[i];
	boundVec(acc, a, 2);
	i32 new_speed = leggMax(abs(vel[0] + acc[0]), abs(vel[1] + acc[1]));
	if (s) v1[2] += tread * new_speed / s;
	if (v1[2]) v_s = leggMin(v_s, getApproachSpeed(abs(v1[2]), a));
	v1[2] = bound1(v1[2], v_s);
	acc[2] = bound1(v1[2] - vel[2], a);
	for (int i = 0; i < 3; i++) w->d_vel[me][i] += acc[i];
}

/* player.c:229-443 without what creates entities (stackem blocks, platforms, baubles: the host's) */
static int equipHands(World *w, i32 e) {
	uint32_t f = w->type[e] & (T_EQUIP | T_EQUIP_SM);
	if (!f) return 0;
	if (f == (T_EQUIP | T_EQUIP_SM)) return 3;
	return 1 + (getSlider(w, e, 0) != 0);
}
static void player_tick(World *w, int g, i32 me) {
	int grounded = getSlider(w, me, 7) > 0;
	uSlider(w, me, 7, 0);
	i32 axis[2] = {getSlider(w, me, 0), getSlider(w, me, 1)}, sl[2] = {getSlider(w, me, 5), getSlider(w, me, 6)}, speed, vel[3];
	if (!grounded) { boundVec(sl, 256, 2); speed = 256; }
	else { speed = abs(sl[0]) > abs(sl[1]) ? abs(sl[0]) : abs(sl[1]); if (speed < 256) speed = 256; }
	vel[2] = getButton(w, me, 0) && grounded ? -384 : 0;
	for (int i = 0; i < 2; i++) vel[i] = speed * axis[i] / 64 - sl[i];
	boundVec(vel, grounded ? 56 : 16, 2);
	for (int i = 0; i < 2; i++) uSlider(w, me, 5 + i, sl[i] + vel[i]);
	for (int i = 0; i < 3; i++) w->d_vel[me][i] += vel[i];
	i32 charge = getSlider(w, me, 8), cooldown = getSlider(w, me, 9);
	int fire = getButton(w, me, 2), fire2 = getButton(w, me, 3), utility = getButton(w, me, 1);
	i32 other = getSlider(w, me, 12);
	int altfire = (other & 1) != 0, altfire2 = (other & 2) != 0;
	if (getSlider(w, me, 11)) {
		for (int mode = 0; mode < 2; mode++) {
			if (!(mode ? altfire2 : altfire)) continue;
			for (i32 h = w->holdee[me]; h >= 0; h = w->LLn[h])
				if ((w->type[h] & 64u) && w->wires[h].num == 1 && mode == getSlider(w, h, 0)) w->newDrop[h] = 1;
		}
		if (cooldown >= 3) {
			if (utility && getButton(w, me, 0)) {
				for (i32 h = w->holdee[me]; h >= 0; h = w->LLn[h])
					if ((w->type[h] & 64u) && w->wires[h].num == 1) pushBtn(w, h, 2);
				cooldown = 0;
			} else if (fire || fire2) cooldown = 0;
		}
	} else {
		int hands = 0, drops = 0;
		for (i32 h = w->holdee[me]; h >= 0; h = w->LLn[h]) { int eh = equipHands(w, h); drops |= hands & eh; hands |= eh; }
		i32 ep = getSlider(w, me, 10);
		ep &= ~1;
		if (altfire) { if (!(ep & 2)) { ep |= 2; drops |= 1; } } else ep &= ~2;
		if (altfire2) { if (!(ep & 4)) { ep |= 4; drops |= 2; } } else ep &= ~4;
		if (fire) { if (hands & 1) ep |= 8; } else ep &= ~8;
		if (fire2) { if (hands & 2) ep |= 16; } else ep &= ~16;
		uSlider(w, me, 10, ep);
		i32 look[3] = {getSlider(w, me, 2), getSlider(w, me, 3), getSlider(w, me, 4)};
		if (drops) {
			if (drops & hands) {
				for (i32 h = w->holdee[me]; h >= 0; h = w->LLn[h])
					if (drops & equipHands(w, h)) {
						for (int i = 0; i < 3; i++) { w->d_vel[h][i] += 200 * look[i] / 64; w->d_center[h][i] += 200 * look[i] / 64; }
						w->newDrop[h] = 1;
					}
			} else {
				for (i32 h = w->holdee[me]; h >= 0; h = w->LLn[h])
					if (w->type[h] & (T_EQUIP | T_EQUIP_SM)) uSlider(w, h, 0, drops - 1);
			}
		}
		if (hands == 3) return;
		if (cooldown >= 5 && charge >= 60) {
			if (fire && !(hands & 1) && !(ep & 8) && !(w->gamerules[g] & (1 << 7))) { charge -= 60; cooldown = 0; }
			else if (charge >= 180 && fire2 && !(hands & 2) && !(ep & 16) && !(w->gamerules[g] & (1 << 8))) { charge -= 180; cooldown = 0; }
		}
	}
	if (charge < 180) uSlider(w, me, 8, charge + 2);
	if (cooldown < 5) uSlider(w, me, 9, cooldown + 1);
}

static void tickOne(World *w, int g, i32 me) {
	int id = w->h[me][w->holder[me] >= 0 ? BT_H_TICKHELD : BT_H_TICK];
	switch (id) {
	case TICK_LOGIC: case TICK_LOGIC_DEBUG: /* logic.c:131-149 */
		if (logic_common_input(w, me, getSlider(w, me, 0))) logic_common_output(w, me, 1);
		break;
	case TICK_TIMER: { /* logic.c:82-99 */
		if (getButton(w, me, 0) ^ getButton(w, me, 1)) { uSlider(w, me, 2, getSlider(w, me, 0)); break; }
		i32 timer = getSlider(w, me, 2);
		if (timer) { uSlider(w, me, 2, timer - 1); break; }
		logic_common_output(w, me, 1);
		break;
	}
	case TICK_ADJUST: /* logic.c:101-109 */
		if (getButton(w, me, 0) ^ getButton(w, me, 1)) {
			i32 sl = getSlider(w, me, 0), value = getSlider(w, me, 1);
			for (int k = 0; k < w->wires[me].num; k++) uSlider(w, w->wires[me].items[k], sl, value);
		}
		break;
	case TICK_RAND: { /* randomazzo_tick logic.c:151-168 */
		if (!logic_common_input(w, me, getSlider(w, me, 0))) break;
		int options = 0;
		for (int k = 0; k < w->wires[me].num; k++) options += !getButton(w, w->wires[me].items[k], 0);
		if (!options) break;
		w->rnd[g] = (i32)(((int64_t)w->rnd[g] * 48271) % 2147483647); /* stepRand random.cpp:7-9 */
		int roll = w->rnd[g] % options;
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k];
			if (getButton(w, x, 0)) continue;
			if (roll--) continue;
			i32 outputs = getSlider(w, me, 1);
			if (outputs & 1) pushBtn(w, x, 0);
			if (outputs & 2) pushBtn(w, x, 1);
			pushBtn(w, me, 2);
			break;
		}
		break;
	}
	case TICK_DEMOLISH: if (getButton(w, me, 0)) uDead(w, me); break; /* logic.c:170-172 */
	case TICK_CURSED: uDead(w, me); break;
	case TICK_DOOR: door_tick(w, me); break;
	case TICK_RESPAWN: { /* respawn.cpp:24-66 with gs->players.num == 0 */
		i32 cooldown = getSlider(w, me, 4);
		if (cooldown > 0) uSlider(w, me, 4, cooldown - 1);
		break;
	}
	case TICK_SEAT_OLD: case TICK_SEAT: /* seat.cpp:38-60 */
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k];
			if (getButton(w, x, 0)) { w->newDrop[x] = 1; continue; }
			if (id == TICK_SEAT_OLD) continue;
			for (int i = 0; i < 5; i++) uSlider(w, me, i, getSlider(w, x, i));
			for (int i = 0; i < 4; i++)
				if (getButton(w, x, i)) pushBtn(w, me, i);
		}
		break;
	case TICK_EXPLOSION: { /* explosion.cpp:12-19 */
		i32 timer = getSlider(w, me, 0);
		if (timer < 3) uSlider(w, me, 0, timer + 1); else uDead(w, me);
		break;
	}
	case TICK_PLATFORM: /* platform.c:15-23 */
		w->h[me][BT_H_TICK] = 0;
		w->h[me][BT_H_CRUSH] = 0;
		for (int k = 0; k < w->wires[me].num; k++) il_add(&w->wiresRm[me], w->wires[me].items[k]); /* uUnwire */
		break;
	case TICK_THRUST: { /* thrust.c:19-37 */
		if (!w->wires[me].num || getButton(w, me, 0)) break;
		i32 divisor = getSlider(w, me, 0);
		if (!divisor) break;
		i32 off[3] = {0, 0, 0};
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k];
			if (!getButton(w, x, 2)) continue;
			for (int i = 0; i < 3; i++) off[i] += w->center[me][i] - w->center[x][i];
		}
		for (int i = 0; i < 3; i++) w->d_vel[w->holdRoot[me]][i] += off[i] * 64 / divisor;
		break;
	}
	case TICK_BULLET: { /* gun.c:116-123 */
		i32 ttl = getSlider(w, me, 0);
		if (ttl <= 0) uDead(w, me); else uSlider(w, me, 0, ttl - 1);
		break;
	}
	case TICK_HELD_BULLET: { /* gun.c:125-152 */
		i32 ttl = getSlider(w, me, 0);
		if (ttl <= -15 || ttl > 0) {
			i32 h = w->holder[me];
			uint32_t t = w->type[h];
			int hits = (t & T_DEBRIS) ? 1 : ((t & T_TERRAIN) ? ((t & T_HEAVY) ? 15 : 5) : 3);
			for (i32 c = w->holdee[h]; c >= 0; c = w->LLn[c]) {
				if (w->h[c][BT_H_TICKHELD] != TICK_HELD_BULLET) continue;
				if (!--hits) { uDead(w, h); break; }
			}
			uDead(w, me);
		} else {
			uSlider(w, me, 0, ttl - 1);
		}
		break;
	}
	case TICK_DUST: { /* dust.c:9-21 */
		i32 time = getSlider(w, me, 3);
		if (time <= -8) { uDead(w, me); break; }
		if (time > 0) for (int i = 0; i < 3; i++) w->d_vel[me][i] += getSlider(w, me, i);
		uSlider(w, me, 3, time - 1);
		break;
	}
	case TICK_INCINERATE: { /* incinerate.cpp:15-32 */
		if (getButton(w, me, 0) ^ getButton(w, me, 1)) break;
		i32 range = getSlider(w, me, 0);
		if (range <= 0) break;
		for (i32 e = w->wbegin[g]; e < w->wend[g]; e++) {
			if (!w->inList[e]) continue;
			for (int i = 0; i < 3; i++)
				if (abs(w->center[e][i] - w->center[me][i]) - w->radius[e][i] >= range) { uDead(w, e); break; }
		}
		break;
	}
	case TICK_EYE: /* eyes.cpp:14-22 */
		w->fullFlush[me] = 1;
		w->ntype[me] = w->type[me] | T_DEBRIS;
		w->ncollide[me] = w->collide[me] | T_TERRAIN | T_OBSTACLE | T_DEBRIS;
		w->h[me][BT_H_TICK] = 0;
		break;
	case TICK_HELD_EYE: { /* eyes.cpp:24-47 */
		i32 p = w->holder[me];
		if (p < 0) break;
		i32 vec[3] = {getSlider(w, p, 2), getSlider(w, p, 3), getSlider(w, p, 4)};
		i32 component = abs(vec[2]), r = w->radius[p][2] - 50;
		for (int i = 0; i < 2; i++) {
			i32 i_comp = abs(vec[i]), i_r = w->radius[p][i] - 50;
			if (i_comp * r > component * i_r) { r = i_r; component = i_comp; }
		}
		if (!component) break;
		for (int i = 0; i < 3; i++)
			w->d_vel[me][i] += r * vec[i] / component - (w->center[me][i] - w->center[p][i]) - (w->vel[me][i] - w->vel[p][i]);
		break;
	}
	case TICK_HELD_VEH_EYE: { /* eyes.cpp:49-66 */
		i32 p = w->holder[me];
		if (p < 0) break;
		i32 vec[3] = {0, 0, 0};
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k];
			for (int j = 0; j < w->wires[x].num; j++)
				for (int i = 0; i < 2; i++) vec[i] += getSlider(w, w->wires[x].items[j], i);
		}
		for (int i = 0; i < 3; i++) w->d_center[me][i] += vec[i] * w->radius[p][i] / 64 - (w->center[me][i] - w->center[p][i]);
		break;
	}
	case TICK_FACTORY: { /* factory.cpp:63-73; spawn() (entity creation) is the host's */
		if (getButton(w, me, 0) || getButton(w, me, 1)) break;
		if (!getButton(w, me, 2) || !w->wires[me].num) break;
		if (w->wires[me].num == 1) { /* finalize :30-61 */
			i32 e = w->wires[me].items[0];
			for (int i = 0; i < 3; i++) w->d_vel[e][i] += getSlider(w, me, 1 + i);
			uint32_t misc = (uint32_t)getSlider(w, me, 7);
			if (misc & 0x80000000u) {
				i32 h = w->holder[me];
				if (h >= 0 && w->wires[h].num == 1) uPickup(w, w->wires[h].items[0], e, w->holdFlags[me]);
			}
			for (int i = 0; i < 4; i++)
				if (misc & (1u << i)) pushBtn(w, e, i);
			w->fullFlush[e] = 1;
			w->ntype[e] = (uint32_t)getSlider(w, me, 4);
			w->ncollide[e] = (uint32_t)getSlider(w, me, 5);
			w->h[e][BT_H_TICK] = getSlider(w, me, 9) & 0xff; w->h[e][BT_H_TICKHELD] = getSlider(w, me, 10) & 0xff;
			w->h[e][BT_H_CRUSH] = getSlider(w, me, 11) & 0xff; w->h[e][BT_H_PUSH] = getSlider(w, me, 12) & 0xff;
			w->h[e][BT_H_PUSHED] = getSlider(w, me, 13) & 0xff; w->h[e][BT_H_ONFUMBLE] = getSlider(w, me, 14) & 0xff;
			w->h[e][BT_H_ONFUMBLED] = getSlider(w, me, 15) & 0xff; w->h[e][BT_H_ONPICKUP] = getSlider(w, me, 16) & 0xff;
			w->h[e][BT_H_ONPICKEDUP] = getSlider(w, me, 17) & 0xff;
		}
		for (int k = 0; k < w->wires[me].num; k++) il_add(&w->wiresRm[me], w->wires[me].items[k]);
		break;
	}
	case TICK_FACTORY_WIRES: { /* factory.cpp:75-91 */
		if (getButton(w, me, 0) || getButton(w, me, 1)) pushBtn(w, me, 2);
		if (!getButton(w, me, 2)) break;
		i32 h = w->holder[me];
		if (h < 0 || w->wires[h].num != 1) break;
		i32 e = w->wires[h].items[0];
		for (int k = 0; k < w->wires[me].num; k++) {
			i32 x = w->wires[me].items[k];
			for (int j = 0; j < w->wires[x].num; j++) { w->fullFlush[e] = 1; il_add(&w->wiresAdd[e], w->wires[x].items[j]); }
		}
		break;
	}
	case TICK_FLAG_SPAWNER: { /* flag.c:36-45; mkFlag (entity creation) is the host's */
		i32 cooldown = getSlider(w, me, 0);
		if (cooldown == 0) break;
		uSlider(w, me, 0, cooldown - 1);
		break;
	}
	case TICK_PLAYER: player_tick(w, g, me); break;
	case TICK_LEGG: legg_tick(w, me, 0); break;
	case TICK_LEGG_OLD: legg_tick(w, me, 1); break;
	case TICK_SPIDER: { /* spider.cpp:58-122 */
		i32 h = w->holder[me];
		if (h < 0) break;
		i32 height_top = getSlider(w, me, 0), height_bottom = getSlider(w, me, 1), width = getSlider(w, me, 2);
		i32 t_top = getSlider(w, me, 3), t_bottom = getSlider(w, me, 4), counter = getSlider(w, me, 5);
		i32 spring = getSlider(w, me, 6), drag = getSlider(w, me, 7), a_max = getSlider(w, me, 8), ratio_legs = getSlider(w, me, 9);
		if (spring <= 0) spring = 1;
		if (ratio_legs <= 0) ratio_legs = 1;
		i32 v1[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, a[3];
		for (int k = 0; k < w->wires[me].num; k++)
			for (int i = 0; i < 2; i++) v1[i] += getSlider(w, w->wires[me].items[k], i) * width / 64;
		int hasInput = v1[0] || v1[1];
		v1[2] = height_bottom;
		i32 t_cycle = t_top + t_bottom;
		if (hasInput && t_cycle) {
			counter++;
			while (counter >= t_cycle) counter -= t_cycle;
			uSlider(w, me, 5, counter);
			i32 numer, denom;
			if (counter >= t_bottom) { v1[2] = height_top; denom = t_top; numer = counter - t_bottom; }
			else { v1[0] *= -1; v1[1] *= -1; denom = t_bottom; numer = counter; }
			for (int i = 0; i < 2; i++) vel[i] = v1[i] / denom;
			for (int i = 0; i < 2; i++) v1[i] = (2 * numer - denom) * v1[i] / denom;
		}
		for (int i = 0; i < 3; i++) v1[i] = (v1[i] - (w->center[me][i] - w->center[h][i])) * 64 / spring;
		boundVec(v1, a_max, 3);
		for (int i = 0; i < 3; i++) vel[i] -= w->vel[me][i] - w->vel[h][i];
		boundVec(vel, drag, 3);
		for (int i = 0; i < 3; i++) a[i] = (v1[i] + vel[i]) / -ratio_legs;
		for (int i = 0; i < 3; i++) w->d_vel[w->holdRoot[me]][i] += a[i];
		for (int i = 0; i < 3; i++) w->d_vel[me][i] += a[i] * -ratio_legs;
		break;
	}
	case TICK_LOADBLOCK: { /* loadblock.cpp:33-36 */
		i32 cooldown = getSlider(w, me, 1);
		if (cooldown > 0) uSlider(w, me, 1, cooldown - 1);
		break;
	}
	case TICK_PUDDLEJUMP: { /* puddlejumper.cpp:94-162; vertical thrust (dust: entity creation) is the host's */
		i32 dvel[3] = {0, 0, 16};
		int flying = !getSlider(w, me, 4);
		uSlider(w, me, 4, 0);
		i32 fuel = getSlider(w, me, 0);
		if (fuel && flying) uSlider(w, me, 0, fuel - 1);
		if (w->wires[me].num == 1) {
			i32 input = w->wires[me].items[0];
			if (fuel && getButton(w, input, 2)) dvel[2] = -8; /* getTrigger(input, 0) entGetters.c:27-29 */
			i32 axis[2] = {getSlider(w, input, 0), getSlider(w, input, 1)}; /* getAxis player.c:48-51 */
			if (axis[0] || axis[1]) {
				i32 sl[2] = {getSlider(w, me, 2), getSlider(w, me, 3)};
				int signX = 2 * (axis[0] >= 0) - 1, signY = 2 * (axis[1] >= 0) - 1;
				i32 denomX = signX * axis[0], numerX = 200 - signX * sl[0];
				i32 denomY = signY * axis[1], numerY = 200 - signY * sl[1];
				if (numerX > 20) numerX = 20; else if (numerX < 0) numerX = 0;
				if (numerY > 20) numerY = 20; else if (numerY < 0) numerY = 0;
				if (denomY && numerY * denomX <= numerX * denomY) { numerX = numerY; denomX = denomY; }
				for (int k = 0; k < 2; k++) {
					dvel[k] = numerX * axis[k] / denomX;
					uSlider(w, me, 2 + k, sl[k] + dvel[k]);
					dvel[k] += axis[k] / 16;
				}
				uSlider(w, me, 5, numerX ? 2 : 1);
			} else {
				uSlider(w, me, 5, 0);
			}
		}
		for (int k = 0; k < 3; k++) w->d_vel[me][k] += dvel[k]; /* uVel */
		break;
	}
	case TICK_SCOREBOARD: { /* scoreboard.cpp:37-90 with gs->players.num == 0: alive == 0 */
		i32 timer = getSlider(w, me, 1);
		if (getSlider(w, me, 0) != 0) { uSlider(w, me, 0, 0); timer = 75; }
		if (!timer) break;
		timer--;
		uSlider(w, me, 1, timer);
		if (timer) break;
		for (i32 e = w->wbegin[g]; e < w->wend[g]; e++) /* uErase entUpdaters.c:25-27 for everything but the board */
			if (e != me && w->inList[e]) w->dead_max[e][w->flip[g]] = 3;
		uSlider(w, me, 0, 0);
		w->n_events++; /* requestReload + showMessage: the host's */
		w->fullFlush[me] = 1; /* uMyCollideMask(me, T_TERRAIN) */
		w->ncollide[me] = T_TERRAIN;
		break;
	}
	case TICK_HELD_GUN: case TICK_HELD_BLINK: case TICK_HELD_JUMPER: {
		/* gun.c:41-88, blink.c:33-60, jumper.c:39-70 up to where the trigger fires (entity creation: the host's) */
		i32 h = w->holder[me];
		if (h < 0 || !(w->type[h] & T_INPUTS) || w->holder[h] >= 0) break;
		int trigger = getButton(w, h, 2 + (getSlider(w, me, 0) != 0)); /* getTrigger entGetters.c:27-29 */
		if (id == TICK_HELD_GUN) {
			i32 cooldown = getSlider(w, me, 2), ammo = getSlider(w, me, 1);
			if (cooldown > 1) uSlider(w, me, 2, cooldown - 1);
			else if (trigger && ammo) break;
			i32 ammoMax = getSlider(w, me, 4);
			if (ammo != ammoMax) {
				i32 reload = getSlider(w, me, 3) + 1;
				if (reload >= getSlider(w, me, 6)) { uSlider(w, me, 3, 0); uSlider(w, me, 1, ammoMax); }
				else uSlider(w, me, 3, reload);
			}
		} else if (id == TICK_HELD_BLINK) {
			i32 cooldown = getSlider(w, me, 1);
			if (cooldown > 0) uSlider(w, me, 1, cooldown - 1);
		} else {
			i32 ammo = getSlider(w, me, 1);
			if (ammo < getSlider(w, me, 4)) {
				i32 reload = 1 + getSlider(w, me, 2);
				if (reload >= getSlider(w, me, 5)) { reload = 0; ammo++; }
				uSlider(w, me, 2, reload);
			}
			i32 cooldown = getSlider(w, me, 3);
			if (cooldown > 0) uSlider(w, me, 3, cooldown - 1);
			else if (ammo && trigger) break;
			uSlider(w, me, 1, ammo);
		}
		break;
	}
	default: break; /* NIL, SIGN; other ids belong to the host */
	}
}

int bto_step(World *w, uint32_t phases) {
	for (int g = 0; g < w->n_worlds; g++) {
		if (phases & BT_PHASE_WORLD_RULES) worldRules(w, g);
		if (phases & BT_PHASE_CTRLS) /* flushCtrls ent.cpp:32-37, 817-821 */
			for (i32 i = w->wbegin[g]; i < w->wend[g]; i++)
				if (w->inList[i])
					for (int k = 0; k < 4; k++) {
						w->btn[i][k] = w->btn[i][4 + k];
						w->btn[i][4 + k] = 0;
					}
		if (phases & BT_PHASE_TICKS)
			for (i32 i = w->wend[g] - 1; i >= w->wbegin[g]; i--)
				if (w->inList[i]) tickOne(w, g, i);
		if (phases & BT_PHASE_PREP) flush1(w, g);
		if ((phases & BT_PHASE_GRAVITY) && (w->gamerules[g] & EFFECT_GRAV)) doGravity(w, g);
		if (phases & BT_PHASE_PHYSICS) doPhysics(w, g);
		if (phases & BT_PHASE_FINISH) finishStep(w, g);
	}
	if (phases & BT_PHASE_PHYSICS) w->stats[ST_TICKS]++;
	return 0;
}

static inline uint64_t mix64(uint64_t x) {
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
	x ^= x >> 27; x *= 0x94d049bb133111ebULL;
	x ^= x >> 31;
	return x;
}

/* state hash, DESIGN.md "State hash" */
void bto_hash(const World *w, uint64_t *out) {
	for (int g = 0; g < w->n_worlds; g++) {
		uint64_t sum = 0, cnt = 0;
		for (i32 e = w->wbegin[g]; e < w->wend[g]; e++) {
			if (!w->inList[e]) continue;
			i32 wd[10] = {w->center[e][0], w->center[e][1], w->center[e][2], w->vel[e][0], w->vel[e][1],
			              w->vel[e][2], (i32)w->type[e], (i32)w->collide[e],
			              w->holder[e] >= 0 && w->inList[w->holder[e]] ? w->uid[w->holder[e]] : -1, /* a dead holder = none */
			              w->holdFlags[e]};
			uint64_t h = mix64((uint64_t)(uint32_t)w->uid[e] + 0x9E3779B97F4A7C15ULL);
			for (int k = 0; k < 10; k++) h = mix64(h + (uint64_t)(uint32_t)wd[k] + 0x9E3779B97F4A7C15ULL);
			sum += h;
			cnt++;
		}
		out[g] = sum + mix64(cnt);
	}
}

void bto_count(const World *w, i32 *n, i32 *ns, i32 *nw) {
	i32 a = 0, b = 0, c = 0;
	for (i32 e = 0; e < w->n; e++)
		if (w->inList[e]) {
			a++;
			b += w->slider_off[e + 1] - w->slider_off[e];
			c += w->wires[e].num;
		}
	*n = a; *ns = b; *nw = c;
}

/* live entities, oldest first, references renumbered (what a .sav would hold) */
int bto_export(const World *w, bt_world_desc *d) {
	i32 *map = (i32 *)malloc(sizeof(i32) * (w->n + 1));
	i32 k = 0;
	for (i32 e = 0; e < w->n; e++) map[e] = w->inList[e] ? k++ : -1;
#define REF(x) ((x) >= 0 ? map[(x)] : -1)
	i32 so = 0, wo = 0;
	for (int g = 0; g < w->n_worlds; g++) {
		d->w_flip_flop[g] = w->flip[g]; d->w_gamerules[g] = w->gamerules[g]; d->w_rand[g] = w->rnd[g];
		i32 pos = 0;
		for (i32 i = w->rootHead[g]; i >= 0; i = w->LLn[i]) d->root_order[map[i]] = pos++;
	}
	for (i32 e = 0; e < w->n; e++) {
		if (!w->inList[e]) continue;
		i32 o = map[e];
		if (d->world_id) d->world_id[o] = w->world[e];
		d->uid[o] = w->uid[e];
		for (int c = 0; c < 3; c++) {
			d->center[3 * o + c] = w->center[e][c]; d->vel[3 * o + c] = w->vel[e][c];
			d->old[3 * o + c] = w->old[e][c]; d->d_center[3 * o + c] = w->d_center[e][c];
			d->d_vel[3 * o + c] = w->d_vel[e][c]; d->radius[3 * o + c] = w->radius[e][c];
		}
		d->type_mask[o] = w->type[e]; d->collide_mask[o] = w->collide[e];
		d->new_type_mask[o] = w->ntype[e]; d->new_collide_mask[o] = w->ncollide[e];
		d->friction[o] = w->friction[e];
		d->holder[o] = REF(w->holder[e]); d->hold_flags[o] = w->holdFlags[e];
		d->holdee_head[o] = REF(w->holdee[e]);
		d->holdee_next[o] = w->holder[e] >= 0 ? REF(w->LLn[e]) : -1;
		if (w->holder[e] >= 0) d->root_order[o] = -1;
		d->new_holder[o] = REF(w->newHolder[e]); d->new_hold_flags[o] = w->newHoldFlags[e];
		memcpy(d->handlers + (size_t)BT_H_NUM * o, w->h[e], BT_H_NUM);
		memcpy(d->btn + 8 * (size_t)o, w->btn[e], 8);
		uint8_t *m = d->misc + (size_t)BT_M_NUM * o;
		m[BT_M_NEWDROP] = w->newDrop[e]; m[BT_M_DEADMAX0] = w->dead_max[e][0];
		m[BT_M_DEADMAX1] = w->dead_max[e][1]; m[BT_M_FULLFLUSH] = w->fullFlush[e];
		d->slider_off[o] = so;
		for (int s = w->slider_off[e]; s < w->slider_off[e + 1]; s++, so++) {
			d->slider_v[so] = w->sv[s]; d->slider_min[so] = w->smin[s]; d->slider_max[so] = w->smax[s];
		}
		d->wire_off[o] = wo;
		for (int c = 0; c < w->wires[e].num; c++, wo++) d->wire_to[wo] = REF(w->wires[e].items[c]);
	}
#undef REF
	d->slider_off[k] = so; d->wire_off[k] = wo;
	d->n_ents = k; d->n_sliders = so; d->n_wires = wo; d->n_worlds = w->n_worlds;
	free(map);
	return 0;
}

/* selectCast raycast.cpp:16-57; indices are live oldest-first ranks (bto_export order) */
typedef struct { i32 n, d; } rat;
#define LT(a, b) ((int64_t)(a).n * (b).d < (int64_t)(b).n * (a).d)
int bto_selectcast_batch(const World *w, const i32 *origin, const i32 *look_in, i32 n_rays, i32 *winner) {
	i32 *live = (i32 *)malloc(sizeof(i32) * (w->n + 1));
	i32 *rank = (i32 *)malloc(sizeof(i32) * (w->n + 1));
	i32 nl = 0;
	for (i32 e = 0; e < w->n; e++)
		if (w->inList[e]) { rank[e] = nl; live[nl++] = e; }
	for (i32 r = 0; r < n_rays; r++) {
		i32 parent = live[origin[r]];
		const i32 *pos = w->center[parent];
		i32 holdRoot = w->holdRoot[parent];
		int g = w->world[parent];
		i32 look[3], flip[3];
		for (int i = 0; i < 3; i++) {
			look[i] = look_in[3 * r + i];
			if (look[i] >= 0) flip[i] = 1;
			else { flip[i] = -1; look[i] *= -1; }
		}
		i32 win = -1;
		rat best = {INT32_MAX, 1};
		for (i32 e = w->wend[g] - 1; e >= w->wbegin[g]; e--) { /* gs->ents order: newest first */
			if (!w->inList[e]) continue;
			if (w->holdRoot[e] == holdRoot) continue;
			rat lower = {0, 1};
			rat upper = best;
			int miss = 0;
			for (int d = 0; d < 3; d++) {
				i32 x = flip[d] * (w->center[e][d] - pos[d]);
				rat f1, f2;
				f1.n = x - w->radius[e][d];
				f2.n = x + w->radius[e][d];
				f1.d = f2.d = look[d];
				if (LT(lower, f1)) lower = f1;
				if (LT(f2, upper)) upper = f2;
				if (!LT(lower, upper)) { miss = 1; break; }
			}
			if (miss) continue;
			if (lower.n) { best = lower; win = e; }
		}
		winner[r] = win >= 0 ? rank[win] : -1;
	}
	free(live); free(rank);
	return 0;
}

/* uids the crush handlers would have run for since the last call, in invokeOnCrush order */
int bto_crush_log(World *w, i32 *out, int max) {
	int n = w->crush_log.num;
	for (int i = 0; i < n && i < max; i++) out[i] = w->crush_log.items[i];
	w->crush_log.num = 0;
	return n;
}

/* the deferred updaters of entUpdaters.c as called by host-side code; kinds = BT_REQ_* */
int bto_request(World *w, int kind, i32 me_uid, i32 other_uid, i32 a0, i32 a1, i32 a2) {
	i32 e = -1, o = -1;
	for (i32 i = 0; i < w->n; i++) {
		if (!w->inList[i]) continue;
		if (w->uid[i] == me_uid) e = i;
		if (w->uid[i] == other_uid) o = i;
	}
	if (e < 0) return -3;
	int g = w->world[e];
	switch (kind) {
	case BT_REQ_DROP: w->newDrop[e] = 1; break;
	case BT_REQ_DEAD: w->dead_max[e][w->flip[g]] |= 1; break;
	case BT_REQ_ERASE: w->dead_max[e][w->flip[g]] = 3; break;
	case BT_REQ_PICKUP: if (o < 0) return -3; uPickup(w, e, o, a0); break;
	case BT_REQ_BUTTON: if (a0 >= 0 && a0 < 4) w->btn[e][4 + a0] = 1; break;
	case BT_REQ_VEL: w->d_vel[e][0] += a0; w->d_vel[e][1] += a1; w->d_vel[e][2] += a2; break;
	case BT_REQ_CENTER: w->d_center[e][0] += a0; w->d_center[e][1] += a1; w->d_center[e][2] += a2; break;
	case BT_REQ_SLIDER: uSlider(w, e, a0, a1); break;
	case BT_REQ_TYPEMASK: w->fullFlush[e] = 1; w->ntype[e] = (uint32_t)a0; break;
	case BT_REQ_COLLIDEMASK: w->fullFlush[e] = 1; w->ncollide[e] = (uint32_t)a0; break;
	default: return -3;
	}
	return 0;
}

int64_t bto_order_ties(const World *w) { return w->order_ties; }

void bto_stats(const World *w, int64_t *out) { memcpy(out, w->stats, sizeof(w->stats)); }

double bto_time_steps(World *w, i32 n, int64_t *entity_ticks) {
	struct timespec a, b;
	int64_t et = 0;
	clock_gettime(CLOCK_MONOTONIC, &a);
	for (int i = 0; i < n; i++) {
		bto_step(w, BT_PHASE_ALL);
		for (i32 e = 0; e < w->n; e++) et += w->inList[e];
	}
	clock_gettime(CLOCK_MONOTONIC, &b);
	if (entity_ticks) *entity_ticks = et;
	return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
