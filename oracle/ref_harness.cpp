/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Headless driver around the UNMODIFIED sboerwinkle/bittoss simulation core.
 * oracle/Makefile compiles this file together with the reference's own
 * sources where they lie under $(REF) (default /root/reference) into
 * oracle/_ref/libbtref.so.  Nothing of the reference is copied into this repo.
 *
 * What it restates (because main.cpp needs GLFW/GL and cannot be built):
 *   - the init order of main()                      main.cpp:1690-1761
 *   - doWholeStep() minus networking/commands       main.cpp:1019-1098
 *   - resetPlayer / mkHeroes / cleanupDeadHeroes    main.cpp:308-317, 609-631
 *   - doInputs                                      main.cpp:337-350
 * and what it adds: export of a gamestate into the plain bt_world_desc arrays
 * of include/bittoss_b200.h, a state hash, synthetic world builders, and a
 * split step (everything before doPhysics / doPhysics itself) for per-tick
 * teacher-forced parity checks of the CUDA path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <unordered_map>
#include <vector>

#include "ent.h"
#include "main.h"
#include "graphics.h"
#include "handlerRegistrar.h"
#include "entFuncs.h"
#include "entUpdaters.h"
#include "entGetters.h"
#include "serialize.h"
#include "effects.h"
#include "file.h"
#include "gamestring.h"
#include "colors.h"
#include "edit.h"
#include "cdump.h"
#include "modules.h"
#include "map.h"
#include "raycast.h"
#include "random.h"
#include "modules/player.h"

#include "../include/bittoss_b200.h"

/* ---- symbols main.cpp / graphics.cpp would have provided (main.h:11-17, graphics.h:38-46) ---- */
volatile char globalRunning = 1;
volatile char dl_srm_ready = 0;
gamestate *rootState = NULL;
void showMessage(gamestate const *const gs, char const *const msg) {}
void rect(int32_t *p, int32_t *radius, float r, float g, float b) {}
void drawTag(const char *str, int32_t *pos, double scale, float *color) {}
void setupStipple() {}
void setupTags() {}

struct RefWorld {
	gamestate *gs;
	std::unordered_map<ent *, int32_t> uid;
	int32_t next_uid;
	int32_t tick;
	int reload_pending;
	std::vector<char> sav; /* bytes of the level this world was loaded from */
	int spacetrash_on;
	uint32_t input_key;
	int32_t pdata[16];
};

static std::vector<RefWorld *> worlds;
static RefWorld *cur = NULL;
static int inited = 0;
static char data_dir[4096];

void requestReload(gamestate const *const gs) {
	for (RefWorld *w : worlds)
		if (w && w->gs == gs) w->reload_pending = 1;
}
void requestLoad(gamestate const *gs, int playerIx, int gameStrIx) {}

/* ---- uid bookkeeping: a stable creation-order id per live ent ---- */
static void assign_new_uids(RefWorld *w) {
	/* new ents are always prepended to gs->ents (ent.cpp:47-52) */
	std::vector<ent *> fresh;
	for (ent *e = w->gs->ents; e; e = e->ll.n) {
		if (w->uid.count(e)) break;
		fresh.push_back(e);
	}
	for (int i = (int)fresh.size() - 1; i >= 0; i--) w->uid[fresh[i]] = w->next_uid++;
}
static void forget_dead(RefWorld *w) {
	for (ent *e = w->gs->deadTail; e; e = e->ll.p) w->uid.erase(e);
}

extern "C" {

int btref_init(const char *dir_with_data) {
	if (inited) return 0;
	/* the reference reports through puts()/printf(); line-buffer C stdout so a caller that
	 * temporarily points fd 1 elsewhere (oracle/pybind.py stdout_to_stderr) really captures it */
	setvbuf(stdout, NULL, _IOLBF, 0);
	char prev[4096];
	if (!getcwd(prev, sizeof prev)) return -1;
	if (chdir(dir_with_data)) { perror("btref_init chdir"); return -1; }
	/* main.cpp:1690-1761 */
	velbox_init();
	init_registrar();
	init_entFuncs();
	cdump_init();
	file_init(); /* chdir("data") */
	gamestring_init();
	colors_init();
	edit_init();
	modules_init();
	ent_init();
	if (!getcwd(data_dir, sizeof data_dir)) return -1;
	/* the reference resolves level names relative to data/; do not leave the
	 * calling process parked there */
	if (chdir(prev)) return -1;
	inited = 1;
	return 0;
}

void btref_flush(void) { fflush(stdout); fflush(stderr); }

int btref_new_world(void) {
	RefWorld *w = new RefWorld();
	w->gs = mkGamestate();
	w->next_uid = 0;
	w->tick = 0;
	w->reload_pending = 0;
	w->spacetrash_on = 0;
	w->input_key = 0;
	memset(w->pdata, 0, sizeof(w->pdata));
	worlds.push_back(w);
	cur = w;
	rootState = w->gs;
	return (int)worlds.size() - 1;
}

int btref_select(int id) {
	if (id < 0 || id >= (int)worlds.size() || !worlds[id]) return -1;
	cur = worlds[id];
	rootState = cur->gs;
	return 0;
}

void btref_free_world(int id) {
	if (id < 0 || id >= (int)worlds.size() || !worlds[id]) return;
	RefWorld *w = worlds[id];
	doCleanup(w->gs);
	free(w->gs);
	if (cur == w) { cur = NULL; rootState = NULL; }
	delete w;
	worlds[id] = NULL;
}

/* resetPlayer main.cpp:308-317 */
static void reset_player(RefWorld *w, int i) {
	player *p = &w->gs->players[i];
	p->reviveCounter = 100000;
	p->data = w->pdata[i & 15];
	p->entity = NULL;
	p->color = 0xCCCCFF;
	snprintf(p->name, NAME_BUF_LEN, "p%d", i);
}

int btref_load_sav(const char *name, int n_players, const int32_t *pdata, int32_t rand_seed) {
	RefWorld *w = cur;
	list<char> data;
	data.init();
	char prev[4096];
	if (!getcwd(prev, sizeof prev) || chdir(data_dir)) return -1;
	char bad = readFile(name, &data);
	if (chdir(prev)) return -1;
	if (bad) { data.destroy(); return -1; }
	w->sav.assign(data.items, data.items + data.num);
	for (int i = 0; i < 16; i++) w->pdata[i] = (pdata && i < n_players) ? pdata[i] : 0;
	w->gs->players.setMaxUp(n_players > 0 ? n_players : 1);
	w->gs->players.num = n_players;
	for (int i = 0; i < n_players; i++) reset_player(w, i);
	list<const char> const fake = {.items = data.items, .num = data.num, .max = data.num};
	deserialize(w->gs, &fake, 0);
	w->gs->rand = rand_seed;
	data.destroy();
	assign_new_uids(w);
	return 0;
}

/* the reference's own writer (serialize.cpp:187-212) on the current world: bytes into `out`, returns
 * their number (also when it exceeds `cap`: call again with a bigger buffer) */
int btref_serialize(char *out, int cap) {
	list<char> data;
	data.init();
	serialize(cur->gs, &data);
	int n = data.num;
	if (n <= cap) memcpy(out, data.items, n);
	data.destroy();
	return n;
}

void btref_set_options(int spacetrash_on, uint32_t input_key) {
	cur->spacetrash_on = spacetrash_on;
	cur->input_key = input_key;
}

/* ---- synthetic box-stack world, BASELINE.json configs[2]; same recipe as
 *      bt_synth_boxstack_fill (bittoss_b200/csrc/bt_synth.cpp) but built through the
 *      reference's own initEnt (entFuncs.c:88-111) ---- */
int btref_make_boxstack(int32_t n_cubes, int32_t H, uint32_t seed, int32_t push_every, int32_t push_max) {
	RefWorld *w = cur;
	gamestate *gs = w->gs;
	int32_t n_cols = (n_cubes + H - 1) / H;
	int32_t G = 1;
	while ((int64_t)G * G < n_cols) G++;
	int32_t span = G * 1000;
	int32_t T = (span + 6400 + 6399) / 6400;
	const int32_t zero[3] = {0, 0, 0};
	const int32_t tile_r[3] = {3200, 3200, 500};
	for (int tx = 0; tx < T; tx++)
		for (int ty = 0; ty < T; ty++) {
			int32_t c[3] = {tx * 6400 - (T * 6400) / 2 + 3200, ty * 6400 - (T * 6400) / 2 + 3200, 600};
			ent *e = initEnt(gs, NULL, c, zero, tile_r, 0, T_TERRAIN | T_HEAVY | T_WEIGHTLESS, 0);
			e->whoMoves = whoMovesHandlers.get(WHOMOVES_ME);
		}
	const int32_t cube_r[3] = {400, 400, 400};
	uint32_t st = seed;
	int32_t made = 0;
	for (int32_t col = 0; col < n_cols; col++) {
		int32_t gx = col / G, gy = col % G;
		int32_t cx = gx * 1000 - span / 2 + 500, cy = gy * 1000 - span / 2 + 500;
		for (int32_t k = 0; k < H && made < n_cubes; k++, made++) {
			int32_t jx = (int32_t)(splitmix32(&st) % 101u) - 50;
			int32_t jy = (int32_t)(splitmix32(&st) % 101u) - 50;
			int32_t jz = (int32_t)(splitmix32(&st) % 64u);
			int32_t c[3] = {cx + jx, cy + jy, 100 - 400 - 900 * k - jz};
			int32_t v[3] = {0, 0, 0};
			if (k == H - 1 && push_every > 0 && col % push_every == 0) {
				v[0] = (int32_t)(splitmix32(&st) % (uint32_t)(2 * push_max + 1)) - push_max;
				v[1] = (int32_t)(splitmix32(&st) % (uint32_t)(2 * push_max + 1)) - push_max;
			}
			ent *e = initEnt(gs, NULL, c, v, cube_r, 0, T_OBSTACLE | T_HEAVY, T_TERRAIN | T_OBSTACLE);
			e->whoMoves = whoMovesHandlers.get(WHOMOVES_WOOD);
		}
	}
	gs->gamerules = EFFECT_GRAV;
	gs->rand = 12345;
	assign_new_uids(w);
	return w->next_uid;
}

/* A world from plain arrays, built through the reference's own initEnt / catalogs / pickupNoHandlers the way its
 * loader does (serialize.cpp:277-334): entity i of the desc becomes the i-th oldest `ent`.  Used to hand a state the
 * CUDA path has reached (bt_world_download) to the reference's selectCast without replaying the ticks on the CPU. */
int btref_build_from_desc(const bt_world_desc *d) {
	RefWorld *w = cur;
	gamestate *gs = w->gs;
	std::vector<ent *> made((size_t)d->n_ents);
	for (int i = 0; i < d->n_ents; i++) {
		const int ns = d->slider_off[i + 1] - d->slider_off[i];
		ent *e = initEnt(gs, NULL, d->center + 3 * i, d->vel + 3 * i, d->radius + 3 * i, ns, (int32_t)d->type_mask[i], (int32_t)d->collide_mask[i]);
		const uint8_t *h = d->handlers + (size_t)BT_H_NUM * i;
		e->whoMoves = whoMovesHandlers.get(h[BT_H_WHOMOVES]);
		e->tick = h[BT_H_TICK] ? tickHandlers.get(h[BT_H_TICK]) : NULL;
		e->tickHeld = h[BT_H_TICKHELD] ? tickHandlers.get(h[BT_H_TICKHELD]) : NULL;
		e->crush = h[BT_H_CRUSH] ? crushHandlers.get(h[BT_H_CRUSH]) : NULL;
		e->push = pushHandlers.get(h[BT_H_PUSH]);
		e->pushed = h[BT_H_PUSHED] ? pushedHandlers.get(h[BT_H_PUSHED]) : NULL;
		e->onFumble = entPairHandlers.get(h[BT_H_ONFUMBLE]); e->onFumbled = entPairHandlers.get(h[BT_H_ONFUMBLED]);
		e->onPickUp = entPairHandlers.get(h[BT_H_ONPICKUP]); e->onPickedUp = entPairHandlers.get(h[BT_H_ONPICKEDUP]);
		e->friction = d->friction[i];
		for (int k = 0; k < 3; k++) e->old[k] = d->old[3 * i + k];
		for (int k = 0; k < ns; k++) e->sliders[k].v = d->slider_v[d->slider_off[i] + k];
		for (int k = 0; k < 4; k++) { e->ctrl.btns[k].v = (char)d->btn[8 * i + k]; e->ctrl.btns[k].v2 = (char)d->btn[8 * i + 4 + k]; }
		made[i] = e;
	}
	for (int i = 0; i < d->n_ents; i++) {
		for (int k = d->wire_off[i]; k < d->wire_off[i + 1]; k++) made[i]->wires.add(made[d->wire_to[k]]);
		if (d->holder[i] >= 0) pickupNoHandlers(gs, made[d->holder[i]], made[i], d->hold_flags[i]);
	}
	gs->flipFlop_death = (byte)d->w_flip_flop[0];
	gs->gamerules = d->w_gamerules[0];
	gs->rand = d->w_rand[0];
	assign_new_uids(w);
	return d->n_ents;
}

/* one entity through the reference's own initEnt (entFuncs.c:88-111); whoMoves by catalog id */
int btref_add_ent(const int32_t *c, const int32_t *v, const int32_t *r, uint32_t typeMask, uint32_t collideMask,
                  int whoMovesId) {
	ent *e = initEnt(cur->gs, NULL, c, v, r, 0, (int32_t)typeMask, (int32_t)collideMask);
	e->whoMoves = whoMovesHandlers.get(whoMovesId);
	assign_new_uids(cur);
	return cur->next_uid - 1;
}

/* ---- "handler lab" setters: build small worlds that exercise one collision-path handler.
 * They only edit the INPUT world through the reference's own catalogs and updaters. ---- */
static ent *by_uid(int32_t uid) {
	for (auto &kv : cur->uid)
		if (kv.second == uid) return kv.first;
	return NULL;
}

/* crush handlers create entities on the host (explode, flag.c:17-19); the lab replaces them by a
 * recorder so that the ORDER of invokeOnCrush (ent.cpp:464-469) can be compared.  An entity whose
 * crush pointer is the recorder exports as catalog id `lab_crush_id[e]`. */
static std::vector<int32_t> lab_crush_log;
static std::unordered_map<ent *, int> lab_crush_id;
static void lab_crush_recorder(gamestate *gs, ent *me) {
	(void)gs;
	lab_crush_log.push_back(cur->uid.count(me) ? cur->uid[me] : -1);
}

/* 1: btref_lab_set_handlers installs the catalog's own crush handler instead of the recorder (free-running
 * tests want flag_crush / platform_crush to really run: they create entities and write sliders) */
static int lab_real_crush = 0;
void btref_lab_real_crush(int on) { lab_real_crush = on; }

/* ids < 0 leave the slot alone.  pair = {onFumble, onFumbled, onPickUp, onPickedUp} */
int btref_lab_set_handlers(int32_t uid, int push, int pushed, int crush, const int32_t *pair, int tick, int tickHeld) {
	ent *e = by_uid(uid);
	if (!e) return -1;
	if (push >= 0) e->push = pushHandlers.get(push);
	if (pushed >= 0) e->pushed = pushed ? pushedHandlers.get(pushed) : NULL;
	if (crush >= 0) {
		if (crush && lab_real_crush) e->crush = crushHandlers.get(crush);
		else if (crush) { e->crush = lab_crush_recorder; lab_crush_id[e] = crush; }
		else e->crush = NULL;
	}
	if (pair) {
		entPair_t *slots[4] = {&e->onFumble, &e->onFumbled, &e->onPickUp, &e->onPickedUp};
		for (int k = 0; k < 4; k++)
			if (pair[k] >= 0) *slots[k] = entPairHandlers.get(pair[k]);
	}
	if (tick >= 0) e->tick = tick ? tickHandlers.get(tick) : NULL;
	if (tickHeld >= 0) e->tickHeld = tickHeld ? tickHandlers.get(tickHeld) : NULL;
	return 0;
}

/* (re)allocate the slider block like initEnt does (entFuncs.c:88-111) and set the values */
int btref_lab_set_sliders(int32_t uid, int n, const int32_t *v) {
	ent *e = by_uid(uid);
	if (!e) return -1;
	e->sliders = (slider *)realloc(e->sliders, sizeof(slider) * (n > 0 ? n : 1));
	e->numSliders = n;
	for (int i = 0; i < n; i++) { e->sliders[i].v = v[i]; e->sliders[i].min = INT32_MAX; e->sliders[i].max = INT32_MIN; }
	return 0;
}

int btref_lab_add_wire(int32_t from, int32_t to) {
	ent *a = by_uid(from), *b = by_uid(to);
	if (!a || !b) return -1;
	a->wires.add(b);
	return 0;
}

/* ctrl.btns[ix].v, what getButton reads (entGetters.c:23); physics-only stepping never flushes it */
int btref_lab_set_button(int32_t uid, int ix, int v) {
	ent *e = by_uid(uid);
	if (!e || ix < 0 || ix > 3) return -1;
	e->ctrl.btns[ix].v = (char)v;
	return 0;
}

/* deferred requests through the reference's own updaters (entUpdaters.c) */
int btref_lab_request(int32_t uid, int what, int32_t other, int32_t arg) {
	ent *e = by_uid(uid);
	if (!e) return -1;
	if (what == 0) uDrop(cur->gs, e);
	else if (what == 1) uDead(cur->gs, e);
	else if (what == 2) { ent *o = by_uid(other); if (!o) return -1; uPickup(cur->gs, e, o, arg); }
	else if (what == 3) pushBtn(e, arg);
	else return -1;
	return 0;
}

/* the remaining deferred updaters, numbered like BT_REQ_* of include/bittoss_b200.h */
int btref_lab_request3(int32_t uid, int kind, int32_t a0, int32_t a1, int32_t a2) {
	ent *e = by_uid(uid);
	if (!e) return -1;
	int32_t v[3] = {a0, a1, a2};
	if (kind == BT_REQ_VEL) uVel(e, v);
	else if (kind == BT_REQ_CENTER) uCenter(e, v);
	else if (kind == BT_REQ_SLIDER) uSlider(e, a0, a1);
	else if (kind == BT_REQ_ERASE) uErase(cur->gs, e);
	else if (kind == BT_REQ_TYPEMASK) uMyTypeMask(e, (uint32_t)a0);
	else if (kind == BT_REQ_COLLIDEMASK) uMyCollideMask(e, (uint32_t)a0);
	else return -1;
	return 0;
}

/* every entity that has a crush handler gets the recorder instead (same catalog id on export) */
int btref_lab_record_all_crushes(void) {
	int k = 0;
	for (ent *e = cur->gs->ents; e; e = e->ll.n)
		if (e->crush && e->crush != lab_crush_recorder) {
			lab_crush_id[e] = crushHandlers.reverseLookup(e->crush);
			e->crush = lab_crush_recorder;
			k++;
		}
	return k;
}

int btref_lab_crush_log(int32_t *out, int max) {
	int n = (int)lab_crush_log.size();
	for (int i = 0; i < n && i < max; i++) out[i] = lab_crush_log[i];
	lab_crush_log.clear();
	return n;
}

/* Replace collision-path handlers that have no device-side implementation by
 * the catalog's NIL entry, and (mode&2) drop every tick handler, so that a
 * shipped level becomes a pure physics fixture.  This edits the INPUT world,
 * not the reference code. keep_* are bitmasks over catalog IDs to keep. */
void btref_strip_handlers(int drop_ticks, uint32_t keep_push, uint32_t keep_pushed, uint32_t keep_pair,
                          uint32_t keep_crush) {
	for (ent *e = cur->gs->ents; e; e = e->ll.n) {
		if (drop_ticks) { e->tick = NULL; e->tickHeld = NULL; }
		int id = pushHandlers.reverseLookup(e->push);
		if (!(keep_push >> id & 1)) e->push = pushHandlers.get(PUSH_NIL);
		if (e->pushed) {
			id = pushedHandlers.reverseLookup(e->pushed);
			if (!(keep_pushed >> id & 1)) e->pushed = NULL;
		}
		if (e->crush) {
			id = e->crush == lab_crush_recorder ? lab_crush_id[e] : crushHandlers.reverseLookup(e->crush);
			if (!(keep_crush >> id & 1)) e->crush = NULL;
		}
		entPair_t *pairs[4] = {&e->onFumble, &e->onFumbled, &e->onPickUp, &e->onPickedUp};
		for (int k = 0; k < 4; k++) {
			id = entPairHandlers.reverseLookup(*pairs[k]);
			if (!(keep_pair >> id & 1)) *pairs[k] = entPairHandlers.get(ENTPAIR_NIL);
		}
	}
}

/* Seeded velocity kicks through the reference's own deferred updater uVel
 * (entUpdaters.c:14-18): every `every`-th root (in uid order) gets d_vel in
 * [-max_v, max_v]^3. */
void btref_kick(uint32_t seed, int32_t max_v, int32_t every) {
	RefWorld *w = cur;
	std::vector<ent *> order;
	for (ent *e = w->gs->ents; e; e = e->ll.n) order.push_back(e);
	uint32_t st = seed;
	int k = 0;
	for (int i = (int)order.size() - 1; i >= 0; i--) {
		ent *e = order[i];
		if (e->holder) continue;
		if ((k++ % every) != 0) continue;
		int32_t v[3];
		for (int d = 0; d < 3; d++) v[d] = (int32_t)(splitmix32(&st) % (uint32_t)(2 * max_v + 1)) - max_v;
		uVel(e, v);
	}
}

void btref_set_gamerules(int32_t rules) { cur->gs->gamerules = rules; }
int32_t btref_get_gamerules(void) { return cur->gs->gamerules; }

/* doInputs main.cpp:337-350 */
static void do_inputs(ent *e, const signed char *data) {
	int32_t otherButtons = 0;
	if (data[0] & 1) pushBtn(e, 0);
	if (data[0] & 2) pushBtn(e, 1);
	if (data[0] & 4) pushBtn(e, 2);
	if (data[0] & 8) otherButtons |= PLAYER_BTN_ALTFIRE1;
	if (data[0] & 16) pushBtn(e, 3);
	if (data[0] & 32) otherButtons |= PLAYER_BTN_ALTFIRE2;
	int32_t axis[2] = {data[1], data[2]};
	player_setAxis(e, axis);
	int32_t look[3] = {data[3], data[4], data[5]};
	player_setLook(e, look);
	player_setButtons(e, otherButtons);
}

static void gen_inputs(RefWorld *w, int player, signed char out[6]) {
	/* splitmix32 keyed by (player, tick, world key) */
	uint32_t st = 0x1234567u * (uint32_t)(player + 1) + (uint32_t)w->tick * 0x9e3779b9u + w->input_key;
	uint32_t a = splitmix32(&st), b = splitmix32(&st), c = splitmix32(&st);
	out[0] = (signed char)(a & 63);
	out[1] = (signed char)((int)((a >> 8) % 129u) - 64);
	out[2] = (signed char)((int)((a >> 16) % 129u) - 64);
	out[3] = (signed char)((int)(b % 129u) - 64);
	out[4] = (signed char)((int)((b >> 8) % 129u) - 64);
	out[5] = (signed char)((int)(c % 65u) - 32);
}

/* Everything of doWholeStep before doPhysics (main.cpp:1019-1090).
 * mode 0: full (world rules, players, tick handlers); mode 1: physics only
 * (flushCtrls restated from ent.cpp:817-824, no tick handlers). */
void btref_pre_physics(int mode) {
	RefWorld *w = cur;
	gamestate *gs = w->gs;
	rootState = gs;
	if (mode == 0) {
		if (w->reload_pending) {
			/* what the real client does on requestReload: BIN_CMD_ADD of the level
			 * (main.cpp:444-461, 923-928) */
			w->reload_pending = 0;
			list<const char> const fake = {.items = w->sav.data(), .num = (int)w->sav.size(), .max = (int)w->sav.size()};
			deserializeSelected(gs, &fake, zeroVec, zeroVec);
		}
		if (gs->gamerules & EFFECT_CRUSH) doCrushtainer(gs);
		if (gs->gamerules & EFFECT_CRUSH_BIG) doBigCrushtainer(gs);
		if (gs->gamerules & EFFECT_LAVA) doLava(gs);
		if (gs->gamerules & EFFECT_BLOCKS) createDebris(gs);
		if (gs->gamerules & EFFECT_BOUNCE) doBouncetainer(gs);
		for (int i = 0; i < gs->players.num; i++) {
			ent *e = gs->players[i].entity;
			if (e) {
				signed char data[6];
				gen_inputs(w, i, data);
				do_inputs(e, data);
			} else {
				gs->players[i].reviveCounter++;
			}
		}
		if (w->spacetrash_on) {
			tick_t st = tickHandlers.get(TICK_SPACETRASH);
			for (ent *e = gs->ents; e; e = e->ll.n)
				if (e->tick == st) e->ctrl.btns[1].v2 = 1;
		}
		doUpdates(gs);
		if (gs->gamerules & EFFECT_SPAWN) {
			/* mkHeroes main.cpp:609-620 */
			int np = gs->players.num;
			for (int i = 0; i < np; i++) {
				player *p = &gs->players[i];
				if (p->entity == NULL) {
					if (p->reviveCounter < 90) continue;
					p->entity = mkHero(gs, i, np);
					p->entity->color = p->color;
					p->reviveCounter = 0;
				}
			}
		}
	} else {
		if (mode == 2) { /* physics-only plus the scan rules of main.cpp:1030-1034 */
			if (gs->gamerules & EFFECT_CRUSH) doCrushtainer(gs);
			if (gs->gamerules & EFFECT_CRUSH_BIG) doBigCrushtainer(gs);
			if (gs->gamerules & EFFECT_LAVA) doLava(gs);
			if (gs->gamerules & EFFECT_BOUNCE) doBouncetainer(gs);
		}
		for (ent *e = gs->ents; e; e = e->ll.n) flushCtrls(e);
	}
	prepPhysics(gs);
	if (gs->gamerules & EFFECT_GRAV) doGravity(gs);
	assign_new_uids(w);
}

/* doPhysics + the tail of doWholeStep (main.cpp:1092-1097). */
void btref_do_physics(void) {
	RefWorld *w = cur;
	gamestate *gs = w->gs;
	rootState = gs;
	doPhysics(gs);
	/* cleanupDeadHeroes main.cpp:622-631 */
	for (int i = 0; i < gs->players.num; i++) {
		player *p = &gs->players[i];
		if (p->entity && p->entity->dead) p->entity = NULL;
	}
	assign_new_uids(w);
	forget_dead(w);
	finishStep(gs);
	w->tick++;
}

void btref_step(int mode) {
	btref_pre_physics(mode);
	btref_do_physics();
}

int btref_count(int32_t *n, int32_t *n_sliders, int32_t *n_wires) {
	int32_t a = 0, b = 0, c = 0;
	for (ent *e = cur->gs->ents; e; e = e->ll.n) {
		a++;
		b += e->numSliders;
		c += e->wires.num;
	}
	*n = a;
	if (n_sliders) *n_sliders = b;
	if (n_wires) *n_wires = c;
	return 0;
}

/* Fill caller-allocated arrays (sizes from btref_count) with the current
 * gamestate.  Entity index = rank from the tail of gs->ents (oldest = 0),
 * identical to the .sav index (serialize.cpp:79-94). World slot `wslot`
 * of the per-world arrays is written; entity arrays are written from
 * `ent_base` with references offset by ent_base (batched export). */
int btref_export_at(bt_world_desc *d, int32_t ent_base, int32_t slider_base, int32_t wire_base, int32_t wslot) {
	RefWorld *w = cur;
	gamestate *gs = w->gs;
	std::vector<ent *> order;
	for (ent *e = gs->ents; e; e = e->ll.n) order.push_back(e);
	int n = (int)order.size();
	std::unordered_map<ent *, int32_t> idx;
	for (int i = 0; i < n; i++) idx[order[n - 1 - i]] = ent_base + i;
	std::unordered_map<ent *, int32_t> rootpos;
	{
		int k = 0;
		for (ent *e = gs->rootEnts; e; e = e->LL.n) rootpos[e] = k++;
	}
	auto ref = [&](ent *p) -> int32_t {
		if (!p) return -1;
		auto it = idx.find(p);
		return it == idx.end() ? -1 : it->second;
	};
	d->w_flip_flop[wslot] = gs->flipFlop_death;
	d->w_gamerules[wslot] = gs->gamerules;
	d->w_rand[wslot] = gs->rand;
	int32_t so = slider_base, wo = wire_base;
	for (int i = 0; i < n; i++) {
		ent *e = order[n - 1 - i];
		int32_t o = ent_base + i;
		if (d->world_id) d->world_id[o] = wslot;
		d->uid[o] = w->uid.count(e) ? w->uid[e] : -1;
		for (int k = 0; k < 3; k++) {
			d->center[3 * o + k] = e->center[k];
			d->vel[3 * o + k] = e->vel[k];
			d->old[3 * o + k] = e->old[k];
			d->d_center[3 * o + k] = e->d_center[k];
			d->d_vel[3 * o + k] = e->d_vel[k];
			d->radius[3 * o + k] = e->radius[k];
		}
		d->type_mask[o] = e->typeMask;
		d->collide_mask[o] = e->collideMask;
		d->new_type_mask[o] = e->newTypeMask;
		d->new_collide_mask[o] = e->newCollideMask;
		d->friction[o] = e->friction;
		d->holder[o] = ref(e->holder);
		d->hold_flags[o] = e->holdFlags;
		d->holdee_head[o] = ref(e->holdee);
		d->holdee_next[o] = e->holder ? ref(e->LL.n) : -1;
		d->root_order[o] = e->holder ? -1 : rootpos[e];
		d->new_holder[o] = ref(e->newHolder);
		d->new_hold_flags[o] = e->newHoldFlags;
		uint8_t *h = d->handlers + (size_t)BT_H_NUM * o;
		h[BT_H_WHOMOVES] = (uint8_t)whoMovesHandlers.reverseLookup(e->whoMoves);
		h[BT_H_TICK] = (uint8_t)(e->tick ? tickHandlers.reverseLookup(e->tick) : 0);
		h[BT_H_TICKHELD] = (uint8_t)(e->tickHeld ? tickHandlers.reverseLookup(e->tickHeld) : 0);
		h[BT_H_CRUSH] = (uint8_t)(e->crush ? (e->crush == lab_crush_recorder ? lab_crush_id[e] : crushHandlers.reverseLookup(e->crush)) : 0);
		h[BT_H_PUSH] = (uint8_t)pushHandlers.reverseLookup(e->push);
		h[BT_H_PUSHED] = (uint8_t)(e->pushed ? pushedHandlers.reverseLookup(e->pushed) : 0);
		h[BT_H_ONFUMBLE] = (uint8_t)entPairHandlers.reverseLookup(e->onFumble);
		h[BT_H_ONFUMBLED] = (uint8_t)entPairHandlers.reverseLookup(e->onFumbled);
		h[BT_H_ONPICKUP] = (uint8_t)entPairHandlers.reverseLookup(e->onPickUp);
		h[BT_H_ONPICKEDUP] = (uint8_t)entPairHandlers.reverseLookup(e->onPickedUp);
		for (int k = 0; k < 4; k++) {
			d->btn[8 * o + k] = (uint8_t)e->ctrl.btns[k].v;
			d->btn[8 * o + 4 + k] = (uint8_t)e->ctrl.btns[k].v2;
		}
		uint8_t *m = d->misc + (size_t)BT_M_NUM * o;
		m[BT_M_NEWDROP] = (uint8_t)e->newDrop;
		m[BT_M_DEADMAX0] = (uint8_t)e->dead_max[0];
		m[BT_M_DEADMAX1] = (uint8_t)e->dead_max[1];
		m[BT_M_FULLFLUSH] = (uint8_t)e->fullFlush;
		d->slider_off[o] = so;
		for (int k = 0; k < e->numSliders; k++, so++) {
			d->slider_v[so] = e->sliders[k].v;
			d->slider_min[so] = e->sliders[k].min;
			d->slider_max[so] = e->sliders[k].max;
		}
		d->wire_off[o] = wo;
		for (int k = 0; k < e->wires.num; k++, wo++) d->wire_to[wo] = ref(e->wires[k]);
	}
	d->slider_off[ent_base + n] = so;
	d->wire_off[ent_base + n] = wo;
	return n;
}

int btref_export(bt_world_desc *d) {
	int n = btref_export_at(d, 0, 0, 0, 0);
	d->n_ents = n;
	d->n_sliders = d->slider_off[n];
	d->n_wires = d->wire_off[n];
	d->n_worlds = 1;
	return 0;
}

/* ---- state hash; must agree bit for bit with oracle/bt_oracle.c and the CUDA path ---- */
static inline uint64_t mix64(uint64_t x) {
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
	x ^= x >> 27; x *= 0x94d049bb133111ebULL;
	x ^= x >> 31;
	return x;
}
static inline uint64_t hash_words(int32_t uid, const int32_t *wd, int n) {
	uint64_t h = mix64((uint64_t)(uint32_t)uid + 0x9E3779B97F4A7C15ULL);
	for (int i = 0; i < n; i++) h = mix64(h + (uint64_t)(uint32_t)wd[i] + 0x9E3779B97F4A7C15ULL);
	return h;
}
uint64_t btref_hash(void) {
	RefWorld *w = cur;
	uint64_t sum = 0;
	int64_t n = 0;
	for (ent *e = w->gs->ents; e; e = e->ll.n) {
		int32_t wd[10] = {e->center[0], e->center[1], e->center[2], e->vel[0], e->vel[1], e->vel[2],
		                  (int32_t)e->typeMask, (int32_t)e->collideMask,
		                  e->holder ? (w->uid.count(e->holder) ? w->uid[e->holder] : -1) : -1, e->holdFlags};
		sum += hash_words(w->uid.count(e) ? w->uid[e] : -1, wd, 10);
		n++;
	}
	return sum + mix64((uint64_t)n);
}

/* full-state hash, bt_world_hash_full of include/bittoss_b200.h restated over the reference's `ent`s: the ten
 * words above, then radius, old, friction, new*Mask, packed handler ids, buttons v / v2, newHolder uid (-2 = the
 * entity itself, entUpdaters.c:51) and flags, newDrop | dead_max[0]<<8 | dead_max[1]<<16 | fullFlush<<24, sliders
 * (count; v, min, max each), wires (count; uid each); plus mix of flipFlop_death, rand, gamerules per world */
static inline uint64_t hword(uint64_t h, int32_t w) { return mix64(h + (uint64_t)(uint32_t)w + 0x9E3779B97F4A7C15ULL); }
uint64_t btref_hash_full(void) {
	RefWorld *w = cur;
	gamestate *gs = w->gs;
	uint64_t sum = 0;
	int64_t n = 0;
	auto uid_live = [&](ent *p) -> int32_t { return p && !p->dead && w->uid.count(p) ? w->uid[p] : -1; };
	for (ent *e = gs->ents; e; e = e->ll.n) {
		int32_t wd[10] = {e->center[0], e->center[1], e->center[2], e->vel[0], e->vel[1], e->vel[2],
		                  (int32_t)e->typeMask, (int32_t)e->collideMask, uid_live(e->holder), e->holdFlags};
		uint64_t h = hash_words(w->uid.count(e) ? w->uid[e] : -1, wd, 10);
		for (int k = 0; k < 3; k++) h = hword(h, e->radius[k]);
		for (int k = 0; k < 3; k++) h = hword(h, e->old[k]);
		h = hword(h, e->friction);
		h = hword(h, (int32_t)e->newTypeMask); h = hword(h, (int32_t)e->newCollideMask);
		uint32_t crush_id = e->crush ? (e->crush == lab_crush_recorder ? lab_crush_id[e] : crushHandlers.reverseLookup(e->crush)) : 0;
		h = hword(h, (int32_t)((uint32_t)whoMovesHandlers.reverseLookup(e->whoMoves) | ((uint32_t)pushHandlers.reverseLookup(e->push) << 8) |
		                       ((uint32_t)(e->pushed ? pushedHandlers.reverseLookup(e->pushed) : 0) << 16) | (crush_id << 24)));
		h = hword(h, (int32_t)((uint32_t)entPairHandlers.reverseLookup(e->onFumble) | ((uint32_t)entPairHandlers.reverseLookup(e->onFumbled) << 8) |
		                       ((uint32_t)entPairHandlers.reverseLookup(e->onPickUp) << 16) | ((uint32_t)entPairHandlers.reverseLookup(e->onPickedUp) << 24)));
		h = hword(h, (int32_t)((uint32_t)(e->tick ? tickHandlers.reverseLookup(e->tick) : 0) | ((uint32_t)(e->tickHeld ? tickHandlers.reverseLookup(e->tickHeld) : 0) << 8)));
		uint32_t bv = 0, bv2 = 0;
		for (int k = 0; k < 4; k++) { bv |= (uint32_t)(uint8_t)e->ctrl.btns[k].v << (8 * k); bv2 |= (uint32_t)(uint8_t)e->ctrl.btns[k].v2 << (8 * k); }
		h = hword(h, (int32_t)bv); h = hword(h, (int32_t)bv2);
		h = hword(h, e->newHolder == e ? -2 : uid_live(e->newHolder));
		h = hword(h, e->newHolder ? e->newHoldFlags : 0);
		h = hword(h, (int32_t)(uint8_t)e->newDrop | ((int32_t)(uint8_t)e->dead_max[0] << 8) | ((int32_t)(uint8_t)e->dead_max[1] << 16) |
		                 ((int32_t)(uint8_t)e->fullFlush << 24));
		h = hword(h, e->numSliders);
		for (int k = 0; k < e->numSliders; k++) { h = hword(h, e->sliders[k].v); h = hword(h, e->sliders[k].min); h = hword(h, e->sliders[k].max); }
		h = hword(h, e->wires.num);
		for (int k = 0; k < e->wires.num; k++) h = hword(h, uid_live(e->wires[k]));
		sum += h;
		n++;
	}
	uint64_t g = mix64((uint64_t)(uint32_t)gs->flipFlop_death + 0x9E3779B97F4A7C15ULL);
	g = hword(g, gs->rand);
	g = hword(g, gs->gamerules);
	return sum + mix64((uint64_t)n) + g;
}

int32_t btref_tick_index(void) { return cur->tick; }

/* selectCast (raycast.cpp:16-57) on entity indices of the current export order. */
int btref_selectcast_batch(const int32_t *origin, const int32_t *look, int32_t n_rays, int32_t *winner) {
	gamestate *gs = cur->gs;
	std::vector<ent *> order;
	for (ent *e = gs->ents; e; e = e->ll.n) order.push_back(e);
	int n = (int)order.size();
	std::unordered_map<ent *, int32_t> idx;
	for (int i = 0; i < n; i++) idx[order[n - 1 - i]] = i;
	for (int r = 0; r < n_rays; r++) {
		int32_t lk[3] = {look[3 * r], look[3 * r + 1], look[3 * r + 2]};
		ent *parent = order[n - 1 - origin[r]];
		ent *hit = selectCast(gs, parent, lk);
		winner[r] = hit ? idx[hit] : -1;
	}
	return 0;
}

/* Wall-clock of n whole steps, for the CPU baseline. Returns seconds, and the
 * sum of live entity counts after each tick in *entity_ticks. */
double btref_time_steps(int mode, int32_t n, int64_t *entity_ticks) {
	struct timespec a, b;
	int64_t et = 0;
	clock_gettime(CLOCK_MONOTONIC, &a);
	for (int i = 0; i < n; i++) {
		btref_step(mode);
		int64_t c = 0;
		for (ent *e = cur->gs->ents; e; e = e->ll.n) c++;
		et += c;
	}
	clock_gettime(CLOCK_MONOTONIC, &b);
	if (entity_ticks) *entity_ticks = et;
	return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}

} /* extern "C" */
