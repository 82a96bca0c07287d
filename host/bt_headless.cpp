/*
 * bt_headless.cpp — a headless bittoss host with its physics on the GPU.
 *
 * Product code: the stand-in for main.cpp in a process without GLFW/GL/network.  host/Makefile compiles it
 * together with the UNMODIFIED reference sources where they lie (the simulation core, every module, the
 * .sav loader) and host/bt_glue.cpp into host/_build/libbthost.so.  What this file restates from main.cpp
 * (which cannot be built here: it needs GLFW) is the init order (main.cpp:1690-1761), resetPlayer / mkHeroes
 * / cleanupDeadHeroes (:308-317, 609-631), doInputs (:337-350) and doWholeStep (:1019-1098) — with the
 * five lines prepPhysics .. finishStep replaced by the binding, exactly as bt_glue.h says:
 *
 *     doUpdates(state); mkHeroes(state);              // host, reference code, unchanged
 *     bt_glue_physics(glue, state);                   // device: prepPhysics, doGravity, doPhysics
 *     cleanupDeadHeroes(state);                       // host
 *     bt_glue_finish(glue, state);                    // finishStep
 *
 * A world that is not attached to a device steps through the reference's own CPU functions instead (same
 * process, same inputs), which is how a host would keep its rollback copy or compare the two.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
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
#include "random.h"
#include "modules/player.h"

#include "bt_glue.h"

/* symbols main.cpp / graphics.cpp provide in the real client (main.h:11-17, graphics.h:38-46) */
volatile char globalRunning = 1;
volatile char dl_srm_ready = 0;
gamestate *rootState = NULL;
void showMessage(gamestate const *const gs, char const *const msg) {}
void rect(int32_t *p, int32_t *radius, float r, float g, float b) {}
void drawTag(const char *str, int32_t *pos, double scale, float *color) {}
void setupStipple() {}
void setupTags() {}

struct HostWorld {
	gamestate *gs;
	bt_glue *glue;          /* NULL: stepped on the CPU by the reference's own doPhysics */
	bt_uid_tracker *uids;   /* creation-order ids for the hashes of a CPU-stepped world */
	int32_t tick;
	int reload_pending;
	std::vector<char> sav;
	int spacetrash_on;
	uint32_t input_key;
	int32_t pdata[16];
};
static std::vector<HostWorld *> worlds;
static HostWorld *cur = NULL;
static int inited = 0;
static char data_dir[4096];
static char host_err[600] = "";

void requestReload(gamestate const *const gs) {
	for (HostWorld *w : worlds)
		if (w && w->gs == gs) w->reload_pending = 1;
}
void requestLoad(gamestate const *gs, int playerIx, int gameStrIx) {}

extern "C" {

const char *bth_error(void) { return host_err[0] ? host_err : bt_glue_error(); }

/* main.cpp:1690-1761; backend_so NULL = the CUDA library of this repository */
int bth_init(const char *dir_with_data, const char *backend_so) {
	if (inited) return 0;
	setvbuf(stdout, NULL, _IOLBF, 0);
	int r = bt_glue_bind(backend_so);
	if (r) return r;
	char prev[4096];
	if (!getcwd(prev, sizeof prev)) return -1;
	if (chdir(dir_with_data)) { snprintf(host_err, sizeof host_err, "cannot enter %s", dir_with_data); return -1; }
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
	if (chdir(prev)) return -1;
	inited = 1;
	return 0;
}

int bth_new_world(void) {
	HostWorld *w = new HostWorld();
	w->gs = mkGamestate();
	w->glue = NULL;
	w->uids = bt_uid_tracker_new();
	w->tick = 0; w->reload_pending = 0; w->spacetrash_on = 0; w->input_key = 0;
	memset(w->pdata, 0, sizeof w->pdata);
	worlds.push_back(w);
	cur = w;
	rootState = w->gs;
	return (int)worlds.size() - 1;
}
int bth_select(int id) {
	if (id < 0 || id >= (int)worlds.size() || !worlds[id]) return -1;
	cur = worlds[id];
	rootState = cur->gs;
	return 0;
}
void bth_free_world(int id) {
	if (id < 0 || id >= (int)worlds.size() || !worlds[id]) return;
	HostWorld *w = worlds[id];
	bt_glue_detach(w->glue);
	bt_uid_tracker_free(w->uids);
	doCleanup(w->gs);
	free(w->gs);
	if (cur == w) { cur = NULL; rootState = NULL; }
	delete w;
	worlds[id] = NULL;
}

/* resetPlayer main.cpp:308-317 */
static void reset_player(HostWorld *w, int i) {
	player *p = &w->gs->players[i];
	p->reviveCounter = 100000;
	p->data = w->pdata[i & 15];
	p->entity = NULL;
	p->color = 0xCCCCFF;
	snprintf(p->name, NAME_BUF_LEN, "p%d", i);
}

/* main.cpp:1808-1813 + setupPlayers :319-323 */
int bth_load_sav(const char *name, int n_players, const int32_t *pdata, int32_t rand_seed) {
	HostWorld *w = cur;
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
	bt_uid_tracker_step(w->uids, w->gs);
	return 0;
}

void bth_set_options(int spacetrash_on, uint32_t input_key) {
	cur->spacetrash_on = spacetrash_on;
	cur->input_key = input_key;
}

/* from here on the current world's physics run on `device` */
int bth_attach(int device, int32_t max_ents) {
	if (cur->glue) return 0;
	cur->glue = bt_glue_attach(cur->gs, device, max_ents);
	if (!cur->glue) { snprintf(host_err, sizeof host_err, "%s", bt_glue_error()); return -1; }
	return 0;
}

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

/* the seeded input stream of SURVEY.md §8(d) config 1: 6 bytes per player and tick */
static void seeded_inputs(HostWorld *w, int player, signed char out[6]) {
	uint32_t st = 0x1234567u * (uint32_t)(player + 1) + (uint32_t)w->tick * 0x9e3779b9u + w->input_key;
	uint32_t a = splitmix32(&st), b = splitmix32(&st), c = splitmix32(&st);
	out[0] = (signed char)(a & 63);
	out[1] = (signed char)((int)((a >> 8) % 129u) - 64);
	out[2] = (signed char)((int)((a >> 16) % 129u) - 64);
	out[3] = (signed char)((int)(b % 129u) - 64);
	out[4] = (signed char)((int)((b >> 8) % 129u) - 64);
	out[5] = (signed char)((int)(c % 65u) - 32);
}

/* doWholeStep main.cpp:1019-1098.  inputs: 6 bytes per player, or NULL for the seeded stream. */
int bth_step(const signed char *inputs) {
	HostWorld *w = cur;
	gamestate *gs = w->gs;
	rootState = gs;
	if (w->reload_pending) { /* what the client does on requestReload: BIN_CMD_ADD of the level (main.cpp:444-461, 923-928) */
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
			if (inputs) memcpy(data, inputs + 6 * i, 6);
			else seeded_inputs(w, i, data);
			do_inputs(e, data);
		} else {
			gs->players[i].reviveCounter++;
		}
	}
	if (w->spacetrash_on) { /* the shipped level holds its spawner down: release it */
		tick_t st = tickHandlers.get(TICK_SPACETRASH);
		for (ent *e = gs->ents; e; e = e->ll.n)
			if (e->tick == st) e->ctrl.btns[1].v2 = 1;
	}
	doUpdates(gs);
	if (gs->gamerules & EFFECT_SPAWN) { /* mkHeroes main.cpp:609-620 */
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
	if (w->glue) {
		if (bt_glue_physics(w->glue, gs)) { snprintf(host_err, sizeof host_err, "%s", bt_glue_error()); return -1; }
	} else {
		prepPhysics(gs);
		if (gs->gamerules & EFFECT_GRAV) doGravity(gs);
		bt_uid_tracker_step(w->uids, gs);
		doPhysics(gs);
	}
	for (int i = 0; i < gs->players.num; i++) { /* cleanupDeadHeroes main.cpp:622-631 */
		player *p = &gs->players[i];
		if (p->entity && p->entity->dead) p->entity = NULL;
	}
	if (w->glue) {
		if (bt_glue_finish(w->glue, gs)) { snprintf(host_err, sizeof host_err, "%s", bt_glue_error()); return -1; }
	} else {
		bt_uid_tracker_step(w->uids, gs);
		bt_uid_tracker_forget_dead(w->uids, gs);
		finishStep(gs);
	}
	w->tick++;
	return 0;
}

/* out[0] = state hash, out[1] = full-state hash, both computed on the host from the `ent`s */
void bth_hash(uint64_t *out) {
	if (cur->glue) bt_glue_host_hash(cur->glue, cur->gs, &out[0], &out[1]);
	else bt_tracked_hash(cur->uids, cur->gs, &out[0], &out[1]);
}
/* the same two from the device arrays (attached worlds) */
int bth_device_hash(uint64_t *out) {
	if (!cur->glue) return -1;
	if (bt_glue_hash(cur->glue, &out[0], &out[1])) { snprintf(host_err, sizeof host_err, "%s", bt_glue_error()); return -1; }
	return 0;
}

/* the reference's own writer (serialize.cpp:187-212) on the current world */
int bth_serialize(char *out, int cap) {
	list<char> data;
	data.init();
	serialize(cur->gs, &data);
	int n = data.num;
	if (n <= cap) memcpy(out, data.items, n);
	data.destroy();
	return n;
}

int bth_count(void) {
	int n = 0;
	for (ent *e = cur->gs->ents; e; e = e->ll.n) n++;
	return n;
}
int32_t bth_tick_index(void) { return cur->tick; }
void bth_counters(int64_t *out) {
	bt_glue_counters c;
	memset(&c, 0, sizeof c);
	if (cur->glue) bt_glue_get_counters(cur->glue, &c);
	out[0] = c.pushes; out[1] = c.pokes; out[2] = c.spawned; out[3] = c.pulls; out[4] = c.deaths; out[5] = c.events_push;
	out[6] = c.events_crush; out[7] = c.rekeys;
}
void *bth_world(void) { return cur->glue ? (void *)bt_glue_world(cur->glue) : NULL; }

} /* extern "C" */
