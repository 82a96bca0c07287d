/*
 * bt_glue.cpp — gamestate <-> device bridge (see bt_glue.h).  Product code, compiled against the
 * unmodified reference headers by host/Makefile.
 *
 * How the two sides stay in step without re-uploading:
 *   - every `ent` of the gamestate owns one device SLOT for life (slot = creation order = the uid the
 *     state hash uses); `shadow` holds, array by array (BT_A_*), what the device holds for every slot;
 *   - bt_glue_push compares each `ent` with the shadow and sends the differences as sparse writes
 *     (bt_world_poke): that catches the deferred updaters of entUpdaters.c (d_vel, d_center, slider
 *     min/max, new*Mask, dead_max, newHolder, newDrop, wiresAdd / wiresRm), direct writes by handlers
 *     (handler pointers, sliders[].v, holdFlags ...), and entities made by initEnt / addEnt
 *     (entFuncs.c:63-111, ent.cpp:107-113), which get fresh slots (bt_world_grow + BT_OP_SPAWN);
 *   - bt_glue_pull reads the arrays back, writes them into the `ent`s and rebuilds the pointer lists
 *     the device edits as index links: holder / holdee / LL (fumble, pickup: ent.cpp:136-188), gs->ents
 *     and gs->deadTail (killEntNoHandlers ent.cpp:190-208), gs->rootEnts in the device's rootKey order.
 */
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "ent.h"
#include "handlerRegistrar.h"
#include "bt_glue.h"

/* ------------------------------------------------------------------ the C ABI, bound at run time */
static struct Api {
	void *lib;
	const char *(*last_error)(void);
	int (*create)(bt_world **, int, int32_t, int32_t, int32_t, int32_t);
	void (*destroy)(bt_world *);
	int (*upload)(bt_world *, const bt_world_desc *);
	int (*step)(bt_world *, uint32_t);
	int (*hash)(bt_world *, uint64_t *);
	int (*hash_full)(bt_world *, uint64_t *);
	int (*drain_events)(bt_world *, bt_event *, int32_t, int32_t *);
	int (*poke)(bt_world *, const bt_poke *, int32_t);
	int (*pull)(bt_world *, int32_t, int32_t, int32_t, void *);
	int (*extent)(bt_world *, int32_t, int32_t *);
	int (*grow)(bt_world *, int32_t, const int32_t *, const int32_t *, int32_t, int32_t *);
	int (*set_wires)(bt_world *, int32_t, const int32_t *, int32_t, int32_t);
	int64_t (*next_root_phase)(bt_world *);
} api;

static char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}
const char *bt_glue_error(void) { return g_err; }
#define API(call) do { int _r = (call); if (_r) return fail(_r, "%s -> %d: %s", #call, _r, api.last_error ? api.last_error() : ""); } while (0)

int bt_glue_bind(const char *so_path) {
	std::string path;
	if (so_path) path = so_path;
	else {
		Dl_info info;
		if (!dladdr((void *)&bt_glue_bind, &info) || !info.dli_fname) return fail(BT_ERR_ARG, "cannot locate this library");
		path = info.dli_fname; /* <repo>/host/_build/libbthost.so -> <repo>/bittoss_b200/libbittoss_b200.so */
		for (int up = 0; up < 3; up++) {
			size_t p = path.find_last_of('/');
			path = p == std::string::npos ? "." : path.substr(0, p);
		}
		path += "/bittoss_b200/libbittoss_b200.so";
	}
	void *lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
	if (!lib) return fail(BT_ERR_NO_DEVICE, "cannot load %s: %s (there is no CPU fallback)", path.c_str(), dlerror());
	memset(&api, 0, sizeof api);
	api.lib = lib;
#define SYM(field, name) do { *(void **)(&api.field) = dlsym(lib, name); if (!api.field) return fail(BT_ERR_ARG, "%s lacks %s", path.c_str(), name); } while (0)
	SYM(last_error, "bt_last_error"); SYM(create, "bt_world_create"); SYM(destroy, "bt_world_destroy");
	SYM(upload, "bt_world_upload"); SYM(step, "bt_world_step"); SYM(hash, "bt_world_hash"); SYM(hash_full, "bt_world_hash_full");
	SYM(drain_events, "bt_world_drain_events"); SYM(poke, "bt_world_poke"); SYM(pull, "bt_world_pull");
	SYM(extent, "bt_world_extent"); SYM(grow, "bt_world_grow"); SYM(set_wires, "bt_world_set_wires");
	SYM(next_root_phase, "bt_world_next_root_phase");
#undef SYM
	return BT_OK;
}

/* ------------------------------------------------------------------ handler pointers <-> catalog ids */
static uint32_t ids_phys(const ent *e) {
	return (uint32_t)whoMovesHandlers.reverseLookup(e->whoMoves) | ((uint32_t)pushHandlers.reverseLookup(e->push) << 8) |
	       ((uint32_t)(e->pushed ? pushedHandlers.reverseLookup(e->pushed) : 0) << 16) |
	       ((uint32_t)(e->crush ? crushHandlers.reverseLookup(e->crush) : 0) << 24);
}
static uint32_t ids_pair(const ent *e) {
	return (uint32_t)entPairHandlers.reverseLookup(e->onFumble) | ((uint32_t)entPairHandlers.reverseLookup(e->onFumbled) << 8) |
	       ((uint32_t)entPairHandlers.reverseLookup(e->onPickUp) << 16) | ((uint32_t)entPairHandlers.reverseLookup(e->onPickedUp) << 24);
}
static uint32_t ids_tick(const ent *e) {
	return (uint32_t)(e->tick ? tickHandlers.reverseLookup(e->tick) : 0) | ((uint32_t)(e->tickHeld ? tickHandlers.reverseLookup(e->tickHeld) : 0) << 8);
}
static uint32_t pack_btn(const ent *e, int second) {
	uint32_t v = 0;
	for (int k = 0; k < 4; k++) v |= (uint32_t)(uint8_t)(second ? e->ctrl.btns[k].v2 : e->ctrl.btns[k].v) << (8 * k);
	return v;
}

/* ------------------------------------------------------------------ export (first upload) */
struct bt_glue_desc {
	bt_world_desc d;
	std::vector<int32_t> i32s[24];
	std::vector<uint32_t> u32s[4];
	std::vector<uint8_t> u8s[3];
	std::vector<ent *> order; /* index -> ent */
};

bt_glue_desc *bt_glue_export(gamestate *gs) {
	bt_glue_desc *D = new bt_glue_desc();
	std::vector<ent *> &order = D->order;
	for (ent *e = gs->ents; e; e = e->ll.n) order.push_back(e);
	std::reverse(order.begin(), order.end()); /* oldest first */
	const int n = (int)order.size();
	std::unordered_map<const ent *, int32_t> idx;
	for (int i = 0; i < n; i++) idx[order[i]] = i;
	std::unordered_map<const ent *, int32_t> rootpos;
	{
		int k = 0;
		for (ent *e = gs->rootEnts; e; e = e->LL.n) rootpos[e] = k++;
	}
	auto ref = [&](const ent *p) -> int32_t {
		if (!p) return -1;
		auto it = idx.find(p);
		return it == idx.end() ? -1 : it->second;
	};
	int ns = 0, nw = 0;
	for (ent *e : order) { ns += e->numSliders; nw += e->wires.num; }
	bt_world_desc &d = D->d;
	memset(&d, 0, sizeof d);
	d.n_ents = n; d.n_sliders = ns; d.n_wires = nw; d.n_worlds = 1;
	int vi = 0;
	auto I = [&](size_t cnt) { D->i32s[vi].assign(cnt ? cnt : 1, 0); return D->i32s[vi++].data(); };
	d.w_flip_flop = I(1); d.w_gamerules = I(1); d.w_rand = I(1);
	d.world_id = NULL; d.uid = I(n);
	d.center = I(3 * n); d.vel = I(3 * n); d.old = I(3 * n); d.d_center = I(3 * n); d.d_vel = I(3 * n); d.radius = I(3 * n);
	d.friction = I(n); d.holder = I(n); d.hold_flags = I(n); d.holdee_head = I(n); d.holdee_next = I(n); d.root_order = I(n);
	d.new_holder = I(n); d.new_hold_flags = I(n);
	d.slider_off = I(n + 1); d.slider_v = I(ns); d.slider_min = I(ns); d.slider_max = I(ns);
	d.wire_off = I(n + 1); d.wire_to = I(nw);
	for (int k = 0; k < 4; k++) D->u32s[k].assign(n ? n : 1, 0);
	d.type_mask = D->u32s[0].data(); d.collide_mask = D->u32s[1].data(); d.new_type_mask = D->u32s[2].data(); d.new_collide_mask = D->u32s[3].data();
	D->u8s[0].assign((size_t)BT_H_NUM * (n ? n : 1), 0); D->u8s[1].assign((size_t)8 * (n ? n : 1), 0); D->u8s[2].assign((size_t)BT_M_NUM * (n ? n : 1), 0);
	d.handlers = D->u8s[0].data(); d.btn = D->u8s[1].data(); d.misc = D->u8s[2].data();
	d.w_flip_flop[0] = gs->flipFlop_death; d.w_gamerules[0] = gs->gamerules; d.w_rand[0] = gs->rand;
	int so = 0, wo = 0;
	for (int i = 0; i < n; i++) {
		const ent *e = order[i];
		d.uid[i] = i;
		for (int k = 0; k < 3; k++) {
			d.center[3 * i + k] = e->center[k]; d.vel[3 * i + k] = e->vel[k]; d.old[3 * i + k] = e->old[k];
			d.d_center[3 * i + k] = e->d_center[k]; d.d_vel[3 * i + k] = e->d_vel[k]; d.radius[3 * i + k] = e->radius[k];
		}
		d.type_mask[i] = e->typeMask; d.collide_mask[i] = e->collideMask;
		d.new_type_mask[i] = e->newTypeMask; d.new_collide_mask[i] = e->newCollideMask;
		d.friction[i] = e->friction;
		d.holder[i] = ref(e->holder); d.hold_flags[i] = e->holdFlags; d.holdee_head[i] = ref(e->holdee);
		d.holdee_next[i] = e->holder ? ref(e->LL.n) : -1;
		d.root_order[i] = e->holder ? -1 : rootpos[e];
		d.new_holder[i] = ref(e->newHolder); d.new_hold_flags[i] = e->newHoldFlags;
		const uint32_t hp = ids_phys(e), hq = ids_pair(e), ht = ids_tick(e);
		uint8_t *h = d.handlers + (size_t)BT_H_NUM * i;
		h[BT_H_WHOMOVES] = hp & 0xff; h[BT_H_PUSH] = (hp >> 8) & 0xff; h[BT_H_PUSHED] = (hp >> 16) & 0xff; h[BT_H_CRUSH] = (hp >> 24) & 0xff;
		h[BT_H_ONFUMBLE] = hq & 0xff; h[BT_H_ONFUMBLED] = (hq >> 8) & 0xff; h[BT_H_ONPICKUP] = (hq >> 16) & 0xff; h[BT_H_ONPICKEDUP] = (hq >> 24) & 0xff;
		h[BT_H_TICK] = ht & 0xff; h[BT_H_TICKHELD] = (ht >> 8) & 0xff;
		for (int k = 0; k < 4; k++) { d.btn[8 * i + k] = (uint8_t)e->ctrl.btns[k].v; d.btn[8 * i + 4 + k] = (uint8_t)e->ctrl.btns[k].v2; }
		uint8_t *m = d.misc + (size_t)BT_M_NUM * i;
		m[BT_M_NEWDROP] = (uint8_t)e->newDrop; m[BT_M_DEADMAX0] = (uint8_t)e->dead_max[0]; m[BT_M_DEADMAX1] = (uint8_t)e->dead_max[1];
		m[BT_M_FULLFLUSH] = (uint8_t)e->fullFlush;
		d.slider_off[i] = so;
		for (int k = 0; k < e->numSliders; k++, so++) { d.slider_v[so] = e->sliders[k].v; d.slider_min[so] = e->sliders[k].min; d.slider_max[so] = e->sliders[k].max; }
		d.wire_off[i] = wo;
		for (int k = 0; k < e->wires.num; k++, wo++) d.wire_to[wo] = ref(e->wires.items[k]);
	}
	d.slider_off[n] = so; d.wire_off[n] = wo;
	return D;
}
const bt_world_desc *bt_glue_desc_get(const bt_glue_desc *d) { return &d->d; }
void bt_glue_desc_free(bt_glue_desc *d) { delete d; }

/* ------------------------------------------------------------------ state hashes over `ent`s */
static inline uint64_t mix64(uint64_t x) {
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
	x ^= x >> 27; x *= 0x94d049bb133111ebULL;
	x ^= x >> 31;
	return x;
}
static inline uint64_t hw(uint64_t h, int32_t w) { return mix64(h + (uint64_t)(uint32_t)w + 0x9E3779B97F4A7C15ULL); }

template <class UidOf> static void hash_ents(gamestate *gs, UidOf uid_of, uint64_t *out, uint64_t *out_full) {
	uint64_t sum = 0, sumf = 0, cnt = 0;
	auto live_uid = [&](const ent *p) -> int32_t { return p && !p->dead ? uid_of(p) : -1; };
	for (ent *e = gs->ents; e; e = e->ll.n) {
		uint64_t h = mix64((uint64_t)(uint32_t)uid_of(e) + 0x9E3779B97F4A7C15ULL);
		for (int k = 0; k < 3; k++) h = hw(h, e->center[k]);
		for (int k = 0; k < 3; k++) h = hw(h, e->vel[k]);
		h = hw(h, (int32_t)e->typeMask); h = hw(h, (int32_t)e->collideMask);
		h = hw(h, live_uid(e->holder)); h = hw(h, e->holdFlags);
		sum += h;
		if (out_full) {
			for (int k = 0; k < 3; k++) h = hw(h, e->radius[k]);
			for (int k = 0; k < 3; k++) h = hw(h, e->old[k]);
			h = hw(h, e->friction);
			h = hw(h, (int32_t)e->newTypeMask); h = hw(h, (int32_t)e->newCollideMask);
			h = hw(h, (int32_t)ids_phys(e)); h = hw(h, (int32_t)ids_pair(e)); h = hw(h, (int32_t)ids_tick(e));
			h = hw(h, (int32_t)pack_btn(e, 0)); h = hw(h, (int32_t)pack_btn(e, 1));
			h = hw(h, e->newHolder == e ? -2 : live_uid(e->newHolder));
			h = hw(h, e->newHolder ? e->newHoldFlags : 0);
			h = hw(h, (int32_t)(uint8_t)e->newDrop | ((int32_t)(uint8_t)e->dead_max[0] << 8) | ((int32_t)(uint8_t)e->dead_max[1] << 16) |
			              ((int32_t)(uint8_t)e->fullFlush << 24));
			h = hw(h, e->numSliders);
			for (int k = 0; k < e->numSliders; k++) { h = hw(h, e->sliders[k].v); h = hw(h, e->sliders[k].min); h = hw(h, e->sliders[k].max); }
			h = hw(h, e->wires.num);
			for (int k = 0; k < e->wires.num; k++) h = hw(h, live_uid(e->wires.items[k]));
			sumf += h;
		}
		cnt++;
	}
	if (out) *out = sum + mix64(cnt);
	if (out_full) {
		uint64_t h = mix64((uint64_t)(uint32_t)gs->flipFlop_death + 0x9E3779B97F4A7C15ULL);
		h = hw(h, gs->rand);
		h = hw(h, gs->gamerules);
		*out_full = sumf + mix64(cnt) + h;
	}
}

struct bt_uid_tracker {
	std::unordered_map<const ent *, int32_t> uid;
	int32_t next;
};
bt_uid_tracker *bt_uid_tracker_new(void) { bt_uid_tracker *t = new bt_uid_tracker(); t->next = 0; return t; }
void bt_uid_tracker_free(bt_uid_tracker *t) { delete t; }
void bt_uid_tracker_step(bt_uid_tracker *t, gamestate *gs) {
	std::vector<ent *> fresh; /* new ents are always prepended to gs->ents (ent.cpp:47-52) */
	for (ent *e = gs->ents; e; e = e->ll.n) {
		if (t->uid.count(e)) break;
		fresh.push_back(e);
	}
	for (size_t i = fresh.size(); i-- > 0;) t->uid[fresh[i]] = t->next++;
}
void bt_uid_tracker_forget_dead(bt_uid_tracker *t, gamestate *gs) {
	for (ent *e = gs->deadTail; e; e = e->ll.p) t->uid.erase(e);
}
void bt_tracked_hash(bt_uid_tracker *t, gamestate *gs, uint64_t *hash, uint64_t *hash_full) {
	hash_ents(gs, [&](const ent *e) -> int32_t { auto it = t->uid.find(e); return it == t->uid.end() ? -1 : it->second; }, hash, hash_full);
}

/* ------------------------------------------------------------------ the binding */
static int elem_bytes(int a) {
	if (a >= BT_A_CENTER && a <= BT_A_RADIUS) return 16;
	if (a >= BT_A_NEWDROP && a <= BT_A_INLIST) return 1;
	if (a == BT_A_ROOTKEY || a == BT_A_DEADLIST) return 8;
	return 4;
}
/* the arrays both directions carry (BT_A_WORLD, BT_A_UID and the CSR headers only at creation) */
static const int SLOT_ARRAYS[] = {BT_A_CENTER, BT_A_VEL, BT_A_OLD, BT_A_DCENTER, BT_A_DVEL, BT_A_RADIUS, BT_A_TYPE, BT_A_COLLIDE, BT_A_NTYPE,
                                  BT_A_NCOLLIDE, BT_A_FRICTION, BT_A_HOLDER, BT_A_HOLDFLAGS, BT_A_HOLDROOT, BT_A_HOLDEE, BT_A_LLN, BT_A_LLP,
                                  BT_A_NEWHOLDER, BT_A_NEWHOLDFLAGS, BT_A_UID, BT_A_HPHYS, BT_A_HPAIR, BT_A_HTICK, BT_A_BTN_V, BT_A_BTN_V2,
                                  BT_A_NEWDROP, BT_A_DEADMAX0, BT_A_DEADMAX1, BT_A_FULLFLUSH, BT_A_DEAD, BT_A_INLIST, BT_A_ROOTKEY,
                                  BT_A_SLIDER_OFF, BT_A_WIRE_OFF, BT_A_WIRE_CAP, BT_A_WIRE_CNT};
static const int OTHER_ARRAYS[] = {BT_A_SV, BT_A_SMIN, BT_A_SMAX, BT_A_WIRE_TO, BT_A_W_FLIP, BT_A_W_GAMERULES, BT_A_W_RAND};

struct bt_glue {
	bt_world *w;
	std::vector<char> sh[BT_A_NUM]; /* shadow: what the device holds */
	std::vector<ent *> slot_ent;    /* slot -> ent (NULL once freed) */
	std::unordered_map<const ent *, int32_t> slot_of;
	std::vector<bt_poke> pokes;
	int32_t dead_seen; /* dead-list entries of the current tick already turned into host deaths */
	bt_glue_counters ctr;
	int32_t *I(int a) { return (int32_t *)sh[a].data(); }
	uint32_t *U(int a) { return (uint32_t *)sh[a].data(); }
	uint8_t *B(int a) { return (uint8_t *)sh[a].data(); }
	uint64_t *Q(int a) { return (uint64_t *)sh[a].data(); }
	int32_t slot(const ent *e) const {
		if (!e) return -1;
		auto it = slot_of.find(e);
		return it == slot_of.end() ? -1 : it->second;
	}
	ent *at(int32_t s) const { return s >= 0 && s < (int32_t)slot_ent.size() ? slot_ent[s] : NULL; }
	void poke(int a, int32_t index, int32_t v0, int32_t v1 = 0, int32_t v2 = 0) {
		bt_poke p; p.array = a; p.index = index; p.v[0] = v0; p.v[1] = v1; p.v[2] = v2;
		pokes.push_back(p);
	}
};
bt_world *bt_glue_world(bt_glue *g) { return g->w; }
void bt_glue_get_counters(bt_glue *g, bt_glue_counters *out) { *out = g->ctr; }

/* size the shadow of array a to the device's extent and read [first, extent) of it */
static int shadow_fetch(bt_glue *g, int a, int32_t first) {
	int32_t ext = 0;
	API(api.extent(g->w, a, &ext));
	const size_t eb = (size_t)elem_bytes(a);
	g->sh[a].resize(eb * (size_t)ext);
	if (first < ext) API(api.pull(g->w, a, first, ext - first, g->sh[a].data() + eb * (size_t)first));
	return BT_OK;
}

bt_glue *bt_glue_attach(gamestate *gs, int device, int32_t max_ents) {
	if (!api.lib) { if (bt_glue_bind(NULL)) return NULL; }
	bt_glue_desc *D = bt_glue_export(gs);
	const bt_world_desc &d = D->d;
	if (max_ents < d.n_ents) max_ents = d.n_ents;
	bt_glue *g = new bt_glue();
	g->w = NULL; g->dead_seen = 0;
	memset(&g->ctr, 0, sizeof g->ctr);
	/* sliders and wires of entities still to come: a few per entity is plenty for the shipped modules */
	int r = api.create(&g->w, device, max_ents, d.n_sliders + 16 * max_ents, d.n_wires + 8 * max_ents, 1);
	if (!r) r = api.upload(g->w, &d);
	if (r) {
		fail(r, "bt_world_create / upload -> %d: %s", r, api.last_error());
		if (g->w) api.destroy(g->w);
		delete g; bt_glue_desc_free(D);
		return NULL;
	}
	g->slot_ent = D->order;
	for (int i = 0; i < d.n_ents; i++) g->slot_of[D->order[i]] = i;
	bt_glue_desc_free(D);
	for (int a : SLOT_ARRAYS) if (shadow_fetch(g, a, 0)) { bt_glue_detach(g); return NULL; }
	for (int a : OTHER_ARRAYS) if (shadow_fetch(g, a, 0)) { bt_glue_detach(g); return NULL; }
	return g;
}
void bt_glue_detach(bt_glue *g) {
	if (!g) return;
	if (g->w) api.destroy(g->w);
	delete g;
}

/* ---- host -> device ---- */

static inline bool diff3(int32_t *sh4, const int32_t *v) { return sh4[0] != v[0] || sh4[1] != v[1] || sh4[2] != v[2]; }

int bt_glue_push(bt_glue *g, gamestate *gs) {
	g->pokes.clear();
	g->ctr.pushes++;
	/* (1) entities made since the last exchange: always at the head of gs->ents (entsLlAdd ent.cpp:47-52) */
	std::vector<ent *> fresh;
	for (ent *e = gs->ents; e; e = e->ll.n) {
		if (g->slot_of.count(e)) break;
		fresh.push_back(e);
	}
	std::reverse(fresh.begin(), fresh.end()); /* oldest first: slot order = creation order */
	int32_t first_new = (int32_t)g->slot_ent.size();
	if (!fresh.empty()) {
		std::vector<int32_t> ns(fresh.size()), wc(fresh.size());
		for (size_t i = 0; i < fresh.size(); i++) {
			ns[i] = fresh[i]->numSliders;
			wc[i] = fresh[i]->wires.num + fresh[i]->wiresAdd.num + 4;
		}
		int32_t first = 0;
		API(api.grow(g->w, (int32_t)fresh.size(), ns.data(), wc.data(), 0, &first));
		if (first != first_new) return fail(BT_ERR_ARG, "device handed out slot %d, the binding expected %d", first, first_new);
		for (size_t i = 0; i < fresh.size(); i++) {
			g->slot_of[fresh[i]] = first + (int32_t)i;
			g->slot_ent.push_back(fresh[i]);
		}
		/* the shadow of a fresh slot is what bt_world_grow left there */
		for (int a : SLOT_ARRAYS) if (int r = shadow_fetch(g, a, first)) return r;
		for (int a : {BT_A_SV, BT_A_SMIN, BT_A_SMAX}) if (int r = shadow_fetch(g, a, g->I(BT_A_SLIDER_OFF)[first])) return r;
		{ int32_t ext = 0; API(api.extent(g->w, BT_A_WIRE_TO, &ext)); g->sh[BT_A_WIRE_TO].resize(4 * (size_t)ext, 0); }
		g->ctr.spawned += (int64_t)fresh.size();
	}
	/* (2) gs->rootEnts order -> rootKey.  Roots the device knows keep their keys as long as the list is still
	 * in descending key order with every keyless root (a fresh entity, or one the host re-rooted) in front. */
	{
		std::vector<ent *> roots;
		for (ent *e = gs->rootEnts; e; e = e->LL.n) roots.push_back(e);
		bool ok = true, seen_keyed = false;
		uint64_t prev = ~0ull;
		size_t keyless = 0;
		for (ent *e : roots) {
			const int32_t s = g->slot(e);
			const bool keyed = s >= 0 && s < first_new && g->I(BT_A_HOLDER)[s] < 0 && g->Q(BT_A_ROOTKEY)[s] != 0;
			if (!keyed) { if (seen_keyed) ok = false; keyless++; continue; }
			seen_keyed = true;
			if (g->Q(BT_A_ROOTKEY)[s] >= prev) ok = false;
			prev = g->Q(BT_A_ROOTKEY)[s];
		}
		if (!ok || keyless) {
			const uint64_t phase = (uint64_t)api.next_root_phase(g->w) << 37;
			const size_t cnt = ok ? keyless : roots.size();
			if (!ok) g->ctr.rekeys++;
			for (size_t i = 0; i < cnt; i++) {
				const int32_t s = g->slot(roots[i]);
				const uint64_t key = phase | (uint64_t)(cnt - i);
				g->Q(BT_A_ROOTKEY)[s] = key;
				g->poke(BT_A_ROOTKEY, s, (int32_t)(uint32_t)key, (int32_t)(uint32_t)(key >> 32));
			}
		}
	}
	/* (3) field by field against the shadow */
	for (ent *e = gs->ents; e; e = e->ll.n) {
		const int32_t s = g->slot(e);
		const bool is_new = s >= first_new;
#define VEC(a, field) do { int32_t *p = g->I(a) + 4 * (size_t)s; if (diff3(p, e->field)) { p[0] = e->field[0]; p[1] = e->field[1]; p[2] = e->field[2]; g->poke(a, s, p[0], p[1], p[2]); } } while (0)
#define W32(a, val) do { const int32_t _v = (int32_t)(val); if (g->I(a)[s] != _v) { g->I(a)[s] = _v; g->poke(a, s, _v); } } while (0)
#define W8(a, val) do { const uint8_t _v = (uint8_t)(val); if (g->B(a)[s] != _v) { g->B(a)[s] = _v; g->poke(a, s, _v); } } while (0)
		VEC(BT_A_CENTER, center); VEC(BT_A_VEL, vel); VEC(BT_A_OLD, old); VEC(BT_A_DCENTER, d_center); VEC(BT_A_DVEL, d_vel); VEC(BT_A_RADIUS, radius);
		W32(BT_A_TYPE, e->typeMask); W32(BT_A_COLLIDE, e->collideMask); W32(BT_A_NTYPE, e->newTypeMask); W32(BT_A_NCOLLIDE, e->newCollideMask);
		W32(BT_A_FRICTION, e->friction);
		W32(BT_A_HOLDER, g->slot(e->holder)); W32(BT_A_HOLDFLAGS, e->holdFlags); W32(BT_A_HOLDROOT, g->slot(e->holdRoot));
		W32(BT_A_HOLDEE, g->slot(e->holdee));
		W32(BT_A_LLN, e->holder ? g->slot(e->LL.n) : -1); W32(BT_A_LLP, e->holder ? g->slot(e->LL.p) : -1);
		W32(BT_A_NEWHOLDER, g->slot(e->newHolder));
		if (e->newHolder) W32(BT_A_NEWHOLDFLAGS, e->newHoldFlags);
		W32(BT_A_UID, s);
		W32(BT_A_HPHYS, ids_phys(e)); W32(BT_A_HPAIR, ids_pair(e)); W32(BT_A_HTICK, ids_tick(e));
		W32(BT_A_BTN_V, pack_btn(e, 0)); W32(BT_A_BTN_V2, pack_btn(e, 1));
		W8(BT_A_NEWDROP, e->newDrop); W8(BT_A_DEADMAX0, e->dead_max[0]); W8(BT_A_DEADMAX1, e->dead_max[1]); W8(BT_A_FULLFLUSH, e->fullFlush);
#undef VEC
#undef W32
#undef W8
		const int32_t so = g->I(BT_A_SLIDER_OFF)[s];
		if (g->I(BT_A_SLIDER_OFF)[s + 1] - so != e->numSliders)
			return fail(BT_ERR_UNSUPPORTED, "entity in slot %d changed its slider count (%d -> %d): not representable without a new slot", s, g->I(BT_A_SLIDER_OFF)[s + 1] - so, e->numSliders);
		for (int k = 0; k < e->numSliders; k++) {
			const slider &sl = e->sliders[k];
			if (g->I(BT_A_SV)[so + k] != sl.v) { g->I(BT_A_SV)[so + k] = sl.v; g->poke(BT_A_SV, so + k, sl.v); }
			if (g->I(BT_A_SMIN)[so + k] != sl.min) { g->I(BT_A_SMIN)[so + k] = sl.min; g->poke(BT_A_SMIN, so + k, sl.min); }
			if (g->I(BT_A_SMAX)[so + k] != sl.max) { g->I(BT_A_SMAX)[so + k] = sl.max; g->poke(BT_A_SMAX, so + k, sl.max); }
		}
		/* e->wires, and room for what wiresAdd is about to add (flushMisc ent.cpp:694-698 runs on the device) */
		{
			int32_t wo = g->I(BT_A_WIRE_OFF)[s], cnt = g->I(BT_A_WIRE_CNT)[s], cap = g->I(BT_A_WIRE_CAP)[s];
			bool same = cnt == e->wires.num;
			for (int k = 0; same && k < cnt; k++) same = g->I(BT_A_WIRE_TO)[wo + k] == g->slot(e->wires.items[k]);
			const int32_t need = e->wires.num + e->wiresAdd.num;
			if (!same || need > cap) {
				std::vector<int32_t> t((size_t)e->wires.num);
				for (int k = 0; k < e->wires.num; k++) t[k] = g->slot(e->wires.items[k]);
				API(api.set_wires(g->w, s, t.data(), (int32_t)t.size(), e->wiresAdd.num + 2));
				API(api.pull(g->w, BT_A_WIRE_OFF, s, 1, &g->I(BT_A_WIRE_OFF)[s]));
				API(api.pull(g->w, BT_A_WIRE_CAP, s, 1, &g->I(BT_A_WIRE_CAP)[s]));
				g->I(BT_A_WIRE_CNT)[s] = (int32_t)t.size();
				wo = g->I(BT_A_WIRE_OFF)[s];
				if (g->sh[BT_A_WIRE_TO].size() < 4 * (size_t)(wo + g->I(BT_A_WIRE_CAP)[s])) g->sh[BT_A_WIRE_TO].resize(4 * (size_t)(wo + g->I(BT_A_WIRE_CAP)[s]), 0);
				for (size_t k = 0; k < t.size(); k++) g->I(BT_A_WIRE_TO)[wo + k] = t[k];
			}
		}
		/* wiresAdd / wiresRm (entUpdaters.c:66-73) move to the device's pending lists, in request order */
		for (int k = 0; k < e->wiresAdd.num; k++) g->poke(BT_OP_UWIRE, s, g->slot(e->wiresAdd.items[k]));
		for (int k = 0; k < e->wiresRm.num; k++) g->poke(BT_OP_UUNWIRE, s, g->slot(e->wiresRm.items[k]));
		e->wiresAdd.num = 0;
		e->wiresRm.num = 0;
		if (is_new) { g->poke(BT_OP_SPAWN, s, 0); g->B(BT_A_INLIST)[s] = 1; }
	}
	/* (4) gamestate scalars */
	if (g->I(BT_A_W_FLIP)[0] != gs->flipFlop_death) { g->I(BT_A_W_FLIP)[0] = gs->flipFlop_death; g->poke(BT_A_W_FLIP, 0, gs->flipFlop_death); }
	if (g->I(BT_A_W_GAMERULES)[0] != gs->gamerules) { g->I(BT_A_W_GAMERULES)[0] = gs->gamerules; g->poke(BT_A_W_GAMERULES, 0, gs->gamerules); }
	if (g->I(BT_A_W_RAND)[0] != gs->rand) { g->I(BT_A_W_RAND)[0] = gs->rand; g->poke(BT_A_W_RAND, 0, gs->rand); }
	g->ctr.pokes += (int64_t)g->pokes.size();
	if (!g->pokes.empty()) API(api.poke(g->w, g->pokes.data(), (int32_t)g->pokes.size()));
	return BT_OK;
}

/* ---- device -> host ---- */

/* killEntNoHandlers ent.cpp:190-208 as far as the HOST lists go (the device has done the fumbles; the
 * hold-tree pointers are rebuilt from its arrays right after) */
static void host_kill(gamestate *gs, ent *e, int dead) {
	entsLlRemove(gs, e);
	if (e->myBox) { velbox_remove(e->myBox); e->myBox = NULL; }
	e->ll.p = gs->deadTail;
	gs->deadTail = e;
	e->dead = (char)dead;
}

int bt_glue_pull(bt_glue *g, gamestate *gs) {
	g->ctr.pulls++;
	const std::vector<char> old_hphys = g->sh[BT_A_HPHYS], old_hpair = g->sh[BT_A_HPAIR], old_htick = g->sh[BT_A_HTICK];
	for (int a : SLOT_ARRAYS) if (int r = shadow_fetch(g, a, 0)) return r;
	for (int a : OTHER_ARRAYS) if (int r = shadow_fetch(g, a, 0)) return r;
	const int32_t n = (int32_t)g->slot_ent.size();
	/* (1) deaths, in the order the device's keys give them (flush deaths newest first, then
	 * pushes by iteration and descending root): invokeOnCrush walks gs->deadTail most recent first */
	{
		int32_t nd = 0;
		API(api.extent(g->w, BT_A_DEADLIST, &nd));
		if (nd > g->dead_seen) {
			std::vector<uint64_t> dl((size_t)nd);
			API(api.pull(g->w, BT_A_DEADLIST, 0, nd, dl.data()));
			std::vector<uint64_t> fresh(dl.begin() + g->dead_seen, dl.end());
			std::sort(fresh.begin(), fresh.end());
			for (uint64_t x : fresh) {
				const int32_t s = (int32_t)(uint32_t)x;
				ent *e = g->at(s);
				if (!e || e->dead) continue;
				host_kill(gs, e, g->B(BT_A_DEAD)[s] ? g->B(BT_A_DEAD)[s] : 1);
				g->ctr.deaths++;
			}
			g->dead_seen = nd;
		}
	}
	/* (2) fields */
	std::vector<std::pair<uint64_t, ent *>> roots;
	for (int32_t s = 0; s < n; s++) {
		ent *e = g->slot_ent[s];
		if (!e) continue;
		const int32_t *c = g->I(BT_A_CENTER) + 4 * (size_t)s, *v = g->I(BT_A_VEL) + 4 * (size_t)s, *o = g->I(BT_A_OLD) + 4 * (size_t)s;
		const int32_t *dc = g->I(BT_A_DCENTER) + 4 * (size_t)s, *dv = g->I(BT_A_DVEL) + 4 * (size_t)s;
		for (int k = 0; k < 3; k++) { e->center[k] = c[k]; e->vel[k] = v[k]; e->old[k] = o[k]; e->d_center[k] = dc[k]; e->d_vel[k] = dv[k]; }
		e->typeMask = g->U(BT_A_TYPE)[s]; e->collideMask = g->U(BT_A_COLLIDE)[s];
		e->newTypeMask = g->U(BT_A_NTYPE)[s]; e->newCollideMask = g->U(BT_A_NCOLLIDE)[s];
		e->holdFlags = g->I(BT_A_HOLDFLAGS)[s];
		for (int k = 0; k < 4; k++) {
			e->ctrl.btns[k].v = (char)((g->U(BT_A_BTN_V)[s] >> (8 * k)) & 0xff);
			e->ctrl.btns[k].v2 = (char)((g->U(BT_A_BTN_V2)[s] >> (8 * k)) & 0xff);
		}
		e->newDrop = (char)g->B(BT_A_NEWDROP)[s]; e->dead_max[0] = (char)g->B(BT_A_DEADMAX0)[s]; e->dead_max[1] = (char)g->B(BT_A_DEADMAX1)[s];
		e->fullFlush = (char)g->B(BT_A_FULLFLUSH)[s];
		e->newHolder = g->at(g->I(BT_A_NEWHOLDER)[s]);
		if (e->newHolder) e->newHoldFlags = g->I(BT_A_NEWHOLDFLAGS)[s];
		const int32_t so = g->I(BT_A_SLIDER_OFF)[s];
		for (int k = 0; k < e->numSliders; k++) { e->sliders[k].v = g->I(BT_A_SV)[so + k]; e->sliders[k].min = g->I(BT_A_SMIN)[so + k]; e->sliders[k].max = g->I(BT_A_SMAX)[so + k]; }
		/* handler ids only change on the device in a handful of places (decor_fumbled clears itself, common.c:61-67) */
		if (s * 4 + 4 <= (int32_t)old_hphys.size()) {
			const uint32_t hp = g->U(BT_A_HPHYS)[s], hq = g->U(BT_A_HPAIR)[s], ht = g->U(BT_A_HTICK)[s];
			if (hp != ((const uint32_t *)old_hphys.data())[s]) {
				e->whoMoves = whoMovesHandlers.get(hp & 0xff); e->push = pushHandlers.get((hp >> 8) & 0xff);
				e->pushed = ((hp >> 16) & 0xff) ? pushedHandlers.get((hp >> 16) & 0xff) : NULL;
				e->crush = (hp >> 24) ? crushHandlers.get(hp >> 24) : NULL;
			}
			if (hq != ((const uint32_t *)old_hpair.data())[s]) {
				e->onFumble = entPairHandlers.get(hq & 0xff); e->onFumbled = entPairHandlers.get((hq >> 8) & 0xff);
				e->onPickUp = entPairHandlers.get((hq >> 16) & 0xff); e->onPickedUp = entPairHandlers.get(hq >> 24);
			}
			if (ht != ((const uint32_t *)old_htick.data())[s]) {
				e->tick = (ht & 0xff) ? tickHandlers.get(ht & 0xff) : NULL;
				e->tickHeld = ((ht >> 8) & 0xff) ? tickHandlers.get((ht >> 8) & 0xff) : NULL;
			}
		}
		/* wires (flushMisc ent.cpp:694-712 ran on the device) */
		{
			const int32_t wo = g->I(BT_A_WIRE_OFF)[s], cnt = g->I(BT_A_WIRE_CNT)[s];
			bool same = cnt == e->wires.num;
			for (int k = 0; same && k < cnt; k++) same = g->at(g->I(BT_A_WIRE_TO)[wo + k]) == e->wires.items[k];
			if (!same) {
				e->wires.num = 0;
				for (int k = 0; k < cnt; k++) e->wires.add(g->at(g->I(BT_A_WIRE_TO)[wo + k]));
			}
		}
		/* hold tree */
		if (e->dead) { e->holder = NULL; e->holdee = NULL; continue; }
		e->holder = g->at(g->I(BT_A_HOLDER)[s]);
		e->holdee = g->at(g->I(BT_A_HOLDEE)[s]);
		e->holdRoot = g->at(g->I(BT_A_HOLDROOT)[s]);
		if (e->holder) { e->LL.n = g->at(g->I(BT_A_LLN)[s]); e->LL.p = g->at(g->I(BT_A_LLP)[s]); }
		else roots.push_back(std::make_pair(g->Q(BT_A_ROOTKEY)[s], e));
	}
	/* (3) gs->rootEnts (ent.cpp:60-66): newest root first = descending key */
	std::sort(roots.begin(), roots.end(), [](const std::pair<uint64_t, ent *> &a, const std::pair<uint64_t, ent *> &b) { return a.first > b.first; });
	gs->rootEnts = roots.empty() ? NULL : roots[0].second;
	for (size_t i = 0; i < roots.size(); i++) {
		roots[i].second->LL.p = i ? roots[i - 1].second : NULL;
		roots[i].second->LL.n = i + 1 < roots.size() ? roots[i + 1].second : NULL;
	}
	/* (4) gamestate scalars */
	gs->flipFlop_death = (byte)g->I(BT_A_W_FLIP)[0];
	gs->rand = g->I(BT_A_W_RAND)[0];
	gs->gamerules = g->I(BT_A_W_GAMERULES)[0];
	return BT_OK;
}

int bt_glue_physics(bt_glue *g, gamestate *gs) {
	if (int r = bt_glue_push(g, gs)) return r;
	API(api.step(g->w, BT_PHASE_PREP | BT_PHASE_GRAVITY | BT_PHASE_PHYSICS_ITER));
	/* the handlers whose effect lives on the host, at the point invokeOnCrush runs them (ent.cpp:907) */
	std::vector<bt_event> ev(256);
	int32_t nev = 0;
	int r = api.drain_events(g->w, ev.data(), (int32_t)ev.size(), &nev);
	if (r == BT_ERR_CAPACITY) { ev.resize((size_t)nev); r = api.drain_events(g->w, ev.data(), (int32_t)ev.size(), &nev); }
	if (r) return fail(r, "bt_world_drain_events -> %d: %s", r, api.last_error());
	if (nev) {
		if (int rr = bt_glue_pull(g, gs)) return rr;
		for (int32_t i = 0; i < nev; i++) {
			ent *me = g->at(ev[i].me_uid), *him = g->at(ev[i].him_uid);
			if (!me) continue;
			if (ev[i].kind == 0) { /* teamselect_push teamselect.cpp:36-45, loadblock_push loadblock.cpp:17-31 */
				if (him) (*me->push)(gs, me, him, 0, 0, 0, 0);
				g->ctr.events_push++;
			} else if (ev[i].kind == 3) { /* flag_crush flag.c:17-19, platform_crush platform.c:26-36 */
				if (me->crush) (*me->crush)(gs, me);
				g->ctr.events_crush++;
			}
		}
		if (int rr = bt_glue_push(g, gs)) return rr;
	}
	API(api.step(g->w, BT_PHASE_PHYSICS_END));
	return bt_glue_pull(g, gs);
}

int bt_glue_finish(bt_glue *g, gamestate *gs) {
	for (ent *e = gs->deadTail; e; e = e->ll.p) { /* clearDeads ent.cpp:471-481 is about to free them */
		auto it = g->slot_of.find(e);
		if (it == g->slot_of.end()) continue;
		g->slot_ent[it->second] = NULL;
		g->slot_of.erase(it);
	}
	finishStep(gs);
	API(api.step(g->w, BT_PHASE_FINISH));
	g->dead_seen = 0;
	return BT_OK;
}

int bt_glue_hash(bt_glue *g, uint64_t *hash, uint64_t *hash_full) {
	if (hash) API(api.hash(g->w, hash));
	if (hash_full) API(api.hash_full(g->w, hash_full));
	return BT_OK;
}
void bt_glue_host_hash(bt_glue *g, gamestate *gs, uint64_t *hash, uint64_t *hash_full) {
	hash_ents(gs, [&](const ent *e) -> int32_t { return g->slot(e); }, hash, hash_full);
}
