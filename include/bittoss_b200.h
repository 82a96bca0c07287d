/*
 * bittoss_b200.h — C ABI of the B200-native bittoss physics hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * A bittoss host keeps its own `ent` / `gamestate` / `catalog<T>` (ent.h,
 * handlerRegistrar.h) and calls these entry points instead of the CPU tick
 * functions.  Every entry point names the reference interface it replaces
 * (file:line inside the sboerwinkle/bittoss tree).  INTEGRATION.md shows the
 * host-side binding.
 *
 * All functions return 0 (BT_OK) or a negative BT_ERR_* code; CUDA failures
 * return BT_ERR_CUDA and the text is available from bt_last_error().  There is
 * no CPU fallback: without a usable CUDA device every compute entry point
 * fails with BT_ERR_NO_DEVICE.
 */
#ifndef BITTOSS_B200_H
#define BITTOSS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_OK 0
#define BT_ERR_NO_DEVICE (-1)   /* no CUDA device / driver                        */
#define BT_ERR_CUDA (-2)        /* a CUDA call failed, see bt_last_error()        */
#define BT_ERR_ARG (-3)         /* bad argument (NULL, negative size, ...)        */
#define BT_ERR_CAPACITY (-4)    /* world does not fit the capacities given        */
#define BT_ERR_UNSUPPORTED (-5) /* handler id with no device-side implementation  */
#define BT_ERR_OVERFLOW (-6)    /* a device work list overflowed its buffer       */

/* Index of each handler slot inside bt_world_desc.handlers (10 bytes / entity).
 * Values are the catalog IDs of handlerRegistrar.h:79-187, i.e. exactly what
 * serialize.cpp:146-161 writes into a .sav file. */
enum {
	BT_H_WHOMOVES = 0, /* whoMovesHandlers  (handlerRegistrar.h:79-90)   */
	BT_H_TICK,         /* tickHandlers      (:93-138) host-side only     */
	BT_H_TICKHELD,     /* tickHandlers                                    */
	BT_H_CRUSH,        /* crushHandlers     (:141-146)                    */
	BT_H_PUSH,         /* pushHandlers      (:163-175)                    */
	BT_H_PUSHED,       /* pushedHandlers    (:149-160)                    */
	BT_H_ONFUMBLE,     /* entPairHandlers   (:178-187)                    */
	BT_H_ONFUMBLED,
	BT_H_ONPICKUP,
	BT_H_ONPICKEDUP,
	BT_H_NUM
};

/* bt_world_desc.misc bytes (4 / entity) */
enum { BT_M_NEWDROP = 0, BT_M_DEADMAX0, BT_M_DEADMAX1, BT_M_FULLFLUSH, BT_M_NUM };

/*
 * Host-side structure-of-arrays image of one or more gamestates
 * (`struct gamestate` ent.h:182-194, `struct ent_blueprint` ent.h:71-165).
 *
 * Entities are the LIVE ones, oldest first: index i is the entity's `.sav`
 * index (serialize.cpp:79-94), i.e. its rank counted from the tail of
 * gs->ents.  Every entity reference (holder, wires, ...) is such an index,
 * -1 for NULL.  vec3 arrays are xyz-interleaved, length 3*n_ents.
 * Several independent worlds may be concatenated (world_id non-NULL): all
 * entities of one world are contiguous and references stay inside a world.
 */
typedef struct bt_world_desc {
	int32_t n_ents;
	int32_t n_sliders;   /* total, == slider_off[n_ents] */
	int32_t n_wires;     /* total, == wire_off[n_ents]   */
	int32_t n_worlds;    /* >= 1                          */

	/* per world [n_worlds]  (gamestate scalars) */
	int32_t *w_flip_flop; /* gs->flipFlop_death */
	int32_t *w_gamerules; /* gs->gamerules (EFFECT_* bits, effects.h:8-17) */
	int32_t *w_rand;      /* gs->rand (carried, not used on the path)     */

	/* per entity [n_ents] unless noted */
	int32_t *world_id;    /* may be NULL when n_worlds == 1 */
	int32_t *uid;         /* caller-chosen stable id (unique inside a world); hashed */
	int32_t *center, *vel, *old, *d_center, *d_vel, *radius; /* [3*n_ents] */
	uint32_t *type_mask, *collide_mask, *new_type_mask, *new_collide_mask;
	int32_t *friction;
	int32_t *holder;      /* e->holder                                      */
	int32_t *hold_flags;  /* e->holdFlags                                   */
	int32_t *holdee_head; /* e->holdee (first of the holdee list)           */
	int32_t *holdee_next; /* e->LL.n when held (next sibling), else -1      */
	int32_t *root_order;  /* position in gs->rootEnts (0 = head) for roots, -1 when held */
	int32_t *new_holder;  /* e->newHolder                                   */
	int32_t *new_hold_flags;
	uint8_t *handlers;    /* [BT_H_NUM*n_ents] catalog IDs                  */
	uint8_t *btn;         /* [8*n_ents] ctrl.btns[0..3].v then [0..3].v2    */
	uint8_t *misc;        /* [BT_M_NUM*n_ents] newDrop, dead_max[0..1], fullFlush */
	int32_t *slider_off;  /* [n_ents+1] CSR offsets into slider_*           */
	int32_t *slider_v, *slider_min, *slider_max; /* [n_sliders]            */
	int32_t *wire_off;    /* [n_ents+1] CSR offsets into wire_to            */
	int32_t *wire_to;     /* [n_wires] e->wires[k] in list order            */
} bt_world_desc;

/* Phases of one simulation step, in reference order (main.cpp:1072-1097). */
#define BT_PHASE_CTRLS   1u  /* flushCtrls part of doUpdates   ent.cpp:817-824 (tick handlers stay on the host) */
#define BT_PHASE_PREP    2u  /* prepPhysics = flush1           ent.cpp:834,794-804 */
#define BT_PHASE_GRAVITY 4u  /* doGravity when EFFECT_GRAV     effects.cpp:35-41   */
#define BT_PHASE_PHYSICS 8u  /* doPhysics                      ent.cpp:842-911     */
#define BT_PHASE_FINISH 16u  /* finishStep = clearDeads        ent.cpp:913-915     */
#define BT_PHASE_ALL    31u
/* world rules of doWholeStep main.cpp:1030-1034 that are plain scans: doCrushtainer /
 * doBigCrushtainer / doLava / doBouncetainer (effects.cpp:43-96).  Runs before everything else.
 * Not part of BT_PHASE_ALL (a host that runs its own rules must not get them twice);
 * bt_world_tick adds it when a world's gamerules ask for one of them.  createDebris
 * (EFFECT_BLOCKS) creates entities and stays on the host. */
#define BT_PHASE_WORLD_RULES 32u
/* The tick handlers of doUpdates (ent.cpp:826-832) that have a device implementation (38 of the catalog's 42): the logic
 * family (logic.c: LOGIC, LOGIC_DEBUG, TIMER, ADJUST, DEMOLISH, RAND), DOOR (door.c), CURSED, SEAT / SEAT_OLD
 * (seat.cpp), SIGN, LOADBLOCK, INCINERATE, THRUST, SPIDER, PUDDLEJUMP (a pilot's controls are sliders and buttons of the
 * wired seat), the eyes (eyes.cpp), EXPLOSION, DUST, BULLET, HELD_BULLET, PLATFORM, BOTTLE; and for a world without
 * players RESPAWN and SCOREBOARD (the end of a round is handed to the host as an event of kind 6), PLAYER (player.c:229-443:
 * walking, jumping, looking, the equipment hands and drops), HELD_GUN / HELD_BLINK / HELD_JUMPER (cooldown and reload in
 * a player's hand), BOTTLE_HELD, LEGG / LEGG_OLD, FACTORY / FACTORY_WIRES / FLAG_SPAWNER -- each of them up to the moment
 * it would CREATE an entity (a shot, a stackem block, a platform, a product), which is counted as unsupported.
 * SPACETRASH, the editor's baubles and DEPLETE stay on the host.
 * Runs right after BT_PHASE_CTRLS.  Not part of BT_PHASE_ALL: a host that runs its own tick handlers must not get
 * them twice.  Fails with BT_ERR_UNSUPPORTED if the world holds any other tick handler (unless
 * the upload was lenient, then those count as unsupported events).  bt_world_tick adds the phase
 * by itself when every tick handler of the world is in that set (option "device_ticks", default
 * 1; 0 = never). */
#define BT_PHASE_TICKS 64u
/* doPhysics in two halves, for a host that runs the crush handlers itself (they create entities):
 * ..._ITER = flush2, integrate, broadphase and the whole resolution loop (ent.cpp:843-905);
 * ..._END  = what follows invokeOnCrush (ent.cpp:907): flush1; flush2 (ent.cpp:909-910).
 * BT_PHASE_PHYSICS == both, back to back.  Between the halves the host drains the events, runs
 * `crush` (and the TEAMSELECT / LOADBLOCK push handlers) on its own `ent`s and sends what they did. */
#define BT_PHASE_PHYSICS_ITER 128u
#define BT_PHASE_PHYSICS_END 256u

typedef struct bt_tick_stats {
	int64_t ticks;
	int64_t iterations;      /* doIteration rounds that found a collision (ent.cpp:866) */
	int64_t candidate_pairs; /* broadphase pairs handed to the narrow phase              */
	int64_t positives;       /* getAxisAndDir != 0 (ent.cpp:403-404)                     */
	int64_t pushes;          /* push() calls (ent.cpp:881)                               */
	int64_t serial_pushes;   /* of those, run in the serial tie/cycle lane               */
	int64_t crushes;         /* crushEnt during resolution (ent.cpp:211)                 */
	int64_t fumbles;         /* fumble() (ent.cpp:136)                                   */
	int64_t kernel_launches; /* CUDA kernels launched by this library                    */
	int64_t tree_builds;     /* full Morton sort + LBVH builds                           */
	int64_t tree_refits;     /* bottom-up refits                                         */
	int64_t live_ents;       /* live entities after the last tick                        */
	int64_t entity_ticks;    /* sum over ticks of live entities after the tick           */
	int64_t pair_builds;     /* full rebuilds of the cached pair list (tree refit + pair search) */
	int64_t pairs_cached;    /* unordered candidate pairs written by those launches      */
	int64_t pair_overflows;  /* ticks that fell back to per-iteration traversal          */
	int64_t unsupported_events; /* lenient uploads only ("allow_unsupported"): times a handler with no device
	                           * implementation would have been invoked; != 0 means the state diverged */
	int64_t pair_patches;    /* incremental updates of the cached pair list (moved entities only)  */
	int64_t order_ties;      /* full collisionBetter ties between different collisions: the reference keeps the
	                          * one its list order met first, the device the larger key */
} bt_tick_stats;

/* Per-kernel-class device timing, filled only while profiling is enabled. */
#define BT_PROF_MAX 24
typedef struct bt_profile {
	int32_t n;                      /* number of kernel classes */
	const char *name[BT_PROF_MAX];
	int64_t launches[BT_PROF_MAX];
	double ms[BT_PROF_MAX];         /* summed CUDA-event time */
} bt_profile;

/* A collision-path handler whose effect lives on the HOST side of the game (gs->players,
 * requestLoad, entity creation) is reported instead of executed: PUSH_TEAMSELECT
 * (teamselect.cpp:36-45), PUSH_LOADBLOCK (loadblock.cpp:17-31; its slider write is done on the
 * device) and the crush handlers CRUSH_FLAG / CRUSH_PLATFORM (flag.c:17-19: explode()).  Drain
 * after every step: push events come in the canonical push order, then the crush events in the
 * order invokeOnCrush (ent.cpp:464-469) calls the handlers, most recent death first. */
typedef struct bt_event {
	int32_t kind;     /* 0 = push handler, 3 = crush handler, 6 = tick handler (BT_PHASE_TICKS only) */
	int32_t handler;  /* catalog id (PUSH_TEAMSELECT = 4, PUSH_LOADBLOCK = 7; CRUSH_FLAG = 1, CRUSH_PLATFORM = 2;
	                   * TICK_SCOREBOARD = 27: the round ended, the host owes requestReload + showMessage, scoreboard.cpp:81-86) */
	int32_t world;
	int32_t me_uid;   /* the entity owning the handler */
	int32_t him_uid;  /* push: the entity that pushed it; crush: me_uid again */
	int32_t arg;      /* LOADBLOCK: getSlider(me, 0), the game-string index to load; crush: ordering key */
} bt_event;

/* One deferred updater call of entUpdaters.c made by code that stays on the host (tick
 * handlers, doInputs, edit commands).  Entities are named by (world, uid). */
enum {
	BT_REQ_DROP = 0,    /* uDrop(me)                               entUpdaters.c:28 */
	BT_REQ_DEAD,        /* uDead(me)                               :20 */
	BT_REQ_PICKUP,      /* uPickup(me, other, holdFlags = a[0])    :32 */
	BT_REQ_BUTTON,      /* pushBtn(me, ix = a[0])                  :6  */
	BT_REQ_VEL,         /* uVel(me, a)                             :14 */
	BT_REQ_CENTER,      /* uCenter(me, a)                          :8  */
	BT_REQ_SLIDER,      /* uSlider(me, ix = a[0], value = a[1])    :54 */
	BT_REQ_ERASE,       /* uErase(me)                              :24 */
	BT_REQ_TYPEMASK,    /* uMyTypeMask(me, a[0])                   :75 */
	BT_REQ_COLLIDEMASK, /* uMyCollideMask(me, a[0])                :80 */
	BT_REQ_NUM
};
typedef struct bt_request {
	int32_t kind, world, me_uid, other_uid;
	int32_t a[3];
} bt_request;

typedef struct bt_world bt_world; /* opaque, owns device memory + one stream */

/* ---- library ---- */
const char *bt_version(void);
const char *bt_last_error(void);
int bt_device_count(void); /* >=0, or BT_ERR_NO_DEVICE */

/* ---- world lifecycle: replaces mkGamestate/resetGamestate/doCleanup (ent.h:216-218),
 *      addEnt (ent.h:206) and velbox_init/getRoot/freeRoot (velbox/box.h:72-75) ---- */
int bt_world_create(bt_world **out, int device, int32_t max_ents, int32_t max_sliders,
                    int32_t max_wires, int32_t max_worlds);
void bt_world_destroy(bt_world *w);
/* dup(gs) (ent.h:222, ent.cpp:1135-1200; the rollback loop main.cpp:1265 clones rootState into
 * phantomState every frame): a second world in the same state, device-to-device.
 * bt_world_assign overwrites an existing world of identical capacities (no allocation). */
int bt_world_dup(bt_world *src, bt_world **out);
int bt_world_assign(bt_world *dst, bt_world *src);
/* Host AoS -> device SoA. Fails with BT_ERR_UNSUPPORTED if an entity names a
 * collision-path handler ID that has no device implementation. */
int bt_world_upload(bt_world *w, const bt_world_desc *desc);
/* Sizes needed to download the current live state. */
int bt_world_count(bt_world *w, int32_t *n_ents, int32_t *n_sliders, int32_t *n_wires);
/* Device SoA -> host arrays (caller-allocated, sized from bt_world_count; the
 * n_* fields are overwritten with what was written).  Dead entities are gone,
 * order is oldest-first as in bt_world_desc. */
int bt_world_download(bt_world *w, bt_world_desc *out);

/* ---- the tick: replaces doUpdates(flushCtrls)/prepPhysics/doPhysics/finishStep
 *      (ent.h:212-215, called from doWholeStep main.cpp:1072-1097) ---- */
int bt_world_step(bt_world *w, uint32_t phases);
/* n_ticks whole steps (BT_PHASE_ALL) with the world resident on the device.
 * hashes: NULL, or [n_ticks*n_worlds] receiving the state hash after each tick. */
int bt_world_tick(bt_world *w, int32_t n_ticks, uint64_t *hashes);
/* Same as bt_world_tick but only enqueues on the world's stream (no host sync,
 * no hash readback) so a caller can bracket it with its own CUDA events. */
int bt_world_tick_async(bt_world *w, int32_t n_ticks);
int bt_world_sync(bt_world *w);
/* n_ticks whole steps bracketed by CUDA events recorded on the world's own stream;
 * *ms receives the device time between them (what bench.py reports). */
int bt_world_tick_timed(bt_world *w, int32_t n_ticks, float *ms);

/* ---- per-tick host exchange for a host that keeps running its tick handlers
 *      (doUpdates, ent.cpp:823-832) on its own `ent` structs ----
 * bt_world_apply_host_deltas: what uCenter / uVel (entUpdaters.c:6-18) accumulated on the
 * host since the last tick, as [3*n] arrays in live oldest-first order; they are ADDED to
 * the device-side d_center / d_vel and consumed by the next flush2 (ent.cpp:716-744).
 * bt_world_read_state: centre and velocity of every live entity after the tick, same
 * order, for the next round of handlers / rendering. Host buffers; copies are issued on
 * the world's stream. */
int bt_world_apply_host_deltas(bt_world *w, const int32_t *d_center, const int32_t *d_vel, int32_t n_ents);
/* Sparse counterpart of bt_world_apply_host_deltas: the deferred updater calls host-side handlers
 * made this tick, applied in the order given before the next step.  Unknown (world, uid): BT_ERR_ARG. */
int bt_world_apply_requests(bt_world *w, const bt_request *reqs, int32_t n);
int bt_world_read_state(bt_world *w, int32_t *center, int32_t *vel, int32_t n_ents);
/* The same exchange, PIPELINED (for a host whose next deltas do not wait for the state it reads back -- input-driven
 * handlers, a renderer or recorder one tick behind): the copies run on two streams of their own, host -> device and
 * device -> host, so that the transfers of neighbouring ticks overlap each other and the tick in between (PCIe is
 * full duplex; a 1 M-entity world moves 24 MB each way per tick, about the time of the tick itself).
 *   bt_world_send_host_deltas: the copy starts at once and nothing else happens -- call it BEFORE the step the deltas
 *     must not reach (bt_world_step waits for its own tick, so a copy started after it cannot overlap it).
 *     bt_world_commit_host_deltas: the oldest batch sent is added on the world's stream as soon as it has arrived, in front
 *     of the next step enqueued; BT_ERR_ARG if an entity died since it was sent (its rows no longer line up: send
 *     again).  bt_world_apply_host_deltas_async = send + commit.  The host arrays (PINNED memory, or the copy is not
 *     asynchronous) must stay untouched until two more batches have been sent or bt_world_sync returned.
 *   bt_world_read_state_async: packs the state behind everything enqueued so far and starts the copy; *ticket (0 / 1)
 *     names it.  bt_world_read_state_wait(ticket) returns once `center` / `vel` hold it.  At most two reads in flight.
 * Both use double buffers on the device (4 x 12 B per entity more, allocated at the first call). */
int bt_world_send_host_deltas(bt_world *w, const int32_t *d_center, const int32_t *d_vel, int32_t n_ents);
int bt_world_commit_host_deltas(bt_world *w);
int bt_world_apply_host_deltas_async(bt_world *w, const int32_t *d_center, const int32_t *d_vel, int32_t n_ents);
int bt_world_read_state_async(bt_world *w, int32_t *center, int32_t *vel, int32_t n_ents, int32_t *ticket);
int bt_world_read_state_wait(bt_world *w, int32_t ticket);
void *bt_world_stream(bt_world *w); /* the cudaStream_t every kernel of w runs on */

/* Commutative 64-bit state hash per world over (uid, center, vel, typeMask,
 * collideMask, holder uid, holdFlags) of every live entity, plus live count.
 * out: [n_worlds]. */
int bt_world_hash(bt_world *w, uint64_t *out);
int bt_world_stats(bt_world *w, bt_tick_stats *out, int reset);
/* live entities after the last finished step, from the host-side bookkeeping: no device access, no synchronisation
 * (what a per-tick exchange loop sizes its buffers with) */
int64_t bt_world_live_ents(bt_world *w);
/* Copies out (and clears) the events queued since the last drain; *n = how many were written,
 * at most max.  Returns BT_ERR_OVERFLOW if the device queue (65536 entries) had overflowed. */
/* per world: full ties since the last stats reset (see bt_tick_stats.order_ties) */
int bt_world_order_ties(bt_world *w, uint32_t *per_world);
int bt_world_drain_events(bt_world *w, bt_event *out, int32_t max, int32_t *n);
/* per world [n_worlds]: unsupported_events since the last bt_world_stats(reset) */
int bt_world_unsupported(bt_world *w, uint32_t *per_world);
int bt_world_profile_enable(bt_world *w, int on);
int bt_world_profile_read(bt_world *w, bt_profile *out, int reset);
/* Knobs (integers).  None of them changes results, only how the work is scheduled:
 * "rebuild_every" ticks between full LBVH rebuilds (16); "pair_cache" 0/1 cached candidate pairs;
 * "cache_margin" inflation of the cached boxes (auto: 32, 128 in small-world mode); "cache_horizon" ticks of motion the cached box
 * is stretched ahead by (0); "pair_patch" 0/1 incremental pair-list patches; "patch_shift" patch
 * while dirty << shift <= leaves (4); "run_ahead" 0/1 speculative iterations; "max_iterations";
 * "device_ticks" 0/1 bt_world_tick runs the tick handlers itself when all have device bodies;
 * "pdl" programmatic dependent launches: 0 off, 1 short kernels only, 3 all kernels, -1 auto (all
 * for batched worlds and up to 786 k entities, short kernels only beyond); "small_worlds" 0/1 brute-force pair search per world
 * when every world holds <= 128 entities (read at upload); "narrow_blocks" / "narrow_threads" grid of
 * the narrow phase (148*8 x 256, the measured optimum).
 * "sort_rays" 0/1 traverse the rays of a batch in the order of their origin's leaf position (device radix sort; 0: no gain measured);
 * "allow_unsupported" 1 = bt_world_upload accepts handler IDs without a device implementation,
 * treats them as NIL and counts every would-be invocation in bt_tick_stats.unsupported_events
 * (used to measure handler coverage on real levels). */
int bt_world_set_option(bt_world *w, const char *name, int64_t value);

/* ---- incremental host binding: what host/bt_glue.cpp (the `gamestate*` <-> device bridge) is built on.
 *      The world stays resident; the host sends only what its own code changed and reads back what
 *      the device changed.  Entities are named by SLOT: the index an upload gave them (slot i =
 *      desc index i), or the one bt_world_grow returned.  Slots are never reused. ----
 *
 * Device arrays, by id.  Element type in brackets; "vec" = 4 x int32 per slot (xyz + pad).
 * Packed words: HPHYS = whoMoves | push<<8 | pushed<<16 | crush<<24, HPAIR = onFumble | onFumbled<<8 |
 * onPickUp<<16 | onPickedUp<<24, HTICK = tick | tickHeld<<8 (catalog ids, handlerRegistrar.h:79-187);
 * BTN_V / BTN_V2 = ctrl.btns[0..3].v / .v2, one byte each.  ROOTKEY: gs->rootEnts order, larger = nearer
 * the head (ent.cpp:60-66); fresh keys for host-made roots come from bt_world_next_root_phase. */
enum {
	BT_A_CENTER = 0, BT_A_VEL, BT_A_OLD, BT_A_DCENTER, BT_A_DVEL, BT_A_RADIUS,          /* vec           */
	BT_A_TYPE, BT_A_COLLIDE, BT_A_NTYPE, BT_A_NCOLLIDE,                                 /* u32 per slot  */
	BT_A_FRICTION, BT_A_HOLDER, BT_A_HOLDFLAGS, BT_A_HOLDROOT, BT_A_HOLDEE, BT_A_LLN, BT_A_LLP, /* i32   */
	BT_A_NEWHOLDER, BT_A_NEWHOLDFLAGS, BT_A_UID, BT_A_WORLD,                            /* i32           */
	BT_A_HPHYS, BT_A_HPAIR, BT_A_HTICK, BT_A_BTN_V, BT_A_BTN_V2,                        /* u32           */
	BT_A_NEWDROP, BT_A_DEADMAX0, BT_A_DEADMAX1, BT_A_FULLFLUSH, BT_A_DEAD, BT_A_INLIST, /* u8            */
	BT_A_ROOTKEY,                                                                       /* u64           */
	BT_A_SLIDER_OFF, BT_A_WIRE_OFF, BT_A_WIRE_CAP, BT_A_WIRE_CNT,                       /* i32 per slot  */
	BT_A_SV, BT_A_SMIN, BT_A_SMAX,                                                      /* i32 per slider */
	BT_A_WIRE_TO,                                                                       /* i32 per wire slot */
	BT_A_W_FLIP, BT_A_W_GAMERULES, BT_A_W_RAND,                                         /* i32 per world */
	BT_A_DEADLIST,                                                                      /* u64: key<<32 | slot, in no order */
	BT_A_NUM,
	/* poke-only operations (array field of a bt_poke) */
	BT_OP_SPAWN = 100,   /* index = fresh slot from bt_world_grow whose fields have been poked: it joins the world
	                      * (addEnt ent.cpp:107-113: in gs->ents, a root, boxed if needsVelbox) */
	BT_OP_UWIRE,         /* uWire(index, v[0])   entUpdaters.c:66: queued for the next flushMisc, order kept */
	BT_OP_UUNWIRE        /* uUnwire(index, v[0]) entUpdaters.c:71 */
};
/* one sparse write: v[0..k) into array `array` at element `index` (k = 3 for vec, 2 for u64 lo/hi, else 1).
 * No two writes of one call may name the same element; the BT_OP_* operations are applied after every
 * write of the call, in the order given. */
typedef struct bt_poke {
	int32_t array, index;
	int32_t v[3];
} bt_poke;
int bt_world_poke(bt_world *w, const bt_poke *pokes, int32_t n);
/* `count` elements of array `array` from element `first` into dst (host memory, the array's element type). */
int bt_world_pull(bt_world *w, int32_t array, int32_t first, int32_t count, void *dst);
/* elements currently in use of the array's index space: slots, sliders, wire slots, worlds, dead-list entries */
int bt_world_extent(bt_world *w, int32_t array, int32_t *count);
/* n_new blank slots behind the ones in use (addEnt's malloc): slider storage for n_sliders[i] sliders and
 * room for wire_cap[i] wires each.  *first_slot = slot of the first.  They stay out of the world until
 * BT_OP_SPAWN.  `world` = the world they belong to (0 for a single gamestate). */
int bt_world_grow(bt_world *w, int32_t n_new, const int32_t *n_sliders, const int32_t *wire_cap, int32_t world,
                  int32_t *first_slot);
/* replaces e->wires of one slot (list order kept); moves the block when it has outgrown its room */
int bt_world_set_wires(bt_world *w, int32_t slot, const int32_t *targets, int32_t n, int32_t reserve);
/* a fresh upper part for BT_A_ROOTKEY values: keys phase<<37 | k sort above every key handed out so far */
int64_t bt_world_next_root_phase(bt_world *w);
/* fumbles + pickups + deaths the device has done since the upload: unchanged = the hold tree did not move */
int bt_world_tree_edits(bt_world *w, uint32_t *count);
/* bt_world_hash widened to everything a `.sav` holds and the per-tick flush state: the words of bt_world_hash
 * plus radius, old, friction, sliders (v, min, max), buttons (v, v2), wire targets (uids, in list order),
 * handler ids, new*Mask, newHolder uid / flags, newDrop, dead_max, fullFlush, forced, and per world
 * flipFlop_death, rand, gamerules.  out: [n_worlds]. */
int bt_world_hash_full(bt_world *w, uint64_t *out);

/* ---- raycast: replaces selectCast(gs, parent, look) (raycast.h:1, raycast.cpp:16-57),
 *      batched.  origin_ent[i] = index of the casting entity (its hold tree is skipped),
 *      look = 3 ints per ray (NOT modified, unlike the reference which sign-folds it
 *      in place), winner[i] = entity index of the nearest hit or -1.  Indices refer to
 *      the live, oldest-first order of bt_world_download. Host buffers. ---- */
int bt_raycast_batch(bt_world *w, const int32_t *origin_ent, const int32_t *look, int32_t n_rays,
                     int32_t *winner);
/* Same with device-resident buffers, enqueued on the world's stream. */
int bt_raycast_batch_device(bt_world *w, const int32_t *d_origin_ent, const int32_t *d_look,
                            int32_t n_rays, int32_t *d_winner);
int bt_raycast_build(bt_world *w); /* (re)build the all-entity tree used by raycasts */

/* ---- synthetic worlds (BASELINE.json configs[2]): fills a caller-allocated desc.
 *      bt_synth_boxstack_size gives the array sizes. ---- */
int bt_synth_boxstack_size(int32_t n_cubes, int32_t height, int32_t *n_ents);
int bt_synth_boxstack_fill(bt_world_desc *desc, int32_t n_cubes, int32_t height, uint32_t seed,
                           int32_t push_every, int32_t push_max);

#ifdef __cplusplus
}
#endif
#endif
