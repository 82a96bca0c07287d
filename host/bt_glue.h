/*
 * bt_glue.h — the binding a bittoss host compiles in to run its physics on the B200 library.
 *
 * Product code.  Compiled against the UNMODIFIED reference headers (ent.h, handlerRegistrar.h);
 * it moves a real `gamestate` (ent.h:182-194) into the C ABI of include/bittoss_b200.h and moves what
 * the device did back into the host's `ent` structures (ent.h:71-169), so that the reference's own
 * host code — tick handlers (doUpdates ent.cpp:823-832), world rules, players, crush handlers,
 * serialize() — keeps running unmodified on those structures, tick after tick, with the world resident
 * on the GPU.  Nothing is re-uploaded between ticks: after the first bt_glue_attach the host sends only
 * what ITS code changed (new entities, deferred-update requests, direct field writes) and reads back
 * what the DEVICE changed.
 *
 * In doWholeStep (main.cpp:1019-1098) the host replaces
 *
 *     prepPhysics(state);                                  // main.cpp:1081
 *     if (state->gamerules & EFFECT_GRAV) doGravity(state); // :1091
 *     doPhysics(state);                                    // :1093
 *     cleanupDeadHeroes(state);                            // :1096  (unchanged, host)
 *     finishStep(state);                                   // :1097
 * by
 *     bt_glue_physics(g, state);       // prepPhysics + doGravity + doPhysics, crush handlers run on the host
 *     cleanupDeadHeroes(state);
 *     bt_glue_finish(g, state);        // finishStep
 *
 * Every function returns 0 or a negative BT_ERR_* code (bt_glue_error() has the text).
 */
#ifndef BT_GLUE_H
#define BT_GLUE_H

#include <stdint.h>
#include "ent.h"
#include "../include/bittoss_b200.h"

struct bt_glue;

/* Binds the entry points of include/bittoss_b200.h.  so_path NULL: libbittoss_b200.so next to the library
 * this code was linked into (../bittoss_b200/ as laid out in this repository).  There is no CPU fallback:
 * without that library, or without a CUDA device, attach fails. */
int bt_glue_bind(const char *so_path);
const char *bt_glue_error(void);

/* gamestate -> bt_world_desc (replaces nothing in the reference: it is what `upload(gamestate*)` of
 * SURVEY §8b needs).  Arrays are owned by the returned object; index = rank from the tail of gs->ents
 * (oldest = 0), identical to the .sav index (serialize.cpp:79-94). */
struct bt_glue_desc;
bt_glue_desc *bt_glue_export(gamestate *gs);
const bt_world_desc *bt_glue_desc_get(const bt_glue_desc *d);
void bt_glue_desc_free(bt_glue_desc *d);

/* Creates the device world for gs (capacity for `max_ents` entities over its lifetime; slots of dead
 * entities are not reused) and uploads it.  From here on gs is stepped with bt_glue_physics. */
bt_glue *bt_glue_attach(gamestate *gs, int device, int32_t max_ents);
void bt_glue_detach(bt_glue *g);

/* host -> device: everything host code changed on the `ent`s since the last exchange. */
int bt_glue_push(bt_glue *g, gamestate *gs);
/* device -> host: positions, velocities, sliders, buttons, masks, hold tree, deaths, wires, flush state. */
int bt_glue_pull(bt_glue *g, gamestate *gs);
/* prepPhysics + doGravity + doPhysics (ent.cpp:834-911, effects.cpp:35-41) on the device.  The handlers
 * that act on host state are run HERE, on the host's own `ent`s, through the reference's function
 * pointers: teamselect / loadblock push handlers and the crush handlers, at the point of doPhysics where
 * invokeOnCrush (ent.cpp:907) runs them. */
int bt_glue_physics(bt_glue *g, gamestate *gs);
/* finishStep (ent.cpp:913-915): frees the dead `ent`s and retires their device slots. */
int bt_glue_finish(bt_glue *g, gamestate *gs);

/* bt_world_hash / bt_world_hash_full of the device world (one world). */
int bt_glue_hash(bt_glue *g, uint64_t *hash, uint64_t *hash_full);
/* The same two hashes computed on the HOST from the `ent`s, with the uids the binding assigned (creation
 * order): equal to the device's when host and device agree about the whole state. */
void bt_glue_host_hash(bt_glue *g, gamestate *gs, uint64_t *hash, uint64_t *hash_full);
/* Hashes of a gamestate that is not attached to anything (a pure CPU run of the reference): uids are
 * assigned in creation order by the tracker the caller keeps alive across ticks. */
struct bt_uid_tracker;
bt_uid_tracker *bt_uid_tracker_new(void);
void bt_uid_tracker_free(bt_uid_tracker *t);
void bt_uid_tracker_step(bt_uid_tracker *t, gamestate *gs); /* call before finishStep and after every creation point */
void bt_uid_tracker_forget_dead(bt_uid_tracker *t, gamestate *gs);
void bt_tracked_hash(bt_uid_tracker *t, gamestate *gs, uint64_t *hash, uint64_t *hash_full);

struct bt_glue_counters {
	int64_t pushes, pokes, spawned, pulls, deaths, events_push, events_crush, rekeys;
};
void bt_glue_get_counters(bt_glue *g, bt_glue_counters *out);
bt_world *bt_glue_world(bt_glue *g);

#endif
