/*
 * bt_types.h — device-side world layout (structure of arrays, int32) and small helpers
 * shared by every kernel of the bittoss_b200 hot path.
 *
 * Layout in HBM (all arrays have `cap` slots; slot index = upload index, never reused):
 *   int4  center/old/vel/radius/d_center/d_vel   xyz + pad  -> one 128-bit load per field
 *   u32   type/collide/ntype/ncollide, i32 links (holder, holdee, LLn, LLp, holdRoot)
 *   u64   pend            one pending collision per hold root, packed (see bt_pack_coll)
 *   u8    flag arrays     needsCollision, hasBox, inList, dead, ...
 *   BvhNode[L-1]          64-byte nodes: both children's boxes + world range + child ids
 *
 * The bodies in bt_phys.cuh are written per logical thread and compile both for
 * sm_100a (the product) and, under BT_EMU, for the host so tests/emu can single-step
 * the exact same logic on a machine without a GPU.  BT_EMU is test tooling only; the
 * shipped library has no host execution path.
 */
#ifndef BT_TYPES_H
#define BT_TYPES_H

#include <stdint.h>

#ifdef BT_EMU
#define BT_HD
#define BT_D
#define BT_NOINLINE
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
#else
#include <cuda_runtime.h>
#define BT_HD __device__ __forceinline__
#define BT_D __device__ __forceinline__
#define BT_NOINLINE __device__ __noinline__
#endif

#include "../../include/bittoss_b200.h"

typedef unsigned long long u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint8_t u8;

/* ent.h:243-284 */
#define MOVE_ME 1
#define MOVE_HIM 2
#define MOVE_BOTH 7
#define MOVE_NONE 15
#define HOLD_PASS 0
#define HOLD_DROP 1
#define HOLD_MOVE 2
#define HOLD_MASK 3
#define HOLD_FREEZE 4
#define HOLD_SINGLE 8
#define T_HEAVY 1u
#define T_TERRAIN 2u
#define T_OBSTACLE 4u
#define T_DEBRIS 8u
#define T_WEIGHTLESS 16u
#define T_FLAG 32u
#define T_INPUTS (1u << 11)
#define TEAM_BIT 256u
#define TEAM_MASK (7u * TEAM_BIT)
#define T_COLLIDABLE (T_HEAVY + T_TERRAIN + T_OBSTACLE + T_DEBRIS + T_FLAG)
#define T_EQUIP (1u << 12)
#define T_EQUIP_SM (1u << 13)
#define EQUIP_MASK (T_EQUIP + T_EQUIP_SM)
#define EFFECT_GRAV (1 << 3)

/* 64-byte LBVH node: boxes of both children, so one node fetch decides two subtrees.
 * wl/wr pack the world-id range of each child as lo | hi<<16 (batched worlds). */
struct __attribute__((aligned(64))) BvhNode {
	i32 llo[3], lhi[3];
	i32 rlo[3], rhi[3];
	i32 left, right; /* >= 0 internal node, < 0 leaf position ~x */
	u32 wl, wr;
};

struct __attribute__((aligned(16))) Box6 { /* own box of a node or leaf, used only while refitting */
	int4 lo; /* xyz min, w = lowest world id */
	int4 hi; /* xyz max, w = highest world id */
};

/* device-side counters / flags block (one per bt_world) */
enum {
	CT_ANY_PROPOSAL = 0, /* doIteration's return value (ent.cpp:456,866): == iter_id of the iteration that proposed */
	CT_SERIAL_N,         /* entries in serialList */
	CT_DEAD_N,           /* entries in deadList */
	CT_ERROR,            /* BT_ERR_* raised on device */
	CT_NEED_REBUILD,     /* leaf set changed (an entity gained a box) */
	CT_ANY_DELTA,        /* some d_center / d_vel is non-zero */
	CT_FLUSH_WORK,       /* some world needs the serial part of flush1 */
	CT_N_LEAVES,
	/* CT_BOUNDS occupies 6 words */
	CT_BOUNDS,           /* 6 ints: min xyz, max xyz of leaf centres */
	CT_PAIR_N = 14,      /* unordered candidate pairs emitted by k_pairs (may exceed pair_cap) */
	CT_CACHE_INVALID = 15, /* an entity left the inflated box its cached pairs were found with */
	CT_STAT0 = 16,       /* (unused since the statistics moved to DevWorld::stat: 16 words kept free) */
	CT_FLUSH_HINT = 32,  /* something happened that flush1 has to act on (slider/mask request, a death) */
	CT_UNSUPPORTED = 33, /* times a handler without a device implementation would have run (lenient uploads only) */
	CT_UNSUP_LAST = 34,  /* kind<<8 | catalog id of the last such handler (0 push, 1 pushed, 2 entPair, 3 crush, 4 pickup flush) */
	CT_PAIR_BUILDS = 35, /* k_pairs runs that actually rebuilt the pair list */
	CT_PAIR_OVERFLOW = 36, /* k_pairs produced more pairs than pair_cap: the iteration's kernels stood down */
	CT_STOP = 37,        /* run-ahead: iter_id of the first speculatively enqueued iteration that stood down */
	CT_ANY_ODD = 38,     /* CT_ANY_PROPOSAL of odd iteration ids (even ones use slot 0): an iteration never
	                      * reads the slot it is itself stamping */
	CT_EVENT_N = 40,     /* entries in events[] */
	CT_CACHE_FULL = 41,  /* the next pair build must be a full one (upload, new tree, a jump, list grown stale) */
	CT_DIRTY_N = 42,     /* entries in dirtyList: entities that left their inflated box since the last build / patch */
	CT_PATCH_EPOCH = 43, /* counts consumed builds / patches; dirtyStamp[e] == epoch + 1 <=> e is in dirtyList */
	CT_PAIR_N_FULL = 44, /* CT_PAIR_N right after the last full build */
	CT_PAIR_PATCHES = 45, /* incremental pair-list patches (k_refit / k_pairs in patch mode) */
	CT_ORDER_TIES = 46,  /* full collisionBetter ties between different collisions (resolved by the canonical key) */
	CT_WQ_N = 47,        /* nodes handed out from the pending-wire pool (wq_next / wq_to) since the last flushMisc */
	CT_TREE_EDITS = 48,  /* fumbles + pickups + deaths so far: a host binding re-reads the hold tree only when it moved */
	CT_RAY_NEXT = 50,    /* k_raycast_dyn: the next unclaimed ray of the batch */
	CT_WORDS = 64
};
enum { SK_ITERS = 0, SK_CAND, SK_POS, SK_PUSH, SK_SERIAL, SK_CRUSH, SK_FUMBLE, SK_ETICKS, SK_NUM };
/* SK_ITERS / SK_ETICKS are only written by the one-block-per-world tick (k_tick_small): doIteration rounds summed over
 * the worlds, and live entities summed over the finished ticks; the generic schedule keeps both on the host */
/* The statistics are sums nobody on the device reads: they live in BT_STAT_ROWS rows, a block adds to row blockIdx.x mod
 * BT_STAT_ROWS, the host adds the rows up when asked.  One row for everybody -- one atomic per warp on the same seven
 * words -- cost the 1 M-cube world 13 % of its tick: 9 472 warps of k_narrow queue on one L2 atomic unit per word
 * (measured: k_narrow 394 -> 315 us, k_push 237 -> 220 us per tick without them). */
#define BT_STAT_ROWS 512

struct DevWorld {
	i32 n;        /* slots in use */
	i32 n_worlds;
	i32 n_leaves; /* leaves of the current tree (host copy) */
	i32 key_shift[3];
	i32 world_bits;
	i32 serial_cap, dead_cap;
	i32 iter_id; /* running doIteration counter, for event ordering keys */
	i32 iter_base; /* iter_id when the current tick's doPhysics began */
	i32 flush_stage; /* 0: a death now still gets its crush handler (ent.cpp:907); 1: the flush1 after invokeOnCrush */
	i32 lenient;  /* uploads with unknown handler IDs were accepted: count would-be invocations instead */
	i32 ray_mode;
	i32 margin;   /* inflation of the boxes candidate pairs are cached with (collision tree only) */
	i32 pair_cap; /* 1: tree over ALL live entities with static boxes (raycast), 0: collision tree */
	i32 horizon;  /* ticks of undisturbed motion (velocity + gravity) the cached box is stretched ahead by:
	               * velbox's validity idea (velbox/box.h:32-34), a box in free flight stays covered */
	i32 patch;    /* incremental pair-list patches allowed (option pair_patch) */
	i32 pdl;        /* bit 0: the short kernels are launched as programmatic dependents (release the successor, wait for
	                 * the predecessor); bit 1: so are the two long ones, k_narrow and k_push (they only wait) */
	i32 small_mode; /* every world holds <= BT_SMALL_WORLD entities: candidate pairs by brute force per world, no tree */
	i32 patch_shift; /* patch while dirty entities << patch_shift <= leaves, full build beyond */
	i32 root_phase;  /* launch-uniform: upper bits of the rootKey a fumble of this launch hands out (host counter) */
	i32 wq_cap;      /* nodes in the pending-wire pool */

	i32 *wbegin, *wend, *flip, *gamerules, *rnd; /* per world */
	u32 *wunsup;                                 /* per world: would-be invocations of unimplemented handlers */
	u32 *wties;                                  /* per world: CT_ORDER_TIES */
	u8 *wflush;                                  /* per world: serial flush1 work pending */

	i32 *world, *uid;
	int4 *center, *old, *vel, *radius, *d_center, *d_vel;
	u32 *type, *collide, *ntype, *ncollide;
	i32 *friction, *holder, *holdFlags, *holdRoot, *holdee, *LLn, *LLp, *newHolder, *newHoldFlags;
	u32 *hphys; /* whoMoves | push<<8 | pushed<<16 | crush<<24  (catalog IDs) */
	u32 *hpair; /* onFumble | onFumbled<<8 | onPickUp<<16 | onPickedUp<<24 */
	u32 *htick; /* tick | tickHeld<<8 (carried for the host) */
	u32 *btn_v, *btn_v2; /* 4 x u8 packed */
	u8 *newDrop, *dead_max0, *dead_max1, *fullFlush, *dead, *inList, *needsCollision, *hasBox;
	u8 *ready; /* 0 no, 1 push in parallel lane, 2 push in serial lane */
	u32 *forced;        /* 3 x int8 packed: x | y<<8 | z<<16 */
	int4 *forcedHoldees;
	u64 *pend;
	i32 *slider_off, *sv, *smin, *smax;
	i32 *wire_off, *wire_cap, *wire_cnt, *wire_to; /* block start / capacity / wires in use; one spare slot behind every
	                                                 * seat (PICKUP_SEAT wires its rider); blocks never move on the device */
	i32 *wireAdd, *wireRm;              /* pending uWire / uUnwire: -1 none, >= 0 the one target, <= -2 a list in the
	                                     * pool starting at node ~v - 1 ... see bt_uwire */
	i32 *wq_next, *wq_to;               /* pending-wire pool: singly linked lists in request order */
	u64 *rootKey;                       /* gs->rootEnts order (ent.cpp:60-66: newest root first) = descending key;
	                                     * phase << 37 | order inside the phase, see bt_fumble */
	u32 *wrootseq;                      /* per world: roots handed a key by the ordered (flush1) lane in this phase */

	/* broadphase */
	u64 *keys, *keys_alt;
	i32 *leafEnt, *leafEnt_alt; /* sorted leaf position -> entity */
	i32 *leafParent, *nodeParent;
	u32 *nodeCount;
	BvhNode *nodes;
	Box6 *nodeBox, *leafBox;
	i32 *nodeHi;     /* last leaf position below each internal node */
	i32 *wroot;      /* per world: root of its subtree (node id, ~leaf for a single leaf, INT32_MIN none) */
	Box6 *ibox;      /* per entity: the inflated swept box of the last refit */
	Box6 *iboxOld;   /* per entity: the box it had before the patch of the current epoch (dirty entities only) */
	i32 *leafOf;     /* entity -> sorted leaf position (inverse of leafEnt) */
	i32 *dirtyList;
	u32 *dirtyStamp;
	u64 *pairs;      /* a | b<<32, a = entity at the earlier leaf position */

	int4 *events;    /* {kind<<8 | handler, me slot, him slot, arg}; only the serial lane appends */
	i32 event_cap;
	i32 *serialList;
	u64 *deadList; /* key<<32 | entity */
	u64 *hash;     /* per world */
	u32 *ctr;      /* CT_WORDS */
	u64 *stat;     /* BT_STAT_ROWS x SK_NUM */
};

/* one deferred updater call made by host-side code (bt_world_apply_requests) */
#define BT_SMALL_WORLD 128

struct BtReq {
	i32 kind, me, other, a0, a1, a2; /* me / other are entity slots here */
};

/* ---- packed pending collision (the 5 collision* fields of ent.h:79-85) ----
 * bit 63 valid | buddy 29 | leaf 29 | axis 2 | dir 1 (1 = +1) | mutual 2
 * The tie-break key of DESIGN.md "Canonical order" is the value without the
 * mutual bits: larger (buddy, leaf, axis, dir) wins a full tie. */
struct Coll {
	i32 leaf, buddy;
	int axis, dir, mutual;
};
#define BT_MAX_ENTS (1 << 29)

static BT_HD u64 bt_pack_coll(const Coll &c) {
	return (1ull << 63) | ((u64)(u32)c.buddy << 34) | ((u64)(u32)c.leaf << 5) | ((u64)c.axis << 3) |
	       ((u64)(c.dir > 0 ? 1 : 0) << 2) | (u64)c.mutual;
}
static BT_HD Coll bt_unpack_coll(u64 p) {
	Coll c;
	c.buddy = (i32)((p >> 34) & 0x1fffffffu);
	c.leaf = (i32)((p >> 5) & 0x1fffffffu);
	c.axis = (int)((p >> 3) & 3);
	c.dir = ((p >> 2) & 1) ? 1 : -1;
	c.mutual = (int)(p & 3);
	return c;
}

static BT_HD int bt_get3(const int4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
static BT_HD void bt_set3(int4 &v, int k, int val) {
	if (k == 0) v.x = val;
	else if (k == 1) v.y = val;
	else v.z = val;
}
static BT_HD int bt_forced_get(u32 f, int axis) { return (int)(int8_t)((f >> (8 * axis)) & 0xff); }
static BT_HD u32 bt_forced_set(u32 f, int axis, int val) {
	return (f & ~(0xffu << (8 * axis))) | (((u32)(uint8_t)(int8_t)val) << (8 * axis));
}

static BT_HD u64 bt_mix64(u64 x) {
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
	x ^= x >> 27; x *= 0x94d049bb133111ebULL;
	x ^= x >> 31;
	return x;
}

#endif
