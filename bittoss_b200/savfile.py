""".sav / .itm reader AND writer straight from / into the plain SoA arrays of `bt_world_desc`
(SURVEY.md §8f-3), without the reference's AoS `ent` round trip.  Restates `deserialize` / `deserializeEnt`
(serialize.cpp:277-334, 385-439, format notes :30-44): big-endian int32 stream "gam0".."gam7",
checksum bytes 128-133, handlers as catalog IDs behind a bitfield.  Host-side plumbing, no
physics; tests compare it field by field with the reference's own loader on all 17 shipped
levels.  `save_sav` restates `serialize` / `serializeEnt` (serialize.cpp:127-212) and is checked byte for
byte against the reference's own writer.  Players, gamestrings and colours are not on the physics path:
the reader hands them back in its info dict and the writer passes them through."""
import struct

import numpy as np

from .desc import WorldArrays

T_EQUIP, T_EQUIP_SM = 1 << 12, 1 << 13
EQUIP_MASK = T_EQUIP | T_EQUIP_SM
HOLD_MASK, HOLD_FREEZE = 3, 4
DEFAULT_FRICTION = 16
INT32_MIN, INT32_MAX = -(1 << 31), (1 << 31) - 1
VERSION = 7
# file order of the handler bitfield (serialize.cpp:34-43) -> slot in bt_world_desc.handlers
_FILE_TO_SLOT = [0, 1, 2, 3, 4, 5, 8, 9, 6, 7]   # whoMoves tick tickHeld crush push pushed onPickUp onPickedUp onFumble onFumbled
_CATALOG_SIZE = [9, 43, 43, 3, 10, 9, 7, 7, 7, 7]  # *_NUM of handlerRegistrar.h:79-187, in file order


class SavError(ValueError):
    pass


class _Reader:
    def __init__(self, data):
        self.d, self.i = data, 0

    def i32(self):  # read32 serialize.cpp:236-241: short reads give 0
        if self.i + 4 > len(self.d):
            return 0
        v = struct.unpack_from(">i", self.d, self.i)[0]
        self.i += 4
        return v

    def u8(self):
        if self.i >= len(self.d):
            return 0
        v = self.d[self.i]
        self.i += 1
        return v

    def str(self):  # readStr: length byte + that many bytes
        ln = self.u8()
        v = bytes(self.d[self.i: self.i + ln])
        self.i += ln
        return v

    def check(self, expected):  # checksum serialize.cpp:262-268 (the reference only warns)
        v = self.u8()
        if v != expected:
            raise SavError("expected checksum byte %d at offset %d, found %d" % (expected, self.i - 1, v))


def load_sav(path_or_bytes, rand_seed=12345):
    """-> (WorldArrays, info dict).  Entity index = file order = oldest first."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    if len(data) < 8 or data[:3] != b"gam":
        raise SavError("not a bittoss save (no 'gam' magic)")
    version = data[3] - ord("0")
    if not 0 <= version <= VERSION:
        raise SavError("unsupported save version %d" % version)
    r = _Reader(data)
    r.i = 4
    n = r.i32()
    if n < 0 or len(data) < n * 60:
        raise SavError("not enough data for %d entities" % n)
    ents = []
    for i in range(n):
        e = {}
        r.check(128)
        e["center"] = [r.i32() for _ in range(3)]
        e["vel"] = [r.i32() for _ in range(3)]
        e["radius"] = [r.i32() for _ in range(3)]
        r.check(129)
        ns = r.i32()
        e["type"] = r.i32() & 0xFFFFFFFF
        e["collide"] = r.i32() & 0xFFFFFFFF
        equip_fix = version < 7 and (e["type"] & EQUIP_MASK) == T_EQUIP
        if equip_fix:
            ns += 1
        n_wires = r.i32()
        r.check(130)
        e["color"] = r.i32()
        btn = r.u8() if version >= 1 else 0
        e["btn_v2"] = [(btn >> k) & 1 for k in range(4)]  # pushBtn: pressed for the coming tick
        e["sliders"] = [0 if (equip_fix and k == 0) else r.i32() for k in range(ns)]
        r.check(131)
        e["wires"] = []
        for _ in range(n_wires):
            w = r.i32()
            if 0 <= w < n:
                e["wires"].append(w)
        r.check(132)
        bits = r.i32() if version >= 4 else 63
        h = [0] * 10
        for k in range(10):
            if bits & (1 << k):
                hid = r.i32()
                if not 0 <= hid < _CATALOG_SIZE[k]:
                    hid = 0  # catalog<T>::get falls back to entry 0 (handlerRegistrar.h:48-59)
                h[_FILE_TO_SLOT[k]] = hid
        e["handlers"] = h
        r.check(133)
        e["friction"] = r.i32() if bits & (1 << 31) else DEFAULT_FRICTION
        e["holder"] = r.i32()
        e["hold_flags_raw"] = r.i32() if version >= 2 else 0
        ents.append(e)
    n_players = r.i32()
    players = []
    for _ in range(n_players):
        p = {"color": r.i32(), "revive": r.i32(), "data": r.i32() if version >= 3 else 0, "name": b""}
        if version >= 5:
            p["name"] = r.str()
        p["entity"] = r.i32()
        players.append(p)
    rand = r.i32()  # only taken over on a full-state load, serialize.cpp:430-432
    gamerules = r.i32()
    gamestrings = [r.str() for _ in range(r.u8())] if version >= 6 else []  # serialize.cpp:372-383

    ns_total = sum(len(e["sliders"]) for e in ents)
    nw_total = sum(len(e["wires"]) for e in ents)
    wa = WorldArrays(n, ns_total, nw_total, 1)
    a = wa.a
    a["w_gamerules"][0] = gamerules
    a["w_rand"][0] = rand_seed
    a["holdee_head"][:n] = -1
    a["holdee_next"][:n] = -1
    a["new_holder"][:n] = -1
    so = wo = 0
    for i, e in enumerate(ents):
        a["uid"][i] = i
        a["center"][3 * i: 3 * i + 3] = e["center"]
        a["vel"][3 * i: 3 * i + 3] = e["vel"]
        a["old"][3 * i: 3 * i + 3] = np.array(e["center"], dtype=np.int64) - np.array(e["vel"], dtype=np.int64)
        a["radius"][3 * i: 3 * i + 3] = e["radius"]
        a["type_mask"][i] = a["new_type_mask"][i] = e["type"]
        a["collide_mask"][i] = a["new_collide_mask"][i] = e["collide"]
        a["friction"][i] = e["friction"]
        a["handlers"][10 * i: 10 * i + 10] = e["handlers"]
        a["btn"][8 * i + 4: 8 * i + 8] = e["btn_v2"]
        a["slider_off"][i] = so
        for v in e["sliders"]:
            a["slider_v"][so], a["slider_min"][so], a["slider_max"][so] = v, INT32_MAX, INT32_MIN
            so += 1
        a["wire_off"][i] = wo
        for w in e["wires"]:
            a["wire_to"][wo] = w
            wo += 1
        # pickupNoHandlers (ent.cpp:166-182) as deserializeEnt calls it, in file order
        hd = e["holder"]
        a["holder"][i] = -1
        if hd != -1 and 0 <= hd < n:
            a["holder"][i] = hd
            a["holdee_next"][i] = a["holdee_head"][hd]      # prepend to the holder's list
            a["holdee_head"][hd] = i
            a["hold_flags"][i] = HOLD_MASK & e["hold_flags_raw"]
            if e["hold_flags_raw"] & HOLD_FREEZE:
                # the holder's vel as loaded SO FAR: a holder later in the file is still zeroed
                hv = ents[hd]["vel"] if hd < i else [0, 0, 0]
                a["d_vel"][3 * i: 3 * i + 3] = np.array(hv, dtype=np.int64) - np.array(e["vel"], dtype=np.int64)
    a["slider_off"][n] = so
    a["wire_off"][n] = wo
    # gs->rootEnts is prepend-ordered (ent.cpp:60-66): newest root first
    roots = [i for i in range(n) if a["holder"][i] < 0]
    a["root_order"][:n] = -1
    for pos, i in enumerate(reversed(roots)):
        a["root_order"][i] = pos
    return wa, {"version": version, "gamerules": gamerules, "n_players": n_players, "players": players, "rand": rand,
                "gamestrings": gamestrings, "colors": [e["color"] for e in ents]}


def save_sav(wa, colors=None, players=(), rand=0, gamerules=None, gamestrings=()):
    """WorldArrays -> the bytes `serialize` (serialize.cpp:187-212) writes for that state: "gam7",
    entity count, one `serializeEnt` record (:127-173) per entity in index order (= oldest first,
    the order `enumerateAll` numbers them, :87-93), players, rand, gamerules, gamestrings.
    Like the reference it writes flushed state only (no pending deltas).  `colors` (render only,
    not part of bt_world_desc), `players` (dicts as `load_sav` returns them) and `gamestrings`
    are passed through; entities of ONE world (the format has no world id)."""
    a, n = wa.a, wa.n
    if n and int(np.max(a["world_id"][:n])) != int(np.min(a["world_id"][:n])):
        raise SavError("save_sav writes one world; split the batch first")
    out = bytearray(b"gam%d" % VERSION)
    p32 = struct.Struct(">i").pack
    u32 = struct.Struct(">I").pack

    def wstr(b):  # writeStr serialize.cpp:68-77
        b = bytes(b)[:255]
        out.append(len(b))
        out.extend(b)

    out += p32(n)
    for i in range(n):
        out.append(128)
        for f in ("center", "vel", "radius"):
            for k in range(3):
                out += p32(int(a[f][3 * i + k]))
        out.append(129)
        s0, s1 = int(a["slider_off"][i]), int(a["slider_off"][i + 1])
        w0, w1 = int(a["wire_off"][i]), int(a["wire_off"][i + 1])
        out += p32(s1 - s0)
        out += u32(int(a["type_mask"][i]) & 0xFFFFFFFF)
        out += u32(int(a["collide_mask"][i]) & 0xFFFFFFFF)
        wires = [int(w) for w in a["wire_to"][w0:w1] if 0 <= int(w) < n]
        out += p32(len(wires))
        out.append(130)
        out += p32(int(colors[i]) if colors is not None else 0)
        out.append(sum(1 << k for k in range(4) if a["btn"][8 * i + k]))  # buttonByte: getButton = btns[k].v
        for k in range(s0, s1):
            out += p32(int(a["slider_v"][k]))
        out.append(131)
        for w in wires:
            out += p32(w)
        out.append(132)
        at = len(out)
        out += p32(0)
        bits = 0
        for k in range(10):  # processEnt serialize.cpp:30-44: only handlers other than catalog entry 0
            hid = int(a["handlers"][10 * i + _FILE_TO_SLOT[k]])
            if hid:
                out += p32(hid)
                bits |= 1 << k
        out.append(133)
        if int(a["friction"][i]) != DEFAULT_FRICTION:
            bits |= 1 << 31
            out += p32(int(a["friction"][i]))
        out[at: at + 4] = u32(bits)
        out += p32(int(a["holder"][i]) if 0 <= int(a["holder"][i]) < n else -1)
        out += p32(int(a["hold_flags"][i]))
    out += p32(len(players))
    for p in players:
        out += p32(p["color"]) + p32(p["revive"]) + p32(p["data"])
        wstr(p["name"])
        out += p32(p["entity"])
    out += p32(int(rand))
    out += p32(int(a["w_gamerules"][0]) if gamerules is None else int(gamerules))
    gs = list(gamestrings)
    while gs and not gs[-1]:
        gs.pop()  # serializeGamestrings :177-185: up to the last non-empty one
    out.append(len(gs))
    for g in gs:
        wstr(g)
    return bytes(out)
