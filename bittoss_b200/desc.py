"""ctypes image of `bt_world_desc` (include/bittoss_b200.h) backed by numpy arrays.

Host-side plumbing only: allocation of the plain SoA buffers that cross the
C ABI.  No physics here.
"""
import ctypes as C

import numpy as np

BT_H_NUM = 10
BT_M_NUM = 4

_I32P = C.POINTER(C.c_int32)
_U32P = C.POINTER(C.c_uint32)
_U8P = C.POINTER(C.c_uint8)


class BtWorldDesc(C.Structure):
    _fields_ = [
        ("n_ents", C.c_int32), ("n_sliders", C.c_int32), ("n_wires", C.c_int32), ("n_worlds", C.c_int32),
        ("w_flip_flop", _I32P), ("w_gamerules", _I32P), ("w_rand", _I32P),
        ("world_id", _I32P), ("uid", _I32P),
        ("center", _I32P), ("vel", _I32P), ("old", _I32P), ("d_center", _I32P), ("d_vel", _I32P),
        ("radius", _I32P),
        ("type_mask", _U32P), ("collide_mask", _U32P), ("new_type_mask", _U32P), ("new_collide_mask", _U32P),
        ("friction", _I32P), ("holder", _I32P), ("hold_flags", _I32P), ("holdee_head", _I32P),
        ("holdee_next", _I32P), ("root_order", _I32P), ("new_holder", _I32P), ("new_hold_flags", _I32P),
        ("handlers", _U8P), ("btn", _U8P), ("misc", _U8P),
        ("slider_off", _I32P), ("slider_v", _I32P), ("slider_min", _I32P), ("slider_max", _I32P),
        ("wire_off", _I32P), ("wire_to", _I32P),
    ]


# (field, dtype, elements per entity) — per-entity arrays
_ENT_FIELDS = [
    ("world_id", np.int32, 1), ("uid", np.int32, 1),
    ("center", np.int32, 3), ("vel", np.int32, 3), ("old", np.int32, 3), ("d_center", np.int32, 3),
    ("d_vel", np.int32, 3), ("radius", np.int32, 3),
    ("type_mask", np.uint32, 1), ("collide_mask", np.uint32, 1), ("new_type_mask", np.uint32, 1),
    ("new_collide_mask", np.uint32, 1),
    ("friction", np.int32, 1), ("holder", np.int32, 1), ("hold_flags", np.int32, 1),
    ("holdee_head", np.int32, 1), ("holdee_next", np.int32, 1), ("root_order", np.int32, 1),
    ("new_holder", np.int32, 1), ("new_hold_flags", np.int32, 1),
    ("handlers", np.uint8, BT_H_NUM), ("btn", np.uint8, 8), ("misc", np.uint8, BT_M_NUM),
]
_WORLD_FIELDS = ["w_flip_flop", "w_gamerules", "w_rand"]
_SLIDER_FIELDS = ["slider_v", "slider_min", "slider_max"]


class WorldArrays:
    """Owns numpy buffers for one bt_world_desc and keeps the ctypes struct in sync."""

    def __init__(self, n_ents, n_sliders=0, n_wires=0, n_worlds=1):
        self.cap = (int(n_ents), int(n_sliders), int(n_wires), int(n_worlds))
        self.a = {}
        for name, dt, k in _ENT_FIELDS:
            self.a[name] = np.zeros(max(1, n_ents * k), dtype=dt)
        for name in _WORLD_FIELDS:
            self.a[name] = np.zeros(max(1, n_worlds), dtype=np.int32)
        for name in _SLIDER_FIELDS:
            self.a[name] = np.zeros(max(1, n_sliders), dtype=np.int32)
        self.a["slider_off"] = np.zeros(n_ents + 1, dtype=np.int32)
        self.a["wire_off"] = np.zeros(n_ents + 1, dtype=np.int32)
        self.a["wire_to"] = np.zeros(max(1, n_wires), dtype=np.int32)
        self.desc = BtWorldDesc()
        self.desc.n_ents, self.desc.n_sliders, self.desc.n_wires, self.desc.n_worlds = self.cap
        for name, arr in self.a.items():
            ptr_t = dict(BtWorldDesc._fields_)[name]
            setattr(self.desc, name, arr.ctypes.data_as(ptr_t))

    def ref(self):
        return C.byref(self.desc)

    @property
    def n(self):
        return int(self.desc.n_ents)

    def vec(self, name):
        """(n,3) view of a vec3 field, trimmed to the live count."""
        return self.a[name][: 3 * self.n].reshape(-1, 3)

    def col(self, name, k=1):
        return self.a[name][: k * self.n].reshape(self.n, k) if k > 1 else self.a[name][: self.n]

    def state_dict(self):
        """Comparable dict of everything that defines the physics state (trimmed)."""
        n, ns, nw = self.n, int(self.desc.n_sliders), int(self.desc.n_wires)
        out = {}
        for name, _, k in _ENT_FIELDS:
            if name in ("world_id", "root_order", "holdee_next", "holdee_head"):
                continue
            out[name] = self.a[name][: n * k].copy()
        for name in _SLIDER_FIELDS:
            out[name] = self.a[name][:ns].copy()
        out["slider_off"] = self.a["slider_off"][: n + 1].copy()
        out["wire_off"] = self.a["wire_off"][: n + 1].copy()
        out["wire_to"] = self.a["wire_to"][:nw].copy()
        return out


_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_G = 0x9E3779B97F4A7C15
_MASK = (1 << 64) - 1


def _mix64(x):
    x = x.copy()
    x ^= x >> np.uint64(30)
    x *= np.uint64(_M1)
    x ^= x >> np.uint64(27)
    x *= np.uint64(_M2)
    x ^= x >> np.uint64(31)
    return x


def state_hash(wa: "WorldArrays"):
    """numpy restatement of the per-world state hash; one world only."""
    n = wa.n
    with np.errstate(over="ignore"):
        uid = wa.a["uid"][:n].astype(np.uint32).astype(np.uint64)
        h = _mix64(uid + np.uint64(_G))
        holder = wa.a["holder"][:n]
        huid = np.where(holder >= 0, wa.a["uid"][:n][np.clip(holder, 0, max(n - 1, 0))], -1).astype(np.int32)
        c, v = wa.vec("center"), wa.vec("vel")
        words = [c[:, 0], c[:, 1], c[:, 2], v[:, 0], v[:, 1], v[:, 2], wa.a["type_mask"][:n],
                 wa.a["collide_mask"][:n], huid, wa.a["hold_flags"][:n]]
        for w in words:
            w64 = np.asarray(w).astype(np.uint32).astype(np.uint64)
            h = _mix64(h + w64 + np.uint64(_G))
        total = int(h.sum(dtype=np.uint64)) if n else 0
        total = (total + int(_mix64(np.array([n], dtype=np.uint64))[0])) & _MASK
    return total
