"""ctypes binding of host/_build/libbthost.so — the headless bittoss host (the reference's own host code:
tick handlers, players, spawners, .sav loader / writer, compiled unmodified) whose physics run on the CUDA
library through host/bt_glue.cpp.  Built by `make -C host` from /root/reference when that tree is present.

There is no CPU fallback here either: `HeadlessHost()` binds libbittoss_b200.so; a different backend library
(the kernel-body emulator of tests/emu) can only be named explicitly, which only tests do.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST_DIR = os.path.join(ROOT, "host")
HOST_SO = os.path.join(HOST_DIR, "_build", "libbthost.so")
# the shipped levels (a plain copy of the reference's data/, made by host/Makefile): the host loads its .sav files
# from <DATA_PARENT>/data like the real client does from its working directory (file.cpp:10-17)
DATA_PARENT = os.environ.get("BT_HOST_DATA", os.path.join(HOST_DIR, "_build"))


def build():
    subprocess.check_call(["make", "-s", "-C", HOST_DIR])


def available():
    return os.path.exists(HOST_SO) and os.path.isdir(os.path.join(DATA_PARENT, "data"))


class HeadlessHost:
    """One per process (the reference keeps file-static state)."""
    _inst = None

    def __new__(cls, backend_so=None):
        if cls._inst is None:
            inst = super().__new__(cls)
            inst._load(backend_so)
            cls._inst = inst
        elif backend_so is not None and backend_so != cls._inst.backend_so:
            raise RuntimeError("the headless host is already bound to %s" % cls._inst.backend_so)
        return cls._inst

    def _load(self, backend_so):
        if not available():
            raise RuntimeError("host/_build/libbthost.so is not built (run `make -C host`)")
        L = C.CDLL(HOST_SO)
        L.bth_error.restype = C.c_char_p
        L.bth_init.argtypes = [C.c_char_p, C.c_char_p]
        L.bth_load_sav.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.c_int32]
        L.bth_set_options.argtypes = [C.c_int, C.c_uint32]
        L.bth_attach.argtypes = [C.c_int, C.c_int32]
        L.bth_step.argtypes = [C.c_char_p]
        L.bth_hash.argtypes = [C.POINTER(C.c_uint64)]
        L.bth_device_hash.argtypes = [C.POINTER(C.c_uint64)]
        L.bth_serialize.argtypes = [C.c_char_p, C.c_int]
        L.bth_counters.argtypes = [C.POINTER(C.c_int64)]
        L.bth_world.restype = C.c_void_p
        self.L = L
        self.backend_so = backend_so
        # the reference chats on stdout (file.cpp:16); keep a caller's machine-readable stdout clean
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            rc = L.bth_init(DATA_PARENT.encode(), backend_so.encode() if backend_so else None)
        finally:
            C.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
        if rc != 0:
            raise RuntimeError("bth_init failed: %s" % L.bth_error().decode())

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self.L.bth_error().decode()))

    def new_world(self):
        return self.L.bth_new_world()

    def select(self, wid):
        self._ck(self.L.bth_select(wid), "bth_select")

    def free_world(self, wid):
        self.L.bth_free_world(wid)

    def load_sav(self, name, players=(), rand_seed=12345):
        arr = (C.c_int32 * max(1, len(players)))(*players)
        self._ck(self.L.bth_load_sav(name.encode(), len(players), arr, rand_seed), "bth_load_sav(%s)" % name)

    def set_options(self, spacetrash_on=0, input_key=0):
        self.L.bth_set_options(spacetrash_on, input_key)

    def attach(self, device=0, max_ents=65536):
        self._ck(self.L.bth_attach(device, max_ents), "bth_attach")

    def step(self, inputs=None):
        """one doWholeStep; inputs = 6 bytes per player, None = the seeded stream of SURVEY §8(d)"""
        self._ck(self.L.bth_step(inputs), "bth_step")

    def hash(self):
        out = (C.c_uint64 * 2)()
        self.L.bth_hash(out)
        return int(out[0]), int(out[1])

    def device_hash(self):
        out = (C.c_uint64 * 2)()
        self._ck(self.L.bth_device_hash(out), "bth_device_hash")
        return int(out[0]), int(out[1])

    def serialize(self):
        cap = 1 << 20
        while True:
            buf = C.create_string_buffer(cap)
            n = self.L.bth_serialize(buf, cap)
            if n <= cap:
                return buf.raw[:n]
            cap = n

    def count(self):
        return int(self.L.bth_count())

    def counters(self):
        out = (C.c_int64 * 8)()
        self.L.bth_counters(out)
        keys = ["pushes", "pokes", "spawned", "pulls", "deaths", "events_push", "events_crush", "rekeys"]
        return dict(zip(keys, [int(x) for x in out]))
