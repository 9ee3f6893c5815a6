"""ctypes front end of the CPU oracle (oracle/swim_oracle.c). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the swim_b200 package."""
import ctypes as C
import os
import subprocess

import numpy as np

from swim_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "swim_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "swim.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.POINTER(A.Config)]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_local_range.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
        L.oracle_set_view.argtypes = [vp, vp]
        L.oracle_inject.argtypes = [vp, vp, sz]
        L.oracle_round_begin.argtypes = [vp]
        L.oracle_round_end.argtypes = [vp]
        L.oracle_step.argtypes = [vp, u32]
        L.oracle_round.restype = u32
        L.oracle_round.argtypes = [vp]
        L.oracle_counters.argtypes = [vp, vp]
        L.oracle_env_words.restype = sz
        L.oracle_env_words.argtypes = [vp]
        L.oracle_outbox_count.restype = sz
        L.oracle_outbox_count.argtypes = [vp, u32]
        L.oracle_outbox_read.argtypes = [vp, u32, vp]
        L.oracle_inbox_add.argtypes = [vp, vp, sz]
        L.oracle_sent_count.restype = sz
        L.oracle_sent_count.argtypes = [vp]
        L.oracle_sent_read.argtypes = [vp, vp]
        L.oracle_array_bytes.restype = sz
        L.oracle_array_bytes.argtypes = [vp, C.c_int]
        L.oracle_get_array.argtypes = [vp, C.c_int, vp, sz]
        L.oracle_set_array.argtypes = [vp, C.c_int, vp, sz]
        L.oracle_digest.restype = u64
        L.oracle_digest.argtypes = [vp]
        L.oracle_mismatches.restype = u64
        L.oracle_mismatches.argtypes = [vp]
        L.oracle_get_members.argtypes = [vp, u32, vp, sz, C.POINTER(sz)]
        L.oracle_set_members.argtypes = [vp, u32, vp, sz]
        L.oracle_k_random_members.argtypes = [vp, u32, u32, vp, sz, vp, sz, C.POINTER(sz)]
        L.oracle_remove_dead_nodes.argtypes = [vp, u32]
        L.oracle_next_seqno.argtypes = [vp, u32, C.POINTER(u32)]
        L.oracle_next_incarnation.argtypes = [vp, u32, C.POINTER(u32)]
        L.oracle_apply_message.argtypes = [vp, u32, C.c_int, C.POINTER(A.Message), C.POINTER(A.Message),
                                           C.POINTER(C.c_int)]
        L.oracle_handle_message.argtypes = [vp, u32, u32, C.c_uint16, C.POINTER(A.Message), vp, sz,
                                            C.POINTER(sz)]
        L.oracle_broadcast.argtypes = [vp, u32, C.POINTER(A.Message)]
        L.oracle_get_broadcasts.argtypes = [vp, u32, vp, sz, C.POINTER(sz)]
        L.oracle_take_broadcasts.argtypes = [vp, u32, vp, sz, C.POINTER(sz)]
        L.oracle_tick_timers.argtypes = [vp, u32, C.POINTER(u32)]
        L.oracle_set_round.argtypes = [vp, u32]
        L.oracle_philox.argtypes = [vp, vp, vp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what}: oracle error {code}")
        self.code = code


def _chk(rc, what):
    if rc != 0:
        raise OracleError(rc, what)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox(c, k, o)
    return list(o)


class Oracle:
    """Same surface as swim_b200.sim.Simulator, backed by the C oracle."""

    def __init__(self, cfg: A.Config):
        self.cfg = cfg
        self._h = lib().oracle_create(C.byref(cfg))
        if not self._h:
            raise OracleError(A.EINVAL, "oracle_create")
        f, n = C.c_uint32(), C.c_uint32()
        lib().oracle_local_range(self._h, C.byref(f), C.byref(n))
        self.first, self.n_local = f.value, n.value

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    __del__ = close

    def set_view(self, nbr):
        nbr = np.ascontiguousarray(nbr, dtype=np.uint32)
        assert nbr.size == self.cfg.n_nodes * self.cfg.view_cap
        _chk(lib().oracle_set_view(self._h, nbr.ctypes.data), "set_view")

    def inject(self, events):
        events = np.ascontiguousarray(events, dtype=A.EVENT_DTYPE)
        _chk(lib().oracle_inject(self._h, events.ctypes.data, len(events)), "inject")

    def step(self, rounds=1):
        _chk(lib().oracle_step(self._h, rounds), "step")

    def round_begin(self):
        _chk(lib().oracle_round_begin(self._h), "round_begin")

    def round_end(self):
        _chk(lib().oracle_round_end(self._h), "round_end")

    def outbox(self, dst_rank):
        w = lib().oracle_env_words(self._h)
        c = lib().oracle_outbox_count(self._h, dst_rank)
        buf = np.zeros((c, w), dtype=np.uint32)
        if c:
            lib().oracle_outbox_read(self._h, dst_rank, buf.ctypes.data)
        return buf

    def sent(self):
        """Envelopes sent in the last round: [(src, dst, records)] with records as RECORD_DTYPE rows."""
        c = lib().oracle_sent_count(self._h)
        w = lib().oracle_env_words(self._h)
        buf = np.zeros((c, w), dtype=np.uint32)
        if c:
            lib().oracle_sent_read(self._h, buf.ctypes.data)
        out = []
        for row in buf:
            recs = row[4:4 + 4 * int(row[2])].copy().view(A.RECORD_DTYPE)
            out.append((int(row[1]), int(row[0]), recs))
        return out

    def inbox_add(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        if words.size:
            lib().oracle_inbox_add(self._h, words.ctypes.data, words.shape[0])

    @property
    def env_words(self):
        return lib().oracle_env_words(self._h)

    @property
    def round(self):
        return lib().oracle_round(self._h)

    def set_round(self, r):
        _chk(lib().oracle_set_round(self._h, r), "set_round")

    def counters(self):
        out = np.zeros(A.CTR_COUNT, dtype=np.uint64)
        lib().oracle_counters(self._h, out.ctypes.data)
        return out

    def get_array(self, arr):
        nb = lib().oracle_array_bytes(self._h, arr)
        dt = A.ARRAY_DTYPES[arr]
        out = np.zeros(nb // dt.itemsize, dtype=dt)
        _chk(lib().oracle_get_array(self._h, arr, out.ctypes.data, nb), "get_array")
        return out

    def set_array(self, arr, data):
        data = np.ascontiguousarray(data, dtype=A.ARRAY_DTYPES[arr])
        _chk(lib().oracle_set_array(self._h, arr, data.ctypes.data, data.nbytes), "set_array")

    def digest(self):
        return lib().oracle_digest(self._h)

    def mismatches(self):
        return lib().oracle_mismatches(self._h)

    # ---- scalar API
    def get_members(self, node):
        buf = (A.Member * self.cfg.view_cap)()
        n = C.c_size_t()
        _chk(lib().oracle_get_members(self._h, node, buf, self.cfg.view_cap, C.byref(n)), "get_members")
        return [_copy(buf[i]) for i in range(n.value)]

    def set_members(self, node, members):
        buf = (A.Member * max(1, len(members)))(*members)
        _chk(lib().oracle_set_members(self._h, node, buf, len(members)), "set_members")

    def k_random_members(self, node, n, excludes=()):
        ex = (A.Member * max(1, len(excludes)))(*excludes)
        buf = (A.Member * self.cfg.view_cap)()
        cnt = C.c_size_t()
        _chk(lib().oracle_k_random_members(self._h, node, n, ex, len(excludes), buf, self.cfg.view_cap,
                                           C.byref(cnt)), "k_random_members")
        return [_copy(buf[i]) for i in range(cnt.value)]

    def remove_dead_nodes(self, node):
        _chk(lib().oracle_remove_dead_nodes(self._h, node), "remove_dead_nodes")

    def next_seqno(self, node):
        o = C.c_uint32()
        _chk(lib().oracle_next_seqno(self._h, node, C.byref(o)), "next_seqno")
        return o.value

    def next_incarnation(self, node):
        o = C.c_uint32()
        _chk(lib().oracle_next_incarnation(self._h, node, C.byref(o)), "next_incarnation")
        return o.value

    def _apply(self, node, want, msg):
        out, has = A.Message(), C.c_int()
        _chk(lib().oracle_apply_message(self._h, node, want, C.byref(msg), C.byref(out), C.byref(has)), "apply")
        return out if has.value else None

    def suspect_node(self, node, msg):
        return self._apply(node, A.MSG_SUSPECT, msg)

    def dead_node(self, node, msg):
        return self._apply(node, A.MSG_DEAD, msg)

    def alive_node(self, node, msg):
        return self._apply(node, A.MSG_ALIVE, msg)

    def broadcast(self, node, msg):
        _chk(lib().oracle_broadcast(self._h, node, C.byref(msg)), "broadcast")

    def get_broadcasts(self, node):
        buf = (A.Message * A.MAX_PB)()
        n = C.c_size_t()
        _chk(lib().oracle_get_broadcasts(self._h, node, buf, A.MAX_PB, C.byref(n)), "get_broadcasts")
        return [_copy(buf[i]) for i in range(n.value)]

    def take_broadcasts(self, node):
        buf = (A.Message * A.MAX_PB)()
        n = C.c_size_t()
        _chk(lib().oracle_take_broadcasts(self._h, node, buf, A.MAX_PB, C.byref(n)), "take_broadcasts")
        return [_copy(buf[i]) for i in range(n.value)]

    def tick_timers(self, node):
        e = C.c_uint32()
        _chk(lib().oracle_tick_timers(self._h, node, C.byref(e)), "tick_timers")
        return e.value

    def handle_message(self, node, sender_addr, sender_port, msg):
        out = (A.Gossip * 4)()
        n = C.c_size_t()
        _chk(lib().oracle_handle_message(self._h, node, sender_addr, sender_port, C.byref(msg), out, 4,
                                         C.byref(n)), "handle_message")
        return [_copy(out[i]) for i in range(n.value)]


def _copy(s):
    t = type(s)()
    C.memmove(C.byref(t), C.byref(s), C.sizeof(s))
    return t


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))
