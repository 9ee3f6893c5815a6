"""Host-side mirror of the reference's Core.hs API for ONE store, executed by the CUDA library.

A `Store` (Types.hs:53-60) is one simulated node of a small Simulator: member names are
interned to ids in ascending name order (== `Map.elems` order, Core.hs:77); every state
transition below runs on the device through the scalar C ABI (swim_suspect_node, ...).

  isAlive/isDead/notAlive Core.hs:33-40      nextSeqNo/nextIncarnation Core.hs:49-53
  removeDeadNodes Core.hs:65-67              kRandomMembers Core.hs:69-74
  members Core.hs:76-77                      handleUDPMessage/process Core.hs:79-121
  suspectNode/deadNode/aliveNode Core.hs:189-218
"""
import bisect
import ctypes as C
from typing import Iterable, List, Optional

from . import _abi as A
from ._lib import check, lib
from .sim import Simulator, default_config
from .types import (Ack, Alive, Broadcast, Config, Dead, Direct, Envelope, Gossip, IndirectPing, Liveness, Member,
                    Message, Ping, SockAddrInet, Suspect, decode)

_UNKNOWN = 0xFFFFFFFE  # id used for a name this store has never interned


def isAlive(m: Member) -> bool:
    return m.memberAlive == Liveness.IsAliveC


def isDead(m: Member) -> bool:
    return m.memberAlive == Liveness.IsDeadC


def notAlive(m: Member) -> bool:
    return not isAlive(m)


class Store:
    """storeSeqNo / storeIncarnation / storeMembers live in HBM; storeSelf and storeCfg here."""

    def __init__(self, self_member: Member, cfg: Config, capacity: int = 31, **sim_kw):
        view_cap = 32
        while view_cap < capacity:
            view_cap *= 2
        self.capacity = view_cap
        self.storeSelf = self_member
        self.storeCfg = cfg
        self._self = view_cap  # the store's own node index; members use ids [0, view_cap)
        self.sim = Simulator(default_config(n_nodes=view_cap + 1, view_cap=view_cap,
                                            k_indirect=min(cfg.numToGossip, A.MAX_K),
                                            fanout=1, **sim_kw))
        self._names: List[str] = []  # sorted; id == index
        self._meta = {}              # name -> (memberHost, memberHostNew) kept host-side

    # ---- name table -----------------------------------------------------------------
    def _id(self, name: str) -> int:
        if name == self.storeSelf.memberName:
            return self._self
        i = bisect.bisect_left(self._names, name)
        return i if i < len(self._names) and self._names[i] == name else _UNKNOWN

    def _name(self, ident: int) -> str:
        return self.storeSelf.memberName if ident == self._self else self._names[ident]

    def _to_c(self, m: Member) -> A.Member:
        c = A.Member()
        c.id = self._id(m.memberName)
        c.addr, c.port = c.id, self.sim.cfg.base_port
        c.liveness, c.incarnation, c.last_change = int(m.memberAlive), m.memberIncarnation, m.memberLastChange
        return c

    def _from_c(self, c: A.Member) -> Member:
        name = self._name(c.id)
        host, addr = self._meta.get(name, ("", SockAddrInet(self.sim.cfg.base_port, c.addr)))
        return Member(name, host, addr, Liveness(c.liveness), c.incarnation, c.last_change)

    def _h(self):
        return self.sim._h

    # ---- storeMembers -----------------------------------------------------------------
    def set_members(self, ms: Iterable[Member]):
        """`swapTVar storeMembers $ membersMap ms` (Spec.hs:101)."""
        ms = list(ms)
        by_name = {m.memberName: m for m in ms if m.memberName != self.storeSelf.memberName}
        self._names = sorted(by_name)
        self._meta = {n: (m.memberHost, m.memberHostNew) for n, m in by_name.items()}
        arr = (A.Member * max(1, len(by_name)))(*[self._to_c(by_name[n]) for n in self._names])
        check(lib().swim_set_members(self._h(), self._self, arr, len(by_name)), "swim_set_members", self._h())

    def _raw_members(self) -> List[A.Member]:
        buf = (A.Member * self.capacity)()
        n = C.c_size_t()
        check(lib().swim_get_members(self._h(), self._self, buf, self.capacity, C.byref(n)), "swim_get_members",
              self._h())
        return [buf[i] for i in range(n.value)]

    def members(self) -> List[Member]:
        """members (Core.hs:76-77): Map.elems, ascending by name."""
        return [self._from_c(c) for c in self._raw_members()]

    def members_map(self):
        return {m.memberName: m for m in self.members()}

    def _intern(self, name: str, host: str, addr: SockAddrInet) -> int:
        """Register a new name keeping ids in name order (ids above the insertion point shift)."""
        pos = bisect.bisect_left(self._names, name)
        raw = self._raw_members()
        for c in raw:
            if c.id >= pos:
                c.id += 1
                c.addr = c.id
        self._names.insert(pos, name)
        self._meta[name] = (host, addr)
        arr = (A.Member * max(1, len(raw)))(*raw)
        check(lib().swim_set_members(self._h(), self._self, arr, len(raw)), "swim_set_members", self._h())
        # records waiting in the piggyback buffer name members by id as well
        B = self.sim.cfg.pb_cap
        cnt = int(self.sim.get_array(A.ARR_PB_CNT)[self._self])
        if cnt:
            pb = self.sim.get_array(A.ARR_PB)
            mine = pb[self._self * B:self._self * B + cnt]
            for f in ("member", "from"):
                shift = (mine[f] >= pos) & (mine[f] < self._self)
                mine[f][shift] += 1
            self.sim.set_array(A.ARR_PB, pb)
        return pos

    # ---- counters -----------------------------------------------------------------------
    @property
    def seqNo(self) -> int:
        return int(self.sim.get_array(A.ARR_SEQNO)[self._self])

    @property
    def incarnation(self) -> int:
        return int(self.sim.get_array(A.ARR_SELF_INC)[self._self])

    def set_incarnation(self, v: int):
        a = self.sim.get_array(A.ARR_SELF_INC)
        a[self._self] = v
        self.sim.set_array(A.ARR_SELF_INC, a)


def nextSeqNo(store: Store) -> int:
    out = C.c_uint32()
    check(lib().swim_next_seqno(store._h(), store._self, C.byref(out)), "swim_next_seqno", store._h())
    return out.value


def nextIncarnation(store: Store) -> int:
    out = C.c_uint32()
    check(lib().swim_next_incarnation(store._h(), store._self, C.byref(out)), "swim_next_incarnation", store._h())
    return out.value


def removeDeadNodes(store: Store) -> None:
    check(lib().swim_remove_dead_nodes(store._h(), store._self), "swim_remove_dead_nodes", store._h())


def members(store: Store) -> List[Member]:
    return store.members()


def kRandomMembers(store: Store, n: int, excludes: Iterable[Member]) -> List[Member]:
    """kRandomMembers (Core.hs:69-74): alive members not structurally equal to an exclude,
    order-preserving shuffle (Util.hs:36-42) on the device, take n."""
    current = store.members()
    ex = [store._to_c(m) for m in excludes if m in current]  # `notElem`: derived Eq on all six fields
    arr = (A.Member * max(1, len(ex)))(*ex)
    buf = (A.Member * store.capacity)()
    cnt = C.c_size_t()
    check(lib().swim_k_random_members(store._h(), store._self, n, arr, len(ex), buf, store.capacity,
                                      C.byref(cnt)), "swim_k_random_members", store._h())
    return [store._from_c(buf[i]) for i in range(cnt.value)]


# ---- state machine -------------------------------------------------------------------------
def _msg_to_c(store: Store, m: Message) -> A.Message:
    c = A.Message()
    if isinstance(m, Suspect):
        c.kind, c.incarnation, c.node = A.MSG_SUSPECT, m.incarnation, store._id(m.node)
    elif isinstance(m, Dead):
        c.kind, c.incarnation, c.node = A.MSG_DEAD, m.incarnation, store._id(m.node)
        c.dead_from = store._id(m.deadFrom)
    elif isinstance(m, Alive):
        c.kind, c.incarnation, c.node, c.target, c.port = A.MSG_ALIVE, m.incarnation, store._id(m.node), m.addr, m.port
    elif isinstance(m, Ping):
        c.kind, c.seq_no, c.node = A.MSG_PING, m.seqNo, store._id(m.node)
    elif isinstance(m, IndirectPing):
        c.kind, c.seq_no, c.target, c.port, c.node = A.MSG_INDIRECT_PING, m.seqNo, m.target, m.port, store._id(m.node)
    elif isinstance(m, Ack):
        c.kind, c.seq_no = A.MSG_ACK, m.seqNo
    return c


def _apply(store: Store, fn, msg: Message) -> Optional[Message]:
    c, out, has = _msg_to_c(store, msg), A.Message(), C.c_int()
    check(fn(store._h(), store._self, C.byref(c), C.byref(out), C.byref(has)), fn.__name__, store._h())
    if not has.value:
        return None
    if out.kind == c.kind and out.node == c.node:
        return msg  # `return $ Just msg` (Core.hs:179): the identical message
    # the refutation (Core.hs:160-166): Alive built from storeSelf's SockAddrInet port host
    sa = store.storeSelf.memberHostNew
    return Alive(int(out.incarnation), store.storeSelf.memberName, sa.host, sa.port)


def suspectNode(store: Store, msg: Message) -> Optional[Message]:
    if not isinstance(msg, Suspect):
        raise TypeError("suspectNode _ _ = undefined (Core.hs:191)")
    return _apply(store, lib().swim_suspect_node, msg)


def deadNode(store: Store, msg: Message) -> Optional[Message]:
    if not isinstance(msg, Dead):
        raise TypeError("deadNode _ _ = undefined (Core.hs:195)")
    return _apply(store, lib().swim_dead_node, msg)


def aliveNode(store: Store, msg: Message) -> Optional[Message]:
    if not isinstance(msg, Alive):
        raise TypeError("aliveNode _ _ = undefined (Core.hs:218)")
    if msg.node != store.storeSelf.memberName and store._id(msg.node) == _UNKNOWN:
        # addNewMember (Core.hs:206-216): memberHost = "", memberHostNew = SockAddrInet port addr
        store._intern(msg.node, "", SockAddrInet(msg.port, msg.addr))
    return _apply(store, lib().swim_alive_node, msg)


def process(store: Store, sender: SockAddrInet, msg: Message) -> List[Gossip]:
    """`process` (Core.hs:89-117) through swim_handle_message."""
    if isinstance(msg, Alive) and msg.node != store.storeSelf.memberName and store._id(msg.node) == _UNKNOWN:
        store._intern(msg.node, "", SockAddrInet(msg.port, msg.addr))
    c = _msg_to_c(store, msg)
    out = (A.Gossip * 2)()
    n = C.c_size_t()
    check(lib().swim_handle_message(store._h(), store._self, sender.host, sender.port, C.byref(c), out, 2,
                                    C.byref(n)), "swim_handle_message", store._h())
    res: List[Gossip] = []
    for i in range(n.value):
        g = out[i]
        if g.is_direct:
            if g.msg.kind == A.MSG_ACK:
                res.append(Direct(Ack(g.msg.seq_no, ()), SockAddrInet(g.dest_port, g.dest_addr)))
            else:  # the forwarded Ping of an IndirectPing: `node'` is passed through verbatim
                res.append(Direct(Ping(g.msg.seq_no, msg.node), SockAddrInet(g.dest_port, g.dest_addr)))
        elif g.msg.kind == c.kind and g.msg.node == c.node:
            res.append(Broadcast(msg))
        else:
            sa = store.storeSelf.memberHostNew
            res.append(Broadcast(Alive(int(g.msg.incarnation), store.storeSelf.memberName, sa.host, sa.port)))
    return res


def _msg_from_c(store: Store, c: A.Message) -> Message:
    name = store._name(c.node) if c.node == store._self or c.node < len(store._names) else f"#{c.node}"
    if c.kind == A.MSG_SUSPECT:
        return Suspect(int(c.incarnation), name)
    if c.kind == A.MSG_DEAD:
        frm = store._name(c.dead_from) if c.dead_from == store._self or c.dead_from < len(store._names) else f"#{c.dead_from}"
        return Dead(int(c.incarnation), name, frm)
    if name == store.storeSelf.memberName:  # our own announcement / refutation carries our real address (Core.hs:160-166)
        addr = store.storeSelf.memberHostNew
    else:
        addr = store._meta.get(name, ("", SockAddrInet(c.port, c.target)))[1]
    return Alive(int(c.incarnation), name, addr.host, addr.port)


def pending_broadcasts(store: Store) -> List[Message]:
    """The store's piggyback buffer, newest first (what the next compound Envelope would carry)."""
    buf = (A.Message * A.MAX_PB)()
    n = C.c_size_t()
    check(lib().swim_get_broadcasts(store._h(), store._self, buf, A.MAX_PB, C.byref(n)), "swim_get_broadcasts", store._h())
    return [_msg_from_c(store, buf[i]) for i in range(n.value)]


def take_broadcasts(store: Store) -> List[Message]:
    """The piggyback payload of the next outgoing message: the buffer (newest first), after which one of each record's
    `retransmit` transmissions is spent (phase T4 of DESIGN.md 2.2; the compound Envelope `disseminate`'s FIXME,
    Core.hs:136, never builds)."""
    buf = (A.Message * A.MAX_PB)()
    n = C.c_size_t()
    check(lib().swim_take_broadcasts(store._h(), store._self, buf, A.MAX_PB, C.byref(n)), "swim_take_broadcasts", store._h())
    return [_msg_from_c(store, buf[i]) for i in range(n.value)]


def tickTimers(store: Store) -> int:
    """One protocol period of the suspicion countdown (`-- FIXME: need a timer to mark this node as dead after suspect
    timeout`, Core.hs:141): expired Suspect entries become Dead and their Dead(inc, member, from = self) is enqueued.
    Returns how many expired."""
    e = C.c_uint32()
    check(lib().swim_tick_timers(store._h(), store._self, C.byref(e)), "swim_tick_timers", store._h())
    return e.value


def disseminate(store: Store, gossip: Iterable[Gossip]):
    """disseminate (Core.hs:127-138): `Direct msg addr` -> a datagram to send now (framed as an Envelope — the
    reference sends a bare `encode msg` that its own receiver cannot decode); `Broadcast msg` ->
    the piggyback buffer (the reference's `enqueue _msg = return ()` FIXME). Returns [(bytes, addr)]."""
    from .types import encode
    out = []
    for g in gossip:
        if isinstance(g, Direct):
            out.append((encode(Envelope((g.msg,))), g.addr))
        else:
            if isinstance(g.msg, Alive) and g.msg.node != store.storeSelf.memberName and store._id(g.msg.node) == _UNKNOWN:
                store._intern(g.msg.node, "", SockAddrInet(g.msg.port, g.msg.addr))
            c = _msg_to_c(store, g.msg)
            check(lib().swim_broadcast(store._h(), store._self, C.byref(c)), "swim_broadcast", store._h())
    return out


def handleUDPMessage(store: Store, datagrams) -> List[Gossip]:
    """handleUDPMessage (Core.hs:79-121) over a list of (bytes, sender) datagrams: decode the
    Envelope, process every message, concatenate the Gossip. A decode failure raises (the
    reference's `handleDecodeErrors = either fail yield`, Core.hs:86-87)."""
    out: List[Gossip] = []
    for data, sender in datagrams:
        for m in decode(data).unEnvelope:
            out.extend(process(store, sender, m))
    return out
