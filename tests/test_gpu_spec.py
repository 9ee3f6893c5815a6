"""The reference's own unit tests (test/Spec.hs) restated against the CUDA implementation through the
host mirror of Core/Types/Util (swim_b200.core / .types / .util), plus the Appendix-E known-answer
vectors and a randomized scalar-API parity run against the oracle. Needs a GPU."""
import numpy as np
import pytest

from spec_fixture import ALIVE_ID, DEAD_ID, N_NODES, SELF, SUSPECT_ID, fixture, member, msg, view
from swim_b200 import _abi as A

pytestmark = pytest.mark.gpu


def make_members(host=0x7F000001):  # makeMembers (Spec.hs:45-56)
    from swim_b200.types import Liveness, Member, SockAddrInet
    seeds = [("alive", Liveness.IsAliveC, 4001), ("suspect", Liveness.IsSuspectC, 4002), ("dead", Liveness.IsDeadC, 4003)]
    return [Member(n, "127.0.0.1", SockAddrInet(p, host), st, 0, 0) for n, st, p in seeds]


@pytest.fixture
def store():  # withStore (Spec.hs:31-34)
    from swim_b200.util import configure
    return configure()


ADDR = None


def addr():
    from swim_b200.types import SockAddrInet
    return SockAddrInet(4000, 0x7F000001)  # Spec.hs:74


def test_remove_dead_nodes(store):  # Spec.hs:98-106
    from swim_b200 import core
    store.set_members(make_members())
    core.removeDeadNodes(store)
    mems = store.members_map()
    assert "dead" not in mems and len(mems) == 2


def test_k_random_members(store):  # Spec.hs:108-139
    from swim_b200 import core
    from swim_b200.types import Liveness, Member, SockAddrInet
    from swim_b200.util import configure
    ms = make_members()
    store.set_members(ms)
    assert core.kRandomMembers(store, 0, ms) == []                      # takes no nodes if n is 0
    rand = core.kRandomMembers(store, 3, [])                            # filters non-alive nodes
    assert len(rand) == 1 and rand[0] == ms[0]
    assert core.kRandomMembers(store, 3, [ms[0]]) == []                 # filters exclusion nodes
    stale = Member("alive", "127.0.0.1", ms[0].memberHostNew, Liveness.IsAliveC, 1, 0)
    assert core.kRandomMembers(store, 3, [stale]) == [ms[0]]            # structural Eq (Types.hs:68)
    big = configure(capacity=200)                                       # shuffles: 200 alive, n = 50
    alives = [Member(f"alive-{i}", "127.0.0.1", SockAddrInet(4001, 1), Liveness.IsAliveC, 0, 0) for i in range(200)]
    big.set_members(alives)
    rand = core.kRandomMembers(big, 50, [])
    assert len(rand) == 50 and len({m.memberName for m in rand}) == 50
    assert rand != alives and rand != sorted(alives)[:50]
    assert [m.memberName for m in core.members(big)] == sorted(m.memberName for m in alives)  # Map.elems order


def test_handle_udp_message(store):  # Spec.hs:141-183
    from swim_b200 import core
    from swim_b200.types import Ack, Direct, Envelope, IndirectPing, Ping, SockAddrInet, encode

    def send(m):
        return core.handleUDPMessage(store, [(encode(Envelope((m,))), addr())])

    assert send(Ping(1, "myself")) == [Direct(Ack(1, ()), addr())]       # gets Ping for us, responds with Ack
    assert send(Ping(1, "unknown-node")) == []                           # gets Ping for someone else
    assert send(Ack(1, ())) == []                                        # gets Ack (pending in the reference)
    before = store.incarnation                                           # gets IndirectPing, sends Ping
    ip = IndirectPing(1, addr().host, addr().port, "other")
    gossip = send(ip)
    assert store.incarnation == before + 1 and store.seqNo == 0
    assert gossip == [Direct(Ping(1, "other"), addr())]


def test_gets_suspect_dead_alive(store):
    """`pending` in the reference (Spec.hs:176-183): SURVEY Appendix E, through datagrams."""
    from swim_b200 import core
    from swim_b200.types import Alive, Broadcast, Dead, Envelope, Liveness, Suspect, encode
    store.set_members(make_members())

    def send(*ms):
        return core.handleUDPMessage(store, [(encode(Envelope(tuple(ms))), addr())])

    assert send(Suspect(0, "alive")) == [Broadcast(Suspect(0, "alive"))]                     # E1
    assert store.members_map()["alive"].memberAlive == Liveness.IsSuspectC
    assert send(Suspect(0, "suspect"), Suspect(5, "suspect"), Suspect(0, "dead"), Suspect(0, "nobody")) == []  # E2-E5
    assert store.members_map()["suspect"].memberIncarnation == 0
    assert send(Dead(0, "suspect", "x")) == [Broadcast(Dead(0, "suspect", "x"))]             # E7, deadFrom intact
    assert send(Dead(0, "dead", "x")) == []                                                  # E8
    # E13: refutation built from storeSelf's (swapped) SockAddrInet 123 4000 (Util.hs:97, Core.hs:160-166)
    assert send(Suspect(0, "myself")) == [Broadcast(Alive(1, "myself", 4000, 123))]
    assert store.incarnation == 1
    # E16: unknown Alive is inserted (Core.hs:206-216) and, per [Q7], re-broadcast; key order kept
    assert send(Alive(7, "bob", 99, 77)) == [Broadcast(Alive(7, "bob", 99, 77))]
    names = [m.memberName for m in core.members(store)]
    assert names == ["alive", "bob", "dead", "suspect"]
    bob = store.members_map()["bob"]
    assert (bob.memberHost, bob.memberHostNew.port, bob.memberHostNew.host, bob.memberIncarnation) == ("", 77, 99, 7)
    assert store.members_map()["alive"].memberAlive == Liveness.IsSuspectC  # untouched by the renumbering
    assert send(Alive(1, "alive", 0, 0)) == [Broadcast(Alive(1, "alive", 0, 0))]             # [Q7] i > j
    assert store.members_map()["alive"].memberAlive == Liveness.IsAliveC


def test_scalar_api_matches_oracle_on_random_sequences():
    """swim_suspect_node / swim_dead_node / swim_alive_node / swim_handle_message / kRandomMembers /
    removeDeadNodes on the device == the oracle, message by message."""
    import ctypes as C
    from oracle.oracle import Oracle, OracleError
    from swim_b200._lib import check, lib
    from swim_b200.sim import Simulator, default_config
    rng = np.random.default_rng(11)
    for cap in (32, 64):
        cfg = default_config(n_nodes=N_NODES * 4, view_cap=cap, suspicion_rounds=7)
        sim, orc = Simulator(cfg), Oracle(cfg)
        node = 100
        ms = [member(int(i), int(rng.integers(0, 3)), int(rng.integers(0, 4))) for i in rng.choice(90, 20, replace=False)]
        for m in ms:
            m.timer = 3 if m.liveness == A.SUSPECT else 0
        arr = (A.Member * len(ms))(*ms)
        check(lib().swim_set_members(sim._h, node, arr, len(ms)), "set", sim._h)
        orc.set_members(node, ms)
        fns = {A.MSG_SUSPECT: lib().swim_suspect_node, A.MSG_DEAD: lib().swim_dead_node, A.MSG_ALIVE: lib().swim_alive_node}
        for step in range(300):
            kind = int(rng.choice([A.MSG_SUSPECT, A.MSG_DEAD, A.MSG_ALIVE]))
            who = int(rng.choice([node, int(rng.integers(0, 90)), int(rng.integers(0, 90))]))
            if who == node and kind == A.MSG_ALIVE and rng.random() < 0.5:
                who = int(rng.integers(0, 90))
            m = msg(kind, who, int(rng.integers(0, 6)), dead_from=int(rng.integers(0, 90)))
            out, has = A.Message(), C.c_int()
            rc = fns[kind](sim._h, node, C.byref(m), C.byref(out), C.byref(has))
            try:
                exp = {A.MSG_SUSPECT: orc.suspect_node, A.MSG_DEAD: orc.dead_node, A.MSG_ALIVE: orc.alive_node}[kind](node, m)
            except OracleError as e:  # row full: both sides must fail the same way
                assert (rc, e.code) == (A.ECAP, A.ECAP), (step, kind, who, rc, e.code, len(orc.get_members(node)),
                                                          len(_members(sim, node, cap)))
                continue
            assert rc == 0, (step, kind, who, rc, lib().swim_last_error(sim._h))
            assert bool(has.value) == (exp is not None), (step, kind, who)
            if exp is not None:
                assert (out.kind, out.node, out.incarnation, out.dead_from) == (exp.kind, exp.node, exp.incarnation, exp.dead_from)
            if step % 50 == 49:
                check(lib().swim_remove_dead_nodes(sim._h, node), "rm", sim._h)
                orc.remove_dead_nodes(node)
            if step % 7 == 3:  # the per-period steps of a real-time node: countdown, piggyback payload (spends a transmission)
                e = C.c_uint32()
                check(lib().swim_tick_timers(sim._h, node, C.byref(e)), "tick", sim._h)
                assert e.value == orc.tick_timers(node), step
                mb, mc = (A.Message * A.MAX_PB)(), C.c_size_t()
                check(lib().swim_take_broadcasts(sim._h, node, mb, A.MAX_PB, C.byref(mc)), "take", sim._h)
                assert [(mb[i].kind, mb[i].node, mb[i].incarnation, mb[i].dead_from) for i in range(mc.value)] == \
                    [(x.kind, x.node, x.incarnation, x.dead_from) for x in orc.take_broadcasts(node)], step
            elif exp is not None:
                check(lib().swim_broadcast(sim._h, node, C.byref(out)), "bc", sim._h)
                orc.broadcast(node, exp)
            got = [(x.id, x.liveness, x.timer, x.incarnation, x.last_change) for x in _members(sim, node, cap)]
            want = [(x.id, x.liveness, x.timer, x.incarnation, x.last_change) for x in orc.get_members(node)]
            assert got == want, step
        # kRandomMembers: same Philox stream, same picks
        for n in (0, 1, 5, 64):
            buf = (A.Member * cap)()
            cnt = C.c_size_t()
            check(lib().swim_k_random_members(sim._h, node, n, None, 0, buf, cap, C.byref(cnt)), "krm", sim._h)
            assert [buf[i].id for i in range(cnt.value)] == [x.id for x in orc.k_random_members(node, n, [])]
        assert sim.get_array(A.ARR_SELF_INC)[node] == orc.get_array(A.ARR_SELF_INC)[node]
        # bulk rounds still work after the membership edits (in-edge index is rebuilt)
        sim.step(3)
        orc.step(3)
        assert sim.digest() == orc.digest()


def _members(sim, node, cap):
    import ctypes as C
    from swim_b200._lib import check, lib
    buf = (A.Member * cap)()
    n = C.c_size_t()
    check(lib().swim_get_members(sim._h, node, buf, cap, C.byref(n)), "get", sim._h)
    return [buf[i] for i in range(n.value)]


def test_disseminate(store):
    """disseminate (Core.hs:127-138) on the device: Direct -> datagram, Broadcast -> piggyback buffer, which a
    bulk round then carries to the ping target."""
    from oracle.oracle import Oracle
    from swim_b200 import core
    from swim_b200.types import Ack, Broadcast, Dead, Direct, Envelope, Ping, Suspect, decode
    store.set_members(make_members())
    wire = core.disseminate(store, [Direct(Ack(7, ()), addr()), Broadcast(Suspect(0, "alive")), Broadcast(Dead(2, "dead", "suspect")),
                                    Broadcast(Suspect(1, "alive"))])
    assert len(wire) == 1 and wire[0][1] == addr() and decode(wire[0][0]) == Envelope((Ack(7, ()),))
    assert core.pending_broadcasts(store) == [Suspect(1, "alive"), Dead(2, "dead", "suspect")]  # newest first, replaced
    # the same queue semantics as the oracle's
    import ctypes as C
    from swim_b200._lib import check, lib
    from swim_b200.sim import Simulator, default_config
    cfg = default_config(n_nodes=N_NODES, pb_cap=4)
    sim, orc = Simulator(cfg), Oracle(cfg)
    rng = np.random.default_rng(4)
    for _ in range(60):
        m = msg(int(rng.choice([A.MSG_SUSPECT, A.MSG_ALIVE, A.MSG_DEAD])), int(rng.integers(0, 9)), int(rng.integers(0, 5)),
                dead_from=int(rng.integers(0, 9)))
        check(lib().swim_broadcast(sim._h, SELF, C.byref(m)), "swim_broadcast", sim._h)
        orc.broadcast(SELF, m)
        assert np.array_equal(sim.get_array(A.ARR_PB), orc.get_array(A.ARR_PB))
        assert np.array_equal(sim.get_array(A.ARR_PB_CNT), orc.get_array(A.ARR_PB_CNT))
