"""Pins the CPU oracle: Philox known-answer vectors, every live assertion of the reference's own
test/Spec.hs (lines cited per test), and the hand-derived known-answer vectors E1..E22 of
SURVEY Appendix E for the state machine the reference leaves `pending` (Spec.hs:176-183)."""
import numpy as np
import pytest

from oracle.oracle import Oracle, OracleError, philox
from spec_fixture import ALIVE_ID, DEAD_ID, N_NODES, SELF, SUSPECT_ID, fixture, member, msg, view
from swim_b200 import _abi as A
from swim_b200.sim import default_config


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors (Salmon et al. SC'11), philox4x32 10 rounds
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.fixture
def store():
    o = Oracle(default_config(n_nodes=N_NODES, suspicion_rounds=5))
    o.set_members(SELF, fixture())
    return o


# ---------------------------------------------------------------- test/Spec.hs restated
def test_spec_remove_dead_nodes(store):  # Spec.hs:98-106
    store.remove_dead_nodes(SELF)
    v = view(store)
    assert DEAD_ID not in v and len(v) == 2


def test_spec_k_random_zero(store):  # Spec.hs:111-115
    assert store.k_random_members(SELF, 0, fixture()) == []


def test_spec_k_random_filters_non_alive(store):  # Spec.hs:117-122
    got = store.k_random_members(SELF, 3, [])
    assert len(got) == 1 and got[0].id == ALIVE_ID


def test_spec_k_random_exclusion(store):  # Spec.hs:124-128 + structural Eq (Types.hs:68)
    alive = store.get_members(SELF)[0]
    assert store.k_random_members(SELF, 3, [alive]) == []
    stale = member(ALIVE_ID, A.ALIVE, inc=1)  # E18: a stale copy does not exclude
    assert [m.id for m in store.k_random_members(SELF, 3, [stale])] == [ALIVE_ID]


def test_spec_k_random_shuffles():  # Spec.hs:130-139: 200 alive members, n = 50
    o = Oracle(default_config(n_nodes=512, view_cap=256))
    o.set_members(300, [member(i, A.ALIVE) for i in range(200)])
    got = [m.id for m in o.k_random_members(300, 50, [])]
    assert len(got) == 50 and len(set(got)) == 50
    assert got != list(range(50))
    assert [m.id for m in o.k_random_members(300, 50, [])] != got  # a fresh draw per call


def test_spec_ping_for_us(store):  # Spec.hs:150-153
    g = store.handle_message(SELF, 0x7F000001, 4000, msg(A.MSG_PING, SELF, seq=1))
    assert len(g) == 1 and g[0].is_direct == 1 and g[0].msg.kind == A.MSG_ACK and g[0].msg.seq_no == 1
    assert g[0].msg.payload_len == 0 and (g[0].dest_addr, g[0].dest_port) == (0x7F000001, 4000)


def test_spec_ping_for_someone_else(store):  # Spec.hs:155-158
    assert store.handle_message(SELF, 1, 4000, msg(A.MSG_PING, 7, seq=1)) == []


def test_spec_ack_emits_nothing(store):  # Spec.hs:160-164 (pending in the reference; Core.hs:92-94)
    assert store.handle_message(SELF, 1, 4000, msg(A.MSG_ACK, 0, seq=1)) == []


def test_spec_indirect_ping(store):  # Spec.hs:166-174 (Q4: seq := new incarnation)
    before = store.get_array(A.ARR_SELF_INC)[SELF]
    g = store.handle_message(SELF, 1, 4000, msg(A.MSG_INDIRECT_PING, 9, seq=1, target=0x7F000001, port=4000))
    assert store.get_array(A.ARR_SELF_INC)[SELF] == before + 1
    assert store.get_array(A.ARR_SEQNO)[SELF] == 0
    assert len(g) == 1 and g[0].is_direct and g[0].msg.kind == A.MSG_PING and g[0].msg.seq_no == 1
    assert g[0].msg.node == 9 and (g[0].dest_addr, g[0].dest_port) == (0x7F000001, 4000)


def test_counters_start_at_zero_and_return_new_value(store):  # Core.hs:42-53, Util.hs:79-80
    assert store.next_seqno(SELF) == 1 and store.next_seqno(SELF) == 2
    assert store.next_incarnation(SELF) == 1


# ---------------------------------------------------------------- SURVEY Appendix E
def test_e1_suspect_alive(store):
    out = store.suspect_node(SELF, msg(A.MSG_SUSPECT, ALIVE_ID, 0))
    assert out is not None and out.kind == A.MSG_SUSPECT and out.node == ALIVE_ID
    assert view(store)[ALIVE_ID] == (A.SUSPECT, 0)
    assert store.get_members(SELF)[0].timer == 5  # [Q8] the countdown is armed


@pytest.mark.parametrize("node,inc", [(SUSPECT_ID, 0), (SUSPECT_ID, 5), (DEAD_ID, 0), (17, 0)])
def test_e2_e5_suspect_ignored(store, node, inc):
    before = view(store)
    assert store.suspect_node(SELF, msg(A.MSG_SUSPECT, node, inc)) is None
    assert view(store) == before  # E3: the stored incarnation stays 0 (Q14)


def test_e6_dead_alive(store):
    out = store.dead_node(SELF, msg(A.MSG_DEAD, ALIVE_ID, 0, dead_from=33))
    assert out is not None and out.kind == A.MSG_DEAD and out.dead_from == 33  # deadFrom intact
    assert view(store)[ALIVE_ID] == (A.DEAD, 0)


def test_e7_dead_suspect(store):
    assert store.dead_node(SELF, msg(A.MSG_DEAD, SUSPECT_ID, 0, dead_from=33)) is not None
    assert view(store)[SUSPECT_ID] == (A.DEAD, 0)


def test_e8_dead_dead(store):
    assert store.dead_node(SELF, msg(A.MSG_DEAD, DEAD_ID, 0, dead_from=33)) is None


@pytest.mark.parametrize("inc,applies", [(2, False), (3, True), (4, True)])
def test_e9_e11_incarnation_compare(inc, applies):
    o = Oracle(default_config(n_nodes=N_NODES))
    o.set_members(SELF, fixture(alive_inc=3))
    out = o.suspect_node(SELF, msg(A.MSG_SUSPECT, ALIVE_ID, inc))
    assert (out is not None) == applies
    assert view(o)[ALIVE_ID] == ((A.SUSPECT, inc) if applies else (A.ALIVE, 3))


def test_e12_stale_dead_ignored():
    o = Oracle(default_config(n_nodes=N_NODES))
    o.set_members(SELF, fixture(alive_inc=3))
    assert o.dead_node(SELF, msg(A.MSG_DEAD, ALIVE_ID, 2, dead_from=1)) is None
    assert view(o)[ALIVE_ID] == (A.ALIVE, 3)


def test_e13_refute(store):
    out = store.suspect_node(SELF, msg(A.MSG_SUSPECT, SELF, 0))
    assert out is not None and out.kind == A.MSG_ALIVE and out.node == SELF and out.incarnation == 1
    assert store.get_array(A.ARR_SELF_INC)[SELF] == 1
    # a Dead about self is refuted the same way; stale accusations are ignored (Core.hs:151)
    out = store.dead_node(SELF, msg(A.MSG_DEAD, SELF, 1, dead_from=3))
    assert out.kind == A.MSG_ALIVE and out.incarnation == 2
    assert store.suspect_node(SELF, msg(A.MSG_SUSPECT, SELF, 1)) is None


def test_e15_refute_exceeds_accusation(store):  # [Q9]: max(storeIncarnation, i) + 1
    out = store.suspect_node(SELF, msg(A.MSG_SUSPECT, SELF, 7))
    assert out.incarnation == 8 and store.get_array(A.ARR_SELF_INC)[SELF] == 8


def test_e16_alive_unknown_inserted(store):  # Core.hs:206-216, then [Q7] re-broadcast
    out = store.alive_node(SELF, msg(A.MSG_ALIVE, 1 + SUSPECT_ID, 7, target=5, port=6))
    assert out is not None and out.kind == A.MSG_ALIVE and out.incarnation == 7
    ms = store.get_members(SELF)
    assert [m.id for m in ms] == [0, 1, 2, 3] and (ms[3].liveness, ms[3].incarnation) == (A.ALIVE, 7)
    # insertion keeps key order
    out = store.alive_node(SELF, msg(A.MSG_ALIVE, 50, 1))
    store.alive_node(SELF, msg(A.MSG_ALIVE, 20, 2))
    assert [m.id for m in store.get_members(SELF)] == [0, 1, 2, 3, 20, 50]


def test_q7_alive_known(store):
    assert store.alive_node(SELF, msg(A.MSG_ALIVE, SUSPECT_ID, 0)) is None        # i == j: not newer
    assert store.alive_node(SELF, msg(A.MSG_ALIVE, SUSPECT_ID, 1)) is not None    # i > j refutes Suspect
    assert view(store)[SUSPECT_ID] == (A.ALIVE, 1)
    assert store.alive_node(SELF, msg(A.MSG_ALIVE, DEAD_ID, 1)) is not None       # and overrides Dead
    assert store.alive_node(SELF, msg(A.MSG_ALIVE, SELF, 9)) is None              # our own refutation echoed


def test_wrong_constructor_is_einval(store):  # Core.hs:191,195,218 `undefined`
    with pytest.raises(OracleError):
        store.suspect_node(SELF, msg(A.MSG_DEAD, ALIVE_ID, 0))


def test_row_full_is_ecap():
    o = Oracle(default_config(n_nodes=N_NODES))
    o.set_members(SELF, [member(i, A.ALIVE) for i in range(32)])
    with pytest.raises(OracleError) as e:
        o.alive_node(SELF, msg(A.MSG_ALIVE, 33, 0))
    assert e.value.code == A.ECAP


# ---------------------------------------------------------------- bulk model sanity
def test_c1_converges_and_is_deterministic():
    from swim_b200.sim import crash_events, generate_topology
    digests = []
    for _ in range(2):
        o = Oracle(default_config(n_nodes=32, seed=0x5EED0002))
        o.set_view(generate_topology("complete", 32, 32))
        o.inject(crash_events(10, [7, 19]))
        o.step(9)
        assert o.mismatches() == 0 and o.counters()[A.CTR_PINGS] == 9 * 32
        o.step(91)
        digests.append(o.digest())
        assert o.mismatches() == 0
        st = o.get_array(A.ARR_VST).reshape(32, 32) & 3
        nbr = o.get_array(A.ARR_NBR).reshape(32, 32)
        alive_rows = [i for i in range(32) if i not in (7, 19)]
        for i in alive_rows:
            assert set(nbr[i][st[i] == A.DEAD]) == {7, 19}
    assert digests[0] == digests[1]


def test_suspicion_timer_counts_rounds():
    """[Q8] the countdown: an entry that turns Suspect during a round's tick/receive phase becomes Dead S rounds
    later; a Suspect injected by the event trace lands before that round's tick, so it expires at r + S - 1."""
    from swim_b200.sim import make_events
    S = 4
    o = Oracle(default_config(n_nodes=8, suspicion_rounds=S, retransmit=3))
    nbr = np.full((8, 32), A.NO_MEMBER, np.uint32)
    nbr[0, :2] = [1, 2]
    nbr[1, :1] = [0]
    nbr[2, :1] = [0]
    o.set_view(nbr)
    o.inject(make_events([2], [0], [A.EV_INJECT], msg_kind=[A.MSG_SUSPECT], msg_node=[2], msg_inc=[0]))
    o.step(2)
    m = {x.id: x for x in o.get_members(0)}
    assert (m[2].liveness, m[2].timer) == (A.SUSPECT, S - 1)
    o.step(S - 2)
    assert {x.id: x for x in o.get_members(0)}[2].liveness == A.SUSPECT
    o.step(1)
    m = {x.id: x for x in o.get_members(0)}
    assert (m[2].liveness, m[2].last_change) == (A.DEAD, 2 + S - 1)
    assert o.counters()[A.CTR_DEAD_TIMEOUT] == 1
    # the Dead record names the declaring node in deadFrom and is now in node 0's piggyback buffer
    pb = o.get_array(A.ARR_PB).reshape(8, -1)[0]
    assert (pb[0]["member"], pb[0]["kind"], pb[0]["from"]) == (2, A.MSG_DEAD, 0)


def test_c5_churn_suspicion_sweep_small():
    """BASELINE config C5 in miniature (10 % of the nodes down at any time, suspicion-timeout sweep): with a
    long suspicion timeout fewer entries die by their own timer (the Dead gossip or the rejoin arrives first); refutations stay zero
    without loss; rejoined nodes are re-admitted through their Alive(incarnation + 1) broadcast."""
    from swim_b200.sim import churn_events, generate_topology
    n, rounds = 256, 300
    nbr = generate_topology("ring", n, 32, 16)  # neighbours share views, so gossip carries
    ev = churn_events(n, rounds, crash_ppm=3300, rejoin_min=10, rejoin_max=50, seed=5)
    assert len(ev) > 100
    dead_declared = []
    for S in (2, 5, 13):
        o = Oracle(default_config(n_nodes=n, suspicion_rounds=S, retransmit=6, seed=9))
        o.set_view(nbr)
        o.inject(ev)
        o.step(rounds)
        c = o.counters()
        dead_declared.append(int(c[A.CTR_DEAD_TIMEOUT]))
        assert c[A.CTR_REFUTES] == 0 and c[A.CTR_SUSPECT_LOCAL] > 0
        inc = o.get_array(A.ARR_SELF_INC)
        assert inc.max() >= 1  # somebody rejoined with a bumped incarnation
        # a rejoined node's higher incarnation has reached at least one of its observers
        vinc = o.get_array(A.ARR_VINC)
        assert vinc.max() >= 1
    assert min(dead_declared) > 0 and dead_declared[2] < dead_declared[0]  # fewer timer-declared deaths with a long timeout


def test_broadcast_queue(store):
    """[Q5] disseminate's Broadcast branch: newest first, a newer record about the same member replaces the older
    one, a full buffer drops its oldest record, Ping/Ack are never enqueued."""
    B = store.cfg.pb_cap
    for i in range(B):
        store.broadcast(SELF, msg(A.MSG_SUSPECT, i, inc=i))
    got = store.get_broadcasts(SELF)
    assert [m.node for m in got] == list(range(B - 1, -1, -1))
    store.broadcast(SELF, msg(A.MSG_DEAD, 3, inc=9, dead_from=7))      # replaces Suspect(3)
    got = store.get_broadcasts(SELF)
    assert len(got) == B and (got[0].kind, got[0].node, got[0].incarnation, got[0].dead_from) == (A.MSG_DEAD, 3, 9, 7)
    assert [m.node for m in got].count(3) == 1
    store.broadcast(SELF, msg(A.MSG_ALIVE, 50, inc=1))                  # full: the oldest (member 0) falls off
    got = store.get_broadcasts(SELF)
    assert len(got) == B and got[0].node == 50 and 0 not in [m.node for m in got]
    with pytest.raises(OracleError):
        store.broadcast(SELF, msg(A.MSG_PING, 1))


def test_scalar_period_steps_tick_timers_and_take_broadcasts(store):
    """The two per-period steps of a real-time node besides the probe: the suspicion countdown (Core.hs:141 FIXME, [Q8]) and
    the piggyback payload with its transmission budget (Core.hs:136 FIXME, [Q5])."""
    # fixture: "suspect" is Suspect with 5 periods left
    for left in (4, 3, 2, 1):
        assert store.tick_timers(SELF) == 0
        assert [m.timer for m in store.get_members(SELF) if m.id == SUSPECT_ID] == [left]
    assert store.get_broadcasts(SELF) == []
    assert store.tick_timers(SELF) == 1                                   # expires: Dead, and the Dead is gossiped
    assert view(store)[SUSPECT_ID] == (A.DEAD, 0)
    got = store.get_broadcasts(SELF)
    assert [(m.kind, m.node, m.incarnation, m.dead_from) for m in got] == [(A.MSG_DEAD, SUSPECT_ID, 0, SELF)]
    assert store.tick_timers(SELF) == 0                                   # nothing left to count down
    # a fresh suspicion arms S = 5 periods
    assert store.suspect_node(SELF, msg(A.MSG_SUSPECT, ALIVE_ID, 0)) is not None
    store.broadcast(SELF, msg(A.MSG_SUSPECT, ALIVE_ID, 0))
    # default retransmit T = 8: each take returns the buffer (newest first) and spends one transmission of every record
    for _ in range(8):
        got = store.take_broadcasts(SELF)
        assert [(m.kind, m.node) for m in got] == [(A.MSG_SUSPECT, ALIVE_ID), (A.MSG_DEAD, SUSPECT_ID)]
    assert store.take_broadcasts(SELF) == [] and store.get_broadcasts(SELF) == []
    # a record enqueued later has its own budget
    store.broadcast(SELF, msg(A.MSG_ALIVE, ALIVE_ID, 3))
    assert len(store.take_broadcasts(SELF)) == 1 and len(store.get_broadcasts(SELF)) == 1
