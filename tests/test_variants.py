"""Protocol variants (include/swim.h SWIM_F_*), pinned on the CPU oracle against an independent
statement of the rules in this file:
  * SWIM_F_STRICT_OVERRIDE — the SWIM paper's §4.2 override order instead of suspectOrDeadNode''s guards
    (Core.hs:151-152,182-184);
  * SWIM_F_ROUND_ROBIN — the `robust scheme` the reference asks for (`-- FIXME: move from random to robust scheme`,
    Core.hs:232; SWIM paper §4.3)."""
import numpy as np
import pytest

from oracle.oracle import Oracle, philox
from spec_fixture import N_NODES, SELF, member, msg, view
from swim_b200 import _abi as A
from swim_b200.sim import crash_events, default_config, generate_topology

M = 7  # the member the records are about
KINDS = {A.MSG_SUSPECT: "suspect", A.MSG_DEAD: "dead", A.MSG_ALIVE: "alive"}


def paper_rule(state, j, kind, i):
    """SWIM (Das, Gupta, Motivala 2002) §4.2, returns the new (state, incarnation) or None when the message is ignored.
    {Alive Ml, i} overrides {Suspect Ml, j}, i > j and {Alive Ml, j}, i > j (and a Dead entry, i > j: rejoin, DESIGN [Q7]);
    {Suspect Ml, i} overrides {Suspect Ml, j}, i > j and {Alive Ml, j}, i >= j;
    {Confirm Ml, i} overrides {Alive Ml, j} and {Suspect Ml, j}, any i, j."""
    if kind == A.MSG_ALIVE:
        return (A.ALIVE, i) if i > j else None
    if kind == A.MSG_SUSPECT:
        if state == A.SUSPECT:
            return (A.SUSPECT, i) if i > j else None
        if state == A.ALIVE:
            return (A.SUSPECT, i) if i >= j else None
        return None
    if state == A.DEAD:
        return None
    return (A.DEAD, max(i, j))


def reference_rule(state, j, kind, i):
    """Core.hs:151-152 with livenessCheck (182-184), and DESIGN [Q7] for Alive — what flags = 0 does."""
    if kind == A.MSG_ALIVE:
        return (A.ALIVE, i) if i > j else None
    if i < j:
        return None
    if kind == A.MSG_SUSPECT:
        return (A.SUSPECT, i) if state == A.ALIVE else None
    return (A.DEAD, i) if state != A.DEAD else None


@pytest.mark.parametrize("flags,rule", [(A.F_NONE, reference_rule), (A.F_STRICT_OVERRIDE, paper_rule)])
def test_override_table(flags, rule):
    for state in (A.ALIVE, A.SUSPECT, A.DEAD):
        for j in (0, 3):
            for kind in KINDS:
                for i in (0, 2, 3, 4):
                    o = Oracle(default_config(n_nodes=N_NODES, suspicion_rounds=6, flags=flags))
                    o.set_members(SELF, [member(M, state, j, timer=2 if state == A.SUSPECT else 0), member(9, A.ALIVE, 1)])
                    out = getattr(o, KINDS[kind] + "_node")(SELF, msg(kind, M, i, dead_from=11))
                    want = rule(state, j, kind, i)
                    where = f"flags={flags} entry=({state},{j}) msg=({KINDS[kind]},{i})"
                    got = view(o)[M]
                    if want is None:
                        assert out is None and got == (state, j), where
                    else:
                        assert out is not None and out.kind == kind and out.node == M and out.incarnation == i, where  # Core.hs:179
                        assert got == want, where
                        tm = [m.timer for m in o.get_members(SELF) if m.id == M][0]
                        assert tm == (6 if want[0] == A.SUSPECT else 0), where      # a fresh suspicion (re-)arms the countdown
                    assert view(o)[9] == (A.ALIVE, 1)


def test_strict_refutes_a_stale_confirm():
    """Others apply a Confirm whatever its incarnation, so the accused must answer even a stale one."""
    for flags, refuted in ((A.F_NONE, False), (A.F_STRICT_OVERRIDE, True)):
        o = Oracle(default_config(n_nodes=N_NODES, flags=flags))
        o.set_members(SELF, [member(M, A.ALIVE)])
        for _ in range(4):
            o.next_incarnation(SELF)
        out = o.dead_node(SELF, msg(A.MSG_DEAD, SELF, 1, dead_from=M))
        if refuted:
            assert out is not None and out.kind == A.MSG_ALIVE and out.incarnation == 5
        else:
            assert out is None
        # a stale Suspect is ignored either way
        assert o.suspect_node(SELF, msg(A.MSG_SUSPECT, SELF, 1)) is None


def test_unknown_flags_are_rejected():
    with pytest.raises(Exception):
        Oracle(default_config(n_nodes=8, flags=1 << 9))


# ---------------------------------------------------------------- round-robin targets
def rr_order(seed, cap, epoch, node):
    key = [seed & 0xFFFFFFFF, seed >> 32]
    word = philox([epoch, node >> 2, 6, 0], key)[node & 3]   # purpose 6 = P_RR
    return word & (cap - 1), (word >> 16) & (cap - 1)


def rr_target(seed, cap, rnd, node, alive_slots):
    b, r = rr_order(seed, cap, rnd // cap, node)
    p = (rnd + r) % cap
    for x in range(cap):
        slot = ((p + x) % cap) ^ b
        if slot in alive_slots:
            return slot
    return None


@pytest.mark.parametrize("cap", [32, 64])
def test_round_robin_probes_follow_the_order_and_cover_the_view(cap):
    """Every node carries a record with a long retransmission budget, so each round's ping target is visible as the
    recipient of its envelope (fan-out 1): it must be the member rr_target names, every round, for three epochs."""
    n, seed = 96, 0xABCDEF0123
    cfg = default_config(n_nodes=n, view_cap=cap, k_indirect=0, fanout=1, retransmit=255, suspicion_rounds=63, seed=seed,
                         flags=A.F_ROUND_ROBIN)
    o = Oracle(cfg)
    nbr = generate_topology("random", n, cap, cap - 3, seed=5)      # 3 vacant slots per row
    o.set_view(nbr)
    for node in range(n):
        o.broadcast(node, msg(A.MSG_ALIVE, node, 0))                 # harmless: Alive(0) about a member held at 0
    alive = [{s for s in range(cap) if nbr[node][s] != A.NO_MEMBER} for node in range(n)]
    probed = [dict() for _ in range(n)]                              # node -> slot -> rounds at which it was probed
    rounds = min(3 * cap, 250)
    for rnd in range(1, rounds + 1):
        o.round_begin()
        sent = {src: dst for src, dst, _ in o.sent()}
        o.round_end()
        assert len(sent) == n
        for node in range(n):
            want = rr_target(seed, cap, rnd, node, alive[node])
            assert sent[node] == nbr[node][want], (rnd, node)
            probed[node].setdefault(want, []).append(rnd)
    # every member of every view comes up at least once per epoch -> the gap between two probes is < 2 cap
    for node in range(n):
        assert set(probed[node]) == alive[node]
        for s, rs in probed[node].items():
            assert np.diff([0] + rs + [rounds + 1]).max() < 2 * cap


def test_round_robin_detection_is_time_bounded():
    """A crashed member is pinged by every observer within one epoch, so its failure is suspected by ALL observers after
    < 2 cap rounds by direct probing alone (dissemination switched off: B = 1, T = 1, fan-out 1, k = 0) — the SWIM paper's
    time-bounded completeness; with uniformly random targets some observer is still waiting after the same number of rounds."""
    n, cap, seed = 64, 32, 77
    nbr = generate_topology("random", n, cap, 31, seed=3)
    victim = 5
    observers = [i for i in range(n) if victim in nbr[i].tolist() and i != victim]
    assert len(observers) > 20
    left = {}
    for flags in (A.F_ROUND_ROBIN, A.F_NONE):
        o = Oracle(default_config(n_nodes=n, view_cap=cap, k_indirect=0, fanout=1, pb_cap=1, retransmit=1,
                                  suspicion_rounds=63, seed=seed, flags=flags))
        o.set_view(nbr)
        o.inject(crash_events(1, [victim]))
        o.step(2 * cap - 1)  # rounds 1 .. 2 cap - 1 contain one whole epoch
        vst = o.get_array(A.ARR_VST).reshape(n, cap)
        unaware = 0
        for i in observers:
            s = nbr[i].tolist().index(victim)
            unaware += (vst[i, s] & 3) == A.ALIVE
        left[flags] = unaware
    assert left[A.F_ROUND_ROBIN] == 0
    assert left[A.F_NONE] > 0  # uniformly random targets: (1 - 1/31)^63 = 13 % of the observers have not probed it yet


def test_round_robin_targets_match_the_independent_order():
    """k = 0, no loss, one crashed node: a node suspects the victim exactly in the round its round-robin position reaches
    the victim's slot (first Alive slot at or after p in the xor order) — checked against rr_target above."""
    n, cap, seed = 40, 32, 0x1234567
    nbr = generate_topology("random", n, cap, 20, seed=9)
    victim = 11
    o = Oracle(default_config(n_nodes=n, view_cap=cap, k_indirect=0, fanout=1, pb_cap=1, retransmit=1,
                              suspicion_rounds=63, seed=seed, flags=A.F_ROUND_ROBIN))
    o.set_view(nbr)
    o.inject(crash_events(1, [victim]))
    suspected_at = {}
    for rnd in range(1, cap + 2):
        o.step(1)
        vst = o.get_array(A.ARR_VST).reshape(n, cap)
        for i in range(n):
            if i == victim or victim not in nbr[i].tolist():
                continue
            s = nbr[i].tolist().index(victim)
            if (vst[i, s] & 3) != A.ALIVE and i not in suspected_at:
                suspected_at[i] = rnd
    for i, rnd in suspected_at.items():
        row = nbr[i].tolist()
        s = row.index(victim)
        # own probing: the first round at which rr_target lands on the victim's slot — unless gossip got there first
        alive = {x for x in range(cap) if row[x] != A.NO_MEMBER}
        first = next(r for r in range(1, cap + 2) if rr_target(seed, cap, r, i, alive) == s)
        assert rnd <= first
    own = [i for i, rnd in suspected_at.items()
           if rnd == next(r for r in range(1, cap + 2)
                          if rr_target(seed, cap, r, i, {x for x in range(cap) if nbr[i][x] != A.NO_MEMBER}) == nbr[i].tolist().index(victim))]
    assert len(own) >= len(suspected_at) // 2  # most observers found out by their own probe (one record, one hop of gossip)


@pytest.mark.parametrize("flags", [A.F_STRICT_OVERRIDE, A.F_ROUND_ROBIN, A.F_STRICT_OVERRIDE | A.F_ROUND_ROBIN])
def test_variants_converge_and_are_deterministic(flags):
    n = 256
    nbr = generate_topology("random", n, 32, 32, seed=2)
    digests = []
    for _ in range(2):
        o = Oracle(default_config(n_nodes=n, seed=99, flags=flags))
        o.set_view(nbr)
        o.inject(crash_events(5, [3, 77, 200]))
        o.step(120)
        assert o.mismatches() == 0
        digests.append(o.digest())
    assert digests[0] == digests[1]
    base = Oracle(default_config(n_nodes=n, seed=99))
    base.set_view(nbr)
    base.inject(crash_events(5, [3, 77, 200]))
    base.step(120)
    if flags & A.F_ROUND_ROBIN:
        assert base.digest() != digests[0]  # a different probing schedule
