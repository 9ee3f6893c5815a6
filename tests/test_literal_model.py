"""Second, independent restatement of Core.hs (oracle/core_model.py: name-keyed dict, the reference's own guard
order and result shapes) against (a) the reference's test-suite + SURVEY Appendix E, including the cases that need
`"myself"` inside the view (E13, E14) and the literal divergences (Q7, Q9), and (b) the C oracle, message by message."""
import random

import pytest

from oracle import core_model as M
from oracle.oracle import Oracle, OracleError, philox
from spec_fixture import member, msg
from swim_b200 import _abi as A
from swim_b200.sim import default_config


def fixture_store(alive_inc=0, with_self=None):
    s = M.Store(M.Member("myself", "localhost", 4000, 123))  # makeSelf: SockAddrInet 123 4000 (Util.hs:97)
    s.members = {"alive": M.Member("alive", "127.0.0.1", 1, 4001, M.ALIVE, alive_inc),
                 "suspect": M.Member("suspect", "127.0.0.1", 1, 4002, M.SUSPECT),
                 "dead": M.Member("dead", "127.0.0.1", 1, 4003, M.DEAD)}
    if with_self is not None:
        s.members["myself"] = M.Member("myself", "localhost", 4000, 123, with_self[0], with_self[1])
    return s


def test_spec_hs_assertions_on_the_literal_model():
    s = fixture_store()
    assert [m.name for m in M.members(s)] == ["alive", "dead", "suspect"]        # Map.elems order
    rnd = random.Random(1)
    rand = lambda lo, hi: rnd.randint(lo, hi)
    assert M.k_random_members(s, 0, [], rand) == []                               # Spec.hs:111-115
    assert M.k_random_members(s, 3, [], rand) == [s.members["alive"]]             # Spec.hs:117-122
    assert M.k_random_members(s, 3, [s.members["alive"]], rand) == []             # Spec.hs:124-128
    stale = M.Member("alive", "127.0.0.1", 1, 4001, M.ALIVE, 1)
    assert M.k_random_members(s, 3, [stale], rand) == [s.members["alive"]]        # E18 structural Eq
    M.remove_dead_nodes(s)                                                        # Spec.hs:98-106
    assert "dead" not in s.members and len(s.members) == 2
    s = fixture_store()
    sender = (4000, 0x7F000001)
    assert M.process(s, sender, M.Msg("Ping", seq_no=1, node="myself")) == [("Direct", M.Msg("Ack", seq_no=1), sender)]
    assert M.process(s, sender, M.Msg("Ping", seq_no=1, node="unknown-node")) == []
    out = M.process(s, sender, M.Msg("IndirectPing", seq_no=1, target=7, port=9, node="other"))   # Spec.hs:166-174
    assert s.incarnation == 1 and s.seq_no == 0 and out == [("Direct", M.Msg("Ping", seq_no=1, node="other"), (9, 7))]
    big = M.Store(M.Member("myself"))
    big.members = {f"alive-{i}": M.Member(f"alive-{i}") for i in range(200)}      # Spec.hs:130-139
    got = M.k_random_members(big, 50, [], rand)
    assert len(got) == 50 and len({m.name for m in got}) == 50


def test_appendix_e_vectors_including_self_in_view():
    sus = lambda i, n: M.Msg("Suspect", incarnation=i, node=n)
    dead = lambda i, n, f="x": M.Msg("Dead", incarnation=i, node=n, dead_from=f)
    s = fixture_store()
    assert M.suspect_node(s, sus(0, "alive")) == sus(0, "alive") and s.members["alive"].alive == M.SUSPECT   # E1
    s = fixture_store()
    for m in (sus(0, "suspect"), sus(5, "suspect"), sus(0, "dead"), sus(0, "nobody")):                      # E2-E5
        assert M.suspect_node(s, m) is None
    assert s.members["suspect"].incarnation == 0
    assert M.dead_node(s, dead(0, "alive")) == dead(0, "alive") and s.members["alive"].alive == M.DEAD       # E6
    assert M.dead_node(s, dead(0, "suspect")) is not None and M.dead_node(s, dead(0, "dead")) is None        # E7, E8
    for inc, applies in ((2, False), (3, True), (4, True)):                                                 # E9-E11
        s = fixture_store(alive_inc=3)
        assert (M.suspect_node(s, sus(inc, "alive")) is not None) == applies
    s = fixture_store(alive_inc=3)
    assert M.dead_node(s, dead(2, "alive")) is None                                                          # E12
    # E13: "myself" in the view, Alive: refute; the Alive is built from storeSelf's swapped SockAddrInet 123 4000
    s = fixture_store(with_self=(M.ALIVE, 0))
    out = M.suspect_node(s, sus(0, "myself"))
    assert out == M.Msg("Alive", incarnation=1, node="myself", addr=4000, port=123)
    assert s.incarnation == 1 and s.members["myself"].incarnation == 1 and s.members["myself"].alive == M.ALIVE
    # E14: "myself" already Suspect in the view: the liveness guard fires BEFORE the self check -> no refutation (Q14)
    s = fixture_store(with_self=(M.SUSPECT, 0))
    assert M.suspect_node(s, sus(0, "myself")) is None and s.incarnation == 0
    # E15 / Q9: the literal nextIncarnation' diverges; the completion bumps past the accusation
    s = fixture_store(with_self=(M.ALIVE, 1))
    with pytest.raises(RecursionError):
        M.suspect_node(s, sus(1, "myself"), literal=True)
    s = fixture_store(with_self=(M.ALIVE, 1))
    assert M.suspect_node(s, sus(1, "myself")).incarnation == 2
    # E16 / Q7: the literal aliveNode inserts and then dies; the completion inserts and re-broadcasts
    s = fixture_store()
    new = M.Msg("Alive", incarnation=7, node="new", addr=5, port=6)
    with pytest.raises(IOError):
        M.alive_node(s, new, literal=True)
    assert s.members["new"].incarnation == 7
    s = fixture_store()
    assert M.alive_node(s, new) == new and [m.name for m in M.members(s)] == ["alive", "dead", "new", "suspect"]
    # a store that never inserted itself (makeStore, Util.hs:78) ignores accusations about itself: "we don't know this node"
    s = fixture_store()
    assert M.suspect_node(s, sus(0, "myself")) is None


@pytest.mark.parametrize("seed", range(5))
def test_c_oracle_equals_literal_model(seed):
    """Random message sequences through both restatements; names map to ids in ascending order. The C oracle's own
    entry is virtual (Alive, storeIncarnation), which is what a literal store holding itself as Alive does."""
    rnd = random.Random(seed)
    names = sorted(f"n{idx:02d}" for idx in rnd.sample(range(60), 18))
    ids = {n: i for i, n in enumerate(names)}
    self_id, self_name = 40, "zz-myself"
    cfg = default_config(n_nodes=64, suspicion_rounds=5)
    orc = Oracle(cfg)
    lit = M.Store(M.Member(self_name, "", self_id, 4000))
    init = []
    for n in names:
        st, inc = rnd.choice([M.ALIVE, M.ALIVE, M.SUSPECT, M.DEAD]), rnd.randint(0, 3)
        lit.members[n] = M.Member(n, "", ids[n], 4000, st, inc)
        init.append(member(ids[n], st, inc))
    lit.members[self_name] = M.Member(self_name, "", self_id, 4000, M.ALIVE, 0)
    orc.set_members(self_id, init)
    kinds = {"Suspect": A.MSG_SUSPECT, "Dead": A.MSG_DEAD, "Alive": A.MSG_ALIVE}
    for step in range(250):
        kind = rnd.choice(list(kinds))
        who = rnd.choice(names + [self_name, "n-unknown"]) if kind != "Alive" else rnd.choice(names + [self_name])
        inc = rnd.randint(0, 6)
        frm = rnd.choice(names)
        lm = M.Msg(kind, incarnation=inc, node=who, dead_from=frm if kind == "Dead" else "", addr=ids.get(who, 0), port=4000)
        cm = msg(kinds[kind], self_id if who == self_name else ids.get(who, 63), inc, dead_from=ids[frm])
        want = M.process(lit, None, lm)
        got = orc.handle_message(self_id, 0, 0, cm)
        assert len(got) == len(want), (step, lm)
        if want:
            w, g = want[0][1], got[0].msg
            assert (kinds[w.kind], w.incarnation) == (g.kind, g.incarnation), (step, lm)
            assert (self_id if w.node == self_name else ids[w.node]) == g.node
            if w.kind == "Dead":
                assert ids[w.dead_from] == g.dead_from
        view = {m.id: (m.liveness, m.incarnation) for m in orc.get_members(self_id)}
        assert view == {ids[m.name]: (m.alive, m.incarnation) for m in M.members(lit) if m.name != self_name}, step
        assert lit.incarnation == orc.get_array(A.ARR_SELF_INC)[self_id]
    # kRandomMembers: the literal shuffle fed with the oracle's Philox draws picks the same members
    for n_take in (1, 4, 18):
        key = [cfg.seed & 0xFFFFFFFF, cfg.seed >> 32]
        call = orc_calls(orc)
        draws = iter(w for blk in range(8) for w in philox([call, self_id, 2, blk], key))
        rand = lambda lo, hi: (next(draws) * (hi - lo + 1)) >> 32
        want = [ids[m.name] for m in M.k_random_members(lit, n_take, [lit.members[self_name]], rand)]
        got = [m.id for m in orc.k_random_members(self_id, n_take, [])]
        assert got == want


def orc_calls(orc):
    """kRandomMembers call number of this oracle handle (the SCALAR stream's counter word 0)."""
    c = getattr(orc, "_krm_calls", 0)
    orc._krm_calls = c + 1
    return c
