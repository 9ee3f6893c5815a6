"""GPU parity for the protocol variants (SWIM_F_STRICT_OVERRIDE, SWIM_F_ROUND_ROBIN): CUDA == oracle, bit for bit,
every round; the rules themselves are pinned on the oracle by tests/test_variants.py. (Named to sort last.)"""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_same_state, crash_events, default_config, generate_topology, make_pair, random_events
from spec_fixture import member, msg
from swim_b200 import _abi as A

pytestmark = pytest.mark.gpu

FLAGS = [A.F_STRICT_OVERRIDE, A.F_ROUND_ROBIN, A.F_STRICT_OVERRIDE | A.F_ROUND_ROBIN]


@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("seed", range(3))
def test_variants_random_small(flags, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2, 200))
    deg = int(rng.integers(1, min(n - 1, 32) + 1))
    k = int(rng.integers(0, 8))
    cfg = default_config(n_nodes=n, view_cap=32, k_indirect=k, fanout=int(rng.integers(1, k + 2)),
                         pb_cap=int(rng.integers(1, 33)), suspicion_rounds=int(rng.integers(1, 12)),
                         retransmit=int(rng.integers(1, 12)), loss_ppm=int(rng.choice([0, 0, 50000, 300000])),
                         seed=int(rng.integers(0, 2 ** 63)), flags=flags)
    kind = rng.choice(["random", "ring"]) if deg < n - 1 else "complete"
    nbr = generate_topology(str(kind), n, 32, deg, seed=seed + 1)
    sim, orc = make_pair(cfg, nbr)
    rounds = 70  # more than two round-robin epochs
    ev = random_events(rng, n, rounds, n_crash=max(1, n // 10), n_rejoin=max(1, n // 30), n_inject=n // 4)
    sim.inject(ev)
    orc.inject(ev)
    for r in range(rounds):
        sim.step(1)
        orc.step(1)
        assert_same_state(sim, orc, f"flags {flags} seed {seed} round {r + 1}")


@pytest.mark.parametrize("cap", [64, 256])
def test_variants_wide_rows(cap):
    rng = np.random.default_rng(cap)
    n = 300
    cfg = default_config(n_nodes=n, view_cap=cap, k_indirect=5, fanout=4, pb_cap=16, suspicion_rounds=3, retransmit=5,
                         loss_ppm=20000, seed=cap, flags=A.F_STRICT_OVERRIDE | A.F_ROUND_ROBIN)
    nbr = generate_topology("random", n, cap, cap - 7, seed=3)
    sim, orc = make_pair(cfg, nbr)
    rounds = cap + 20
    ev = random_events(rng, n, rounds, n_crash=30, n_rejoin=10, n_inject=40)
    sim.inject(ev)
    orc.inject(ev)
    for r in range(rounds):
        sim.step(1)
        orc.step(1)
        assert_same_state(sim, orc, f"cap {cap} round {r + 1}")


def test_variants_multi_round_launch_c2():
    """65 536 nodes, 1 % crash, 200 rounds in ONE call (round_kernel runs the event-free stretches in single launches)."""
    n = 65536
    cfg = default_config(n_nodes=n, seed=0x5EED0001 + 2, flags=A.F_STRICT_OVERRIDE | A.F_ROUND_ROBIN)
    nbr = generate_topology("random", n, 32, 32, seed=2)
    sim, orc = make_pair(cfg, nbr)
    rng = np.random.default_rng(5)
    ev = crash_events(10, rng.choice(n, 655, replace=False))
    sim.inject(ev)
    orc.inject(ev)
    sim.step(200)
    orc.step(200)
    assert sim.digest() == orc.digest()
    assert sim.counters().tolist() == orc.counters().tolist()
    assert sim.mismatches() == orc.mismatches()


def test_strict_override_scalar_calls_match_oracle():
    from oracle.oracle import Oracle
    from swim_b200._lib import check, lib
    from swim_b200.sim import Simulator
    rng = np.random.default_rng(3)
    cfg = default_config(n_nodes=256, suspicion_rounds=7, flags=A.F_STRICT_OVERRIDE)
    sim, orc = Simulator(cfg), Oracle(cfg)
    node = 100
    ms = [member(int(i), int(rng.integers(0, 3)), int(rng.integers(0, 4))) for i in rng.choice(90, 20, replace=False)]
    for m in ms:
        m.timer = 3 if m.liveness == A.SUSPECT else 0
    check(lib().swim_set_members(sim._h, node, (A.Member * len(ms))(*ms), len(ms)), "set", sim._h)
    orc.set_members(node, ms)
    fns = {A.MSG_SUSPECT: (lib().swim_suspect_node, orc.suspect_node), A.MSG_DEAD: (lib().swim_dead_node, orc.dead_node),
           A.MSG_ALIVE: (lib().swim_alive_node, orc.alive_node)}
    known = [m.id for m in ms]
    for step in range(400):
        kind = int(rng.choice([A.MSG_SUSPECT, A.MSG_DEAD, A.MSG_ALIVE]))
        who = int(rng.choice([node, int(rng.choice(known)), int(rng.choice(known))]))
        if who == node and kind == A.MSG_ALIVE:
            who = int(rng.choice(known))
        m = msg(kind, who, int(rng.integers(0, 6)), dead_from=int(rng.integers(0, 90)))
        out, has = A.Message(), C.c_int()
        rc = fns[kind][0](sim._h, node, C.byref(m), C.byref(out), C.byref(has))
        exp = fns[kind][1](node, m)
        assert rc == 0
        assert bool(has.value) == (exp is not None), (step, kind, who)
        if exp is not None:
            assert (out.kind, out.node, out.incarnation, out.dead_from) == (exp.kind, exp.node, exp.incarnation, exp.dead_from)
        buf, cnt = (A.Member * 32)(), C.c_size_t()
        check(lib().swim_get_members(sim._h, node, buf, 32, C.byref(cnt)), "get", sim._h)
        got = [(buf[i].id, buf[i].liveness, buf[i].timer, buf[i].incarnation) for i in range(cnt.value)]
        assert got == [(x.id, x.liveness, x.timer, x.incarnation) for x in orc.get_members(node)], step
    assert sim.get_array(A.ARR_SELF_INC)[node] == orc.get_array(A.ARR_SELF_INC)[node]


def test_checkpoint_and_resume():
    """State arrays + round counter reproduce a run exactly (swim_sim_set_round): a fresh handle restored from a checkpoint
    continues with the digests of the run it was taken from and of the oracle."""
    from swim_b200.sim import Simulator
    rng = np.random.default_rng(5)
    n = 4000
    kw = dict(n_nodes=n, seed=31, loss_ppm=30000)
    nbr = generate_topology("random", n, 32, 20, seed=8)
    ev = random_events(rng, n, 60, n_crash=200, n_rejoin=60, n_inject=300)
    a, orc = make_pair(default_config(**kw), nbr)
    a.inject(ev)
    orc.inject(ev)
    a.step(25)
    orc.step(25)
    ck = a.checkpoint()
    b = Simulator(default_config(**kw))
    b.restore(ck)
    b.inject(ev[ev["round"] > 25])
    assert b.round == 25 and b.digest() == a.digest()
    for _ in range(7):
        a.step(5)
        b.step(5)
        orc.step(5)
        assert a.digest() == b.digest() == orc.digest()
