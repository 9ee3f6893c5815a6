"""Shared helpers for the parity tests: build the same scenario on the CUDA library (through
the C ABI) and on the CPU oracle, and compare every state array bit for bit."""
import numpy as np

from swim_b200 import _abi as A
from swim_b200.sim import concat_events, crash_events, default_config, generate_topology, make_events  # noqa: F401


def make_pair(cfg, nbr):
    from oracle.oracle import Oracle
    from swim_b200.sim import Simulator
    sim = Simulator(cfg)
    sim.set_view(nbr)
    orc = Oracle(cfg)
    orc.set_view(nbr)
    return sim, orc


def assert_same_state(sim, orc, where=""):
    for arr in range(A.ARR_COUNT):
        a, b = sim.get_array(arr), orc.get_array(arr)
        if not np.array_equal(a, b):
            idx = np.flatnonzero(a != b)[:8]
            raise AssertionError(f"{where}: array {A.ARRAY_NAMES[arr]} differs at {idx.tolist()}: "
                                 f"cuda={a[idx].tolist()} oracle={b[idx].tolist()}")
    ca, cb = sim.counters(), orc.counters()
    assert ca.tolist() == cb.tolist(), f"{where}: counters differ cuda={dict(zip(A.CTR_NAMES, ca.tolist()))} " \
                                       f"oracle={dict(zip(A.CTR_NAMES, cb.tolist()))}"
    assert sim.digest() == orc.digest(), f"{where}: digest differs"
    assert sim.mismatches() == orc.mismatches(), f"{where}: mismatch count differs"


def random_events(rng, n_nodes, rounds, n_crash, n_rejoin=0, n_inject=0):
    """A seeded event trace: crashes, later rejoins of some crashed nodes, injected messages."""
    evs = []
    crashed = rng.choice(n_nodes, size=min(n_crash, n_nodes), replace=False)
    cr = rng.integers(1, max(2, rounds // 2), size=len(crashed))
    evs.append(make_events(cr.astype(np.uint32), crashed.astype(np.uint32), np.full(len(crashed), A.EV_CRASH, np.uint8)))
    if n_rejoin:
        who = rng.choice(len(crashed), size=min(n_rejoin, len(crashed)), replace=False)
        rr = cr[who] + rng.integers(1, max(2, rounds // 2), size=len(who))
        evs.append(make_events(rr.astype(np.uint32), crashed[who].astype(np.uint32),
                               np.full(len(who), A.EV_REJOIN, np.uint8)))
    if n_inject:
        evs.append(make_events(rng.integers(1, rounds, size=n_inject).astype(np.uint32),
                               rng.integers(0, n_nodes, size=n_inject).astype(np.uint32),
                               np.full(n_inject, A.EV_INJECT, np.uint8),
                               msg_kind=rng.choice([A.MSG_SUSPECT, A.MSG_ALIVE, A.MSG_DEAD], size=n_inject).astype(np.uint8),
                               msg_node=rng.integers(0, n_nodes, size=n_inject).astype(np.uint32),
                               msg_inc=rng.integers(0, 4, size=n_inject).astype(np.int64),
                               msg_from=rng.integers(0, n_nodes, size=n_inject).astype(np.uint32)))
    return concat_events(evs)
