"""GPU parity: the CUDA path (through the C ABI) must equal the CPU oracle bit for bit on every
state array, counter and digest, after every round, on the same seed and event trace."""
import numpy as np
import pytest

from helpers import assert_same_state, crash_events, default_config, generate_topology, make_pair, random_events
from swim_b200 import _abi as A

pytestmark = pytest.mark.gpu


def test_c1_every_round():
    """BASELINE config C1: N=32 complete view (D=31), k=3, B=8, S=5, nodes {7,19} crash at round 10."""
    cfg = default_config(n_nodes=32, seed=0x5EED0001 + 1)
    nbr = generate_topology("complete", 32, 32)
    sim, orc = make_pair(cfg, nbr)
    ev = crash_events(10, [7, 19])
    sim.inject(ev)
    orc.inject(ev)
    for r in range(100):
        sim.step(1)
        orc.step(1)
        assert_same_state(sim, orc, f"round {r + 1}")
    assert sim.mismatches() == 0  # converged: both crashed nodes Dead everywhere
    c = sim.counters()
    assert c[A.CTR_DEAD_TIMEOUT] > 0 and c[A.CTR_RECS_APPLIED] > 0


@pytest.mark.parametrize("seed", range(6))
def test_random_small(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 200))
    cap = 32
    deg = int(rng.integers(1, min(n - 1, cap) + 1))
    k = int(rng.integers(0, 8))
    cfg = default_config(n_nodes=n, view_cap=cap, k_indirect=k, fanout=int(rng.integers(1, k + 2)),
                         pb_cap=int(rng.integers(1, 33)), suspicion_rounds=int(rng.integers(1, 12)),
                         retransmit=int(rng.integers(1, 12)), loss_ppm=int(rng.choice([0, 0, 50000, 300000])),
                         seed=int(rng.integers(0, 2 ** 63)))
    kind = rng.choice(["random", "ring"]) if deg < n - 1 else "complete"
    nbr = generate_topology(str(kind), n, cap, deg, seed=seed + 1)
    sim, orc = make_pair(cfg, nbr)
    rounds = 40
    ev = random_events(rng, n, rounds, n_crash=max(1, n // 10), n_rejoin=max(1, n // 30), n_inject=n // 4)
    sim.inject(ev)
    orc.inject(ev)
    for r in range(rounds):
        sim.step(1)
        orc.step(1)
        assert_same_state(sim, orc, f"seed {seed} round {r + 1}")


@pytest.mark.parametrize("cap", [64, 128, 256])
def test_wide_rows(cap):
    """view_cap > 32: a lane owns several slots of the row."""
    rng = np.random.default_rng(cap)
    n = 300
    deg = cap - 7
    cfg = default_config(n_nodes=n, view_cap=cap, k_indirect=5, fanout=4, pb_cap=16, suspicion_rounds=3,
                         retransmit=5, loss_ppm=20000, seed=cap)
    nbr = generate_topology("random", n, cap, deg, seed=3)
    sim, orc = make_pair(cfg, nbr)
    ev = random_events(rng, n, 30, n_crash=30, n_rejoin=10, n_inject=40)
    sim.inject(ev)
    orc.inject(ev)
    for r in range(30):
        sim.step(1)
        orc.step(1)
        assert_same_state(sim, orc, f"cap {cap} round {r + 1}")


def test_multi_round_steps_equal_single_steps():
    cfg = default_config(n_nodes=500, seed=99, loss_ppm=10000)
    nbr = generate_topology("random", 500, 32, 20, seed=5)
    sim, orc = make_pair(cfg, nbr)
    ev = crash_events(3, list(range(0, 500, 17)))
    sim.inject(ev)
    orc.inject(ev)
    sim.step(64)
    orc.step(64)
    assert_same_state(sim, orc, "after 64 rounds in one call")


def test_c2_digest():
    """BASELINE config C2: N=65,536, D=32 random views, k=3, 1 % crash at round 10."""
    n = 65536
    cfg = default_config(n_nodes=n, seed=0x5EED0001 + 2)
    nbr = generate_topology("random", n, 32, 32, seed=2)
    sim, orc = make_pair(cfg, nbr)
    rng = np.random.default_rng(2)
    ev = crash_events(10, rng.choice(n, size=n // 100, replace=False))
    sim.inject(ev)
    orc.inject(ev)
    for chunk in (9, 1, 5, 1, 1, 23, 60):
        sim.step(chunk)
        orc.step(chunk)
        assert sim.digest() == orc.digest(), f"digest differs at round {sim.round}"
        assert sim.counters().tolist() == orc.counters().tolist()
    assert_same_state(sim, orc, "C2 round 100")


def test_c3_properties():
    """BASELINE config C3 (N=1,048,576): size-independent properties at full size.
    determinism (two runs, same digest), counters consistent, no false positives without loss."""
    n = 1 << 20
    cfg = default_config(n_nodes=n, seed=0x5EED0001 + 3)
    nbr = generate_topology("random", n, 32, 32, seed=3)
    rng = np.random.default_rng(3)
    crashed = rng.choice(n, size=n // 1000, replace=False)
    from swim_b200.sim import Simulator
    digests = []
    for rep in range(2):
        sim = Simulator(cfg)
        sim.set_view(nbr)
        sim.inject(crash_events(10, crashed))
        sim.step(40)
        digests.append(sim.digest())
        c = sim.counters()
        st = sim.get_array(A.ARR_VST).reshape(n, 32) & 3
        alive = sim.get_array(A.ARR_ALIVE)
        sim.close()
    assert digests[0] == digests[1]
    assert c[A.CTR_PINGS] == 9 * n + 31 * (n - len(crashed))
    assert c[A.CTR_REFUTES] == 0
    # no loss => only crashed members are ever non-Alive
    nonalive_members = np.unique(nbr[st != A.ALIVE])
    assert alive[nonalive_members].max(initial=0) == 0
    assert c[A.CTR_SUSPECT_LOCAL] > 0 and c[A.CTR_DEAD_TIMEOUT] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("qbatch", [None, "0", "8"])
def test_c3_full_size_against_the_oracle(qbatch, monkeypatch):
    """BASELINE config C3 at full size (N = 1,048,576, the bench workload: 1,048 crashes at round 10), 40 rounds through
    the burst: digest, every counter and the convergence count equal the oracle's at chunk boundaries, on the default
    launch path (one fused kernel per event-free stretch, batched quiet scans) and with batching off / at 8 rounds."""
    if qbatch is not None:
        monkeypatch.setenv("SWIM_QUIET_BATCH", qbatch)
    n = 1 << 20
    cfg = default_config(n_nodes=n, seed=0x5EED0001 + 3)
    nbr = generate_topology("random", n, 32, 32, seed=3)
    rng = np.random.default_rng(3)
    crashed = np.sort(rng.choice(n, size=n // 1000, replace=False)).astype(np.uint32)
    sim, orc = make_pair(cfg, nbr)
    ev = crash_events(10, crashed)
    sim.inject(ev)
    orc.inject(ev)
    for chunk in (5, 4, 1, 2, 8, 20):
        sim.step(chunk)
        orc.step(chunk)
        assert sim.digest() == orc.digest(), f"digest differs at round {sim.round}"
        assert sim.counters().tolist() == orc.counters().tolist(), f"counters differ at round {sim.round}"
        assert sim.mismatches() == orc.mismatches()
    for arr in (A.ARR_VST, A.ARR_VINC, A.ARR_PB_CNT, A.ARR_SELF_INC):
        assert np.array_equal(sim.get_array(arr), orc.get_array(arr)), A.ARRAY_NAMES[arr]
    sim.close()


@pytest.mark.parametrize("kind,n", [("random", 4096), ("ring", 20000)])
def test_device_side_churn(kind, n):
    """BASELINE config C5 in miniature: 1 % of the processes crash per round and rejoin after U[3, 12] rounds (incarnation
    + 1, Alive broadcast), generated on the device from Philox purpose 7; every array equals the oracle's phase C."""
    cfg = default_config(n_nodes=n, seed=0x5EED0001 + 5, churn_ppm=10000, rejoin_min=3, rejoin_max=12, suspicion_rounds=3)
    nbr = generate_topology(kind, n, 32, 32 if kind == "random" else 16, seed=5)
    sim, orc = make_pair(cfg, nbr)
    for chunk in (1, 2, 5, 12, 40):
        sim.step(chunk)
        orc.step(chunk)
        assert_same_state(sim, orc, f"{kind} churn after {sim.round} rounds")
    c = sim.counters()
    assert c[A.CTR_DEAD_TIMEOUT] > 0 and c[A.CTR_RECS_APPLIED] > 0


def test_set_array_alive_rebuilds_crash_bitmaps():
    """Bulk edits of alive[] (swim_sim_set_array) must reach the per-row crashed-member bitmaps."""
    n = 400
    cfg = default_config(n_nodes=n, seed=5)
    nbr = generate_topology("random", n, 32, 16, seed=8)
    sim, orc = make_pair(cfg, nbr)
    sim.step(2)
    orc.step(2)
    alive = sim.get_array(A.ARR_ALIVE)
    alive[::7] = 0
    sim.set_array(A.ARR_ALIVE, alive)
    orc.set_array(A.ARR_ALIVE, alive)
    for r in range(12):
        sim.step(1)
        orc.step(1)
        assert_same_state(sim, orc, f"after alive edit, round {r + 3}")
    alive[::7] = 1
    sim.set_array(A.ARR_ALIVE, alive)
    orc.set_array(A.ARR_ALIVE, alive)
    sim.step(10)
    orc.step(10)
    assert_same_state(sim, orc, "after revive")


@pytest.mark.parametrize("seed", range(6, 14))
def test_random_pipelined_chunks(seed):
    """Multi-round calls run K2 of round r fused with K1a of round r+1 (receivers are skipped by the scan and
    re-scanned by K1b); single-round calls do not. Both must equal the oracle at every chunk boundary."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 3000))
    cap = int(rng.choice([32, 32, 64]))
    deg = int(rng.integers(2, min(n - 1, cap) + 1))
    k = int(rng.integers(0, 8))
    cfg = default_config(n_nodes=n, view_cap=cap, k_indirect=k, fanout=int(rng.integers(1, k + 2)),
                         pb_cap=int(rng.integers(1, 17)), suspicion_rounds=int(rng.integers(1, 9)),
                         retransmit=int(rng.integers(1, 9)), loss_ppm=int(rng.choice([0, 30000, 200000])),
                         seed=int(rng.integers(0, 2 ** 63)))
    kind = str(rng.choice(["random", "ring"])) if deg < n - 1 else "complete"
    nbr = generate_topology(kind, n, cap, deg, seed=seed + 1)
    sim, orc = make_pair(cfg, nbr)
    chunks = [3, 7, 1, 13, 2, 16, 5]
    rounds = sum(chunks)
    ev = random_events(rng, n, rounds, n_crash=max(1, n // 10), n_rejoin=max(1, n // 30), n_inject=n // 4)
    sim.inject(ev)
    orc.inject(ev)
    for c in chunks:
        sim.step(c)
        orc.step(c)
        assert_same_state(sim, orc, f"seed {seed} round {sim.round}")


def test_dense_gossip_pipelined():
    """Complete views: almost every receiver applies news and becomes a sender in the next round, so K1b's
    re-scan path (receivers turned work items) carries most of the traffic."""
    n = 33
    cfg = default_config(n_nodes=n, view_cap=32, k_indirect=3, fanout=4, pb_cap=8, suspicion_rounds=3, retransmit=6,
                         loss_ppm=150000, seed=12345)
    nbr = generate_topology("complete", n, 32)
    sim, orc = make_pair(cfg, nbr)
    ev = crash_events(2, [1, 5, 9, 20])
    sim.inject(ev)
    orc.inject(ev)
    for c in (40, 1, 25, 30):
        sim.step(c)
        orc.step(c)
        assert_same_state(sim, orc, f"dense round {sim.round}")
    assert sim.counters()[A.CTR_RECS_APPLIED] > 100 and sim.counters()[A.CTR_REFUTES] > 0


def test_c5_churn_parity():
    """BASELINE config C5 at test size: continuous churn (crash / rejoin with incarnation + 1), ring views so
    gossip carries, a few suspicion timeouts; every array equal to the oracle at every chunk boundary."""
    from swim_b200.sim import churn_events
    n = 4096
    nbr = generate_topology("ring", n, 32, 16)
    ev = churn_events(n, 120, crash_ppm=3000, rejoin_min=10, rejoin_max=40, seed=3)
    for S in (2, 8):
        cfg = default_config(n_nodes=n, suspicion_rounds=S, retransmit=6, loss_ppm=5000, seed=21 + S)
        sim, orc = make_pair(cfg, nbr)
        sim.inject(ev)
        orc.inject(ev)
        for c in (1, 30, 9, 80):
            sim.step(c)
            orc.step(c)
            assert_same_state(sim, orc, f"churn S={S} round {sim.round}")
        assert sim.counters()[A.CTR_RECS_APPLIED] > 0


@pytest.mark.parametrize("qbatch", ["0", "3", "8"])
def test_batched_quiet_scans(qbatch, monkeypatch):
    """round_kernel's batched quiet scans (SWIM_QUIET_BATCH rounds per pass after a round without work):
    a few crashed nodes, each in 32 views, make quiet and busy rounds alternate inside long launches for a while (a batch
    ends wherever some observer's draw hits a crashed member), then the cluster converges and whole batches commit.
    State, digest and counters — every Ping counted exactly once — equal the oracle's; 0 turns batching off."""
    monkeypatch.setenv("SWIM_QUIET_BATCH", qbatch)
    n = 6000
    cfg = default_config(n_nodes=n, seed=777 + int(qbatch), suspicion_rounds=2, retransmit=2)
    nbr = generate_topology("random", n, 32, 32, seed=13)
    sim, orc = make_pair(cfg, nbr)
    ev = crash_events(4, [101, 4000])
    sim.inject(ev)
    orc.inject(ev)
    for chunk in (1, 2, 3, 61, 9, 40, 2, 130, 64, 11):
        sim.step(chunk)
        orc.step(chunk)
        assert_same_state(sim, orc, f"qbatch {qbatch} after {sim.round} rounds")


@pytest.mark.parametrize("xmode", ["1", "0"])
def test_one_barrier_round_kernel_equals_oracle(xmode, monkeypatch):
    """round_kernel_x (one grid barrier per round; the default for launches of >= 32 rounds on one shard) and the two-phase
    round_kernel, each forced for every launch, against the oracle: ring-lattice views (mail every round) and sparse random
    views through long event-free stretches, every array compared at the chunk boundaries."""
    import numpy as np
    from helpers import assert_same_state, crash_events, default_config, generate_topology, make_pair
    monkeypatch.setenv("SWIM_XMODE", xmode)
    for n, topo, deg, loss, chunks in [(20000, "ring", 24, 0, [2, 6, 60]), (50000, "random", 32, 0, [2, 1, 130]),
                                       (6000, "ring", 16, 30000, [2, 40])]:
        cfg = default_config(n_nodes=n, k_indirect=3, fanout=4, pb_cap=8, suspicion_rounds=5, retransmit=8, seed=4242 + n,
                             loss_ppm=loss, device=0)
        nbr = generate_topology(topo, n, 32, deg, seed=5)
        sim, orc = make_pair(cfg, nbr)
        rng = np.random.default_rng(n)
        ev = crash_events(3, np.sort(rng.choice(n, size=n // 50, replace=False)).astype(np.uint32))
        sim.inject(ev)
        orc.inject(ev)
        for c in chunks:
            sim.step(c)
            orc.step(c)
            assert_same_state(sim, orc, f"xmode {xmode} n {n} {topo} after {c} more rounds")
        sim.close()
