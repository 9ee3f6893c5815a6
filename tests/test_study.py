"""swim_b200/study.py (the C5 measurements: detection latency, false positives, convergence) checked on the CPU with the
oracle standing in for the CUDA handle (same interface; the study code only reads arrays back)."""
import numpy as np

from oracle.oracle import Oracle
from swim_b200 import _abi as A
from swim_b200.sim import churn_events, crash_events, default_config, generate_topology
from swim_b200.study import event_history, latency_stats, merge_reports, run_sweep_point, view_report


def test_static_crash_latency_tracks_the_suspicion_timeout():
    n = 512
    nbr = generate_topology("random", n, 32, 32, seed=4)
    ev = crash_events(5, [3, 100, 200, 300, 400, 500])
    means = []
    for S in (2, 5, 9):
        o = Oracle(default_config(n_nodes=n, suspicion_rounds=S, seed=21))
        o.set_view(nbr)
        o.inject(ev)
        res = run_sweep_point(o, ev, 400, sample_every=20)
        rep = res["report"]
        assert res["mismatch_series"][-1] == (400, 0)                       # converged
        assert rep["undetected"] == rep["suspected"] == rep["false_dead"] == rep["stale_dead"] == rep["stale_dead_up"] == 0
        assert rep["detected"] == rep["down_entries"] > 100
        assert rep["mismatches"] == o.mismatches() == 0
        lat = rep["latency"]
        assert lat["n"] == rep["detected"] and S <= lat["p50"] and lat["max"] <= 256
        assert res["counters"]["refutes"] == 0
        means.append((lat["mean"], lat["p50"]))
    # a longer timeout detects later: the median moves by about the difference in S (the mean is dominated by the tail of
    # observers that the gossip missed and that have to probe the dead member themselves: 1/32 per round)
    assert means[0] < means[1] < means[2]
    assert 1 <= means[1][1] - means[0][1] <= 4 and 3 <= means[2][1] - means[1][1] <= 5


def test_report_agrees_with_the_library_mismatch_count_under_churn():
    n, rounds = 384, 160
    nbr = generate_topology("random", n, 32, 24, seed=8)                     # 8 vacant slots per row
    ev = churn_events(n, rounds, crash_ppm=4000, rejoin_min=10, rejoin_max=30, seed=3)
    o = Oracle(default_config(n_nodes=n, suspicion_rounds=4, loss_ppm=30000, seed=5))
    o.set_view(nbr)
    o.inject(ev)
    for upto in (40, 90, 160):
        o.step(upto - o.round)
        rep = view_report(o, ev, upto)
        assert rep["mismatches"] == o.mismatches()
        # the classes partition the entries of live observers
        assert rep["down_entries"] == rep["detected"] + rep["stale_dead"] + rep["undetected"]
        assert rep["entries"] == rep["down_entries"] + _alive_entries_of_up_members(o, n) + rep["suspected"] + \
            rep["false_dead"] + rep["stale_dead_up"]
    assert rep["detected"] > 0 and rep["suspected"] + rep["false_dead"] + rep["stale_dead_up"] > 0


def _alive_entries_of_up_members(o, n):
    alive = o.get_array(A.ARR_ALIVE).astype(bool)
    st = o.get_array(A.ARR_VST).reshape(n, 32) & 3
    nbr = o.get_array(A.ARR_NBR).reshape(n, 32)
    occ = (st != A.VACANT) & alive[:, None]
    m_up = alive[np.where(occ, nbr, 0)]
    return int((occ & m_up & (st == A.ALIVE)).sum())


def test_event_history_and_merge():
    ev = np.concatenate([crash_events(3, [1, 2]), crash_events(9, [2])])
    lc, lr = event_history(ev, 4, 5)
    assert lc.tolist() == [0, 3, 3, 0] and lr.tolist() == [0, 0, 0, 0]
    lc, _ = event_history(ev, 4, 9)
    assert lc.tolist() == [0, 3, 9, 0]
    a = {"x": 1, "latency_hist": np.array([1, 2])}
    b = {"x": 5, "latency_hist": np.array([0, 3])}
    m = merge_reports([a, b])
    assert m["x"] == 6 and m["latency_hist"].tolist() == [1, 5]
    assert latency_stats(np.array([0, 0, 4, 0, 4]))["mean"] == 3.0
    assert latency_stats(np.zeros(4, dtype=np.int64))["n"] == 0


def test_round_robin_bounds_the_detection_latency():
    """SWIM_F_ROUND_ROBIN: every observer probes the dead member within 2 cap - 1 rounds, so the worst latency is bounded by
    2 cap - 1 + S (+1: the crash lands before the tick of its round) — with random targets the tail is geometric (the test above
    sees 250+ rounds at the same size)."""
    n, cap = 512, 32
    nbr = generate_topology("random", n, cap, 32, seed=4)
    ev = crash_events(5, [3, 100, 200, 300, 400, 500])
    for S in (2, 5, 9):
        o = Oracle(default_config(n_nodes=n, suspicion_rounds=S, seed=21, flags=A.F_ROUND_ROBIN))
        o.set_view(nbr)
        o.inject(ev)
        res = run_sweep_point(o, ev, 2 * cap + S + 6, sample_every=10)
        assert res["mismatch_series"][-1][1] == 0
        assert res["report"]["latency"]["max"] <= 2 * cap - 1 + S + 1


def test_report_with_seeded_churn_needs_no_event_trace():
    """cfg.churn_ppm: crashes and rejoins are drawn inside the simulator; the report takes the transition rounds from
    SWIM_ARR_LAST_CRASH / SWIM_ARR_LAST_REJOIN."""
    n = 384
    nbr = generate_topology("ring", n, 32, 16, seed=8)
    o = Oracle(default_config(n_nodes=n, suspicion_rounds=3, seed=5, churn_ppm=5000, rejoin_min=5, rejoin_max=20))
    o.set_view(nbr)
    for upto in (30, 80, 150):
        o.step(upto - o.round)
        rep = view_report(o, None, upto)
        assert rep["mismatches"] == o.mismatches()
        assert rep["down_entries"] == rep["detected"] + rep["stale_dead"] + rep["undetected"]
    lc, lr = o.get_array(A.ARR_LAST_CRASH), o.get_array(A.ARR_LAST_REJOIN)
    assert (lc > 0).sum() > 20 and (lr > 0).sum() > 5 and rep["detected"] > 0
    res = run_sweep_point(o, None, 20, sample_every=10)
    assert res["report"]["latency"]["n"] > 0
