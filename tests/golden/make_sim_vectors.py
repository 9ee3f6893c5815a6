"""Generates tests/golden/sim_vectors.json: per-round state digests and final counters of two fixed scenarios,
produced by the CPU oracle (oracle/swim_oracle.c). They freeze the synchronous-round SPEC: any later
change of the oracle OR the CUDA kernels that alters a single bit of state shows up against these numbers.
Run: python tests/golden/make_sim_vectors.py   (CPU only)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import random_events  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from swim_b200.sim import crash_events, default_config, generate_topology  # noqa: E402

SCENARIOS = {
    # BASELINE config C1: N=32 complete view (D=31), k=3, B=8, S=5, nodes {7,19} crash at round 10, 100 rounds
    "c1": dict(cfg=dict(n_nodes=32, seed=0x5EED0002), topo=("complete", 32, 32, 31, 1), rounds=100, events="c1"),
    # sparse random views with loss, churn-like events and injected messages
    "mixed": dict(cfg=dict(n_nodes=200, k_indirect=5, fanout=3, pb_cap=5, suspicion_rounds=4, retransmit=6, loss_ppm=60000,
                           seed=424242), topo=("random", 200, 32, 14, 9), rounds=60, events="random"),
    # the protocol variants of DESIGN.md 2.5 on the same kind of workload: SWIM 4.2 override order, round-robin probe order
    # (more than two epochs of 32 rounds), and both together on 64-slot rows
    "mixed_strict": dict(cfg=dict(n_nodes=200, k_indirect=5, fanout=3, pb_cap=5, suspicion_rounds=4, retransmit=6, loss_ppm=60000,
                                  seed=424243, flags=1), topo=("random", 200, 32, 14, 9), rounds=60, events="random"),
    "mixed_round_robin": dict(cfg=dict(n_nodes=200, k_indirect=5, fanout=3, pb_cap=5, suspicion_rounds=4, retransmit=6,
                                       loss_ppm=60000, seed=424244, flags=2), topo=("random", 200, 32, 14, 9), rounds=80,
                              events="random"),
    "wide_both": dict(cfg=dict(n_nodes=150, view_cap=64, k_indirect=4, fanout=4, pb_cap=9, suspicion_rounds=3, retransmit=5,
                               loss_ppm=20000, seed=424245, flags=3), topo=("random", 150, 64, 50, 4), rounds=140, events="random"),
}


def build(name):
    sc = SCENARIOS[name]
    cfg = default_config(**sc["cfg"])
    kind, n, cap, deg, seed = sc["topo"]
    nbr = generate_topology(kind, n, cap, deg, seed=seed)
    if sc["events"] == "c1":
        ev = crash_events(10, [7, 19])
    else:
        ev = random_events(np.random.default_rng(2024), n, sc["rounds"], n_crash=20, n_rejoin=8, n_inject=40)
    return cfg, nbr, ev, sc["rounds"]


def main():
    out = {}
    for name in SCENARIOS:
        cfg, nbr, ev, rounds = build(name)
        o = Oracle(cfg)
        o.set_view(nbr)
        o.inject(ev)
        digests = []
        for _ in range(rounds):
            o.step(1)
            digests.append(f"{o.digest():016x}")
        out[name] = {"digests": digests, "counters": [int(x) for x in o.counters()], "mismatches": int(o.mismatches())}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim_vectors.json")
    json.dump(out, open(path, "w"), indent=1)
    print({k: (v["digests"][-1], v["counters"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
