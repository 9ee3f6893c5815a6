"""Developer tool: the batched quiet scans of round_kernel at bench size on the emulator (592-CTA-sized grids, 2^20 nodes):
a handful of crashed nodes keep quiet and busy rounds alternating inside long launches; digest, counters and the convergence
count are compared with the oracle after every call.
    python tests/emu/big_quiet.py [N_NODES] [SWIM_QUIET_BATCH]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
if len(sys.argv) > 2:
    os.environ["SWIM_QUIET_BATCH"] = sys.argv[2]

import build_emu  # noqa: E402
import swim_b200._lib as L  # noqa: E402

L.SO_PATH, L._lib = build_emu.build(), None

from helpers import crash_events, default_config, generate_topology, make_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cfg = default_config(n_nodes=n, seed=4242, suspicion_rounds=2, retransmit=2)
nbr = generate_topology("random", n, 32, 32, seed=3)
sim, orc = make_pair(cfg, nbr)
rng = np.random.default_rng(1)
ev = crash_events(5, rng.choice(n, size=4, replace=False))
sim.inject(ev)
orc.inject(ev)
bad = 0
for chunk in (4, 1, 20, 40, 35):
    sim.step(chunk)
    orc.step(chunk)
    same = sim.digest() == orc.digest() and sim.counters().tolist() == orc.counters().tolist() and \
        sim.mismatches() == orc.mismatches()
    print(f"round {sim.round}: {'equal' if same else 'DIFFERENT'}  wrong entries {sim.mismatches()}", flush=True)
    bad += not same
sys.exit(1 if bad else 0)
