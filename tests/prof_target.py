"""Short C3 run for ncu (not a test): N=1,048,576, crash at round 10, `argv[1]` rounds."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from swim_b200.sim import Simulator, crash_events, default_config, generate_topology  # noqa: E402

n = 1 << 20
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cfg = default_config(n_nodes=n, seed=0x5EED0004)
nbr = generate_topology("random", n, 32, 32, seed=3)
rng = np.random.default_rng(3)
sim = Simulator(cfg)
sim.set_view(nbr)
sim.inject(crash_events(10, np.sort(rng.choice(n, n // 1000, replace=False))))
sim.step(rounds)
print("rounds", sim.round, "digest", hex(sim.digest()))
