"""Developer tool: fuzz the device code on the emulator against the oracle — random sizes, view widths, fan-outs, buffer
capacities, loss, crashes / rejoins / injected messages, protocol variants, single- and multi-round launches.
    python tests/emu/soak.py FIRST_SEED N_SEEDS        (prints one line per failing seed; exit code 1 if any)"""
import os
import sys
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import build_emu  # noqa: E402
import swim_b200._lib as L  # noqa: E402

L.SO_PATH, L._lib = build_emu.build(), None

from helpers import assert_same_state, default_config, generate_topology, make_pair, random_events  # noqa: E402


def one(seed):
    rng = np.random.default_rng(seed)
    cap = int(rng.choice([32, 32, 32, 64, 128, 256]))
    n = int(rng.integers(2, 260))
    deg = int(rng.integers(1, min(n - 1, cap) + 1))
    k = int(rng.integers(0, 8))
    flags = int(rng.integers(0, 4))
    cfg = default_config(n_nodes=n, view_cap=cap, k_indirect=k, fanout=int(rng.integers(1, k + 2)), pb_cap=int(rng.integers(1, 33)),
                         suspicion_rounds=int(rng.integers(1, 14)), retransmit=int(rng.integers(1, 14)),
                         loss_ppm=int(rng.choice([0, 0, 20000, 200000, 600000])), seed=int(rng.integers(0, 2 ** 63)), flags=flags)
    kind = str(rng.choice(["random", "ring"])) if deg < n - 1 else "complete"
    nbr = generate_topology(kind, n, cap, deg, seed=int(rng.integers(1, 1000)))
    sim, orc = make_pair(cfg, nbr)
    rounds = int(rng.integers(20, 90))
    ev = random_events(rng, n, rounds, n_crash=max(1, n // int(rng.integers(3, 20))), n_rejoin=max(1, n // 25), n_inject=n // 3)
    sim.inject(ev)
    orc.inject(ev)
    done = 0
    while done < rounds:
        chunk = int(rng.choice([1, 1, 1, 2, 5, 17]))
        chunk = min(chunk, rounds - done)
        sim.step(chunk)
        orc.step(chunk)
        done += chunk
        assert_same_state(sim, orc, f"seed {seed} round {done} (n={n} cap={cap} k={k} flags={flags})")
    sim.close()


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + count):
        try:
            one(seed)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print(f"FAIL seed {seed}: {e}", flush=True)
            traceback.print_exc(limit=2)
    print(f"soak {first}..{first + count - 1}: {count - bad} ok, {bad} failed", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
