"""Developer tool: fuzz the device code on the emulator against the oracle — random sizes, view widths, fan-outs, buffer
capacities, loss, crashes / rejoins / injected messages, protocol variants, single- and multi-round launches.
    python tests/emu/soak.py FIRST_SEED N_SEEDS [sharded]   (prints one line per failing seed; exit code 1 if any)"""
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
os.environ["SWIM_NCCL_LIB"] = os.path.join(HERE, "libfake_nccl.so")

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
    if rng.random() < 0.4:   # seeded device-side churn
        cfg.churn_ppm, cfg.rejoin_min, cfg.rejoin_max = int(rng.choice([3000, 30000, 150000])), 1 + int(rng.integers(0, 4)), 5 + int(rng.integers(0, 12))
    if rng.random() < 0.35:
        cfg.probes_per_round = int(rng.integers(2, 5))
    if rng.random() < 0.4 and cfg.suspicion_rounds <= 15:   # Lifeguard-style dynamic suspicion timeout
        cfg.suspicion_max = int(rng.integers(cfg.suspicion_rounds, 16))
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


def one_sharded(seed):
    """Several ranks in one process (fused exchange), stepped by one host thread each, against the single-shard oracle."""
    import threading
    from oracle.oracle import Oracle
    from swim_b200 import _abi as A
    from swim_b200.sim import Simulator
    rng = np.random.default_rng(seed)
    world = int(rng.integers(2, 5))
    n = int(rng.integers(world * 2, 400))
    deg = int(rng.integers(1, min(n - 1, 32) + 1))
    k = int(rng.integers(0, 8))
    kw = dict(n_nodes=n, k_indirect=k, fanout=int(rng.integers(1, k + 2)), pb_cap=int(rng.integers(1, 17)),
              suspicion_rounds=int(rng.integers(1, 9)), retransmit=int(rng.integers(1, 9)),
              loss_ppm=int(rng.choice([0, 20000, 200000])), seed=int(rng.integers(0, 2 ** 63)), flags=int(rng.integers(0, 4)))
    os.environ["SWIM_ROUND_KERNEL"] = "1" if rng.random() < 0.6 else "0"
    if rng.random() < 0.4:
        kw.update(churn_ppm=int(rng.choice([3000, 30000])), rejoin_min=2, rejoin_max=9)
    if rng.random() < 0.4:
        kw.update(suspicion_max=int(rng.integers(kw["suspicion_rounds"], 16)))
    if rng.random() < 0.3:
        kw.update(probes_per_round=int(rng.integers(2, 5)))
    nbr = generate_topology("random" if deg < n - 1 else "complete", n, 32, deg, seed=int(rng.integers(1, 1000)))
    rounds = int(rng.integers(10, 50))
    ev = random_events(rng, n, rounds, n_crash=max(1, n // 8), n_rejoin=max(1, n // 30), n_inject=n // 5)
    sims = [Simulator(default_config(rank=r, world=world, **kw)) for r in range(world)]
    for s_ in sims:
        s_.set_view(nbr)
    staged = rng.random() < 0.3  # the NCCL baseline path, NCCL replaced by tests/emu/fake_nccl.cpp
    if staged:
        from swim_b200.sim import nccl_unique_id
        uid = nccl_unique_id()
        ts = [threading.Thread(target=s_.connect, args=(uid,)) for s_ in sims]
        [t.start() for t in ts]
        [t.join() for t in ts]
    else:
        blobs = [s_.ipc_export() for s_ in sims]
        for s_ in sims:
            s_.ipc_connect(blobs)
    for s_ in sims:
        s_.inject(ev)
    ref = Oracle(default_config(**kw))
    ref.set_view(nbr)
    ref.inject(ev)
    done = 0
    while done < rounds:
        chunk = min(int(rng.choice([1, 1, 3, 11])), rounds - done)
        errs = []

        def work(x):
            try:
                x.step(chunk)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=work, args=(x,)) for x in sims]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
        ref.step(chunk)
        done += chunk
        where = f"seed {seed} round {done} (world={world} n={n} flags={kw['flags']} rk={os.environ.get('SWIM_ROUND_KERNEL')} staged={staged})"
        assert sum(x.digest() for x in sims) % 2 ** 64 == ref.digest(), where
        assert sum(x.mismatches() for x in sims) == ref.mismatches(), where
    assert np.sum([x.counters() for x in sims], axis=0).tolist() == ref.counters().tolist()
    for a in range(A.ARR_COUNT):
        got = sims[0].get_array(a) if a in A.REPLICATED_ARRAYS else np.concatenate([x.get_array(a) for x in sims])
        assert np.array_equal(got, ref.get_array(a)), A.ARRAY_NAMES[a]
    for x in sims:
        x.close()


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    sharded = len(sys.argv) > 3 and sys.argv[3] == "sharded"
    bad = 0
    for seed in range(first, first + count):
        try:
            (one_sharded if sharded else one)(seed)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print(f"FAIL seed {seed}: {e}", flush=True)
            traceback.print_exc(limit=2)
    print(f"soak {first}..{first + count - 1}: {count - bad} ok, {bad} failed", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
