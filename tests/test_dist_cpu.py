"""N>1 path on CPU: a world_size-2 gloo job shards the node set over two oracle processes that exchange
cross-shard envelopes once per round (swim_b200.dist.run_sharded_rounds). Results must be identical to
the single-shard run: partitioning must not change a single bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import default_config, generate_topology, random_events
from oracle.oracle import Oracle
from swim_b200 import _abi as A
from swim_b200 import dist as sdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    for n, w in ((10, 3), (1 << 20, 8), (7, 8), (33, 2)):
        got = [sdist.shard_range(n, w, r) for r in range(w)]
        assert sum(c for _, c in got) == n
        for r, (f, c) in enumerate(got):
            for node in (f, f + c - 1):
                if c:
                    assert sdist.owner_of(n, w, node) == r


@pytest.mark.parametrize("world,loss", [(2, 0), (2, 30000), (3, 0)])
def test_sharded_oracle_equals_single(tmp_path, world, loss):
    n, rounds, seed = 301, 30, 77
    rng = np.random.default_rng(world * 10 + (loss > 0))
    nbr = generate_topology("random", n, 32, 20, seed=4)
    events = random_events(rng, n, rounds, n_crash=25, n_rejoin=8, n_inject=30)
    np.savez(tmp_path / "case.npz", n=n, rounds=rounds, seed=seed, loss=loss, nbr=nbr, events=np.frombuffer(events.tobytes(), dtype=np.uint8))
    out = tmp_path / "result.npz"
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
           os.path.join(ROOT, "tests", "dist_cpu_worker.py"), str(out)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    ref = Oracle(default_config(n_nodes=n, k_indirect=3, fanout=4, pb_cap=6, suspicion_rounds=4, retransmit=5,
                                loss_ppm=loss, seed=seed))
    ref.set_view(nbr)
    ref.inject(events)
    ref.step(rounds)
    assert int(got["digest"]) == ref.digest()
    assert got["counters"].tolist() == ref.counters().tolist()
    assert int(got["mismatches"]) == ref.mismatches()
    for a in range(A.ARR_COUNT):
        name = A.ARRAY_NAMES[a]
        assert np.array_equal(got[name], ref.get_array(a)), name
    assert ref.counters()[A.CTR_MSGS] > 0
