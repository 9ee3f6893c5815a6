"""Multi-GPU parity: the node set sharded over G GPUs (one process each, one all-to-all of cross-shard
envelopes per round inside swim_sim_step) must produce exactly the single-shard oracle's state."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import default_config, generate_topology, random_events
from swim_b200 import _abi as A

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run(tmp_path, world, n, rounds_chunks, loss, deg, gather=True, n_crash=None, mode="p2p"):
    from oracle.oracle import Oracle
    rng = np.random.default_rng(world * 100 + n)
    seed = 4242
    nbr = generate_topology("random", n, 32, deg, seed=6)
    total = sum(rounds_chunks)
    events = random_events(rng, n, total, n_crash=n_crash or max(2, n // 12), n_rejoin=max(1, n // 40), n_inject=n // 10)
    np.savez(tmp_path / "case.npz", n=n, seed=seed, loss=loss, nbr=nbr, chunks=np.array(rounds_chunks),
             gather=int(gather), mode=mode, events=np.frombuffer(events.tobytes(), dtype=np.uint8))
    out = tmp_path / "result.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-6000:]
    got = np.load(out)
    ref = Oracle(default_config(n_nodes=n, k_indirect=3, fanout=4, pb_cap=6, suspicion_rounds=4, retransmit=5,
                                loss_ppm=loss, seed=seed))
    ref.set_view(nbr)
    ref.inject(events)
    for c, dg in zip(rounds_chunks, got["digests"]):
        ref.step(c)
        assert int(dg) == ref.digest(), f"digest differs at round {ref.round}"
    assert got["counters"].tolist() == ref.counters().tolist()
    assert int(got["mismatches"]) == ref.mismatches()
    if gather:
        for a in range(A.ARR_COUNT):
            assert np.array_equal(got[A.ARRAY_NAMES[a]], ref.get_array(a)), A.ARRAY_NAMES[a]
    assert ref.counters()[A.CTR_MSGS_RECV] > 0


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cuda_equals_oracle_small(tmp_path, world, mode):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(tmp_path, world, n=1003, rounds_chunks=[1] * 12 + [20], loss=20000, deg=24, mode=mode)


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_sharded_cuda_equals_oracle_64k(tmp_path, mode):
    world = min(_gpus(), 4)
    if world < 2:
        pytest.skip("needs 2 GPUs")
    _run(tmp_path, world, n=65536, rounds_chunks=[5, 5, 10, 20], loss=0, deg=32, gather=False, n_crash=655, mode=mode)
