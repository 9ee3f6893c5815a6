"""Worker of tests/test_gpu_dist.py: one rank (one GPU) of a sharded CUDA run; rank 0 gathers the
state and writes it for comparison with the single-shard oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from swim_b200 import _abi as A  # noqa: E402
from swim_b200 import dist as sdist  # noqa: E402
from swim_b200.sim import Simulator, default_config  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, local = sdist.init_from_env("nccl")
    data = np.load(os.path.join(os.path.dirname(out_path), "case.npz"))
    cfg = default_config(n_nodes=int(data["n"]), k_indirect=3, fanout=4, pb_cap=6, suspicion_rounds=4, retransmit=5,
                         loss_ppm=int(data["loss"]), seed=int(data["seed"]), rank=rank, world=world, device=local)
    sim = Simulator(cfg)
    sim.set_view(data["nbr"])
    mode = sdist.connect(sim, str(data["mode"]))
    assert mode == str(data["mode"]), f"asked for {data['mode']} exchange, got {mode}"
    sim.inject(np.ascontiguousarray(data["events"]).reshape(-1).view(A.EVENT_DTYPE))
    chunks = [int(c) for c in data["chunks"]]
    digests = []
    for c in chunks:
        sim.step(c)
        digests.append(sdist.global_digest(sim.digest()))
    counters = sdist.global_sum(sim.counters())
    mism = int(sdist.global_sum([sim.mismatches()])[0])
    arrays = {A.ARRAY_NAMES[a]: sim.get_array(a) for a in range(A.ARR_COUNT)} if int(data["gather"]) else {}
    gathered = [None] * world
    dist.all_gather_object(gathered, arrays)
    if rank == 0:
        REPLICATED = {A.ARRAY_NAMES[a] for a in A.REPLICATED_ARRAYS}  # [N] on every rank
        full = {}
        for name in arrays:
            full[name] = gathered[0][name] if name in REPLICATED else np.concatenate([g[name] for g in gathered])
        np.savez(out_path, digests=np.array(digests, dtype=np.uint64), counters=counters, mismatches=mism, **full)
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
