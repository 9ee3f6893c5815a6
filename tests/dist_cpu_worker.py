"""Worker of tests/test_dist_cpu.py: one rank of a world_size-2 gloo job driving one oracle shard."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402
from swim_b200 import _abi as A  # noqa: E402
from swim_b200 import dist as sdist  # noqa: E402
from swim_b200.sim import default_config  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, _ = sdist.init_from_env("gloo")
    data = np.load(os.path.join(os.path.dirname(out_path), "case.npz"))
    cfg = default_config(n_nodes=int(data["n"]), k_indirect=3, fanout=4, pb_cap=6, suspicion_rounds=4, retransmit=5,
                         loss_ppm=int(data["loss"]), seed=int(data["seed"]), rank=rank, world=world)
    shard = Oracle(cfg)
    assert (shard.first, shard.n_local) == sdist.shard_range(cfg.n_nodes, world, rank)
    shard.set_view(data["nbr"])
    shard.inject(np.ascontiguousarray(data["events"]).reshape(-1).view(A.EVENT_DTYPE))
    sdist.run_sharded_rounds(shard, int(data["rounds"]))
    digest = sdist.global_digest(shard.digest())
    counters = sdist.global_sum(shard.counters())
    mism = int(sdist.global_sum([shard.mismatches()])[0])
    arrays = {A.ARRAY_NAMES[a]: shard.get_array(a) for a in range(A.ARR_COUNT)}
    gathered = [None] * world
    dist.all_gather_object(gathered, arrays)
    if rank == 0:
        REPLICATED = {A.ARRAY_NAMES[a] for a in A.REPLICATED_ARRAYS}  # [N] on every rank
        full = {}
        for name in arrays:
            full[name] = gathered[0][name] if name in REPLICATED else np.concatenate([g[name] for g in gathered])
        np.savez(out_path, digest=np.uint64(digest), counters=counters, mismatches=mism, **full)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
