"""Host-side plumbing for multi-GPU runs: one process per GPU (torchrun), ranks own contiguous
node ranges. The per-round all-to-all of cross-shard piggyback envelopes (the UDP hop of
Core.hs:280,286) happens inside swim_sim_step (csrc/swim_dist.cu, NCCL); this module only carries
the rendezvous (NCCL unique id), the reductions of per-rank results, and a backend-agnostic
round-by-round driver used to check on CPU (gloo) that sharding does not change results."""
import os

import numpy as np


def shard_range(n_nodes: int, world: int, rank: int):
    """Contiguous shards of ceil(N/world) nodes (mirror of shard_first in csrc/swim_sim.cu)."""
    per = (n_nodes + world - 1) // world
    first = min(per * rank, n_nodes)
    return first, min(per * (rank + 1), n_nodes) - first


def owner_of(n_nodes: int, world: int, node: int) -> int:
    return node // ((n_nodes + world - 1) // world)


def init_from_env(backend=None):
    """torchrun environment -> (rank, world, local_rank); initialises torch.distributed if needed."""
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def connect(sim, mode=None):
    """Wire up the per-round exchange of a sharded simulator; call AFTER sim.set_view().
    mode "p2p" (default): fused exchange over peer memory — every rank exports CUDA IPC handles of
    its mail arrays, the blobs are all-gathered, every rank maps its peers (swim_sim_ipc_connect).
    mode "nccl": staged all-to-all — rank 0 creates the NCCL unique id, everyone joins
    (swim_sim_connect). If peer mapping fails on any rank, all ranks fall back to "nccl".
    Returns the mode in use."""
    import torch.distributed as dist
    from ._lib import SwimError
    from .sim import nccl_unique_id
    if sim.cfg.world == 1:
        return "single"
    mode = mode or os.environ.get("SWIM_EXCHANGE", "p2p")
    if mode == "p2p":
        ok = 1
        try:
            blob = sim.ipc_export()
        except SwimError:
            blob, ok = b"", 0
        blobs = [None] * dist.get_world_size()
        dist.all_gather_object(blobs, blob)
        if ok and all(blobs):
            try:
                sim.ipc_connect(blobs)
            except SwimError:
                ok = 0
        else:
            ok = 0
        if int(global_sum([ok])[0]) == dist.get_world_size():
            dist.barrier()
            return "p2p"
        mode = "nccl"
    ids = [nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.connect(ids[0])
    return "nccl"


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def global_digest(local_digest: int) -> int:
    """Shard digests add up modulo 2^64 to the single-GPU digest."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_digest
    halves = torch.tensor([local_digest & 0xFFFFFFFF, local_digest >> 32], dtype=torch.int64, device=_device())
    dist.all_reduce(halves)
    lo, hi = int(halves[0].item()), int(halves[1].item())
    return (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF


def global_sum(values) -> np.ndarray:
    import torch
    import torch.distributed as dist
    a = np.asarray(values).astype(np.int64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return a.astype(np.uint64)
    t = torch.tensor(a, device=_device())
    dist.all_reduce(t)
    return t.cpu().numpy().astype(np.uint64)


def exchange_envelopes(outboxes):
    """All-to-all of per-destination uint32 [count, words] arrays; works on any backend."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world
    dist.all_gather_object(gathered, outboxes)
    return [gathered[src][rank] for src in range(world) if src != rank]


def run_sharded_rounds(shard, rounds: int):
    """Drive one shard (anything with round_begin / outbox / inbox_add / round_end, e.g. the CPU oracle)
    through `rounds` protocol periods with one envelope all-to-all per round."""
    import torch.distributed as dist
    world = dist.get_world_size()
    for _ in range(rounds):
        shard.round_begin()
        out = [shard.outbox(r) if r != dist.get_rank() else None for r in range(world)]
        for words in exchange_envelopes(out):
            shard.inbox_add(words)
        shard.round_end()
