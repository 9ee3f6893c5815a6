#!/usr/bin/env python
"""bench.py — simulated node·rounds/sec of the SWIM per-round hot path on B200.

  python bench.py --gpus N --steps K --warmup W          # CUDA path (this repo)
  python bench.py --impl reference --steps K --warmup W  # CPU arm: the restated oracle, all host threads

A "step" is ONE protocol round over all N simulated nodes of BASELINE config C3
(N = 1,048,576, D = 32 uniform-random views, k = 3, piggyback fan-out 4, B = 8, S = 5,
0.1 % of the nodes crash at round 10). For --gpus G > 1 the node set is sharded G ways
(weak scaling: 1,048,576 nodes per GPU, config C4 at G = 4) with one all-to-all of cross-shard
piggyback envelopes per round.

Prints ONE JSON line (rank 0). `value` = node·rounds/s with state resident in HBM, timed with
CUDA events on the stream the kernels run on, max over ranks. `e2e` = the same metric through
the C ABI one round per call with host buffers: every step uploads that round's event trace,
runs one round and reads the counters and the convergence count back
(swim_sim_inject + swim_sim_step_async(1) + swim_sim_observe, one synchronisation per round).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from swim_b200 import _abi as A  # noqa: E402

def usable_cpus():
    """CPUs this process may actually use (affinity mask and cgroup quota), for the CPU arm."""
    n = len(os.sched_getaffinity(0))
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(int(q) / int(per))))
    except (OSError, ValueError):
        pass
    return n


os.environ.setdefault("OMP_NUM_THREADS", str(usable_cpus()))

N_PER_GPU = 1 << 20
CRASH_ROUND = 10
CRASH_PPM = 1000  # 0.1 %
SEED = 0x5EED0001 + 3


def workload(n_gpus, nodes_per_gpu=N_PER_GPU):
    from swim_b200.sim import crash_events, default_config, generate_topology
    n = nodes_per_gpu * n_gpus
    cfg_kw = dict(n_nodes=n, view_cap=32, k_indirect=3, fanout=4, pb_cap=8, suspicion_rounds=5, retransmit=8,
                  loss_ppm=0, seed=SEED)
    nbr = generate_topology("random", n, 32, 32, seed=3)
    rng = np.random.default_rng(3)
    crashed = np.sort(rng.choice(n, size=n * CRASH_PPM // 1000000, replace=False)).astype(np.uint32)
    return cfg_kw, nbr, crash_events(CRASH_ROUND, crashed), n


def algorithmic_bytes(cfg_kw, n_nodes, rounds, ctr_delta):
    """AB = 7·D + 16·B + 4·F + 8 + 2·(8 + 8·b̄)·m̄ bytes per node·round, canonical
    widths (state 1 B, incarnation 4 B, timer 2 B, record 8 B). Returns (AB per node·round, tick share)."""
    D, B, F = 32, cfg_kw["pb_cap"], cfg_kw["fanout"]
    msgs = float(ctr_delta[A.CTR_MSGS])
    recs = float(ctr_delta[A.CTR_RECS_SENT])
    m_bar = msgs / (n_nodes * rounds)
    b_bar = recs / msgs if msgs else 0.0
    floor = 7 * D + 16 * B + 4 * F + 8
    msg_side = (8 + 8 * b_bar) * m_bar  # written once (tick) and read once (receive)
    return floor + 2 * msg_side, floor + msg_side, m_bar, b_bar


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def cpu_arm(cfg_kw, nbr, events, n_nodes, steps, warmup):
    """The CPU arm: the restated oracle (oracle/swim_oracle.c, OpenMP over nodes) on the same config, with all the
    host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: override it before the
    default through omp_set_num_threads; SWIM_CPU_THREADS pins a number)."""
    from oracle.oracle import Oracle, num_threads, set_num_threads
    set_num_threads(int(os.environ.get("SWIM_CPU_THREADS", usable_cpus())))
    from swim_b200.sim import default_config
    orc = Oracle(default_config(**cfg_kw))
    orc.set_view(nbr)
    orc.inject(events)
    orc.step(warmup)
    t0 = time.perf_counter()
    orc.step(steps)
    dt = time.perf_counter() - t0
    return n_nodes * steps / dt, dt, num_threads(), orc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cfg_kw, nbr, events, n = workload(1)
    val, dt, threads, _ = cpu_arm(cfg_kw, nbr, events, n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "simulated node-rounds/sec", "value": val, "unit": "node-rounds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
        "config": config_dict(cfg_kw, n, 1),
        "cpu_baseline": {"value": val, "unit": "node-rounds/s", "cores": threads, "kind": "port",
                         "sample": f"all {n} nodes of C3, rounds {args.warmup + 1}..{args.warmup + args.steps} "
                                   "(restated C oracle, OpenMP; the Haskell reference cannot be built here)"},
        "e2e": {"value": val, "unit": "node-rounds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    return line


def config_dict(cfg_kw, n, n_gpus, exchange_mode="single"):
    return {"workload": f"C3 x{n_gpus}: N={n} simulated nodes ({n // n_gpus}/GPU), D=32 uniform-random views, k=3, "
                        f"fanout=4, B=8, S=5, T=8, {CRASH_PPM / 1e4:.1f}% crash at round {CRASH_ROUND}; step = 1 round",
            "n_nodes": n, "view_degree": 32, "k_indirect": 3, "fanout": 4, "pb_cap": 8, "suspicion_rounds": 5,
            "retransmit": 8, "crash_round": CRASH_ROUND, "seed": SEED,
            "parallelism": f"shard{n_gpus}" if n_gpus > 1 else "single", "exchange": exchange_mode,
            # rounds decided per batched quiet scan of round_kernel (single shard; 0 = off), DESIGN.md section 5
            "quiet_batch": (min(8, max(0, int(os.environ.get("SWIM_QUIET_BATCH", "4")))) if n_gpus == 1 else 0),
            "launch_switches": {k: os.environ[k] for k in ("SWIM_PIPELINE", "SWIM_SPLIT", "SWIM_ROUND_KERNEL",
                                                            "SWIM_ONE_ROUND_PER_LAUNCH", "SWIM_WPB", "SWIM_QUIET_BATCH") if k in os.environ},
            "l2": "no flush between rounds: consecutive rounds of one simulation share state by definition; "
                  "state arrays total 0.5 GB/GPU (> 126 MB L2), the per-round hot set (packed state rows 32 MB "
                  "+ flags) is L2-resident by design"}


def run_cuda(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the swim_b200 compute path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from swim_b200 import dist as sdist
    from swim_b200.sim import Simulator, default_config
    cfg_kw, nbr, events, n = workload(world, args.nodes_per_gpu)

    exchange = {"mode": "single"}

    def fresh(inject=True, flags=0):
        sim = Simulator(default_config(rank=rank, world=world, device=local, flags=flags, **cfg_kw))
        sim.set_view(nbr)
        exchange["mode"] = sdist.connect(sim, args.exchange)
        if inject:
            sim.inject(events)
        return sim

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------ device-resident timing (value)
    sim = fresh()
    # a non-default torch stream: its handle is what the library launches on, so the torch events
    # below bracket the kernels (handle 0 would mean "the handle's private stream" to the C ABI)
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    sim.set_stream(stream.cuda_stream)
    clocks = ClockSampler(local) if rank == 0 else None  # runs until the end of the e2e region
    barrier()
    sim.step(args.warmup)
    c0, l0 = sim.counters(), sim.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    sim.step_async(args.steps)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    c1, l1 = sim.counters(), sim.launch_count()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cd = torch.tensor((c1 - c0).astype(np.int64), device="cuda")
        dist.all_reduce(cd)
        ctr_delta = cd.cpu().numpy().astype(np.uint64)
    else:
        ctr_delta = c1 - c0
    value = n * args.steps / (ms * 1e-3)
    launches = int(l1 - l0)

    # ------------------------------------------------ per-kernel timing of the same rounds (roofline)
    sim.close()
    sim = fresh()
    sim.set_stream(stream.cuda_stream)
    sim.step(args.warmup)
    sim.set_profile(True)
    sim.step(args.steps)
    prof = sim.profile_ms()
    sim.set_profile(False)
    sim.close()
    ab_round, ab_tick, m_bar, b_bar = algorithmic_bytes(cfg_kw, n, args.steps, ctr_delta)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    tick_ms = prof["tick_scan"] / max(1.0, prof["rounds"])
    n_local = n // world
    achieved = ab_tick * n_local / (tick_ms * 1e-3) / 1e9 if tick_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "tick_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "tick_scan_kernel<1>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)",
                "traffic": traffic,
                "dram_gbs": (traffic / (tick_ms * 1e-3) / 1e9) if traffic and tick_ms > 0 else None,
                "dram_frac": (traffic / (tick_ms * 1e-3) / 1e9 / peak) if traffic and tick_ms > 0 else None,
                "algorithmic_bytes_per_launch": ab_tick * n_local,
                "ab_per_node_round": {"round": ab_round, "tick": ab_tick, "m_bar": m_bar, "b_bar": b_bar},
                "tick_ms_per_launch": tick_ms, "tick_work_ms_per_launch": prof["tick_work"] / max(1.0, prof["rounds"]),
                "recv_ms_per_launch": prof["recv"] / max(1.0, prof["rounds"]),
                "exchange_ms_per_round": prof["exchange"] / max(1.0, prof["rounds"]),
                "note": "achieved/frac use SURVEY 8(d)'s canonical bytes (376 B/node-round); K1a really moves 16 B/node-round "
                        "(one uint4 meta record: alive/suspect/crashed-member bitmaps + flags), so frac > 1 means 'faster "
                        "than streaming the canonical arrays could be', while dram_gbs/dram_frac (ncu traffic / measured "
                        "launch time) are the real HBM utilisation"}

    # ------------------------------------------------ end to end through the C ABI, host buffers
    e2e = None
    if True:
        sim = fresh(inject=False)
        by_round = {}
        for e in events:
            by_round.setdefault(int(e["round"]), []).append(e)
        h2d = d2h = 0

        def one_round(r):
            nonlocal h2d, d2h
            evs = by_round.get(r)
            if evs:
                arr = np.array(evs, dtype=A.EVENT_DTYPE)
                sim.inject(arr)  # host buffer -> library -> device (uploaded by the step below)
                h2d += arr.nbytes
            sim.step_async(1)
            # the round's result as a convergence study reads it: counters + convergence count, one read-back and ONE
            # synchronisation (the state digest is a parity tool: 370 MB of reads per call, not part of the metric)
            c, dg, mm = sim.observe(digest=False)
            d2h += c.nbytes + 8
            return c, dg, mm

        for r in range(1, args.warmup + 1):
            one_round(r)
        h2d = d2h = 0
        barrier()
        t0 = time.perf_counter()
        for r in range(args.warmup + 1, args.warmup + args.steps + 1):
            one_round(r)
        sim.sync()  # surfaces a watchdog report, if any; the stream is already idle
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n * args.steps / dt, "unit": "node-rounds/s", "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps,
               "what": "per round: swim_sim_inject(host events) + swim_sim_step_async(1) + swim_sim_observe (counters, "
                       "convergence count; one synchronisation) — host wall clock, max over ranks"}
        sim.close()
    clk = clocks.stop() if clocks else None

    # ------------------------------------------------ convergence metric (second half of BASELINE's metric)
    conv = None
    if rank == 0 or world > 1:
        def rounds_to_convergence(flags):
            sim = fresh(flags=flags)
            sim.step(CRASH_ROUND)
            r = CRASH_ROUND
            mm = None
            stride = 8  # coarse while thousands of view entries are wrong, exact (every round) in the tail
            while r < args.converge_limit:
                sim.step(stride)
                r += stride
                mm = sim.mismatches()
                if world > 1:
                    t = torch.tensor([mm], device="cuda", dtype=torch.int64)
                    dist.all_reduce(t)
                    mm = int(t.item())
                if mm == 0:
                    break
                if mm < 64:
                    stride = 1
            sim.close()
            return (r if mm == 0 else None), mm

        r0, mm0 = rounds_to_convergence(0)
        # the same workload with the paper's round-robin probe order (SWIM_F_ROUND_ROBIN, `-- FIXME: move from random to
        # robust scheme`, Core.hs:232): every observer reaches the crashed member within 2 view_cap - 1 rounds
        r1, mm1 = rounds_to_convergence(A.F_ROUND_ROBIN)
        conv = {"rounds_to_convergence": r0, "checked_every": "8 rounds, every round once fewer than 64 view entries are wrong", "limit": args.converge_limit, "mismatches_at_end": mm0,
                "crash_round": CRASH_ROUND, "rounds_to_convergence_round_robin": r1, "mismatches_at_end_round_robin": mm1}

    # ------------------------------------------------ CPU baseline (rank 0, N=1 only): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rounds = 512  # a bounded sample of roughly 10-30 s of CPU work on the box's host cores
        val, dt, threads, _ = cpu_arm(cfg_kw, nbr, events, n, rounds, 3)
        cpu = {"value": val, "unit": "node-rounds/s", "cores": threads, "kind": "port",
               "sample": f"all {n} nodes of C3, rounds 4..{3 + rounds} ({dt:.1f} s of CPU wall time; restated C "
                         "oracle with OpenMP; the Haskell reference cannot be built here: no GHC)"}

    if rank == 0:
        line = {"metric": "simulated node-rounds/sec", "value": value, "unit": "node-rounds/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32",
                "data": "synthetic", "config": config_dict(cfg_kw, n, world, exchange["mode"]), "clocks": clk, "e2e": e2e,
                "gpu_launches": launches,
                "gpu_launches_note": "round_kernel<1> runs K1a, K1b and K2 of every consecutive event-free round of a call "
                                     "in ONE launch (grid barriers between phases), so the timed region of K rounds is a "
                                     "handful of launches, not 3K",
                "roofline": roofline, "cpu_baseline": cpu, "convergence": conv,
                "counters_timed_region": dict(zip(A.CTR_NAMES, [int(x) for x in ctr_delta]))}
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line


class StdoutToStderr:
    """Everything third parties print on fd 1 during the run (e.g. NCCL's version banner) goes to stderr, so that
    stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=448)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--nodes-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--converge-limit", type=int, default=1200)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exchange", default=None, choices=[None, "p2p", "nccl"],
                    help="cross-shard exchange: fused peer-memory (default) or staged NCCL all-to-all")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    with StdoutToStderr():
        line = run_reference(args) if args.impl == "reference" else run_cuda(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
