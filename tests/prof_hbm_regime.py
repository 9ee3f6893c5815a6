"""HBM-regime point of the per-round scan (VERDICT r1 / SURVEY 8(d)): at C3 the 16-byte-per-node records (16 MB) sit in L2,
so nothing there shows the scan against the HBM roofline. Here one GPU holds N = 2^24 nodes (256 MB of records, twice the
L2): quiescent rounds stream them from HBM every round.

    python tests/prof_hbm_regime.py [log2_nodes] [out.json]         # CUDA-event / in-kernel timing
    SWIM_SPLIT=1 ncu --set full -k regex:tick_scan -c 3 ... python tests/prof_hbm_regime.py 24   # the scan as its own kernel

Reports per quiescent round: in-kernel time of the fused kernel (scan + one grid barrier), the split scan kernel's
CUDA-event time, bytes the scan must read (16 B x N) and the fraction of MEASURED_PEAKS.json's hbm_gbs that is."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    out = sys.argv[2] if len(sys.argv) > 2 else "-"
    n = 1 << lg
    os.environ["SWIM_QUIET_BATCH"] = "0"  # every round scans
    from swim_b200.sim import Simulator, default_config, generate_topology
    t0 = time.time()
    nbr = generate_topology("random", n, 32, 32, seed=3)
    t_topo = time.time() - t0
    t0 = time.time()
    sim = Simulator(default_config(n_nodes=n, seed=11, device=0))
    sim.set_view(nbr)
    del nbr
    t_view = time.time() - t0
    peak = 6650.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk)).get("hbm_gbs", peak))
    res = {"n_nodes": n, "record_bytes": 16 * n, "setup_s": {"topology": t_topo, "create_set_view": t_view}, "peak_gbs": peak}
    sim.step(8)  # warm: clocks, derive_meta
    rounds = 64
    sim.set_timeline(rounds)
    sim.step(rounds)
    tl = sim.timeline(rounds).astype(np.int64)
    sim.set_timeline(0)
    per = (tl[:, 2] - tl[:, 0])[tl[:, 0] > 0]
    own = (tl[:, 1] - tl[:, 0])[tl[:, 0] > 0]
    if len(per):  # (no stamps on the split launch path: SWIM_SPLIT=1)
        res["fused_quiet_round_us"] = {"scan_plus_barrier_mean": float(per.mean()) / 1e3, "min": float(per.min()) / 1e3,
                                       "cta0_scan_mean": float(own.mean()) / 1e3}
    sim.step(4)
    ms = sim.last_step_ms()
    sim.step(256)
    ms = sim.last_step_ms()
    res["fused_256_rounds_ms"] = ms
    res["fused_us_per_round"] = ms * 1e3 / 256
    res["fused_dram_gbs"] = 16.0 * n / (ms * 1e-3 / 256) / 1e9
    res["fused_dram_frac"] = res["fused_dram_gbs"] / peak
    # batched quiet scans (default path): one pass decides up to 4 rounds
    sim.close()
    os.environ["SWIM_QUIET_BATCH"] = "4"
    print(json.dumps(res))
    if out != "-":
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
