"""bench.py's CUDA arm end to end WITHOUT a GPU: the library is the emulator build (tests/emu: the CUDA sources compiled for
the CPU), `torch.cuda` is replaced by a handful of stand-ins (streams and events are wall-clock stamps). Not a measurement
— the numbers mean nothing — but every leg of the single-GPU flow runs: the one shared handle with its checkpoint, the timed
windows, the phase timeline parser, the split-kernel profile, the parity leg against the oracle, the end-to-end loop through
swim_sim_step_observe, both convergence legs (parameter change on the live handle), the ring-lattice workload and the CPU arm.
What the driver runs on hardware must at least be free of Python-level mistakes."""
import argparse
import json
import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Stream:
    cuda_stream = 0x1000  # any non-zero handle: the emulated runtime never looks inside a stream


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


@pytest.fixture()
def emulated_bench(monkeypatch):
    import torch
    import swim_b200._lib as L
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    saved = (L.SO_PATH, L._lib)
    L.SO_PATH, L._lib = build_emu.build(), None
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_a, **_k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_a, **_k: None)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    sys.path.insert(0, ROOT)
    import bench
    yield bench
    L.SO_PATH, L._lib = saved


@pytest.mark.parametrize("steps,xmode", [(20, None), (40, None), (12, "1")])
def test_cuda_arm_runs_every_leg_on_the_emulator(emulated_bench, steps, xmode, monkeypatch, capsys):
    bench = emulated_bench
    if xmode is not None:
        monkeypatch.setenv("SWIM_XMODE", xmode)
    args = argparse.Namespace(gpus=1, steps=steps, warmup=5, impl="cuda", nodes_per_gpu=8192, converge_limit=160, no_cpu=False,
                              no_parity=False, no_ring=False, windows=2, spinup=0.0, exchange=None)
    line = bench.run_cuda(args)
    json.dumps(line)  # serialisable
    assert line["metric"] == "simulated node-rounds/sec" and line["n_gpus"] == 1 and line["steps"] == steps
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
    assert line["parity_check"] == "ok", line["parity"]
    assert line["gpu_launches"] >= 1
    e = line["e2e"]
    assert e["value"] > 0 and e["api"] == "swim_sim_step_observe" and not e["notes"] and len(e["windows_ms"]) == 2
    assert e["d2h_bytes_per_step"] > 0  # counters + convergence count every round
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "latency_floor", "split_kernels_us"):
        assert k in r
    assert r["kernel"].startswith("round_kernel")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    c = line["convergence"]
    assert c["crash_round"] == bench.CRASH_ROUND and "rounds_to_convergence_round_robin" in c
    g = line["state_machine_workload"]
    assert g["parity_check"] == "ok" and g["value"] > 0
    assert line["config"]["n_nodes"] == 8192 and "workload" in line["config"]
