"""swim_b200/daemon.py — the real-time single-node daemon (Core.main completed) — on loopback UDP, with the
library's scalar calls running on the SIMT emulator (tests/emu; the same test on hardware is tests/test_gpu_zz_daemon.py)."""
import socket
import time

import pytest

from swim_b200 import _abi as A


@pytest.fixture(scope="module", autouse=True)
def emu_library():
    import os
    import sys
    import swim_b200._lib as L
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "emu"))
    import build_emu
    so = build_emu.build()
    saved = (L.SO_PATH, L._lib)
    L.SO_PATH, L._lib = so, None
    assert L.lib().swim_abi_version() == A.ABI_VERSION
    yield
    L.SO_PATH, L._lib = saved


from daemon_scenarios import scenario_live_cluster_detects_a_crash, scenario_probe_escalation_and_relay  # noqa: E402


def test_probe_escalation_and_relay():
    scenario_probe_escalation_and_relay()


def test_live_cluster_detects_a_crash():
    scenario_live_cluster_detects_a_crash(period=0.25)
