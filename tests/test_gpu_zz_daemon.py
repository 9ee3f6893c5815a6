"""swim_b200/daemon.py on hardware: the same scenarios as tests/test_daemon_emu.py, the scalar calls running on the GPU."""
import pytest

from daemon_scenarios import scenario_live_cluster_detects_a_crash, scenario_probe_escalation_and_relay

pytestmark = pytest.mark.gpu


def test_probe_escalation_and_relay():
    scenario_probe_escalation_and_relay()


def test_live_cluster_detects_a_crash():
    scenario_live_cluster_detects_a_crash(period=0.1)
