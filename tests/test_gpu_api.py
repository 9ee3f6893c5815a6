"""C-ABI behaviour on the device beyond parity: error codes and texts, state access, profiling hooks,
stream hand-over, event validation."""
import numpy as np
import pytest

from helpers import crash_events, default_config, generate_topology, make_events
from swim_b200 import _abi as A
from swim_b200._lib import SwimError

pytestmark = pytest.mark.gpu


def test_step_needs_a_view_and_inject_validates():
    from swim_b200.sim import Simulator
    sim = Simulator(default_config(n_nodes=64))
    with pytest.raises(SwimError) as e:
        sim.step(1)
    assert e.value.code == A.ESTATE and "no view installed" in str(e.value)
    sim.set_view(generate_topology("ring", 64, 32, 6))
    sim.step(3)
    with pytest.raises(SwimError) as e:  # an event for a round that already ran
        sim.inject(crash_events(3, [1]))
    assert e.value.code == A.EINVAL
    with pytest.raises(SwimError):       # node out of range
        sim.inject(crash_events(9, [64]))
    with pytest.raises(SwimError):       # only Suspect/Alive/Dead can be injected
        sim.inject(make_events([9], [1], [A.EV_INJECT], msg_kind=[A.MSG_PING], msg_node=[2], msg_inc=[0]))
    with pytest.raises(SwimError) as e:  # incarnation must fit the device width
        sim.inject(make_events([9], [1], [A.EV_INJECT], msg_kind=[A.MSG_ALIVE], msg_node=[2], msg_inc=[2 ** 40]))
    assert e.value.code == A.ERANGE
    bad = generate_topology("ring", 64, 32, 6)
    bad[5, 0], bad[5, 1] = bad[5, 1], bad[5, 0]  # not ascending
    with pytest.raises(SwimError):
        sim.set_view(bad)
    with pytest.raises(SwimError):
        sim.set_array(A.ARR_NBR, bad)
    with pytest.raises(SwimError):
        sim.set_array(A.ARR_VINC, np.zeros(3, np.uint32))  # wrong size


def test_state_round_trip_and_counters():
    from swim_b200.sim import Simulator
    n = 300
    cfg = default_config(n_nodes=n, seed=3)
    nbr = generate_topology("random", n, 32, 10, seed=3)
    a = Simulator(cfg)
    a.set_view(nbr)
    a.inject(crash_events(2, list(range(0, n, 9))))
    a.step(12)
    # copy the whole state into a fresh handle: it must continue identically (checkpoint/resume by hand)
    b = Simulator(cfg)
    b.set_view(nbr)
    for arr in range(A.ARR_COUNT):
        if arr != A.ARR_NBR:
            b.set_array(arr, a.get_array(arr))
    assert a.digest() == b.digest()
    c, dg, mm = a.observe()
    assert c.tolist() == a.counters().tolist() and dg == a.digest() and mm == a.mismatches()
    assert a.round == 12 and b.round == 0
    st = a.state()
    assert set(st) == set(A.ARRAY_NAMES.values()) and st["vst"].shape == (n * 32,)
    assert a.launch_count() >= 2  # round_kernel covers every event-free stretch of a call in one launch


def test_profile_hooks_and_caller_stream():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("torch CUDA streams and events: hardware only (SWIM_TEST_EMU dry run)")
    from swim_b200.sim import Simulator
    n = 20000
    sim = Simulator(default_config(n_nodes=n))
    sim.set_view(generate_topology("random", n, 32, 16, seed=1))
    stream = torch.cuda.Stream()
    sim.set_stream(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step_async(40)
    e1.record(stream)
    sim.sync()
    assert e0.elapsed_time(e1) > 0 and sim.last_step_ms() > 0
    sim.set_profile(True)
    sim.step(10)
    p = sim.profile_ms()
    assert p["rounds"] == 10 and p["tick_scan"] > 0 and p["recv"] > 0 and p["tick_work"] > 0
    sim.set_profile(False)
    sim.set_stream(0)
    sim.step(1)
    assert sim.round == 51


def test_step_observe_and_device_checkpoint():
    from swim_b200.sim import Simulator
    """swim_sim_step_observe (rounds + counters + convergence count through mapped host memory, no stream synchronisation)
    and swim_sim_save / swim_sim_load / swim_sim_set_params on hardware, against the oracle."""
    from helpers import random_events
    from oracle.oracle import Oracle
    rng = np.random.default_rng(12)
    n = 5000
    base = dict(n_nodes=n, seed=8, churn_ppm=3000, rejoin_min=3, rejoin_max=9)
    nbr = generate_topology("ring", n, 32, 16, seed=3)
    sim = Simulator(default_config(suspicion_rounds=3, **base))
    sim.set_view(nbr)
    ev = random_events(rng, n, 40, n_crash=50, n_rejoin=10, n_inject=40)
    sim.inject(ev)
    sim.save()
    for S, smax in ((3, 0), (5, 12)):
        sim.load()
        sim.set_params(suspicion_rounds=S, suspicion_max=smax)
        orc = Oracle(default_config(suspicion_rounds=S, suspicion_max=smax, **base))
        orc.set_view(nbr)
        orc.inject(ev)
        for chunk in (1, 1, 1, 7, 1, 25):
            c, mm = sim.step_observe(chunk)
            orc.step(chunk)
            assert c.tolist() == orc.counters().tolist() and mm == orc.mismatches(), (S, sim.round)
        assert sim.digest() == orc.digest()
    sim.close()
