"""Frozen per-round digests of two fixed scenarios (tests/golden/sim_vectors.json, made by the oracle with
tests/golden/make_sim_vectors.py): the oracle must keep producing them (CPU), and the CUDA path must hit the same
numbers (GPU) — the SPEC cannot drift silently on either side."""
import importlib.util
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "sim_vectors.json")))
_spec = importlib.util.spec_from_file_location("make_sim_vectors", os.path.join(HERE, "golden", "make_sim_vectors.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


def run(make, name):
    cfg, nbr, ev, rounds = gen.build(name)
    s = make(cfg)
    s.set_view(nbr)
    s.inject(ev)
    want = GOLD[name]
    for r in range(rounds):
        s.step(1)
        assert f"{s.digest():016x}" == want["digests"][r], f"{name}: digest differs at round {r + 1}"
    assert [int(x) for x in s.counters()] == want["counters"]
    assert int(s.mismatches()) == want["mismatches"]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_golden(name):
    from oracle.oracle import Oracle
    run(Oracle, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_matches_golden(name):
    from swim_b200.sim import Simulator
    run(Simulator, name)
