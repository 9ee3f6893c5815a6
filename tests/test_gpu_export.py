"""a round's simulated piggyback traffic exported as real datagrams in the reference's wire format,
and captured datagrams replayed into the simulator. The exported bytes are decoded with the Types.hs mirror and must
be exactly the envelopes the oracle sent in that round."""
import numpy as np
import pytest

from helpers import assert_same_state, crash_events, default_config, generate_topology, make_pair
from swim_b200 import _abi as A

pytestmark = pytest.mark.gpu

KIND = {A.MSG_SUSPECT: "Suspect", A.MSG_ALIVE: "Alive", A.MSG_DEAD: "Dead"}


def as_tuples(msgs):
    from swim_b200.types import Alive, Dead, Suspect
    out = []
    for m in msgs:
        if isinstance(m, Suspect):
            out.append((A.MSG_SUSPECT, int(m.node[1:]), m.incarnation, 0))
        elif isinstance(m, Dead):
            out.append((A.MSG_DEAD, int(m.node[1:]), m.incarnation, int(m.deadFrom[1:])))
        else:
            assert isinstance(m, Alive) and m.addr == int(m.node[1:]) and m.port == 4000
            out.append((A.MSG_ALIVE, int(m.node[1:]), m.incarnation, 0))
    return out


@pytest.mark.parametrize("kind,n,deg", [("complete", 32, 31), ("ring", 600, 12)])
def test_exported_datagrams_are_the_oracles_envelopes(kind, n, deg):
    from swim_b200.types import decode
    cfg = default_config(n_nodes=n, seed=77, loss_ppm=40000, suspicion_rounds=3)
    nbr = generate_topology(kind, n, 32, deg)
    sim, orc = make_pair(cfg, nbr)
    ev = crash_events(2, list(range(1, n, 7)))
    sim.inject(ev)
    orc.inject(ev)
    seen = 0
    for chunk in (3, 1, 4, 1, 1, 6, 1):
        sim.step(chunk)
        orc.step(chunk)
        got = sorted((s, d, tuple(as_tuples(decode(b).unEnvelope)), len(b)) for s, d, b in sim.export_round())
        want = sorted((s, d, tuple((int(r["kind"]), int(r["member"]), int(r["incarnation"]), int(r["from"]) if r["kind"] == A.MSG_DEAD else 0)
                                   for r in recs)) for s, d, recs in orc.sent())
        assert [(g[0], g[1], g[2]) for g in got] == want, f"round {sim.round}"
        seen += len(got)
    assert seen > 50
    # a compound envelope really is the reference's compound framing
    multi = [b for _, _, b in sim.export_round() if b[0] == 6]
    single = [b for _, _, b in sim.export_round() if b[0] != 6]
    assert all(b[1] >= 2 for b in multi) and all(b[0] in (3, 4, 5) for b in single)


def test_replay_of_captured_datagrams():
    """Datagrams captured from one run, delivered into another simulator through swim_sim_inject_datagram, act
    exactly like the same messages injected as events (oracle)."""
    from swim_b200.sim import make_events
    from swim_b200.types import Ack, Dead, Envelope, Ping, Suspect, encode
    n = 64
    cfg = default_config(n_nodes=n, seed=5)
    nbr = generate_topology("ring", n, 32, 8)
    sim, orc = make_pair(cfg, nbr)
    data = encode(Envelope((Suspect(0, "n5"), Ping(9, "n1"), Dead(0, "n6", "n2"), Ack(9, ()))))
    sim.inject_datagram(3, 4, data)  # node 4 knows n5 and n6 on a ring of degree 8
    orc.inject(make_events([3, 3], [4, 4], [A.EV_INJECT, A.EV_INJECT], msg_kind=[A.MSG_SUSPECT, A.MSG_DEAD], msg_node=[5, 6],
                           msg_inc=[0, 0], msg_from=[0, 2]))
    sim.step(12)
    orc.step(12)
    assert_same_state(sim, orc, "after replay")
    assert sim.counters()[A.CTR_RECS_APPLIED] >= 2
    from swim_b200._lib import SwimError
    with pytest.raises(SwimError) as e:
        sim.inject_datagram(20, 4, bytes([6, 0]))
    assert e.value.code == A.EDECODE
    with pytest.raises(SwimError):
        sim.inject_datagram(20, 4, encode(Envelope((Suspect(0, "bob"),))))  # not a simulated node name
