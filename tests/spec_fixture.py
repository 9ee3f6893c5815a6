"""The reference's test fixture (test/Spec.hs:45-56) in id form: members "alive" < "dead" <
"suspect" (ascending key order of the Map) get ids 0, 1, 2; self = node `SELF`."""
from swim_b200 import _abi as A

ALIVE_ID, DEAD_ID, SUSPECT_ID = 0, 1, 2
SELF = 40
N_NODES = 64


def member(ident, liveness, inc=0, port=4000, last=0, timer=0):
    m = A.Member()
    m.id, m.addr, m.port, m.liveness, m.timer, m.incarnation, m.last_change = ident, ident, port, liveness, timer, inc, last
    return m


def fixture(alive_inc=0):
    # makeMembers (Spec.hs:45-56): all incarnation 0, lastChange = zeroTime
    return [member(ALIVE_ID, A.ALIVE, alive_inc), member(DEAD_ID, A.DEAD), member(SUSPECT_ID, A.SUSPECT, timer=5)]


def msg(kind, node, inc=0, dead_from=0, seq=0, target=0, port=0):
    m = A.Message()
    m.kind, m.node, m.incarnation, m.dead_from, m.seq_no, m.target, m.port = kind, node, inc, dead_from, seq, target, port
    return m


def view(store_like, node=SELF):
    return {m.id: (m.liveness, m.incarnation) for m in store_like.get_members(node)}
