"""Literal, name-keyed model of the reference's Core.hs state machine. TEST INFRASTRUCTURE ONLY.

Where oracle/swim_oracle.c restates the rules over id-indexed arrays, this file keeps the reference's own
shapes — a `Map String Member` (a dict iterated in ascending key order), `Maybe Message` results, the exact
guard order of `suspectOrDeadNode'` — so that the C oracle can be cross-checked against a second, independent
restatement (tests/test_literal_model.py). Pure-Python loops: small cases only.

Each function cites the reference lines it follows; completions of the unfinished loop carry the same
[Qn] tags as the oracle. `literal=True` switches a completion off and reproduces what
the reference does as written, where that is observable without crashing.
"""
from dataclasses import dataclass, replace
from typing import Callable, Dict, List, Optional

ALIVE, SUSPECT, DEAD = 0, 1, 2  # Types.hs:76-77


@dataclass(frozen=True)
class Member:  # Types.hs:62-68 (structural Eq over every field)
    name: str
    host: str = ""
    addr: int = 0
    port: int = 0
    alive: int = ALIVE
    incarnation: int = 0
    last_change: int = 0


@dataclass(frozen=True)
class Msg:  # Types.hs:122-145
    kind: str  # "Ping" "IndirectPing" "Ack" "Suspect" "Alive" "Dead"
    seq_no: int = 0
    node: str = ""
    target: int = 0
    port: int = 0
    incarnation: int = 0
    dead_from: str = ""
    addr: int = 0


class Store:  # Types.hs:53-60 / makeStore Util.hs:76-91
    def __init__(self, self_member: Member, suspicion_rounds: int = 5):
        self.seq_no = 0        # Util.hs:79
        self.incarnation = 0   # Util.hs:80
        self.members: Dict[str, Member] = {}  # Util.hs:78 Map.empty
        self.self = self_member
        self.now = 0           # getCurrentTime -> the round counter
        self.S = suspicion_rounds
        self.timers: Dict[str, int] = {}


def is_alive(m): return m.alive == ALIVE          # Core.hs:33-34
def is_dead(m): return m.alive == DEAD            # Core.hs:36-37
def not_alive(m): return not is_alive(m)          # Core.hs:39-40


def next_seq_no(s: Store) -> int:                 # Core.hs:42-50: returns the NEW value
    s.seq_no += 1
    return s.seq_no


def next_incarnation(s: Store) -> int:            # Core.hs:52-53
    s.incarnation += 1
    return s.incarnation


def next_incarnation_prime(s: Store, n: int, literal=False) -> int:
    """nextIncarnation' (Core.hs:56-63). Literal: loops forever when n >= inc + 1 — reported as an
    exception here. Completion [Q9]: max(storeIncarnation, n) + 1."""
    if literal:
        if n >= s.incarnation + 1:
            raise RecursionError("nextIncarnation' diverges (Core.hs:60-61)")
        s.incarnation += 1
        return s.incarnation
    s.incarnation = max(s.incarnation, n) + 1
    return s.incarnation


def members(s: Store) -> List[Member]:            # Core.hs:76-77: Map.elems = ascending key order
    return [s.members[k] for k in sorted(s.members)]


def remove_dead_nodes(s: Store) -> None:          # Core.hs:65-67
    s.members = {k: m for k, m in s.members.items() if not is_dead(m)}


def shuffle(xs: list, rand: Callable[[int, int], int]) -> list:
    """Util.hs:36-42: pick index r in [0, len-1], emit it, continue with the rest in order."""
    xs = list(xs)
    out = []
    while xs:
        r = rand(0, len(xs) - 1)
        out.append(xs[r])
        xs = xs[:r] + xs[r + 1:]  # let (l, a:r) = splitAt rand as; l <> r
    return out


def k_random_members(s: Store, n: int, excludes: List[Member], rand) -> List[Member]:
    """Core.hs:69-74: take n <$> shuffle (filter (\\m -> notElem m excludes && isAlive m) ms)."""
    cand = [m for m in members(s) if m not in excludes and is_alive(m)]
    return shuffle(cand, rand)[:n]


def suspect_or_dead_node(s: Store, msg: Msg, name: str, i: int, kind: int, literal=False) -> Optional[Msg]:
    """suspectOrDeadNode' (Core.hs:142-187), guards in the reference's order."""
    m = next((x for x in members(s) if x.name == name), None)    # Core.hs:144-145
    if m is None:                                                # 147-148: we don't know this node. ignore.
        return None
    liveness_check = (m.alive != ALIVE) if kind == SUSPECT else (m.alive == DEAD)   # 182-184
    if i < m.incarnation or liveness_check:                      # 151-152
        return None
    if name == s.self.name:                                      # 155-166: refute
        inc = next_incarnation_prime(s, i if not literal else m.incarnation, literal)
        # (the reference passes memberIncarnation m, Core.hs:157; [Q9] bumps past the accusing i)
        s.members[name] = replace(m, incarnation=inc)            # saveMember m'
        return Msg("Alive", incarnation=inc, node=name, addr=s.self.addr, port=s.self.port)
    s.members[name] = replace(m, incarnation=i, alive=kind, last_change=s.now)      # 169-177
    if kind == SUSPECT:
        s.timers[name] = s.S                                     # [Q8]
    else:
        s.timers.pop(name, None)
    return msg                                                   # 179: Just msg


def suspect_node(s, msg, literal=False):          # Core.hs:189-191
    if msg.kind != "Suspect":
        raise TypeError("undefined")
    return suspect_or_dead_node(s, msg, msg.node, msg.incarnation, SUSPECT, literal)


def dead_node(s, msg, literal=False):             # Core.hs:193-195
    if msg.kind != "Dead":
        raise TypeError("undefined")
    return suspect_or_dead_node(s, msg, msg.node, msg.incarnation, DEAD, literal)


def alive_node(s: Store, msg: Msg, literal=False) -> Optional[Msg]:
    """aliveNode (Core.hs:197-218). Literal: adds an unknown member, then `fail "READ THE PAPER"` (Q7).
    Completion [Q7]: unknown -> insert and re-broadcast; known -> applies iff incarnation is newer."""
    if msg.kind != "Alive":
        raise TypeError("undefined")
    known = next((x for x in members(s) if x.name == msg.node), None)
    if known is None:                                            # addNewMember, 206-216
        s.members[msg.node] = Member(msg.node, "", msg.addr, msg.port, ALIVE, msg.incarnation, s.now)
        if literal:
            raise IOError("READ THE PAPER")                      # Core.hs:203
        return msg
    if literal:
        raise IOError("READ THE PAPER")
    if msg.node == s.self.name:
        return None                                              # our own refutation coming back
    if msg.incarnation <= known.incarnation:
        return None
    s.members[msg.node] = replace(known, incarnation=msg.incarnation, alive=ALIVE, last_change=s.now)
    s.timers.pop(msg.node, None)
    return msg


def process(s: Store, sender, msg: Msg, literal=False) -> list:
    """process (Core.hs:89-117) -> list of ("Direct", msg, addr) / ("Broadcast", msg)."""
    if msg.kind == "Ack":
        return []                                                # 92-94
    if msg.kind == "Ping":
        return [("Direct", Msg("Ack", seq_no=msg.seq_no), sender)] if msg.node == s.self.name else []   # 97-101
    if msg.kind == "IndirectPing":                               # 105-108 (Q4 kept: seq := nextIncarnation)
        nxt = next_incarnation(s)
        return [("Direct", Msg("Ping", seq_no=nxt, node=msg.node), (msg.port, msg.target))]
    fn = {"Suspect": suspect_node, "Dead": dead_node, "Alive": alive_node}[msg.kind]
    out = fn(s, msg, literal)
    return [("Broadcast", out)] if out is not None else []      # maybeBroadcast, 119-121
