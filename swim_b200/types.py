"""Host-side mirror of the reference's Types.hs (data model + wire codec), same names.

  Member / Liveness        Types.hs:62-77
  Message (6 constructors) Types.hs:122-145     msgIndex / MsgType  Types.hs:159-178
  Envelope encode/decode   Types.hs:90-119      Gossip              Types.hs:42-44
  Config                   Types.hs:46-51

Encoding and decoding run in the C library (swim_envelope_encode / swim_envelope_decode)."""
import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum
from typing import List, Tuple, Union

from . import _abi as A
from ._lib import SwimError, check, lib


class Liveness(IntEnum):  # data Liveness = IsAliveC | IsSuspectC | IsDeadC (Types.hs:76)
    IsAliveC = 0
    IsSuspectC = 1
    IsDeadC = 2


@dataclass(frozen=True)
class SockAddrInet:  # Network.Socket.SockAddrInet PortNumber HostAddress
    port: int
    host: int


@dataclass(frozen=True)
class Member:  # Types.hs:62-68; derived structural Eq over every field
    memberName: str
    memberHost: str
    memberHostNew: SockAddrInet
    memberAlive: Liveness
    memberIncarnation: int
    memberLastChange: int  # UTCTime in the reference; a round number here

    def __lt__(self, other):  # instance Ord Member: by name only (Types.hs:72-73)
        return self.memberName < other.memberName


@dataclass(frozen=True)
class Ping:
    seqNo: int
    node: str


@dataclass(frozen=True)
class IndirectPing:
    seqNo: int
    target: int
    port: int
    node: str


@dataclass(frozen=True)
class Ack:
    seqNo: int
    payload: Tuple[int, ...] = ()


@dataclass(frozen=True)
class Suspect:
    incarnation: int
    node: str


@dataclass(frozen=True)
class Alive:
    incarnation: int
    node: str
    addr: int
    port: int


@dataclass(frozen=True)
class Dead:
    incarnation: int
    node: str
    deadFrom: str


Message = Union[Ping, IndirectPing, Ack, Suspect, Alive, Dead]


class MsgType(IntEnum):  # Types.hs:159-167
    PingMsg = 0
    IndirectPingMsg = 1
    AckMsg = 2
    SuspectMsg = 3
    AliveMsg = 4
    DeadMsg = 5
    CompoundMsg = 6


_INDEX = {Ping: 0, IndirectPing: 1, Ack: 2, Suspect: 3, Alive: 4, Dead: 5}


def msgIndex(m: Message) -> int:  # Types.hs:169-178
    return _INDEX[type(m)]


@dataclass(frozen=True)
class Direct:  # Gossip = Direct Message SockAddr | Broadcast Message (Types.hs:42-44)
    msg: Message
    addr: SockAddrInet


@dataclass(frozen=True)
class Broadcast:
    msg: Message


Gossip = Union[Direct, Broadcast]


@dataclass
class Config:  # Types.hs:46-51 with the constants of parseConfig (Util.hs:44-50)
    bindHost: str = "udp://127.0.0.1:4002"
    joinHosts: List[str] = field(default_factory=lambda: ["udp://127.0.0.1:4000"])
    udpBufferSize: int = 65336
    numToGossip: int = 10
    gossipInterval: int = 200 * 1000  # `milliseconds 200` (Util.hs:23-24,49)


# ---------------------------------------------------------------- wire messages <-> C structs
def _to_wire(m: Message) -> A.WireMessage:
    w = A.WireMessage()
    w.kind = msgIndex(m)

    def name(s):
        b = s.encode("utf-8")
        if len(b) > A.NAME_MAX or b"\0" in b:
            raise SwimError(A.ERANGE, "name", "names are limited to 255 bytes without NUL")
        return b

    if isinstance(m, Ping):
        w.seq_no, w.node = m.seqNo, name(m.node)
    elif isinstance(m, IndirectPing):
        w.seq_no, w.target, w.port, w.node = m.seqNo, m.target, m.port, name(m.node)
    elif isinstance(m, Ack):
        w.seq_no = m.seqNo
        if len(m.payload) > A.ACK_PAYLOAD_MAX:
            raise SwimError(A.ERANGE, "Ack.payload", "payload longer than SWIM_ACK_PAYLOAD_MAX")
        w.payload_len = len(m.payload)
        for i, b in enumerate(m.payload):
            w.payload[i] = b
    elif isinstance(m, Suspect):
        w.incarnation, w.node = m.incarnation, name(m.node)
    elif isinstance(m, Alive):
        w.incarnation, w.node, w.target, w.port = m.incarnation, name(m.node), m.addr, m.port
    elif isinstance(m, Dead):
        w.incarnation, w.node, w.dead_from = m.incarnation, name(m.node), name(m.deadFrom)
    else:
        raise TypeError(f"not a Message: {m!r}")
    return w


def _from_wire(w: A.WireMessage) -> Message:
    node = bytes(w.node).split(b"\0", 1)[0].decode("utf-8")
    if w.kind == A.MSG_PING:
        return Ping(w.seq_no, node)
    if w.kind == A.MSG_INDIRECT_PING:
        return IndirectPing(w.seq_no, w.target, w.port, node)
    if w.kind == A.MSG_ACK:
        return Ack(w.seq_no, tuple(w.payload[i] for i in range(w.payload_len)))
    if w.kind == A.MSG_SUSPECT:
        return Suspect(w.incarnation, node)
    if w.kind == A.MSG_ALIVE:
        return Alive(w.incarnation, node, w.target, w.port)
    if w.kind == A.MSG_DEAD:
        return Dead(w.incarnation, node, bytes(w.dead_from).split(b"\0", 1)[0].decode("utf-8"))
    raise SwimError(A.EDECODE, "decode", f"unknown kind {w.kind}")


@dataclass(frozen=True)
class Envelope:  # newtype Envelope = Envelope { unEnvelope :: NonEmpty Message } (Types.hs:90)
    unEnvelope: Tuple[Message, ...]

    def __post_init__(self):
        object.__setattr__(self, "unEnvelope", tuple(self.unEnvelope))
        if not self.unEnvelope:
            raise ValueError("Envelope is a NonEmpty list")


def encode(env: Envelope) -> bytes:
    """`Data.Serialize.encode` of an Envelope (Types.hs:96-103)."""
    n = len(env.unEnvelope)
    arr = (A.WireMessage * n)(*[_to_wire(m) for m in env.unEnvelope])
    cap = 4 + n * (96 + 2 * A.NAME_MAX)
    buf = (C.c_uint8 * cap)()
    ln = C.c_size_t()
    check(lib().swim_envelope_encode(arr, n, buf, cap, C.byref(ln)), "swim_envelope_encode")
    return bytes(buf[:ln.value])


def decode(data: bytes) -> Envelope:
    """`Data.Serialize.decode :: Either String Envelope` (Types.hs:105-119); raises SwimError
    (code EDECODE, message = the reference's error text) on `Left`."""
    arr = (A.WireMessage * 255)()
    n = C.c_size_t()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if data else b"\0")
    check(lib().swim_envelope_decode(buf, len(data), arr, 255, C.byref(n)), "swim_envelope_decode")
    return Envelope(tuple(_from_wire(arr[i]) for i in range(n.value)))
