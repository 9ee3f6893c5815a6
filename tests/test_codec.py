"""Wire codec (host code of the product library; no GPU needed): Types.hs:90-178 parity.
 - the three round trips of the reference's own test (Spec.hs:77-96)
 - byte-for-byte equality with vectors produced by Python msgpack (tests/golden/make_codec_vectors.py)
 - decode of any key order / any integer width, and the reference's error strings (Types.hs:115,118)."""
import json
import os
import struct

import msgpack
import pytest

from swim_b200 import _abi as A
from swim_b200._lib import SwimError
from swim_b200.types import (Ack, Alive, Dead, Envelope, IndirectPing, Ping, Suspect, decode, encode, msgIndex)

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "codec_vectors.json")))


def to_msg(d):
    t = d["tag"]
    if t == "Ping":
        return Ping(d["seqNo"], d["node"])
    if t == "IndirectPing":
        return IndirectPing(d["seqNo"], d["target"], d["port"], d["node"])
    if t == "Ack":
        return Ack(d["seqNo"], tuple(d["payload"]))
    if t == "Suspect":
        return Suspect(d["incarnation"], d["node"])
    if t == "Alive":
        return Alive(d["incarnation"], d["node"], d["addr"], d["port"])
    return Dead(d["incarnation"], d["node"], d["deadFrom"])


def test_spec_wire_protocol_round_trips():  # Spec.hs:77-96
    ping, iping = Ping(1, "a"), IndirectPing(2, 1, 4000, "b")
    ack, ping2, ack2 = Ack(2, ()), Ping(3, "b"), Ack(4, ())
    for msgs in ([ping], [iping], [ping, ack, ping2, ack2]):
        assert decode(encode(Envelope(tuple(msgs)))) == Envelope(tuple(msgs))
    raw = encode(Envelope((ping, ack, ping2, ack2)))
    assert raw[0] == 6 and raw[1] == 4
    lens = struct.unpack(">4H", raw[2:10])
    assert sum(lens) == len(raw) - 10


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: f"{len(c['messages'])}x{c['messages'][0]['tag']}")
def test_golden_vectors(case):
    msgs = tuple(to_msg(m) for m in case["messages"])
    assert encode(Envelope(msgs)).hex() == case["hex"]
    assert decode(bytes.fromhex(case["hex"])) == Envelope(msgs)


def test_msg_index_order():  # Types.hs:159-178
    assert [msgIndex(m) for m in (Ping(0, ""), IndirectPing(0, 0, 0, ""), Ack(0), Suspect(0, ""), Alive(0, "", 0, 0),
                                  Dead(0, "", ""))] == [0, 1, 2, 3, 4, 5]
    assert (A.MSG_COMPOUND, A.ALIVE, A.SUSPECT, A.DEAD) == (6, 0, 1, 2)


def test_decode_accepts_any_key_order_and_int_width():
    # aeson objects are HashMaps: the sender's key order is unspecified
    body = msgpack.packb({"node": "n1", "deadFrom": "n2", "incarnation": 5, "tag": "Dead"})
    assert decode(bytes([5]) + body) == Envelope((Dead(5, "n1", "n2"),))
    # a wider-than-necessary integer and an unknown key are accepted, as aeson's parser would
    body = b"\x84" + msgpack.packb("tag") + msgpack.packb("Ping") + msgpack.packb("seqNo") + b"\xcf" + (7).to_bytes(8, "big") \
        + msgpack.packb("node") + msgpack.packb("x") + msgpack.packb("extra") + msgpack.packb([1, {"a": 2}])
    assert decode(bytes([0]) + body) == Envelope((Ping(7, "x"),))


def test_single_message_type_byte_is_ignored():  # Types.hs:93-94,119 FIXME
    body = encode(Envelope((Ack(9, ()),)))[1:]
    assert decode(bytes([0]) + body) == Envelope((Ack(9, ()),))


def test_decode_errors():
    def err(data):
        with pytest.raises(SwimError) as e:
            decode(data)
        assert e.value.code == A.EDECODE
        return str(e.value)
    assert "compound message is truncated" in err(bytes([6, 2, 0]))        # SURVEY E24, Types.hs:113-115
    assert "compound mesage with zero messages" in err(bytes([6, 0]))      # SURVEY E25, Types.hs:116-118 [sic]
    assert "invalid message type" in err(bytes([9, 0x80]))                 # toEnum out of range (Types.hs:110)
    assert "Could not parse" in err(bytes([0, 0x83]))                      # truncated body
    assert "Could not parse" in err(bytes([0]) + msgpack.packb({"tag": "Ping", "seqNo": 1}))  # missing field
    assert "Could not parse" in err(bytes([0]) + msgpack.packb({"tag": "Nope"}))
    assert "Could not parse" in err(bytes([0]) + msgpack.packb({"tag": "Ping", "seqNo": 2 ** 32, "node": "a"}))
    assert "too few bytes" in err(b"")
    two = encode(Envelope((Ping(1, "a"), Ping(2, "b"))))
    assert "too few bytes" in err(two[:-1])                                # `isolate` runs out of input


def test_encode_limits():
    with pytest.raises(SwimError):
        encode(Envelope(tuple(Ping(i, "x") for i in range(256))))  # the count is one byte (Types.hs:100)
    with pytest.raises(SwimError):
        encode(Envelope((Ping(1, "x" * 256),)))
    with pytest.raises(ValueError):
        Envelope(())


def test_cross_check_with_python_msgpack_random():
    import random
    rng = random.Random(7)
    for _ in range(200):
        msgs = []
        for _ in range(rng.randint(1, 6)):
            name = "".join(rng.choice("abcxyz-0123") for _ in range(rng.randint(0, 40)))
            k = rng.randint(0, 5)
            big = rng.choice([0, 1, 127, 128, 255, 256, 65535, 65536, 2 ** 31, 2 ** 32 - 1])
            inc = rng.choice([0, 1, -1, -32, -33, -129, 2 ** 31, -2 ** 31 - 1, 2 ** 62, -2 ** 62])
            msgs.append([Ping(big, name), IndirectPing(big, big, big & 0xFFFF, name),
                         Ack(big, tuple(rng.randrange(256) for _ in range(rng.randint(0, 16)))),
                         Suspect(inc, name), Alive(inc, name, big, big & 0xFFFF), Dead(inc, name, name[::-1])][k])
        raw = encode(Envelope(tuple(msgs)))
        assert decode(raw) == Envelope(tuple(msgs))
        # independent decoder: every body is a msgpack map with the aeson-generic shape
        if len(msgs) == 1:
            bodies = [raw[1:]]
        else:
            n = raw[1]
            lens = struct.unpack(f">{n}H", raw[2:2 + 2 * n])
            pos, bodies = 2 + 2 * n, []
            for ln in lens:
                bodies.append(raw[pos:pos + ln])
                pos += ln
            assert pos == len(raw)
        for m, b in zip(msgs, bodies):
            d = msgpack.unpackb(b)
            assert d.pop("tag") == type(m).__name__
            exp = {k: (list(v) if isinstance(v, tuple) else v) for k, v in m.__dict__.items()}
            assert d == exp


def ping_with_raw_name(name: bytes) -> bytes:
    n = len(name)
    hdr = bytes([0xA0 | n]) if n < 32 else bytes([0xD9, n])
    return bytes([0]) + b"\x83" + msgpack.packb("tag") + msgpack.packb("Ping") + msgpack.packb("seqNo") + b"\x01" \
        + msgpack.packb("node") + hdr + name


def test_names_must_be_utf8():
    """A name is a Haskell `String` (Types.hs:70) carried as msgpack str: bytes that are not well-formed UTF-8 have
    no String, so the datagram fails to parse — never a name the host side cannot represent."""
    assert decode(ping_with_raw_name("nœud-é-東京-🜁".encode())) == Envelope((Ping(1, "nœud-é-東京-🜁"),))
    for bad in (b"\x80", b"a\xc3", b"\xc0\x80", b"\xe0\x80\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xf8\x88\x80\x80\x80",
                b"\xff", b"ab\xe2\x82"):
        with pytest.raises(SwimError) as e:
            decode(ping_with_raw_name(bad))
        assert e.value.code == A.EDECODE and "Could not parse" in str(e.value)


def test_utf8_validation_agrees_with_python():
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies
    pieces = st.one_of(st.binary(min_size=1, max_size=4), st.text(min_size=1, max_size=2).map(str.encode),
                       st.sampled_from([b"\xc0\xaf", b"\xed\xbf\xbf", b"\xf4\x8f\xbf\xbf", b"\xf0\x90\x80\x80", b"\xef\xbf\xbf"]))

    @hyp.settings(max_examples=600, deadline=None)
    @hyp.given(st.lists(pieces, max_size=6).map(b"".join).filter(lambda b: len(b) <= 200 and b"\0" not in b))
    def check(raw):
        try:
            want = raw.decode("utf-8")
        except UnicodeDecodeError:
            want = None
        try:
            got = decode(ping_with_raw_name(raw)).unEnvelope[0].node
        except SwimError:
            got = None
        assert got == want
    check()
