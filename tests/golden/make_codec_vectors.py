"""Generates tests/golden/codec_vectors.json with Python `msgpack` as an INDEPENDENT encoder of the
reference's wire format (Types.hs:96-119 framing around the aeson-generic msgpack body,
Types.hs:147-155). Key order: "tag" first, then the record's declaration order (Types.hs:122-145).
Run: python tests/golden/make_codec_vectors.py  (no GPU, no library needed)."""
import json
import os
import struct

import msgpack

FIELDS = {"Ping": ["seqNo", "node"], "IndirectPing": ["seqNo", "target", "port", "node"], "Ack": ["seqNo", "payload"],
          "Suspect": ["incarnation", "node"], "Alive": ["incarnation", "node", "addr", "port"],
          "Dead": ["incarnation", "node", "deadFrom"]}
INDEX = {"Ping": 0, "IndirectPing": 1, "Ack": 2, "Suspect": 3, "Alive": 4, "Dead": 5}


def body(m):
    d = {"tag": m["tag"]}
    for f in FIELDS[m["tag"]]:
        d[f] = m[f]
    return msgpack.packb(d, use_bin_type=True)


def envelope(msgs):
    if len(msgs) == 1:
        return bytes([INDEX[msgs[0]["tag"]]]) + body(msgs[0])
    bodies = [body(m) for m in msgs]
    return bytes([6, len(msgs)]) + b"".join(struct.pack(">H", len(b)) for b in bodies) + b"".join(bodies)


ping = {"tag": "Ping", "seqNo": 1, "node": "a"}
iping = {"tag": "IndirectPing", "seqNo": 2, "target": 1, "port": 4000, "node": "b"}
ack = {"tag": "Ack", "seqNo": 2, "payload": []}
ping2 = {"tag": "Ping", "seqNo": 3, "node": "b"}
ack2 = {"tag": "Ack", "seqNo": 4, "payload": []}
CASES = [
    [ping], [iping], [ping, ack, ping2, ack2],  # the three envelopes of Spec.hs:79-96
    [ack],
    [{"tag": "Ack", "seqNo": 4294967295, "payload": [0, 127, 128, 255]}],
    [{"tag": "Suspect", "incarnation": 0, "node": "alive"}],
    [{"tag": "Suspect", "incarnation": -5, "node": "n" * 40}],
    [{"tag": "Alive", "incarnation": 1, "node": "myself", "addr": 4000, "port": 123}],
    [{"tag": "Alive", "incarnation": 2 ** 40, "node": "x", "addr": 2130706433, "port": 65535}],
    [{"tag": "Dead", "incarnation": 70000, "node": "dead", "deadFrom": "x"}],
    [{"tag": "Dead", "incarnation": 300, "node": "é" * 20, "deadFrom": "z" * 200}],
    [{"tag": "Ping", "seqNo": 128, "node": ""}, {"tag": "Suspect", "incarnation": -33, "node": "q"}],
    [{"tag": "Ping", "seqNo": 65536, "node": "p%d" % i} for i in range(255)],
]

out = [{"messages": c, "hex": envelope(c).hex()} for c in CASES]
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "codec_vectors.json")
json.dump(out, open(path, "w"), indent=1)
print(len(out), "vectors ->", path)
