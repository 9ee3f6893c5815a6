"""Memory-safety checks of the host-side C/C++ code on CPU: the oracle under AddressSanitizer +
UBSan on a scenario that exercises every phase, and the wire decoder fuzzed with hypothesis (it parses bytes
from the network: it must reject garbage, never crash)."""
import os
import subprocess
import sys

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from swim_b200 import _abi as A
from swim_b200._lib import SwimError
from swim_b200.types import Ack, Alive, Dead, Envelope, IndirectPing, Ping, Suspect, decode, encode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCENARIO = r'''
import ctypes, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import oracle.oracle as O
O._SO = os.path.join(sys.argv[1], "oracle", "liboracle_asan.so")
from oracle.oracle import Oracle
from swim_b200 import _abi as A
from swim_b200.sim import default_config, generate_topology, make_events, concat_events, crash_events, churn_events
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from spec_fixture import member, msg
n = 257
for world in (1, 3):
    shards = [Oracle(default_config(n_nodes=n, k_indirect=7, fanout=8, pb_cap=3, suspicion_rounds=2, retransmit=2,
                                    loss_ppm=200000, seed=9, rank=r, world=world)) for r in range(world)]
    nbr = generate_topology("random", n, 32, 31, seed=2)
    ev = concat_events([churn_events(n, 40, 20000, 2, 6, seed=1),
                        make_events([3, 3, 9], [5, 5, 200], [A.EV_INJECT] * 3, msg_kind=[A.MSG_SUSPECT, A.MSG_DEAD, A.MSG_ALIVE],
                                    msg_node=[5, 9, 77], msg_inc=[4, 0, 9], msg_from=[1, 2, 3])])
    for s in shards:
        s.set_view(nbr); s.inject(ev)
    for r in range(40):
        for s in shards: s.round_begin()
        for a in shards:
            for b in shards:
                if a is not b: b.inbox_add(a.outbox(b.cfg.rank))
        for s in shards: s.round_end()
    print(world, hex(sum(s.digest() for s in shards) & (2**64 - 1)), sum(s.mismatches() for s in shards))
    s = shards[0]
    s.sent(); s.get_array(A.ARR_PB)
    for i in range(40):   # scalar API incl. row-full and insertion paths
        try: s.alive_node(0, msg(A.MSG_ALIVE, 100 + i, i))
        except Exception: pass
    s.remove_dead_nodes(0); s.k_random_members(0, 300, []); s.broadcast(0, msg(A.MSG_DEAD, 3, 1, dead_from=2)); s.get_broadcasts(0)
print("ok")
'''


def test_oracle_under_asan_ubsan(tmp_path):
    so = os.path.join(ROOT, "oracle", "liboracle_asan.so")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_asan.so"], capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(so):
        pytest.skip("no sanitizer runtime for this compiler: " + r.stderr[-300:])
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan):
        pytest.skip("libasan.so not found")
    script = tmp_path / "scenario.py"
    script.write_text(SCENARIO)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1",
               OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-1500:], r.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].split()[1:] == lines[1].split()[1:]  # 1 shard and 3 shards: same digest, same mismatch count


@settings(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.binary(min_size=0, max_size=200))
def test_decoder_rejects_garbage_without_crashing(data):
    try:
        env = decode(data)
    except SwimError as e:
        assert e.code == A.EDECODE
    else:
        assert 1 <= len(env.unEnvelope) <= 255
        assert decode(encode(env)) == env  # whatever it accepted re-encodes to an equal envelope


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.data())
def test_mutated_valid_datagrams(data):
    name = st.text(alphabet="abcxyz0189-", min_size=0, max_size=30)
    u32, u16 = st.integers(0, 2 ** 32 - 1), st.integers(0, 65535)
    inc = st.integers(-2 ** 63, 2 ** 63 - 1)
    one = st.one_of(st.builds(Ping, u32, name), st.builds(IndirectPing, u32, u32, u16, name),
                    st.builds(Ack, u32, st.lists(st.integers(0, 255), max_size=16).map(tuple)),
                    st.builds(Suspect, inc, name), st.builds(Alive, inc, name, u32, u16), st.builds(Dead, inc, name, name))
    env = Envelope(tuple(data.draw(st.lists(one, min_size=1, max_size=6))))
    raw = bytearray(encode(env))
    assert decode(bytes(raw)) == env
    for _ in range(data.draw(st.integers(1, 4))):  # flip / truncate
        if raw and data.draw(st.booleans()):
            raw[data.draw(st.integers(0, len(raw) - 1))] = data.draw(st.integers(0, 255))
        else:
            raw = raw[:data.draw(st.integers(0, len(raw)))]
    try:
        decode(bytes(raw))
    except SwimError as e:
        assert e.code == A.EDECODE


CODEC_FUZZ = r'''
import ctypes as C, os, random, sys
sys.path.insert(0, sys.argv[1])
from swim_b200 import _abi as A
L = C.CDLL(sys.argv[2])
L.swim_envelope_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
L.swim_envelope_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
rnd = random.Random(12345)
msgs = (A.WireMessage * 255)()
out = (A.WireMessage * 255)()
buf = (C.c_uint8 * 70000)()
n_ok = n_bad = 0
for it in range(6000):
    k = rnd.choice([1, 1, 2, 3, 7, 255])
    for i in range(k):
        m = msgs[i]
        m.kind = rnd.randrange(6); m.seq_no = rnd.getrandbits(32); m.target = rnd.getrandbits(32); m.port = rnd.getrandbits(16)
        m.incarnation = rnd.choice([0, 1, -1, 127, 128, -33, 2**31, -2**31 - 1, 2**62, -2**62])
        m.payload_len = rnd.randrange(17)
        m.node = bytes(rnd.choice(b"abcxyz09-") for _ in range(rnd.choice([0, 1, 5, 31, 32, 200, 255])))
        m.dead_from = bytes(rnd.choice(b"abcxyz09-") for _ in range(rnd.choice([0, 3, 40, 255])))
    ln = C.c_size_t()
    rc = L.swim_envelope_encode(msgs, k, buf, rnd.choice([70000, 70000, 40, 3, 0]), C.byref(ln))
    if rc != 0:
        continue
    raw = bytearray(bytes(buf[:ln.value]))
    for _ in range(rnd.randrange(0, 4)):
        if raw and rnd.random() < 0.6:
            raw[rnd.randrange(len(raw))] = rnd.randrange(256)
        else:
            raw = raw[:rnd.randrange(len(raw) + 1)]
    data = (C.c_uint8 * max(1, len(raw))).from_buffer_copy(bytes(raw) if raw else b"\0")
    cnt = C.c_size_t()
    rc = L.swim_envelope_decode(data, len(raw), out, rnd.choice([255, 255, 1, 0]), C.byref(cnt))
    n_ok += rc == 0
    n_bad += rc != 0
for it in range(20000):  # pure garbage
    raw = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 64)))
    data = (C.c_uint8 * max(1, len(raw))).from_buffer_copy(raw if raw else b"\0")
    cnt = C.c_size_t()
    L.swim_envelope_decode(data, len(raw), out, 255, C.byref(cnt))
print("ok", n_ok, n_bad)
'''


def test_codec_under_asan_ubsan(tmp_path):
    """The wire codec (product host code) compiled with AddressSanitizer + UBSan and driven with valid, mutated and
    random datagrams: no out-of-bounds access, no undefined behaviour."""
    stub = tmp_path / "stub.cpp"
    stub.write_text('#include <string>\nnamespace swim { thread_local std::string g_last_error; }\n'
                    'extern "C" __attribute__((visibility("default"))) const char *swim_last_error(const void *) { return swim::g_last_error.c_str(); }\n')
    so = tmp_path / "libcodec_asan.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "swim_b200", "csrc", "swim_codec.cpp"), str(stub),
                        "-o", str(so)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cannot build the sanitizer variant: " + r.stderr[-300:])
    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan):
        pytest.skip("libasan.so not found")
    script = tmp_path / "fuzz.py"
    script.write_text(CODEC_FUZZ)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([sys.executable, str(script), ROOT, str(so)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout[-500:], r.stderr[-3000:])
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
    _, n_ok, n_bad = r.stdout.split()
    assert int(n_ok) > 300 and int(n_bad) > 300


EMU_ASAN_DRIVER = r'''
import sys
root, so = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, root + "/tests", root + "/tests/emu"]
import swim_b200._lib as L
import soak                      # scenario generators (builds / selects the plain emulator library on import) ...
L.SO_PATH, L._lib = so, None     # ... but this process runs the AddressSanitizer build
for seed in (31001, 31002, 31003):
    soak.one(seed)
soak.one_sharded(32001)
print("ok")
'''


def test_device_code_under_asan_on_the_emulator(tmp_path):
    """The CUDA sources compiled for the CPU (tests/emu) WITH AddressSanitizer: every cudaMalloc is an instrumented heap
    block, so an out-of-bounds access of a kernel — e.g. the speculative loads K2 issues before it knows an in-list's
    length — aborts here. Random scenarios against the oracle, one of them sharded over several ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan):
        pytest.skip("libasan.so not found")
    saved = (build_emu.OBJ, build_emu.SO)
    try:
        build_emu.OBJ, build_emu.SO = str(tmp_path / "build"), str(tmp_path / "libswim_emu_asan.so")
        try:
            so = build_emu.build(force=True, extra=("-fsanitize=address", "-fno-omit-frame-pointer"))
        except RuntimeError as e:
            pytest.skip(f"cannot build the sanitizer variant: {e}")
    finally:
        build_emu.OBJ, build_emu.SO = saved
    script = tmp_path / "drive.py"
    script.write_text(EMU_ASAN_DRIVER)
    env = dict(os.environ, LD_PRELOAD=libasan,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1")
    r = subprocess.run([sys.executable, str(script), ROOT, so], capture_output=True, text=True, env=env, timeout=1200)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-500:], r.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in r.stderr
