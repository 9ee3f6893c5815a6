"""The pure integer helpers of swim_b200/csrc/swim_device.cuh (Philox4x32-10, r-th-set-bit pick = `shuffle` step,
round-robin pick) are host+device functions: tests/device_helpers_harness.cu builds the SAME source for the CPU
and this file checks it against the oracle's Philox and against plain-Python statements of the selection rules.
No GPU needed (nvcc compiles the host side)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle.oracle import philox

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dh(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    so = tmp_path_factory.mktemp("dh") / "libdevice_helpers.so"
    r = subprocess.run(["nvcc", "-O1", "-std=c++17", "-arch=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(so),
                        os.path.join(HERE, "device_helpers_harness.cu")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("host build of the device helpers failed:\n" + r.stderr[-2000:])
    L = C.CDLL(str(so))
    L.h_nth_set.restype = L.h_xor_permute.restype = L.h_pick_remove.restype = L.h_rr_pick.restype = L.h_bounded.restype = C.c_uint32
    L.h_nth_set.argtypes = L.h_xor_permute.argtypes = L.h_bounded.argtypes = [C.c_uint32, C.c_uint32]
    L.h_pick_remove.argtypes = [C.c_int, C.c_void_p, C.c_uint32]
    L.h_rr_pick.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
    L.h_philox.argtypes = [C.c_void_p] * 3
    return L


def test_philox_matches_oracle_and_random123(dh):
    rng = np.random.default_rng(0)
    cases = [([0, 0, 0, 0], [0, 0]), ([0xffffffff] * 4, [0xffffffff] * 2),
             ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])]
    cases += [(rng.integers(0, 2 ** 32, 4).tolist(), rng.integers(0, 2 ** 32, 2).tolist()) for _ in range(200)]
    for ctr, key in cases:
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        dh.h_philox(c, k, o)
        assert list(o) == philox(ctr, key)
    c, k, o = (C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), (C.c_uint32 * 4)()
    dh.h_philox(c, k, o)
    assert list(o) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]  # Random123 known answer


def test_bounded_draw(dh):  # randomR (0, L-1) (Util.hs:40) as mulhi(x, L)
    rng = np.random.default_rng(4)
    for _ in range(2000):
        x, L = int(rng.integers(0, 2 ** 32)), int(rng.integers(1, 300))
        got = dh.h_bounded(x, L)
        assert got == (x * L) >> 32 and 0 <= got < L


def bits(words):
    return [w * 32 + b for w, x in enumerate(words) for b in range(32) if x >> b & 1]


def test_nth_set_and_pick_remove_are_the_shuffle_step(dh):
    """Util.hs:36-42: pick index r of the remaining candidates (ascending slot order), remove it, keep the order."""
    rng = np.random.default_rng(1)
    for _ in range(2000):
        m = int(rng.integers(1, 2 ** 32))
        r = int(rng.integers(0, bin(m).count("1")))
        assert dh.h_nth_set(m, r) == bits([m])[r]
    for W in (1, 2, 4, 8):
        for _ in range(300):
            words = [int(x) for x in rng.integers(0, 2 ** 32, W)]
            if rng.random() < 0.3:
                words = [w & int(rng.integers(0, 2 ** 32)) & int(rng.integers(0, 2 ** 32)) for w in words]
            cand = bits(words)
            arr = (C.c_uint32 * W)(*words)
            while cand:
                r = int(rng.integers(0, len(cand)))
                assert dh.h_pick_remove(W, arr, r) == cand.pop(r)
                assert bits(list(arr)) == cand


def test_xor_permute(dh):
    rng = np.random.default_rng(2)
    for _ in range(2000):
        m, b = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 32))
        want = sum(((m >> (i ^ b)) & 1) << i for i in range(32))
        assert dh.h_xor_permute(m, b) == want


def test_rr_pick_is_first_alive_slot_in_xor_order(dh):
    """slot(p) = p xor b, p = (round + r) mod cap, first Alive slot at or after p (cyclic) — the statement the oracle
    implements as a literal loop (oracle/swim_oracle.c tick_node) and tests/test_variants.py pins."""
    rng = np.random.default_rng(3)
    for W in (1, 2, 4, 8):
        cap = 32 * W
        for it in range(1500):
            words = [int(x) for x in rng.integers(0, 2 ** 32, W)]
            dens = it % 4
            for _ in range(dens * 2):
                words = [w & int(rng.integers(0, 2 ** 32)) for w in words]
            if not any(words):
                words[int(rng.integers(0, W))] = 1 << int(rng.integers(0, 32))
            alive = set(bits(words))
            word, rnd = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32))
            b, r = word & (cap - 1), (word >> 16) & (cap - 1)
            p = (rnd + r) % cap
            want = next(((p + x) % cap) ^ b for x in range(cap) if (((p + x) % cap) ^ b) in alive)
            arr = (C.c_uint32 * W)(*words)
            assert dh.h_rr_pick(W, arr, word, rnd) == want, (W, words, word, rnd)
