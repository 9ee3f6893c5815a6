"""The pure integer helpers of swim_b200/csrc/swim_device.cuh (Philox4x32-10, the bounded draw, the r-th-set-bit pick
= one `shuffle` step) are host+device functions: tests/device_helpers_harness.cu builds the SAME source for the CPU
and this file checks it against the oracle's Philox and a plain-Python statement of Util.hs:36-42.
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
    L.h_nth_set.restype = L.h_bounded.restype = L.h_pick_remove.restype = C.c_uint32
    L.h_nth_set.argtypes = L.h_bounded.argtypes = [C.c_uint32, C.c_uint32]
    L.h_pick_remove.argtypes = [C.c_int, C.c_void_p, C.c_uint32]
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
