"""The C-ABI library loads and exports every symbol include/swim.h declares; the ctypes mirror
(swim_b200/_abi.py) has the header's struct layouts. No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from swim_b200 import _abi as A
from swim_b200._lib import SO_PATH, SwimError, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swim.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swim_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 35
    L = C.CDLL(SO_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in swim.h but not exported: {missing}"


def test_struct_layouts_match_header():
    probe = r'''
#include <stdio.h>
#include <stddef.h>
#include "swim.h"
#define S(t) printf(#t " %zu\n", sizeof(t))
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(swim_config_t); S(swim_member_t); S(swim_message_t); S(swim_gossip_t); S(swim_record_t); S(swim_event_t);
  S(swim_wire_message_t);
  O(swim_config_t, seed); O(swim_config_t, base_port); O(swim_member_t, last_change); O(swim_message_t, incarnation);
  O(swim_message_t, dead_from); O(swim_message_t, payload); O(swim_gossip_t, msg); O(swim_record_t, kind);
  O(swim_event_t, msg); O(swim_wire_message_t, node); O(swim_wire_message_t, dead_from);
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    py = {"swim_config_t": A.Config, "swim_member_t": A.Member, "swim_message_t": A.Message, "swim_gossip_t": A.Gossip,
          "swim_record_t": A.Record, "swim_event_t": A.Event, "swim_wire_message_t": A.WireMessage}
    for k, v in out.items():
        if "." in k:
            t, f = k.split(".")
            f = {"from": "from_"}.get(f, f)
            assert getattr(py[t], f).offset == int(v), k
        else:
            assert C.sizeof(py[k]) == int(v), k
    assert A.RECORD_DTYPE.itemsize == C.sizeof(A.Record) and A.EVENT_DTYPE.itemsize == C.sizeof(A.Event)


def test_abi_version_and_strerror():
    L = lib()
    assert L.swim_abi_version() == A.ABI_VERSION
    assert b"no CUDA device" in L.swim_strerror(A.ENODEV)
    cfg = A.Config()
    assert L.swim_config_default(C.byref(cfg)) == 0
    assert (cfg.k_indirect, cfg.fanout, cfg.pb_cap, cfg.suspicion_rounds, cfg.view_cap, cfg.world) == (3, 4, 8, 5, 32, 1)


def test_create_rejects_bad_config_and_has_no_cpu_fallback():
    from swim_b200.sim import Simulator, default_config
    for kw in (dict(view_cap=48), dict(k_indirect=8), dict(fanout=0), dict(fanout=5), dict(pb_cap=33),
               dict(suspicion_rounds=64), dict(retransmit=0), dict(n_nodes=0), dict(rank=1)):
        with pytest.raises(SwimError) as e:
            Simulator(default_config(**kw))
        assert e.value.code == A.EINVAL
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(SwimError) as e:  # the product path fails loudly: no CPU fallback
            Simulator(default_config())
        assert e.value.code == A.ENODEV
        assert "no CPU fallback" in str(e.value)


def test_topology_generators_are_valid_rows():
    import numpy as np
    from swim_b200.sim import generate_topology
    for kind, n, cap, deg in (("complete", 32, 32, 31), ("random", 5000, 32, 32), ("random", 300, 64, 50), ("ring", 1000, 32, 9)):
        nbr = generate_topology(kind, n, cap, deg, seed=5)
        assert nbr.shape == (n, cap)
        for i in (0, 1, n // 2, n - 1):
            row = nbr[i][nbr[i] != A.NO_MEMBER]
            assert len(row) == deg and i not in row and np.all(np.diff(row.astype(np.int64)) > 0) and row.max() < n
            assert np.all(nbr[i][deg:] == A.NO_MEMBER)
    a, b = generate_topology("random", 2000, 32, 32, seed=9), generate_topology("random", 2000, 32, 32, seed=9)
    assert np.array_equal(a, b) and not np.array_equal(a, generate_topology("random", 2000, 32, 32, seed=10))
    with pytest.raises(SwimError):
        generate_topology("complete", 64, 32)


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under swim_b200/ may reference it."""
    pkg = os.path.join(ROOT, "swim_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
