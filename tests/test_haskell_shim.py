"""haskell/SwimFFI.hs cannot be compiled here (no ghc). What can be checked without a compiler is: every
`foreign import` names a function include/swim.h declares, with the header's parameter count, and each
Storable instance has the C struct's size and pokes the C struct's field offsets."""
import ctypes as C
import os
import re

from swim_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS = open(os.path.join(ROOT, "haskell", "SwimFFI.hs")).read()
HEADER = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "swim.h")).read(), flags=re.S)


def c_prototypes():
    out = {}
    for name, params in re.findall(r"\b(swim_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", HEADER, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def hs_imports():
    out = {}
    for cname, sig in re.findall(r'foreign import ccall \w+\s+"(swim_[a-z0-9_]+)"\s+\w+\s*::\s*(.*)', HS):
        out[cname] = len([t for t in sig.split("->")]) - 1      # the last arrow type is the IO result
    return out


def test_foreign_imports_match_header():
    protos, imports = c_prototypes(), hs_imports()
    assert len(imports) >= 20
    for name, nargs in imports.items():
        assert name in protos, f"{name} is not declared in swim.h"
        assert protos[name] == nargs, f"{name}: header has {protos[name]} parameters, the shim passes {nargs}"


def instance_block(hs_type):
    m = re.search(r"instance Storable %s where(.*?)(?=\n\S)" % hs_type, HS, flags=re.S)
    assert m, hs_type
    return m.group(1)


def test_storable_layouts_match_structs():
    camel = lambda prefix, f: prefix + "".join(p.capitalize() for p in f.split("_"))
    for hs_type, prefix, ct in (("CConfig", "cfg", A.Config), ("CMember", "m", A.Member),
                                ("CMessage", "msg", A.Message), ("CGossip", "g", A.Gossip)):
        blk = instance_block(hs_type)
        assert int(re.search(r"sizeOf _ = (\d+)", blk).group(1)) == C.sizeof(ct)
        pokes = dict((v, int(o)) for o, v in re.findall(r"pokeByteOff p (\d+) (\w+)", blk))
        assert pokes, hs_type
        for fname, _ in ct._fields_:
            hs_field = camel(prefix, fname)
            if hs_field in pokes:
                assert pokes[hs_field] == getattr(ct, fname).offset, f"{hs_type}.{fname}"
        # every poked field is a struct field
        known = {camel(prefix, f) for f, _ in ct._fields_}
        assert set(pokes) <= known, set(pokes) - known
        peeks = [int(o) for o in re.findall(r"peekByteOff p (\d+)", blk)]
        assert sorted(peeks) == sorted(pokes.values()), hs_type
    assert "poke (p `plusPtr` 8) gMsg" in HS and A.Gossip.msg.offset == 8
