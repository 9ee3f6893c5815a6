"""Host-side mirror of the reference's Util.hs: configuration and Store construction.

  milliseconds Util.hs:23-24    parseConfig Util.hs:44-50    makeSelf Util.hs:93-101
  makeStore Util.hs:76-91       configure Util.hs:103-107    dumpStore Util.hs:64-74
`shuffle` (Util.hs:36-42) is device code (pick_remove in csrc/swim_device.cuh); sockets and
real-time sleeping (after / timeout / bindUDP / withSocket) have no simulator counterpart."""
from .types import Config, Liveness, Member, SockAddrInet


def milliseconds(n: int) -> int:
    return 1000 * n


def parseConfig() -> Config:
    return Config()


def makeSelf(_cfg: Config) -> Member:
    # Util.hs:93-101 — note the swapped port/host of `SockAddrInet 123 4000`
    return Member("myself", "localhost", SockAddrInet(123, 4000), Liveness.IsAliveC, 0, 0)


def makeStore(self_member: Member, cfg: Config, capacity: int = 31, **sim_kw):
    from .core import Store
    return Store(self_member, cfg, capacity=capacity, **sim_kw)


def configure(capacity: int = 31, **sim_kw):
    """configure (Util.hs:103-107): parseConfig -> makeSelf -> makeStore."""
    cfg = parseConfig()
    return makeStore(makeSelf(cfg), cfg, capacity=capacity, **sim_kw)


def dumpStore(store) -> str:
    s, i = store.seqNo, store.incarnation
    return f"(seqNo, inc) {(s, i)}\nmembers: {store.members_map()}\nself: {store.storeSelf}"
