"""Scenarios for swim_b200/daemon.py shared by the emulated (CPU) and the hardware test."""
import socket
import time

from swim_b200.types import Ack, Envelope, IndirectPing, Liveness, Ping, SockAddrInet, decode, encode


def require_loopback():
    """Loopback UDP must work (it does in the build container and on the GPU boxes); skip instead of failing otherwise."""
    import pytest
    try:
        a = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        b = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        a.bind(("127.0.0.1", 0))
        b.bind(("127.0.0.1", 0))
        a.settimeout(2.0)
        b.sendto(b"x", a.getsockname())
        a.recvfrom(16)
        a.close()
        b.close()
    except OSError as e:
        pytest.skip(f"no loopback UDP here: {e}")


def recv_msgs(node, timeout=2.0):
    node.sock.settimeout(timeout)
    data, (ip, port) = node.sock.recvfrom(65535)
    from swim_b200.daemon import ip_to_int
    return list(decode(data).unEnvelope), SockAddrInet(port, ip_to_int(ip)), data


def scenario_probe_escalation_and_relay():
    """No threads: every datagram is moved by hand, so each step of probeNode' (Core.hs:243-269) is visible."""
    require_loopback()
    from swim_b200.daemon import Node
    a, b, c = (Node(n, period=0.09) for n in ("a", "b", "c"))
    try:
        a.join([b.member(), c.member()])
        b.join([a.member(), c.member()])
        c.join([a.member(), b.member()])
        # 1. a Ping meant for us is acknowledged to its sender (Core.hs:97-99); the join announcement rides on nothing yet
        b.handle_datagram(encode(Envelope((Ping(5, "b"),))), a.addr)
        msgs, frm, _ = recv_msgs(a)
        assert msgs == [Ack(5, ())] and frm == b.addr
        # ... one for somebody else is ignored (Core.hs:100-101)
        b.handle_datagram(encode(Envelope((Ping(6, "zzz"),))), a.addr)
        # 2. IndirectPing: the proxy pings the target with ITS OWN sequence number (Q4) and relays the Ack back
        c.handle_datagram(encode(Envelope((IndirectPing(7, b.addr.host, b.addr.port, "b"),))), a.addr)
        msgs, frm, raw = recv_msgs(b)
        assert len(msgs) == 1 and isinstance(msgs[0], Ping) and msgs[0].node == "b" and frm == c.addr
        proxy_seq = msgs[0].seqNo
        # relaying bumped c's storeIncarnation (Q4: the forwarded Ping's seqNo IS the new incarnation, Spec.hs:166-174); c
        # announces it, otherwise a later Suspect(old incarnation, c) would be dropped as stale by c itself (Core.hs:151)
        # and never refuted. The announcement overrides such a suspicion at everybody who hears it (Alive i > Suspect j).
        from swim_b200 import core as _core
        from swim_b200.types import Alive, Suspect
        with c.lock:
            pend = _core.pending_broadcasts(c.store)
            assert any(isinstance(m, Alive) and m.node == "c" and m.incarnation == proxy_seq for m in pend), pend
            assert _core.suspectNode(c.store, Suspect(proxy_seq - 1, "c")) is None        # stale for c itself ...
            refute = _core.suspectNode(c.store, Suspect(proxy_seq, "c"))                   # ... the announced one is refuted
            assert isinstance(refute, Alive) and refute.incarnation == proxy_seq + 1
        b.handle_datagram(raw, c.addr)
        msgs, frm, raw = recv_msgs(c)
        assert msgs == [Ack(proxy_seq, ())] and frm == b.addr
        # own probes count storeSeqNo, relayed ones storeIncarnation (Q4): the same number can be in flight twice. An own probe
        # of `a` with that very sequence number must not take b's Ack for its own
        import threading
        own = threading.Event()
        with c.lock:
            c.acks[proxy_seq] = (own, {a.addr})
        c.handle_datagram(raw, b.addr)
        assert not own.is_set()
        with c.lock:
            del c.acks[proxy_seq]
        msgs, frm, _ = recv_msgs(a)
        assert msgs == [Ack(7, ())] and frm == c.addr                       # the requester gets its own sequence number back
        assert c.stats["relayed"] == 1
        # 3. a probe nobody answers: Ping (with the piggybacked join announcement), IndirectPings, then local suspicion
        target, verdict = a.tick()
        assert verdict == "suspect" and target in ("b", "c")
        other = "c" if target == "b" else "b"
        view = a.members()
        assert view[target].memberAlive == Liveness.IsSuspectC and view[other].memberAlive == Liveness.IsAliveC
        tnode, onode = (b, c) if target == "b" else (c, b)
        msgs, frm, _ = recv_msgs(tnode)
        assert isinstance(msgs[0], Ping) and msgs[0].node == target
        assert any(type(m).__name__ == "Alive" and m.node == "a" for m in msgs[1:])  # compound Envelope: Ping + gossip
        msgs, frm, _ = recv_msgs(onode)
        assert isinstance(msgs[0], IndirectPing) and msgs[0].node == target and msgs[0].port == tnode.addr.port
        assert a.stats == {**a.stats, "pings": 1, "indirect": 1, "suspected": 1}
        # 4. S = 5 periods later the entry is Dead and the Dead is queued for gossip (Core.hs:141 FIXME)
        from swim_b200 import core
        for _ in range(5):
            with a.lock:
                core.tickTimers(a.store)
        assert a.members()[target].memberAlive == Liveness.IsDeadC
        with a.lock:
            kinds = [type(m).__name__ for m in core.pending_broadcasts(a.store)]
        assert "Dead" in kinds
    finally:
        for n in (a, b, c):
            n.stop()


def scenario_live_cluster_detects_a_crash(period=0.25, deadline=60.0):
    """Three daemons on loopback: they learn of each other through the join gossip, then one stops and the others declare
    it Dead (probe -> indirect probe -> Suspect -> S periods -> Dead), while staying Alive to each other."""
    require_loopback()
    from swim_b200.daemon import Node
    nodes = {n: Node(n, period=period) for n in ("a", "b", "c")}
    a, b, c = nodes["a"], nodes["b"], nodes["c"]
    try:
        # a knows nobody; b and c know only a: everything else must travel as gossip
        a.join([])
        b.join([a.member()])
        c.join([a.member()])
        for n in nodes.values():
            n.start()
        t0 = time.monotonic()
        while time.monotonic() - t0 < deadline:
            if all(len(n.members()) == 2 and all(m.memberAlive == Liveness.IsAliveC for m in n.members().values())
                   for n in nodes.values()):
                break
            time.sleep(period)
        else:
            raise AssertionError("the cluster did not form: " + repr({k: list(n.members()) for k, n in nodes.items()}))
        c.stop()
        t0 = time.monotonic()
        while time.monotonic() - t0 < deadline:
            va, vb = a.members(), b.members()
            # ... and the survivors hold each other Alive (on a loaded machine an Ack can miss its deadline: the false
            # suspicion is then refuted with a higher incarnation, which takes a few more periods)
            if va["c"].memberAlive == Liveness.IsDeadC and vb["c"].memberAlive == Liveness.IsDeadC and \
                    va["b"].memberAlive == Liveness.IsAliveC and vb["a"].memberAlive == Liveness.IsAliveC:
                break
            time.sleep(period)
        else:
            raise AssertionError("the crash was not detected: " + repr((a.members(), b.members())))
        assert a.stats["acks"] > 0 and b.stats["acks"] > 0
        assert a.stats["decode_errors"] == b.stats["decode_errors"] == 0
    finally:
        for n in (a, b):
            n.stop()
