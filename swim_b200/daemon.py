"""A real-time single-node SWIM daemon on the scalar C ABI — SURVEY §8(f)-3: the equivalent of `Core.main`
(Core.hs:272-287) with the loop the reference leaves unfinished completed:

  receiver          UDP.sourceSocket $$ handleUDPMessage store =$= sinkTMChan gossip           (Core.hs:279-280)
  disseminate'      sourceTMChan gossip $$ disseminate store =$= UDP.sinkToSocket sock         (Core.hs:285-286)
  failureDetector'  failureDetector store (Core.hs:233-241) over probeNode' (Core.hs:243-269)

What is filled in: ack waits that time out and escalate (Q1/Q2: `unlessAck` has the polarity of its name here), the relay
of an indirect probe's Ack back to the requester (the comment at Core.hs:103-104), framing every datagram as an Envelope
(Q6), a gossip period in milliseconds (Q10), the suspicion timer (Core.hs:141 FIXME -> swim_tick_timers) and the piggyback
queue (Core.hs:136 FIXME -> swim_broadcast / swim_take_broadcasts): the buffered records ride on the period's Ping and
IndirectPings as one compound Envelope (Types.hs:96-119). The membership state machine — suspectNode / deadNode / aliveNode,
kRandomMembers, process — is the CUDA library's (swim_b200.core -> include/swim.h), the same code the bulk simulator runs.

One `Node` = one `Store` = one handle of the library; the handle is externally synchronised (one lock)."""
import socket
import struct
import threading
import time
from typing import Dict, List, Optional, Tuple

from . import core
from .types import (Ack, Alive, Broadcast, Config, Direct, Envelope, IndirectPing, Liveness, Member, Ping, SockAddrInet,
                    Suspect, decode, encode)


def ip_to_int(ip: str) -> int:
    return struct.unpack("!I", socket.inet_aton(ip))[0]


def int_to_ip(v: int) -> str:
    return socket.inet_ntoa(struct.pack("!I", v & 0xFFFFFFFF))


class Node:
    def __init__(self, name: str, host: str = "127.0.0.1", port: int = 0, period: float = 0.2, cfg: Optional[Config] = None,
                 capacity: int = 31, **sim_kw):
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)  # withSocket (bindUDP "127.0.0.1" 4000) (Core.hs:278)
        self.sock.bind((host, port))
        self.sock.settimeout(0.05)
        self.addr = SockAddrInet(self.sock.getsockname()[1], ip_to_int(host))
        self.period = period                       # gossipInterval (Util.hs:49), in seconds
        self.cfg = cfg or Config()
        me = Member(name, host, self.addr, Liveness.IsAliveC, 0, 0)
        self.store = core.Store(me, self.cfg, capacity=capacity, **sim_kw)
        self.lock = threading.RLock()
        # storeAckHandler (Types.hs:58): seqNo -> (event, senders whose Ack counts: the target, later the proxies). Own probes
        # take their seqNo from storeSeqNo, relayed ones from storeIncarnation (Q4), so the two number spaces overlap: an Ack
        # is matched by sequence number AND sender
        self.acks: Dict[int, Tuple[threading.Event, set]] = {}
        # proxy side: my ping's seqNo -> (requester, its seqNo, the target I pinged, when)
        self.relays: Dict[int, Tuple[SockAddrInet, int, SockAddrInet, float]] = {}
        self.stop_flag = threading.Event()
        self.threads: List[threading.Thread] = []
        self.stats = {"pings": 0, "acks": 0, "indirect": 0, "suspected": 0, "relayed": 0, "datagrams": 0, "decode_errors": 0, "send_errors": 0}

    # ------------------------------------------------------------------ membership bootstrap
    def join(self, seeds: List[Member]):
        """Start from a seed list and announce ourselves: Alive(incarnation, name, addr, port) is queued for gossip, the
        receivers' aliveNode adds the unknown member (Core.hs:206-216)."""
        with self.lock:
            self.store.set_members(seeds)
            core.disseminate(self.store, [Broadcast(Alive(self.store.incarnation, self.name, self.addr.host, self.addr.port))])

    @property
    def name(self) -> str:
        return self.store.storeSelf.memberName

    def member(self) -> Member:
        return self.store.storeSelf

    def members(self) -> Dict[str, Member]:
        with self.lock:
            return self.store.members_map()

    # ------------------------------------------------------------------ wire
    def _sendto(self, data: bytes, to: SockAddrInet):
        try:
            self.sock.sendto(data, (int_to_ip(to.host), to.port))
        except OSError:  # an unreachable / malformed address is a lost datagram, not the end of the daemon
            self.stats["send_errors"] += 1

    def _send(self, msgs, to: SockAddrInet):
        self._sendto(encode(Envelope(tuple(msgs))), to)

    def _flush(self, gossip):
        """disseminate (Core.hs:127-138): Direct -> the socket now; Broadcast -> the piggyback buffer."""
        with self.lock:
            out = core.disseminate(self.store, gossip)
        for data, to in out:
            self._sendto(data, to)

    # ------------------------------------------------------------------ receiver (Core.hs:279-280, 79-121)
    def handle_datagram(self, data: bytes, sender: SockAddrInet):
        self.stats["datagrams"] += 1
        try:
            msgs = decode(data).unEnvelope
        except Exception:  # handleDecodeErrors = either fail yield (Core.hs:86-87): a bad datagram is dropped, not fatal
            self.stats["decode_errors"] += 1
            return
        gossip = []
        for m in msgs:
            if isinstance(m, Ack):
                self.stats["acks"] += 1
                with self.lock:  # invokeAckHandler (Core.hs:220-221) ... and the relay promised at Core.hs:103-104
                    mine = self.acks.get(m.seqNo)
                    relay = self.relays.get(m.seqNo)
                    if relay is not None and relay[2] == sender:
                        del self.relays[m.seqNo]
                    else:
                        relay = None
                if mine and sender in mine[1]:
                    mine[0].set()
                if relay:
                    self.stats["relayed"] += 1
                    self._send([Ack(relay[1], ())], relay[0])
                continue
            with self.lock:
                out = core.process(self.store, sender, m)
            if isinstance(m, IndirectPing):
                for g in list(out):  # the forwarded Ping carries OUR sequence number (Q4): remember whom to answer
                    if isinstance(g, Direct) and isinstance(g.msg, Ping):
                        with self.lock:
                            self.relays[g.msg.seqNo] = (sender, m.seqNo, g.addr, time.monotonic())
                        # Q4 also means that relaying bumped storeIncarnation (Core.hs:105-108, pinned by Spec.hs:166-174):
                        # announce it. The state machine drops a Suspect/Dead about self whose incarnation is below
                        # storeIncarnation (Core.hs:151), so a proxy whose peers still hold the old number would never
                        # refute a false suspicion and would be declared Dead while alive.
                        sa = self.store.storeSelf.memberHostNew
                        out.append(Broadcast(Alive(g.msg.seqNo, self.store.storeSelf.memberName, sa.host, sa.port)))
            gossip.extend(out)
        self._flush(gossip)

    def _receiver(self):
        while not self.stop_flag.is_set():
            try:
                data, (ip, port) = self.sock.recvfrom(65535)  # UDP.sourceSocket sock 65535 (Core.hs:280)
            except socket.timeout:
                continue
            except OSError:
                return
            self.handle_datagram(data, SockAddrInet(port, ip_to_int(ip)))

    # ------------------------------------------------------------------ failure detector (Core.hs:233-269)
    def _wait_ack(self, seq: int, ev: threading.Event, seconds: float) -> bool:
        return ev.wait(seconds)  # race (timeout ...) (waitForAckOf store currSeqNo) (Core.hs:258-259)

    def tick(self):
        """One protocol period: timers, one probe (Q11), escalation, local suspicion."""
        with self.lock:
            core.tickTimers(self.store)                                   # Core.hs:141 FIXME
            seq = core.nextSeqNo(self.store)                              # Core.hs:238
            targets = core.kRandomMembers(self.store, 1, [])              # Core.hs:239
            payload = core.take_broadcasts(self.store)                    # Core.hs:136 FIXME: the compound message
            ev = threading.Event()
            allowed = {targets[0].memberHostNew} if targets else set()
            self.acks[seq] = (ev, allowed)
            now = time.monotonic()                                        # relays nobody answered are forgotten
            for k in [k for k, v in self.relays.items() if now - v[3] > 4 * self.period]:
                del self.relays[k]
        try:
            if not targets:
                return None
            m = targets[0]
            self.stats["pings"] += 1
            self._send([Ping(seq, m.memberName)] + payload, m.memberHostNew)                    # Core.hs:246
            if self._wait_ack(seq, ev, self.period / 3):
                return m.memberName, "ack"
            with self.lock:
                proxies = core.kRandomMembers(self.store, self.cfg.numToGossip, [])            # Core.hs:249
            ip = IndirectPing(seq, m.memberHostNew.host, m.memberHostNew.port, m.memberName)   # Core.hs:262-268
            for p in proxies:
                if p.memberName != m.memberName:
                    self.stats["indirect"] += 1
                    with self.lock:
                        allowed.add(p.memberHostNew)                                            # its relayed Ack counts
                    self._send([ip] + payload, p.memberHostNew)                                 # Core.hs:250
            if self._wait_ack(seq, ev, self.period / 3):
                return m.memberName, "indirect-ack"
            with self.lock:                                                                    # Core.hs:253-254
                suspect = core.suspectNode(self.store, Suspect(m.memberIncarnation, m.memberName))
                if suspect is not None:
                    self.stats["suspected"] += 1
                    core.disseminate(self.store, [Broadcast(suspect)])
            return m.memberName, "suspect"
        finally:
            with self.lock:
                self.acks.pop(seq, None)

    def _ticker(self):
        nxt = time.monotonic()
        while not self.stop_flag.is_set():
            nxt += self.period                       # after $ milliseconds (gossipInterval storeCfg) (Core.hs:237)
            self.tick()
            delay = nxt - time.monotonic()
            if delay > 0:
                self.stop_flag.wait(delay)
            else:
                nxt = time.monotonic()

    # ------------------------------------------------------------------ lifecycle (udpReceiver `race_` failureDetector')
    def start(self):
        for fn in (self._receiver, self._ticker):
            t = threading.Thread(target=fn, daemon=True)
            t.start()
            self.threads.append(t)
        return self

    def stop(self):
        self.stop_flag.set()
        for t in self.threads:
            t.join(timeout=5)
        self.sock.close()
        with self.lock:
            self.store.sim.close()


def main(argv=None):
    """python -m swim_b200.daemon NAME PORT [SEED_NAME:SEED_IP:SEED_PORT ...] — runs until interrupted (needs a CUDA device)."""
    import sys
    args = list(sys.argv[1:] if argv is None else argv)
    if len(args) < 2:
        raise SystemExit(main.__doc__)
    node = Node(args[0], port=int(args[1]))
    seeds = []
    for s in args[2:]:
        n, ip, p = s.split(":")
        seeds.append(Member(n, ip, SockAddrInet(int(p), ip_to_int(ip)), Liveness.IsAliveC, 0, 0))
    node.join(seeds)
    node.start()
    try:
        while True:
            time.sleep(5)
            from .util import dumpStore
            print(dumpStore(node.store), flush=True)    # what SIGUSR1 prints in the reference (Core.hs:275)
    except KeyboardInterrupt:
        node.stop()


if __name__ == "__main__":
    main()
