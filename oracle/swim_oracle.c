/*
 * swim_oracle.c — CPU ORACLE for swim-b200.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
 * legs may build, load or call this file. The product library (swim_b200/csrc) never
 * includes or links it and has no CPU fallback.
 *
 * What it is: a plain-C, loop-per-node restatement of the protocol rules that
 * jpfuentes2/swim (Haskell, commit 4320f07) defines in src/Core.hs, src/Types.hs and
 * src/Util.hs, executed under the synchronous-round model of DESIGN.md §2 (the reference
 * is a real-time UDP daemon and cannot be compiled here: no GHC/stack/cabal, no network).
 *
 * PARITY PINNING
 *   pinned   : the rules exercised by the reference's own test/Spec.hs (removeDeadNodes
 *              98-106, kRandomMembers 111-139, Ping->Ack 150-153, Ping-other 155-158,
 *              IndirectPing->Ping 166-174) — tests/test_oracle_kat.py replays them.
 *   derived  : suspectOrDeadNode' (Core.hs:142-187) known-answer vectors E1..E16 worked by
 *              hand from the source (the reference leaves those tests `pending`,
 *              Spec.hs:176-183).
 *   UNPINNED : the random stream (reference uses the unseedable global StdGen, Util.hs:40;
 *              here Philox4x32-10), and the multi-round loop itself, because the
 *              reference's loop never escalates, drops broadcasts and has no timer
 *              (SURVEY §0.2 Q1,Q2,Q5,Q7,Q8). "parity unpinned" for those parts: they are
 *              completed per the reference's own comments, each completion tagged [Qn].
 *
 * Completions of the unfinished reference loop:
 *   [Q1,Q2] escalate on NO ack (Core.hs:245-253 comments; literal polarity is inverted)
 *   [Q3]    IndirectPing.node = memberName m (literal: `show m`)
 *   [Q4]    bulk rounds do not bump a proxy's incarnation; the scalar handle_message does
 *           (pinned by Spec.hs:166-174)
 *   [Q5]    Broadcast -> bounded piggyback buffer (Core.hs:136 FIXME), B records, each sent
 *           T times, newest first, newer record about a member replaces the older one
 *   [Q7]    Alive(i) about a known member applies iff i > stored incarnation (SWIM §4.2);
 *           unknown member: inserted by the scalar call (Core.hs:206-216), ignored by bulk
 *           rounds (static view graph)
 *   [Q8]    suspicion timer: S rounds after entering Suspect the entry becomes Dead and
 *           Dead(inc, name, from=self) is broadcast (Core.hs:141 FIXME)
 *   [Q9]    refutation incarnation = max(storeIncarnation, accused) + 1 (literal diverges)
 *   [Q11]   one probe target per node per period; k = numToGossip proxies
 *   [Q14]   guard order and >= / ignore rules of Core.hs:151 kept VERBATIM
 *   self    a node's own entry is virtual: (Alive, storeIncarnation); it is not in its row
 *           (Util.hs:78 makeStore starts with an empty map that never holds self)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/swim.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Philox4x32-10
 * Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11). Written with a
 * 64-bit product on purpose (the CUDA side uses __umulhi) so the two are independent. */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

EXPORT void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr, key, out);
}

/* counter layout: (a, id, purpose, block); key = seed lo, hi. The target draw and the
 * direct-leg loss draw of node i are word (i & 3) of the block with id = i >> 2 (four nodes share a
 * block); proxy draws and indirect-leg loss draws use per-node blocks. */
enum { P_TARGET = 0, P_LOSS0 = 1, P_SCALAR = 2, P_TOPO = 3, P_PROXY = 4, P_LOSS = 5, P_RR = 6, P_CHURN = 7, P_TARGETS = 8, P_LOSSD = 9 };

/* bounded draw: floor(x * L / 2^32) — `randomR (0, L-1)` of Util.hs:40 on our stream */
static uint32_t bounded(uint32_t x, uint32_t L) { return (uint32_t)(((uint64_t)x * L) >> 32); }

/* ------------------------------------------------------------------ state */
typedef swim_record_t rec_t;

typedef struct {
  uint32_t dst, src, cnt;
  rec_t recs[SWIM_MAX_PB];
} env_t; /* one piggyback Envelope (Types.hs:90) in flight */

typedef struct oracle {
  swim_config_t cfg;
  uint32_t N, cap, k, fanout, B, S, T, loss_ppm;
  uint32_t first, n; /* local shard [first, first+n) */
  uint32_t key[2];
  uint32_t round;
  uint64_t scalar_calls;
  uint8_t *alive;     /* [N] global truth */
  uint32_t *back_at;  /* [N] churn: round at which a crashed process rejoins (0 = none) */
  uint32_t *last_crash, *last_rejoin; /* [N] round of the last up -> down / down -> up transition (0 = never) */
  uint32_t *self_inc; /* [n] */
  uint32_t *seqno;    /* [n] */
  uint32_t *nbr;      /* [n*cap] */
  uint8_t *state;     /* [n*cap] */
  uint8_t *timer;     /* [n*cap] */
  uint8_t *conf;      /* [n*cap] suspicion_max > 0: further Suspect messages seen about a suspected member (0..3) */
  uint32_t s_arm;     /* rounds a new suspicion starts with */
  uint32_t lg_delta[4]; /* lg_delta[c]: what the c-th confirmation takes off the countdown */
  uint32_t *vinc;     /* [n*cap] */
  uint32_t *vlast;    /* [n*cap] */
  rec_t *pb;          /* [n*B] */
  uint8_t *pb_cnt;    /* [n] */
  rec_t *out;         /* [n*B] snapshot sent this round */
  uint8_t *out_cnt;   /* [n] */
  uint32_t *send_to;  /* [n*(1+k)] recipient ids this round, SWIM_NO_MEMBER = none */
  /* events */
  swim_event_t *ev;
  size_t n_ev, cap_ev;
  /* envelopes */
  env_t *outbox; /* cross-shard envelopes produced by round_begin */
  size_t n_outbox, cap_outbox;
  env_t *inbox; /* envelopes for local receivers this round */
  size_t n_inbox, cap_inbox;
  uint64_t ctr[SWIM_CTR__COUNT];
  int view_set;
} oracle_t;

static uint32_t owner_first(uint32_t N, uint32_t world, uint32_t rank) {
  /* contiguous shards of ceil(N/world) nodes (the last one may be shorter) */
  uint64_t per = ((uint64_t)N + world - 1) / world;
  uint64_t f = per * rank;
  return (uint32_t)(f > N ? N : f);
}

EXPORT oracle_t *oracle_create(const swim_config_t *cfg) {
  if (!cfg || cfg->abi_version != SWIM_ABI_VERSION) return NULL;
  if (cfg->n_nodes == 0 || cfg->world == 0 || cfg->rank >= cfg->world) return NULL;
  if (cfg->view_cap == 0 || cfg->view_cap > SWIM_MAX_VIEW || cfg->view_cap % 32) return NULL;
  if (cfg->k_indirect > SWIM_MAX_K || cfg->fanout < 1 || cfg->fanout > 1 + cfg->k_indirect) return NULL;
  if (cfg->pb_cap < 1 || cfg->pb_cap > SWIM_MAX_PB) return NULL;
  if (cfg->suspicion_rounds < 1 || cfg->suspicion_rounds > SWIM_MAX_TIMER) return NULL;
  if (cfg->retransmit < 1 || cfg->retransmit > 255 || cfg->loss_ppm > 1000000u) return NULL;
  if (cfg->flags & ~SWIM_F__ALL) return NULL;
  if ((cfg->flags & SWIM_F_ROUND_ROBIN) && (cfg->view_cap & (cfg->view_cap - 1))) return NULL; /* xor order */
  if (cfg->churn_ppm > 1000000u || (cfg->churn_ppm && (cfg->rejoin_min < 1 || cfg->rejoin_max < cfg->rejoin_min))) return NULL;
  if (cfg->probes_per_round < 1 || cfg->probes_per_round > SWIM_MAX_PROBES) return NULL;
  if (cfg->suspicion_max && (cfg->suspicion_max < cfg->suspicion_rounds || cfg->suspicion_max > SWIM_MAX_TIMER_LIFEGUARD)) return NULL;
  oracle_t *o = (oracle_t *)calloc(1, sizeof *o);
  o->cfg = *cfg;
  o->N = cfg->n_nodes; o->cap = cfg->view_cap; o->k = cfg->k_indirect; o->fanout = cfg->fanout;
  o->B = cfg->pb_cap; o->S = cfg->suspicion_rounds; o->T = cfg->retransmit; o->loss_ppm = cfg->loss_ppm;
  o->key[0] = (uint32_t)cfg->seed; o->key[1] = (uint32_t)(cfg->seed >> 32);
  o->first = owner_first(o->N, cfg->world, cfg->rank);
  o->n = owner_first(o->N, cfg->world, cfg->rank + 1) - o->first;
  size_t n = o->n ? o->n : 1, slots = n * o->cap;
  o->alive = (uint8_t *)malloc(o->N); memset(o->alive, 1, o->N); /* every node up */
  o->back_at = (uint32_t *)calloc(o->N, 4);
  o->last_crash = (uint32_t *)calloc(o->N, 4);
  o->last_rejoin = (uint32_t *)calloc(o->N, 4);
  o->self_inc = (uint32_t *)calloc(n, 4);                        /* Util.hs:80 */
  o->seqno = (uint32_t *)calloc(n, 4);                           /* Util.hs:79 */
  o->nbr = (uint32_t *)malloc(slots * 4); memset(o->nbr, 0xFF, slots * 4); /* Util.hs:78: empty */
  o->state = (uint8_t *)malloc(slots); memset(o->state, SWIM_VACANT, slots);
  o->timer = (uint8_t *)calloc(slots, 1);
  o->conf = (uint8_t *)calloc(slots, 1);
  o->s_arm = cfg->suspicion_max ? cfg->suspicion_max : cfg->suspicion_rounds;
  if (cfg->suspicion_max) { /* Lifeguard's timeout(c) = max - (max - min) log(c+1)/log(K+1), K = 3, in 1/256ths: 0, .5, log(3)/log(4), 1 */
    static const uint32_t frac[4] = {0, 128, 203, 256};
    uint32_t T[4];
    for (int c = 0; c < 4; ++c) T[c] = cfg->suspicion_max - ((cfg->suspicion_max - cfg->suspicion_rounds) * frac[c] + 128) / 256;
    for (int c = 1; c < 4; ++c) o->lg_delta[c] = T[c - 1] - T[c];
  }
  o->vinc = (uint32_t *)calloc(slots, 4);
  o->vlast = (uint32_t *)calloc(slots, 4);
  o->pb = (rec_t *)calloc(n * o->B, sizeof(rec_t));
  o->pb_cnt = (uint8_t *)calloc(n, 1);
  o->out = (rec_t *)calloc(n * o->B, sizeof(rec_t));
  o->out_cnt = (uint8_t *)calloc(n, 1);
  o->send_to = (uint32_t *)malloc(n * (1 + o->k) * 4);
  return o;
}

EXPORT void oracle_destroy(oracle_t *o) {
  if (!o) return;
  free(o->alive); free(o->back_at); free(o->last_crash); free(o->last_rejoin); free(o->self_inc); free(o->seqno); free(o->nbr); free(o->state); free(o->timer); free(o->conf);
  free(o->vinc); free(o->vlast); free(o->pb); free(o->pb_cnt); free(o->out); free(o->out_cnt);
  free(o->send_to); free(o->ev); free(o->outbox); free(o->inbox); free(o);
}

EXPORT void oracle_local_range(const oracle_t *o, uint32_t *first, uint32_t *count) {
  *first = o->first; *count = o->n;
}

/* bulk `swapTVar storeMembers` (Spec.hs:101): global id matrix, every member Alive/inc 0 */
EXPORT int oracle_set_view(oracle_t *o, const uint32_t *nbr_global) {
  for (uint32_t l = 0; l < o->n; ++l) {
    const uint32_t *row = nbr_global + (size_t)(o->first + l) * o->cap;
    uint32_t prev = 0; int seen = 0, vacant = 0;
    for (uint32_t s = 0; s < o->cap; ++s) {
      uint32_t m = row[s];
      if (m == SWIM_NO_MEMBER) { vacant = 1; continue; }
      if (vacant || m >= o->N || m == o->first + l || (seen && m <= prev)) return SWIM_EINVAL;
      prev = m; seen = 1;
    }
    for (uint32_t s = 0; s < o->cap; ++s) {
      size_t x = (size_t)l * o->cap + s;
      o->nbr[x] = row[s];
      o->state[x] = row[s] == SWIM_NO_MEMBER ? SWIM_VACANT : SWIM_ALIVE;
      o->timer[x] = 0; o->conf[x] = 0; o->vinc[x] = 0; o->vlast[x] = 0;
    }
  }
  o->view_set = 1;
  return SWIM_OK;
}

/* ------------------------------------------------------------------ piggyback buffer [Q5]
 * enqueue = what `Broadcast m` should have done (Core.hs:131,136-138). */
static void pb_enqueue(oracle_t *o, uint32_t l, rec_t r, uint64_t *ctr) {
  rec_t *q = o->pb + (size_t)l * o->B;
  uint32_t cnt = o->pb_cnt[l];
  r.ttl = (uint8_t)o->T; r._pad = 0;
  /* a newer record about the same member replaces the older one */
  for (uint32_t x = 0; x < cnt; ++x)
    if (q[x].member == r.member) {
      memmove(q + x, q + x + 1, (cnt - x - 1) * sizeof(rec_t));
      --cnt;
      break;
    }
  if (cnt == o->B) { --cnt; ctr[SWIM_CTR_PB_DROPPED]++; } /* oldest falls off */
  memmove(q + 1, q, cnt * sizeof(rec_t));
  q[0] = r;
  ++cnt;
  memset(q + cnt, 0, (o->B - cnt) * sizeof(rec_t));
  o->pb_cnt[l] = (uint8_t)cnt;
}

/* ------------------------------------------------------------------ state machine
 * suspectOrDeadNode' (Core.hs:142-187) + aliveNode (Core.hs:197-218, [Q7]).
 * Returns 1 and fills *rb with the message to re-broadcast (`Just`), else 0 (`Nothing`).
 * allow_insert: scalar aliveNode adds unknown members (Core.hs:206-216); *err on full row. */
/* net: the record came in a datagram (receive phase, injected event, scalar `process`), not from this node's own probe.
 * [Lifeguard] with suspicion_max, a Suspect received about a member that is already Suspect (same or newer incarnation)
 * is still `Nothing` (Core.hs:151,183) but counts as a confirmation: the countdown loses lg_delta[c], never below 1. */
static void confirm_suspicion(oracle_t *o, size_t x) {
  if (o->conf[x] >= 3) return;
  const uint32_t c = ++o->conf[x], dlt = o->lg_delta[c];
  o->timer[x] = (uint8_t)(o->timer[x] > dlt ? o->timer[x] - dlt : 1);
}

static int apply_record(oracle_t *o, uint32_t l, rec_t r, int allow_insert, int net, rec_t *rb, int *err, uint64_t *ctr) {
  uint32_t self = o->first + l;
  uint32_t *ids = o->nbr + (size_t)l * o->cap;
  uint8_t *st = o->state + (size_t)l * o->cap;
  uint8_t *tm = o->timer + (size_t)l * o->cap;
  uint32_t *inc = o->vinc + (size_t)l * o->cap;
  uint32_t *last = o->vlast + (size_t)l * o->cap;
  uint8_t *cf = o->conf + (size_t)l * o->cap;
  const int lg = o->cfg.suspicion_max != 0;
  const int strict = (o->cfg.flags & SWIM_F_STRICT_OVERRIDE) != 0;
  if (err) *err = 0;
  if (r.member == self) {
    /* own entry is virtual (Alive, storeIncarnation) */
    if (r.kind == SWIM_MSG_ALIVE) return 0; /* our own refutation coming back */
    /* Core.hs:151: `i < memberIncarnation m || livenessCheck m` -> ignore (self is Alive).
     * STRICT_OVERRIDE: a Confirm overrides whatever the others hold, so even a stale one must be refuted. */
    if (r.incarnation < o->self_inc[l] && !(strict && r.kind == SWIM_MSG_DEAD)) return 0;
    /* Core.hs:155-166 refute; [Q9] terminating nextIncarnation' */
    uint32_t base = o->self_inc[l] > r.incarnation ? o->self_inc[l] : r.incarnation;
    o->self_inc[l] = base + 1;
    ctr[SWIM_CTR_REFUTES]++;
    rb->member = self; rb->incarnation = base + 1; rb->from = 0; rb->kind = SWIM_MSG_ALIVE;
    rb->ttl = 0; rb->_pad = 0;
    return 1;
  }
  /* Core.hs:144-145: find ((== name) . memberName) ms */
  uint32_t s;
  for (s = 0; s < o->cap; ++s)
    if (st[s] != SWIM_VACANT && ids[s] == r.member) break;
  if (s == o->cap) {
    /* Core.hs:147-148 we don't know this node. ignore.  (Alive: Core.hs:206-216 adds it) */
    if (r.kind != SWIM_MSG_ALIVE || !allow_insert) return 0;
    uint32_t used = 0;
    while (used < o->cap && st[used] != SWIM_VACANT) ++used;
    if (used == o->cap) { if (err) *err = SWIM_ECAP; return 0; }
    uint32_t pos = 0;
    while (pos < used && ids[pos] < r.member) ++pos; /* Map.insert keeps key order */
    for (uint32_t x = used; x > pos; --x) {
      ids[x] = ids[x - 1]; st[x] = st[x - 1]; tm[x] = tm[x - 1]; cf[x] = cf[x - 1]; inc[x] = inc[x - 1]; last[x] = last[x - 1];
    }
    ids[pos] = r.member; st[pos] = SWIM_ALIVE; tm[pos] = 0; cf[pos] = 0; inc[pos] = r.incarnation; last[pos] = o->round;
    *rb = r;
    return 1;
  }
  if (strict) {
    /* SWIM paper §4.2 (SWIM_F_STRICT_OVERRIDE):
     *   {Suspect Mj, i} overrides {Suspect Mj, j} i > j and {Alive Mj, j} i >= j
     *   {Confirm Mj, i} overrides {Alive Mj, j} and {Suspect Mj, j}, any i and j
     *   {Alive Mj, i}   overrides {Suspect Mj, j} and {Alive Mj, j}, i > j (and Dead, so that a rejoin is seen: [Q7]) */
    switch (r.kind) {
      case SWIM_MSG_SUSPECT:
        if (st[s] == SWIM_DEAD) return 0;
        if (st[s] == SWIM_ALIVE ? r.incarnation < inc[s] : r.incarnation <= inc[s]) {
          if (lg && net && st[s] == SWIM_SUSPECT && r.incarnation == inc[s]) confirm_suspicion(o, (size_t)l * o->cap + s);
          return 0;
        }
        inc[s] = r.incarnation; st[s] = SWIM_SUSPECT; tm[s] = (uint8_t)o->s_arm; cf[s] = 0; last[s] = o->round;
        *rb = r;
        return 1;
      case SWIM_MSG_DEAD:
        if (st[s] == SWIM_DEAD) return 0;
        if (r.incarnation > inc[s]) inc[s] = r.incarnation; /* keeps max(i, j) */
        st[s] = SWIM_DEAD; tm[s] = 0; cf[s] = 0; last[s] = o->round;
        *rb = r;
        return 1;
      case SWIM_MSG_ALIVE:
        if (r.incarnation <= inc[s]) return 0;
        inc[s] = r.incarnation; st[s] = SWIM_ALIVE; tm[s] = 0; cf[s] = 0; last[s] = o->round;
        *rb = r;
        return 1;
    }
    return 0;
  }
  switch (r.kind) {
    case SWIM_MSG_SUSPECT:
      /* Core.hs:151 + livenessCheck IsSuspect = memberAlive /= IsAliveC (Core.hs:183) */
      if (r.incarnation < inc[s] || st[s] != SWIM_ALIVE) {
        if (lg && net && st[s] == SWIM_SUSPECT && r.incarnation >= inc[s]) confirm_suspicion(o, (size_t)l * o->cap + s);
        return 0;
      }
      inc[s] = r.incarnation; st[s] = SWIM_SUSPECT; tm[s] = (uint8_t)o->s_arm; cf[s] = 0; /* [Q8] arm timer */
      last[s] = o->round;                                                  /* Core.hs:176 */
      *rb = r;                                                             /* Core.hs:179 */
      return 1;
    case SWIM_MSG_DEAD:
      /* livenessCheck IsDead = memberAlive == IsDeadC (Core.hs:184) */
      if (r.incarnation < inc[s] || st[s] == SWIM_DEAD) return 0;
      inc[s] = r.incarnation; st[s] = SWIM_DEAD; tm[s] = 0; cf[s] = 0; last[s] = o->round;
      *rb = r; /* deadFrom preserved */
      return 1;
    case SWIM_MSG_ALIVE:
      /* [Q7] SWIM §4.2: Alive(i) overrides Suspect(j)/Alive(j)/(Dead j) iff i > j */
      if (r.incarnation <= inc[s]) return 0;
      inc[s] = r.incarnation; st[s] = SWIM_ALIVE; tm[s] = 0; cf[s] = 0; last[s] = o->round;
      *rb = r;
      return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------ selection
 * kRandomMembers (Core.hs:69-74) over `shuffle` (Util.hs:36-42): the candidate list is the
 * row in ascending key order filtered by isAlive; each draw picks index r in [0,len) and
 * removes it preserving order. Literal list surgery on purpose. */
static uint32_t shuffle_take(uint32_t *list, uint32_t len, uint32_t n, const uint32_t *draws, uint32_t *outp) {
  uint32_t taken = 0;
  while (len > 0 && taken < n) {
    uint32_t r = bounded(draws[taken], len); /* rand <- randomR (0, length as - 1) */
    outp[taken++] = list[r];                 /* let (l, a:r) = splitAt rand as */
    memmove(list + r, list + r + 1, (len - r - 1) * sizeof(uint32_t)); /* l <> r */
    --len;
  }
  return taken;
}

static void draws_for(const oracle_t *o, uint32_t a, uint32_t node, uint32_t purpose, uint32_t n, uint32_t *out) {
  for (uint32_t b = 0; b * 4 < n; ++b) {
    uint32_t ctr[4] = {a, node, purpose, b}, w[4];
    philox4x32_10(ctr, o->key, w);
    for (uint32_t x = 0; x < 4 && b * 4 + x < n; ++x) out[b * 4 + x] = w[x];
  }
}

/* T1 [Q8] suspicion countdown of node `l`, ascending slot order (`-- FIXME: need a timer to mark this node as dead
 * after suspect timeout`, Core.hs:141). Returns the number of entries that expired. */
static uint32_t tick_timers(oracle_t *o, uint32_t l, uint64_t *ctr) {
  uint32_t self = o->first + l, expired = 0;
  uint32_t *ids = o->nbr + (size_t)l * o->cap;
  uint8_t *st = o->state + (size_t)l * o->cap;
  uint8_t *tm = o->timer + (size_t)l * o->cap;
  uint32_t *inc = o->vinc + (size_t)l * o->cap;
  uint32_t *last = o->vlast + (size_t)l * o->cap;
  for (uint32_t s = 0; s < o->cap; ++s)
    if (st[s] == SWIM_SUSPECT && --tm[s] == 0) {
      st[s] = SWIM_DEAD; last[s] = o->round; o->conf[(size_t)l * o->cap + s] = 0;
      rec_t d = {ids[s], inc[s], self, SWIM_MSG_DEAD, 0, 0};
      pb_enqueue(o, l, d, ctr);
      ctr[SWIM_CTR_DEAD_TIMEOUT]++;
      ++expired;
    }
  return expired;
}

/* ------------------------------------------------------------------ phase T: one tick
 * failureDetector (Core.hs:233-241) + probeNode' (Core.hs:243-269) for node `l`. */
static void tick_node(oracle_t *o, uint32_t l, uint64_t *ctr) {
  uint32_t self = o->first + l;
  uint32_t *to = o->send_to + (size_t)l * (1 + o->k);
  for (uint32_t f = 0; f <= o->k; ++f) to[f] = SWIM_NO_MEMBER;
  o->out_cnt[l] = 0;
  if (!o->alive[self]) return; /* a crashed process does nothing */
  uint32_t *ids = o->nbr + (size_t)l * o->cap;
  uint8_t *st = o->state + (size_t)l * o->cap;
  uint32_t *inc = o->vinc + (size_t)l * o->cap;

  tick_timers(o, l, ctr); /* T1 */

  /* T2 kRandomMembers store P [] (Core.hs:239; P = probes_per_round, [Q11]: 1 by default, the reference's literal
   * numToGossip otherwise) and, per failed probe, kRandomMembers store k [] (Core.hs:249) */
  const uint32_t P = o->cfg.probes_per_round, K = o->k;
  uint32_t cand[SWIM_MAX_VIEW], L = 0;
  for (uint32_t s = 0; s < o->cap; ++s)
    if (st[s] == SWIM_ALIVE) cand[L++] = s; /* filter isAlive over Map.elems */
  if (L == 0) return;
  uint32_t grp[4], tdraw[SWIM_MAX_PROBES], pdraw[SWIM_MAX_PROBES * SWIM_MAX_K + 1];
  draws_for(o, o->round, self >> 2, P_TARGET, 4, grp);
  tdraw[0] = grp[self & 3];                                      /* probe 0: the group stream (four nodes per block) */
  if (P > 1) draws_for(o, o->round, self, P_TARGETS, P - 1, tdraw + 1); /* probes 1..: a per-node stream */
  draws_for(o, o->round, self, P_PROXY, P * K, pdraw);           /* probe j's proxies: draws j*k .. j*k + k-1 */
  uint32_t targets[SWIM_MAX_PROBES], nt = P < L ? P : L, tmp[SWIM_MAX_VIEW];
  if (o->cfg.flags & SWIM_F_ROUND_ROBIN) {
    /* `-- FIXME: move from random to robust scheme` (Core.hs:232), SWIM paper §4.3. Rounds are grouped in epochs of
     * `cap` rounds; within epoch e node i walks its view in the order  slot(p) = p xor b,  p = (round + r) mod cap,
     * (b, r) drawn once per (epoch, node); the target is the first Alive slot at or after p in that order (cyclic).
     * Every slot position comes up exactly once per epoch, so an Alive member waits < 2 cap rounds for a probe.
     * Further probes of the period take the next Alive slots of the same walk. */
    draws_for(o, o->round / o->cap, self >> 2, P_RR, 4, grp);
    const uint32_t word = grp[self & 3], b = word & (o->cap - 1), r = (word >> 16) & (o->cap - 1);
    const uint32_t p = (o->round + r) & (o->cap - 1);
    uint32_t got = 0;
    for (uint32_t x = 0; x < o->cap && got < nt; ++x) {
      const uint32_t slot = ((p + x) & (o->cap - 1)) ^ b;
      if (st[slot] == SWIM_ALIVE) targets[got++] = slot;
    }
  } else {
    memcpy(tmp, cand, L * 4);
    shuffle_take(tmp, L, nt, tdraw, targets); /* ONE shuffle, take P (Core.hs:239) */
  }
  memcpy(tmp, cand, L * 4); /* a fresh shuffle: target and self are not excluded (Core.hs:249) */
  uint32_t prox0[SWIM_MAX_K];
  const uint32_t np0 = shuffle_take(tmp, L, K, pdraw, prox0); /* probe 0's proxies double as the piggyback recipients (T4) */

  /* T3 the probes, one after the other (mapM_ probeNode', Core.hs:240): Ping (Core.hs:246), unlessAck -> IndirectPings
   * (250), unlessAck -> suspect (253). A suspicion raised by an earlier probe of the period is in the store when a later
   * probe draws its proxies; the incarnations were captured when the targets were chosen (Core.hs:239, 243). */
  uint32_t lossd[SWIM_MAX_PROBES], lossi[SWIM_MAX_PROBES * SWIM_MAX_K + 1];
  if (o->loss_ppm) {
    draws_for(o, o->round, self >> 2, P_LOSS0, 4, grp);
    lossd[0] = grp[self & 3];
    if (P > 1) draws_for(o, o->round, self, P_LOSSD, P - 1, lossd + 1);
    draws_for(o, o->round, self, P_LOSS, P * K, lossi);
  }
#define LOST(w) (o->loss_ppm && bounded((w), 1000000u) < o->loss_ppm)
  uint32_t tn0 = ids[targets[0]], tincs[SWIM_MAX_PROBES];
  for (uint32_t j = 0; j < nt; ++j) tincs[j] = inc[targets[j]];
  for (uint32_t j = 0; j < nt; ++j) {
    const uint32_t t = targets[j], tn = ids[t];
    ctr[SWIM_CTR_PINGS]++;
    int acked = o->alive[tn] && !LOST(lossd[j]);
    if (!acked) {
      uint32_t prox[SWIM_MAX_K], np;
      if (j == 0) { np = np0; memcpy(prox, prox0, sizeof prox); }
      else { /* kRandomMembers on the store as it is now */
        uint32_t Lc = 0;
        for (uint32_t s = 0; s < o->cap; ++s)
          if (st[s] == SWIM_ALIVE) tmp[Lc++] = s;
        np = shuffle_take(tmp, Lc, K, pdraw + j * K, prox);
      }
      ctr[SWIM_CTR_DIRECT_FAIL]++;
      ctr[SWIM_CTR_INDIRECT_PINGS] += np;
      for (uint32_t x = 0; x < np; ++x)
        if (o->alive[ids[prox[x]]] && o->alive[tn] && !LOST(lossi[j * K + x])) acked = 1;
    }
    if (!acked) {
      /* suspectNode store $ Suspect (memberIncarnation m) (memberName m) (Core.hs:253) */
      rec_t sus = {tn, tincs[j], 0, SWIM_MSG_SUSPECT, 0, 0}, rb;
      if (apply_record(o, l, sus, 0, 0, &rb, NULL, ctr)) {
        pb_enqueue(o, l, rb, ctr); /* yield . Broadcast (Core.hs:254) */
        ctr[SWIM_CTR_SUSPECT_LOCAL]++;
      }
    }
  }
#undef LOST

  /* T4 [Q5] piggyback: the buffer rides on the messages to the target and the proxies */
  uint32_t cnt = o->pb_cnt[l];
  if (cnt == 0) return;
  /* recipients: the probe targets in order, then probe 0's proxies that are not targets, the first `fanout` of them */
  uint32_t nr = 0;
  (void)tn0;
  for (uint32_t j = 0; j < nt && nr < o->fanout; ++j) to[nr++] = ids[targets[j]];
  for (uint32_t x = 0; x < np0 && nr < o->fanout; ++x) {
    int is_target = 0;
    for (uint32_t j = 0; j < nt; ++j) is_target |= prox0[x] == targets[j];
    if (!is_target) to[nr++] = ids[prox0[x]];
  }
  rec_t *q = o->pb + (size_t)l * o->B, *snap = o->out + (size_t)l * o->B;
  memcpy(snap, q, cnt * sizeof(rec_t));
  o->out_cnt[l] = (uint8_t)cnt;
  ctr[SWIM_CTR_MSGS] += nr;
  ctr[SWIM_CTR_RECS_SENT] += (uint64_t)nr * cnt;
  uint32_t w = 0;
  for (uint32_t x = 0; x < cnt; ++x)
    if (q[x].ttl > 1) { q[w] = q[x]; q[w].ttl--; ++w; }
  memset(q + w, 0, (o->B - w) * sizeof(rec_t));
  o->pb_cnt[l] = (uint8_t)w;
}

/* ------------------------------------------------------------------ phase E: events */
EXPORT int oracle_inject(oracle_t *o, const swim_event_t *ev, size_t n) {
  for (size_t x = 0; x < n; ++x) {
    if (ev[x].round <= o->round || ev[x].node >= o->N || ev[x].kind > SWIM_EV_INJECT) return SWIM_EINVAL;
    if (ev[x].kind == SWIM_EV_INJECT) {
      const swim_message_t *m = &ev[x].msg;
      if (m->kind != SWIM_MSG_SUSPECT && m->kind != SWIM_MSG_ALIVE && m->kind != SWIM_MSG_DEAD) return SWIM_EINVAL;
      if (m->incarnation < 0 || m->incarnation > 0xFFFFFFFFll) return SWIM_ERANGE;
    }
  }
  if (o->n_ev + n > o->cap_ev) {
    o->cap_ev = (o->n_ev + n) * 2;
    o->ev = (swim_event_t *)realloc(o->ev, o->cap_ev * sizeof(swim_event_t));
  }
  memcpy(o->ev + o->n_ev, ev, n * sizeof(swim_event_t));
  o->n_ev += n;
  return SWIM_OK;
}

static rec_t rec_of_msg(const swim_message_t *m) {
  rec_t r = {m->node, (uint32_t)m->incarnation, m->kind == SWIM_MSG_DEAD ? m->dead_from : 0, m->kind, 0, 0};
  return r;
}

static void run_events(oracle_t *o) {
  size_t w = 0;
  for (size_t x = 0; x < o->n_ev; ++x) {
    swim_event_t *e = &o->ev[x];
    if (e->round != o->round) { o->ev[w++] = *e; continue; }
    int local = e->node >= o->first && e->node < o->first + o->n;
    uint32_t l = e->node - o->first;
    switch (e->kind) {
      case SWIM_EV_CRASH:
        if (o->alive[e->node]) { o->alive[e->node] = 0; o->last_crash[e->node] = o->round; }
        break;
      case SWIM_EV_REJOIN:
        if (!o->alive[e->node]) {
          o->alive[e->node] = 1;
          o->last_rejoin[e->node] = o->round;
          if (local) { /* restart: incarnation+1 and announce Alive (BASELINE config C5) */
            o->self_inc[l]++;
            rec_t a = {e->node, o->self_inc[l], 0, SWIM_MSG_ALIVE, 0, 0};
            pb_enqueue(o, l, a, o->ctr);
          }
        }
        break;
      case SWIM_EV_INJECT:
        if (local && o->alive[e->node]) { /* one datagram through `process` (Core.hs:110-117) */
          rec_t rb;
          if (apply_record(o, l, rec_of_msg(&e->msg), 0, 1, &rb, NULL, o->ctr)) {
            pb_enqueue(o, l, rb, o->ctr);
            o->ctr[SWIM_CTR_RECS_APPLIED]++;
          }
        }
        break;
    }
  }
  o->n_ev = w;
}

/* ------------------------------------------------------------------ phase C: seeded churn (BASELINE config C5)
 * At the start of round r, before the events of r: every live process crashes with probability churn_ppm / 1e6; a
 * crashed process whose rejoin round has come restarts with incarnation + 1 and announces Alive (the effects of
 * SWIM_EV_CRASH / SWIM_EV_REJOIN). Draws: words (i & 3) of the Philox blocks (r, i >> 2, P_CHURN, 0) — crash — and
 * (r, i >> 2, P_CHURN, 1) — rejoin delay, uniform in [rejoin_min, rejoin_max]. Every shard runs it for all N nodes. */
static void run_churn(oracle_t *o) {
  const uint32_t ppm = o->cfg.churn_ppm;
  if (!ppm) return;
  const uint32_t span = o->cfg.rejoin_max - o->cfg.rejoin_min + 1;
  for (uint32_t g = 0; 4 * g < o->N; ++g) {
    uint32_t c0[4] = {o->round, g, P_CHURN, 0}, c1[4] = {o->round, g, P_CHURN, 1}, x[4], y[4];
    philox4x32_10(c0, o->key, x);
    int have_y = 0;
    for (uint32_t j = 0; j < 4 && 4 * g + j < o->N; ++j) {
      const uint32_t i = 4 * g + j;
      if (o->alive[i]) {
        if (bounded(x[j], 1000000u) < ppm) {
          if (!have_y) { philox4x32_10(c1, o->key, y); have_y = 1; }
          o->alive[i] = 0;
          o->last_crash[i] = o->round;
          o->back_at[i] = o->round + o->cfg.rejoin_min + bounded(y[j], span);
        }
      } else if (o->back_at[i] == o->round) {
        o->alive[i] = 1;
        o->last_rejoin[i] = o->round;
        o->back_at[i] = 0;
        if (i >= o->first && i < o->first + o->n) { /* restart: incarnation + 1, announce Alive */
          const uint32_t l = i - o->first;
          o->self_inc[l]++;
          rec_t a = {i, o->self_inc[l], 0, SWIM_MSG_ALIVE, 0, 0};
          pb_enqueue(o, l, a, o->ctr);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------ round driver */
static void push_env(env_t **arr, size_t *n, size_t *cap, const env_t *e) {
  if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *arr = (env_t *)realloc(*arr, *cap * sizeof(env_t)); }
  (*arr)[(*n)++] = *e;
}

static uint32_t owner_of(const oracle_t *o, uint32_t node) {
  uint64_t per = ((uint64_t)o->N + o->cfg.world - 1) / o->cfg.world;
  return (uint32_t)(node / per);
}

/* events + tick for every local node; fills the inbox with local envelopes and the outbox
 * with cross-shard ones. */
EXPORT int oracle_round_begin(oracle_t *o) {
  if (!o->view_set) return SWIM_ESTATE;
  o->round++;
  run_churn(o);
  run_events(o);
  o->n_outbox = 0; o->n_inbox = 0;
  uint64_t tot[SWIM_CTR__COUNT] = {0};
#ifdef _OPENMP
#pragma omp parallel
  {
    uint64_t loc[SWIM_CTR__COUNT] = {0};
#pragma omp for schedule(static)
    for (int64_t l = 0; l < (int64_t)o->n; ++l) tick_node(o, (uint32_t)l, loc);
#pragma omp critical
    for (int c = 0; c < SWIM_CTR__COUNT; ++c) tot[c] += loc[c];
  }
#else
  for (uint32_t l = 0; l < o->n; ++l) tick_node(o, l, tot);
#endif
  for (int c = 0; c < SWIM_CTR__COUNT; ++c) o->ctr[c] += tot[c];
  /* route: senders in ascending id order => each receiver's envelopes arrive sorted by src */
  for (uint32_t l = 0; l < o->n; ++l) {
    if (!o->out_cnt[l]) continue;
    const uint32_t *to = o->send_to + (size_t)l * (1 + o->k);
    for (uint32_t f = 0; f <= o->k; ++f) {
      if (to[f] == SWIM_NO_MEMBER) continue;
      env_t e; e.dst = to[f]; e.src = o->first + l; e.cnt = o->out_cnt[l];
      memcpy(e.recs, o->out + (size_t)l * o->B, e.cnt * sizeof(rec_t));
      if (owner_of(o, e.dst) == o->cfg.rank) push_env(&o->inbox, &o->n_inbox, &o->cap_inbox, &e);
      else push_env(&o->outbox, &o->n_outbox, &o->cap_outbox, &e);
    }
  }
  return SWIM_OK;
}

/* cross-shard traffic as flat words: per envelope [dst, src, cnt, pad] + B records (4 words each) */
EXPORT size_t oracle_env_words(const oracle_t *o) { return 4 + 4 * (size_t)o->B; }
EXPORT size_t oracle_outbox_count(const oracle_t *o, uint32_t dst_rank) {
  size_t c = 0;
  for (size_t x = 0; x < o->n_outbox; ++x) c += owner_of(o, o->outbox[x].dst) == dst_rank;
  return c;
}
EXPORT void oracle_outbox_read(const oracle_t *o, uint32_t dst_rank, uint32_t *words) {
  size_t W = oracle_env_words(o), c = 0;
  for (size_t x = 0; x < o->n_outbox; ++x) {
    const env_t *e = &o->outbox[x];
    if (owner_of(o, e->dst) != dst_rank) continue;
    uint32_t *w = words + c++ * W;
    memset(w, 0, W * 4);
    w[0] = e->dst; w[1] = e->src; w[2] = e->cnt;
    memcpy(w + 4, e->recs, e->cnt * sizeof(rec_t));
  }
}
EXPORT void oracle_inbox_add(oracle_t *o, const uint32_t *words, size_t count) {
  size_t W = oracle_env_words(o);
  for (size_t x = 0; x < count; ++x) {
    const uint32_t *w = words + x * W;
    env_t e; e.dst = w[0]; e.src = w[1]; e.cnt = w[2];
    memcpy(e.recs, w + 4, e.cnt * sizeof(rec_t));
    push_env(&o->inbox, &o->n_inbox, &o->cap_inbox, &e);
  }
}

/* every envelope sent in the last round (local and cross-shard), ascending sender, recipient order as chosen:
 * what the reference's `disseminate` would have put on the wire had it had its piggyback queue */
EXPORT size_t oracle_sent_count(const oracle_t *o) {
  size_t c = 0;
  for (uint32_t l = 0; l < o->n; ++l)
    if (o->out_cnt[l])
      for (uint32_t f = 0; f <= o->k; ++f) c += o->send_to[(size_t)l * (1 + o->k) + f] != SWIM_NO_MEMBER;
  return c;
}
EXPORT void oracle_sent_read(const oracle_t *o, uint32_t *words) {
  size_t W = oracle_env_words(o), c = 0;
  for (uint32_t l = 0; l < o->n; ++l) {
    if (!o->out_cnt[l]) continue;
    for (uint32_t f = 0; f <= o->k; ++f) {
      uint32_t dst = o->send_to[(size_t)l * (1 + o->k) + f];
      if (dst == SWIM_NO_MEMBER) continue;
      uint32_t *w = words + c++ * W;
      memset(w, 0, W * 4);
      w[0] = dst; w[1] = o->first + l; w[2] = o->out_cnt[l];
      memcpy(w + 4, o->out + (size_t)l * o->B, o->out_cnt[l] * sizeof(rec_t));
    }
  }
}

static int env_cmp(const void *a, const void *b) {
  const env_t *x = (const env_t *)a, *y = (const env_t *)b;
  if (x->dst != y->dst) return x->dst < y->dst ? -1 : 1;
  if (x->src != y->src) return x->src < y->src ? -1 : 1;
  return 0;
}

/* phase R: every live receiver consumes its envelopes in ascending sender order, records in
 * buffer order, through `process` (Core.hs:110-117) -> maybeBroadcast (Core.hs:119-121). */
EXPORT int oracle_round_end(oracle_t *o) {
  qsort(o->inbox, o->n_inbox, sizeof(env_t), env_cmp); /* (dst, src) unique per round */
  for (size_t x = 0; x < o->n_inbox; ++x) {
    const env_t *e = &o->inbox[x];
    uint32_t l = e->dst - o->first;
    if (!o->alive[e->dst]) continue; /* datagram to a crashed process is lost */
    o->ctr[SWIM_CTR_MSGS_RECV]++;
    for (uint32_t q = 0; q < e->cnt; ++q) {
      rec_t rb;
      if (apply_record(o, l, e->recs[q], 0, 1, &rb, NULL, o->ctr)) {
        pb_enqueue(o, l, rb, o->ctr);
        o->ctr[SWIM_CTR_RECS_APPLIED]++;
      }
    }
  }
  o->n_inbox = 0;
  return SWIM_OK;
}

EXPORT int oracle_step(oracle_t *o, uint32_t rounds) {
  if (o->cfg.world != 1) return SWIM_ESTATE; /* sharded runs are driven round by round */
  for (uint32_t r = 0; r < rounds; ++r) {
    int rc = oracle_round_begin(o);
    if (rc) return rc;
    oracle_round_end(o);
  }
  return SWIM_OK;
}

EXPORT uint32_t oracle_round(const oracle_t *o) { return o->round; }
EXPORT int oracle_set_round(oracle_t *o, uint32_t round) {
  for (size_t x = 0; x < o->n_ev; ++x)
    if (o->ev[x].round <= round) return SWIM_EINVAL;
  o->round = round;
  return SWIM_OK;
}
EXPORT void oracle_counters(const oracle_t *o, uint64_t *out) { memcpy(out, o->ctr, sizeof o->ctr); }

/* ------------------------------------------------------------------ state access */
EXPORT size_t oracle_array_bytes(const oracle_t *o, int arr) {
  size_t n = o->n, slots = n * o->cap;
  switch (arr) {
    case SWIM_ARR_ALIVE: return o->N;
    case SWIM_ARR_BACK_AT: case SWIM_ARR_LAST_CRASH: case SWIM_ARR_LAST_REJOIN: return (size_t)o->N * 4;
    case SWIM_ARR_SELF_INC: case SWIM_ARR_SEQNO: return n * 4;
    case SWIM_ARR_NBR: case SWIM_ARR_VINC: case SWIM_ARR_VLAST: return slots * 4;
    case SWIM_ARR_VST: return slots;
    case SWIM_ARR_PB: return n * o->B * sizeof(rec_t);
    case SWIM_ARR_PB_CNT: return n;
  }
  return 0;
}

EXPORT int oracle_get_array(const oracle_t *o, int arr, void *buf, size_t bytes) {
  if (bytes != oracle_array_bytes(o, arr) || arr < 0 || arr >= SWIM_ARR__COUNT) return SWIM_EINVAL;
  switch (arr) {
    case SWIM_ARR_ALIVE: memcpy(buf, o->alive, bytes); break;
    case SWIM_ARR_SELF_INC: memcpy(buf, o->self_inc, bytes); break;
    case SWIM_ARR_SEQNO: memcpy(buf, o->seqno, bytes); break;
    case SWIM_ARR_NBR: memcpy(buf, o->nbr, bytes); break;
    case SWIM_ARR_VST:
      for (size_t x = 0; x < bytes; ++x) ((uint8_t *)buf)[x] = (uint8_t)(o->state[x] | (o->timer[x] << 2) | (o->conf[x] << 6));
      break;
    case SWIM_ARR_VINC: memcpy(buf, o->vinc, bytes); break;
    case SWIM_ARR_VLAST: memcpy(buf, o->vlast, bytes); break;
    case SWIM_ARR_PB: memcpy(buf, o->pb, bytes); break;
    case SWIM_ARR_PB_CNT: memcpy(buf, o->pb_cnt, bytes); break;
    case SWIM_ARR_BACK_AT: memcpy(buf, o->back_at, bytes); break;
    case SWIM_ARR_LAST_CRASH: memcpy(buf, o->last_crash, bytes); break;
    case SWIM_ARR_LAST_REJOIN: memcpy(buf, o->last_rejoin, bytes); break;
  }
  return SWIM_OK;
}

EXPORT int oracle_set_array(oracle_t *o, int arr, const void *buf, size_t bytes) {
  if (bytes != oracle_array_bytes(o, arr) || arr < 0 || arr >= SWIM_ARR__COUNT) return SWIM_EINVAL;
  switch (arr) {
    case SWIM_ARR_ALIVE: memcpy(o->alive, buf, bytes); break;
    case SWIM_ARR_SELF_INC: memcpy(o->self_inc, buf, bytes); break;
    case SWIM_ARR_SEQNO: memcpy(o->seqno, buf, bytes); break;
    case SWIM_ARR_NBR: return SWIM_EINVAL;
    case SWIM_ARR_VST:
      for (size_t x = 0; x < bytes; ++x) {
        const uint8_t b = ((const uint8_t *)buf)[x];
        o->state[x] = b & 3;
        o->timer[x] = o->cfg.suspicion_max ? (b >> 2) & 15 : b >> 2;
        o->conf[x] = o->cfg.suspicion_max ? b >> 6 : 0;
      }
      break;
    case SWIM_ARR_VINC: memcpy(o->vinc, buf, bytes); break;
    case SWIM_ARR_VLAST: memcpy(o->vlast, buf, bytes); break;
    case SWIM_ARR_PB: memcpy(o->pb, buf, bytes); break;
    case SWIM_ARR_PB_CNT: memcpy(o->pb_cnt, buf, bytes); break;
    case SWIM_ARR_BACK_AT: memcpy(o->back_at, buf, bytes); break;
    case SWIM_ARR_LAST_CRASH: memcpy(o->last_crash, buf, bytes); break;
    case SWIM_ARR_LAST_REJOIN: memcpy(o->last_rejoin, buf, bytes); break;
  }
  return SWIM_OK;
}

/* ------------------------------------------------------------------ digest / convergence */
static uint64_t fmix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
/* digest term of one element group: tag 1 = node scalars, 2 = view slot, 3 = piggyback record */
static uint64_t dg3(uint64_t tag, uint64_t idx, uint64_t w0, uint64_t w1) {
  return fmix64(fmix64(fmix64(idx + (tag << 56)) ^ w0) ^ w1);
}

EXPORT uint64_t oracle_digest(const oracle_t *o) {
  uint64_t d = 0;
  for (uint32_t l = 0; l < o->n; ++l) {
    uint64_t g = o->first + l;
    d += dg3(1, g, (uint64_t)o->self_inc[l] | ((uint64_t)o->seqno[l] << 32), (uint64_t)o->alive[g] | ((uint64_t)o->pb_cnt[l] << 8));
    for (uint32_t s = 0; s < o->cap; ++s) {
      size_t x = (size_t)l * o->cap + s;
      uint64_t st = (uint64_t)(o->state[x] | (o->timer[x] << 2) | (o->conf[x] << 6));
      d += dg3(2, g * o->cap + s, (uint64_t)o->nbr[x] | ((uint64_t)o->vinc[x] << 32), st | ((uint64_t)o->vlast[x] << 8));
    }
    for (uint32_t q = 0; q < o->pb_cnt[l]; ++q) {
      const rec_t *r = &o->pb[(size_t)l * o->B + q];
      d += dg3(3, g * o->B + q, (uint64_t)r->member | ((uint64_t)r->incarnation << 32),
               (uint64_t)r->from | ((uint64_t)r->kind << 32) | ((uint64_t)r->ttl << 40));
    }
  }
  return d;
}

EXPORT uint64_t oracle_mismatches(const oracle_t *o) {
  uint64_t bad = 0;
  for (uint32_t l = 0; l < o->n; ++l) {
    if (!o->alive[o->first + l]) continue;
    for (uint32_t s = 0; s < o->cap; ++s) {
      size_t x = (size_t)l * o->cap + s;
      if (o->state[x] == SWIM_VACANT) continue;
      bad += o->alive[o->nbr[x]] ? o->state[x] != SWIM_ALIVE : o->state[x] != SWIM_DEAD;
    }
  }
  return bad;
}

/* ------------------------------------------------------------------ scalar API (Core.hs) */
static void fill_member(const oracle_t *o, uint32_t l, uint32_t s, swim_member_t *m) {
  size_t x = (size_t)l * o->cap + s;
  memset(m, 0, sizeof *m);
  m->id = o->nbr[x]; m->addr = o->nbr[x]; m->port = (uint16_t)o->cfg.base_port;
  m->liveness = o->state[x]; m->timer = o->timer[x]; m->incarnation = o->vinc[x]; m->last_change = o->vlast[x];
}

/* members (Core.hs:76-77) */
EXPORT int oracle_get_members(const oracle_t *o, uint32_t node, swim_member_t *out, size_t cap, size_t *n_out) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  uint32_t l = node - o->first; size_t c = 0;
  for (uint32_t s = 0; s < o->cap; ++s) {
    if (o->state[(size_t)l * o->cap + s] == SWIM_VACANT) continue;
    if (c == cap) return SWIM_ECAP;
    fill_member(o, l, s, &out[c++]);
  }
  *n_out = c;
  return SWIM_OK;
}

static int member_cmp(const void *a, const void *b) {
  uint32_t x = ((const swim_member_t *)a)->id, y = ((const swim_member_t *)b)->id;
  return x < y ? -1 : x > y;
}

/* swapTVar storeMembers (Spec.hs:101) */
EXPORT int oracle_set_members(oracle_t *o, uint32_t node, const swim_member_t *ms, size_t n) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  if (n > o->cap) return SWIM_ECAP;
  swim_member_t tmp[SWIM_MAX_VIEW];
  memcpy(tmp, ms, n * sizeof *ms);
  qsort(tmp, n, sizeof *tmp, member_cmp); /* Map.fromList orders by key */
  for (size_t x = 0; x < n; ++x) {
    if (tmp[x].id == SWIM_NO_MEMBER || tmp[x].id == node || tmp[x].liveness > SWIM_DEAD ||
        tmp[x].timer > (o->cfg.suspicion_max ? SWIM_MAX_TIMER_LIFEGUARD : SWIM_MAX_TIMER) || (x && tmp[x].id == tmp[x - 1].id))
      return SWIM_EINVAL;
  }
  uint32_t l = node - o->first;
  for (uint32_t s = 0; s < o->cap; ++s) {
    size_t x = (size_t)l * o->cap + s;
    if (s < n) {
      /* the countdown only exists while Suspect; a Suspect member given without one is armed with S */
      o->nbr[x] = tmp[s].id; o->state[x] = tmp[s].liveness;
      o->timer[x] = tmp[s].liveness != SWIM_SUSPECT ? 0 : tmp[s].timer ? tmp[s].timer : (uint8_t)o->s_arm;
      o->conf[x] = 0;
      o->vinc[x] = tmp[s].incarnation; o->vlast[x] = (uint32_t)tmp[s].last_change;
    } else {
      o->nbr[x] = SWIM_NO_MEMBER; o->state[x] = SWIM_VACANT; o->timer[x] = 0; o->conf[x] = 0; o->vinc[x] = 0; o->vlast[x] = 0;
    }
  }
  o->view_set = 1;
  return SWIM_OK;
}

static int member_eq(const swim_member_t *a, const swim_member_t *b) {
  /* derived structural Eq on every field (Types.hs:68), not the name-only Ord (72-73) */
  return a->id == b->id && a->addr == b->addr && a->port == b->port && a->liveness == b->liveness &&
         a->timer == b->timer && a->incarnation == b->incarnation && a->last_change == b->last_change;
}

/* kRandomMembers (Core.hs:69-74) */
EXPORT int oracle_k_random_members(oracle_t *o, uint32_t node, uint32_t n, const swim_member_t *ex, size_t n_ex,
                                   swim_member_t *out, size_t cap, size_t *n_out) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  uint32_t l = node - o->first, cand[SWIM_MAX_VIEW], L = 0;
  for (uint32_t s = 0; s < o->cap; ++s) {
    if (o->state[(size_t)l * o->cap + s] != SWIM_ALIVE) continue; /* isAlive m */
    swim_member_t m; fill_member(o, l, s, &m);
    int excluded = 0;
    for (size_t e = 0; e < n_ex; ++e) excluded |= member_eq(&m, &ex[e]); /* notElem m excludes */
    if (!excluded) cand[L++] = s;
  }
  uint32_t want = n < L ? n : L;
  if (want > cap) return SWIM_ECAP;
  uint32_t draws[SWIM_MAX_VIEW], picks[SWIM_MAX_VIEW];
  uint64_t call = o->scalar_calls++;
  draws_for(o, (uint32_t)call, node, P_SCALAR, want, draws);
  uint32_t got = shuffle_take(cand, L, want, draws, picks);
  for (uint32_t x = 0; x < got; ++x) fill_member(o, l, picks[x], &out[x]);
  *n_out = got;
  return SWIM_OK;
}

/* removeDeadNodes (Core.hs:65-67): Map.filter (not . isDead) */
EXPORT int oracle_remove_dead_nodes(oracle_t *o, uint32_t node) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  uint32_t l = node - o->first, w = 0;
  for (uint32_t s = 0; s < o->cap; ++s) {
    size_t x = (size_t)l * o->cap + s, y = (size_t)l * o->cap + w;
    if (o->state[x] == SWIM_VACANT || o->state[x] == SWIM_DEAD) continue;
    o->nbr[y] = o->nbr[x]; o->state[y] = o->state[x]; o->timer[y] = o->timer[x]; o->conf[y] = o->conf[x];
    o->vinc[y] = o->vinc[x]; o->vlast[y] = o->vlast[x];
    ++w;
  }
  for (; w < o->cap; ++w) {
    size_t y = (size_t)l * o->cap + w;
    o->nbr[y] = SWIM_NO_MEMBER; o->state[y] = SWIM_VACANT; o->timer[y] = 0; o->conf[y] = 0; o->vinc[y] = 0; o->vlast[y] = 0;
  }
  return SWIM_OK;
}

/* nextSeqNo / nextIncarnation: atomicIncr returns the NEW value (Core.hs:42-53) */
EXPORT int oracle_next_seqno(oracle_t *o, uint32_t node, uint32_t *out) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  *out = ++o->seqno[node - o->first];
  return SWIM_OK;
}
EXPORT int oracle_next_incarnation(oracle_t *o, uint32_t node, uint32_t *out) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  *out = ++o->self_inc[node - o->first];
  return SWIM_OK;
}

static void msg_of_rec(const oracle_t *o, const rec_t *r, swim_message_t *m) {
  memset(m, 0, sizeof *m);
  m->kind = r->kind; m->node = r->member; m->incarnation = r->incarnation;
  if (r->kind == SWIM_MSG_DEAD) m->dead_from = r->from;
  if (r->kind == SWIM_MSG_ALIVE) { m->target = r->member; m->port = (uint16_t)o->cfg.base_port; }
}

/* suspectNode / deadNode / aliveNode (Core.hs:189-218); `want` = required constructor */
EXPORT int oracle_apply_message(oracle_t *o, uint32_t node, int want, const swim_message_t *msg,
                                swim_message_t *out, int *has_out) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  if (msg->kind != want) return SWIM_EINVAL; /* reference: `undefined` (Core.hs:191,195,218) */
  if (msg->incarnation < 0 || msg->incarnation > 0xFFFFFFFFll) return SWIM_ERANGE;
  rec_t rb; int err = 0;
  int applied = apply_record(o, node - o->first, rec_of_msg(msg), 1, 1, &rb, &err, o->ctr);
  if (err) return err;
  *has_out = applied;
  if (applied) {
    if (rb.kind == msg->kind && rb.member == msg->node) *out = *msg; /* `Just msg`: the identical message */
    else msg_of_rec(o, &rb, out);                                   /* the Alive refutation */
  }
  return SWIM_OK;
}

/* disseminate, Broadcast branch (Core.hs:131,136-138) */
EXPORT int oracle_broadcast(oracle_t *o, uint32_t node, const swim_message_t *msg) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  if (msg->kind != SWIM_MSG_SUSPECT && msg->kind != SWIM_MSG_ALIVE && msg->kind != SWIM_MSG_DEAD) return SWIM_EINVAL;
  if (msg->incarnation < 0 || msg->incarnation > 0xFFFFFFFFll) return SWIM_ERANGE;
  uint64_t sink[SWIM_CTR__COUNT] = {0};
  pb_enqueue(o, node - o->first, rec_of_msg(msg), sink);
  return SWIM_OK;
}

EXPORT int oracle_get_broadcasts(const oracle_t *o, uint32_t node, swim_message_t *out, size_t cap, size_t *n_out) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  uint32_t l = node - o->first;
  if (o->pb_cnt[l] > cap) return SWIM_ECAP;
  for (uint32_t q = 0; q < o->pb_cnt[l]; ++q) msg_of_rec(o, &o->pb[(size_t)l * o->B + q], &out[q]);
  *n_out = o->pb_cnt[l];
  return SWIM_OK;
}

/* one protocol period's suspicion countdown of one store (phase T1; counters untouched, like every scalar call) */
EXPORT int oracle_tick_timers(oracle_t *o, uint32_t node, uint32_t *n_expired) {
  if (node < o->first || node >= o->first + o->n) return SWIM_EINVAL;
  uint64_t sink[SWIM_CTR__COUNT] = {0};
  uint32_t e = tick_timers(o, node - o->first, sink);
  if (n_expired) *n_expired = e;
  return SWIM_OK;
}

/* the piggyback payload of the next outgoing message (phase T4): the buffer, then one transmission spent on every record */
EXPORT int oracle_take_broadcasts(oracle_t *o, uint32_t node, swim_message_t *out, size_t cap, size_t *n_out) {
  int rc = oracle_get_broadcasts(o, node, out, cap, n_out);
  if (rc) return rc;
  uint32_t l = node - o->first, cnt = o->pb_cnt[l], w = 0;
  rec_t *q = o->pb + (size_t)l * o->B;
  for (uint32_t x = 0; x < cnt; ++x)
    if (q[x].ttl > 1) { q[w] = q[x]; q[w].ttl--; ++w; }
  memset(q + w, 0, (o->B - w) * sizeof(rec_t));
  o->pb_cnt[l] = (uint8_t)w;
  return SWIM_OK;
}

/* process (Core.hs:89-117) */
EXPORT int oracle_handle_message(oracle_t *o, uint32_t node, uint32_t sender_addr, uint16_t sender_port,
                                 const swim_message_t *msg, swim_gossip_t *out, size_t cap, size_t *n_out) {
  if (node < o->first || node >= o->first + o->n || cap < 1) return SWIM_EINVAL;
  *n_out = 0;
  memset(out, 0, sizeof *out);
  switch (msg->kind) {
    case SWIM_MSG_ACK: return SWIM_OK; /* invokeAckHandler; emits nothing (Core.hs:92-94) */
    case SWIM_MSG_PING:
      if (msg->node != node) return SWIM_OK; /* Core.hs:100-101 */
      out->is_direct = 1; out->dest_addr = sender_addr; out->dest_port = sender_port;
      out->msg.kind = SWIM_MSG_ACK; out->msg.seq_no = msg->seq_no; out->msg.payload_len = 0; /* Core.hs:99 */
      *n_out = 1;
      return SWIM_OK;
    case SWIM_MSG_INDIRECT_PING: { /* [Q4] kept: seq := nextIncarnation (Core.hs:105-108, Spec.hs:166-174) */
      uint32_t next = ++o->self_inc[node - o->first];
      out->is_direct = 1; out->dest_addr = msg->target; out->dest_port = msg->port;
      out->msg.kind = SWIM_MSG_PING; out->msg.seq_no = next; out->msg.node = msg->node;
      *n_out = 1;
      return SWIM_OK;
    }
    case SWIM_MSG_SUSPECT: case SWIM_MSG_DEAD: case SWIM_MSG_ALIVE: {
      int has = 0;
      int rc = oracle_apply_message(o, node, msg->kind, msg, &out->msg, &has);
      if (rc) return rc;
      if (has) { out->is_direct = 0; *n_out = 1; } /* maybeBroadcast (Core.hs:119-121) */
      return SWIM_OK;
    }
  }
  return SWIM_EINVAL;
}

EXPORT void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

EXPORT int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
