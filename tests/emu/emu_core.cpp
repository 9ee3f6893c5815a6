// tests/emu/emu_core.cpp — the scheduler of the SIMT emulator declared in tests/emu/include/cuda_runtime.h.
// TEST INFRASTRUCTURE (see that header). One OS thread per warp, one ucontext fiber per lane.
#include <cuda_runtime.h>

#include <memory>

thread_local swim_emu::Idx threadIdx, blockIdx, blockDim, gridDim;

namespace swim_emu {

namespace {

constexpr size_t kStackBytes = 256u << 10;

struct CtaBarrier { // generation barrier whose party count shrinks when a warp retires
  std::mutex m;
  std::condition_variable cv;
  unsigned parties, arrived = 0, gen = 0;
  explicit CtaBarrier(unsigned n) : parties(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived >= parties) { arrived = 0; ++gen; cv.notify_all(); return; }
    cv.wait(lk, [&] { return gen != g; });
  }
  void drop() {
    std::unique_lock<std::mutex> lk(m);
    --parties;
    if (parties && arrived >= parties) { arrived = 0; ++gen; cv.notify_all(); }
  }
};

struct Cta {
  unsigned bid;
  CtaBarrier bar;
  std::mutex mu;
  std::map<const void *, void *> shared_mem;
  Cta(unsigned b, unsigned nwarps) : bid(b), bar(nwarps) {}
  ~Cta() { for (auto &kv : shared_mem) free(kv.second); }
};

struct Warp {
  Cta *cta = nullptr;
  unsigned wid = 0;
  ucontext_t sched;
  ucontext_t ctx[32];
  uint32_t live = 0, waiting = 0;
  uint32_t op[32];
  uint64_t pub[32], res[32];
  uint32_t res_mask = 0;
  int cur = -1;
  const std::function<void()> *body = nullptr;
};

thread_local Warp *tl_warp = nullptr;

void lane_entry() {
  Warp *w = tl_warp;
  const int lane = w->cur;
  (*w->body)();
  w->live &= ~(1u << lane); // uc_link returns to the scheduler
}

void run_warp(Cta *cta, unsigned wid, unsigned nlanes, Idx grid, Idx block, const std::function<void()> *body) {
  auto w = std::make_unique<Warp>();
  w->cta = cta; w->wid = wid; w->body = body;
  tl_warp = w.get();
  ::blockIdx = Idx{cta->bid, 0, 0};
  ::blockDim = block;
  ::gridDim = grid;
  char *stacks = (char *)malloc(kStackBytes * nlanes);
  for (unsigned l = 0; l < nlanes; ++l) {
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = stacks + kStackBytes * l;
    w->ctx[l].uc_stack.ss_size = kStackBytes;
    w->ctx[l].uc_link = &w->sched;
    makecontext(&w->ctx[l], lane_entry, 0);
    w->live |= 1u << l;
  }
  while (w->live) {
    for (unsigned l = 0; l < 32; ++l) {
      if (!(w->live >> l & 1u) || (w->waiting >> l & 1u)) continue;
      w->cur = (int)l;
      ::threadIdx = Idx{wid * 32 + l, 0, 0};
      swapcontext(&w->sched, &w->ctx[l]); // back here when the lane blocks in a collective or returns
    }
    if (!w->live) break;
    // every live lane is waiting: they must all be in the same collective
    const unsigned first = (unsigned)__builtin_ctz(w->waiting);
    for (unsigned l = 0; l < 32; ++l)
      if ((w->waiting >> l & 1u) && w->op[l] != w->op[first]) {
        fprintf(stderr, "swim_emu: divergent warp collective: CTA %u warp %u lane %u is in op %u, lane %u in op %u\n", cta->bid,
                wid, first, w->op[first], l, w->op[l]);
        abort();
      }
    if (w->waiting != w->live) { fprintf(stderr, "swim_emu: scheduler invariant broken\n"); abort(); }
    memcpy(w->res, w->pub, sizeof w->res);
    w->res_mask = w->waiting;
    if (w->op[first] == OP_SYNCTHREADS) cta->bar.wait();
    w->waiting = 0;
  }
  cta->bar.drop();
  free(stacks);
  tl_warp = nullptr;
}

} // namespace

void exchange(uint32_t op, uint64_t val, uint64_t out[32], uint32_t *mask) {
  Warp *w = tl_warp;
  if (!w) { fprintf(stderr, "swim_emu: warp collective outside a kernel\n"); abort(); }
  const int lane = w->cur;
  w->op[lane] = op;
  w->pub[lane] = val;
  w->waiting |= 1u << lane;
  swapcontext(&w->ctx[lane], &w->sched);
  memcpy(out, w->res, sizeof w->res);
  *mask = w->res_mask;
}

void *shared(const void *tag, size_t bytes) {
  Cta *c = tl_warp->cta;
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->shared_mem.find(tag);
  if (it != c->shared_mem.end()) return it->second;
  void *p = aligned_alloc(16, (bytes + 15) & ~(size_t)15);
  memset(p, 0, bytes);
  c->shared_mem[tag] = p;
  return p;
}

void launch(unsigned grid, unsigned block, std::function<void()> body) {
  if (grid == 0 || block == 0) return;
  const unsigned nwarps = (block + 31) / 32;
  std::vector<std::unique_ptr<Cta>> ctas;
  for (unsigned b = 0; b < grid; ++b) ctas.emplace_back(new Cta(b, nwarps));
  std::vector<std::thread> threads;
  threads.reserve((size_t)grid * nwarps);
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned wi = 0; wi < nwarps; ++wi) {
      const unsigned lanes = block - wi * 32 < 32 ? block - wi * 32 : 32;
      threads.emplace_back(run_warp, ctas[b].get(), wi, lanes, Idx{grid, 1, 1}, Idx{block, 1, 1}, &body);
    }
  for (auto &t : threads) t.join();
}

} // namespace swim_emu
