// Host build of the pure integer helpers of swim_b200/csrc/swim_device.cuh (Philox, r-th-set-bit pick, round-robin
// pick): the same source the kernels compile, run on the CPU by tests/test_device_helpers.py. Test infrastructure.
#include "../swim_b200/csrc/swim_device.cuh"

using namespace swim;

template <int W>
static uint32_t pick_w(uint32_t *am, uint32_t r) {
  uint32_t m[W];
  for (int w = 0; w < W; ++w) m[w] = am[w];
  uint32_t s = pick_remove<W>(m, r);
  for (int w = 0; w < W; ++w) am[w] = m[w];
  return s;
}
template <int W>
static uint32_t rr_w(const uint32_t *am, uint32_t word, uint32_t round) {
  uint32_t m[W];
  for (int w = 0; w < W; ++w) m[w] = am[w];
  return rr_pick<W>(m, word, round);
}

extern "C" {
void h_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  uint4 r = philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
uint32_t h_bounded(uint32_t x, uint32_t L) { return bounded(x, L); }
uint32_t h_nth_set(uint32_t m, uint32_t r) { return nth_set(m, r); }
uint32_t h_xor_permute(uint32_t m, uint32_t b) { return xor_permute(m, b); }

uint32_t h_pick_remove(int W, uint32_t *am, uint32_t r) {
  switch (W) { case 1: return pick_w<1>(am, r); case 2: return pick_w<2>(am, r); case 4: return pick_w<4>(am, r); default: return pick_w<8>(am, r); }
}

uint32_t h_rr_pick(int W, const uint32_t *am, uint32_t word, uint32_t round) {
  switch (W) { case 1: return rr_w<1>(am, word, round); case 2: return rr_w<2>(am, word, round); case 4: return rr_w<4>(am, word, round); default: return rr_w<8>(am, word, round); }
}
}
