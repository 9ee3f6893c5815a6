// swim_dist.cu — cross-shard exchange (placeholder until the multi-GPU path lands).
#include "swim_host.h"

namespace swim {
int dist_exchange(swim_sim *sim) {
  set_error(sim, "multi-shard exchange not connected");
  return SWIM_ESTATE;
}
void dist_teardown(swim_sim *) {}
} // namespace swim

extern "C" int swim_nccl_unique_id(uint8_t *) { return SWIM_ENCCL; }
extern "C" int swim_sim_connect(swim_sim_t *, const uint8_t *) { return SWIM_ENCCL; }
