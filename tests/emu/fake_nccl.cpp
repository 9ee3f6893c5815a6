// tests/emu/fake_nccl.cpp — TEST INFRASTRUCTURE. An in-process stand-in for the nine NCCL entry points that
// swim_b200/csrc/swim_dist.cu binds with dlopen/dlsym (SWIM_NCCL_LIB points at this library in the emulated tests), so the
// STAGED exchange — bucket counts all-gathered, envelopes sent/received per peer, deliver_kernel — runs on the CPU with the
// ranks as threads of one process. Semantics kept: collectives block until every rank of the communicator has called them;
// point-to-point operations issued between ncclGroupStart and ncclGroupEnd complete at ncclGroupEnd; sends are buffered.
#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <deque>

namespace {

struct Group {
  std::mutex m;
  std::condition_variable cv;
  int world = 0, joined = 0, left = 0;
  // all-gather
  std::vector<std::vector<uint8_t>> stage;
  int ag_arrived = 0, ag_done = 0;
  uint64_t ag_gen = 0;
  // buffered point-to-point messages: box[src * world + dst]
  std::vector<std::deque<std::vector<uint8_t>>> box;
};

struct Comm { Group *g; int rank; };

struct PendingRecv { void *dst; size_t bytes; int src; Comm *c; };

std::mutex g_reg_m;
std::map<uint64_t, Group *> g_groups;
std::atomic<uint64_t> g_next_id{0x5157494D00000001ull};
thread_local int tl_group_depth = 0;
thread_local std::vector<PendingRecv> tl_recvs;

size_t dtype_bytes(ncclDataType_t t) {
  switch ((int)t) {
    case 0: case 1: return 1;       // int8 / uint8
    case 2: case 3: return 4;       // int32 / uint32
    case 4: case 5: return 8;       // int64 / uint64
    case 6: return 2;               // half
    case 7: return 4;               // float
    case 8: return 8;               // double
    default: return 2;
  }
}

void finish_recvs() {
  for (PendingRecv &r : tl_recvs) {
    Group *g = r.c->g;
    std::unique_lock<std::mutex> lk(g->m);
    auto &q = g->box[(size_t)r.src * g->world + r.c->rank];
    g->cv.wait(lk, [&] { return !q.empty(); });
    std::vector<uint8_t> msg = std::move(q.front());
    q.pop_front();
    lk.unlock();
    memcpy(r.dst, msg.data(), msg.size() < r.bytes ? msg.size() : r.bytes);
  }
  tl_recvs.clear();
}

} // namespace

#define API extern "C" __attribute__((visibility("default")))

API const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL error"; }

API ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof *id);
  const uint64_t v = g_next_id.fetch_add(1);
  memcpy(id->internal, &v, sizeof v);
  return ncclSuccess;
}

API ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  uint64_t key;
  memcpy(&key, id.internal, sizeof key);
  Group *g;
  {
    std::lock_guard<std::mutex> lk(g_reg_m);
    Group *&slot = g_groups[key];
    if (!slot) {
      slot = new Group();
      slot->world = nranks;
      slot->stage.resize(nranks);
      slot->box.resize((size_t)nranks * nranks);
    }
    g = slot;
  }
  {
    std::unique_lock<std::mutex> lk(g->m);
    ++g->joined;
    g->cv.notify_all();
    g->cv.wait(lk, [&] { return g->joined >= g->world; }); // a communicator exists once every rank has joined
  }
  *comm = reinterpret_cast<ncclComm_t>(new Comm{g, rank});
  return ncclSuccess;
}

API ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  delete c; // the group itself stays registered: a handful of bytes per test
  return ncclSuccess;
}

API ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t comm, cudaStream_t) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  Group *g = c->g;
  const size_t bytes = count * dtype_bytes(t);
  std::unique_lock<std::mutex> lk(g->m);
  g->cv.wait(lk, [&] { return g->ag_done == 0; }); // the previous all-gather has been read by everybody
  g->stage[c->rank].assign((const uint8_t *)send, (const uint8_t *)send + bytes);
  const uint64_t gen = g->ag_gen;
  if (++g->ag_arrived == g->world) { g->ag_arrived = 0; g->ag_done = g->world; ++g->ag_gen; g->cv.notify_all(); }
  else g->cv.wait(lk, [&] { return g->ag_gen != gen; });
  for (int r = 0; r < g->world; ++r) memcpy((uint8_t *)recv + (size_t)r * bytes, g->stage[r].data(), bytes);
  if (--g->ag_done == 0) g->cv.notify_all();
  return ncclSuccess;
}

API ncclResult_t ncclGroupStart() { ++tl_group_depth; return ncclSuccess; }

API ncclResult_t ncclGroupEnd() {
  if (tl_group_depth > 0 && --tl_group_depth == 0) finish_recvs();
  return ncclSuccess;
}

API ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  Group *g = c->g;
  const size_t bytes = count * dtype_bytes(t);
  {
    std::lock_guard<std::mutex> lk(g->m);
    g->box[(size_t)c->rank * g->world + peer].emplace_back((const uint8_t *)buf, (const uint8_t *)buf + bytes);
  }
  g->cv.notify_all();
  return ncclSuccess;
}

API ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t) {
  tl_recvs.push_back(PendingRecv{buf, count * dtype_bytes(t), peer, reinterpret_cast<Comm *>(comm)});
  if (tl_group_depth == 0) finish_recvs();
  return ncclSuccess;
}
