// tests/emu/include/cuda_runtime.h — TEST INFRASTRUCTURE. A SIMT emulator and a stub of the CUDA runtime API, so that the
// library's own CUDA sources (swim_b200/csrc/*.cu, swim_device.cuh) compile with g++ and run on the CPU:
//
//     g++ -DSWIM_EMU -I tests/emu/include -x c++ swim_b200/csrc/swim_sim.cu ...      (tests/emu/build_emu.py)
//
// This header shadows <cuda_runtime.h>. It exists to run the device code in a container without a GPU (parity against
// the oracle, the cross-GPU barrier logic with several "ranks" in one process, ASan/TSan over the kernels). It is never
// part of the product: swim_b200/_lib.py loads libswim_b200.so only, and libswim_emu.so is built under tests/emu/.
//
// Execution model
//   * a launch runs every CTA of the grid concurrently; a warp is one OS thread, its 32 lanes are fibers (ucontext) that
//     the warp thread runs one after the other until each blocks in a warp collective (ballot, shfl, reduce, syncwarp,
//     syncthreads) or returns; when every live lane has arrived the collective completes and all lanes resume. Lanes of a
//     warp that arrive at DIFFERENT collectives (divergent use of a full-mask intrinsic) abort the process: a bug detector.
//   * __syncthreads = warp collective + a barrier among the CTA's warp threads.
//   * launches are synchronous (the call returns when the grid has finished), so streams and events are trivial.
//     Persistent kernels of several ranks that wait for each other must therefore be launched from different host threads.
//   * memory: cudaMalloc = aligned malloc, every "device" pointer is a host pointer; CUDA IPC handles carry the raw pointer
//     (all ranks live in one process); atomics and fences map to the GCC __atomic builtins.
#pragma once
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

// ------------------------------------------------------------------ language
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct alignas(8) uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace swim_emu {

struct Idx { unsigned x, y, z; };

enum Op : uint32_t { OP_BALLOT = 1, OP_SHFL, OP_SHFL_XOR, OP_ANY, OP_RED_ADD, OP_RED_OR, OP_SYNCWARP, OP_SYNCTHREADS };

// publish `val`, wait until every live lane of the warp has arrived at a collective of kind `op`, return all values;
// *mask = lanes that took part
void exchange(uint32_t op, uint64_t val, uint64_t out[32], uint32_t *mask);
void *shared(const void *tag, size_t bytes);
void launch(unsigned grid, unsigned block, std::function<void()> body);

} // namespace swim_emu

// the built-in index variables: real (thread-local) objects, not macros — `cfg.gridDim` must stay a member access
extern thread_local swim_emu::Idx threadIdx, blockIdx, blockDim, gridDim;

// ------------------------------------------------------------------ warp / CTA collectives
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t v[32]; uint32_t m;
  swim_emu::exchange(swim_emu::OP_BALLOT, pred != 0, v, &m);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if ((m >> i & 1u) && v[i]) r |= 1u << i;
  return r;
}
static inline int __any_sync(unsigned, int pred) {
  uint64_t v[32]; uint32_t m;
  swim_emu::exchange(swim_emu::OP_ANY, pred != 0, v, &m);
  for (int i = 0; i < 32; ++i) if ((m >> i & 1u) && v[i]) return 1;
  return 0;
}
template <typename T>
static inline T __shfl_sync(unsigned, T var, int src) {
  static_assert(sizeof(T) <= 8, "shfl of at most 64 bits");
  uint64_t bits = 0, v[32]; uint32_t m;
  memcpy(&bits, &var, sizeof(T));
  swim_emu::exchange(swim_emu::OP_SHFL, bits, v, &m);
  src &= 31;
  if (!(m >> src & 1u)) return var; // reading an exited lane is undefined on the device; keep it harmless here
  T r;
  memcpy(&r, &v[src], sizeof(T));
  return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T var, int lane_mask) {
  uint64_t bits = 0, v[32]; uint32_t m;
  memcpy(&bits, &var, sizeof(T));
  swim_emu::exchange(swim_emu::OP_SHFL_XOR, bits, v, &m);
  const int src = (int)((threadIdx.x & 31u) ^ (unsigned)lane_mask) & 31;
  if (!(m >> src & 1u)) return var;
  T r;
  memcpy(&r, &v[src], sizeof(T));
  return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned x) {
  uint64_t v[32]; uint32_t m;
  swim_emu::exchange(swim_emu::OP_RED_ADD, x, v, &m);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (m >> i & 1u) r += (unsigned)v[i];
  return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned x) {
  uint64_t v[32]; uint32_t m;
  swim_emu::exchange(swim_emu::OP_RED_OR, x, v, &m);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (m >> i & 1u) r |= (unsigned)v[i];
  return r;
}
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { uint64_t v[32]; uint32_t m; swim_emu::exchange(swim_emu::OP_SYNCWARP, 0, v, &m); }
static inline void __syncthreads() { uint64_t v[32]; uint32_t m; swim_emu::exchange(swim_emu::OP_SYNCTHREADS, 0, v, &m); }

// ------------------------------------------------------------------ scalar intrinsics, atomics, fences
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned n) { return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (n & 31u)); }

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAnd(unsigned *p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned short atomicCAS(unsigned short *p, unsigned short cmp, unsigned short v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { sched_yield(); }
// "cycles": a quarter of a nanosecond-clock tick, so the watchdog budgets of the kernels (written for ~2 GHz) become
// four times longer in wall time — emulated grids are slow and oversubscribed
static inline long long clock64() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ((long long)ts.tv_sec * 1000000000ll + ts.tv_nsec) >> 2;
}

// ------------------------------------------------------------------ runtime API (host side)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInsufficientDriver = 35, cudaErrorNoDevice = 100, cudaErrorNotReady = 600, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
struct EmuStream { int id; };
struct EmuEvent { double t_ms; };
typedef EmuStream *cudaStream_t;
typedef EmuEvent *cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; int major, minor; size_t totalGlobalMem; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttribute { int id; union { int programmaticStreamSerializationAllowed; char pad[64]; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute *attrs; unsigned numAttrs; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof *p);
  snprintf(p->name, sizeof p->name, "swim SIMT emulator");
  const char *s = getenv("SWIM_EMU_SMS");
  p->multiProcessorCount = s ? atoi(s) : 1;
  p->major = 10;
  return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n) {
  n = (n + 255) & ~(size_t)255;
  *p = aligned_alloc(256, n ? n : 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new EmuStream{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline double swim_emu_now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent{0.0}; return cudaSuccess; }
#define cudaEventDisableTiming 2u
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t_ms = swim_emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
template <typename K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) {
  const char *s = getenv("SWIM_EMU_CTAS_PER_SM");
  *n = s ? atoi(s) : 2;
  return cudaSuccess;
}
template <typename K, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, K kernel, Args... args) {
  swim_emu::launch(cfg->gridDim.x, cfg->blockDim.x, [=] { kernel(args...); });
  return cudaSuccess;
}
// all ranks are handles of one process: an IPC handle is the pointer itself
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
