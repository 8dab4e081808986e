// TEST INFRASTRUCTURE ONLY.  A just-enough CUDA execution model for running simple kernels of csrc/ on the CPU: every CUDA
// thread of a block is an OS thread, __syncthreads() is a barrier, blocks run one after another, __shared__ is a function-local
// static.  Serves kernels made of plain loads / stores / arithmetic / shared memory / __syncthreads (no warp shuffles, no
// mbarrier / bulk copies, no tensor cores) - the gather, scatter and fused-epilogue kernels whose index arithmetic is what can
// go wrong.  Built by tests/test_kernels_on_cpu_shim.py with g++ -std=c++20.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
inline std::barrier<>* g_block_barrier = nullptr;
inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }

// ---- warp shuffles: the 32 OS threads of a warp meet at a per-warp barrier around an exchange buffer.  Serves full warps in which
// every lane executes the shuffle (true for the reductions of csrc/); the mask argument is ignored.
struct ShimWarp {
  std::barrier<> bar{32};
  uint32_t slot[32];
};
inline std::vector<std::unique_ptr<ShimWarp>>* g_warps = nullptr;
inline thread_local int t_linear_tid = 0;
template <typename T>
inline T shim_exchange(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  ShimWarp& w = *(*g_warps)[t_linear_tid >> 5];
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  w.slot[t_linear_tid & 31] = bits;
  w.bar.arrive_and_wait();
  bits = w.slot[src_lane & 31];
  w.bar.arrive_and_wait();
  T out;
  std::memcpy(&out, &bits, 4);
  return out;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return shim_exchange(v, (t_linear_tid & 31) ^ lane_mask); }
template <typename T> inline T __ldg(const T* p) { return *p; }  // read-only data path: a plain load here
template <typename T> inline T __shfl_sync(unsigned, T v, int src_lane) { return shim_exchange(v, src_lane); }
inline void __syncwarp() { (*g_warps)[t_linear_tid >> 5]->bar.arrive_and_wait(); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __expf(float x) { return std::exp(x); }
struct __nv_bfloat16 { uint16_t v; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline __nv_bfloat16 shim_f2bf16_rn(float f) {  // round to nearest even (finite inputs)
  uint32_t u = __float_as_uint(f);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return {(uint16_t)(u >> 16)};
}
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return {shim_f2bf16_rn(a), shim_f2bf16_rn(b)}; }
inline __nv_bfloat16 __float2bfloat16_rn(float a) { return shim_f2bf16_rn(a); }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
using std::max;
using std::min;
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801, cudaErrorLaunchFailure = 719 };
typedef void* cudaStream_t;
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* s) { *s = cudaStreamCaptureStatusNone; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t) { std::memset(dst, v, n); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline int g_shim_sm_count = 148;  // tests shrink it so that persistent kernels walk many work items per CTA
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = g_shim_sm_count; return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "ok" : "shim error"; }

namespace shim {
// dynamic shared memory of the running block when the launcher states its size (run_grid_smem): an exactly-sized heap block, so an
// address sanitizer sees overruns, refilled with 0xFF bytes (NaN as float) before every block since its contents are undefined
inline dim3 g_last_grid;  // grid of the most recent launch (tests check which kernel variant a launcher chose)
inline void* g_dyn_smem = nullptr;
inline size_t g_dyn_smem_bytes = 0;

// ---- mbarrier / bulk-copy emulation (csrc/ua2_common.cuh wrappers).  State per barrier word, guarded by one mutex: the shim has
// no asynchronous copy engine - bulk_copy() is a checked memcpy done by the issuing thread, and its complete_tx follows at once.
struct MbarState { uint32_t phase = 0, count = 1, pending = 1; int64_t tx = 0; };
inline std::mutex g_mbar_mu;
inline std::map<const void*, MbarState> g_mbar;
inline void mbar_init(uint64_t* bar, uint32_t arrivals) {
  std::lock_guard<std::mutex> l(g_mbar_mu);
  g_mbar[bar] = MbarState{0, arrivals, arrivals, 0};
}
inline void mbar_update(uint64_t* bar, uint32_t arrive, int64_t tx_delta) {
  std::lock_guard<std::mutex> l(g_mbar_mu);
  auto it = g_mbar.find(bar);
  if (it == g_mbar.end()) { std::fprintf(stderr, "shim: mbarrier used before mbarrier.init\n"); std::abort(); }
  MbarState& b = it->second;
  b.tx += tx_delta;
  if (arrive > b.pending) { std::fprintf(stderr, "shim: more arrivals than the mbarrier expects\n"); std::abort(); }
  b.pending -= arrive;
  if (b.pending == 0 && b.tx == 0) {  // phase completes
    b.phase ^= 1u;
    b.pending = b.count;
  }
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (;;) {
    {
      std::lock_guard<std::mutex> l(g_mbar_mu);
      auto it = g_mbar.find(bar);
      if (it != g_mbar.end() && it->second.phase != (parity & 1u)) return;  // the phase with this parity has completed
    }
    std::this_thread::yield();
  }
}
inline void bulk_copy(void* dst, const void* src, uint32_t bytes) {
  if ((bytes & 15u) || ((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u)) {  // cp.async.bulk: 16-byte size and alignment
    std::fprintf(stderr, "shim: cp.async.bulk needs 16-byte aligned addresses and size (dst %p src %p bytes %u)\n", dst, src, bytes);
    std::abort();
  }
  std::memcpy(dst, src, bytes);
}

// run `kernel(args...)` for every thread of every block of the grid
template <typename... KArgs, typename... Args>
inline cudaError_t run_grid(void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
  const unsigned nthreads = block.x * block.y * block.z;
  g_last_grid = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        if (g_dyn_smem) std::memset(g_dyn_smem, 0xFF, g_dyn_smem_bytes);
        std::barrier<> bar(nthreads);
        g_block_barrier = &bar;
        std::vector<std::unique_ptr<ShimWarp>> warps;
        for (unsigned w = 0; w < (nthreads + 31) / 32; ++w) warps.push_back(std::make_unique<ShimWarp>());
        g_warps = &warps;
        std::vector<std::thread> ts;  // (a persistent worker pool was tried: slower - the cost is the futex traffic of the barriers)
        ts.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          ts.emplace_back([=, &bar]() {
            t_linear_tid = (int)t;
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            blockIdx = dim3(bx, by, bz);
            blockDim = block;
            gridDim = grid;
            kernel(static_cast<KArgs>(args)...);
            bar.arrive_and_drop();  // a thread that has returned no longer takes part in later barriers
          });
        for (auto& th : ts) th.join();
      }
  return cudaSuccess;
}

template <typename... KArgs, typename... Args>
inline cudaError_t run_grid_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  if (smem_bytes > 227 * 1024) return cudaErrorInvalidValue;  // the sm_100a opt-in limit per CTA
  void* buf = nullptr;
  if (smem_bytes && posix_memalign(&buf, 128, smem_bytes) != 0) return 2;
  g_dyn_smem = buf;
  g_dyn_smem_bytes = smem_bytes;
  const cudaError_t e = run_grid(kernel, grid, block, args...);
  g_dyn_smem = nullptr;
  g_dyn_smem_bytes = 0;
  std::free(buf);
  return e;
}
}  // namespace shim
