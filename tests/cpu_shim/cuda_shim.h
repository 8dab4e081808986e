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
#include <memory>
#include <string>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
inline std::barrier<>* g_block_barrier = nullptr;
inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
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
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "ok" : "shim error"; }

namespace shim {
// run `kernel(args...)` for every thread of every block of the grid
template <typename... KArgs, typename... Args>
inline cudaError_t run_grid(void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nthreads);
        g_block_barrier = &bar;
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          ts.emplace_back([=, &bar]() {
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
}  // namespace shim
