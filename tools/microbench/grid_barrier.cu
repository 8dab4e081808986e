// Microbenchmark: cost of one grid-wide barrier on a co-resident persistent grid (2 CTAs/SM x 256 threads), the building
// block of a multi-op persistent kernel.  Variants: 0 = atomicAdd + ld.acquire spin (csrc/ua2_chain.cu), 1 = same with
// nanosleep back-off, 2 = red.release + ld.acquire, 3 = 0 plus a dependent 12 KB activation read from L2 after the barrier
// (what a fused linear does next), 4 = cooperative_groups grid.sync().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_barrier grid_barrier.cu && ./grid_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int VAR>
__global__ void __launch_bounds__(256, 2) bar_kernel(unsigned* ctr, int n_iter, const float* x, float* sink) {
  __shared__ float xs[3072];
  unsigned arrivals = 0;
  float acc = 0.f;
  cg::grid_group grid = cg::this_grid();
  for (int it = 0; it < n_iter; ++it) {
    arrivals += gridDim.x;
    if (VAR == 4) {
      grid.sync();
    } else {
      __syncthreads();
      if (threadIdx.x == 0) {
        if (VAR == 2) {
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        } else {
          __threadfence();
          atomicAdd(ctr, 1u);
        }
        while (ld_acquire(ctr) < arrivals) {
          if (VAR == 1) __nanosleep(32);
        }
        if (VAR != 2) __threadfence();
      }
      __syncthreads();
    }
    if (VAR == 3) {
      for (int k = threadIdx.x * 4; k < 3072; k += 1024) *reinterpret_cast<float4*>(xs + k) = *reinterpret_cast<const float4*>(x + k);
      __syncthreads();
      acc += xs[(threadIdx.x * 7 + it) % 3072];
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int VAR>
void run(const char* name, int grid, int n_iter) {
  unsigned* ctr;
  float *x, *sink;
  cudaMalloc(&ctr, 4);
  cudaMalloc(&x, 3072 * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(x, 0, 3072 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(ctr, 0, 4);
    void* args[] = {&ctr, &n_iter, &x, &sink};
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bar_kernel<VAR>, dim3(grid), dim3(256), args, 0, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      printf("%s: launch failed\n", name);
      return;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-58s grid %4d: %.3f us per barrier\n", name, grid, best * 1e3f / n_iter);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int per = 1; per <= 2; ++per) {
    const int g = sms * per;
    run<0>("atomicAdd + ld.acquire spin (thread 0), 2 __syncthreads", g, 2000);
    run<1>("same + nanosleep(32) back-off", g, 2000);
    run<2>("red.release + ld.acquire spin", g, 2000);
    run<3>("variant 0 + dependent 12 KB L2 read after the barrier", g, 2000);
    run<4>("cooperative_groups grid.sync()", g, 2000);
  }
  return 0;
}
