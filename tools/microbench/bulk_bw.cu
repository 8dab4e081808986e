// Microbenchmark: HBM read bandwidth through cp.async.bulk rings vs register LDG streaming on B200.
// For every (bytes per bulk copy, ring depth, warps per CTA, CTAs per SM) prints achieved GB/s over a 2 GiB buffer.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}

// per-warp rings: each warp streams a contiguous share of the buffer in copies of `cbytes`
__global__ void warp_ring(const float* __restrict__ src, size_t n_floats, int cbytes, int depth, float* out) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bars[32][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t gw = (size_t)blockIdx.x * nw + warp, tw = (size_t)gridDim.x * nw;
  const size_t cf = cbytes / 4;
  const size_t n_chunks = n_floats / cf;
  const size_t my = (n_chunks - gw + tw - 1) / tw;  // chunks gw, gw+tw, ...
  float* ring = sm + (size_t)warp * depth * cf;
  if (lane == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (size_t t = 0; t < my && t < (size_t)depth; ++t) {
      mbar_expect_tx(&bars[warp][t], cbytes);
      bulk_g2s(ring + t * cf, src + (gw + t * tw) * cf, cbytes, &bars[warp][t]);
    }
  }
  __syncwarp();
  float acc = 0.f;
  for (size_t t = 0; t < my; ++t) {
    const int st = t % depth;
    mbar_wait(&bars[warp][st], (uint32_t)((t / depth) & 1));
    const float* s = ring + (size_t)st * cf;
    for (size_t k = lane * 4; k < cf; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(s + k);
      acc += v.x + v.y + v.z + v.w;
    }
    __syncwarp();
    if (lane == 0 && t + depth < my) {
      mbar_expect_tx(&bars[warp][st], cbytes);
      bulk_g2s(ring + (size_t)st * cf, src + (gw + (t + depth) * tw) * cf, cbytes, &bars[warp][st]);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// CTA ring: one producer thread issues copies of `cbytes`, all warps consume each stage cooperatively
__global__ void cta_ring(const float* __restrict__ src, size_t n_floats, int cbytes, int depth, float* out) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t full[8], empty[8];
  const int tid = threadIdx.x, nw = blockDim.x >> 5, warp = tid >> 5, lane = tid & 31;
  const size_t cf = cbytes / 4;
  const size_t n_chunks = n_floats / cf;
  const size_t gw = blockIdx.x, tw = gridDim.x;
  const size_t my = (n_chunks - gw + tw - 1) / tw;
  if (tid == 0) {
    for (int s = 0; s < depth; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nw - 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  if (warp == nw - 1) {  // producer warp
    if (lane == 0) {
      for (size_t t = 0; t < my; ++t) {
        const int st = t % depth;
        if (t >= (size_t)depth) mbar_wait(&empty[st], (uint32_t)(((t / depth) - 1) & 1));
        mbar_expect_tx(&full[st], cbytes);
        bulk_g2s(sm + (size_t)st * cf, src + (gw + t * tw) * cf, cbytes, &full[st]);
      }
    }
  } else {
    const int cw = nw - 1;
    for (size_t t = 0; t < my; ++t) {
      const int st = t % depth;
      mbar_wait(&full[st], (uint32_t)((t / depth) & 1));
      const float* s = sm + (size_t)st * cf;
      for (size_t k = (warp * 32 + lane) * 4; k < cf; k += cw * 128) {
        const float4 v = *reinterpret_cast<const float4*>(s + k);
        acc += v.x + v.y + v.z + v.w;
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// register streaming: each warp reads contiguous 512 B per instruction, UNROLL loads in flight
template <int UNROLL>
__global__ void ldg_stream(const float* __restrict__ src, size_t n_floats, float* out) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x, tt = (size_t)gridDim.x * blockDim.x;
  const float4* p = reinterpret_cast<const float4*>(src);
  const size_t n4 = n_floats / 4;
  float acc = 0.f;
  for (size_t i = gt; i + (UNROLL - 1) * tt < n4; i += UNROLL * tt) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i + u * tt));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t n = (size_t)512 << 20;  // 2 GiB of floats
  float *d, *out;
  cudaMalloc(&d, n * 4); cudaMalloc(&out, 16);
  cudaMemset(d, 0, n * 4);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(warp_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(cta_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const double gb = n * 4.0 / 1e9;
  printf("SMs %d, buffer %.2f GB\n", sms, gb);
  for (int u : {4, 8, 16}) {
    for (int wps : {8, 16, 32}) {
      float ms = 0;
      auto run = [&](auto k) { ms = time_ms([&] { k<<<sms * (wps / 8), 256>>>(d, n, out); }); };
      if (u == 4) run(ldg_stream<4>); else if (u == 8) run(ldg_stream<8>); else run(ldg_stream<16>);
      printf("ldg unroll=%2d warps/SM=%2d inflight/SM=%4d KB : %7.1f GB/s\n", u, wps, u * wps * 512 / 1024, gb / (ms * 1e-3));
    }
  }
  for (int cb : {1024, 2048, 4096, 8192, 16384}) {
    for (int depth : {2, 4}) {
      for (int nw : {4, 8}) {
        for (int cps : {1, 2}) {
          const size_t smem = (size_t)nw * depth * cb;
          if (smem * cps > 200 * 1024) continue;
          float ms = time_ms([&] { warp_ring<<<sms * cps, nw * 32, smem>>>(d, n, cb, depth, out); });
          printf("warp_ring copy=%5d B depth=%d warps=%d ctas/SM=%d inflight/SM=%4zu KB : %7.1f GB/s\n", cb, depth, nw, cps, smem * cps / 1024, gb / (ms * 1e-3));
        }
      }
    }
  }
  for (int cb : {4096, 8192, 16384, 32768}) {
    for (int depth : {2, 3, 4, 6}) {
      for (int cps : {1, 2}) {
        const size_t smem = (size_t)depth * cb;
        if (smem * cps > 200 * 1024 || depth > 8) continue;
        float ms = time_ms([&] { cta_ring<<<sms * cps, 9 * 32, smem>>>(d, n, cb, depth, out); });
        printf("cta_ring  copy=%5d B depth=%d ctas/SM=%d inflight/SM=%4zu KB : %7.1f GB/s\n", cb, depth, cps, smem * cps / 1024, gb / (ms * 1e-3));
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
