// Fused SEANet residual block (option "resblock_fused", default 0 - written at the end of round 1, NOT yet run on a B200).
//
// Replaces, for the 64-channel blocks that run at the full 24 kHz rate, the two launches of SEANetResnetBlock
// (modules/seanet.py:21-94; dilation 1, true_skip):   y = x + conv_k1( ELU( conv_k3( ELU(x) ) ) )
// by one kernel that keeps the 32-channel intermediate in shared memory.  profiles/r1_kernel_rooflines.md: as two implicit
// GEMMs these layers take 4.5 + 1.6 ms at batch 16 x 10 s for 63 GFLOP (14 % of the fp32 pipe) and move 3 x 1.5 GB; fused,
// the traffic is one read and one write of the 64-channel activation (2 x 0.98 GB, 0.3 ms at the HBM peak) and the bound
// becomes the fp32 FMA pipe (63 GFLOP = 0.85 ms at peak).
//
// CTA = 128 consecutive positions of one clip, 256 threads:
//   stage 0: ELU(x) tile (64 x 130, two causal halo columns on the left; zeros before t = 0) and both weights -> shared memory
//   stage 1: hidden (32 x 128) = b1 + W1 * ELU(x): thread = 2 hidden channels x 8 positions, 48 FMA per 16 shared loads
//   stage 2: y (64 x 128) = x + b2 + W2 * ELU(hidden): thread = 4 channels x 8 positions, 32 FMA per 6 shared loads; the skip
//            operand is re-read from global memory (an L2 hit: the tile was just loaded) and the stores are 128-bit
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

constexpr int RB_C = 64, RB_H = 32, RB_T = 128, RB_XS = RB_T + 4;  // row stride of the ELU(x) tile: 130 used, 132 for alignment

__device__ __forceinline__ float elu1r(float x) { return x > 0.f ? x : expm1f(x); }

// w1: (H, C, 3) torch Conv1d layout; w2: (C, H, 1)
__global__ void __launch_bounds__(256) resblock64_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ y,
                                                         int T) {
  extern __shared__ __align__(16) float rb_smem[];
  float* xe = rb_smem;                       // [C][RB_XS]   ELU(x), column j = position t0 - 2 + j
  float* w1s = xe + RB_C * RB_XS;            // [C][3][H]    h fastest
  float* w2s = w1s + RB_C * 3 * RB_H;        // [H][C]       c fastest
  float* he = w2s + RB_H * RB_C;             // [H][RB_T]    ELU(hidden)
  const int tid = threadIdx.x;
  const int b = blockIdx.y, t0 = blockIdx.x * RB_T;
  const float* xb = x + (size_t)b * RB_C * T;
  pdl_launch_dependents();
  // weights first: they do not depend on the producer kernel
  for (int i = tid; i < RB_H * RB_C * 3; i += 256) {
    const int h = i / (RB_C * 3), r = i - h * RB_C * 3, ci = r / 3, tap = r - ci * 3;
    w1s[(ci * 3 + tap) * RB_H + h] = w1[i];
  }
  for (int i = tid; i < RB_C * RB_H; i += 256) {
    const int c = i / RB_H, h = i - c * RB_H;
    w2s[h * RB_C + c] = w2[i];
  }
  pdl_wait();
  for (int i = tid; i < RB_C * (RB_T + 2); i += 256) {
    const int ci = i / (RB_T + 2), j = i - ci * (RB_T + 2);
    const int t = t0 - 2 + j;
    xe[ci * RB_XS + j] = (t >= 0 && t < T) ? elu1r(xb[(size_t)ci * T + t]) : 0.f;
  }
  __syncthreads();

  // ---- stage 1: thread -> hidden channels {2*th, 2*th+1}, positions 8*tt .. 8*tt+7
  {
    const int tt = tid & 15, th = tid >> 4;
    float acc[2][8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      acc[0][p] = b1[2 * th];
      acc[1][p] = b1[2 * th + 1];
    }
    for (int ci = 0; ci < RB_C; ++ci) {
      float xv[10];
      const float* xr = xe + ci * RB_XS + 8 * tt;
      const float4 a0 = *reinterpret_cast<const float4*>(xr), a1 = *reinterpret_cast<const float4*>(xr + 4);
      const float2 a2 = *reinterpret_cast<const float2*>(xr + 8);
      xv[0] = a0.x; xv[1] = a0.y; xv[2] = a0.z; xv[3] = a0.w; xv[4] = a1.x; xv[5] = a1.y; xv[6] = a1.z; xv[7] = a1.w;
      xv[8] = a2.x; xv[9] = a2.y;
#pragma unroll
      for (int tap = 0; tap < 3; ++tap) {
        const float2 w = *reinterpret_cast<const float2*>(w1s + (ci * 3 + tap) * RB_H + 2 * th);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          acc[0][p] = fmaf(w.x, xv[p + tap], acc[0][p]);  // position 8*tt + p reads columns p + tap (t - 2 + tap)
          acc[1][p] = fmaf(w.y, xv[p + tap], acc[1][p]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float* dst = he + (2 * th + r) * RB_T + 8 * tt;
      *reinterpret_cast<float4*>(dst) = make_float4(elu1r(acc[r][0]), elu1r(acc[r][1]), elu1r(acc[r][2]), elu1r(acc[r][3]));
      *reinterpret_cast<float4*>(dst + 4) = make_float4(elu1r(acc[r][4]), elu1r(acc[r][5]), elu1r(acc[r][6]), elu1r(acc[r][7]));
    }
  }
  __syncthreads();

  // ---- stage 2: thread -> channels 4*tc .. 4*tc+3, positions 8*tt .. 8*tt+7
  {
    const int tt = tid & 15, tc = tid >> 4;
    float acc[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int p = 0; p < 8; ++p) acc[c][p] = b2[4 * tc + c];
    for (int h = 0; h < RB_H; ++h) {
      const float4 h0 = *reinterpret_cast<const float4*>(he + h * RB_T + 8 * tt), h1 = *reinterpret_cast<const float4*>(he + h * RB_T + 8 * tt + 4);
      const float4 w = *reinterpret_cast<const float4*>(w2s + h * RB_C + 4 * tc);
      const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = fmaf(wv[c], hv[p], acc[c][p]);
    }
    const int t = t0 + 8 * tt;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const size_t o = ((size_t)b * RB_C + 4 * tc + c) * T + t;
      if (t + 7 < T && (o & 3) == 0) {
        const float4 r0 = *reinterpret_cast<const float4*>(x + o), r1 = *reinterpret_cast<const float4*>(x + o + 4);
        *reinterpret_cast<float4*>(y + o) = make_float4(r0.x + acc[c][0], r0.y + acc[c][1], r0.z + acc[c][2], r0.w + acc[c][3]);
        *reinterpret_cast<float4*>(y + o + 4) = make_float4(r1.x + acc[c][4], r1.y + acc[c][5], r1.z + acc[c][6], r1.w + acc[c][7]);
      } else {
#pragma unroll
        for (int p = 0; p < 8; ++p)
          if (t + p < T) y[o + p] = x[o + p] + acc[c][p];
      }
    }
  }
}

int g_resblock_fused = 1;  // taken by the codec handle only while conv_umma is off (measured slower than the default pair of launches)

}  // namespace

void set_resblock_fused(int v) { g_resblock_fused = v ? 1 : 0; }
int get_resblock_fused() { return g_resblock_fused; }

// x, y (B, C, T); y must not alias x (positions of a tile read their left halo from x).  cudaErrorNotSupported unless C = 64, H = 32.
cudaError_t launch_resblock_fused(const LaunchCtx& lc, const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  float* y, int B, int C, int H, int T) {
  if (C != RB_C || H != RB_H || b1 == nullptr || b2 == nullptr || x == y || B > 65535) return cudaErrorNotSupported;
  const size_t smem = (size_t)(RB_C * RB_XS + RB_C * 3 * RB_H + RB_H * RB_C + RB_H * RB_T) * sizeof(float);
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(resblock64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return launch(lc, resblock64_kernel, dim3((T + RB_T - 1) / RB_T, B), dim3(256), smem, x, w1, b1, w2, b2, y, T);
}

}  // namespace ua2
