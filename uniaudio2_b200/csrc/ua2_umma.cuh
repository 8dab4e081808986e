// Interface of the hand-written tcgen05 3xTF32 GEMM (ua2_umma.cu): the schedule shared by the kernel, its launchers and the
// consumers that add the side slots of split tiles.
#pragma once
#include <algorithm>

#include "ua2_common.cuh"

namespace ua2 {

struct UmmaPlan {
  int NT = 0;          // activation rows per tile (MMA N)
  int KB = 0;          // 32-float k-blocks per tile
  int nt_per_mat = 0;  // 128-row weight tiles per weight matrix
  int n_nt = 0;        // weight tiles over all matrices (2 matrices for SwiGLU)
  int n_mt = 0;        // activation-row tiles
  int GM = 1;          // activation-row tiles per raster group
  int n_tiles = 0;
  int grid = 0;
  int full_waves = 0;  // whole tiles dealt round-robin: tiles [0, dp_tiles), CTA c takes c, c + grid, ...
  int dp_tiles = 0;
  int rem_units = 0;   // (tile, k-block) units of the remaining tiles, split evenly: CTA c takes [c Lr, (c + 1) Lr)
  int Lr = 0;
  size_t slot_floats = 0;  // side-slot scratch the launch may write: grid * NT * 128
};

int umma_pick_nt(int M);
UmmaPlan umma_plan(int M, int N, int n_mat, int K, bool bf16 = false);
cudaError_t run_umma_tf32x3(const LaunchCtx& lc, const float* X2, const float* W, const float* W2, float* C, int ldc, float* slots, int M, int N,
                            int K, const UmmaPlan& pl);
// Tensor-core attention (ua2_flash.cu): q16 / k16 / v16 (B, H, T, 64) bf16 -> out (B, T, H * 64) fp32, unmasked softmax(q k^T / 8) v
// (out16 != NULL: the result as bf16 instead, the next linear's operand)
// fp32 SIMT attention of ua2_dit.cu (head size 32 / 64 / 128): q (B * T, H * hs), k / v (B, H, T, hs) -> out (B * T, H * hs)
cudaError_t launch_dense_attn_f32(const LaunchCtx& lc, const float* q, const float* kc, const float* vc, float* out, int B, int T, int H, int hs);
// + gate[b, h, i] * tab[h, j - i + T - 1] on the scaled scores (WavLM's gated relative position bias): gate (B, H, T), tab (H, 2 T - 1)
cudaError_t launch_dense_attn_bias_f32(const LaunchCtx& lc, const float* q, const float* kc, const float* vc, float* out, int B, int T, int H, int hs,
                                       const float* gate, const float* tab);
void set_flash_sbuf(int v);
cudaError_t launch_flash_bf16(const LaunchCtx& lc, const void* q16, const void* k16, const void* v16, float* out, void* out16, int B, int T, int H,
                              int hs);
// + gate[b, h, i] * tab[h, j - i + T - 1] on the scaled scores (gate (B, H, T), tab (H, 2 T - 1) fp32; both NULL = no bias)
cudaError_t launch_flash_bf16_bias(const LaunchCtx& lc, const void* q16, const void* k16, const void* v16, float* out, void* out16, int B, int T, int H,
                                   int hs, const float* gate, const float* tab);
cudaError_t run_umma_bf16(const LaunchCtx& lc, const void* X16, const void* W16, float* C, int ldc, float* slots, int M, int N, int K,
                          const UmmaPlan& pl);
cudaError_t run_umma_fixup(const LaunchCtx& lc, float* C, int ldc, const float* slots, int M, int N, int n_mat, const UmmaPlan& pl);
inline bool umma_has_split_tiles(const UmmaPlan& pl) { return pl.rem_units > 0 && (pl.Lr % pl.KB) != 0; }

#ifdef __CUDACC__
// Sum, in CTA order, of the partial tiles that continuation CTAs left in their side slots for element (m, col) of C; `col` counts
// over the concatenated matrices (col / N = matrix).  The CTA holding the tile's first k-block wrote C itself.
__device__ __forceinline__ float umma_side_sum(const UmmaPlan& pl, const float* __restrict__ slots, int m, int col, int N) {
  const int mat = col / N, cn = col - mat * N;
  const int mt = m / pl.NT, nt = mat * pl.nt_per_mat + (cn >> 7);
  const int g = mt / pl.GM;
  const int gmg = pl.n_mt - g * pl.GM < pl.GM ? pl.n_mt - g * pl.GM : pl.GM;
  const int tile = g * pl.GM * pl.n_nt + nt * gmg + (mt - g * pl.GM);  // rasterised tile number (inverse of um_tile_coords)
  if (tile < pl.dp_tiles) return 0.f;
  const int u0 = (tile - pl.dp_tiles) * pl.KB, u1 = u0 + pl.KB - 1;
  const int c_first = u0 / pl.Lr, c_last = u1 / pl.Lr;
  float s = 0.f;
  for (int c = c_first + 1; c <= c_last; ++c) s += slots[((size_t)c * pl.NT + (m - mt * pl.NT)) * 128 + (cn & 127)];
  return s;
}
// the same for 4 consecutive columns (col % 4 == 0: one tile, one slot row)
__device__ __forceinline__ float4 umma_side_sum4(const UmmaPlan& pl, const float* __restrict__ slots, int m, int col, int N) {
  const int mat = col / N, cn = col - mat * N;
  const int mt = m / pl.NT, nt = mat * pl.nt_per_mat + (cn >> 7);
  const int g = mt / pl.GM;
  const int gmg = pl.n_mt - g * pl.GM < pl.GM ? pl.n_mt - g * pl.GM : pl.GM;
  const int tile = g * pl.GM * pl.n_nt + nt * gmg + (mt - g * pl.GM);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tile < pl.dp_tiles) return s;
  const int u0 = (tile - pl.dp_tiles) * pl.KB, u1 = u0 + pl.KB - 1;
  const int c_first = u0 / pl.Lr, c_last = u1 / pl.Lr;
  for (int c = c_first + 1; c <= c_last; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(slots + ((size_t)c * pl.NT + (m - mt * pl.NT)) * 128 + (cn & 127));
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
  }
  return s;
}
#endif

}  // namespace ua2
