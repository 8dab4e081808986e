// Interface of the hand-written tcgen05 3xTF32 GEMM (ua2_umma.cu): the stream-K plan shared by the kernel, its launchers and the
// consumers that add the side slots of split tiles.
#pragma once
#include "ua2_common.cuh"

namespace ua2 {

struct UmmaPlan {
  int NT = 0;          // activation rows per tile (MMA N)
  int KB = 0;          // 32-float k-blocks per tile
  int nt_per_mat = 0;  // 128-row weight tiles per weight matrix
  int n_nt = 0;        // weight tiles over all matrices (2 matrices for SwiGLU)
  int n_tiles = 0;
  int grid = 0;
  long long L = 0;      // (tile, k-block) units per CTA
  long long total = 0;  // n_tiles * KB
  size_t slot_floats = 0;  // side-slot scratch the launch may write: grid * NT * 128
};

int umma_pick_nt(int M);
UmmaPlan umma_plan(int M, int N, int n_mat, int K);
cudaError_t run_umma_tf32x3(const LaunchCtx& lc, const float* X2, const float* W, const float* W2, float* C, int ldc, float* slots, int M, int N,
                            int K, const UmmaPlan& pl);
cudaError_t run_umma_fixup(const LaunchCtx& lc, float* C, int ldc, const float* slots, int M, int N, int n_mat, const UmmaPlan& pl);

#ifdef __CUDACC__
// Sum, in CTA order, of the partial tiles that continuation CTAs left in their side slots for element (m, col) of C; `col` counts
// over the concatenated matrices (col / N = matrix).  CTA c_first (the one holding the tile's first k-block) wrote C itself.
__device__ __forceinline__ float umma_side_sum(const UmmaPlan& pl, const float* __restrict__ slots, int m, int col, int N) {
  const int mat = col / N, cn = col - mat * N;
  const int mt = m / pl.NT, nt = mat * pl.nt_per_mat + (cn >> 7);
  const long long u0 = ((long long)mt * pl.n_nt + nt) * pl.KB, u1 = u0 + pl.KB - 1;
  const int c_first = (int)(u0 / pl.L), c_last = (int)(u1 / pl.L);
  float s = 0.f;
  for (int c = c_first + 1; c <= c_last; ++c) s += slots[((size_t)c * pl.NT + (m - mt * pl.NT)) * 128 + (cn & 127)];
  return s;
}
#endif

}  // namespace ua2
