// Tiled fp32 GEMM for the many-row linears: C[m, n'] = sum_k A(m, k) * W(n', k), fp32 FMA, fused prologue / epilogue.
//
// Used when a linear has M >= 128 rows (forward_prefix of long / batched prompts, model_new.py:456-507; the codec's
// Moshi-family transformer, llm_modules/transformer.py:430-588, M = B * frames).  The skinny (GEMV) kernels stream the
// whole weight matrix once per 8 rows; here a 128 x 128 output tile reuses every weight element 128 times, so the op
// leaves the HBM/L2 roofline and becomes fp32-SIMT compute bound (the precision contract is fp32, DESIGN.md section 2 -
// tf32 tensor-core MMA would break bit-exact VQ indices / greedy token ids).
//
// Design: 256 threads, 128 x 128 x 16 tiles, 8 x 8 register tile per thread (4 LDS.128 per 64 FMA), double-buffered
// shared memory with register prefetch of the next k-tile.  Output columns are enumerated as n' = 2 * unit + r so that the
// two rows of a fused-epilogue pair (RoPE rotation pair, SwiGLU fc_1/fc_2, or simply adjacent rows) land in the same
// thread; the epilogues are the ones of the skinny kernels (ua2_gemv_dev.cuh).  Row-wise prologues: RMSNorm is applied as
// x*g in the loader and rsqrt(mean(x^2)+eps) in the epilogue; LayerNorm uses per-row (mean, rstd) from a statistics pass.
#include "ua2_gemv_dev.cuh"
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

using namespace v1dev;

constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;
constexpr int SG_THREADS = 256;

// one warp per row: RMSNorm -> stats[m] = rsqrt(mean(x^2)+eps); LayerNorm -> stats[2m] = mean, stats[2m+1] = rstd
__global__ void row_stats_kernel(const float* __restrict__ x, int ldx, int M, int K, float eps, int layernorm, float* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* r = x + (size_t)warp * ldx;
  float s = 0.f, ss = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(r + k);
    s += (v.x + v.y) + (v.z + v.w);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (!layernorm) {
    if (lane == 0) stats[warp] = rsqrtf(ss / (float)K + eps);
    return;
  }
  const float mean = s / (float)K;
  float sq = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(r + k);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    sq += a * a + b * b + c * c + d * d;
  }
  sq = warp_sum(sq);
  if (lane == 0) {
    stats[2 * warp] = mean;
    stats[2 * warp + 1] = rsqrtf(sq / (float)K + eps);
  }
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(SG_THREADS, 2) sgemm_linear_kernel(const GemvParams p, const float* __restrict__ stats) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;  // n0 in n' space (pairs adjacent)
  const int K = p.K;
  const int n_units = (EPI == EPI_SWIGLU) ? p.N : (p.N >> 1);
  const int Np = 2 * n_units;

  pdl_launch_dependents();
  pdl_wait();

  // ---- loader state: thread t moves 8 consecutive k (two 128-bit loads) of row t/2 of both operands per k-tile
  const int lrow = tid >> 1, lk0 = (tid & 1) * 8;
  const float* arow = nullptr;
  float amean = 0.f, arstd = 1.f;
  {
    const int m = m0 + lrow;
    if (m < p.M) {
      if (PRO == PRO_GATHER) {
        const long long row = (long long)p.gidx[(size_t)m * p.gidx_stride] + p.gidx_offset;
        arow = p.emb + (size_t)row * K;
      } else {
        arow = p.X + (size_t)m * p.ldx;
      }
      if (PRO == PRO_LAYERNORM) {
        amean = stats[2 * m];
        arstd = stats[2 * m + 1];
      }
    }
  }
  const float* brow = nullptr;
  {
    const int np = n0 + lrow;
    if (np < Np) {
      const float *ra, *rb;
      int nA, nB;
      unit_rows<EPI>(p, np >> 1, ra, rb, nA, nB);
      brow = (np & 1) ? rb : ra;
    }
  }
  auto load_a = [&](int k, float (&v)[8]) {
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (arow != nullptr && k < K) {  // K % 8 == 0 (checked by the launcher)
      x0 = *reinterpret_cast<const float4*>(arow + k);
      x1 = *reinterpret_cast<const float4*>(arow + k + 4);
      if (PRO == PRO_RMSNORM || PRO == PRO_LAYERNORM) {
        const float4 g0 = *reinterpret_cast<const float4*>(p.norm_w + k), g1 = *reinterpret_cast<const float4*>(p.norm_w + k + 4);
        if (PRO == PRO_RMSNORM) {  // rs[m] is applied in the epilogue
          x0 = make_float4(x0.x * g0.x, x0.y * g0.y, x0.z * g0.z, x0.w * g0.w);
          x1 = make_float4(x1.x * g1.x, x1.y * g1.y, x1.z * g1.z, x1.w * g1.w);
        } else {
          const float4 c0 = *reinterpret_cast<const float4*>(p.norm_b + k), c1 = *reinterpret_cast<const float4*>(p.norm_b + k + 4);
          x0 = make_float4((x0.x - amean) * arstd * g0.x + c0.x, (x0.y - amean) * arstd * g0.y + c0.y,
                           (x0.z - amean) * arstd * g0.z + c0.z, (x0.w - amean) * arstd * g0.w + c0.w);
          x1 = make_float4((x1.x - amean) * arstd * g1.x + c1.x, (x1.y - amean) * arstd * g1.y + c1.y,
                           (x1.z - amean) * arstd * g1.z + c1.z, (x1.w - amean) * arstd * g1.w + c1.w);
        }
      }
    }
    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
  };
  auto load_b = [&](int k, float (&v)[8]) {
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (brow != nullptr && k < K) {
      x0 = ldg_stream(brow + k);
      x1 = ldg_stream(brow + k + 4);
    }
    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  load_a(lk0, ra);
  load_b(lk0, rb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    As[0][lk0 + i][lrow] = ra[i];
    Bs[0][lk0 + i][lrow] = rb[i];
  }
  __syncthreads();
  const int nkt = (K + BK - 1) / BK;
  for (int kt = 0; kt < nkt; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nkt) {
      load_a((kt + 1) * BK + lk0, ra);
      load_b((kt + 1) * BK + lk0, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nkt) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        As[cur ^ 1][lk0 + i][lrow] = ra[i];
        Bs[cur ^ 1][lk0 + i][lrow] = rb[i];
      }
    }
    __syncthreads();
  }

  // ---- epilogue: thread owns rows {ty*4+i, 64+ty*4+i} and n' columns {tx*4+j, 64+tx*4+j}: pairs (0,1), (2,3) of each half
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    const float rs = (PRO == PRO_RMSNORM) ? stats[m] : 1.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = 2 * q;
      const int np = n0 + (q < 2 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (np >= Np) continue;
      const float *rA, *rB;
      int nA, nB;
      unit_rows<EPI>(p, np >> 1, rA, rB, nA, nB);
      epilogue<EPI>(p, 0, 1, m, acc[i][j] * rs, acc[i][j + 1] * rs, nA, nB);
    }
  }
}

}  // namespace

// returns cudaErrorNotSupported when the (pro, epi) combination has no tiled instance (caller falls back to the skinny path)
cudaError_t launch_sgemm_linear(const LaunchCtx& lc, int pro, int epi, const GemvParams& p, float* stats_ws) {
  if ((p.K & 7) || (p.ldx & 3)) return cudaErrorNotSupported;  // 2 x 128-bit loads per thread along K
  const int n_units = (epi == EPI_SWIGLU) ? p.N : p.N / 2;
  const dim3 grid((2 * n_units + BN - 1) / BN, (p.M + BM - 1) / BM);
  if (pro == PRO_RMSNORM || pro == PRO_LAYERNORM) {
    if (stats_ws == nullptr) return cudaErrorNotSupported;
    cudaError_t e = launch(lc, row_stats_kernel, dim3((p.M * 32 + 255) / 256), dim3(256), 0, p.X, p.ldx, p.M, p.K, p.eps,
                           pro == PRO_LAYERNORM ? 1 : 0, stats_ws);
    if (e != cudaSuccess) return e;
  }
#define UA2_SG(P, E) \
  if (pro == P && epi == E) return launch(lc, sgemm_linear_kernel<P, E>, grid, dim3(SG_THREADS), 0, p, (const float*)stats_ws);
  UA2_SG(PRO_PLAIN, EPI_STORE)
  UA2_SG(PRO_PLAIN, EPI_RESADD)
  UA2_SG(PRO_PLAIN, EPI_SWIGLU)
  UA2_SG(PRO_PLAIN, EPI_QKV)
  UA2_SG(PRO_PLAIN, EPI_SCALE_RESADD)
  UA2_SG(PRO_PLAIN, EPI_QKV_IL)
  UA2_SG(PRO_RMSNORM, EPI_STORE)
  UA2_SG(PRO_RMSNORM, EPI_RESADD)
  UA2_SG(PRO_RMSNORM, EPI_SWIGLU)
  UA2_SG(PRO_RMSNORM, EPI_QKV)
  UA2_SG(PRO_LAYERNORM, EPI_QKV_IL)
  UA2_SG(PRO_LAYERNORM, EPI_GELU)
  UA2_SG(PRO_GATHER, EPI_STORE)
#undef UA2_SG
  return cudaErrorNotSupported;
}

}  // namespace ua2
