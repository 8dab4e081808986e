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

// H = 2: 128 x 128 tile, 8 x 8 register tile per thread (two 4-wide halves per dimension);  H = 1: 64 x 64 tile, 4 x 4 per thread,
// used when the 128-tile grid would leave most SMs idle (e.g. the codec transformer: 2000 rows x 512 outputs = 64 CTAs).
// Every output element is the same k-ordered fp32 FMA chain in both shapes, so the results are bit-identical.
template <int PRO, int EPI, int H>
__global__ void __launch_bounds__(SG_THREADS, 2) sgemm_linear_kernel(const GemvParams p, const float* __restrict__ stats) {
  constexpr int TBM = 64 * H, TBN = 64 * H, R = 4 * H;  // tile and per-thread register tile
  constexpr int LK = 4 * H;                             // consecutive k one thread stages per operand row and k-tile
  __shared__ __align__(16) float As[2][BK][TBM + PADM];
  __shared__ __align__(16) float Bs[2][BK][TBN + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;  // n0 in n' space (pairs adjacent)
  const int K = p.K;
  const int n_units = (EPI == EPI_SWIGLU) ? p.N : (p.N >> 1);
  const int Np = 2 * n_units;

  pdl_launch_dependents();
  pdl_wait();

  // ---- loader state: thread t moves LK consecutive k of row t / (16 / LK) of both operands per k-tile
  const int lrow = tid / (BK / LK), lk0 = (tid % (BK / LK)) * LK;
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
  auto load_a = [&](int k, float (&v)[LK]) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = k + 4 * h;
      if (arow != nullptr && kk < K) {  // K % 8 == 0 (checked by the launcher)
        x0 = *reinterpret_cast<const float4*>(arow + kk);
        if (PRO == PRO_RMSNORM || PRO == PRO_LAYERNORM) {
          const float4 g0 = *reinterpret_cast<const float4*>(p.norm_w + kk);
          if (PRO == PRO_RMSNORM) {  // rs[m] is applied in the epilogue
            x0 = make_float4(x0.x * g0.x, x0.y * g0.y, x0.z * g0.z, x0.w * g0.w);
          } else {
            const float4 c0 = *reinterpret_cast<const float4*>(p.norm_b + kk);
            x0 = make_float4((x0.x - amean) * arstd * g0.x + c0.x, (x0.y - amean) * arstd * g0.y + c0.y,
                             (x0.z - amean) * arstd * g0.z + c0.z, (x0.w - amean) * arstd * g0.w + c0.w);
          }
        }
      }
      v[4 * h] = x0.x;
      v[4 * h + 1] = x0.y;
      v[4 * h + 2] = x0.z;
      v[4 * h + 3] = x0.w;
    }
  };
  auto load_b = [&](int k, float (&v)[LK]) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (brow != nullptr && k + 4 * h < K) x0 = ldg_stream(brow + k + 4 * h);
      v[4 * h] = x0.x;
      v[4 * h + 1] = x0.y;
      v[4 * h + 2] = x0.z;
      v[4 * h + 3] = x0.w;
    }
  };

  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  float ra[LK], rb[LK];
  load_a(lk0, ra);
  load_b(lk0, rb);
#pragma unroll
  for (int i = 0; i < LK; ++i) {
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
      float av[R], bv[R];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[cur][k][64 * h + ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 * h + tx * 4]);
        av[4 * h] = a4.x;
        av[4 * h + 1] = a4.y;
        av[4 * h + 2] = a4.z;
        av[4 * h + 3] = a4.w;
        bv[4 * h] = b4.x;
        bv[4 * h + 1] = b4.y;
        bv[4 * h + 2] = b4.z;
        bv[4 * h + 3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nkt) {
#pragma unroll
      for (int i = 0; i < LK; ++i) {
        As[cur ^ 1][lk0 + i][lrow] = ra[i];
        Bs[cur ^ 1][lk0 + i][lrow] = rb[i];
      }
    }
    __syncthreads();
  }

  // ---- epilogue: thread owns rows {64h + ty*4 + i} and n' columns {64h + tx*4 + j}: pairs (0,1), (2,3) of each half
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + 64 * (i >> 2) + ty * 4 + (i & 3);
    if (m >= p.M) continue;
    const float rs = (PRO == PRO_RMSNORM) ? stats[m] : 1.f;
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
      const int j = 2 * q;
      const int np = n0 + 64 * (j >> 2) + tx * 4 + (j & 3);
      if (np >= Np) continue;
      const float *rA, *rB;
      int nA, nB;
      unit_rows<EPI>(p, np >> 1, rA, rB, nA, nB);
      epilogue<EPI>(p, 0, 1, m, acc[i][j] * rs, acc[i][j + 1] * rs, nA, nB);
    }
  }
}

int g_sms_sg = 0;
int sm_count_sg() {
  if (g_sms_sg == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_sg, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_sg <= 0) g_sms_sg = 148;
  }
  return g_sms_sg;
}

}  // namespace

// returns cudaErrorNotSupported when the (pro, epi) combination has no tiled instance (caller falls back to the skinny path)
cudaError_t launch_sgemm_linear(const LaunchCtx& lc, int pro, int epi, const GemvParams& p, float* stats_ws) {
  if ((p.K & 7) || (p.ldx & 3)) return cudaErrorNotSupported;  // 2 x 128-bit loads per thread along K
  const int n_units = (epi == EPI_SWIGLU) ? p.N : p.N / 2;
  dim3 grid((2 * n_units + BN - 1) / BN, (p.M + BM - 1) / BM);
  const bool small = (int)(grid.x * grid.y) < sm_count_sg();  // few big tiles: quarter them so every SM gets work
  if (small) grid = dim3((2 * n_units + 63) / 64, (p.M + 63) / 64);
  if (pro == PRO_RMSNORM || pro == PRO_LAYERNORM) {
    if (stats_ws == nullptr) return cudaErrorNotSupported;
    cudaError_t e = launch(lc, row_stats_kernel, dim3((p.M * 32 + 255) / 256), dim3(256), 0, p.X, p.ldx, p.M, p.K, p.eps,
                           pro == PRO_LAYERNORM ? 1 : 0, stats_ws);
    if (e != cudaSuccess) return e;
  }
#define UA2_SG(P, E)                                                                                                        \
  if (pro == P && epi == E)                                                                                                 \
    return small ? launch(lc, sgemm_linear_kernel<P, E, 1>, grid, dim3(SG_THREADS), 0, p, (const float*)stats_ws)           \
                 : launch(lc, sgemm_linear_kernel<P, E, 2>, grid, dim3(SG_THREADS), 0, p, (const float*)stats_ws);
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

// =====================================================================================================
// Implicit-GEMM causal Conv1d on the same register-tiled core.
//   C[n, m] = sum_kk W[n, kk] * A(m, kk),  m = (b, t_out),  kk = ci * Ktaps + tap   (torch weight layout (Cout, Cin, Ktaps))
//   A(m, kk) = f(x[b, ci, t_out * stride + tap * dilation - pad_left])   f = identity | ELU, out of range -> 0 | edge value
// Replaces StreamingConv1d.forward + the pre-activation ELU + the resblock skip add (llm_modules/conv.py:232-254,
// seanet.py:52-66, :92-94; Mimi twin).  Lanes run along m (time), so activation loads and the (B, Cout, T) stores are
// coalesced; the 128-position x BN-channel tile reuses each weight 128x and each activation BN x.
// =====================================================================================================
namespace ua2 {
namespace {

struct ConvGemmParams {
  const float* x;     // (B, Cin, T_in)
  const float* w;     // (Cout, Cin, Ktaps) torch layout, or (Cin, Cout, Ktaps) for the transposed-conv mode
  const float* bias;  // (Cout) or null
  const float* res;   // (B, Cout, T_out) or null
  float* y;           // (B, Cout, T_out)
  int B, Cin, Cout, T_in, T_out, Ktaps, stride, dilation, pad_left;
  int pre_elu, replicate;
  int out_stride;  // > 1: transposed-conv mode, blockIdx.z = output phase: weights w + z*Cout*Cin*Ktaps, store at t*out_stride + z
  int out_offset;  // transposed conv: samples cropped on the left of the full output (non-causal padding)
  int T_store;     // time length of y
  const float* prelu;  // nullable: single-slope nn.PReLU applied after bias, before the residual add
};

__device__ __forceinline__ float elu1g(float x) { return x > 0.f ? x : expm1f(x); }

template <int BNC>  // output channels per CTA: 128 / 64 / 32
__global__ void __launch_bounds__(SG_THREADS, 2) sgemm_conv_kernel(const ConvGemmParams p) {
  constexpr int TN = BNC / 16;  // channels per thread
  __shared__ __align__(16) float Xs[2][BK][BM + PADM];
  __shared__ __align__(16) float Ws[2][BK][BNC + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.B * p.T_out;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BNC;
  const int KT = p.Cin * p.Ktaps;
  const int phase = blockIdx.z;
  const float* wbase = p.w + (size_t)phase * p.Cout * KT;
  const int T_store = p.T_store;
  pdl_launch_dependents();
  pdl_wait();

  // ---- activation loader: thread -> position m0 + (tid % 128), 8 consecutive kk starting at (tid / 128) * 8
  const int lm = tid & 127, lkx = (tid >> 7) * 8;
  const float* xb = nullptr;
  int tin0 = 0;
  {
    const long long m = m0 + lm;
    if (m < M) {
      const int b = (int)(m / p.T_out), t = (int)(m - (long long)b * p.T_out);
      xb = p.x + (size_t)b * p.Cin * p.T_in;
      tin0 = t * p.stride - p.pad_left;
    }
  }
  auto load_x = [&](int kk0, float (&v)[8]) {
    int ci = kk0 / p.Ktaps, tap = kk0 - ci * p.Ktaps;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float val = 0.f;
      if (xb != nullptr && kk0 + i < KT) {
        int pos = tin0 + tap * p.dilation;
        bool inside = pos >= 0 && pos < p.T_in;
        if (!inside && p.replicate) {
          pos = pos < 0 ? 0 : p.T_in - 1;
          inside = true;
        }
        if (inside) {
          val = xb[(size_t)ci * p.T_in + pos];
          if (p.pre_elu) val = elu1g(val);
        }
      }
      v[i] = val;
      if (++tap == p.Ktaps) {
        tap = 0;
        ++ci;
      }
    }
  };
  // ---- weight loader: BNC x 16 elements per k-tile
  constexpr int W_PER_THREAD = BNC * BK / SG_THREADS;  // 8 / 4 / 2
  auto load_w = [&](int kk0, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < W_PER_THREAD; ++i) {
      const int e = tid + SG_THREADS * i;
      const int k = e & 15, n = e >> 4;
      const int kk = kk0 + k;
      v[i] = (n0 + n < p.Cout && kk < KT) ? wbase[(size_t)(n0 + n) * KT + kk] : 0.f;
    }
  };
  auto store_w = [&](int buf, const float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < W_PER_THREAD; ++i) {
      const int e = tid + SG_THREADS * i;
      Ws[buf][e & 15][e >> 4] = v[i];
    }
  };

  float acc[TN][8];
#pragma unroll
  for (int i = 0; i < TN; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float rx[8], rw[8];
  load_x(lkx, rx);
  load_w(0, rw);
#pragma unroll
  for (int i = 0; i < 8; ++i) Xs[0][lkx + i][lm] = rx[i];
  store_w(0, rw);
  __syncthreads();
  const int nkt = (KT + BK - 1) / BK;
  for (int kt = 0; kt < nkt; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nkt) {
      load_x((kt + 1) * BK + lkx, rx);
      load_w((kt + 1) * BK, rw);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(&Xs[cur][k][tx * 4]);
      const float4 x1 = *reinterpret_cast<const float4*>(&Xs[cur][k][64 + tx * 4]);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      float wv[TN];
      if (TN == 8) {
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[cur][k][ty * 4]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + ty * 4]);
        wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
        wv[TN > 4 ? 4 : 0] = w1.x; wv[TN > 5 ? 5 : 0] = w1.y; wv[TN > 6 ? 6 : 0] = w1.z; wv[TN > 7 ? 7 : 0] = w1.w;
      } else if (TN == 4) {
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[cur][k][ty * 4]);
        wv[0] = w0.x; wv[1] = w0.y; wv[TN > 2 ? 2 : 0] = w0.z; wv[TN > 3 ? 3 : 0] = w0.w;
      } else {
        const float2 w0 = *reinterpret_cast<const float2*>(&Ws[cur][k][ty * 2]);
        wv[0] = w0.x; wv[1] = w0.y;
      }
#pragma unroll
      for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
    }
    if (kt + 1 < nkt) {
#pragma unroll
      for (int i = 0; i < 8; ++i) Xs[cur ^ 1][lkx + i][lm] = rx[i];
      store_w(cur ^ 1, rw);
    }
    __syncthreads();
  }

  // ---- epilogue: thread owns channels (TN == 8: ty*4+i and 64+ty*4+i; else ty*TN+i) and positions {tx*4+j, 64+tx*4+j}
#pragma unroll
  for (int i = 0; i < TN; ++i) {
    int nl;
    if (TN == 8) nl = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    else nl = ty * TN + i;
    const int n = n0 + nl;
    if (n >= p.Cout) continue;
    const float bv = p.bias ? p.bias[n] : 0.f;
    const float slope = p.prelu ? p.prelu[0] : 1.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long mb = m0 + h * 64 + tx * 4;
      if (mb >= M) continue;
      const int b = (int)(mb / p.T_out), t = (int)(mb - (long long)b * p.T_out);
      const size_t o = ((size_t)b * p.Cout + n) * T_store + t;
      if (p.out_stride == 1 && p.out_offset == 0 && t + 3 < p.T_out && mb + 3 < M && ((o & 3) == 0)) {
        float4 v = make_float4(acc[i][h * 4 + 0] + bv, acc[i][h * 4 + 1] + bv, acc[i][h * 4 + 2] + bv, acc[i][h * 4 + 3] + bv);
        if (p.prelu) v = make_float4(v.x > 0.f ? v.x : slope * v.x, v.y > 0.f ? v.y : slope * v.y, v.z > 0.f ? v.z : slope * v.z,
                                     v.w > 0.f ? v.w : slope * v.w);
        if (p.res) {
          const float4 r = *reinterpret_cast<const float4*>(p.res + o);
          v = make_float4(r.x + v.x, r.y + v.y, r.z + v.z, r.w + v.w);
        }
        *reinterpret_cast<float4*>(p.y + o) = v;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long m = mb + j;
          if (m >= M) break;
          const int b2 = (int)(m / p.T_out), t2 = (int)(m - (long long)b2 * p.T_out);
          const long long ts = (long long)t2 * p.out_stride + phase - p.out_offset;
          if (ts < 0 || ts >= T_store) continue;
          const size_t o2 = ((size_t)b2 * p.Cout + n) * T_store + (size_t)ts;
          float v = acc[i][h * 4 + j] + bv;
          if (p.prelu) v = v > 0.f ? v : slope * v;
          if (p.res) v = p.res[o2] + v;
          p.y[o2] = v;
        }
      }
    }
  }
}

}  // namespace

namespace {
int g_conv_pointwise = 1;  // option "conv_pointwise": measurement switch for the k = 1 streaming kernel below
}
void set_conv_pointwise(int v) { g_conv_pointwise = v ? 1 : 0; }
int get_conv_pointwise() { return g_conv_pointwise; }

// ---------------------------------------------------------------------------------------------------------------------------
// The two thin ends of the SEANet stacks at the full 24 kHz rate (modules/seanet.py:147-150, :365-371): conv k7 1 -> 64 (encoder
// input) and ELU + conv k7 64 -> 1 (decoder output).  One operand is a single channel, so both are pure streaming kernels (HBM
// bound: 4 B x 64 channels per sample on the wide side); as implicit GEMMs they wasted 31 of 32 output columns (4.0 ms for the
// decoder output at batch 16 x 10 s).  Causal zero padding, stride 1.
constexpr int THIN_T = 1024;     // outputs per CTA (4 per thread)
constexpr int THIN_MAX_K = 16;

// Cout = 1:  y[b, 0, t] = bias + sum_{ci, k} w[0, ci, k] f(x[b, ci, t + k - pad])
__global__ void __launch_bounds__(256) conv1d_cout1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ y, int Cin, int T, int K, int pad_left, int pre_elu) {
  constexpr int CH = 8;  // channels per shared-memory pass
  __shared__ __align__(16) float tile[CH][THIN_T + THIN_MAX_K];
  __shared__ float ws[CH][THIN_MAX_K];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * THIN_T;
  const int span = THIN_T + K - 1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < Cin; c0 += CH) {
    __syncthreads();
    for (int i = tid; i < CH * span; i += 256) {
      const int c = i / span, j = i - c * span;
      const int t = t0 + j - pad_left;
      float v = 0.f;
      if (c0 + c < Cin && t >= 0 && t < T) {
        v = x[((size_t)b * Cin + c0 + c) * T + t];
        if (pre_elu) v = v > 0.f ? v : expm1f(v);
      }
      tile[c][j] = v;
    }
    for (int i = tid; i < CH * K; i += 256) {
      const int c = i / K, k = i - c * K;
      ws[c][k] = (c0 + c < Cin) ? w[(size_t)(c0 + c) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float xv[4 + THIN_MAX_K - 1];
#pragma unroll
      for (int j = 0; j < 4 + THIN_MAX_K - 1; ++j) xv[j] = (j < 4 + K - 1) ? tile[c][4 * tid + j] : 0.f;
#pragma unroll
      for (int k = 0; k < THIN_MAX_K; ++k) {
        if (k < K) {
          const float wk = ws[c][k];
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(wk, xv[o + k], acc[o]);
        }
      }
    }
  }
  const float bv = bias ? bias[0] : 0.f;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int t = t0 + 4 * tid + o;
    if (t < T) y[(size_t)b * T + t] = acc[o] + bv;
  }
}

// Cin = 1:  y[b, co, t] = bias[co] + sum_k w[co, 0, k] f(x[b, 0, t + k - pad])
__global__ void __launch_bounds__(256) conv1d_cin1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ y, int Cout, int T, int K, int pad_left, int pre_elu) {
  extern __shared__ float thin_ws[];  // [Cout][K] then bias [Cout]
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * THIN_T + 4 * tid;
  for (int i = tid; i < Cout * K; i += 256) thin_ws[i] = w[i];
  for (int i = tid; i < Cout; i += 256) thin_ws[Cout * K + i] = bias ? bias[i] : 0.f;
  float xv[4 + THIN_MAX_K - 1];
#pragma unroll
  for (int j = 0; j < 4 + THIN_MAX_K - 1; ++j) {
    const int t = t0 + j - pad_left;
    float v = (j < 4 + K - 1 && t >= 0 && t < T) ? x[(size_t)b * T + t] : 0.f;
    if (pre_elu) v = v > 0.f ? v : expm1f(v);
    xv[j] = v;
  }
  __syncthreads();
  if (t0 >= T) return;
  const bool vec = (T & 3) == 0 && t0 + 3 < T;
  for (int co = 0; co < Cout; ++co) {
    float a[4];
    const float bv = thin_ws[Cout * K + co];
#pragma unroll
    for (int o = 0; o < 4; ++o) a[o] = bv;
#pragma unroll
    for (int k = 0; k < THIN_MAX_K; ++k) {
      if (k < K) {
        const float wk = thin_ws[co * K + k];
#pragma unroll
        for (int o = 0; o < 4; ++o) a[o] = fmaf(wk, xv[o + k], a[o]);
      }
    }
    float* dst = y + ((size_t)b * Cout + co) * T + t0;
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (t0 + o < T) dst[o] = a[o];
    }
  }
}

// Pointwise convolution of a narrow SEANet residual block (k = 1, Cin in {32, 64, 96, 128}: modules/seanet.py:53-58, the second conv of the
// block, 32 -> 64 at 24 kHz, 64 -> 128 at 6 kHz, 128 -> 256 at 1.2 kHz) with the block's skip add:  y[b, co, t] = bias[co] + sum_ci w[co, ci] f(x[b, ci, t]) + res.
// 16-32 FLOP per byte of activations: a streaming kernel.  Thread = one position (all loads and stores coalesced along t), 64 output
// channels in registers, the (Cin x 64) weight slice in shared memory read as broadcast float4.  As an implicit GEMM with K = 32 the
// tiled core spent its time on tile overhead (2.98 ms for 32 -> 64 at batch 16 x 10 s; 2.46 GB of traffic = 0.38 ms at the HBM peak).
template <int PW_CO>  // output channels per CTA: 64, 48 (the 48 / 96-channel stages of ScalarModel) or 32
__global__ void __launch_bounds__(256, 2) conv1d_pointwise_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                               const float* __restrict__ res, float* __restrict__ y, int Cin, int Cout, int T,
                                                               int pre_elu, const float* __restrict__ prelu) {
  extern __shared__ __align__(16) float pw_ws[];  // [Cin][PW_CO]
  pdl_launch_dependents();
  const int tid = threadIdx.x, b = blockIdx.z, co0 = blockIdx.y * PW_CO, t = blockIdx.x * 256 + tid;
  for (int i = tid; i < Cin * PW_CO; i += 256) {  // weights do not depend on the previous kernel: staged before the wait
    const int co = i / Cin, ci = i - co * Cin;
    pw_ws[ci * PW_CO + co] = w[(size_t)(co0 + co) * Cin + ci];
  }
  pdl_wait();
  __syncthreads();
  if (t >= T) return;
  const float slope = prelu ? __ldg(prelu) : 1.f;  // scalar nn.PReLU after the bias, before the skip add (ScalarModel: activation2(conv2(.)) + x, scalar24k.py:139-150)
  float acc[PW_CO];
#pragma unroll
  for (int c = 0; c < PW_CO; ++c) acc[c] = bias ? bias[co0 + c] : 0.f;
  const float* xp = x + (size_t)b * Cin * T + t;
  for (int c0 = 0; c0 < Cin; c0 += 16) {  // Cin % 32 == 0 (launcher): 16 independent loads in flight per thread (2 KB per warp, 32 KB per SM)
    float xv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) xv[i] = __ldg(xp + (size_t)(c0 + i) * T);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v = xv[i];
      if (pre_elu) v = v > 0.f ? v : expm1f(v);
      const float4* wr = reinterpret_cast<const float4*>(pw_ws + (c0 + i) * PW_CO);
#pragma unroll
      for (int c4 = 0; c4 < PW_CO / 4; ++c4) {
        const float4 w4 = wr[c4];
        acc[4 * c4 + 0] = fmaf(w4.x, v, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(w4.y, v, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(w4.z, v, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(w4.w, v, acc[4 * c4 + 3]);
      }
    }
  }
  const size_t o = ((size_t)b * Cout + co0) * T + t;
#pragma unroll
  for (int c = 0; c < PW_CO; ++c) {
    float r = acc[c];
    if (r < 0.f) r *= slope;
    if (res) r += res[o + (size_t)c * T];
    y[o + (size_t)c * T] = r;
  }
}

cudaError_t launch_conv1d_gemm(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res,
                               float* y, int B, int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation,
                               int pad_left, int pre_elu, int replicate, const float* prelu) {
  if (prelu == nullptr && res == nullptr && stride == 1 && dilation == 1 && !replicate && Ktaps <= THIN_MAX_K && T_out == T_in && B <= 65535) {
    const dim3 grid((T_out + THIN_T - 1) / THIN_T, B);
    if (Cout == 1 && Cin > 1) return launch(lc, conv1d_cout1_kernel, grid, dim3(256), 0, x, w_torch, bias, y, Cin, T_in, Ktaps, pad_left, pre_elu);
    if (Cin == 1 && Cout > 1 && (size_t)Cout * (Ktaps + 1) * 4 <= 40 * 1024)
      return launch(lc, conv1d_cin1_kernel, grid, dim3(256), (size_t)Cout * (Ktaps + 1) * 4, x, w_torch, bias, y, Cout, T_in, Ktaps, pad_left, pre_elu);
  }
  if (!(pre_elu && prelu) && Ktaps == 1 && stride == 1 && !replicate && T_out == T_in && pad_left == 0 && Cin <= 128 && (Cin % 16) == 0 && B <= 65535 &&
      get_conv_pointwise()) {
    const dim3 blk(256);
    const unsigned gx = (unsigned)((T_out + 255) / 256);
    if ((Cout % 64) == 0)
      return launch(lc, conv1d_pointwise_kernel<64>, dim3(gx, Cout / 64, B), blk, (size_t)Cin * 64 * 4, x, w_torch, bias, res, y, Cin, Cout, T_in, pre_elu,
                    prelu);
    if ((Cout % 48) == 0)
      return launch(lc, conv1d_pointwise_kernel<48>, dim3(gx, Cout / 48, B), blk, (size_t)Cin * 48 * 4, x, w_torch, bias, res, y, Cin, Cout, T_in, pre_elu,
                    prelu);
    if ((Cout % 32) == 0)
      return launch(lc, conv1d_pointwise_kernel<32>, dim3(gx, Cout / 32, B), blk, (size_t)Cin * 32 * 4, x, w_torch, bias, res, y, Cin, Cout, T_in, pre_elu,
                    prelu);
  }
  {  // option "conv_umma" (default 1): implicit GEMM on tcgen05 straight from (B, C, T), ua2_convumma.cu
    const cudaError_t e = launch_conv1d_umma(lc, x, w_torch, bias, res, y, B, Cin, Cout, T_in, T_out, Ktaps, stride, dilation, pad_left, pre_elu,
                                             replicate, prelu);
    if (e != cudaErrorNotSupported) return e;
  }
  if (get_conv_tc()) {  // option "conv_tc" (default 1): wide layers as im2col + tensor-core GEMM, ua2_convtc.cu
    const cudaError_t e = launch_conv1d_tc(lc, x, w_torch, bias, res, y, B, Cin, Cout, T_in, T_out, Ktaps, stride, dilation, pad_left, pre_elu,
                                           replicate, prelu);
    if (e != cudaErrorNotSupported) return e;
  }
  ConvGemmParams p{x, w_torch, bias, res, y, B, Cin, Cout, T_in, T_out, Ktaps, stride, dilation, pad_left, pre_elu, replicate, 1, 0, T_out,
                   prelu};
  const long long M = (long long)B * T_out;
  const unsigned gx = (unsigned)((M + BM - 1) / BM);
  if (Cout > 64) return launch(lc, sgemm_conv_kernel<128>, dim3(gx, (Cout + 127) / 128), dim3(SG_THREADS), 0, p);
  if (Cout > 32) return launch(lc, sgemm_conv_kernel<64>, dim3(gx, 1), dim3(SG_THREADS), 0, p);
  return launch(lc, sgemm_conv_kernel<32>, dim3(gx, 1), dim3(SG_THREADS), 0, p);
}

// Transposed conv (kernel = 2*stride, causal trim) as `stride` phase GEMMs in one launch (grid.z = phase):
//   y[b, n, j*s + ph] = bias[n] + sum_ci ( w[ci, n, ph] * f(x[b, ci, j]) + w[ci, n, ph + s] * f(x[b, ci, j-1]) )
// w_phase: (s, Cout, Cin, 2) repacked from torch's (Cin, Cout, 2s) by repack_convtr_phase_kernel.
cudaError_t launch_convtr1d_gemm(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B,
                                 int Cin, int Cout, int T_in, int stride, int pre_elu, int crop_left, int T_out,
                                 const float* prelu) {
  // as a conv over the input grid: 2 taps, tap 0 -> x[j], tap 1 -> x[j-1]  (dilation -1, no padding, input stride 1);
  // j runs to T_in inclusive when the right tail (x[T_in - 1] * w[ph + s]) survives the crop
  if (prelu == nullptr && Cout <= 64) {  // the narrow 24 kHz layer (128 -> 64): implicit GEMM straight from (B, C, T), ua2_convumma.cu (option "conv_umma"); wider ones measured equal or better as im2col + GEMM
    const cudaError_t e = launch_convtr1d_umma(lc, x, w_phase, bias, y, B, Cin, Cout, T_in, stride, pre_elu, crop_left, T_out);
    if (e != cudaErrorNotSupported) return e;
  }
  if (get_conv_tc() && prelu == nullptr) {  // option "conv_tc" (default 1): all phases as one tensor-core GEMM, ua2_convtc.cu
    const cudaError_t e = launch_convtr1d_tc(lc, x, w_phase, bias, y, B, Cin, Cout, T_in, stride, pre_elu, crop_left, T_out);
    if (e != cudaErrorNotSupported) return e;
  }
  const int Tj = (crop_left + T_out > T_in * stride) ? T_in + 1 : T_in;
  ConvGemmParams p{x, w_phase, bias, nullptr, y, B, Cin, Cout, T_in, Tj, 2, 1, -1, 0, pre_elu, 0, stride, crop_left, T_out, prelu};
  const long long M = (long long)B * Tj;
  const unsigned gx = (unsigned)((M + BM - 1) / BM);
  if (Cout > 64) return launch(lc, sgemm_conv_kernel<128>, dim3(gx, (Cout + 127) / 128, stride), dim3(SG_THREADS), 0, p);
  if (Cout > 32) return launch(lc, sgemm_conv_kernel<64>, dim3(gx, 1, stride), dim3(SG_THREADS), 0, p);
  return launch(lc, sgemm_conv_kernel<32>, dim3(gx, 1, stride), dim3(SG_THREADS), 0, p);
}

}  // namespace ua2
