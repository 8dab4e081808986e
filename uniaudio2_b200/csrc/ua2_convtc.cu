// Wide SEANet convolutions on the tensor cores (option "conv_tc", default 0 - written at the end of round 1, NOT yet run on a
// B200).
//
// profiles/r1_kernel_rooflines.md: the register-tiled fp32 core runs the strided down-convolutions (k = 2 * stride, 64..1024
// channels) at 25-30 TFLOP/s (34-40 % of the fp32 pipe) while the tcgen05 3xTF32 GEMM of the linears sustains ~170
// fp32-equivalent TFLOP/s.  For layers whose reduction length Cin * K is >= 1024 the convolution is therefore run as
//     im2col  : A[(b, t)][ci * K + tap] = f(x[b, ci, t * stride - pad + tap * dilation])      (f = identity | ELU, zero / replicate pad)
//     GEMM    : C = A W^T with W = the torch weight (Cout, Cin, K) read as (Cout, Cin * K) - no repack - on the 3xTF32 path
//     epilogue: y[b, co, t] = C[(b, t)][co] + bias[co] (+ residual), transposed back to the (B, C, T) layout through shared memory
// in chunks of rows that bound the scratch.  Transposed convolutions (kernel = 2 * stride) are ONE GEMM over all output phases
// (K = 2 * Cin, N = stride * Cout) with a scattering epilogue; there the im2col matrix is only twice the input, so they are
// served from Cin >= 128 on (the SIMT phase GEMMs run the 128 -> 64 stage at 13 TFLOP/s).  Same arithmetic class as the transformer linears of the codec (VQ indices stayed
// bit-equal there).  The im2col matrix costs 4 * Cin * K bytes per output position of extra traffic, which is why the narrow
// 24 kHz layers (Cin * K <= 512) stay on the SIMT core.
#include <algorithm>

#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

__device__ __forceinline__ float elu1c(float x) { return x > 0.f ? x : expm1f(x); }

// rows [m0, m0 + rows) of the im2col matrix; consecutive threads walk k (coalesced stores; loads are runs of K taps)
__global__ void conv_im2col_kernel(const float* __restrict__ x, float* __restrict__ A, int Cin, int T_in, int T_out, int Ktaps, int stride,
                                   int dilation, int pad_left, int pre_elu, int replicate, long long m0, int rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int KT = Cin * Ktaps;
  const long long n = (long long)rows * KT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % KT);
    const long long m = m0 + i / KT;
    const int b = (int)(m / T_out), t = (int)(m - (long long)b * T_out);
    const int ci = k / Ktaps, tap = k - ci * Ktaps;
    int pos = t * stride - pad_left + tap * dilation;
    bool inside = pos >= 0 && pos < T_in;
    if (!inside && replicate) {
      pos = pos < 0 ? 0 : T_in - 1;
      inside = true;
    }
    float v = 0.f;
    if (inside) {
      v = x[((size_t)b * Cin + ci) * T_in + pos];
      if (pre_elu) v = elu1c(v);
    }
    A[i] = v;
  }
}

// y[b, co, t] = act(C[r][co] + bias[co]) (+ res[b, co, t]) for rows r of the chunk, act = identity or a scalar nn.PReLU (ScalarModel);
// 32 x 32 tiles through shared memory
__global__ void conv_tc_epilogue_kernel(const float* __restrict__ C, const float* __restrict__ bias, const float* __restrict__ res,
                                        float* __restrict__ y, int Cout, int T_out, long long m0, int rows, const float* __restrict__ prelu) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const float slope = prelu ? prelu[0] : 1.f;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // read: threadIdx.x walks channels
    const int r = r0 + i, co = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && co < Cout) ? C[(size_t)r * Cout + co] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // write: threadIdx.x walks positions
    const int co = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && co < Cout) {
      const long long m = m0 + r;
      const int b = (int)(m / T_out), t = (int)(m - (long long)b * T_out);
      const size_t o = ((size_t)b * Cout + co) * T_out + t;
      float v = tile[threadIdx.x][i] + (bias ? bias[co] : 0.f);
      if (v < 0.f) v *= slope;
      if (res) v += res[o];
      y[o] = v;
    }
  }
}


// ---- transposed convolution (kernel = 2 * stride): every output phase is a 2-tap filter over the input grid (ua2_sgemm.cu,
// launch_convtr1d_gemm), so all `stride` phases are ONE GEMM:  A[(b, j)][ci * 2 + tap] = f(x[b, ci, j - tap]),
// W = w_phase (stride, Cout, Cin, 2) read as (stride * Cout, Cin * 2),  C[(b, j)][ph * Cout + co] -> y[b, co, j * stride + ph - crop]
// 32 rows x 32 input channels per CTA through shared memory: loads walk j (coalesced), stores walk k (coalesced float2)
__global__ void convtr_im2col_kernel(const float* __restrict__ x, float* __restrict__ A, int Cin, int T_in, int Tj, int pre_elu,
                                     long long m0, int rows) {
  __shared__ float t0[32][33], t1[32][33];  // [ci][row]: tap 0 = x[j], tap 1 = x[j - 1]
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int r = r0 + threadIdx.x;
  int b = 0, j = 0;
  if (r < rows) {
    const long long m = m0 + r;
    b = (int)(m / Tj);
    j = (int)(m - (long long)b * Tj);
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ci = c0 + i;
    float a = 0.f, bb = 0.f;
    if (r < rows && ci < Cin) {
      const float* xr = x + ((size_t)b * Cin + ci) * T_in;
      if (j < T_in) a = pre_elu ? elu1c(xr[j]) : xr[j];
      if (j >= 1 && j - 1 < T_in) bb = pre_elu ? elu1c(xr[j - 1]) : xr[j - 1];
    }
    t0[i][threadIdx.x] = a;
    t1[i][threadIdx.x] = bb;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int rr = r0 + i, ci = c0 + threadIdx.x;
    if (rr < rows && ci < Cin)
      *reinterpret_cast<float2*>(A + (size_t)rr * 2 * Cin + 2 * ci) = make_float2(t0[threadIdx.x][i], t1[threadIdx.x][i]);
  }
}

constexpr int TR_J = 8;      // input positions per CTA of the epilogue
constexpr int TR_MAX_S = 8;  // largest stride served (SEANet ratios are <= 8)

// y[b, co, j * s + ph - crop] = C[(b, j)][ph * Cout + co] + bias[co]; reads walk co, writes walk the 8 * s consecutive samples
__global__ void convtr_tc_epilogue_kernel(const float* __restrict__ C, const float* __restrict__ bias, float* __restrict__ y, int Cout, int s,
                                          int Tj, int T_out, int crop_left, long long m0, int rows) {
  __shared__ float tile[TR_J][TR_MAX_S][33];
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * TR_J, c0 = blockIdx.y * 32;
  const int N = s * Cout;
  {
    const int jl = threadIdx.y, co = c0 + threadIdx.x;  // blockDim = (32, TR_J)
    for (int ph = 0; ph < s; ++ph)
      tile[jl][ph][threadIdx.x] = (r0 + jl < rows && co < Cout) ? C[(size_t)(r0 + jl) * N + ph * Cout + co] : 0.f;
  }
  __syncthreads();
  for (int cl = threadIdx.y; cl < 32; cl += blockDim.y) {
    const int co = c0 + cl;
    if (co >= Cout) continue;
    const float bv = bias ? bias[co] : 0.f;
    for (int e = threadIdx.x; e < TR_J * s; e += 32) {
      const int jl = e / s, ph = e - jl * s;
      if (r0 + jl >= rows) continue;
      const long long m = m0 + r0 + jl;
      const int b = (int)(m / Tj), j = (int)(m - (long long)b * Tj);
      const int t = j * s + ph - crop_left;
      if (t >= 0 && t < T_out) y[((size_t)b * Cout + co) * T_out + t] = tile[jl][ph][cl] + bv;
    }
  }
}

struct ConvTcScratch {
  float *a = nullptr, *c = nullptr;
  size_t a_floats = 0, c_floats = 0;
  TcWorkspace tc;
};
ConvTcScratch g_ws;
int g_conv_tc = 1;  // default since round 2 (first hardware run: profiles/r2_first_call.md)

cudaError_t grow(float** p, size_t* have, size_t want) {
  if (want <= *have) return cudaSuccess;
  if (*p) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    cudaFree(*p);
    *p = nullptr;
    *have = 0;
  }
  cudaError_t e = cudaMalloc((void**)p, want * sizeof(float));
  if (e == cudaSuccess) *have = want;
  return e;
}

}  // namespace

void set_conv_tc(int v) { g_conv_tc = v ? 1 : 0; }
int get_conv_tc() { return g_conv_tc; }

// returns cudaErrorNotSupported when the layer is not served here (the caller continues on the SIMT core)
cudaError_t launch_conv1d_tc(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res, float* y,
                             int B, int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation, int pad_left, int pre_elu,
                             int replicate, const float* prelu) {
  const int KT = Cin * Ktaps;
  const long long M = (long long)B * T_out;
  // reduction length: >= 1024, or a pointwise convolution of >= 256 channels (the "im2col" is then just the transpose to rows)
  const bool long_k = KT >= 1024 || (Ktaps == 1 && KT >= 256 && M >= 1024);
  if (!tc_gemm_available() || !get_tc_gemm() || !long_k || (KT & 3) || (Cout & 3) || M < 128 || (pre_elu && prelu)) return cudaErrorNotSupported;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(lc.stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return cudaErrorNotSupported;  // the scratch may have to grow
  // rows per chunk: im2col chunk <= 64 Mi floats (256 MB; its tf32 split is 3x that)
  long long R = std::max<long long>(128, ((64LL << 20) / KT) / 128 * 128);
  R = std::min(R, M);
  cudaError_t e;
  if ((e = grow(&g_ws.a, &g_ws.a_floats, (size_t)R * KT)) != cudaSuccess) return e;
  if ((e = grow(&g_ws.c, &g_ws.c_floats, (size_t)R * Cout)) != cudaSuccess) return e;
  if ((e = grow(&g_ws.tc.a, &g_ws.tc.a_floats, (size_t)R * 2 * KT)) != cudaSuccess) return e;
  if ((e = grow(&g_ws.tc.slots, &g_ws.tc.slots_floats, tc_slots_max_floats())) != cudaSuccess) return e;
  if ((e = grow(&g_ws.tc.c, &g_ws.tc.c_floats, (size_t)R * Cout)) != cudaSuccess) return e;
  for (long long m0 = 0; m0 < M; m0 += R) {
    const int rows = (int)std::min<long long>(R, M - m0);
    const long long n = (long long)rows * KT;
    const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 64);
    if ((e = launch(lc, conv_im2col_kernel, dim3(grid), dim3(256), 0, x, g_ws.a, Cin, T_in, T_out, Ktaps, stride, dilation, pad_left, pre_elu,
                    replicate, m0, rows)) != cudaSuccess)
      return e;
    GemvParams p;
    p.W = w_torch;
    p.N = Cout;
    p.K = KT;
    p.M = rows;
    p.X = g_ws.a;
    p.ldx = KT;
    p.Y = g_ws.c;
    p.ldy = Cout;
    p.tc = &g_ws.tc;
    const float* raw = nullptr;
    p.raw_out = &raw;
    if (rows >= get_tc_min_rows()) {
      e = launch_tc_linear(lc, PRO_PLAIN, EPI_STORE, p);
    } else {
      e = cudaErrorNotSupported;
    }
    if (e == cudaErrorNotSupported) {  // tail chunk too small for the tensor-core path: skinny fp32 kernels into g_ws.c
      p.raw_out = nullptr;
      p.tc = nullptr;
      e = launch_gemv(lc, PRO_PLAIN, EPI_STORE, p);
    }
    if (e != cudaSuccess) return e;
    const float* src = raw ? raw : g_ws.c;
    if ((e = launch(lc, conv_tc_epilogue_kernel, dim3((rows + 31) / 32, (Cout + 31) / 32), dim3(32, 8), 0, src, bias, res, y, Cout, T_out, m0,
                    rows, prelu)) != cudaSuccess)
      return e;
  }
  return cudaSuccess;
}


// Transposed conv of launch_convtr1d_gemm (kernel = 2 * stride) on the tensor cores; cudaErrorNotSupported -> SIMT phase GEMMs
cudaError_t launch_convtr1d_tc(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin,
                               int Cout, int T_in, int stride, int pre_elu, int crop_left, int T_out) {
  const int KT = 2 * Cin, N = stride * Cout;
  const int Tj = (crop_left + T_out > T_in * stride) ? T_in + 1 : T_in;  // same input grid as the SIMT launcher
  const long long M = (long long)B * Tj;
  if (!tc_gemm_available() || !get_tc_gemm() || KT < 256 || (KT & 3) || (N & 3) || stride > TR_MAX_S || M < 128) return cudaErrorNotSupported;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(lc.stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return cudaErrorNotSupported;
  long long R = std::max<long long>(128, ((64LL << 20) / std::max(KT, N)) / 128 * 128);
  R = std::min(R, M);
  cudaError_t e;
  if ((e = grow(&g_ws.a, &g_ws.a_floats, (size_t)R * KT)) != cudaSuccess) return e;
  if ((e = grow(&g_ws.c, &g_ws.c_floats, (size_t)R * N)) != cudaSuccess) return e;
  if ((e = grow(&g_ws.tc.a, &g_ws.tc.a_floats, (size_t)R * 2 * KT)) != cudaSuccess) return e;
  if ((e = grow(&g_ws.tc.slots, &g_ws.tc.slots_floats, tc_slots_max_floats())) != cudaSuccess) return e;
  if ((e = grow(&g_ws.tc.c, &g_ws.tc.c_floats, (size_t)R * N)) != cudaSuccess) return e;
  for (long long m0 = 0; m0 < M; m0 += R) {
    const int rows = (int)std::min<long long>(R, M - m0);
    if ((e = launch(lc, convtr_im2col_kernel, dim3((rows + 31) / 32, (Cin + 31) / 32), dim3(32, 8), 0, x, g_ws.a, Cin, T_in, Tj, pre_elu, m0,
                    rows)) != cudaSuccess)
      return e;
    GemvParams p;
    p.W = w_phase;
    p.N = N;
    p.K = KT;
    p.M = rows;
    p.X = g_ws.a;
    p.ldx = KT;
    p.Y = g_ws.c;
    p.ldy = N;
    p.tc = &g_ws.tc;
    const float* raw = nullptr;
    p.raw_out = &raw;
    e = rows >= get_tc_min_rows() ? launch_tc_linear(lc, PRO_PLAIN, EPI_STORE, p) : cudaErrorNotSupported;
    if (e == cudaErrorNotSupported) {
      p.raw_out = nullptr;
      p.tc = nullptr;
      e = launch_gemv(lc, PRO_PLAIN, EPI_STORE, p);
    }
    if (e != cudaSuccess) return e;
    const float* src = raw ? raw : g_ws.c;
    if ((e = launch(lc, convtr_tc_epilogue_kernel, dim3((rows + TR_J - 1) / TR_J, (Cout + 31) / 32), dim3(32, TR_J), 0, src, bias, y, Cout, stride,
                    Tj, T_out, crop_left, m0, rows)) != cudaSuccess)
      return e;
  }
  return cudaSuccess;
}

}  // namespace ua2
