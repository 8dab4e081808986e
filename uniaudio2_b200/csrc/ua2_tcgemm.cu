// Many-row linears on the 5th-generation tensor cores with fp32-class accuracy ("3xTF32").
//
// Used (option "tc_gemm") for linears with M >= 128 rows: forward_prefix of long / batched prompts (model_new.py:456-507,
// SURVEY section 8 a13 and config 3).  The precision contract of the path is fp32: a plain tf32 MMA
// (10-bit mantissa) would move logits by ~1e-3 and flip greedy ids, so every operand is split into two tf32 numbers
//     x = hi + lo,   hi = rna_tf32(x),   lo = rna_tf32(x - hi)            (x - hi is exact in fp32)
// and the product is evaluated as  A_hi W_hi + A_lo W_hi + A_hi W_lo  (the dropped lo*lo term is ~2^-22 relative), with
// fp32 accumulation in tensor memory.  The three partial products are ONE GEMM over a 3x longer inner dimension:
//     A3 = [A_hi | A_lo | A_hi]  (M x 3K),    W3 = [W_hi | W_hi | W_lo]  (N x 3K),    C = A3 W3^T.
//
//   tc_split_a_kernel<PRO>   activation rows -> A3, with the row prologue fused (RMSNorm lit_model.py:883-890, embedding
//                            gather model_new.py:662-663)
//   tc_split_w_kernel        weight rows -> W3 (per call, into scratch: 16 B of traffic per weight against the >= 128-fold
//                            reuse of the GEMM; no persistent second copy of the weights)
//   CUTLASS 4.x sm_100a collective mainloop (headers vendored in the image): TMA (UTMALDG) feeds a shared-memory ring,
//                            one elected thread issues tcgen05.mma kind::tf32 (UTCHMMA) on 128 x 128 x 32 tiles,
//                            accumulators in TMEM, tcgen05.ld (LDTM) epilogue -> C (M x N) fp32
//   tc_epilogue_kernel<EPI>  the fused epilogues of the skinny kernels applied to C: store / +residual / SwiGLU /
//                            split + half-split RoPE + KV-cache append (ua2_gemv_dev.cuh)
#include <algorithm>
#include <unordered_map>

#include "ua2_gemv_dev.cuh"
#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"


namespace ua2 {
namespace {

using namespace v1dev;

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
  lo = make_float4(tf32_rna(v.x - hi.x), tf32_rna(v.y - hi.y), tf32_rna(v.z - hi.z), tf32_rna(v.w - hi.w));
}

// one CTA per activation row
// planes = 0: A3 = [hi | lo | hi] rows of 3K (library mainloop); planes = 1: [2][M][K] hi / lo planes (ua2_umma.cu)
template <int PRO>
__global__ void __launch_bounds__(256) tc_split_a_kernel(const GemvParams p, float* __restrict__ A3, int planes) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, tid = threadIdx.x, K = p.K;
  const float* src;
  if (PRO == PRO_GATHER) {
    const long long row = (long long)p.gidx[(size_t)m * p.gidx_stride] + p.gidx_offset;
    src = p.emb + (size_t)row * K;
  } else {
    src = p.X + (size_t)m * p.ldx;
  }
  float rs = 1.f, mean = 0.f;
  if (PRO == PRO_LAYERNORM) {  // two-pass statistics like row_stats_kernel (ua2_sgemm.cu): mean, then centred sum of squares
    float s = 0.f;
    for (int k = tid * 4; k < K; k += 256 * 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + k);
      s += (v.x + v.y) + (v.z + v.w);
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    mean = tot / (float)K;
    __syncthreads();
    float sq = 0.f;
    for (int k = tid * 4; k < K; k += 256 * 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + k);
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      sq += a * a + b * b + c * c + d * d;
    }
    sq = warp_sum(sq);
    if ((tid & 31) == 0) red[tid >> 5] = sq;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    rs = rsqrtf(tot / (float)K + p.eps);
  }
  if (PRO == PRO_RMSNORM) {
    float ss = 0.f;
    for (int k = tid * 4; k < K; k += 256 * 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + k);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    rs = rsqrtf(tot / (float)K + p.eps);
  }
  float* out = planes ? A3 + (size_t)m * K : A3 + (size_t)m * 3 * K;
  const size_t lo_off = planes ? (size_t)p.M * K : (size_t)K;
  for (int k = tid * 4; k < K; k += 256 * 4) {
    float4 v = *reinterpret_cast<const float4*>(src + k);
    if (PRO == PRO_RMSNORM) {
      const float4 g = *reinterpret_cast<const float4*>(p.norm_w + k);
      v = make_float4((v.x * rs) * g.x, (v.y * rs) * g.y, (v.z * rs) * g.z, (v.w * rs) * g.w);
    }
    if (PRO == PRO_LAYERNORM) {  // same expression as the SIMT loader (ua2_sgemm.cu)
      const float4 g = *reinterpret_cast<const float4*>(p.norm_w + k), c = *reinterpret_cast<const float4*>(p.norm_b + k);
      v = make_float4((v.x - mean) * rs * g.x + c.x, (v.y - mean) * rs * g.y + c.y, (v.z - mean) * rs * g.z + c.z,
                      (v.w - mean) * rs * g.w + c.w);
    }
    float4 hi, lo;
    split4(v, hi, lo);
    *reinterpret_cast<float4*>(out + k) = hi;
    *reinterpret_cast<float4*>(out + lo_off + k) = lo;
    if (!planes) *reinterpret_cast<float4*>(out + 2 * K + k) = hi;
  }
}

__global__ void __launch_bounds__(256) tc_split_w_kernel(const float* __restrict__ W, long long n_vec, int K,
                                                         float* __restrict__ W3) {
  pdl_launch_dependents();
  pdl_wait();
  const int kv = K >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / kv;
    const int k = (int)(i - n * kv) * 4;
    const float4 v = ldg_stream(W + n * K + k);
    float4 hi, lo;
    split4(v, hi, lo);
    float* out = W3 + n * 3 * K;
    *reinterpret_cast<float4*>(out + k) = hi;
    *reinterpret_cast<float4*>(out + K + k) = hi;
    *reinterpret_cast<float4*>(out + 2 * K + k) = lo;
  }
}

// thread = (row m, output unit u): the pair of sums the skinny kernels' epilogue expects, read back from C
template <int EPI>
__global__ void __launch_bounds__(256) tc_epilogue_kernel(const GemvParams p, const float* __restrict__ C, int ldc, int n_units,
                                                          const float* __restrict__ slots, const UmmaPlan pl) {
  pdl_launch_dependents();
  pdl_wait();
  const int u = blockIdx.y * blockDim.x + threadIdx.x, m = blockIdx.x;  // rows on grid.x: no 65535 limit
  if (u >= n_units) return;
  int nA, nB, cA, cB;
  if (EPI == EPI_SWIGLU) {
    nA = nB = u;
    cA = u;
    cB = p.N + u;  // fc_2 rows follow the fc_1 rows in W3
  } else if (EPI == EPI_QKV) {
    const int half = p.hs >> 1;
    const int hh = u / half, i = u - hh * half;
    nA = hh * p.hs + i;
    nB = nA + half;
    cA = nA;
    cB = nB;
  } else {
    nA = 2 * u;
    nB = nA + 1;
    cA = nA;
    cB = nB;
  }
  float a = C[(size_t)m * ldc + cA], b = C[(size_t)m * ldc + cB];
  if (slots != nullptr) {  // stream-K tiles continued by other CTAs (ua2_umma.cuh)
    a += umma_side_sum(pl, slots, m, cA, p.N);
    b += umma_side_sum(pl, slots, m, cB, p.N);
  }
  epilogue<EPI>(p, 0, 1, m, a, b, nA, nB);
}

#ifdef UA2_HAVE_CUTLASS
}  // namespace
// the CUTLASS instantiations live in their own translation units (ua2_tcgemm_t128.cu / ua2_tcgemm_t64.cu) so they build in parallel
cudaError_t run_tf32_gemm_128x128(cudaStream_t st, const float* A, const float* B, float* C, int M, int N, int K);
cudaError_t run_tf32_gemm_64x32(cudaStream_t st, const float* A, const float* B, float* C, int M, int N, int K);
namespace {
// 128 x 128 x 32 tiles for prefill passes; 64 x 32 x 32 tiles for decode-sized M (<= 64 rows): N / 32 CTAs keep every SM
// streaming weights where N / 128 would leave most of them idle
cudaError_t run_tf32_gemm(cudaStream_t st, const float* A, const float* B, float* C, int M, int N, int K) {
  return M <= 64 ? run_tf32_gemm_64x32(st, A, B, C, M, N, K) : run_tf32_gemm_128x128(st, A, B, C, M, N, K);
}
#endif

int g_tc_gemm = 1;  // many-row linears (M >= tc_min_rows) on tcgen05; 0 = fp32 SIMT tiles (ua2_sgemm.cu)
int g_tc_impl = 1;  // 1 = hand-written tcgen05 mainloop with the on-chip weight split (ua2_umma.cu); 0 = library collective over pre-split copies
int g_tc_persistent = 0;
int g_tc_min_rows = 32;  // measured at 32 rows: 61 ms (skinny kernels, weights re-streamed per 8-row tile) -> 35 ms per frame

}  // namespace

// Split weights kept across calls (option "tc_persistent_weights"): W3 = [W_hi | W_hi | W_lo] of every weight matrix the
// tensor-core path has served, keyed by the fp32 weight pointer.  Trades 12 B of HBM per parameter for not re-splitting
// (16 B of traffic per parameter per call) - what batched DECODE frames need, where each weight is used by only 16-64 rows.
struct TcWeightCache {
  std::unordered_map<const float*, float*> w3;
  size_t bytes = 0;
  bool full = false;  // a cudaMalloc failed: serve the rest from scratch
};
TcWeightCache* tc_cache_create() { return new TcWeightCache(); }
void tc_cache_destroy(TcWeightCache* c) {
  if (c == nullptr) return;
  for (auto& kv : c->w3) cudaFree(kv.second);
  delete c;
}
size_t tc_cache_bytes(const TcWeightCache* c) { return c ? c->bytes : 0; }
void set_tc_persistent(int v) { g_tc_persistent = v ? 1 : 0; }
int get_tc_persistent() { return g_tc_persistent; }
void set_tc_min_rows(int v) { g_tc_min_rows = v < 1 ? 1 : v; }
int get_tc_min_rows() { return g_tc_min_rows; }

void set_tc_impl(int v) { g_tc_impl = v ? 1 : 0; }
int get_tc_impl() { return g_tc_impl; }
size_t tc_slots_max_floats() { return (size_t)148 * 256 * 128; }
void set_tc_gemm(int v) { g_tc_gemm = v ? 1 : 0; }
int get_tc_gemm() { return g_tc_gemm; }
bool tc_gemm_available() { return true; }  // the hand-written mainloop needs no third-party headers

// returns cudaErrorNotSupported when this (pro, epi, shape, workspace) is not served (caller falls back to the SIMT core)
cudaError_t launch_tc_linear(const LaunchCtx& lc, int pro, int epi, const GemvParams& p) {
  if (p.tc == nullptr || (p.K & 3) || (p.ldx & 3)) return cudaErrorNotSupported;
  if (!(pro == PRO_PLAIN || pro == PRO_RMSNORM || pro == PRO_GATHER || pro == PRO_LAYERNORM)) return cudaErrorNotSupported;
  if (!(epi == EPI_STORE || epi == EPI_RESADD || epi == EPI_SWIGLU || epi == EPI_QKV || epi == EPI_GELU || epi == EPI_SCALE_RESADD ||
        epi == EPI_QKV_IL))
    return cudaErrorNotSupported;
  const TcWorkspace& ws = *p.tc;
  const int M = p.M, K = p.K, K3 = 3 * p.K;
  const int Ntot = epi == EPI_SWIGLU ? 2 * p.N : p.N;
  if ((Ntot & 3) || (size_t)M * K3 > ws.a_floats || (size_t)M * Ntot > ws.c_floats) return cudaErrorNotSupported;
  if (g_tc_impl == 1) {
    const UmmaPlan pl = umma_plan(M, p.N, epi == EPI_SWIGLU ? 2 : 1, K);
    if (pl.grid < 1 || pl.slot_floats > ws.w_floats) return cudaErrorNotSupported;
    cudaError_t e = cudaErrorNotSupported;
#define UA2_TCA(P) \
  if (pro == P) e = launch(lc, tc_split_a_kernel<P>, dim3(M), dim3(256), 0, p, ws.a, 1);
    UA2_TCA(PRO_PLAIN)
    UA2_TCA(PRO_RMSNORM)
    UA2_TCA(PRO_GATHER)
    UA2_TCA(PRO_LAYERNORM)
#undef UA2_TCA
    if (e != cudaSuccess) return e;
    if ((e = run_umma_tf32x3(lc, ws.a, p.W, epi == EPI_SWIGLU ? p.W2 : nullptr, ws.c, Ntot, ws.w, M, p.N, K, pl)) != cudaSuccess) return e;
    if (p.raw_out != nullptr && epi == EPI_STORE) {  // the caller's own epilogue consumes the product in place
      if ((e = run_umma_fixup(lc, ws.c, Ntot, ws.w, M, p.N, 1, pl)) != cudaSuccess) return e;
      *p.raw_out = ws.c;
      return cudaSuccess;
    }
    const float* slots = umma_has_split_tiles(pl) ? ws.w : nullptr;
    const int n_units = epi == EPI_SWIGLU ? p.N : p.N / 2;
    const dim3 grid(M, (n_units + 255) / 256);
#define UA2_TCE(E) \
  if (epi == E) return launch(lc, tc_epilogue_kernel<E>, grid, dim3(256), 0, p, (const float*)ws.c, Ntot, n_units, slots, pl);
    UA2_TCE(EPI_STORE)
    UA2_TCE(EPI_RESADD)
    UA2_TCE(EPI_SWIGLU)
    UA2_TCE(EPI_QKV)
    UA2_TCE(EPI_GELU)
    UA2_TCE(EPI_SCALE_RESADD)
    UA2_TCE(EPI_QKV_IL)
#undef UA2_TCE
    return cudaErrorNotSupported;
  }
#ifndef UA2_HAVE_CUTLASS
  return cudaErrorNotSupported;
#else
  // where the split weights live: the persistent cache (filled on first use, outside any stream capture) or scratch
  const float* w3 = nullptr;
  bool need_split = true;
  float* w3_dst = ws.w;
  TcWeightCache* cache = (g_tc_persistent || ws.force_persistent) ? ws.cache : nullptr;
  if (cache != nullptr) {
    auto it = cache->w3.find(p.W);
    if (it != cache->w3.end()) {
      w3 = it->second;
      need_split = false;
    } else if (!cache->full) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(lc.stream, &cs);
      float* buf = nullptr;
      if (cs == cudaStreamCaptureStatusNone && cudaMalloc(&buf, (size_t)Ntot * K3 * 4) == cudaSuccess) {
        cache->w3.emplace(p.W, buf);
        cache->bytes += (size_t)Ntot * K3 * 4;
        w3 = w3_dst = buf;
      } else {
        cudaGetLastError();  // out of memory: clear the error, use scratch from now on
        if (cs == cudaStreamCaptureStatusNone) cache->full = true;
      }
    }
  }
  if (w3 == nullptr) {
    if ((size_t)Ntot * K3 > ws.w_floats) return cudaErrorNotSupported;
    w3 = ws.w;
  }
  cudaError_t e;
#define UA2_TCA(P) \
  if (pro == P) e = launch(lc, tc_split_a_kernel<P>, dim3(M), dim3(256), 0, p, ws.a, 0);
  e = cudaErrorNotSupported;
  UA2_TCA(PRO_PLAIN)
  UA2_TCA(PRO_RMSNORM)
  UA2_TCA(PRO_GATHER)
  UA2_TCA(PRO_LAYERNORM)
#undef UA2_TCA
  if (e != cudaSuccess) return e;
  if (need_split) {
    const long long n_vec = (long long)p.N * K / 4;
    const int grid = (int)std::min<long long>((n_vec + 255) / 256, 148 * 16);
    if ((e = launch(lc, tc_split_w_kernel, dim3(grid), dim3(256), 0, p.W, n_vec, K, w3_dst)) != cudaSuccess) return e;
    if (epi == EPI_SWIGLU &&
        (e = launch(lc, tc_split_w_kernel, dim3(grid), dim3(256), 0, p.W2, n_vec, K, w3_dst + (size_t)p.N * K3)) != cudaSuccess)
      return e;
  }
  if ((e = run_tf32_gemm(lc.stream, ws.a, w3, ws.c, M, Ntot, K3)) != cudaSuccess) return e;
  if (lc.launch_counter) ++*lc.launch_counter;
  if (p.raw_out != nullptr && epi == EPI_STORE) {  // the caller's own epilogue consumes the product in place
    *p.raw_out = ws.c;
    return cudaSuccess;
  }
  const int n_units = epi == EPI_SWIGLU ? p.N : p.N / 2;
  const dim3 grid(M, (n_units + 255) / 256);
#define UA2_TCE(E) \
  if (epi == E) return launch(lc, tc_epilogue_kernel<E>, grid, dim3(256), 0, p, (const float*)ws.c, Ntot, n_units, (const float*)nullptr, UmmaPlan());
  UA2_TCE(EPI_STORE)
  UA2_TCE(EPI_RESADD)
  UA2_TCE(EPI_SWIGLU)
  UA2_TCE(EPI_QKV)
  UA2_TCE(EPI_GELU)
  UA2_TCE(EPI_SCALE_RESADD)
  UA2_TCE(EPI_QKV_IL)
#undef UA2_TCE
  return cudaErrorNotSupported;
#endif
}

}  // namespace ua2
