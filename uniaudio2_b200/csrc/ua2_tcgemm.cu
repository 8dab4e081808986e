// Many-row linears on the 5th-generation tensor cores with fp32-class accuracy ("3xTF32").
//
// Used (option "tc_gemm") for linears with M >= tc_min_rows rows: forward_prefix of long / batched prompts (model_new.py:456-507,
// SURVEY section 8 a13 and config 3), batched decode frames, the codec transformer, the flow decoder and the wide convolutions.
// The precision contract of the path is fp32: a plain tf32 MMA (10-bit mantissa) would move logits by ~1e-3
// and flip greedy ids, so every operand is split into two tf32 numbers
//     x = hi + lo,   hi = rna_tf32(x),   lo = rna_tf32(x - hi)            (x - hi is exact in fp32)
// and the product is evaluated as  A_hi W_hi + A_lo W_hi + A_hi W_lo  (the dropped lo*lo term is ~2^-22 relative), with
// fp32 accumulation in tensor memory.
//
//   tc_split_a_kernel<PRO>   activation rows -> [2][M][K] hi / lo planes, with the row prologue fused (RMSNorm lit_model.py:883-890,
//                            LayerNorm, embedding gather model_new.py:662-663)
//   umma_kernel (ua2_umma.cu) hand-written TMA / tcgen05 / TMEM mainloop; the fp32 WEIGHTS are read once and split on chip
//   tc_epilogue_kernel<EPI>  the fused epilogues of the skinny kernels applied to the product (+ the stream-K side slots):
//                            store / +residual / SwiGLU / split + half-split RoPE + KV-cache append (ua2_gemv_dev.cuh)
#include <algorithm>

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

// one CTA per activation row; X2 = [2][M][K] hi / lo planes
template <int PRO>
__global__ void __launch_bounds__(256) tc_split_a_kernel(const GemvParams p, float* __restrict__ X2) {
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
  float* out = X2 + (size_t)m * K;
  const size_t lo_off = (size_t)p.M * K;
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

int g_tc_gemm = 1;      // many-row linears (M >= tc_min_rows) on tcgen05; 0 = fp32 SIMT tiles (ua2_sgemm.cu)
int g_tc_min_rows = 32;  // measured at 32 rows: 61 ms per frame on the skinny kernels (weights re-streamed per 8-row tile)

}  // namespace

static unsigned long long g_option_epoch = 0;
void bump_option_epoch() { ++g_option_epoch; }
unsigned long long option_epoch() { return g_option_epoch; }
void set_tc_min_rows(int v) { g_tc_min_rows = v < 1 ? 1 : v; }
int get_tc_min_rows() { return g_tc_min_rows; }
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
  const int M = p.M, K = p.K;
  const int Ntot = epi == EPI_SWIGLU ? 2 * p.N : p.N;
  if ((Ntot & 3) || (size_t)M * 2 * K > ws.a_floats || (size_t)M * Ntot > ws.c_floats) return cudaErrorNotSupported;
  const UmmaPlan pl = umma_plan(M, p.N, epi == EPI_SWIGLU ? 2 : 1, K);
  if (pl.grid < 1 || pl.slot_floats > ws.slots_floats) return cudaErrorNotSupported;
  cudaError_t e = cudaErrorNotSupported;
#define UA2_TCA(P) \
  if (pro == P) e = launch(lc, tc_split_a_kernel<P>, dim3(M), dim3(256), 0, p, ws.a);
  UA2_TCA(PRO_PLAIN)
  UA2_TCA(PRO_RMSNORM)
  UA2_TCA(PRO_GATHER)
  UA2_TCA(PRO_LAYERNORM)
#undef UA2_TCA
  if (e != cudaSuccess) return e;
  if ((e = run_umma_tf32x3(lc, ws.a, p.W, epi == EPI_SWIGLU ? p.W2 : nullptr, ws.c, Ntot, ws.slots, M, p.N, K, pl)) != cudaSuccess) return e;
  if (p.raw_out != nullptr && epi == EPI_STORE) {  // the caller's own epilogue consumes the product in place
    if ((e = run_umma_fixup(lc, ws.c, Ntot, ws.slots, M, p.N, 1, pl)) != cudaSuccess) return e;
    *p.raw_out = ws.c;
    return cudaSuccess;
  }
  const float* slots = umma_has_split_tiles(pl) ? ws.slots : nullptr;
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

}  // namespace ua2
