// TEST INFRASTRUCTURE ONLY: the launcher that the shim cannot run (tcgen05 GEMM) replaced by a CPU GEMM, plus
// the error plumbing of ua2_llm.cu, for building csrc/ua2_codec.cu + ua2_sgemm.cu + ua2_convtc.cu + ua2_resblock.cu with -DUA2_CPU_SHIM.
#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"

namespace ua2 {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
size_t tc_slots_max_floats() { return 16; }
static int g_tc_avail = 1;
bool tc_gemm_available() { return g_tc_avail != 0; }
int get_tc_gemm() { return g_tc_avail; }
int get_tc_min_rows() { return 32; }
static void cpu_gemm(const GemvParams& p, float* C, int ldc) {
  for (int m = 0; m < p.M; ++m)
    for (int n = 0; n < p.N; ++n) {
      double acc = 0.0;
      for (int k = 0; k < p.K; ++k) acc += (double)p.X[(size_t)m * p.ldx + k] * p.W[(size_t)n * p.K + k];
      C[(size_t)m * ldc + n] = (float)acc;
    }
}
cudaError_t launch_tc_linear(const LaunchCtx&, int pro, int epi, const GemvParams& p) {
  if (pro != PRO_PLAIN || epi != EPI_STORE || p.tc == nullptr || (size_t)p.M * p.N > p.tc->c_floats) return cudaErrorNotSupported;
  cpu_gemm(p, p.tc->c, p.N);
  if (p.raw_out) {
    *p.raw_out = p.tc->c;
  } else {
    for (int m = 0; m < p.M; ++m) std::memcpy(p.Y + (size_t)m * p.ldy, p.tc->c + (size_t)m * p.N, sizeof(float) * p.N);
  }
  return cudaSuccess;
}
// the tcgen05 convolution path does not exist on the shim: the option reads 0 and the launchers decline, so the codec model
// takes the fp32 SIMT / conv_tc kernels that the shim can run
int get_conv_umma() { return 0; }
void set_conv_umma(int) {}
void set_conv_umma_staged(int) {}
cudaError_t launch_conv1d_umma(const LaunchCtx&, const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int,
                               int, int, int, int, const float*) {
  return cudaErrorNotSupported;
}
cudaError_t launch_convtr1d_umma(const LaunchCtx&, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int) {
  return cudaErrorNotSupported;
}
}  // namespace ua2

extern "C" {
const char* ua2_last_error(void) { return ua2::g_err.c_str(); }
void shim_set_sm_count(int n) { g_shim_sm_count = n; }  // before the first launch of a persistent kernel (launchers cache it)
// ua2_attn_f32 of ua2_ops.cu (outside this build) + the Moshi context window and the "attn_ring" option
int shim_attn(const float* q, const float* k_cache, const float* v_cache, const int32_t* pos, const int32_t* bidx, float* y, float* workspace,
              int M, int n_head, int n_groups, int hs, int S_max, int window, int ring, int* grid_x) {
  ua2::LaunchCtx lc;
  ua2::AttnParams a;
  a.q = q; a.k_cache = k_cache; a.v_cache = v_cache; a.pos = pos; a.bidx = bidx;
  a.M = M; a.n_head = n_head; a.n_groups = n_groups; a.hs = hs; a.S_max = S_max; a.window = window;
  a.max_splits = (S_max + ua2::ATTN_CHUNK - 1) / ua2::ATTN_CHUNK;
  a.n_splits_launch = a.max_splits;
  a.o_part = workspace;
  a.ml_part = workspace + (size_t)M * n_head * a.max_splits * hs;
  ua2::set_attn_ring(ring);
  cudaError_t e = ua2::launch_attn(lc, a);
  ua2::set_attn_ring(0);
  *grid_x = (int)shim::g_last_grid.x;
  if (e != cudaSuccess) return (int)e;
  return (int)ua2::launch_attn_combine(lc, a, y);
}
// the decode-frame linear (csrc/ua2_gemv3.cu) as ua2_linear_f32 / ua2_swiglu_f32 of ua2_ops.cu set it up: norm_w -> PRO_RMSNORM,
// residual -> EPI_RESADD, W2 -> EPI_SWIGLU
int shim_gemv3(const float* x, const float* W, const float* W2, const float* norm_w, float eps, const float* residual, float* y, int M, int N,
               int K, int* grid_y) {
  ua2::LaunchCtx lc;
  ua2::GemvParams p;
  p.W = W; p.W2 = W2; p.N = N; p.K = K; p.M = M; p.X = x; p.ldx = K; p.norm_w = norm_w; p.eps = eps;
  p.Y = y; p.ldy = N; p.R = residual; p.ldr = N;
  const int epi = W2 ? ua2::EPI_SWIGLU : (residual ? ua2::EPI_RESADD : ua2::EPI_STORE);
  const cudaError_t e = ua2::launch_gemv3(lc, norm_w ? ua2::PRO_RMSNORM : ua2::PRO_PLAIN, epi, p, 1);
  *grid_y = (int)shim::g_last_grid.y;
  return (int)e;
}
void shim_set_conv_tc(int v) { ua2::set_conv_tc(v); }
void shim_set_tc_available(int v) { ua2::g_tc_avail = v; }  // 0: the build behaves like one without the CUTLASS headers  // ua2_set_global_option("conv_tc") lives in ua2_ops.cu, outside this build
int ua2_linear_f32(const float* x, const float* W, const float* norm_w, float, const float* residual, float* y, int M, int N, int K, void*) {
  if (norm_w || residual) return UA2_ERR_INVALID;
  ua2::GemvParams p;
  p.W = W; p.N = N; p.K = K; p.M = M; p.X = x; p.ldx = K; p.Y = y; p.ldy = N;
  ua2::cpu_gemm(p, y, N);
  return UA2_OK;
}
}
