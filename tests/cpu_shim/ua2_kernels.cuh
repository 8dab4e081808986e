// TEST INFRASTRUCTURE ONLY: stand-in for csrc/ua2_kernels.cuh + ua2_common.cuh when a .cu of the product is compiled for the CPU
// shim (tests/cpu_shim/cuda_shim.h).  Same names and signatures for what csrc/ua2_convtc.cu and csrc/ua2_resblock.cu use; the
// tensor-core / skinny GEMM launchers are replaced by a plain CPU GEMM (they have their own GPU parity suites).
#pragma once
#include "cuda_shim.h"

namespace ua2 {
enum : int { PRO_PLAIN = 0 };
enum : int { EPI_STORE = 0 };
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = std::fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
struct DeviceOnce {  // one device on the shim
  bool done = false;
  bool need() {
    const bool n = !done;
    done = true;
    return n;
  }
};
struct LaunchCtx {
  cudaStream_t stream = nullptr;
  bool pdl = false;
  int* launch_counter = nullptr;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch(const LaunchCtx& lc, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t, Args... args) {
  if (lc.launch_counter) ++*lc.launch_counter;
  return shim::run_grid(kernel, grid, block, args...);
}
inline size_t tc_slots_max_floats() { return 16; }
struct TcWorkspace {
  float* a = nullptr;
  size_t a_floats = 0;
  float* slots = nullptr;
  size_t slots_floats = 0;
  float* c = nullptr;
  size_t c_floats = 0;
};
struct GemvParams {
  const float* W = nullptr;
  int N = 0, K = 0, M = 0;
  const float* X = nullptr;
  int ldx = 0;
  float* Y = nullptr;
  int ldy = 0;
  const TcWorkspace* tc = nullptr;
  const float** raw_out = nullptr;
};
inline bool tc_gemm_available() { return true; }
inline int get_tc_gemm() { return 1; }
inline int get_tc_min_rows() { return 32; }
inline void cpu_gemm(const GemvParams& p, float* C, int ldc) {
  for (int m = 0; m < p.M; ++m)
    for (int n = 0; n < p.N; ++n) {
      double acc = 0.0;
      for (int k = 0; k < p.K; ++k) acc += (double)p.X[(size_t)m * p.ldx + k] * p.W[(size_t)n * p.K + k];
      C[(size_t)m * ldc + n] = (float)acc;
    }
}
// the product's tensor-core path: the raw product stays in the workspace and its address is reported through raw_out
inline cudaError_t launch_tc_linear(const LaunchCtx&, int, int, const GemvParams& p) {
  if (p.tc == nullptr || (size_t)p.M * p.N > p.tc->c_floats) return cudaErrorNotSupported;
  cpu_gemm(p, p.tc->c, p.N);
  if (p.raw_out) *p.raw_out = p.tc->c;
  return cudaSuccess;
}
inline cudaError_t launch_gemv(const LaunchCtx&, int, int, const GemvParams& p) {
  cpu_gemm(p, p.Y, p.ldy);
  return cudaSuccess;
}
cudaError_t launch_conv1d_tc(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res, float* y,
                             int B, int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation, int pad_left, int pre_elu,
                             int replicate, const float* prelu = nullptr);
cudaError_t launch_convtr1d_tc(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin,
                               int Cout, int T_in, int stride, int pre_elu, int crop_left, int T_out);
cudaError_t launch_resblock_fused(const LaunchCtx& lc, const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  float* y, int B, int C, int H, int T);
void set_conv_tc(int v);
int get_conv_tc();
void set_resblock_fused(int v);
int get_resblock_fused();
}  // namespace ua2
