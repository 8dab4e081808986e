// TEST INFRASTRUCTURE ONLY: C entry points that run the product's launchers of ua2_convtc.cu / ua2_resblock.cu on the CPU shim.
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {
alignas(16) float rb_smem[1 << 16];  // the dynamic shared memory of resblock64_kernel (`extern __shared__` in the kernel)
}
}  // namespace ua2

#include "ua2_convtc_shim.inc"
#include "ua2_resblock_shim.inc"

extern "C" {
int shim_conv1d_tc(const float* x, const float* w, const float* bias, const float* res, float* y, int B, int Cin, int Cout, int T_in, int K,
                   int stride, int dilation, int pre_elu, int replicate) {
  const int k_eff = (K - 1) * dilation + 1, pad = k_eff - stride, T_out = (T_in + stride - 1) / stride;  // ua2_conv1d_causal_gemm_f32
  ua2::LaunchCtx lc;
  return ua2::launch_conv1d_tc(lc, x, w, bias, res, y, B, Cin, Cout, T_in, T_out, K, stride, dilation, pad, pre_elu, replicate);
}
int shim_convtr1d_tc(const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin, int Cout, int T_in, int stride,
                     int pre_elu) {
  ua2::LaunchCtx lc;
  return ua2::launch_convtr1d_tc(lc, x, w_phase, bias, y, B, Cin, Cout, T_in, stride, pre_elu, 0, T_in * stride);  // ua2_convtr1d_causal_gemm_f32
}
int shim_resblock(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y, int B, int C, int H, int T) {
  ua2::LaunchCtx lc;
  return ua2::launch_resblock_fused(lc, x, w1, b1, w2, b2, y, B, C, H, T);
}
}
