// TEST INFRASTRUCTURE ONLY: C entry points that run the kernels of csrc/ua2_frontend.cu and csrc/ua2_wavlm.cu (the parts of those
// files in front of their host-side C-ABI / handle code) on the CPU shim: resampler, log-mel features, and the WavLM encoder's own
// kernels (first convolution + GroupNorm, im2col, weight repacks, positional convolution, gate, bias table, hidden-state mean).
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {
// csrc/ua2_enc_dev.cuh (not part of this build: its epilogue kernels need the tensor-core GEMM's headers) defines this helper
inline uint2 pack4_bf16(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 o;
  std::memcpy(&o.x, &lo, 4);
  std::memcpy(&o.y, &hi, 4);
  return o;
}
}  // namespace
}  // namespace ua2

#include "ua2_frontend_kernels.inc"
#include "ua2_wavlm_kernels.inc"

extern "C" {
int shim_fe_resample(const float* x, long long ldx, const float* kern, float* y, long long ldy, int B, int L, int n_valid, int n_store, int orig,
                     int newf, int width) {
  ua2::LaunchCtx lc;
  return ua2::launch_fe_resample(lc, x, ldx, kern, y, ldy, B, L, n_valid, n_store, orig, newf, width);
}
int shim_fe_logmel(const float* wav, long long ld, const float* window, const float* filt, float* out, int B, int L, int n_fft, int hop, int n_mels,
                   int n_frames) {
  int rc = shim::run_grid(ua2::fe_logmel_kernel, dim3((unsigned)((n_frames + ua2::FE_FRAMES - 1) / ua2::FE_FRAMES), (unsigned)B), dim3(256), wav, ld,
                          window, filt, out, L, n_fft, hop, n_mels, n_frames);
  if (rc) return rc;
  return shim::run_grid(ua2::fe_logmel_norm_kernel, dim3((unsigned)B), dim3(128), out, (long long)n_mels * n_frames);  // fewer OS threads than 1024
}
int shim_wl_conv0(const float* x, long long ld, const float* w0, const float* b0, const float* gamma, const float* beta, double* part, float* stat,
                  float* out, int B, int L, int T0, int C0, int k0, int s0, float eps) {
  ua2::LaunchCtx lc;
  return ua2::launch_wl_conv0(lc, x, ld, w0, b0, gamma, beta, part, stat, out, B, L, T0, C0, k0, s0, eps);
}
int shim_wl_im2col(const float* in, float* col, int B, int Tin, int Tout, int C, int k, int s) {
  ua2::LaunchCtx lc;
  return ua2::launch_wl_im2col<float>(lc, in, col, B, Tin, Tout, C, k, s);
}
int shim_wl_repack_conv(const float* w, float* out, int Cout, int Cin, int k) {
  return shim::run_grid(ua2::wl_repack_conv_kernel, dim3(ua2::wl_grid((long long)Cout * Cin * k)), dim3(256), w, out, Cout, Cin, k);
}
int shim_wl_posconv(const float* h, const float* w, const float* bias, float* wr, float* p, int B, int T, int D, int cg, int K) {
  int rc = shim::run_grid(ua2::wl_repack_posconv_kernel, dim3(ua2::wl_grid((long long)D * cg * K)), dim3(256), w, wr, D, cg, K);
  if (rc) return rc;
  ua2::LaunchCtx lc;
  return ua2::launch_wl_posconv(lc, h, wr, bias, p, B, T, D, cg, K);
}
int shim_wl_gate(const float* h, const float* Wg, const float* bg, const float* cst, float* gate, int B, int T, int H, int hs) {
  ua2::LaunchCtx lc;
  return ua2::launch_wl_gate(lc, h, Wg, bg, cst, gate, B, T, H, hs);
}
int shim_wl_bias_table(const float* emb, float* tab, int H, int T, int num_buckets, int max_distance) {
  const int n = 2 * T - 1;
  std::vector<int32_t> bucket(n);
  for (int r = 0; r < n; ++r) bucket[r] = ua2::wl_rel_bucket(r - (T - 1), num_buckets, max_distance);
  return shim::run_grid(ua2::wl_bias_table_kernel, dim3((unsigned)((H * n + 255) / 256)), dim3(256), emb, (const int32_t*)bucket.data(), tab, H, n);
}
int shim_wl_axpy(const float* h, float* out, float alpha, int first, long long n4) {
  return shim::run_grid(ua2::wl_axpy_kernel, dim3(ua2::wl_grid(n4)), dim3(256), h, out, alpha, first, n4);
}
}
