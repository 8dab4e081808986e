// TEST INFRASTRUCTURE ONLY: C entry points that run kernels of csrc/ua2_stream.cu and csrc/ua2_dit.cu (the parts of those files
// in front of their host-side handle code) on the CPU shim.  These kernels are parity-green on the GPU; running their source on
// the CPU as well lets the CPU-only test tier notice a regression in them.
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {
alignas(16) unsigned char st_dyn[1 << 16];  // dynamic shared memory of sample_token_kernel (top-p sort buffers)
}
}  // namespace ua2

#include "ua2_philox.cuh"
#include "ua2_stream_kernels.inc"
#include "ua2_dit_kernels.inc"

extern "C" {
int shim_rope_ring_append(const float* qkv, int ld, const int32_t* pos, const int32_t* bidx, const float* freqs, float* q_out, float* kc,
                          float* vc, int M, int H, int hs, int cap, int ring) {
  const long long total = (long long)M * H * (hs / 2);
  return shim::run_grid(ua2::rope_ring_append_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), qkv, ld, pos, bidx, freqs, q_out, kc, vc,
                        M, H, hs, cap, ring);
}
int shim_ring_attn(const float* q, const float* kc, const float* vc, const int32_t* pos, const int32_t* bidx, float* y, int M, int H, int hs,
                   int cap, long long end, int ring, int causal, int context, int n_splits, float* part_ml, float* part_acc) {
  ua2::RingAttnParams p{};
  p.q = q; p.kc = kc; p.vc = vc; p.pos = pos; p.bidx = bidx; p.y = y; p.M = M; p.H = H; p.cap = cap; p.end = end;
  p.ring = ring; p.causal = causal; p.context = context; p.n_splits = n_splits; p.part_ml = part_ml; p.part_acc = part_acc;
  ua2::LaunchCtx lc;
  return ua2::launch_ring_attn(lc, p, hs);
}
int shim_sample_token(const float* logits, int R, int V, int use_sampling, float temp, int top_k, float top_p, int end_token,
                      const float* noise, long long* out) {
  ua2::SampleTokenArgs a{logits, V, use_sampling, temp, top_k, top_p, end_token, noise, 0ull, 0ull, out};
  return shim::run_grid(ua2::sample_token_kernel, dim3(R), dim3(ua2::ST_THREADS), a);
}
int shim_dit_attn(const float* q, const float* k, const float* v, float* out, int B, int T, int H, int hs) {
  ua2::LaunchCtx lc;
  return ua2::launch_dit_attn(lc, q, k, v, out, B, T, H, hs);
}
int shim_dit_attn_bias(const float* q, const float* k, const float* v, float* out, int B, int T, int H, int hs, const float* gate, const float* tab) {
  ua2::LaunchCtx lc;
  return ua2::launch_dit_attn_bias(lc, q, k, v, out, B, T, H, hs, gate, tab);
}
int shim_dit_ln_mod(const float* x, float* out, const float* table, const float* t, int t_row, int t_stride, int shift_idx, int scale_idx,
                    float eps, int M, int T, int D) {
  return shim::run_grid(ua2::dit_ln_mod_kernel, dim3(M), dim3(256), x, out, (__nv_bfloat16*)nullptr, table, t, t_row, t_stride, shift_idx, scale_idx, eps, T, D);
}
int shim_dit_im2col3(const float* x, float* out, int B, int T, int C) {
  return shim::run_grid(ua2::dit_im2col3_kernel, dim3(ua2::grid_for((long long)B * T * 3 * C)), dim3(256), x, out, B, T, C);
}
int shim_dit_gate_res(const float* src, const float* bias, float* res, const float* table, const float* t6, int gate_idx, int M, int N, int T) {
  ua2::DitEpi e{};
  e.src = src; e.y = nullptr; e.bias = bias; e.M = M; e.N = N; e.T = T; e.y2 = res; e.table = table; e.t6 = t6; e.gate_idx = gate_idx;
  ua2::LaunchCtx lc;
  return ua2::launch_epi<ua2::DE_GATE_RES>(lc, e);
}
int shim_dit_qkv_split(const float* src, const float* bias, float* q, float* k, float* v, int M, int D, int T, int H, int hs) {
  ua2::DitEpi e{};
  e.src = src; e.y = nullptr; e.bias = bias; e.M = M; e.N = 3 * D; e.T = T; e.q = q; e.k = k; e.v = v; e.H = H; e.hs = hs;
  ua2::LaunchCtx lc;
  return ua2::launch_epi<ua2::DE_QKV_SPLIT>(lc, e);
}
int shim_dit_bias_gelu(const float* src, const float* bias, float* y, int M, int N) {
  ua2::DitEpi e{};
  e.src = src; e.y = y; e.bias = bias; e.M = M; e.N = N; e.T = 1;
  ua2::LaunchCtx lc;
  return ua2::launch_epi<ua2::DE_BIAS_GELU>(lc, e);
}
int shim_dit_euler(float* x, const float* noise, const float* incontext, const float* mu, float* inp, const float* d, int T, int lat, int cond,
                   int ic, float t, float one_minus_sigma, float g, float dt, int do_update) {
  if (!do_update)
    return shim::run_grid(ua2::dit_euler_pack_kernel, dim3(ua2::grid_for((long long)T * (2 * lat + cond))), dim3(256), x, noise, incontext, mu, inp, T,
                          lat, cond, ic, t, one_minus_sigma);
  return shim::run_grid(ua2::dit_euler_update_kernel, dim3(ua2::grid_for((long long)T * lat)), dim3(256), x, d, (long long)T * lat, g, dt);
}
}
