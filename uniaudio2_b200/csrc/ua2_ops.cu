// Stand-alone operator entry points of the C ABI (unit-parity surface).  Same kernels as the handle path.
#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"

using namespace ua2;

// scratch for the tiled (many-row) path of the stand-alone operators: row statistics + attention combine
static float* g_ops_ws = nullptr;
static const size_t kOpsWsFloats = (size_t)16 << 20;
static int ops_ws(GemvParams& p) {
  if (p.M < 128) return UA2_OK;
  if (!g_ops_ws) UA2_CHECK_CUDA(cudaMalloc(&g_ops_ws, kOpsWsFloats * sizeof(float)));
  p.ws = g_ops_ws;
  p.ws_floats = kOpsWsFloats;
  return UA2_OK;
}

extern "C" {

int ua2_set_global_option(const char* name, int value) {
  UA2_REQUIRE(name, "null name");
  bump_option_epoch();  // handles drop their captured frame graphs: a replay would still run the kernels chosen under the old options
  if (std::string(name) == "gemv3_balance_grid") {
    set_gemv3_balance_grid(value);
    return UA2_OK;
  }
  if (std::string(name) == "sgemm_min_rows") {
    set_sgemm_min_rows(value);
    return UA2_OK;
  }
  if (std::string(name) == "gemv3_kcw") {
    set_gemv3_kcw(value);
    return UA2_OK;
  }
  if (std::string(name) == "tc_gemm") {
    UA2_REQUIRE(!value || tc_gemm_available(), "no tensor-core path in this build");
    set_tc_gemm(value);
    return UA2_OK;
  }
  if (std::string(name) == "resblock_fused") {
    set_resblock_fused(value);
    return UA2_OK;
  }
  if (std::string(name) == "attn_rows") {  // row-tile attention kernel for many-row launches (ua2_attn.cu); default 1
    set_attn_rows(value);
    return UA2_OK;
  }
  if (std::string(name) == "attn_ring") {  // persistent K/V chunk ring for long batched contexts (ua2_attn.cu); default 0
    set_attn_ring(value);
    return UA2_OK;
  }
  if (std::string(name) == "conv_umma_staged") {
    set_conv_umma_staged(value);
    return UA2_OK;
  }
  if (std::string(name) == "conv_pointwise") {
    set_conv_pointwise(value);
    return UA2_OK;
  }
  if (std::string(name) == "flash_sbuf") {
    set_flash_sbuf(value);
    return UA2_OK;
  }
  if (std::string(name) == "conv_umma") {
    set_conv_umma(value);
    return UA2_OK;
  }
  if (std::string(name) == "conv_tc") {
    UA2_REQUIRE(!value || tc_gemm_available(), "no tensor-core path in this build");
    set_conv_tc(value);
    return UA2_OK;
  }
  if (std::string(name) == "tc_min_rows") {
    set_tc_min_rows(value);
    return UA2_OK;
  }
  if (std::string(name) == "gemv3_prefetch_mb") {
    set_gemv3_prefetch_mb(value, -1);
    return UA2_OK;
  }
  if (std::string(name) == "gemv3_prefetch_idle_mb") {
    set_gemv3_prefetch_mb(-1, value);
    return UA2_OK;
  }
  if (std::string(name) == "gemv3_budget_kb") {
    set_gemv3_budget_kb(value);
    return UA2_OK;
  }
  if (std::string(name) == "gemv3_max_stages") {
    set_gemv3_max_stages(value);
    return UA2_OK;
  }
  if (std::string(name) == "gemv3_ctas_per_sm") {
    set_gemv3_ctas_per_sm(value);
    return UA2_OK;
  }
  UA2_REQUIRE(false, std::string("unknown global option ") + name);
}

int ua2_linear_f32(const float* x, const float* W, const float* norm_w, float eps, const float* residual, float* y,
                   int M, int N, int K, void* stream) {
  UA2_REQUIRE(x && W && y, "null argument");
  UA2_REQUIRE(M >= 1 && N >= 2 && (N % 2) == 0 && K >= 4 && (K % 4) == 0, "need M>=1, even N, K % 4 == 0");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  GemvParams p;
  p.W = W;
  p.N = N;
  p.K = K;
  p.M = M;
  p.X = x;
  p.ldx = K;
  p.norm_w = norm_w;
  p.eps = eps;
  p.Y = y;
  p.ldy = N;
  p.R = residual;
  p.ldr = N;
  if (int rc = ops_ws(p)) return rc;
  UA2_CHECK_CUDA(launch_gemv(lc, norm_w ? PRO_RMSNORM : PRO_PLAIN, residual ? EPI_RESADD : EPI_STORE, p));
  return UA2_OK;
}

// Many-row linear forced onto the tensor-core path (ua2_tcgemm.cu / ua2_umma.cu) whatever M is; scratch owned by the library.
static TcWorkspace g_tc_op_ws;
static int tc_op_ws(size_t a, size_t c) {
  auto grow = [](float** p, size_t* have, size_t want) -> cudaError_t {
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
  };
  UA2_CHECK_CUDA(grow(&g_tc_op_ws.a, &g_tc_op_ws.a_floats, a));
  UA2_CHECK_CUDA(grow(&g_tc_op_ws.c, &g_tc_op_ws.c_floats, c));
  UA2_CHECK_CUDA(grow(&g_tc_op_ws.slots, &g_tc_op_ws.slots_floats, tc_slots_max_floats()));
  return UA2_OK;
}

int ua2_flash_attn_bf16(const void* q16, const void* k16, const void* v16, float* out, int B, int T, int H, int hs, void* stream) {
  UA2_REQUIRE(q16 && k16 && v16 && out, "null argument");
  UA2_REQUIRE(hs == 64, "head size must be 64");
  UA2_REQUIRE(B >= 1 && T >= 1 && H >= 1, "empty problem");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_flash_bf16(lc, q16, k16, v16, out, nullptr, B, T, H, hs));
  return UA2_OK;
}

int ua2_tc_linear_f32(const float* x, const float* W, const float* W2, const float* norm_w, float eps, const float* residual, float* y, int M,
                      int N, int K, void* stream) {
  UA2_REQUIRE(x && W && y, "null argument");
  UA2_REQUIRE(M >= 1 && N >= 4 && (N % 4) == 0 && K >= 4 && (K % 4) == 0, "need M>=1, N % 4 == 0, K % 4 == 0");
  UA2_REQUIRE(!(W2 && residual), "SwiGLU form takes no residual");
  const int Ntot = W2 ? 2 * N : N;
  if (int rc = tc_op_ws((size_t)M * 2 * K, (size_t)M * Ntot)) return rc;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  GemvParams p;
  p.W = W;
  p.W2 = W2;
  p.N = N;
  p.K = K;
  p.M = M;
  p.X = x;
  p.ldx = K;
  p.norm_w = norm_w;
  p.eps = eps;
  p.Y = y;
  p.ldy = N;
  p.R = residual;
  p.ldr = N;
  p.tc = &g_tc_op_ws;
  UA2_CHECK_CUDA(launch_tc_linear(lc, norm_w ? PRO_RMSNORM : PRO_PLAIN, W2 ? EPI_SWIGLU : residual ? EPI_RESADD : EPI_STORE, p));
  return UA2_OK;
}

// one application of the device-side TTS state machine to a sampled row (unit-parity surface of ua2_llm_tts_frames)
int ua2_tts_state_step(const int32_t* sample, int nq, int32_t* state, int32_t* frames_out, int frames_cap, int reason_eos, int end_tok,
                       int reason_card, int fixed_switch, void* stream) {
  UA2_REQUIRE(sample && state && frames_out && nq >= 1 && frames_cap >= 1, "bad argument");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_tts_state(lc, sample, nq, state, frames_out, frames_cap, reason_eos, end_tok, reason_card, fixed_switch));
  return UA2_OK;
}

int ua2_swiglu_f32(const float* x, const float* W1, const float* W2, const float* norm_w, float eps, float* y, int M,
                   int N, int K, void* stream) {
  UA2_REQUIRE(x && W1 && W2 && y, "null argument");
  UA2_REQUIRE(M >= 1 && N >= 1 && K >= 4 && (K % 4) == 0, "need M>=1, K % 4 == 0");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  GemvParams p;
  p.W = W1;
  p.W2 = W2;
  p.N = N;
  p.K = K;
  p.M = M;
  p.X = x;
  p.ldx = K;
  p.norm_w = norm_w;
  p.eps = eps;
  p.Y = y;
  p.ldy = N;
  if (int rc = ops_ws(p)) return rc;
  UA2_CHECK_CUDA(launch_gemv(lc, norm_w ? PRO_RMSNORM : PRO_PLAIN, EPI_SWIGLU, p));
  return UA2_OK;
}

int ua2_qkv_rope_f32(const float* x, const float* Wqkv, const float* norm_w, float eps, const int32_t* pos,
                     const int32_t* bidx, const float* cos, const float* sin, float* q_out, float* k_cache,
                     float* v_cache, int M, int K, int n_head, int n_groups, int hs, int S_max, void* stream) {
  UA2_REQUIRE(x && Wqkv && pos && bidx && cos && sin && q_out && k_cache && v_cache, "null argument");
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "head_size must be 32/64/128");
  UA2_REQUIRE(M >= 1 && K >= 4 && (K % 4) == 0 && n_groups >= 1 && n_head % n_groups == 0, "bad shape");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  GemvParams p;
  p.W = Wqkv;
  p.N = (n_head + 2 * n_groups) * hs;
  p.K = K;
  p.M = M;
  p.X = x;
  p.ldx = K;
  p.norm_w = norm_w;
  p.eps = eps;
  p.pos = pos;
  p.bidx = bidx;
  p.n_head = n_head;
  p.n_groups = n_groups;
  p.hs = hs;
  p.q_out = q_out;
  p.k_cache = k_cache;
  p.v_cache = v_cache;
  p.cos = cos;
  p.sin = sin;
  p.S_max = S_max;
  if (int rc = ops_ws(p)) return rc;
  UA2_CHECK_CUDA(launch_gemv(lc, norm_w ? PRO_RMSNORM : PRO_PLAIN, EPI_QKV, p));
  return UA2_OK;
}

int64_t ua2_attn_workspace_floats(int M, int n_head, int hs, int S_max) {
  const int64_t splits = (S_max + ATTN_CHUNK - 1) / ATTN_CHUNK;
  return (int64_t)M * n_head * splits * (hs + 2);
}

int ua2_attn_f32(const float* q, const float* k_cache, const float* v_cache, const int32_t* pos, const int32_t* bidx,
                 float* y, float* workspace, int M, int n_head, int n_groups, int hs, int S_max, void* stream) {
  UA2_REQUIRE(q && k_cache && v_cache && pos && bidx && y && workspace, "null argument");
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "head_size must be 32/64/128");
  UA2_REQUIRE(M >= 1 && M <= 65535 && n_groups >= 1 && n_head % n_groups == 0 && n_head / n_groups <= 4, "bad shape");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  AttnParams a;
  a.q = q;
  a.k_cache = k_cache;
  a.v_cache = v_cache;
  a.pos = pos;
  a.bidx = bidx;
  a.M = M;
  a.n_head = n_head;
  a.n_groups = n_groups;
  a.hs = hs;
  a.S_max = S_max;
  a.max_splits = (S_max + ATTN_CHUNK - 1) / ATTN_CHUNK;
  a.n_splits_launch = a.max_splits;
  a.o_part = workspace;
  a.ml_part = workspace + (size_t)M * n_head * a.max_splits * hs;
  UA2_CHECK_CUDA(launch_attn(lc, a));
  UA2_CHECK_CUDA(launch_attn_combine(lc, a, y));
  return UA2_OK;
}

int ua2_sample_topk_f32(const float* logits, int R, int V, float temperature, int topk, int forbid_prefix,
                        float cfg_scale, const float* noise, uint64_t seed, uint64_t offset, int32_t* out,
                        void* stream) {
  UA2_REQUIRE(logits && out, "null argument");
  UA2_REQUIRE(temperature > 0.f, "temperature must be > 0");          // model_new.py:165-166
  UA2_REQUIRE(forbid_prefix >= 0, "forbid_prefix must be >= 0");      // :167-168
  UA2_REQUIRE(forbid_prefix < V, "forbid_prefix must be smaller than vocab size");  // :171-172
  UA2_REQUIRE(topk >= 1 && topk <= V - forbid_prefix, "topk must be in 1..effective_vocab given forbid_prefix");  // :179-180
  UA2_REQUIRE(R >= 1, "R >= 1");
  const bool use_cfg = cfg_scale > 1.0f;
  UA2_REQUIRE(!use_cfg || R == 1, "CFG samples one row from (cond, uncond)");
  static FrameScalars* d_fs = nullptr;
  if (!d_fs) UA2_CHECK_CUDA(cudaMalloc(&d_fs, sizeof(FrameScalars)));
  FrameScalars fs;
  fs.temperature = temperature;
  fs.topk = topk;
  fs.forbid_prefix = forbid_prefix;
  fs.cfg_scale = cfg_scale;
  fs.seed = seed;
  fs.offset = offset;
  fs.noise = noise;
  fs.out = out;
  fs.rows = R;
  fs.B = use_cfg ? 2 : R;
  cudaStream_t s = (cudaStream_t)stream;
  UA2_CHECK_CUDA(cudaMemcpyAsync(d_fs, &fs, sizeof(fs), cudaMemcpyHostToDevice, s));
  UA2_CHECK_CUDA(cudaStreamSynchronize(s));  // fs lives on the host stack
  LaunchCtx lc;
  lc.stream = s;
  // under CFG the single sampled token is replicated to both rows (out_ld = 1 -> out[0], out[1])
  UA2_CHECK_CUDA(launch_sampler(lc, logits, V, d_fs, 1, 0, 1, 0, 0, fs.B, R));
  return UA2_OK;
}

}  // extern "C"
