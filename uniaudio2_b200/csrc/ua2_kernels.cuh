// Kernel launch surface shared between the handle API (ua2_llm.cu) and the stand-alone operator API.
#pragma once
#include <vector>
#include "ua2_common.cuh"

namespace ua2 {

// ---------------------------------------------------------------- skinny linear (GEMV family)
// PRO_ATTN_DIRECT: the whole single-query attention over a SHORT cache (<= ATTN_DIRECT_MAX_KEYS keys, head size 64: the
// local decoder, model_new.py:626-643) is computed inside the proj kernel's prologue - no attention launch at all.
enum : int { PRO_PLAIN = 0, PRO_RMSNORM = 1, PRO_GATHER = 2, PRO_ATTN = 3, PRO_LAYERNORM = 4, PRO_ATTN_DIRECT = 5 };
constexpr int ATTN_DIRECT_MAX_KEYS = 8;
// EPI_GELU / EPI_SCALE_RESADD / EPI_QKV_IL serve the Moshi-family transformer layer (llm_modules/transformer.py:430-588)
enum : int { EPI_STORE = 0, EPI_RESADD = 1, EPI_QKV = 2, EPI_SWIGLU = 3, EPI_GELU = 4, EPI_SCALE_RESADD = 5, EPI_QKV_IL = 6 };

constexpr int ATTN_CHUNK = 64;  // keys per split CTA of the attention kernel

// L2 prefetch of the NEXT linears' weights, issued by the tail of the current kernel (fire-and-forget
// cp.async.bulk.prefetch.L2): the HBM pipe keeps streaming through the kernel boundary / the attention and sampler
// kernels, and the next kernel's first ring fills hit L2.  One spec describes the first `n` units of every CTA slab of
// a following gemv3 launch (same slab partition as that kernel computes for itself).
struct PfSpec {
  const float* W = nullptr;
  const float* W2 = nullptr;  // SwiGLU second matrix
  int K = 0;
  int n_units = 0;  // units of the next kernel
  int G = 0;        // its grid (CTA slabs)
  int mode = 0;     // 0 one row per unit, 1 SwiGLU (row u of W and of W2), 2 QKV rotation pair (rows nA, nA + hs/2)
  int hs = 0;
  int n = 0;  // units to prefetch per slab
};
constexpr int PF_MAX = 3;

// scratch of the tensor-core path for many-row linears (ua2_tcgemm.cu / ua2_umma.cu): split activations, stream-K side slots, raw product
void bump_option_epoch();          // process-wide options changed (ua2_set_global_option)
unsigned long long option_epoch();
void set_tc_min_rows(int v);
int get_tc_min_rows();
size_t tc_slots_max_floats();  // largest side-slot scratch a launch may need: 148 CTAs x 256 x 128 floats
struct TcWorkspace {
  float* a = nullptr;  // [2][M][K] hi / lo planes of the activation rows
  size_t a_floats = 0;
  float* slots = nullptr;  // partial tiles of stream-K continuation CTAs
  size_t slots_floats = 0;
  float* c = nullptr;  // (M, N_total)
  size_t c_floats = 0;
};

struct GemvParams {
  // weights: W (N x K) row-major fp32 (nn.Linear layout); W2 second matrix for SwiGLU
  const float* W = nullptr;
  const float* W2 = nullptr;
  int N = 0, K = 0, M = 0;
  // ---- prologue (how the M x K activation tile is produced in shared memory)
  const float* X = nullptr;  // PLAIN / RMSNORM source, row stride ldx
  int ldx = 0;
  const float* norm_w = nullptr;  // RMSNORM / LAYERNORM weight (K)
  const float* norm_b = nullptr;  // LAYERNORM bias (K)
  float eps = 0.f;
  const float* emb = nullptr;  // GATHER: X[m] = emb[(gidx[m*gidx_stride] + gidx_offset) * K ...]
  const int32_t* gidx = nullptr;
  int gidx_stride = 0, gidx_offset = 0;
  const float* o_part = nullptr;  // ATTN: split-softmax partials of the attention kernel
  const float* ml_part = nullptr;
  int max_splits = 0;
  int n_splits = 0;  // splits launched by the attention kernel (empty ones carry (m=-inf, l=0))
  const int32_t* pos = nullptr;   // (M) cache slot / position of each row
  const int32_t* bidx = nullptr;  // (M) batch row of each row
  int bidx_identity = 0;          // decode frames: row m IS batch row m (kernels may skip the bidx load)
  int n_head = 0, n_groups = 0, hs = 0;
  // ---- epilogue
  float* Y = nullptr;  // STORE / RESADD / SWIGLU destination, row stride ldy
  int ldy = 0;
  const float* R = nullptr;  // RESADD residual, row stride ldr (may alias Y)
  int ldr = 0;
  const float* scale = nullptr;  // SCALE_RESADD: per-output-channel LayerScale (N)
  float* ws = nullptr;           // optional scratch for the tiled path (row statistics 2*M floats, attention combine M*K floats)
  size_t ws_floats = 0;
  float rope_max_period = 10000.f;  // QKV_IL: interleaved-pair RoPE computed on the fly (llm_modules/rope.py:11-68)
  float* q_out = nullptr;  // QKV: roped queries (M, n_head*hs)
  float* k_cache = nullptr;  // (B, G, S_max, hs)
  float* v_cache = nullptr;
  const float* cos = nullptr;  // (positions, hs)
  const float* sin = nullptr;
  int S_max = 0;
  const TcWorkspace* tc = nullptr;  // optional: enables the tcgen05 3xTF32 path for M >= sgemm_min_rows
  // optional (host side only, EPI_STORE): when the tensor-core path serves the call, leave the raw product (M x N, row stride
  // N) in the workspace, skip the copy-out epilogue and report its address here - for callers that run their own fused
  // epilogue over it (ua2_dit.cu).  Left untouched (nullptr) when another path ran: the result is then in Y as usual.
  const float** raw_out = nullptr;
  // ---- tail prefetch (filled by launch_gemv from the recorded launch sequence; see GemvSeq)
  int n_pf = 0;
  PfSpec pf[PF_MAX];
};

// The sequence of skinny linears of one frame, recorded on the first (eager) run of a shape and used from then on to
// tell every launch which weights come next.  idle_after: the launches after this op that do not touch HBM much
// (1 = attention, 2 = sampler) - the prefetch budget grows accordingly.
struct GemvSeqEntry {
  const float* W = nullptr;
  const float* W2 = nullptr;
  int N = 0, K = 0, M = 0, epi = 0, hs = 0;
  int idle_after = 0;
};
struct GemvSeq {
  std::vector<GemvSeqEntry> ops;
  int pos = 0;
  bool recorded = false;
};
void set_gemv3_prefetch_mb(int mb, int idle_mb);
size_t gemv3_prefetch_budget(int idle_after);
int gemv3_make_pf(const GemvSeqEntry* next, int n_next, size_t budget_bytes, PfSpec* out);
cudaError_t launch_gemv(const LaunchCtx& lc, int pro, int epi, const GemvParams& p);
cudaError_t launch_sgemm_linear(const LaunchCtx& lc, int pro, int epi, const GemvParams& p, float* stats_ws);
cudaError_t launch_tc_linear(const LaunchCtx& lc, int pro, int epi, const GemvParams& p);
void set_tc_gemm(int v);
int get_tc_gemm();
bool tc_gemm_available();
int get_sgemm_min_rows();
void set_sgemm_min_rows(int v);
cudaError_t launch_convtr1d_gemm(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B,
                                 int Cin, int Cout, int T_in, int stride, int pre_elu, int crop_left, int T_out,
                                 const float* prelu = nullptr);
// implicit-GEMM causal conv1d (weights in torch layout (Cout, Cin, Ktaps))
cudaError_t launch_conv1d_gemm(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res,
                               float* y, int B, int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation,
                               int pad_left, int pre_elu, int replicate, const float* prelu = nullptr);  // 1 = register-streamed LDG, 2 = per-warp bulk-copy rings, 3 = persistent slab + K-split rings (default)
// wide convolutions (Cin * Ktaps >= 1024) as im2col + tcgen05 3xTF32 GEMM (ua2_convtc.cu; option "conv_tc", default 0)
cudaError_t launch_conv1d_tc(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res, float* y,
                             int B, int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation, int pad_left, int pre_elu,
                             int replicate, const float* prelu = nullptr);
cudaError_t launch_convtr1d_tc(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin,
                               int Cout, int T_in, int stride, int pre_elu, int crop_left, int T_out);
// narrow / strided causal convolutions (Cin % 32 == 0, 16 <= Cout <= 256) as an implicit GEMM on tcgen05 with the activations going
// global -> registers -> tensor memory (ua2_convumma.cu; option "conv_umma", default 1)
cudaError_t launch_conv1d_umma(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res, float* y, int B,
                               int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation, int pad_left, int pre_elu,
                               int replicate, const float* prelu = nullptr);
cudaError_t launch_convtr1d_umma(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin, int Cout,
                                 int T_in, int stride, int pre_elu, int crop_left, int T_out);
void set_conv_pointwise(int v);
int get_conv_pointwise();
void set_conv_umma(int v);
void set_conv_umma_staged(int v);
int get_conv_umma();
// fused SEANet residual block for the 64-channel / 24 kHz stages (ua2_resblock.cu; option "resblock_fused", default 0)
cudaError_t launch_resblock_fused(const LaunchCtx& lc, const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  float* y, int B, int C, int H, int T);
void set_resblock_fused(int v);
int get_resblock_fused();
void set_conv_tc(int v);
int get_conv_tc();
cudaError_t launch_gemv3(const LaunchCtx& lc, int pro, int epi, const GemvParams& p, int n_splits);
void set_gemv3_ctas_per_sm(int v);
void set_gemv3_max_stages(int v);
void set_gemv3_kcw(int v);
void set_gemv3_balance_grid(int v);
void set_gemv3_budget_kb(int v);

// ---------------------------------------------------------------- attention over the KV cache
struct AttnParams {
  const float* q = nullptr;  // (M, n_head*hs) roped
  const float* k_cache = nullptr;
  const float* v_cache = nullptr;
  const int32_t* pos = nullptr;
  const int32_t* bidx = nullptr;
  float* o_part = nullptr;   // (M, n_head, max_splits, hs)
  float* ml_part = nullptr;  // (M, n_head, max_splits, 2)
  int M = 0, n_head = 0, n_groups = 0, hs = 0, S_max = 0, max_splits = 0;
  int window = 0;  // > 0: only keys with pos - j < window are visible (Moshi `context`, transformer.py:404-408)
  int n_splits_launch = 0;  // grid.x (>= splits needed by the largest pos in this launch)
  int bidx_identity = 0;    // decode frames: row m is batch row m
};
cudaError_t launch_attn(const LaunchCtx& lc, const AttnParams& p);
cudaError_t launch_attn_combine(const LaunchCtx& lc, const AttnParams& p, float* y);
// persistent 3-slot K/V chunk ring for long batched contexts (ua2_attn.cu; option "attn_ring", default 0)
void set_attn_rows(int v);
int get_attn_rows();
void set_attn_ring(int v);
int get_attn_ring();

// ---------------------------------------------------------------- small fused elementwise kernels
struct FrameScalars {  // per-call scalars living in device memory so captured graphs stay valid
  float temperature;
  int topk;
  int forbid_prefix;
  float cfg_scale;
  unsigned long long seed;
  unsigned long long offset;
  const float* noise;  // nullable
  int32_t* out;        // (B, 1+nq)
  int rows;            // sampled rows R
  int B;
};

// embedding merge of model_new.py:598-604: audio_in[m] = sum_c mask[m,c] * E[tok[m,c] + c*V]; text_emb[m] = wte[tok[m,nq]]
cudaError_t launch_embed(const LaunchCtx& lc, const int64_t* tokens, const uint8_t* mask, const float* audio_emb,
                         const float* wte, float* audio_in, float* text_emb, int M, int nq, int V, int D, int text_vocab, int* err_flag);
// out[m] = RMSNorm(x[m]; w) * ma[m] + add[m] * mt[m]; optional normed copy kept in `keep`
//   ma = mask[m*(nq+1)+0], mt = mask[m*(nq+1)+nq] (audio-step / text-step masks, model_new.py:594-595)
cudaError_t launch_norm_mix(const LaunchCtx& lc, const float* x, const float* w, float eps, const uint8_t* mask,
                            int nq, const float* add, float* keep, float* out, int M, int D, int mode);
enum : int { MIX_UND_TO_BACKBONE = 0, MIX_BACKBONE_TO_GEN = 1, MIX_FINAL = 2, MIX_NORM_ONLY = 3 };

// per-frame begin: copy caller tokens/mask into the handle's fixed buffers and write scalars
cudaError_t launch_frame_begin(const LaunchCtx& lc, const int64_t* tokens, const uint8_t* mask, int n_tok,
                               int64_t* d_tokens, uint8_t* d_mask, int32_t* d_pos, int32_t* d_bidx, int B,
                               int32_t pos_value, FrameScalars* d_fs, FrameScalars fs);
// device-side phase / EOS state machine of the TTS task loop (ua2_llm_tts_frames)
cudaError_t launch_tts_begin(const LaunchCtx& lc, const int64_t* tokens0, const uint8_t* mask0, const int32_t* prev_sample, int nq,
                             int64_t* d_tokens, uint8_t* d_mask, int32_t* d_pos, int32_t* d_bidx, int32_t pos_value, FrameScalars* d_fs,
                             FrameScalars fs, const int32_t* state);
cudaError_t launch_tts_state(const LaunchCtx& lc, const int32_t* sample, int nq, int32_t* state, int32_t* frames_out, int frames_cap,
                             int reason_eos, int end_tok, int reason_card, int fixed_switch);
cudaError_t launch_prefill_begin(const LaunchCtx& lc, const int64_t* pos64, int32_t* d_pos, int32_t* d_bidx, int M,
                                 int T, int row0);
// audio_head (nq, d, V) -> (nq, V, d)
cudaError_t launch_transpose_head(const LaunchCtx& lc, const float* src, float* dst, int nq, int d, int V);

// ---------------------------------------------------------------- sampler
// logits: (rows_in, V); rows sampled = fs->rows; CFG mixes row 0 (cond) and row 1 (uncond).
// out_col: column of FrameScalars::out to write (row stride out_ld); noise_off: float offset into fs->noise
cudaError_t launch_sampler(const LaunchCtx& lc, const float* logits, int V, const FrameScalars* d_fs, int is_audio,
                           int out_col, int out_ld, long long noise_off, unsigned long long stream_id, int B,
                           int rows);

}  // namespace ua2
