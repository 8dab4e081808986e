/*
 * ua2_b200.h - C ABI of the B200-native (sm_100a) UniAudio2 inference hot path.
 *
 * The reference (yangdongchao/UniAudio2) is 100% Python/PyTorch and has NO FFI for this path
 * (SURVEY.md section 0 / 8b), so there is no existing binding to mirror.  Each entry point below names
 * the reference Python interface it replaces (paths relative to the reference root).  The Python host
 * side (uniaudio2_b200/llm_models/model_new.py etc.) binds these with ctypes and re-exposes the
 * reference's class / method names.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success, <0 on error; ua2_last_error() gives the message of the
 *     last failure on the calling thread;
 *   - a handle is thread-compatible (one stream at a time), not thread-safe - same contract as the
 *     reference's nn.Module with persistent KV caches (lit_model.py:814-860);
 *   - no host allocation / cudaMalloc happens after ua2_llm_setup_caches() returns.
 */
#ifndef UA2_B200_H
#define UA2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UA2_OK 0
#define UA2_ERR_INVALID (-1)
#define UA2_ERR_CUDA (-2)
#define UA2_ERR_STATE (-3)

const char* ua2_last_error(void);
/* library / device probe: returns the SM count of the current device (e.g. 148), <0 on error */
int ua2_device_sm_count(void);
const char* ua2_version(void);
/* process-wide knobs:
 * "gemv3_ctas_per_sm" (1..3, default 2), "gemv3_max_stages" (2..6, default 3), "gemv3_kcw" (floats per bulk copy, default 1024),
 * "gemv3_budget_kb" (shared memory per decode CTA, default 110), "gemv3_balance_grid" (0/1, default 1): the skinny weight-streaming linear;
 * "sgemm_min_rows" (rows from which linears use the fp32 register-tiled GEMM core when the tensor-core path is off, default 128);
 * "tc_gemm" (0/1, default 1: linears with >= "tc_min_rows" (default 32) rows run on the hand-written tcgen05 mainloop - 3xTF32,
 * fp32-class accuracy, fp32 weights read once and split on chip: csrc/ua2_umma.cu, csrc/ua2_tcgemm.cu);
 * "resblock_fused" (0/1, default 1, but the codec handle takes it only while "conv_umma" is 0: the 64-channel SEANet residual blocks as
 * one SIMT kernel, csrc/ua2_resblock.cu - measured slower than the pair of tensor-core / streaming launches, profiles/r2_kernel_rooflines.md);
 * "conv_umma" (0/1, default 1: convolutions with Cin % 16 == 0 and >= 512 output positions run as implicit GEMMs on tcgen05 straight
 * from the (B, C, T) layout - dilation, ELU prologue, PReLU epilogue, residual - and transposed convolutions of <= 64 channels as one
 * GEMM over all phases: csrc/ua2_convumma.cu); "conv_umma_staged" (0/1, default 1: its stride-1 layers stage the activations in
 * shared memory by TMA instead of gathering them into registers);
 * "conv_pointwise" (0/1, default 1: k = 1 convolutions of <= 128 input channels on the streaming SIMT kernel of csrc/ua2_sgemm.cu);
 * "conv_tc" (0/1, default 1: causal convolutions with Cin * K >= 1024 (or k = 1 with >= 256 channels) and transposed convolutions with
 * Cin >= 128 that the kernels above do not take run as im2col + the tcgen05 GEMM instead of the fp32 register-tiled core -
 * csrc/ua2_convtc.cu);
 * "attn_rows" (0/1, default 1: causal attention of many-row passes on the row-tile kernel); "flash_sbuf" (0 = by grid size, 1 / 2:
 * variant of the bf16 tensor-core attention, csrc/ua2_flash.cu);
 * "attn_ring" (0/1, default 1: KV-cache attention launches with >= 592 (row, group, 64-key split) items run on persistent CTAs that stream the
 * K / V chunks through a 3-slot bulk-copy ring instead of one fetch-compute-exit CTA per item - csrc/ua2_attn.cu);
 * "gemv3_prefetch_mb" / "gemv3_prefetch_idle_mb" (tail L2 prefetch budgets, default 0; only effective in builds with
 * -DUA2_GEMV3_TAIL_PREFETCH=1 - measured slower, see profiles/r1_l2_prefetch_experiment.md) */
int ua2_set_global_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------------
 * AR decode: llm_models/model_new.py::Model_stage3 over llm_models/lit_model.py::GPT
 * ---------------------------------------------------------------------------------------------- */

/* One litgpt GPT stack (llm_models/config.py:785-899, the Llama-3.2 entries). */
typedef struct ua2_gpt_cfg {
  int32_t n_layer;
  int32_t n_embd;
  int32_t n_head;
  int32_t n_query_groups;
  int32_t head_size;          /* 64 or 128 (32 accepted for test configs) */
  int32_t intermediate_size;
  float norm_eps;             /* config.py:38, 1e-5 */
} ua2_gpt_cfg;

/* Model_stage3.__init__ (model_new.py:340-355). */
typedef struct ua2_llm_cfg {
  ua2_gpt_cfg backbone;       /* llm_name, e.g. Llama-3.2-3B */
  ua2_gpt_cfg decoder;        /* decoder_name / local_model, e.g. Llama-3.2-300M */
  ua2_gpt_cfg understanding;  /* hard-wired 'Llama-3.2-Understanding' (model_new.py:351) */
  ua2_gpt_cfg generation;     /* hard-wired 'Llama-3.2-Generation'   (model_new.py:354) */
  int32_t text_vocab;         /* backbone padded_vocab_size (128256) */
  int32_t audio_vocab;        /* audio_semantic_vocab_size + audio_reason_vocab_size */
  int32_t num_codebooks;      /* audio_num_codebooks (8) */
  int32_t max_seq_length;     /* KV slots of the three global stacks (2048, model_new.py:560-565) */
} ua2_llm_cfg;

typedef struct ua2_llm ua2_llm;

/* Model_stage3(config) - model_new.py:340.  Allocates nothing on the device yet. */
int ua2_llm_create(const ua2_llm_cfg* cfg, ua2_llm** out);
int ua2_llm_destroy(ua2_llm* h);

/* load_state_dict / resume_for_inference (llm_utils/train_utils.py:159-177): register one fp32
 * parameter by its reference state-dict key ("backbone.transformer.h.3.attn.qkv.weight",
 * "audio_embeddings.weight", "projection.weight", "audio_head", ...).  The tensor must stay alive
 * and unchanged for the life of the handle, EXCEPT "audio_head" (num_codebooks, d, V_a), which is
 * repacked into the library's own (num_codebooks, V_a, d) layout at ua2_llm_setup_caches() time.
 * `rope_cos` / `rope_sin` tables are registered the same way under the pseudo keys
 * "<stack>.rope_cos" / "<stack>.rope_sin" with shape (max_positions, head_size) - they are the
 * output of lit_model.py:634-706 build_rope_cache, computed by the host side. */
int ua2_llm_load_weight(ua2_llm* h, const char* key, const float* dptr, const int64_t* shape, int ndim);

/* Model_stage3.setup_caches(max_batch_size) - model_new.py:554-565.  Allocates KV caches
 * (B, n_query_groups, max_seq, head_size) fp32 per layer, workspaces, and validates that every
 * parameter has been registered.  max_prefill_rows bounds the rows processed per prefill chunk. */
int ua2_llm_setup_caches(ua2_llm* h, int max_batch_size, void* stream);

/* Model_stage3.reset_caches() - model_new.py:647-651 (zero-fills every KV buffer). */
int ua2_llm_reset_caches(ua2_llm* h, void* stream);

/* Model_stage3.forward_prefix - model_new.py:456-507, KV-cache side effect only (every caller
 * discards the returned logits, e.g. evaluation/tts_task.py:244).
 *   tokens   (B, T, num_codebooks+1) int64   audio streams in cols 0..nq-1, text in col nq
 *   mask     (B, T, num_codebooks+1) uint8   already sliced to the T processed rows
 *   pos      (B, T) int64                    cache slot / RoPE position of each row
 *   max_pos  host copy of max(pos) (bounds the attention span like input_pos_maxp1); -1 = unknown */
int ua2_llm_prefill(ua2_llm* h, const int64_t* tokens, const uint8_t* mask, const int64_t* pos, int B, int T,
                    int64_t max_pos, void* stream);

/* Model_stage3.generate_frame - model_new.py:568-645.
 *   tokens (B,1,nq+1) int64, mask (B,1,nq+1) uint8, input_pos scalar (the reference passes a
 *   1-element tensor shared by all rows, tts_task.py:245), temperature > 0, topk >= 1,
 *   forbid_prefix >= 0, cfg_scale (CFG active iff cfg_scale > 1 and B > 1, model_new.py:618).
 *   noise: nullable.  If non-null: Exp(1) draws, laid out [R x text_vocab | nq x (R x audio_vocab)]
 *   with R = sampled rows (B, or 1 under CFG) - the draws the reference would take from
 *   torch.empty_like(probs).exponential_(1) (model_new.py:141-143).  If null the library draws
 *   Exp(1) itself from Philox4x32-10 keyed by (seed, frame counter).
 *   out (B, 1+nq) int32: col 0 text token, cols 1.. audio tokens (merged id space).
 * Error behaviour mirrors model_new.py:165-180 (ValueError cases -> UA2_ERR_INVALID). */
int ua2_llm_generate_frame(ua2_llm* h, const int64_t* tokens, const uint8_t* mask, int B, int64_t input_pos,
                           float temperature, int topk, int forbid_prefix, float cfg_scale, const float* noise,
                           uint64_t seed, int32_t* out, void* stream);

/* The hot loop of Generator.generate_tts (evaluation/tts_task.py:253-279, B = 1) with its phase / EOS state machine ON THE DEVICE:
 * `n_frames` consecutive frames from input_pos, each fed the previous frame's sample (audio tokens -> columns 0..nq-1, text token ->
 * column nq, audio mask), no host round trip in between.  tokens0 / mask0 (1,1,nq+1): the prompt's last row for the first frame of an
 * utterance, NULL to continue from the previous sample.  After every frame the sampled row is tested like the reference does:
 * all audio tokens == end_tok -> done (nothing recorded, later frames of the call are ignored); otherwise the row is appended to
 * frames_out (frames_cap x (1+nq) int32, device); all == reason_eos (or, with fixed_switch >= 0, the fixed_switch-th recorded frame)
 * -> the following frames run with forbid_prefix = reason_card.
 * state (device int32[4], zero it before an utterance): [0] forbid_prefix of the next frame, [1] done, [2] frames recorded,
 * [3] 1-based index of the frame that switched the phase.  noise: NULL (Philox) or n_frames blocks of noise_stride floats laid out as
 * for ua2_llm_generate_frame. */
int ua2_llm_tts_frames(ua2_llm* h, const int64_t* tokens0, const uint8_t* mask0, int64_t input_pos, int n_frames, float temperature, int topk,
                       const float* noise, int64_t noise_stride, uint64_t seed, int reason_eos, int end_tok, int reason_card, int fixed_switch,
                       int32_t* state, int32_t* frames_out, int frames_cap, void* stream);

/* Introspection for parity tests: device pointers of internal buffers (valid until destroy).
 * which: 0 backbone, 1 decoder, 2 understanding, 3 generation. */
int ua2_llm_get_kv(ua2_llm* h, int which, int layer, float** k, float** v);
/* name: "h_final" (B x n_embd), "text_logits" (B x text_vocab), "audio_logits" (nq x B x audio_vocab) */
int ua2_llm_get_buffer(ua2_llm* h, const char* name, float** ptr, int64_t* numel);
/* knobs: "graph" (0/1, default 1: replay the frame as a CUDA graph), "pdl" (0/1, default 1: programmatic dependent launch),
 * "chain" (0/1, default 0: B = 1 frames run as persistent multi-op cooperative kernels, see csrc/ua2_chain.cu),
 * "attn_direct" (0/1, default 0: the local decoder's <= 8-key attention runs inside the proj kernel's prologue; measured 0.3 % slower) */
int ua2_llm_set_option(ua2_llm* h, const char* name, int value);
/* number of kernels launched (or graph kernel nodes replayed) by the last prefill / generate_frame */
int ua2_llm_last_launch_count(ua2_llm* h);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone operators (unit-parity surface; the handle API above is built from these kernels)
 * ---------------------------------------------------------------------------------------------- */

/* y[m, n] = sum_k f(x)[m,k] * W[n,k]    (lit_model.py:424/511/592-595 F.linear, bias-free)
 *   norm_w != NULL : f = RMSNorm(x; norm_w, eps) (lit_model.py:883-890) fused in front
 *   residual != NULL : y += residual (Block.forward residual adds, lit_model.py:344-349)        */
int ua2_linear_f32(const float* x, const float* W, const float* norm_w, float eps, const float* residual, float* y,
                   int M, int N, int K, void* stream);
/* The same linear (W2 == NULL; optional RMSNorm prologue and residual) or SwiGLU pair (W2 != NULL) forced onto the tensor-core
 * path for any M: tcgen05 3xTF32 with the fp32 weights split on chip (csrc/ua2_umma.cu) - what forward_prefix
 * (model_new.py:456-507) and batched frames use for M >= tc_min_rows.  Scratch is owned by the library. */
int ua2_tc_linear_f32(const float* x, const float* W, const float* W2, const float* norm_w, float eps, const float* residual, float* y,
                      int M, int N, int K, void* stream);
/* Unmasked softmax(q k^T / sqrt(hs)) v on tensor cores, bf16 operands, fp32 scores / statistics / output: the attention of the
 * flow-matching decoder's blocks (ReasoningCodec_film/models/attention.py:308-357 -> diffusers Attention -> SDPA, which the reference
 * runs in bf16 under reason_tokenizer.py:265's autocast).  q16, k16, v16: (B, H, T, 64) bf16; out: (B, T, H * 64) fp32.
 * csrc/ua2_flash.cu: tcgen05 kind::f16 for both contractions, P from tensor memory.  UA2_ERR_INVALID unless hs == 64. */
int ua2_flash_attn_bf16(const void* q16, const void* k16, const void* v16, float* out, int B, int T, int H, int hs, void* stream);
/* One application of the TTS loop's break / phase-switch rules (evaluation/tts_task.py:259-271) to a sampled row (1+nq int32, device):
 * the state machine that ua2_llm_tts_frames runs after every frame. */
int ua2_tts_state_step(const int32_t* sample, int nq, int32_t* state, int32_t* frames_out, int frames_cap, int reason_eos, int end_tok,
                       int reason_card, int fixed_switch, void* stream);
/* y[m, n] = silu(sum_k f(x) W1[n,k]) * (sum_k f(x) W2[n,k])   (LLaMAMLP fc_1/fc_2, lit_model.py:591-594) */
int ua2_swiglu_f32(const float* x, const float* W1, const float* W2, const float* norm_w, float eps, float* y, int M,
                   int N, int K, void* stream);
/* fused RMSNorm -> QKV linear -> half-split RoPE -> KV-cache append (lit_model.py:424-467).
 *   pos (M) int32 cache slot per row, bidx (M) int32 batch row in the cache;
 *   q_out (M, n_head*hs); k_cache/v_cache (B, G, S_max, hs); cos/sin (>=max pos, hs). */
int ua2_qkv_rope_f32(const float* x, const float* Wqkv, const float* norm_w, float eps, const int32_t* pos,
                     const int32_t* bidx, const float* cos, const float* sin, float* q_out, float* k_cache,
                     float* v_cache, int M, int K, int n_head, int n_groups, int hs, int S_max, void* stream);
/* causal GQA attention of M query rows against the cache: softmax(q k^T / sqrt(hs)) v over slots
 * 0..pos[m] (lit_model.py:468-532).  y (M, n_head*hs).  workspace: see ua2_attn_workspace_floats. */
int ua2_attn_f32(const float* q, const float* k_cache, const float* v_cache, const int32_t* pos, const int32_t* bidx,
                 float* y, float* workspace, int M, int n_head, int n_groups, int hs, int S_max, void* stream);
int64_t ua2_attn_workspace_floats(int M, int n_head, int hs, int S_max);
/* sample_topk / audio_sample_topk (model_new.py:146-187) on R rows of V logits.
 *   cfg_scale > 1: logits has 2R rows (cond rows first R? no: row 0 = cond, row 1 = uncond, R must be 1)
 *   noise nullable (R x V Exp(1) draws); out (R) int32. */
int ua2_sample_topk_f32(const float* logits, int R, int V, float temperature, int topk, int forbid_prefix,
                        float cfg_scale, const float* noise, uint64_t seed, uint64_t offset, int32_t* out,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Codec operators: SEANet causal convolutions and residual VQ (tools/tokenizer/MimiCodec/model/*, the
 * importable twin of llm_modules/{seanet,conv,resample}.py + quantization/{core_vq,vq}.py)
 * ---------------------------------------------------------------------------------------------- */

/* StreamingConv1d.forward, non-streaming causal branch (modules/conv.py:232-254): left pad k_eff - stride, right
 * "extra" pad to a full last window, T_out = ceil(T_in / stride).  w_ckc: weights repacked (Cin, K, Cout).
 * pre_elu: apply the nn.ELU that SEANet places in front of the conv (modules/seanet.py:52-66, :196, :214).
 * residual != NULL: y = residual + conv (SEANetResnetBlock skip, modules/seanet.py:92-94).
 * replicate_pad: pad_mode 'replicate' (ConvDownsample1d, modules/resample.py:48) instead of zeros. */
int ua2_conv1d_causal_f32(const float* x, const float* w_ckc, const float* bias, const float* residual, float* y, int B,
                          int Cin, int Cout, int T_in, int K, int stride, int dilation, int pre_elu, int replicate_pad,
                          void* stream);
/* Same operator as ua2_conv1d_causal_f32 computed as an implicit GEMM (128 x 128 register tiles); weights stay in the
 * reference's (Cout, Cin, K) layout.  This is what the codec handle calls. */
int ua2_conv1d_causal_gemm_f32(const float* x, const float* w_torch, const float* bias, const float* residual, float* y, int B,
                               int Cin, int Cout, int T_in, int K, int stride, int dilation, int pre_elu, int replicate_pad,
                               void* stream);
/* SEANetResnetBlock.forward (modules/seanet.py:21-94, dilation 1, true_skip) as ONE kernel that keeps the hidden activation on
 * chip: y = x + conv_k1(ELU(conv_k3(ELU(x)))).  w1 (H, C, 3), w2 (C, H, 1) in torch's Conv1d layout; served for C = 64, H = 32
 * (the blocks that run at 24 kHz).  The codec handle uses it when the global option "resblock_fused" is 1 and "conv_umma"
 * is 0 (measured on a B200 at 4.1 ms for batch 16 x 10 s against 0.88 + 0.83 ms for the two default launches it would replace). */
int ua2_resblock_f32(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y, int B, int C, int H,
                     int T, void* stream);
/* StreamingConvTranspose1d.forward, causal, trim_right_ratio = 1 (modules/conv.py:306-329): kernel = 2*stride,
 * T_out = T_in * stride.  w_ckc: torch's (Cin, Cout, K) weights repacked to (Cin, K, Cout). */
int ua2_convtr1d_causal_f32(const float* x, const float* w_ckc, const float* bias, float* y, int B, int Cin, int Cout,
                            int T_in, int stride, int pre_elu, void* stream);
/* Transposed conv as `stride` phase GEMMs (ua2_sgemm.cu).  w_phase (stride, Cout, Cin, 2) comes from
 * ua2_convtr1d_repack_phase_f32(torch weight (Cin, Cout, 2*stride)). */
int ua2_convtr1d_repack_phase_f32(const float* w_torch, float* w_phase, int Cin, int Cout, int stride, void* stream);
int ua2_convtr1d_causal_gemm_f32(const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin, int Cout,
                                 int T_in, int stride, int pre_elu, void* stream);
/* General forms (explicit left / right zero padding, optional single-slope PReLU after the bias, optional residual add):
 * building blocks of ScalarModel (tools/tokenizer/ReasoningCodec_film/models/scalar24k.py: Conv1d :30-69, ResidualUnit
 * :139-150, UpsampleLayer :232-265) and of the non-overlapping strided convs of AudioDiffusion1D.py:244-251.
 * ua2_convtr1d_f32: kernel = 2*stride; y[t] = full[t + crop_left], t < T_out, full length (T_in + 1) * stride. */
int ua2_conv1d_f32(const float* x, const float* w_torch, const float* bias, const float* prelu_slope, const float* residual, float* y,
                   int B, int Cin, int Cout, int T_in, int K, int stride, int dilation, int pad_left, int pad_right, void* stream);
int ua2_convtr1d_f32(const float* x, const float* w_phase, const float* bias, const float* prelu_slope, float* y, int B, int Cin,
                     int Cout, int T_in, int stride, int crop_left, int T_out, void* stream);
/* op 0: round(param * x) / param (round_func9, scalar24k.py:279-288);  op 1: tanh(x) */
int ua2_elementwise_f32(const float* x, float* y, long long n, int op, float param, void* stream);
/* time_film (ReasoningCodec_film/models/AudioDiffusion1D.py:428-438): out = gamma * features + beta with
 * gamma = 1 + gamma_scale * tanh(params[..., :C]), beta = params[..., C:]; batches with zero_mask[b] != 0 get (1, 0).
 * params (B, T, 2C), features / out (B, T, C); zero_mask (B) uint8 nullable - the reference's torch.rand(B,1,1) < 0.2 draw. */
int ua2_film_f32(const float* params, const float* features, const uint8_t* zero_mask, float* out, int B, int T, int C, float gamma_scale,
                 void* stream);
/* F.interpolate(x (B, C, T_in), scale_factor, mode='nearest') (AudioDiffusion1D.py:523, :589); T_out = floor(T_in * scale) */
int ua2_interp_nearest_f32(const float* x, float* y, int B, int C, int T_in, int T_out, float scale_factor, void* stream);
/* nn.Linear with bias (cond_fusion_layer_*, cond_feature_emb: AudioDiffusion1D.py:278-280, :547) */
int ua2_linear_bias_f32(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, void* stream);
/* ConvTrUpsample1d(learnt, channel_wise) (modules/resample.py:68-119): depthwise, w (C, 1, 2*stride). */
int ua2_convtr1d_depthwise_f32(const float* x, const float* w, float* y, int B, int C, int T_in, int stride, void* stream);
/* ResidualVectorQuantization.encode (quantization/core_vq.py:365-376) on an already projected input x (B, D, T):
 * for q in 0..n_q-1: code = argmin_j ||r - emb[q][j]||  (EuclideanCodebook._quantize :179-185); r -= emb[q][code].
 * emb (n_q, K, D) = embedding_sum / clamp(cluster_usage, eps) (:142-150); emb_sqnorm (n_q, K) = row squared norms.
 * Writes codes[b, q_off + q, t] (int64) of a (B, n_q_total, T) tensor. */
int ua2_rvq_encode_f32(const float* x, const float* emb, const float* emb_sqnorm, int64_t* codes, int B, int D, int T, int K,
                       int n_q, int n_q_total, int q_off, void* stream);
/* Same algorithm for many frames: per quantizer a tiled fp32 GEMM (scores) + an argmin / residual-update kernel.
 * r_md: projected input in frame-major layout (B*T, D), updated in place; S: (B*T, K) scratch. */
int ua2_rvq_encode_gemm_f32(float* r_md, const float* emb, const float* emb_sqnorm, float* S, int64_t* codes, int B, int D, int T,
                            int K, int n_q, int n_q_total, int q_off, void* stream);
/* ResidualVectorQuantization.decode (core_vq.py:378-384): out (B, D, T) = sum_q emb[q][codes[b, q_off + q, t]]. */
int ua2_rvq_decode_f32(const int64_t* codes, const float* emb, float* out, int B, int D, int T, int K, int n_q, int n_q_total,
                       int q_off, void* stream);

/* ---- codec handle: tools/tokenizer/MimiCodec/model/models/MimiCodec.py::MimiCodec (encode :93-101, decode :103-110) ---- */
typedef struct ua2_codec_cfg {
  int32_t n_filters;        /* SEANet base width (64) */
  int32_t ratios[8];        /* encoder_rates in DECODER order, e.g. {8,6,5,4}; the encoder walks them reversed */
  int32_t n_ratios;
  int32_t latent_dim;       /* 512 */
  int32_t codebook_size;    /* 2048 */
  int32_t codebook_dim;     /* 256 */
  int32_t rvq_layers;       /* 32 = 1 semantic + 31 acoustic quantizers (SplitResidualVectorQuantizer) */
  int32_t num_heads;        /* 8 */
  int32_t num_layers;       /* 8 */
  int32_t context;          /* 250: causal attention window of the Moshi-family transformer */
  int32_t dim_feedforward;  /* 2048 */
  int32_t resample_stride;  /* int(encoder_frame_rate / target_frame_rate) = 2 (MimiCodec.py:66-67) */
  float max_period;         /* RoPE max period, 10000 */
} ua2_codec_cfg;
typedef struct ua2_codec ua2_codec;

int ua2_codec_create(const ua2_codec_cfg* cfg, ua2_codec** out);
int ua2_codec_destroy(ua2_codec* h);
/* register a parameter / buffer by its reference state-dict key ("encoder.model.3.conv.conv.weight",
 * "quantizer.rvq_rest.vq.layers.4._codebook.embedding_sum", ...); tensors must outlive the handle */
int ua2_codec_load_weight(ua2_codec* h, const char* key, const float* dptr, const int64_t* shape, int ndim);
/* repack conv weights to (Cin, K, Cout), materialise codebooks (embedding_sum / clamp(cluster_usage, eps)) */
int ua2_codec_finalize(ua2_codec* h, void* stream);
/* number of code frames for T samples: ceil through every strided stage */
int64_t ua2_codec_frames(ua2_codec* h, int64_t T_samples);
/* MimiCodec.encode: wav (B, 1, T) fp32 -> codes (B, rvq_layers, frames) int64 */
int ua2_codec_encode(ua2_codec* h, const float* wav, int B, int T, int64_t* codes, void* stream);
/* MimiCodec.decode: codes (B, rvq_layers, Tq) int64 -> wav (B, 1, Tq * resample_stride * hop_length) fp32 */
int ua2_codec_decode(ua2_codec* h, const int64_t* codes, int B, int Tq, float* wav, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Moshi-family streaming transformer and sampler: llm_modules/transformer.py (StreamingTransformer,
 * StreamingTransformerLayer, StreamingMultiheadAttention, RingKVCache, multi_linear, _rms_norm), llm_modules/gating.py
 * (ActivationGating), llm_modules/rope.py (apply_rope) and llm_utils/sampling.py (sample_token) - the modules
 * BASELINE.json's north_star names for the AR decode.
 * ---------------------------------------------------------------------------------------------- */

/* Position held by ring slot j after `end_offset` keys were written (RingKVCache.complete, transformer.py:254-276):
 *   j >= end_offset -> -1 (never written); delta = j - end_offset % cap; delta <= 0 ? end_offset + delta
 *   : end_offset + delta - cap.  (The reference's `<=` makes the slot at end_offset % cap report a future position once
 *   the ring has wrapped, so the oldest key is never visible - reproduced as is.)  ring = 0: linear cache, position = j.
 * StreamingMultiheadAttention.forward, transformer.py:375-419, after the in_proj: interleaved-pair RoPE of q and k
 * (rope.py:40-58; angle = freqs[i] * pos, freqs (hs/2) = exp(i * -ln(max_period) * 2 / hs) supplied by the host, NULL =
 * no RoPE) and the append of k / v to the cache (B, H, cap, hs) at slot pos % cap (ring) or pos (linear).
 *   qkv (M rows of 3*H*hs, row stride ld_qkv), ordered (p h d) as `rearrange(projected, "b t (p h d) -> p b h t d")`;
 *   pos (M) int32 absolute position of each row, bidx (M) int32 batch row; q_out (M, H*hs). */
int ua2_rope_ring_append_f32(const float* qkv, int ld_qkv, const int32_t* pos, const int32_t* bidx, const float* freqs,
                             float* q_out, float* k_cache, float* v_cache, int M, int H, int hs, int cap, int ring,
                             void* stream);
/* F.scaled_dot_product_attention with the streaming mask of transformer.py:399-410: key slot j of batch row bidx[m] is
 * visible to row m iff its position p_j >= 0 and (causal == 0 or (0 <= pos[m] - p_j and (context <= 0 or pos[m] - p_j <
 * context))).  end_offset = keys written so far INCLUDING this call's.  y (M, H*hs). */
int ua2_ring_attn_f32(const float* q, const float* k_cache, const float* v_cache, const int32_t* pos, const int32_t* bidx,
                      float* y, int M, int H, int hs, int cap, int64_t end_offset, int ring, int causal, int context,
                      void* stream);

/* sample_token (llm_utils/sampling.py:84-105) and sample_token_audio (:107-130) on R rows of V logits -> out (R) int64.
 *   use_sampling == 0 or temp <= 0: argmax(logits) (first maximum).  Otherwise probs = softmax(logits / temp), ids >=
 *   end_token get probability -inf when end_token >= 0, and
 *     top_p > 0 : sort descending, keep ranks whose exclusive cumulative sum is <= top_p, renormalise      (:64-81; V <= 4096)
 *     top_k > 0 : the k largest probabilities in descending order                                         (:49-61; k <= 1024)
 *     else      : all V probabilities                                                                     (:15-46)
 *   then token = argmax(p / q) with q ~ Exp(1) (multinomial without replacement, :41-43), first maximum wins.
 *   noise: the Exp(1) draws, (R, n) with n = top_k for the top-k branch (one draw per RANK, as torch.topk orders them)
 *   and n = V otherwise; NULL = drawn by the library from Philox4x32-10 keyed by (seed, offset). */
int ua2_sample_token_f32(const float* logits, int R, int V, int use_sampling, float temp, int top_k, float top_p,
                         int end_token, const float* noise, uint64_t seed, uint64_t offset, int64_t* out, void* stream);

/* StreamingTransformer(d_model, num_heads, num_layers, dim_feedforward, causal, context, positional_embedding, max_period,
 * positional_scale, norm, layer_scale, gating, weights_per_step) - transformer.py:616-669, :449-543. */
#define UA2_STX_MAX_STEPS 64
typedef struct ua2_stx_cfg {
  int32_t d_model;
  int32_t num_heads;          /* head size d_model / num_heads must be 32, 64 or 128 */
  int32_t num_layers;
  int32_t causal;
  int32_t context;            /* 0 = None */
  int32_t positional_embedding; /* 0 none, 1 sin, 2 rope, 3 sin_rope */
  int32_t norm;               /* 0 layer_norm (eps 1e-5), 1 layer_norm_f32 (1e-8), 2 rms_norm (1e-5), 3 rms_norm_f32 (1e-8) */
  int32_t gating;             /* 0 none: linear2(gelu(linear1 x));  1 silu: ActivationGating(F.silu), gating.py:24-51 */
  int32_t weights_per_step;   /* 0, or the number of per-step weight slabs (multi_linear, transformer.py:155-179) */
  int32_t layer_scale;        /* 0 / 1 (LayerScale parameters present) */
  int32_t dim_feedforward[UA2_STX_MAX_STEPS]; /* [0] unless weights_per_step > 0 with a per-step list */
  float max_period;
  float positional_scale;
} ua2_stx_cfg;
typedef struct ua2_stx ua2_stx;

int ua2_stx_create(const ua2_stx_cfg* cfg, ua2_stx** out);
int ua2_stx_destroy(ua2_stx* h);
/* one fp32 parameter by its reference state-dict key ("layers.0.self_attn.in_proj_weight", "layers.0.norm1.alpha",
 * "layers.1.gating.3.linear_in.weight", "layers.0.layer_scale_1.scale", ...) plus the host-computed tables "rope_freqs"
 * (hs/2; rope.py:37-38) and "sin_denoms" (d_model/2; max_period ** (i / (half - 1)), transformer.py:150); tensors must
 * outlive the handle */
int ua2_stx_load_weight(ua2_stx* h, const char* key, const float* dptr, const int64_t* shape, int ndim);
/* validates that every parameter of the configuration has been registered */
int ua2_stx_finalize(ua2_stx* h);
/* StreamingModule._start_streaming(batch_size) (llm_modules/streaming.py:86-91): allocates one zeroed ring
 * (batch, H, capacity, hs) x {k, v} per layer, capacity = context, or weights_per_step when context is None
 * (transformer.py:337-346); offsets = 0.  _stop_streaming / reset_streaming (streaming.py:93-126; reset of a module
 * that is not streaming -> UA2_ERR_INVALID like the reference's ValueError). */
int ua2_stx_start_streaming(ua2_stx* h, int batch_size, void* stream);
int ua2_stx_stop_streaming(ua2_stx* h);
int ua2_stx_reset_streaming(ua2_stx* h);
/* StreamingTransformer.forward (transformer.py:671-692): x (B, T, d_model) -> y (B, T, d_model); y may alias x.
 * Streaming: B must equal the streaming batch size, T <= capacity, and offset + T <= weights_per_step when per-step
 * weights are used (the reference indexes past the weight slab otherwise). */
int ua2_stx_forward(ua2_stx* h, const float* x, float* y, int B, int T, void* stream);
/* knobs: "graph" (0/1, default 1: streaming calls replay the layer stack as a CUDA graph per (batch, T, weight step)),
 * "pdl" (0/1, default 1: programmatic dependent launch between the kernels of a streaming call) */
int ua2_stx_set_option(ua2_stx* h, const char* name, int value);
/* kernels launched (or graph kernel nodes replayed) by the last forward */
int ua2_stx_last_launch_count(ua2_stx* h);
/* introspection: ring buffers (batch, H, capacity, hs) of a layer and the number of keys written so far */
int ua2_stx_get_kv(ua2_stx* h, int layer, float** k, float** v, int64_t* end_offset, int* capacity);

/* ------------------------------------------------------------------------------------------------
 * Flow-matching decoder of ReasoningCodec_film: the DiT estimator
 * tools/tokenizer/ReasoningCodec_film/models/transformer_1d_flow.py::Transformer1DModel (adaLN-single blocks of
 * models/attention.py::BasicTransformerBlock) and models/AudioDiffusion1D.py::BASECFM.solve_euler (:89-129).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ua2_dit_cfg {     /* models/model_config.json */
  int32_t num_attention_heads;   /* 24 */
  int32_t attention_head_dim;    /* 64 (32 / 64 / 128 served) */
  int32_t in_channels;           /* 1040 = latent 136 + in-context latent 136 + condition 768 */
  int32_t out_channels;          /* 136 */
  int32_t num_layers;            /* 32 */
  int32_t num_positional_embeddings; /* rows of pos_embed.pe (3000, transformer_1d_flow.py:195) */
  int32_t flow_t_size;           /* 512 (transformer_1d_flow.py:47) */
  float norm_eps;                /* 1e-6 */
} ua2_dit_cfg;
typedef struct ua2_dit ua2_dit;

int ua2_dit_create(const ua2_dit_cfg* cfg, ua2_dit** out);
int ua2_dit_destroy(ua2_dit* h);
/* one fp32 parameter / buffer by its reference state-dict key ("proj_in.ffn_1.weight", "pos_embed.pe",
 * "transformer_blocks.3.attn1.to_q.bias", "adaln_single.linear.weight", ...) plus the host-evaluated table "tfreqs"
 * (flow_t_size / 2: exp(-ln(10000) * i / half), transformer_1d_flow.py:67); tensors must outlive the handle */
int ua2_dit_load_weight(ua2_dit* h, const char* key, const float* dptr, const int64_t* shape, int ndim);
/* validates the parameter set; repacks the k = 3 convolutions of ProjectLayer to GEMM form and concatenates to_q/k/v */
int ua2_dit_finalize(ua2_dit* h, void* stream);
/* Transformer1DModel.forward(hidden_states (B, T, in_channels), timestep (B) fp32 on the device).sample -> (B, T, out_channels) */
int ua2_dit_forward(ua2_dit* h, const float* hidden_states, const float* timestep, float* out, int B, int T, void* stream);
/* BASECFM.solve_euler for batch 1 with classifier-free guidance (the reference's only working configuration: it repeats the
 * timestep twice, AudioDiffusion1D.py:114): x (1, T, out_channels) noise in, solution out (in place);
 * incontext_x (1, T, out_channels); mu (1, T, in_channels - 2 * out_channels); t_span: HOST array of n_span time points
 * (torch.linspace(0, 1, steps + 1)); guidance_scale > 1; sigma_min = 1e-4 (:68). */
int ua2_dit_solve_euler(ua2_dit* h, float* x, const float* incontext_x, int incontext_length, const float* t_span, int n_span,
                        const float* mu, int T, float guidance_scale, float sigma_min, void* stream);
/* knobs: "bf16" (0/1, default 0): linears of >= 32 rows run on bf16 operands with fp32 accumulation (tcgen05 kind::f16) - the
 * arithmetic of the reference, which calls this model under torch.autocast(bfloat16) (reason_tokenizer.py:265); the default
 * keeps fp32-class accuracy (3xTF32).  A bf16 copy of every weight is kept (2 B per parameter). */
int ua2_dit_set_option(ua2_dit* h, const char* name, int value);
int ua2_dit_last_launch_count(ua2_dit* h);

/* ----------------------------------------------------------------------------------------------
 * Whisper encoder: the first SSL front-end of ReasoningCodec_film's tokenize.  Replaces
 * `WhisperModel.from_pretrained(path).encoder(mels, return_dict=True).last_hidden_state` as called from
 * models/AudioDiffusion1D.py:223, :334-343 - the reference's own copy of the model code is
 * models/modeling_whisper.py: WhisperEncoder.forward :766-867, WhisperEncoderLayer.forward :394-443, WhisperAttention.forward :255-374.
 * csrc/ua2_enc.cu.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ua2_whisper_cfg {     /* WhisperConfig fields the encoder reads (whisper-medium: 1024 / 16 / 4096 / 24 / 1500 / 80) */
  int32_t d_model;                   /* multiple of 8, <= 4096; d_model / heads in {32, 64, 128} (64 for the bf16 mode) */
  int32_t encoder_attention_heads;
  int32_t encoder_ffn_dim;
  int32_t encoder_layers;
  int32_t max_source_positions;      /* output frames per clip; the input has exactly twice as many mel frames (:808-811) */
  int32_t num_mel_bins;
} ua2_whisper_cfg;
typedef struct ua2_whisper ua2_whisper;

int ua2_whisper_create(const ua2_whisper_cfg* cfg, ua2_whisper** out);
int ua2_whisper_destroy(ua2_whisper* h);
/* one fp32 parameter by its state-dict key relative to the encoder ("conv1.weight", "embed_positions.weight",
 * "layers.3.self_attn.k_proj.weight", "layers.3.final_layer_norm.bias", "layer_norm.weight", ...); tensors must outlive the handle */
int ua2_whisper_load_weight(ua2_whisper* h, const char* key, const float* dptr, const int64_t* shape, int ndim);
/* validates the parameter set; repacks the two k = 3 convolutions to GEMM form and concatenates q / k / v projections */
int ua2_whisper_finalize(ua2_whisper* h, void* stream);
/* input_features (B, num_mel_bins, 2 * max_source_positions) fp32 -> last_hidden_state (B, max_source_positions, d_model) fp32 */
int ua2_whisper_forward(ua2_whisper* h, const float* input_features, float* out, int B, void* stream);
/* "bf16" (0/1, default 0): the reference's autocast arithmetic (reason_tokenizer.py:114-118) - linears on bf16 operands with fp32
 * accumulation (tcgen05 kind::f16), attention on tensor cores (csrc/ua2_flash.cu); default: fp32 class (3xTF32 linears, fp32 attention) */
int ua2_whisper_set_option(ua2_whisper* h, const char* name, int value);
int ua2_whisper_last_launch_count(ua2_whisper* h);

/* ----------------------------------------------------------------------------------------------
 * Waveform front-end of ReasoningCodec_film's tokenize (csrc/ua2_frontend.cu).  The reference leaves the
 * device here: reason_tokenizer.py:67-72 get_whisper_features = torchaudio Resample(24000, 16000) -> .cpu().numpy() ->
 * transformers WhisperFeatureExtractor (torch.stft on the host) -> .to(device); AudioDiffusion1D.py:363-365 resamples again for WavLM.
 * ---------------------------------------------------------------------------------------------- */
/* torchaudio.transforms.Resample (sinc_interp_hann): y[b, new * m + p] = sum_k kernel[p, k] * x[b, orig * m + k - width], zero outside
 * the clip; kernel (newf, 2 * width + orig) fp32 as torchaudio builds it, orig / newf already divided by their gcd.  Positions
 * [n_valid, n_store) of every output row are written as zeros (padding to 30 s; the 160 zeros appended for WavLM).
 * n_valid <= ceil(newf * L / orig). */
int ua2_resample_f32(const float* x, long long ldx, const float* kernel, float* y, long long ldy, int B, int L, int n_valid, int n_store, int orig,
                     int newf, int width, void* stream);
/* WhisperFeatureExtractor._torch_extract_fbank_features: wav16 (B, L) fp32 (already padded / cut to 30 s) -> out (B, n_mels, n_frames)
 * = (max(log10(max(mel_filters^T |stft|^2, 1e-10)), clip max - 8) + 4) / 4 with stft = torch.stft(n_fft, hop, window, center, reflect);
 * window (n_fft), mel_filters (n_fft / 2 + 1, n_mels) fp32 device arrays.  n_fft <= 512 and even; n_frames <= 1 + L / hop (the
 * reference drops the last of those frames: pass L / hop). */
int ua2_whisper_logmel_f32(const float* wav16, long long ld, const float* window, const float* mel_filters, float* out, int B, int L, int n_fft,
                           int hop, int n_mels, int n_frames, void* stream);

/* ----------------------------------------------------------------------------------------------
 * WavLM encoder: the second SSL front-end of tokenize.  Replaces `AutoModel.from_pretrained(wav_lm_path)(wav_16k,
 * output_hidden_states=True).hidden_states` + the [6:10] mean of models/AudioDiffusion1D.py:226, :359-370 (transformers WavLMModel:
 * modeling_wavlm.py WavLMFeatureEncoder / WavLMFeatureProjection / WavLMPositionalConvEmbedding / WavLMEncoder / WavLMAttention).
 * Served: feat_extract_norm "group", do_stable_layer_norm false (wavlm-base, wavlm-base-plus), no attention mask, eval.
 * csrc/ua2_wavlm.cu.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ua2_wavlm_cfg {       /* WavLMConfig fields (base-plus: 768 / 12 / 3072 / 12 / 7 / 512.. / 10,3,3,3,3,2,2 / 5,2,2,2,2,2,2 / 0 / 128 / 16 / 320 / 800 / 1e-5) */
  int32_t hidden_size;               /* multiple of 8, <= 4096; hidden_size / heads in {32, 64, 128} */
  int32_t num_attention_heads;
  int32_t intermediate_size;
  int32_t num_hidden_layers;
  int32_t num_feat_extract_layers;   /* 2 .. 8 */
  int32_t conv_dim[8];               /* multiples of 8 */
  int32_t conv_kernel[8];            /* conv_kernel[0] <= 16 */
  int32_t conv_stride[8];            /* conv_stride[0] <= 8 */
  int32_t conv_bias;
  int32_t num_conv_pos_embeddings;   /* even, <= 128 */
  int32_t num_conv_pos_embedding_groups; /* hidden_size / groups <= 48 */
  int32_t num_buckets;
  int32_t max_bucket_distance;
  float layer_norm_eps;
} ua2_wavlm_cfg;
typedef struct ua2_wavlm ua2_wavlm;

int ua2_wavlm_create(const ua2_wavlm_cfg* cfg, ua2_wavlm** out);
int ua2_wavlm_destroy(ua2_wavlm* h);
/* one fp32 parameter by its WavLMModel state-dict key ("feature_extractor.conv_layers.2.conv.weight", "feature_projection.projection.bias",
 * "encoder.layers.0.attention.rel_attn_embed.weight", "encoder.layers.4.attention.gru_rel_pos_const", ...).  The positional convolution
 * is given with its weight normalisation applied: "encoder.pos_conv_embed.conv.weight" (hidden, hidden / groups, kernel) =
 * g * v / ||v|| over dims (0, 1) of ...parametrizations.weight.original0 / original1.  Tensors must outlive the handle. */
int ua2_wavlm_load_weight(ua2_wavlm* h, const char* key, const float* dptr, const int64_t* shape, int ndim);
/* validates the parameter set; repacks the strided convolutions to GEMM form, the positional convolution to per-tap slabs, and
 * concatenates q / k / v projections */
int ua2_wavlm_finalize(ua2_wavlm* h, void* stream);
/* encoder frames for a clip of L samples (0: shorter than the receptive field) */
long long ua2_wavlm_frames(ua2_wavlm* h, long long L);
/* wav16 (B, L) fp32, row stride ld -> out (B, T, hidden) = mean of hidden_states[hs_lo:hs_hi] (hidden_states[0] = the encoder's input
 * after its LayerNorm, [i] = output of layer i - 1; only the first hs_hi - 1 layers run).  all_hidden (optional): every hidden state
 * [0, hs_hi) as (hs_hi, B, T, hidden). */
int ua2_wavlm_forward(ua2_wavlm* h, const float* wav16, long long ld, int B, int L, int hs_lo, int hs_hi, float* out, float* all_hidden,
                      void* stream);
/* "bf16" (0/1, default 0): the reference's autocast arithmetic (reason_tokenizer.py:114-118) - GEMMs on bf16 operands with fp32
 * accumulation (tcgen05 kind::f16), attention on tensor cores with the position bias added in the softmax (csrc/ua2_flash.cu, head
 * size 64); default: fp32 class (3xTF32 GEMMs, fp32 attention) */
int ua2_wavlm_set_option(ua2_wavlm* h, const char* name, int value);
int ua2_wavlm_last_launch_count(ua2_wavlm* h);
/* host-only: WavLMAttention._relative_positions_bucket for relative positions -(T - 1) .. T - 1 -> out_host[2 T - 1] */
int ua2_wavlm_rel_bucket_table(int T, int num_buckets, int max_distance, int32_t* out_host);
/* the encoder's own kernels one at a time, for operator-level parity tests (op 0 positional convolution, 1 gate, 2 biased attention,
 * 3 biased attention on tensor cores from bf16 q / k / v;
 * argument meaning per op in csrc/ua2_wavlm.cu) */
int ua2_wavlm_ops_f32(int op, const float* a, const float* b, const float* c, const float* d, float* y, int i0, int i1, int i2, int i3, int i4,
                      void* stream);

/* ----------------------------------------------------------------------------------------------
 * Reasoning encoder (`AudioThinking`) of tokenize, up to the query tokens that go into reasoning_vq.  Replaces
 * AudioDiffusion1D.encode_reasoning_part (models/AudioDiffusion1D.py:372-390: down_sampling_layer_whisper, concatenation with the
 * BEST-RQ features, semantic_merge_proj, set_masking :458-477, 5 x modules/transformer.py TransformerBlock :645-783 with
 * power-normalised linears, per-head LayerNorm of q / k, partial rotary embedding, sigmoid GLU and LayerScale) before
 * extract_mask_positions (:479-486, a strided slice) and the residual VQ (ua2_rvq_*).  csrc/ua2_thinking.cu.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ua2_thinking_cfg {    /* the reference's values: 768 / 128 / 5 / 5 / 1024 / 1024 / 4 */
  int32_t dim;                       /* multiple of dim_heads */
  int32_t dim_heads;                 /* 128 */
  int32_t depth;
  int32_t interval;                  /* frames per query token */
  int32_t whisper_dim;
  int32_t mu_dim;                    /* BEST-RQ feature width */
  int32_t ff_mult;
} ua2_thinking_cfg;
typedef struct ua2_thinking ua2_thinking;

int ua2_thinking_create(const ua2_thinking_cfg* cfg, ua2_thinking** out);
int ua2_thinking_destroy(ua2_thinking* h);
/* one fp32 parameter by its AudioThinking state-dict key ("cls_token", "down_sampling_layer_whisper.weight", "semantic_merge_proj.bias",
 * "encoder_transformers.2.self_attn.q_norm.weight", "encoder_transformers.2.rope.inv_freq", ...).  Weight-normed linears are given
 * with the normalisation applied: "...self_attn.to_qkv.weight", "...self_attn.to_out.weight", "...ff.ff.0.proj.weight",
 * "...ff.ff.2.weight" = g * v / ||v||_row of the parametrizations.weight.original0 / original1 pair. */
int ua2_thinking_load_weight(ua2_thinking* h, const char* key, const float* dptr, const int64_t* shape, int ndim);
int ua2_thinking_finalize(ua2_thinking* h, void* stream);
/* rows per clip of the encoder's sequence, T + T / interval with T = min(Tw / 2, Tb); 0 when T is not a positive multiple of interval */
long long ua2_thinking_rows(ua2_thinking* h, int Tw, int Tb);
/* whisper (B, whisper_dim, Tw), mu (B, mu_dim, Tb) fp32 channels-first -> out (B, rows, dim): the encoder's output sequence; the query
 * tokens are its rows interval, 2 interval + 1, ... (out[:, interval :: interval + 1]) */
int ua2_thinking_encode(ua2_thinking* h, const float* whisper, const float* mu, int B, int Tw, int Tb, float* out, void* stream);
int ua2_thinking_last_launch_count(ua2_thinking* h);

#ifdef __cplusplus
}
#endif
#endif /* UA2_B200_H */
