// Small fused elementwise kernels around the three global stacks of Model_stage3 (llm_models/model_new.py).
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

// model_new.py:598-604 (generate_frame) / :475-487 (forward_prefix):
//   audio_in[m] = sum_c mask[m,c] * audio_embeddings[tok[m,c] + c*V]   (masked codebook merge)
//   text_emb[m] = wte[tok[m,nq]]
// The reference multiplies by the mask (0/1) and sums the 8 streams in order c = 0..7 (torch.sum over dim 2);
// the same order is kept here so the fp32 sum is reproduced exactly.
__global__ void embed_kernel(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ mask,
                             const float* __restrict__ audio_emb, const float* __restrict__ wte,
                             float* __restrict__ audio_in, float* __restrict__ text_emb, int nq, int V, int D, int text_vocab,
                             int* __restrict__ err_flag) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.y;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (k >= D) return;
  const int64_t* tk = tokens + (size_t)m * (nq + 1);
  const uint8_t* mk = mask + (size_t)m * (nq + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // nn.Embedding raises an index error for ids outside the table (also on masked streams: the reference gathers before it masks);
  // here such an id reads row 0 instead of foreign memory and raises the handle's error flag, which the next C-ABI call reports
  bool bad = false;
  for (int c = 0; c < nq; ++c) {
    long long id = tk[c];
    if (id < 0 || id >= V) {
      bad = true;
      id = 0;
    }
    const float4 e = *reinterpret_cast<const float4*>(audio_emb + ((size_t)id + (size_t)c * V) * D + k);
    const float w = mk[c] ? 1.f : 0.f;
    acc.x += e.x * w;
    acc.y += e.y * w;
    acc.z += e.z * w;
    acc.w += e.w * w;
  }
  *reinterpret_cast<float4*>(audio_in + (size_t)m * D + k) = acc;
  long long tid_text = tk[nq];
  if (tid_text < 0 || tid_text >= text_vocab) {
    bad = true;
    tid_text = 0;
  }
  *reinterpret_cast<float4*>(text_emb + (size_t)m * D + k) = *reinterpret_cast<const float4*>(wte + (size_t)tid_text * D + k);
  if (bad && k == 0 && err_flag != nullptr) *err_flag = 1;
}

// ln_f of one stack fused with the mask-mix feeding the next one (model_new.py:607, :610, :613):
//   n = RMSNorm(x; w)                                   (GPT.forward ln_f, lit_model.py:169)
//   MIX_UND_TO_BACKBONE : out = n*ma + add*mt           (add = text_emb)
//   MIX_BACKBONE_TO_GEN : keep = n ; out = n*ma
//   MIX_FINAL           : out = n*ma + add*mt           (add = kept backbone output)
//   MIX_NORM_ONLY       : out = n
__global__ void __launch_bounds__(256) norm_mix_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       float eps, const uint8_t* __restrict__ mask, int nq,
                                                       const float* __restrict__ add, float* __restrict__ keep,
                                                       float* __restrict__ out, int D, int mode) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xr = x + (size_t)m * D;
  float ss = 0.f;
  for (int k = tid * 4; k < D; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rs = rsqrtf(tot / (float)D + eps);
  float ma = 1.f, mt = 0.f;
  if (mode != MIX_NORM_ONLY) {
    ma = mask[(size_t)m * (nq + 1)] ? 1.f : 0.f;
    mt = mask[(size_t)m * (nq + 1) + nq] ? 1.f : 0.f;
  }
  for (int k = tid * 4; k < D; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    const float4 g = *reinterpret_cast<const float4*>(w + k);
    float4 n;
    n.x = (v.x * rs) * g.x;
    n.y = (v.y * rs) * g.y;
    n.z = (v.z * rs) * g.z;
    n.w = (v.w * rs) * g.w;
    float4 o = n;
    if (mode == MIX_UND_TO_BACKBONE || mode == MIX_FINAL) {
      const float4 a = *reinterpret_cast<const float4*>(add + (size_t)m * D + k);
      o.x = __fadd_rn(__fmul_rn(n.x, ma), __fmul_rn(a.x, mt));
      o.y = __fadd_rn(__fmul_rn(n.y, ma), __fmul_rn(a.y, mt));
      o.z = __fadd_rn(__fmul_rn(n.z, ma), __fmul_rn(a.z, mt));
      o.w = __fadd_rn(__fmul_rn(n.w, ma), __fmul_rn(a.w, mt));
    } else if (mode == MIX_BACKBONE_TO_GEN) {
      *reinterpret_cast<float4*>(keep + (size_t)m * D + k) = n;
      o.x = n.x * ma;
      o.y = n.y * ma;
      o.z = n.z * ma;
      o.w = n.w * ma;
    }
    *reinterpret_cast<float4*>(out + (size_t)m * D + k) = o;
  }
}

__global__ void frame_begin_kernel(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ mask, int n_tok,
                                   int64_t* __restrict__ d_tokens, uint8_t* __restrict__ d_mask,
                                   int32_t* __restrict__ d_pos, int32_t* __restrict__ d_bidx, int B, int32_t pos_value,
                                   FrameScalars* __restrict__ d_fs, FrameScalars fs) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x;
  for (int i = t; i < n_tok; i += blockDim.x) {
    d_tokens[i] = tokens[i];
    d_mask[i] = mask[i] ? 1 : 0;
  }
  for (int i = t; i < B; i += blockDim.x) {
    d_pos[i] = pos_value;
    d_bidx[i] = i;
  }
  if (t == 0) *d_fs = fs;
}

// ---- device-side phase / EOS state machine of Generator.generate_tts (evaluation/tts_task.py:253-279, B = 1) ----
// state[0] forbid_prefix of the next frame, [1] done, [2] frames recorded, [3] 1-based index of the frame that switched the phase (0: none)
__global__ void tts_begin_kernel(const int64_t* __restrict__ tokens0, const uint8_t* __restrict__ mask0, const int32_t* __restrict__ prev_sample,
                                 int nq, int64_t* __restrict__ d_tokens, uint8_t* __restrict__ d_mask, int32_t* __restrict__ d_pos,
                                 int32_t* __restrict__ d_bidx, int32_t pos_value, FrameScalars* __restrict__ d_fs, FrameScalars fs,
                                 const int32_t* __restrict__ state) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x;
  if (t <= nq) {
    if (tokens0 != nullptr) {  // first frame of the utterance: the last row of the prompt
      d_tokens[t] = tokens0[t];
      d_mask[t] = mask0[t] ? 1 : 0;
    } else {  // feedback (tts_task.py:276-279): audio tokens of the previous sample in columns 0..nq-1, its text token in column nq
      d_tokens[t] = t < nq ? (int64_t)prev_sample[1 + t] : (int64_t)prev_sample[0];
      d_mask[t] = t < nq ? 1 : 0;
    }
  }
  if (t == 0) {
    d_pos[0] = pos_value;
    d_bidx[0] = 0;
    fs.forbid_prefix = state[0];
    *d_fs = fs;
  }
}

// after the samplers of a frame: apply the reference's break / phase-switch tests to the sampled row and record it
__global__ void tts_state_kernel(const int32_t* __restrict__ sample, int nq, int32_t* __restrict__ state, int32_t* __restrict__ frames_out,
                                 int frames_cap, int reason_eos, int end_tok, int reason_card, int fixed_switch) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x != 0 || state[1]) return;
  bool all_end = true, all_reason_eos = true;
  for (int i = 1; i <= nq; ++i) {
    all_end = all_end && sample[i] == end_tok;
    all_reason_eos = all_reason_eos && sample[i] == reason_eos;
  }
  if (fixed_switch < 0 && all_end) {  // `break` before anything is recorded (tts_task.py:261-262)
    state[1] = 1;
    return;
  }
  const int n = state[2];
  if (n < frames_cap)
    for (int i = 0; i <= nq; ++i) frames_out[(size_t)n * (nq + 1) + i] = sample[i];
  state[2] = n + 1;
  const bool sw = fixed_switch >= 0 ? (n + 1 == fixed_switch) : all_reason_eos;
  if (sw) {  // :263-266
    state[0] = reason_card;
    state[3] = n + 1;
  }
}

__global__ void prefill_begin_kernel(const int64_t* __restrict__ pos64, int32_t* __restrict__ d_pos,
                                     int32_t* __restrict__ d_bidx, int M, int T, int row0) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) {
    d_pos[i] = (int32_t)pos64[row0 + i];
    d_bidx[i] = (row0 + i) / T;
  }
}

// audio_head (nq, d, V) -> (nq, V, d): one-time repack so the per-codebook head (model_new.py:632 torch.mm)
// streams rows like every other nn.Linear weight.
__global__ void transpose_head_kernel(const float* __restrict__ src, float* __restrict__ dst, int d, int V) {
  __shared__ float tile[32][33];
  const int cb = blockIdx.z;
  const float* s = src + (size_t)cb * d * V;
  float* t = dst + (size_t)cb * d * V;
  const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, v = v0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < d && v < V) ? s[(size_t)k * V + v] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int v = v0 + r, k = k0 + threadIdx.x;
    if (v < V && k < d) t[(size_t)v * d + k] = tile[threadIdx.x][r];
  }
}

}  // namespace

static void misc_attrs_once() {
  static DeviceOnce once;
  if (!once.need()) return;
  prefer_max_smem(embed_kernel);
  prefer_max_smem(norm_mix_kernel);
  prefer_max_smem(frame_begin_kernel);
  prefer_max_smem(prefill_begin_kernel);
}

cudaError_t launch_embed(const LaunchCtx& lc, const int64_t* tokens, const uint8_t* mask, const float* audio_emb,
                         const float* wte, float* audio_in, float* text_emb, int M, int nq, int V, int D, int text_vocab, int* err_flag) {
  misc_attrs_once();
  const int threads = 128;
  const dim3 grid((D / 4 + threads - 1) / threads, M);
  return launch(lc, embed_kernel, grid, dim3(threads), 0, tokens, mask, audio_emb, wte, audio_in, text_emb, nq, V, D, text_vocab, err_flag);
}

cudaError_t launch_norm_mix(const LaunchCtx& lc, const float* x, const float* w, float eps, const uint8_t* mask,
                            int nq, const float* add, float* keep, float* out, int M, int D, int mode) {
  misc_attrs_once();
  return launch(lc, norm_mix_kernel, dim3(M), dim3(256), 0, x, w, eps, mask, nq, add, keep, out, D, mode);
}

cudaError_t launch_frame_begin(const LaunchCtx& lc, const int64_t* tokens, const uint8_t* mask, int n_tok,
                               int64_t* d_tokens, uint8_t* d_mask, int32_t* d_pos, int32_t* d_bidx, int B,
                               int32_t pos_value, FrameScalars* d_fs, FrameScalars fs) {
  misc_attrs_once();
  return launch(lc, frame_begin_kernel, dim3(1), dim3(128), 0, tokens, mask, n_tok, d_tokens, d_mask, d_pos, d_bidx, B,
                pos_value, d_fs, fs);
}

cudaError_t launch_tts_begin(const LaunchCtx& lc, const int64_t* tokens0, const uint8_t* mask0, const int32_t* prev_sample, int nq,
                             int64_t* d_tokens, uint8_t* d_mask, int32_t* d_pos, int32_t* d_bidx, int32_t pos_value, FrameScalars* d_fs,
                             FrameScalars fs, const int32_t* state) {
  return launch(lc, tts_begin_kernel, dim3(1), dim3(32), 0, tokens0, mask0, prev_sample, nq, d_tokens, d_mask, d_pos, d_bidx, pos_value, d_fs, fs,
                state);
}
cudaError_t launch_tts_state(const LaunchCtx& lc, const int32_t* sample, int nq, int32_t* state, int32_t* frames_out, int frames_cap,
                             int reason_eos, int end_tok, int reason_card, int fixed_switch) {
  return launch(lc, tts_state_kernel, dim3(1), dim3(32), 0, sample, nq, state, frames_out, frames_cap, reason_eos, end_tok, reason_card,
                fixed_switch);
}

cudaError_t launch_prefill_begin(const LaunchCtx& lc, const int64_t* pos64, int32_t* d_pos, int32_t* d_bidx, int M,
                                 int T, int row0) {
  return launch(lc, prefill_begin_kernel, dim3((M + 127) / 128), dim3(128), 0, pos64, d_pos, d_bidx, M, T, row0);
}

cudaError_t launch_transpose_head(const LaunchCtx& lc, const float* src, float* dst, int nq, int d, int V) {
  const dim3 grid((V + 31) / 32, (d + 31) / 32, nq), block(32, 8);
  return launch(lc, transpose_head_kernel, grid, block, 0, src, dst, d, V);
}

}  // namespace ua2
