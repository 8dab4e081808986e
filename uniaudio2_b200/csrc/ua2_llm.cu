// Handle API of the AR-decode path: Model_stage3.{setup_caches, reset_caches, forward_prefix, generate_frame}
// (llm_models/model_new.py:554-651) sequenced as sm_100a kernels on one stream, replayed as a CUDA graph.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"

namespace ua2 {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

struct LayerW {
  const float *norm1 = nullptr, *qkv = nullptr, *proj = nullptr, *norm2 = nullptr, *fc1 = nullptr, *fc2 = nullptr,
              *mproj = nullptr;
};

struct Stack {
  ua2_gpt_cfg cfg{};
  std::string prefix;
  std::vector<LayerW> layers;
  const float* ln_f = nullptr;
  const float *cos = nullptr, *sin = nullptr;
  int64_t rope_rows = 0;
  int S_max = 0;
  std::vector<float*> kc, vc;  // per layer (B, G, S_max, hs)
  size_t cache_floats = 0;
};

}  // namespace ua2

using namespace ua2;

struct ua2_llm {
  ua2_llm_cfg cfg{};
  Stack st[4];  // 0 backbone, 1 decoder, 2 understanding, 3 generation
  const float *wte = nullptr, *lm_head = nullptr, *audio_emb = nullptr, *projection = nullptr, *audio_head_src = nullptr;
  float* audio_head_t = nullptr;  // (nq, V_a, d) repacked
  bool ready = false;
  int B_max = 0;
  int M_cap = 0;  // rows per launch (prefill chunk)
  int max_splits = 0;
  // fixed device buffers
  int64_t* d_tokens = nullptr;
  uint8_t* d_mask = nullptr;
  int32_t *d_pos = nullptr, *d_bidx = nullptr, *d_pos_local = nullptr;
  FrameScalars* d_fs = nullptr;
  int* err_flag = nullptr;      // pinned, device-mapped: set by embed_kernel when a token id is outside its embedding table
  int* err_flag_dev = nullptr;  // the device alias of the same word
  float *x = nullptr, *audio_in = nullptr, *text_emb = nullptr, *hb = nullptr, *h_final = nullptr, *qbuf = nullptr,
        *hmlp = nullptr, *sg_ws = nullptr, *o_part = nullptr, *ml_part = nullptr, *dec_x = nullptr, *text_logits = nullptr,
        *audio_logits = nullptr;
  size_t sg_ws_floats = 0;
  int64_t h_final_numel = 0, text_logits_numel = 0, audio_logits_numel = 0;
  std::vector<void*> owned;
  // options / stats
  TcWorkspace tcws;        // scratch of the tcgen05 3xTF32 path (many-row linears of forward_prefix)
  int opt_chunk_rows = 0;  // prefill rows per pass (0 = M_cap)
  bool rows_are_batch = false;  // generate_frame: activation row m belongs to batch row m (prefill flattens B x T)
  int opt_attn_direct = 0;  // local decoder: attention inside the proj prologue (32 fewer launches; measured 1466 vs 1471 tok/s: off)
  int opt_graph = 1, opt_pdl = 1;  // a persistent one-kernel "chain" form measured 18 % slower than graph + PDL (profiles/r1_chain_experiment.md) and was removed
  int last_launches = 0;
  unsigned long long frame_counter = 0;
  unsigned long long opt_epoch = 0;  // option_epoch() the captured graphs / recorded sequences were made under
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
  };  // an entry with exec == nullptr means "seen once, ran eagerly" (lazy kernel attributes are set by then)
  std::map<unsigned long long, GraphEntry> graphs;
  std::map<unsigned long long, GemvSeq> seqs;  // per frame shape: the sequence of linears (tail L2 prefetch planning)
};

namespace {

int alloc(ua2_llm* h, void** p, size_t bytes) {
  UA2_CHECK_CUDA(cudaMalloc(p, bytes));
  h->owned.push_back(*p);
  return UA2_OK;
}

bool parse_layer_key(const std::string& rest, int& layer, std::string& leaf) {
  // rest = "transformer.h.<i>.<leaf>"
  const std::string pre = "transformer.h.";
  if (rest.compare(0, pre.size(), pre) != 0) return false;
  size_t dot = rest.find('.', pre.size());
  if (dot == std::string::npos) return false;
  layer = atoi(rest.substr(pre.size(), dot - pre.size()).c_str());
  leaf = rest.substr(dot + 1);
  return true;
}

const char* kPrefix[4] = {"backbone.", "decoder.", "audio_understanding_expert.", "audio_generation_expert."};

// One transformer Block (lit_model.py:307-349) on M rows: 5 kernels.
// the kernel just launched barely touches HBM (1 attention, 2 sampler): the preceding linear may prefetch more
inline void mark_idle(const LaunchCtx& lc, int kind) {
  if (lc.seq && !lc.seq->recorded && !lc.seq->ops.empty()) lc.seq->ops.back().idle_after |= kind;
}

cudaError_t run_block(ua2_llm* h, const LaunchCtx& lc, Stack& s, int l, float* x, int M, const int32_t* pos,
                      const int32_t* bidx, int n_splits) {
  const ua2_gpt_cfg& c = s.cfg;
  const LayerW& w = s.layers[l];
  const int D = c.n_embd, hs = c.head_size, QD = c.n_head * hs;
  cudaError_t e;
  {  // A: RMSNorm -> QKV -> RoPE -> cache append
    GemvParams p;
    p.W = w.qkv;
    p.N = (c.n_head + 2 * c.n_query_groups) * hs;
    p.K = D;
    p.M = M;
    p.ws = h->sg_ws;
    p.ws_floats = h->sg_ws_floats;
    p.tc = h->tcws.a ? &h->tcws : nullptr;
    p.X = x;
    p.ldx = D;
    p.norm_w = w.norm1;
    p.eps = c.norm_eps;
    p.pos = pos;
    p.bidx = bidx;
    p.bidx_identity = h->rows_are_batch ? 1 : 0;
    p.n_head = c.n_head;
    p.n_groups = c.n_query_groups;
    p.hs = hs;
    p.q_out = h->qbuf;
    p.k_cache = s.kc[l];
    p.v_cache = s.vc[l];
    p.cos = s.cos;
    p.sin = s.sin;
    p.S_max = s.S_max;
    if ((e = launch_gemv(lc, PRO_RMSNORM, EPI_QKV, p)) != cudaSuccess) return e;
  }
  // short caches (the local decoder's <= 8 codebook steps): attention runs inside the proj kernel's prologue
  const bool direct = h->opt_attn_direct && s.S_max <= ATTN_DIRECT_MAX_KEYS && hs == 64 && M <= 16;
  if (!direct) {  // B: split-softmax attention over the cache
    AttnParams a;
    a.q = h->qbuf;
    a.k_cache = s.kc[l];
    a.v_cache = s.vc[l];
    a.pos = pos;
    a.bidx = bidx;
    a.bidx_identity = h->rows_are_batch ? 1 : 0;
    a.o_part = h->o_part;
    a.ml_part = h->ml_part;
    a.M = M;
    a.n_head = c.n_head;
    a.n_groups = c.n_query_groups;
    a.hs = hs;
    a.S_max = s.S_max;
    a.max_splits = h->max_splits;
    a.n_splits_launch = n_splits;
    if ((e = launch_attn(lc, a)) != cudaSuccess) return e;
    mark_idle(lc, 1);
  }
  {  // C: combine splits -> proj -> + residual
    GemvParams p;
    p.W = w.proj;
    p.N = D;
    p.K = QD;
    p.M = M;
    p.ws = h->sg_ws;
    p.ws_floats = h->sg_ws_floats;
    p.tc = h->tcws.a ? &h->tcws : nullptr;
    p.o_part = h->o_part;
    p.ml_part = h->ml_part;
    p.max_splits = h->max_splits;
    p.n_splits = n_splits;
    p.pos = pos;
    p.n_head = c.n_head;
    p.hs = hs;
    p.Y = x;
    p.ldy = D;
    p.R = x;
    p.ldr = D;
    if (direct) {
      p.X = h->qbuf;
      p.ldx = QD;
      p.k_cache = s.kc[l];
      p.v_cache = s.vc[l];
      p.bidx = bidx;
      p.n_groups = c.n_query_groups;
      p.S_max = s.S_max;
      if ((e = launch_gemv(lc, PRO_ATTN_DIRECT, EPI_RESADD, p)) != cudaSuccess) return e;
    } else if ((e = launch_gemv(lc, PRO_ATTN, EPI_RESADD, p)) != cudaSuccess) {
      return e;
    }
  }
  {  // D: RMSNorm -> fc_1 | fc_2 -> silu * mul
    GemvParams p;
    p.W = w.fc1;
    p.W2 = w.fc2;
    p.N = c.intermediate_size;
    p.K = D;
    p.M = M;
    p.ws = h->sg_ws;
    p.ws_floats = h->sg_ws_floats;
    p.tc = h->tcws.a ? &h->tcws : nullptr;
    p.X = x;
    p.ldx = D;
    p.norm_w = w.norm2;
    p.eps = c.norm_eps;
    p.Y = h->hmlp;
    p.ldy = c.intermediate_size;
    if ((e = launch_gemv(lc, PRO_RMSNORM, EPI_SWIGLU, p)) != cudaSuccess) return e;
  }
  {  // E: mlp.proj + residual
    GemvParams p;
    p.W = w.mproj;
    p.N = D;
    p.K = c.intermediate_size;
    p.M = M;
    p.ws = h->sg_ws;
    p.ws_floats = h->sg_ws_floats;
    p.tc = h->tcws.a ? &h->tcws : nullptr;
    p.X = h->hmlp;
    p.ldx = c.intermediate_size;
    p.Y = x;
    p.ldy = D;
    p.R = x;
    p.ldr = D;
    if ((e = launch_gemv(lc, PRO_PLAIN, EPI_RESADD, p)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// embedding merge + understanding expert + backbone + generation expert (model_new.py:593-613 / :474-497)
// on M rows whose tokens/mask/pos/bidx already sit in the handle's fixed buffers.  Leaves h_final (M x D).
cudaError_t run_global(ua2_llm* h, const LaunchCtx& lc, int M, int n_splits, bool need_final) {
  const int D = h->cfg.backbone.n_embd, nq = h->cfg.num_codebooks;
  cudaError_t e;
  if ((e = launch_embed(lc, h->d_tokens, h->d_mask, h->audio_emb, h->wte, h->audio_in, h->text_emb, M, nq,
                        h->cfg.audio_vocab, D, h->cfg.text_vocab, h->err_flag_dev)) != cudaSuccess)
    return e;
  Stack& und = h->st[2];
  for (int l = 0; l < und.cfg.n_layer; ++l)
    if ((e = run_block(h, lc, und, l, h->audio_in, M, h->d_pos, h->d_bidx, n_splits)) != cudaSuccess) return e;
  if ((e = launch_norm_mix(lc, h->audio_in, und.ln_f, und.cfg.norm_eps, h->d_mask, nq, h->text_emb, nullptr, h->x, M, D,
                           MIX_UND_TO_BACKBONE)) != cudaSuccess)
    return e;
  Stack& bb = h->st[0];
  for (int l = 0; l < bb.cfg.n_layer; ++l)
    if ((e = run_block(h, lc, bb, l, h->x, M, h->d_pos, h->d_bidx, n_splits)) != cudaSuccess) return e;
  // gen expert input reuses audio_in as its residual stream
  if ((e = launch_norm_mix(lc, h->x, bb.ln_f, bb.cfg.norm_eps, h->d_mask, nq, nullptr, h->hb, h->audio_in, M, D,
                           MIX_BACKBONE_TO_GEN)) != cudaSuccess)
    return e;
  Stack& gen = h->st[3];
  for (int l = 0; l < gen.cfg.n_layer; ++l)
    if ((e = run_block(h, lc, gen, l, h->audio_in, M, h->d_pos, h->d_bidx, n_splits)) != cudaSuccess) return e;
  if (need_final)
    if ((e = launch_norm_mix(lc, h->audio_in, gen.ln_f, gen.cfg.norm_eps, h->d_mask, nq, h->hb, nullptr, h->h_final, M,
                             D, MIX_FINAL)) != cudaSuccess)
      return e;
  return cudaSuccess;
}

// text head + the 8-step local decoder (model_new.py:615-645) for B rows
cudaError_t run_heads(ua2_llm* h, const LaunchCtx& lc, int B, int rows) {
  const int D = h->cfg.backbone.n_embd, nq = h->cfg.num_codebooks, Vt = h->cfg.text_vocab, Va = h->cfg.audio_vocab;
  Stack& dec = h->st[1];
  const int d = dec.cfg.n_embd;
  cudaError_t e;
  {  // lm_head(last_h) -> sample_topk
    GemvParams p;
    p.W = h->lm_head;
    p.N = Vt;
    p.K = D;
    p.M = B;
    p.X = h->h_final;
    p.ldx = D;
    p.Y = h->text_logits;
    p.ldy = Vt;
    p.ws = h->sg_ws;
    p.ws_floats = h->sg_ws_floats;
    p.tc = h->tcws.a ? &h->tcws : nullptr;
    if ((e = launch_gemv(lc, PRO_PLAIN, EPI_STORE, p)) != cudaSuccess) return e;
    if ((e = launch_sampler(lc, h->text_logits, Vt, h->d_fs, 0, 0, nq + 1, 0, 0, B, rows)) != cudaSuccess) return e;
    mark_idle(lc, 2);
  }
  // the sampler writes the int32 frame into a mirror buffer placed right after the audio logits; the next step's
  // projection gathers audio_embeddings[tok_{i-1} + (i-1)*V_a] (model_new.py:640-641, :662-663) straight from it
  const int32_t* mirror = reinterpret_cast<const int32_t*>(h->audio_logits + h->audio_logits_numel);
  for (int i = 0; i < nq; ++i) {
    {  // projection(curr_h): curr_h = last_h (i == 0) or the embedding of the token sampled at step i-1
      GemvParams p;
      p.W = h->projection;
      p.N = d;
      p.K = D;
      p.M = B;
      p.Y = h->dec_x;
      p.ldy = d;
      p.ws = h->sg_ws;
      p.ws_floats = h->sg_ws_floats;
      p.tc = h->tcws.a ? &h->tcws : nullptr;
      if (i == 0) {
        p.X = h->h_final;
        p.ldx = D;
        if ((e = launch_gemv(lc, PRO_PLAIN, EPI_STORE, p)) != cudaSuccess) return e;
      } else {
        p.emb = h->audio_emb;
        p.gidx = mirror + i;  // column 1 + (i-1) of the (B, nq+1) frame
        p.gidx_stride = nq + 1;
        p.gidx_offset = (i - 1) * Va;
        if ((e = launch_gemv(lc, PRO_GATHER, EPI_STORE, p)) != cudaSuccess) return e;
      }
    }
    const int32_t* pos_i = h->d_pos_local + (size_t)i * h->B_max;
    for (int l = 0; l < dec.cfg.n_layer; ++l)
      if ((e = run_block(h, lc, dec, l, h->dec_x, B, pos_i, h->d_bidx, 1)) != cudaSuccess) return e;
    {  // ln_f -> audio_head[i] -> audio_sample_topk
      GemvParams p;
      p.W = h->audio_head_t + (size_t)i * Va * d;
      p.N = Va;
      p.K = d;
      p.M = B;
      p.X = h->dec_x;
      p.ldx = d;
      p.norm_w = dec.ln_f;
      p.eps = dec.cfg.norm_eps;
      p.Y = h->audio_logits + (size_t)i * B * Va;
      p.ldy = Va;
      p.ws = h->sg_ws;
      p.ws_floats = h->sg_ws_floats;
      p.tc = h->tcws.a ? &h->tcws : nullptr;
      if ((e = launch_gemv(lc, PRO_RMSNORM, EPI_STORE, p)) != cudaSuccess) return e;
      if ((e = launch_sampler(lc, h->audio_logits + (size_t)i * B * Va, Va, h->d_fs, 1, 1 + i, nq + 1,
                              (long long)rows * Vt + (long long)i * rows * Va, 1 + i, B, rows)) != cudaSuccess)
        return e;
      mark_idle(lc, 2);
    }
  }
  return cudaSuccess;
}


}  // namespace

extern "C" {

const char* ua2_last_error(void) { return g_err.c_str(); }
const char* ua2_version(void) { return "uniaudio2_b200 0.1 (sm_100a)"; }

int ua2_device_sm_count(void) {
  int dev = 0, n = 0;
  UA2_CHECK_CUDA(cudaGetDevice(&dev));
  UA2_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

int ua2_llm_create(const ua2_llm_cfg* cfg, ua2_llm** out) {
  UA2_REQUIRE(cfg && out, "null cfg/out");
  const ua2_gpt_cfg* g[4] = {&cfg->backbone, &cfg->decoder, &cfg->understanding, &cfg->generation};
  for (int i = 0; i < 4; ++i) {
    UA2_REQUIRE(g[i]->n_layer > 0 && g[i]->n_embd > 0 && g[i]->n_head > 0 && g[i]->n_query_groups > 0, "bad gpt cfg");
    UA2_REQUIRE(g[i]->head_size == 128 || g[i]->head_size == 64 || g[i]->head_size == 32, "head_size must be 32/64/128");
    UA2_REQUIRE(g[i]->n_head % g[i]->n_query_groups == 0 && g[i]->n_head / g[i]->n_query_groups <= 4,
                "n_head / n_query_groups must be an integer <= 4");
    UA2_REQUIRE(g[i]->n_embd % 4 == 0 && g[i]->intermediate_size % 4 == 0, "sizes must be multiples of 4");
  }
  UA2_REQUIRE(cfg->understanding.n_embd == cfg->backbone.n_embd && cfg->generation.n_embd == cfg->backbone.n_embd,
              "experts must match the backbone width (model_new.py:351-355)");
  UA2_REQUIRE(cfg->num_codebooks >= 1 && cfg->num_codebooks <= 32, "num_codebooks out of range");
  UA2_REQUIRE(cfg->text_vocab % 2 == 0 && cfg->audio_vocab % 2 == 0, "vocab sizes must be even");
  UA2_REQUIRE(cfg->max_seq_length >= 1, "max_seq_length");
  ua2_llm* h = new ua2_llm();
  h->cfg = *cfg;
  for (int i = 0; i < 4; ++i) {
    h->st[i].cfg = *g[i];
    h->st[i].prefix = kPrefix[i];
    h->st[i].layers.resize(g[i]->n_layer);
    h->st[i].S_max = (i == 1) ? cfg->num_codebooks : cfg->max_seq_length;
  }
  *out = h;
  return UA2_OK;
}

int ua2_llm_destroy(ua2_llm* h) {
  if (!h) return UA2_OK;
  for (auto& kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : h->owned) cudaFree(p);
  if (h->err_flag) cudaFreeHost(h->err_flag);
  delete h;
  return UA2_OK;
}

int ua2_llm_load_weight(ua2_llm* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape, "null argument");
  const std::string key(key_c);
  auto numel = [&]() {
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    return n;
  };
  auto expect2 = [&](int64_t a, int64_t b) { return ndim == 2 && shape[0] == a && shape[1] == b; };
  const int D = h->cfg.backbone.n_embd;
  if (key == "audio_embeddings.weight") {
    UA2_REQUIRE(expect2((int64_t)h->cfg.audio_vocab * h->cfg.num_codebooks, D), key + ": bad shape");
    h->audio_emb = dptr;
    return UA2_OK;
  }
  if (key == "projection.weight") {
    UA2_REQUIRE(expect2(h->cfg.decoder.n_embd, D), key + ": bad shape");
    h->projection = dptr;
    return UA2_OK;
  }
  if (key == "audio_head") {
    UA2_REQUIRE(ndim == 3 && shape[0] == h->cfg.num_codebooks && shape[1] == h->cfg.decoder.n_embd &&
                    shape[2] == h->cfg.audio_vocab,
                key + ": bad shape");
    h->audio_head_src = dptr;
    return UA2_OK;
  }
  for (int si = 0; si < 4; ++si) {
    Stack& s = h->st[si];
    if (key.compare(0, s.prefix.size(), s.prefix) != 0) continue;
    const std::string rest = key.substr(s.prefix.size());
    const ua2_gpt_cfg& c = s.cfg;
    if (rest == "rope_cos" || rest == "rope_sin") {
      UA2_REQUIRE(ndim == 2 && shape[1] == c.head_size && shape[0] >= s.S_max, key + ": need (>=S_max, head_size)");
      (rest == "rope_cos" ? s.cos : s.sin) = dptr;
      s.rope_rows = shape[0];
      return UA2_OK;
    }
    if (rest == "transformer.ln_f.weight") {
      UA2_REQUIRE(numel() == c.n_embd, key + ": bad shape");
      s.ln_f = dptr;
      return UA2_OK;
    }
    if (si == 0 && rest == "transformer.wte.weight") {
      UA2_REQUIRE(expect2(h->cfg.text_vocab, D), key + ": bad shape");
      h->wte = dptr;
      return UA2_OK;
    }
    if (si == 0 && rest == "lm_head.weight") {
      UA2_REQUIRE(expect2(h->cfg.text_vocab, D), key + ": bad shape");
      h->lm_head = dptr;
      return UA2_OK;
    }
    int l = -1;
    std::string leaf;
    if (parse_layer_key(rest, l, leaf)) {
      UA2_REQUIRE(l >= 0 && l < c.n_layer, key + ": layer index out of range");
      LayerW& w = s.layers[l];
      const int64_t qkv_out = (int64_t)(c.n_head + 2 * c.n_query_groups) * c.head_size;
      if (leaf == "norm_1.weight") {
        UA2_REQUIRE(numel() == c.n_embd, key + ": bad shape");
        w.norm1 = dptr;
      } else if (leaf == "norm_2.weight") {
        UA2_REQUIRE(numel() == c.n_embd, key + ": bad shape");
        w.norm2 = dptr;
      } else if (leaf == "attn.qkv.weight") {
        UA2_REQUIRE(expect2(qkv_out, c.n_embd), key + ": bad shape");
        w.qkv = dptr;
      } else if (leaf == "attn.proj.weight") {
        UA2_REQUIRE(expect2(c.n_embd, (int64_t)c.n_head * c.head_size), key + ": bad shape");
        w.proj = dptr;
      } else if (leaf == "mlp.fc_1.weight") {
        UA2_REQUIRE(expect2(c.intermediate_size, c.n_embd), key + ": bad shape");
        w.fc1 = dptr;
      } else if (leaf == "mlp.fc_2.weight") {
        UA2_REQUIRE(expect2(c.intermediate_size, c.n_embd), key + ": bad shape");
        w.fc2 = dptr;
      } else if (leaf == "mlp.proj.weight") {
        UA2_REQUIRE(expect2(c.n_embd, c.intermediate_size), key + ": bad shape");
        w.mproj = dptr;
      } else {
        UA2_REQUIRE(false, key + ": unknown parameter");
      }
      return UA2_OK;
    }
  }
  UA2_REQUIRE(false, key + ": unknown parameter");
}

int ua2_llm_setup_caches(ua2_llm* h, int max_batch_size, void* stream_v) {
  UA2_REQUIRE(h, "null handle");
  UA2_REQUIRE(max_batch_size >= 1 && max_batch_size <= 256, "max_batch_size must be in 1..256");
  if (h->ready) {
    set_error("setup_caches called twice (create a new handle to change the batch size)");
    return UA2_ERR_STATE;
  }
  cudaStream_t stream = (cudaStream_t)stream_v;
  // every parameter present?  (load_state_dict strict=True, llm_utils/train_utils.py:159-177)
  UA2_REQUIRE(h->wte && h->lm_head && h->audio_emb && h->projection && h->audio_head_src, "missing top-level weights");
  for (int si = 0; si < 4; ++si) {
    Stack& s = h->st[si];
    UA2_REQUIRE(s.ln_f && s.cos && s.sin, s.prefix + ": missing ln_f / rope tables");
    for (auto& w : s.layers)
      UA2_REQUIRE(w.norm1 && w.qkv && w.proj && w.norm2 && w.fc1 && w.fc2 && w.mproj, s.prefix + ": missing layer weights");
  }
  const int B = max_batch_size, nq = h->cfg.num_codebooks, D = h->cfg.backbone.n_embd, d = h->cfg.decoder.n_embd;
  h->B_max = B;
  h->M_cap = B > 1024 ? B : 1024;  // prefill rows per pass: large enough for the tiled / tensor-core GEMMs to fill the GPU
  h->max_splits = (h->cfg.max_seq_length + ATTN_CHUNK - 1) / ATTN_CHUNK;
  const int Mc = h->M_cap;
  int rc;
  for (int si = 0; si < 4; ++si) {
    Stack& s = h->st[si];
    s.cache_floats = (size_t)B * s.cfg.n_query_groups * s.S_max * s.cfg.head_size;
    for (int l = 0; l < s.cfg.n_layer; ++l) {
      float *k = nullptr, *v = nullptr;
      if ((rc = alloc(h, (void**)&k, s.cache_floats * 4))) return rc;
      if ((rc = alloc(h, (void**)&v, s.cache_floats * 4))) return rc;
      UA2_CHECK_CUDA(cudaMemsetAsync(k, 0, s.cache_floats * 4, stream));
      UA2_CHECK_CUDA(cudaMemsetAsync(v, 0, s.cache_floats * 4, stream));
      s.kc.push_back(k);
      s.vc.push_back(v);
    }
  }
  int max_qd = 0, max_inter = 0, max_heads_hs = 0, max_nhead = 0;
  for (int si = 0; si < 4; ++si) {
    const ua2_gpt_cfg& c = h->st[si].cfg;
    max_qd = std::max(max_qd, c.n_head * c.head_size);
    max_inter = std::max(max_inter, c.intermediate_size);
    max_heads_hs = std::max(max_heads_hs, c.n_head * c.head_size);
    max_nhead = std::max(max_nhead, c.n_head);
  }
  if ((rc = alloc(h, (void**)&h->d_tokens, (size_t)Mc * (nq + 1) * 8))) return rc;
  if ((rc = alloc(h, (void**)&h->d_mask, (size_t)Mc * (nq + 1)))) return rc;
  if ((rc = alloc(h, (void**)&h->d_pos, (size_t)Mc * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->d_bidx, (size_t)Mc * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->d_pos_local, (size_t)nq * B * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->d_fs, sizeof(FrameScalars)))) return rc;
  if (h->err_flag == nullptr) {
    UA2_CHECK_CUDA(cudaHostAlloc((void**)&h->err_flag, sizeof(int), cudaHostAllocMapped));
    *h->err_flag = 0;
    UA2_CHECK_CUDA(cudaHostGetDevicePointer((void**)&h->err_flag_dev, h->err_flag, 0));
  }
  if ((rc = alloc(h, (void**)&h->x, (size_t)Mc * D * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->audio_in, (size_t)Mc * D * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->text_emb, (size_t)Mc * D * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->hb, (size_t)Mc * D * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->h_final, (size_t)Mc * D * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->qbuf, (size_t)Mc * max_qd * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->hmlp, (size_t)Mc * max_inter * 4))) return rc;
  h->sg_ws_floats = (size_t)Mc * (2 + max_qd) + 4;  // tiled-GEMM scratch: row statistics + attention combine
  if ((rc = alloc(h, (void**)&h->sg_ws, h->sg_ws_floats * 4))) return rc;
  if (tc_gemm_available()) {
    size_t kmax = 0, nmax = 0;
    for (int si = 0; si < 4; ++si) {
      const ua2_gpt_cfg& c = h->st[si].cfg;
      const size_t Dm = c.n_embd, QD = (size_t)c.n_head * c.head_size, QKV = (size_t)(c.n_head + 2 * c.n_query_groups) * c.head_size,
                   F = c.intermediate_size;
      kmax = std::max({kmax, Dm, QD, F});
      nmax = std::max({nmax, QKV, Dm, 2 * F});
    }
    h->tcws.a_floats = (size_t)Mc * 2 * kmax;
    h->tcws.slots_floats = tc_slots_max_floats();
    const size_t nheads = std::max({nmax, (size_t)h->cfg.text_vocab, (size_t)h->cfg.audio_vocab});
    h->tcws.c_floats = std::max((size_t)Mc * nmax, (size_t)B * nheads);  // prefill passes never run the heads; frames have M <= B
    if ((rc = alloc(h, (void**)&h->tcws.a, h->tcws.a_floats * 4))) return rc;
    if ((rc = alloc(h, (void**)&h->tcws.slots, h->tcws.slots_floats * 4))) return rc;
    if ((rc = alloc(h, (void**)&h->tcws.c, h->tcws.c_floats * 4))) return rc;
  }
  if ((rc = alloc(h, (void**)&h->o_part, (size_t)Mc * max_heads_hs * h->max_splits * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->ml_part, (size_t)Mc * max_nhead * h->max_splits * 2 * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->dec_x, (size_t)B * d * 4))) return rc;
  if ((rc = alloc(h, (void**)&h->text_logits, (size_t)B * h->cfg.text_vocab * 4))) return rc;
  // audio logits (nq, B, V_a) followed by an int32 mirror of the sampled frame (B, nq+1)
  h->audio_logits_numel = (int64_t)nq * B * h->cfg.audio_vocab;
  if ((rc = alloc(h, (void**)&h->audio_logits, (size_t)h->audio_logits_numel * 4 + (size_t)B * (nq + 1) * 4))) return rc;
  h->h_final_numel = (int64_t)B * D;
  h->text_logits_numel = (int64_t)B * h->cfg.text_vocab;
  // local-decoder positions: step i uses position i for every row (model_new.py:627, :643)
  std::vector<int32_t> pl((size_t)nq * B);
  for (int i = 0; i < nq; ++i)
    for (int b = 0; b < B; ++b) pl[(size_t)i * B + b] = i;
  UA2_CHECK_CUDA(cudaMemcpyAsync(h->d_pos_local, pl.data(), pl.size() * 4, cudaMemcpyHostToDevice, stream));
  // audio_head repack (nq, d, V) -> (nq, V, d)
  if ((rc = alloc(h, (void**)&h->audio_head_t, (size_t)nq * d * h->cfg.audio_vocab * 4))) return rc;
  LaunchCtx lc;
  lc.stream = stream;
  UA2_CHECK_CUDA(launch_transpose_head(lc, h->audio_head_src, h->audio_head_t, nq, d, h->cfg.audio_vocab));
  UA2_CHECK_CUDA(cudaStreamSynchronize(stream));
  h->ready = true;
  return UA2_OK;
}

// nn.Embedding's IndexError analogue: an earlier call gathered with an id outside [0, vocab); reported by the next call (the flag is
// written by the device into mapped host memory - no synchronisation on the hot path) and cleared
#define UA2_CHECK_IDS(h)                                                                                                   \
  do {                                                                                                                     \
    if ((h)->err_flag && *(volatile int*)(h)->err_flag) {                                                                  \
      *(volatile int*)(h)->err_flag = 0;                                                                                   \
      UA2_REQUIRE(false, "index out of range in self: a token id of an earlier call lies outside its embedding table");   \
    }                                                                                                                      \
  } while (0)

int ua2_llm_reset_caches(ua2_llm* h, void* stream_v) {
  UA2_REQUIRE(h, "null handle");
  UA2_CHECK_IDS(h);
  if (!h->ready) {
    set_error("You need to call setup_caches() first");  // lit_model.py:134-135 TypeError analogue
    return UA2_ERR_STATE;
  }
  cudaStream_t stream = (cudaStream_t)stream_v;
  for (int si = 0; si < 4; ++si) {
    Stack& s = h->st[si];
    for (int l = 0; l < s.cfg.n_layer; ++l) {
      UA2_CHECK_CUDA(cudaMemsetAsync(s.kc[l], 0, s.cache_floats * 4, stream));
      UA2_CHECK_CUDA(cudaMemsetAsync(s.vc[l], 0, s.cache_floats * 4, stream));
    }
  }
  return UA2_OK;
}

int ua2_llm_prefill(ua2_llm* h, const int64_t* tokens, const uint8_t* mask, const int64_t* pos, int B, int T,
                    int64_t max_pos, void* stream_v) {
  UA2_REQUIRE(h && tokens && mask && pos, "null argument");
  if (!h->ready) {
    set_error("You need to call setup_caches() first");
    return UA2_ERR_STATE;
  }
  UA2_CHECK_IDS(h);
  UA2_REQUIRE(B >= 1 && B <= h->B_max, "batch size exceeds setup_caches(max_batch_size)");
  UA2_REQUIRE(T >= 1 && T <= h->cfg.max_seq_length,
              "Cannot forward sequence of length T, max seq length is only max_seq_length");  // lit_model.py:120-121
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int nq = h->cfg.num_codebooks;
  const int Mtot = B * T;
  int launches = 0;
  LaunchCtx lc;
  lc.stream = stream;
  lc.pdl = h->opt_pdl != 0;
  lc.launch_counter = &launches;
  UA2_REQUIRE(max_pos < h->cfg.max_seq_length, "Positions in 'input_pos' must be in [0,max_seq_length)");
  const int n_splits = max_pos < 0 ? h->max_splits : (int)(max_pos / ATTN_CHUNK) + 1;
  // rows are processed in chunks of <= M_cap, t-major inside each batch row, so causality is preserved:
  // a chunk appends all of its K/V (kernel A) before its own attention (kernel B) runs.
  int chunk = (h->opt_chunk_rows > 0 && h->opt_chunk_rows < h->M_cap) ? h->opt_chunk_rows : h->M_cap;
  if (h->opt_chunk_rows <= 0 && Mtot > chunk) {  // equal passes instead of a short tail pass that would fall off the GEMM paths
    const int n_pass = (Mtot + chunk - 1) / chunk;
    chunk = std::min(chunk, ((Mtot + n_pass - 1) / n_pass + 7) / 8 * 8);
  }
  for (int row0 = 0; row0 < Mtot; row0 += chunk) {
    const int M = std::min(chunk, Mtot - row0);
    UA2_CHECK_CUDA(cudaMemcpyAsync(h->d_tokens, tokens + (size_t)row0 * (nq + 1), (size_t)M * (nq + 1) * 8,
                                   cudaMemcpyDeviceToDevice, stream));
    UA2_CHECK_CUDA(cudaMemcpyAsync(h->d_mask, mask + (size_t)row0 * (nq + 1), (size_t)M * (nq + 1),
                                   cudaMemcpyDeviceToDevice, stream));
    LaunchCtx lc0 = lc;
    lc0.pdl = false;
    h->rows_are_batch = false;
    UA2_CHECK_CUDA(launch_prefill_begin(lc0, pos, h->d_pos, h->d_bidx, M, T, row0));
    UA2_CHECK_CUDA(run_global(h, lc, M, n_splits, false));
  }
  h->last_launches = launches;
  return UA2_OK;
}

}  // extern "C"

// The kernels of one frame after frame_begin (global stacks, heads, local decoder, samplers): eager on the first use of a shape,
// captured on the second, replayed as a CUDA graph from then on.  lc.launch_counter must be set.
static int run_frame_body(ua2_llm* h, LaunchCtx lc, int B, int rows, int64_t input_pos, bool use_cfg) {
  cudaStream_t stream = lc.stream;
  if (h->opt_epoch != option_epoch()) {  // a process-wide option changed: the captured frames were built from the old choice
    for (auto& kv : h->graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
    h->seqs.clear();
    h->opt_epoch = option_epoch();
  }
  const int n_splits = (int)(input_pos / ATTN_CHUNK) + 1;
  lc.pdl = h->opt_pdl != 0;
  const unsigned long long key = ((unsigned long long)B << 32) | ((unsigned long long)n_splits << 8) |
                                 (use_cfg ? 2ull : 0ull) | (h->opt_pdl ? 1ull : 0ull) | (h->opt_attn_direct ? 4ull : 0ull) |
                                 ((get_tc_gemm() && B >= get_tc_min_rows()) ? 8ull : 0ull);
  // the frame's sequence of linears: recorded by the first run of this shape, replayed (with tail prefetch specs of the
  // following weights) by every later run / by the graph capture
  GemvSeq& sq = h->seqs[key];
  sq.pos = 0;
  struct SeqDone {
    GemvSeq& s;
    ~SeqDone() { s.recorded = s.recorded || !s.ops.empty(); }
  };
  if (!h->opt_graph) {
    lc.seq = &sq;
    SeqDone done{sq};
    UA2_CHECK_CUDA(run_global(h, lc, B, n_splits, true));
    UA2_CHECK_CUDA(run_heads(h, lc, B, rows));
  } else {
    lc.seq = &sq;
    SeqDone done{sq};
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
      // first use of this shape: run eagerly (also sets the lazily-initialised kernel attributes)
      UA2_CHECK_CUDA(run_global(h, lc, B, n_splits, true));
      UA2_CHECK_CUDA(run_heads(h, lc, B, rows));
      h->graphs.emplace(key, ua2_llm::GraphEntry());
      return UA2_OK;
    }
    if (it->second.exec == nullptr) {
      // second use: capture on a private stream so the caller's stream state is untouched
      cudaStream_t cs;
      UA2_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      LaunchCtx lcc = lc;
      lcc.stream = cs;
      int glaunches = 0;
      lcc.launch_counter = &glaunches;
      cudaGraph_t graph = nullptr;
      UA2_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      cudaError_t e1 = run_global(h, lcc, B, n_splits, true);
      cudaError_t e2 = (e1 == cudaSuccess) ? run_heads(h, lcc, B, rows) : e1;
      cudaError_t e3 = cudaStreamEndCapture(cs, &graph);
      if (e2 != cudaSuccess || e3 != cudaSuccess) {
        set_error(std::string("graph capture failed: ") + cudaGetErrorString(e2 != cudaSuccess ? e2 : e3));
        cudaStreamDestroy(cs);
        return UA2_ERR_CUDA;
      }
      UA2_CHECK_CUDA(cudaGraphInstantiate(&it->second.exec, graph, 0));
      it->second.launches = glaunches;
      cudaGraphDestroy(graph);
      cudaStreamDestroy(cs);
    }
    UA2_CHECK_CUDA(cudaGraphLaunch(it->second.exec, stream));
    *lc.launch_counter += it->second.launches;
  }
  return UA2_OK;
}

extern "C" {

int ua2_llm_generate_frame(ua2_llm* h, const int64_t* tokens, const uint8_t* mask, int B, int64_t input_pos,
                           float temperature, int topk, int forbid_prefix, float cfg_scale, const float* noise,
                           uint64_t seed, int32_t* out, void* stream_v) {
  UA2_REQUIRE(h && tokens && mask && out, "null argument");
  if (!h->ready) {
    set_error("You need to call setup_caches() first");
    return UA2_ERR_STATE;
  }
  UA2_CHECK_IDS(h);
  UA2_REQUIRE(B >= 1 && B <= h->B_max, "batch size exceeds setup_caches(max_batch_size)");
  UA2_REQUIRE(input_pos >= 0 && input_pos < h->cfg.max_seq_length, "Positions in 'input_pos' must be in [0,max_seq_length)");
  // model_new.py:165-180
  UA2_REQUIRE(temperature > 0.f, "temperature must be > 0");
  UA2_REQUIRE(forbid_prefix >= 0, "forbid_prefix must be >= 0");
  UA2_REQUIRE(forbid_prefix < h->cfg.audio_vocab, "forbid_prefix must be smaller than vocab size");
  UA2_REQUIRE(topk >= 1 && topk <= h->cfg.audio_vocab - forbid_prefix, "topk must be in 1..effective_vocab given forbid_prefix");
  UA2_REQUIRE(topk <= h->cfg.text_vocab, "topk exceeds the text vocabulary");  // torch.topk would raise
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int nq = h->cfg.num_codebooks;
  const bool use_cfg = cfg_scale > 1.0f && B > 1;
  UA2_REQUIRE(!use_cfg || B == 2, "CFG expects exactly 2 rows (cond, uncond) - tts_task.py:230-234");
  const int rows = use_cfg ? 1 : B;

  FrameScalars fs;
  fs.temperature = temperature;
  fs.topk = topk;
  fs.forbid_prefix = forbid_prefix;
  fs.cfg_scale = cfg_scale;
  fs.seed = seed;
  fs.offset = h->frame_counter++;
  fs.noise = noise;
  // the sampler writes the int32 frame into the mirror buffer that the next step's embedding gather reads;
  // the caller's `out` receives a copy at the end of the frame.
  int32_t* mirror = reinterpret_cast<int32_t*>(h->audio_logits + h->audio_logits_numel);
  fs.out = mirror;
  fs.rows = rows;
  fs.B = B;

  int launches = 0;
  LaunchCtx lc;
  lc.stream = stream;
  lc.launch_counter = &launches;
  h->rows_are_batch = true;
  UA2_CHECK_CUDA(launch_frame_begin(lc, tokens, mask, B * (nq + 1), h->d_tokens, h->d_mask, h->d_pos, h->d_bidx, B,
                                    (int32_t)input_pos, h->d_fs, fs));
  lc.pdl = h->opt_pdl != 0;
  if (int rc = run_frame_body(h, lc, B, rows, input_pos, use_cfg)) return rc;
  UA2_CHECK_CUDA(cudaMemcpyAsync(out, mirror, (size_t)B * (nq + 1) * 4, cudaMemcpyDeviceToDevice, stream));
  h->last_launches = launches;
  return UA2_OK;
}

int ua2_llm_tts_frames(ua2_llm* h, const int64_t* tokens0, const uint8_t* mask0, int64_t input_pos, int n_frames, float temperature, int topk,
                       const float* noise, int64_t noise_stride, uint64_t seed, int reason_eos, int end_tok, int reason_card, int fixed_switch,
                       int32_t* state, int32_t* frames_out, int frames_cap, void* stream_v) {
  UA2_REQUIRE(h && state && frames_out, "null argument");
  if (!h->ready) {
    set_error("You need to call setup_caches() first");
    return UA2_ERR_STATE;
  }
  UA2_REQUIRE((tokens0 == nullptr) == (mask0 == nullptr), "tokens0 and mask0 go together");
  UA2_REQUIRE(n_frames >= 1 && input_pos >= 0 && input_pos + n_frames <= h->cfg.max_seq_length, "Positions in 'input_pos' must be in [0,max_seq_length)");
  UA2_REQUIRE(temperature > 0.f, "temperature must be > 0");
  UA2_REQUIRE(reason_card >= 0 && reason_card < h->cfg.audio_vocab, "forbid_prefix must be smaller than vocab size");
  UA2_REQUIRE(topk >= 1 && topk <= h->cfg.audio_vocab - reason_card && topk <= h->cfg.text_vocab, "topk must be in 1..effective_vocab given forbid_prefix");
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int nq = h->cfg.num_codebooks;
  int32_t* mirror = reinterpret_cast<int32_t*>(h->audio_logits + h->audio_logits_numel);
  int launches = 0;
  for (int f = 0; f < n_frames; ++f) {
    FrameScalars fs;
    fs.temperature = temperature;
    fs.topk = topk;
    fs.forbid_prefix = 0;  // overwritten on the device from state[0]
    fs.cfg_scale = 1.f;
    fs.seed = seed;
    fs.offset = h->frame_counter++;
    fs.noise = noise ? noise + (size_t)f * noise_stride : nullptr;
    fs.out = mirror;
    fs.rows = 1;
    fs.B = 1;
    LaunchCtx lc;
    lc.stream = stream;
    lc.launch_counter = &launches;
    h->rows_are_batch = true;
    const bool first = f == 0 && tokens0 != nullptr;
    UA2_CHECK_CUDA(launch_tts_begin(lc, first ? tokens0 : nullptr, first ? mask0 : nullptr, mirror, nq, h->d_tokens, h->d_mask, h->d_pos, h->d_bidx,
                                    (int32_t)(input_pos + f), h->d_fs, fs, state));
    lc.pdl = h->opt_pdl != 0;
    if (int rc = run_frame_body(h, lc, 1, 1, input_pos + f, false)) return rc;
    lc.pdl = false;
    UA2_CHECK_CUDA(launch_tts_state(lc, mirror, nq, state, frames_out, frames_cap, reason_eos, end_tok, reason_card, fixed_switch));
  }
  h->last_launches = launches;
  return UA2_OK;
}

int ua2_llm_get_kv(ua2_llm* h, int which, int layer, float** k, float** v) {
  UA2_REQUIRE(h && h->ready && which >= 0 && which < 4 && k && v, "bad argument");
  UA2_REQUIRE(layer >= 0 && layer < h->st[which].cfg.n_layer, "layer out of range");
  *k = h->st[which].kc[layer];
  *v = h->st[which].vc[layer];
  return UA2_OK;
}

int ua2_llm_get_buffer(ua2_llm* h, const char* name, float** ptr, int64_t* numel) {
  UA2_REQUIRE(h && h->ready && name && ptr && numel, "bad argument");
  const std::string n(name);
  if (n == "h_final") {
    *ptr = h->h_final;
    *numel = h->h_final_numel;
  } else if (n == "text_logits") {
    *ptr = h->text_logits;
    *numel = h->text_logits_numel;
  } else if (n == "audio_logits") {
    *ptr = h->audio_logits;
    *numel = h->audio_logits_numel;
  } else {
    UA2_REQUIRE(false, "unknown buffer " + n);
  }
  return UA2_OK;
}

int ua2_llm_set_option(ua2_llm* h, const char* name, int value) {
  UA2_REQUIRE(h && name, "bad argument");
  const std::string n(name);
  if (n == "graph")
    h->opt_graph = value;
  else if (n == "pdl")
    h->opt_pdl = value;
  else if (n == "prefill_chunk_rows")
    h->opt_chunk_rows = value;
  else if (n == "attn_direct")
    h->opt_attn_direct = value ? 1 : 0;
  else
    UA2_REQUIRE(false, "unknown option " + n);
  return UA2_OK;
}

int ua2_llm_last_launch_count(ua2_llm* h) { return h ? h->last_launches : 0; }

}  // extern "C"
