// Op descriptors of the persistent chain kernel (ua2_chain.cu).
#pragma once
#include "ua2_kernels.cuh"

namespace ua2 {

enum : int { OP_GEMV = 0, OP_ATTN = 1, OP_EMBED = 2, OP_NORMMIX = 3 };
constexpr int CHAIN_ATTN_CHUNK = 32;  // keys per attention work item inside the chain (K/V staged in the activation tile area)

struct V3Cfg_public {
  int nsl, ngrp, SL, KCW, stages, n_splits;
};

struct ChainOp {
  int type = OP_GEMV;
  int pro = 0, epi = 0;
  int next_gemv = -1;  // index of the next OP_GEMV in the chain (weight look-ahead), -1 if none
  GemvParams g;
  V3Cfg_public c{};
  AttnParams a;
  // OP_EMBED / OP_NORMMIX (one row)
  const int64_t* tokens = nullptr;
  const uint8_t* mask = nullptr;
  const float* audio_emb = nullptr;
  const float* wte = nullptr;
  float* audio_in = nullptr;
  float* text_emb = nullptr;
  int nq = 0, V = 0, D = 0;
  const float* nm_x = nullptr;
  const float* nm_w = nullptr;
  float nm_eps = 0.f;
  const float* nm_add = nullptr;
  float* nm_keep = nullptr;
  float* nm_out = nullptr;
  int nm_mode = 0;
};

size_t chain_smem_bytes();
void chain_cfg_for(int K, V3Cfg_public* out);
constexpr int CHAIN_PROF_OPS = 512;  // ops per launch the optional profile buffer covers (2 CTAs x CHAIN_PROF_OPS x 5 clocks)
cudaError_t launch_chain(cudaStream_t stream, const ChainOp* d_ops, int n_ops, unsigned* d_sync, int n_ctas,
                         unsigned long long* d_prof = nullptr);
int chain_max_ctas();

}  // namespace ua2
