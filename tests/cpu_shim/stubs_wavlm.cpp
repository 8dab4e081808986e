// TEST INFRASTRUCTURE ONLY: what the whole-handle CPU build of csrc/ua2_wavlm.cu needs from csrc/ua2_dit.cu - the dense attention
// launchers launch_dense_attn_f32 / launch_dense_attn_bias_f32 - comes with that file's kernel part (the handle code behind it needs
// the tensor-core GEMM and stays out).
#include "ua2_kernels.cuh"
#include "ua2_dit_kernels.inc"
