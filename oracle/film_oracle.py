"""CPU oracle for the own-code conditioning glue of ReasoningCodec_film.
TEST INFRASTRUCTURE ONLY - never imported by the product path.

Restated from /root/reference/tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py:
    time_film        :428-438   gamma = 1 + self.gamma * tanh(dgamma), beta; a per-sample draw `torch.rand(B, 1, 1) < 0.2`
                                replaces (gamma, beta) by (1, 0); out = gamma * features + beta
    feature_combine  :440-456   reason_adaptor (Linear) -> nearest x2.5 interpolation over time -> crop to T -> + rec_feature

Parity status: PINNED.  oracle/make_golden_film.py executes the UNMODIFIED source text of these two methods (the module itself
cannot be imported: whisper, peft, fairseq ... are absent) on a stand-in `self` and asserts bit-identical results; fixtures in
tests/golden/film_golden.pt.  The product operators ua2_film_f32 / ua2_interp_nearest_f32 / ua2_linear_bias_f32 are compared
with the same formulas on the GPU (tests/test_scalar_gpu.py::test_film_interp_linear_bias).
"""
import torch.nn.functional as F


def time_film(params, features, zero_mask, gamma_scale):
    """params (B, T, 2C) = layer(cond_seq); zero_mask (B,) {0, 1} = the reference's `torch.rand(B, 1, 1) < 0.2` draw."""
    delta_gamma, beta = params.chunk(2, dim=-1)
    gamma = 1.0 + gamma_scale * delta_gamma.tanh()
    mask = zero_mask.float().view(-1, 1, 1)
    gamma = gamma * (1 - mask) + 1.0 * mask
    beta = beta * (1 - mask) + 0.0 * mask
    return gamma * features + beta


def feature_combine(reason_w, reason_b, reasoning_feature, rec_feature):
    B, T, D = rec_feature.shape
    r = F.linear(reasoning_feature, reason_w, reason_b)
    r = F.interpolate(r.permute(0, 2, 1), scale_factor=2.5, mode="nearest").permute(0, 2, 1)
    if r.shape[1] != T:
        r = r[:, :T, :]
    return rec_feature + r
