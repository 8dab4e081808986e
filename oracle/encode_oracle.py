"""CPU oracle for the own-code ENCODE chain of ReasoningCodec_film: everything `fetch_codes_batch`
(tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py:492-551) does AFTER the three SSL front-ends and the reasoning
encoder have produced their features.  TEST INFRASTRUCTURE ONLY - never imported by the product path.

    whisper (B, 1024, Tw) --d_conv_whisper k4 s4--------------------.
    bestrq acoustic (B, 1024, Tb) --d_conv_embedding_acoustic k2 s2--+- cat -> cond_fusion_layer_acoustic -> time_film -> vq_acoustic (6 q)
    bestrq semantic (B, 1024, Tb) --d_conv_embedding_semantic k2 s2--> cond_fusion_layer_semantic -> time_film -> vq_structure_semantic (1 q)
    wavlm (B, 768, Tw) --d_conv_wavlm k4 s4--------------------------> cond_fusion_layer_phone    -> time_film -> vq_pronunciation_semantic (1 q)
    quantized reasoning (B, Tq, 768) --reason_adaptor, nearest x2.5--> conditions the three time_film heads
    codes = [phone | semantic | 6 x acoustic] (B, T, 8);  merge_features = cond_feature_emb(sum of the three quantized outputs)

Parity status: the CHAIN is PINNED - oracle/make_golden_encode.py executes the UNMODIFIED source text of fetch_codes_batch / time_film
on a stand-in `self` (the module cannot be imported: whisper, peft, fairseq ... are absent) and asserts bit-identical codes and
features.  `ResidualVQ` itself is third-party (vector_quantize_pytorch==1.27.15, pyproject.toml:31, absent here): its eval-mode forward
is restated below from the published algorithm and is UNPINNED.
"""
import math

import torch
import torch.nn.functional as F

from oracle import film_oracle as FO

CODEC_DIM, WHISPER_DIM, WAVLM_DIM, BESTRQ_DIM = 768, 1024, 768, 1024
CB_SIZE, CB_DIM = 8192, 32
VQS = (("vq_pronunciation_semantic", 1), ("vq_structure_semantic", 1), ("vq_acoustic", 6))


def residual_vq_forward(x, p, name, nq):
    """vector_quantize_pytorch.ResidualVQ.forward in eval mode (restated, unpinned): project_in (dim != codebook_dim), then per
    quantizer EuclideanCodebook: dist = -cdist(residual, embed) with cdist = sqrt(clamp(x2 + y2 - 2 xy, min=0)), index = argmax (first
    maximum), quantized = embed[index]; residual -= quantized; output = project_out(sum of quantized).  x (B, T, dim) ->
    (quantized (B, T, dim), indices (B, T, nq))."""
    h = F.linear(x, p[f"{name}.project_in.weight"], p[f"{name}.project_in.bias"])
    residual, total, idx = h, torch.zeros_like(h), []
    for i in range(nq):
        e = p[f"{name}.codebooks"][i]
        x2 = (residual ** 2).sum(-1, keepdim=True)
        y2 = (e ** 2).sum(-1)
        d = (x2 + y2 + (residual @ e.t()) * -2).clamp(min=0).sqrt()
        ind = (-d).argmax(dim=-1)
        q = e[ind]
        residual = residual - q
        total = total + q
        idx.append(ind)
    return F.linear(total, p[f"{name}.project_out.weight"], p[f"{name}.project_out.bias"]), torch.stack(idx, dim=-1)


def min_margin(x, p, name, nq):
    """Smallest gap between the best and the second-best squared distance over all frames and quantizers (how much slack the argmin
    has against a different summation order)."""
    h = F.linear(x, p[f"{name}.project_in.weight"], p[f"{name}.project_in.bias"])
    residual, m = h, float("inf")
    for i in range(nq):
        e = p[f"{name}.codebooks"][i]
        d2 = ((residual ** 2).sum(-1, keepdim=True) + (e ** 2).sum(-1) - 2 * residual @ e.t())
        two = d2.topk(2, dim=-1, largest=False)[0]
        m = min(m, float((two[..., 1] - two[..., 0]).min()))
        residual = residual - e[d2.argmin(-1)]
    return m


def random_params(seed: int):
    """Seeded stand-in parameters under the reference's state-dict names (conv / linear ~ U(+-1/sqrt(fan_in)), codebooks ~ N(0, 1))."""
    g = torch.Generator().manual_seed(seed)

    def u(*shape, fan_in):
        return (torch.rand(*shape, generator=g) * 2 - 1) / math.sqrt(fan_in)

    p = {}
    for name, c, k in (("d_conv_whisper", WHISPER_DIM, 4), ("d_conv_wavlm", WAVLM_DIM, 4), ("d_conv_embedding_semantic", BESTRQ_DIM, 2),
                       ("d_conv_embedding_acoustic", BESTRQ_DIM, 2)):
        p[f"{name}.weight"], p[f"{name}.bias"] = u(c, c, k, fan_in=c * k), u(c, fan_in=c * k)
    for name, cin in (("cond_fusion_layer_semantic", BESTRQ_DIM), ("cond_fusion_layer_acoustic", BESTRQ_DIM + WHISPER_DIM),
                      ("cond_fusion_layer_phone", WAVLM_DIM), ("reason_adaptor", CODEC_DIM), ("cond_feature_emb", CODEC_DIM)):
        p[f"{name}.weight"], p[f"{name}.bias"] = u(CODEC_DIM, cin, fan_in=cin), u(CODEC_DIM, fan_in=cin)
    for name in ("time_film_phone", "time_film_semantic", "time_film_acoustic"):
        p[f"{name}.weight"], p[f"{name}.bias"] = u(2 * CODEC_DIM, CODEC_DIM, fan_in=CODEC_DIM), u(2 * CODEC_DIM, fan_in=CODEC_DIM)
    for name, nq in VQS:
        p[f"{name}.project_in.weight"], p[f"{name}.project_in.bias"] = u(CB_DIM, CODEC_DIM, fan_in=CODEC_DIM), u(CB_DIM, fan_in=CODEC_DIM)
        p[f"{name}.project_out.weight"], p[f"{name}.project_out.bias"] = u(CODEC_DIM, CB_DIM, fan_in=CB_DIM), u(CODEC_DIM, fan_in=CB_DIM)
        p[f"{name}.codebooks"] = torch.randn(nq, CB_SIZE, CB_DIM, generator=g) * 0.5
    return p


def stand_in_features(seed, B, Tw, Tb, Tq):
    """Seeded stand-ins for what the SSL front-ends / reasoning encoder hand to the chain (fixtures store the seed, not the tensors)."""
    g = torch.Generator().manual_seed(seed)
    return dict(whisper=torch.randn(B, WHISPER_DIM, Tw, generator=g), wavlm=torch.randn(B, WAVLM_DIM, Tw, generator=g),
                bestrq_acoustic=torch.randn(B, BESTRQ_DIM, Tb, generator=g), bestrq_semantic=torch.randn(B, BESTRQ_DIM, Tb, generator=g),
                quantized_reasoning=torch.randn(B, Tq, CODEC_DIM, generator=g))


def fetch_codes_from_features(p, whisper, wavlm, bestrq_acoustic, bestrq_semantic, quantized_reasoning, film_masks, gamma=0.1):
    """The chain of AudioDiffusion1D.fetch_codes_batch :515-551.  film_masks: three (B,) {0, 1} tensors = the `torch.rand(B, 1, 1) <
    0.2` draws of the phone / semantic / acoustic time_film calls, in that order.  Returns (codes (B, T, 8), merge_features (B, T, 768))."""
    whisper_rec = F.conv1d(whisper, p["d_conv_whisper.weight"], p["d_conv_whisper.bias"], stride=4)
    wavlm_feat = F.conv1d(wavlm, p["d_conv_wavlm.weight"], p["d_conv_wavlm.bias"], stride=4)
    sem_rec = F.conv1d(bestrq_semantic, p["d_conv_embedding_semantic.weight"], p["d_conv_embedding_semantic.bias"], stride=2)
    acoustic = F.conv1d(bestrq_acoustic, p["d_conv_embedding_acoustic.weight"], p["d_conv_embedding_acoustic.bias"], stride=2)
    reasoning = F.linear(quantized_reasoning, p["reason_adaptor.weight"], p["reason_adaptor.bias"])
    reasoning = F.interpolate(reasoning.permute(0, 2, 1), scale_factor=2.5, mode="nearest").permute(0, 2, 1)

    def branch(feat_bct, fusion, film, vq, nq, mask):
        f = F.linear(feat_bct.transpose(1, 2), p[f"{fusion}.weight"], p[f"{fusion}.bias"])
        params = F.linear(reasoning, p[f"{film}.weight"], p[f"{film}.bias"])
        f = FO.time_film(params, f, mask, gamma)
        return residual_vq_forward(f, p, vq, nq)

    q_phone, c_phone = branch(wavlm_feat, "cond_fusion_layer_phone", "time_film_phone", "vq_pronunciation_semantic", 1, film_masks[0])
    q_sem, c_sem = branch(sem_rec, "cond_fusion_layer_semantic", "time_film_semantic", "vq_structure_semantic", 1, film_masks[1])
    n = min(acoustic.shape[-1], whisper_rec.shape[-1])
    q_ac, c_ac = branch(torch.cat([acoustic[:, :, :n], whisper_rec[:, :, :n]], dim=1), "cond_fusion_layer_acoustic", "time_film_acoustic",
                        "vq_acoustic", 6, film_masks[2])
    merge = F.linear(q_phone + q_sem + q_ac, p["cond_feature_emb.weight"], p["cond_feature_emb.bias"])
    return torch.cat([c_phone, c_sem, c_ac], dim=-1), merge
