"""Generate tests/golden/encode_golden.pt by executing the UNMODIFIED source of AudioDiffusion1D.fetch_codes_batch and time_film
(tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py:428-438, :492-551) on a stand-in `self` - the three SSL front-ends and
the reasoning encoder replaced by recorded stand-in features, `ResidualVQ` by the restatement of oracle/encode_oracle.py (third-party,
unpinned) - and assert that oracle/encode_oracle.fetch_codes_from_features is bit-identical.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_encode
"""
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import encode_oracle as EO  # noqa: E402
from oracle.make_golden_film import load_methods  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "encode_golden.pt")
SEED = 41


class _VQ(nn.Module):  # call surface of vector_quantize_pytorch.ResidualVQ as fetch_codes_batch uses it: (quantized, indices, loss)
    def __init__(self, p, name, nq):
        super().__init__()
        self.p, self.name, self.nq = p, name, nq

    def forward(self, x):
        q, idx = EO.residual_vq_forward(x, self.p, self.name, self.nq)
        return q, idx, torch.zeros(1)


def _module(kind, p, name, **kw):
    w, b = p[f"{name}.weight"], p[f"{name}.bias"]
    m = nn.Conv1d(w.shape[1], w.shape[0], w.shape[2], **kw) if kind == "conv" else nn.Linear(w.shape[1], w.shape[0])
    m.weight.data.copy_(w)
    m.bias.data.copy_(b)
    return m


def main():
    torch.set_num_threads(4)
    m = load_methods(["fetch_codes_batch", "time_film"])
    p = EO.random_params(SEED)
    cases = []
    for ci, (B, Tw, Tb, Tq) in enumerate( ((2, 200, 100, 20), (3, 120, 60, 12))):  # 50 / 30 code frames (the reference needs Tw / 4 == Tb / 2 == 2.5 Tq: time_film broadcasts)
        feats = EO.stand_in_features(SEED + 1 + ci, B, Tw, Tb, Tq)
        assert Tw // 4 == Tb // 2 == int(Tq * 2.5)
        self_ = types.SimpleNamespace(gamma=0.1)
        self_.pretrained_model = types.SimpleNamespace(eval=lambda: None, extract_continous_embeds_multiple=lambda a: (feats["bestrq_acoustic"], feats["bestrq_semantic"]))
        self_.wavlm_encoder = types.SimpleNamespace(eval=lambda: None)
        self_.whisper_encoder = types.SimpleNamespace(eval=lambda: None)
        self_.get_whisper_feature = lambda spec, n, ls: feats["whisper"]
        self_.get_wavlm_feature = lambda wav, ls: feats["wavlm"]
        self_.encode_reasoning_part = lambda w, b: (feats["quantized_reasoning"], torch.zeros(B, Tq, 8, dtype=torch.long), None)
        self_.time_film = lambda cond, f, layer: m["time_film"](self_, cond, f, layer)
        for name, k, s in (("d_conv_whisper", 4, 4), ("d_conv_wavlm", 4, 4), ("d_conv_embedding_semantic", 2, 2), ("d_conv_embedding_acoustic", 2, 2)):
            setattr(self_, name, _module("conv", p, name, stride=s))
        for name in ("cond_fusion_layer_semantic", "cond_fusion_layer_acoustic", "cond_fusion_layer_phone", "reason_adaptor", "cond_feature_emb",
                     "time_film_phone", "time_film_semantic", "time_film_acoustic"):
            setattr(self_, name, _module("linear", p, name))
        for name, nq in EO.VQS:
            setattr(self_, name, _VQ(p, name, nq))
        for seed in range(100):  # a generator state whose three zero-condition draws are neither all-on nor all-off
            torch.manual_seed(seed)
            masks = [(torch.rand(B, 1, 1) < 0.2).float().view(-1) for _ in range(3)]
            if 0 < sum(int(mk.sum()) for mk in masks) < 3 * B:
                break
        torch.manual_seed(seed)
        with torch.no_grad():
            reason_codes, merge_codes, merge_feats = m["fetch_codes_batch"](self_, torch.zeros(B, 1, 240), None)
            codes, merge = EO.fetch_codes_from_features(p, film_masks=masks, **feats)
        assert torch.equal(merge_codes[0], codes), "restated chain: codes differ from the reference source"
        assert torch.equal(merge_feats[0], merge), "restated chain: merge features differ from the reference source"
        cases.append(dict(feat_seed=SEED + 1 + ci, shape=(B, Tw, Tb, Tq), film_masks=[mk.to(torch.uint8) for mk in masks], codes=codes,
                          merge=merge, rand_seed=seed))
        print(f"[ok] B={B} code frames={codes.shape[1]}: fetch_codes_batch source == restatement bit-exact; zero-condition draws "
              f"{[mk.int().tolist() for mk in masks]}")
    assert cases
    torch.save(dict(param_seed=SEED, cases=cases), GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
