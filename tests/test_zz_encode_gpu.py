"""GPU parity of the own-code ENCODE chain of ReasoningCodec_film: the product's
AudioDiffusion1D.fetch_codes_from_features (strided convolutions, fusion linears, FiLM, nearest interpolation, residual VQ through the
C ABI) against the fixtures produced from the UNMODIFIED source of fetch_codes_batch (oracle/make_golden_encode.py).
Bar: VQ indices bit-equal, merged condition features within 1e-4 relative to their scale."""
import os

import pytest
import torch

from oracle import dit_oracle as DO
from oracle import encode_oracle as EO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product(p):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import AudioDiffusion1D
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    dit = DO.DitCfg(num_attention_heads=2, attention_head_dim=32, in_channels=EO.CODEC_DIM + 2 * 136, out_channels=136, num_layers=1)
    m = AudioDiffusion1D(Transformer1DModel(**dit.ctor_kwargs()), codec_dim=EO.CODEC_DIM, codebook_size=EO.CB_SIZE, codebook_dim=EO.CB_DIM)
    sd = dict(m.state_dict())
    for k, v in p.items():
        if k.endswith(".codebooks"):
            name = k[: -len(".codebooks")]
            for i in range(v.shape[0]):
                sd[f"{name}.layers.{i}._codebook.embed"] = v[i:i + 1].clone()
        else:
            assert k in sd, k
            sd[k] = v
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def test_fetch_codes_from_features_matches_reference_golden():
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "encode_golden.pt"), weights_only=False)
    p = EO.random_params(gold["param_seed"])
    m = _product(p)
    for c in gold["cases"]:
        feats = EO.stand_in_features(c["feat_seed"], *c["shape"])
        codes, merge = m.fetch_codes_from_features(feats["whisper"], feats["wavlm"], feats["bestrq_acoustic"], feats["bestrq_semantic"],
                                                   feats["quantized_reasoning"], film_masks=c["film_masks"])
        torch.cuda.synchronize()
        assert codes.shape == c["codes"].shape and codes.dtype == torch.int64
        agree = float((codes.cpu() == c["codes"]).float().mean())
        assert torch.equal(codes.cpu(), c["codes"]), f"VQ index agreement {agree:.4f}"
        assert float((merge.cpu() - c["merge"]).abs().max()) < 1e-4 * max(1.0, float(c["merge"].abs().max()))


def test_fetch_codes_draws_its_own_zero_condition_masks_and_checks_lengths():
    p = EO.random_params(3)
    m = _product(p)
    f = EO.stand_in_features(4, 2, 40, 20, 4)
    codes, merge = m.fetch_codes_from_features(f["whisper"], f["wavlm"], f["bestrq_acoustic"], f["bestrq_semantic"], f["quantized_reasoning"])
    assert codes.shape == (2, 10, 8) and merge.shape == (2, 10, EO.CODEC_DIM) and bool(torch.isfinite(merge).all())
    assert int(codes.min()) >= 0 and int(codes.max()) < EO.CB_SIZE
    with pytest.raises(ValueError):  # 12 feature frames against 10 reasoning frames: the reference's time_film cannot broadcast either
        g = EO.stand_in_features(4, 2, 48, 24, 4)
        m.fetch_codes_from_features(g["whisper"], g["wavlm"], g["bestrq_acoustic"], g["bestrq_semantic"], g["quantized_reasoning"])
