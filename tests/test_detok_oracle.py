"""CPU suite for the codes -> waveform caller (ReasoningTokenizer.token2audio_no_reason / AudioDiffusion1D.inference_codes):
  * oracle/detok_oracle.py reproduces the fixtures written by oracle/make_golden_detok.py from the UNMODIFIED reference source
    of those methods (tests/golden/detok_golden.pt) bit-exactly;
  * the PRODUCT's host logic (uniaudio2_b200...reason_tokenizer.ReasoningTokenizer: windowing, in-context continuation,
    cross-fade) reproduces the same fixtures bit-exactly when its two device-side collaborators are replaced by oracle-backed
    stand-ins - the part of the product that is plain host code is thereby checked here, without a GPU."""
import os

import pytest
import torch

from oracle import detok_oracle as TO
from oracle import dit_oracle as DO
from oracle.make_golden_detok import CB_DIM, CB_SIZE, CODEC_DIM, DIT, THREADS, VQS, random_params, sq_decode_standin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    n = torch.get_num_threads()
    torch.set_num_threads(THREADS)
    yield
    torch.set_num_threads(n)


@pytest.fixture(scope="module")
def detok_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "detok_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def oracle_parts(detok_golden):
    p = random_params(31)
    assert {k: float(v.double().sum()) for k, v in p.items()} == detok_golden["checksum"]
    return p, TO.DetokOracle(p, DO.DitOracle(DIT, DO.random_state_dict(DIT, seed=32)), sq_decode_standin(p))


def test_oracle_matches_reference(detok_golden, oracle_parts):
    _, orc = oracle_parts
    with torch.no_grad():
        for c in detok_golden["cases"]:
            n = c["codes"].shape[-1]
            torch.manual_seed(1000 + n)
            wav = orc.token2audio_no_reason(c["codes"], duration=c["duration"], num_steps=c["steps"])
            assert torch.equal(wav, c["wav"]), n
            it = iter(c["draws"])  # ... and with the recorded draws replayed explicitly
            wav2 = orc.token2audio_no_reason(c["codes"], duration=c["duration"], num_steps=c["steps"], randn=lambda shape: next(it))
            assert torch.equal(wav2, c["wav"]), n
            assert wav.shape[-1] == int(n / 12.5 * 24000)


def test_product_host_logic_matches_reference(detok_golden, oracle_parts):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.reason_tokenizer import ReasoningTokenizer

    p, orc = oracle_parts

    class FakeModel:  # AudioDiffusion1D.inference_codes served by the oracle (CPU)
        def inference_codes(self, codes, spk_embeds, true_latents, latent_length, incontext_length, additional_feats, guidance_scale=2,
                            num_steps=20, disable_progress=True, scenario="start_seg"):
            assert len(codes) == 1 and spk_embeds is None and additional_feats == [] and scenario == "other_seg"
            return orc.inference_codes(codes[0], true_latents, latent_length, incontext_length, guidance_scale, num_steps,
                                       lambda shape: torch.randn(*shape))

    class FakeSQ:
        decode = staticmethod(sq_decode_standin(p))

    tok = ReasoningTokenizer(FakeModel(), FakeSQ(), device=torch.device("cpu"))
    for c in detok_golden["cases"]:
        n = c["codes"].shape[-1]
        torch.manual_seed(1000 + n)
        wav = tok.token2audio_no_reason(c["codes"], False, duration=c["duration"], num_steps=c["steps"], disable_progress=True)
        assert wav.dtype == c["wav"].dtype and torch.equal(wav, c["wav"]), n
    # detokenize_no_reason (:399-404): (8, T2) -> batch of one, default 20 s windows, `steps` -> num_steps
    seen = {}
    tok.token2audio_no_reason = lambda rec, **kw: seen.update(shape=tuple(rec.shape), **kw) or "wave"
    assert tok.detokenize_no_reason(c["codes"][0], False, steps=7) == "wave"
    assert seen == dict(shape=(1, 8, c["codes"].shape[-1]), return_reasoning_text=False, guidance_scale=1.5, num_steps=7, disable_progress=False)


def test_residual_vq_restatement_properties():
    g = torch.Generator().manual_seed(0)
    cb = torch.randn(3, 10, 4, generator=g)
    w, b = torch.randn(6, 4, generator=g), torch.randn(6, generator=g)
    idx = torch.randint(0, 10, (2, 5, 3), generator=g)
    out = TO.residual_vq_output_from_indices(cb, w, b, idx)
    manual = torch.zeros(2, 5, 4)
    for q in range(3):
        manual += cb[q][idx[..., q]]
    assert torch.allclose(out, manual @ w.t() + b, atol=1e-6)


def test_product_modules_keys_and_refusals():
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import AudioDiffusion1D, ResidualVQ
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    est = Transformer1DModel(**DIT.ctor_kwargs())
    m = AudioDiffusion1D(est, codec_dim=CODEC_DIM, codebook_size=CB_SIZE, codebook_dim=CB_DIM)
    keys = set(m.state_dict().keys())
    for name, nq in VQS:
        assert {f"{name}.project_out.weight", f"{name}.project_out.bias"} <= keys
        assert {f"{name}.layers.{i}._codebook.embed" for i in range(nq)} <= keys
        assert m.state_dict()[f"{name}.layers.0._codebook.embed"].shape == (1, CB_SIZE, CB_DIM)
    assert {"cond_feature_emb.weight", "cond_feature_emb.bias", "zero_cond_embedding1", "cfm_wrapper.estimator.scale_shift_table"} <= keys
    vq = ResidualVQ(dim=CODEC_DIM, codebook_size=CB_SIZE, codebook_dim=CB_DIM, num_quantizers=2, decay=0.9, commitment_weight=1.0)
    sd = dict(vq.state_dict())
    sd["project_in.weight"] = torch.zeros(CB_DIM, CODEC_DIM)        # entries of the real package that the decode path ignores
    sd["layers.0._codebook.cluster_size"] = torch.zeros(1, CB_SIZE)
    vq.load_state_dict(sd, strict=True)
    with pytest.raises(Exception):  # CPU tensors: refuse, no fallback
        vq.lookup_sum(torch.zeros(1, 2, 4, dtype=torch.int64), 0)
    with pytest.raises(Exception):
        m.inference_codes([torch.zeros(1, 8, 4, dtype=torch.int64)], None, torch.zeros(1, 8, 136), 8, 0, additional_feats=[],
                          guidance_scale=1.5, num_steps=1, scenario="other_seg")
