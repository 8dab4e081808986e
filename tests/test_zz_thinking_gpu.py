"""GPU: the reasoning encoder drop-in (csrc/ua2_thinking.cu behind tools/tokenizer/ReasoningCodec_film/models/audio_thinking.py) against
query tokens produced by the UNMODIFIED reference code (tests/golden/thinking_golden.pt, oracle/make_golden_thinking.py) and against
the oracle (bit-equal to that code) at the reference's geometry.

Bar (floating point, fp32 class: 3xTF32 GEMMs, fp32 attention): 2e-4 of the output scale; reasoning codes bit-equal to the restated
ResidualVQ on the same query tokens."""
import os

import pytest
import torch

from oracle import thinking_oracle as TO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(cfg, sd):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.audio_thinking import AudioThinking

    m = AudioThinking(dim=cfg["dim"], interval=cfg["interval"], encoder_depth=cfg["depth"], whisper_fea_dim=cfg["whisper_dim"], mu_dim=cfg["mu_dim"],
                      dim_heads=cfg["dim_heads"], ff_mult=cfg["ff_mult"])
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0")


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("name", ["small", "ragged"])
def test_thinking_matches_reference_golden(name):
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "thinking_golden.pt"), weights_only=False)
    c = gold["cases"][name]
    m = _model(c["cfg"], TO.random_state_dict(c["cfg"], c["seed"]))
    q = m.query_tokens(c["whisper"].cuda(), c["mu"].cuda())
    assert q.shape == c["query_tokens"].shape and bool(torch.isfinite(q).all())
    assert _rel(q.cpu(), c["query_tokens"]) < 2e-4, _rel(q.cpu(), c["query_tokens"])
    assert torch.equal(m.query_tokens(c["whisper"].cuda(), c["mu"].cuda()), q)  # deterministic


def test_thinking_reference_geometry_vs_oracle_and_codes():
    """dim 768, 6 heads of 128, 5 blocks, 30 s window: 1500 Whisper frames + 750 BEST-RQ frames -> 900 rows -> 150 query tokens -> 8 codes each."""
    from oracle import encode_oracle as EO

    cfg = dict(TO.CFG)
    sd = TO.random_state_dict(cfg, 9)
    g = torch.Generator().manual_seed(10)
    B = 2
    whisper, mu = torch.randn(B, 1024, 1500, generator=g), torch.randn(B, 1024, 750, generator=g)
    with torch.no_grad():
        ref = TO.encode(sd, cfg, whisper, mu)
    m = _model(cfg, sd)
    # a codebook for the quantiser
    vq = {"project_in.weight": torch.randn(64, 768, generator=g) / 768 ** 0.5, "project_in.bias": 0.1 * torch.randn(64, generator=g),
          "project_out.weight": torch.randn(768, 64, generator=g) / 8, "project_out.bias": 0.1 * torch.randn(768, generator=g)}
    books = torch.randn(8, 4096, 64, generator=g)
    for i in range(8):
        vq[f"layers.{i}._codebook.embed"] = books[i:i + 1] * (0.7 ** i)
    m.reasoning_vq.load_state_dict(vq)
    m = m.to("cuda:0")
    q = m.query_tokens(whisper.cuda(), mu.cuda())
    assert q.shape == ref.shape == (B, 150, 768)
    assert _rel(q.cpu(), ref) < 2e-4, _rel(q.cpu(), ref)
    quantized, codes, loss = m.encode_reasoning_part(whisper.cuda(), mu.cuda())
    assert loss is None and codes.shape == (B, 150, 8) and quantized.shape == (B, 150, 768)
    p = {"reasoning_vq.project_in.weight": vq["project_in.weight"], "reasoning_vq.project_in.bias": vq["project_in.bias"],
         "reasoning_vq.project_out.weight": vq["project_out.weight"], "reasoning_vq.project_out.bias": vq["project_out.bias"],
         "reasoning_vq.codebooks": torch.stack([books[i] * (0.7 ** i) for i in range(8)])}
    want_q, want_codes = EO.residual_vq_forward(q.cpu(), p, "reasoning_vq", 8)  # the restated quantiser on the SAME query tokens
    assert torch.equal(codes.cpu(), want_codes)
    assert _rel(quantized.cpu(), want_q) < 1e-4


def test_thinking_through_audio_diffusion_and_interface_errors():
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import AudioDiffusion1D
    from oracle.make_golden_thinking import SMALL

    m = _model(SMALL, TO.random_state_dict(SMALL, 21))
    host = AudioDiffusion1D.__new__(AudioDiffusion1D)  # the method reads audio_thinking only
    torch.nn.Module.__init__(host)
    host.audio_thinking = None
    with pytest.raises(_lib.Ua2Error):
        AudioDiffusion1D.encode_reasoning_part(host, torch.zeros(1, 64, 20, device="cuda"), torch.zeros(1, 48, 10, device="cuda"))
    AudioDiffusion1D.attach_audio_thinking(host, m)
    with pytest.raises(RuntimeError):  # 11 frames cannot be grouped by 5: the reference's set_masking fails in its reshape
        m.query_tokens(torch.zeros(1, 64, 22, device="cuda"), torch.zeros(1, 48, 11, device="cuda"))
    with pytest.raises(ValueError):
        m.query_tokens(torch.zeros(1, 32, 20, device="cuda"), torch.zeros(1, 48, 10, device="cuda"))
    q = m.query_tokens(torch.randn(1, 64, 20, device="cuda"), torch.randn(1, 48, 10, device="cuda"))
    assert q.shape == (1, 2, 256)
