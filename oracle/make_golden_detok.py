"""Generate tests/golden/detok_golden.pt by executing the UNMODIFIED source text of
    ReasoningTokenizer.token2audio_no_reason   (tools/tokenizer/ReasoningCodec_film/reason_tokenizer.py:228-306)
    AudioDiffusion1D.inference_codes / prepare_latents   (models/AudioDiffusion1D.py:553-624, :652-655)
    BASECFM (solve_euler)                       (models/AudioDiffusion1D.py:62-129)
over the unmodified in-repo Transformer1DModel (imported over oracle/diffusers_stub.py) and stand-ins for what cannot exist here
(vector_quantize_pytorch.ResidualVQ -> the restatement in oracle/detok_oracle.py; the SQ-codec decoder -> a seeded transposed
convolution with the real hop of 960 samples per latent frame), and assert that oracle/detok_oracle.py produces the bit-identical
waveform.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_detok
"""
import ast
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import detok_oracle as TO  # noqa: E402
from oracle import dit_oracle as DO  # noqa: E402
from oracle.diffusers_stub import install_diffusers_stub  # noqa: E402
from oracle.make_golden_dit import load_basecfm  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "detok_golden.pt")
THREADS = 4
CODEC_DIM, CB_DIM, CB_SIZE, LAT = 48, 16, 64, 136   # production: 768, 32, 8192, 136
DIT = DO.DitCfg(num_attention_heads=2, attention_head_dim=64, in_channels=2 * LAT + CODEC_DIM, out_channels=LAT, num_layers=1,
                num_positional_embeddings=64)
VQS = (("vq_pronunciation_semantic", 1), ("vq_structure_semantic", 1), ("vq_acoustic", 6))


def extract_method(rel_path, cls_name, fn_name, ns):
    path = os.path.join(REF_ROOT, "tools", "tokenizer", "ReasoningCodec_film", rel_path)
    src = open(path).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == fn_name][0]
    text = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))  # from the `def` line: decorators dropped
    exec(compile(text, path, "exec"), ns)
    return ns[fn_name]


def random_params(seed):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, nq in VQS:
        p[f"{name}.codebooks"] = torch.randn(nq, CB_SIZE, CB_DIM, generator=g)
        p[f"{name}.project_out.weight"] = torch.randn(CODEC_DIM, CB_DIM, generator=g) / math.sqrt(CB_DIM)
        p[f"{name}.project_out.bias"] = 0.1 * torch.randn(CODEC_DIM, generator=g)
    p["cond_feature_emb.weight"] = torch.randn(CODEC_DIM, CODEC_DIM, generator=g) / math.sqrt(CODEC_DIM)
    p["cond_feature_emb.bias"] = 0.1 * torch.randn(CODEC_DIM, generator=g)
    p["zero_cond_embedding1"] = torch.randn(CODEC_DIM, generator=g)
    p["sq_decode.weight"] = torch.randn(LAT, 1, 960, generator=g) / math.sqrt(LAT)  # stand-in SQ-codec decoder (hop 960)
    return p


def sq_decode_standin(p):
    return lambda latent: F.conv_transpose1d(latent, p["sq_decode.weight"], stride=960)


class ResidualVQStandIn:
    """What the extracted inference_codes needs from vector_quantize_pytorch.ResidualVQ: .eval() and .get_output_from_indices."""

    def __init__(self, p, name):
        self.p, self.name = p, name

    def eval(self):
        return self

    def get_output_from_indices(self, indices):
        p, n = self.p, self.name
        return TO.residual_vq_output_from_indices(p[f"{n}.codebooks"], p[f"{n}.project_out.weight"], p[f"{n}.project_out.bias"], indices)


def build_reference(p, dit_sd):
    install_diffusers_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    est = Transformer1DModel(**DIT.ctor_kwargs())
    est.load_state_dict(dit_sd, strict=True)
    est = est.float().eval()
    BASECFM = load_basecfm()
    ns = {"torch": torch, "F": F, "nn": nn, "math": math, "np": np,
          "randn_tensor": lambda shape, generator=None, device=None, dtype=None: torch.randn(shape, generator=generator, device=device, dtype=dtype)}
    model = types.SimpleNamespace(device=torch.device("cpu"), dtype=torch.float32, sq_codec_latent=LAT, max_t_len=30 * 50,
                                  cfm_wrapper=BASECFM(est), zero_cond_embedding1=p["zero_cond_embedding1"])
    for name, _ in VQS:
        setattr(model, name, ResidualVQStandIn(p, name))
    lin = nn.Linear(CODEC_DIM, CODEC_DIM)
    lin.weight.data.copy_(p["cond_feature_emb.weight"])
    lin.bias.data.copy_(p["cond_feature_emb.bias"])
    model.cond_feature_emb = lin
    model.prepare_latents = types.MethodType(extract_method("models/AudioDiffusion1D.py", "AudioDiffusion1D", "prepare_latents", ns), model)
    model.inference_codes = types.MethodType(extract_method("models/AudioDiffusion1D.py", "AudioDiffusion1D", "inference_codes", ns), model)
    tok = types.SimpleNamespace(device=torch.device("cpu"), sample_rate=24000, sq_codec_hz=25, rec_frame_rate=12.5, reason_frame_rate=5,
                                model=model, SQCodec=types.SimpleNamespace(decode=sq_decode_standin(p)))
    tok.token2audio_no_reason = types.MethodType(extract_method("reason_tokenizer.py", "ReasoningTokenizer", "token2audio_no_reason", ns), tok)
    return tok


def main():
    torch.set_num_threads(THREADS)
    p = random_params(31)
    dit_sd = DO.random_state_dict(DIT, seed=32)
    tok = build_reference(p, dit_sd)
    orc = TO.DetokOracle(p, DO.DitOracle(DIT, dit_sd), sq_decode_standin(p))
    g = torch.Generator().manual_seed(33)
    out = {"cases": []}
    with torch.no_grad():
        # duration 2 s: windows of 25 codes (hop 18, overlap 7) / 50 latent frames / 48000 samples
        for n_codes, steps in ((40, 2), (25, 3), (9, 2), (61, 2)):
            codes = torch.randint(0, CB_SIZE, (1, 8, n_codes), generator=g)
            torch.manual_seed(1000 + n_codes)
            wav_ref = tok.token2audio_no_reason(codes, False, duration=2, guidance_scale=1.5, num_steps=steps, disable_progress=True)
            torch.manual_seed(1000 + n_codes)
            draws = []

            def randn(shape):
                t = torch.randn(*shape)
                draws.append(t)
                return t

            wav = orc.token2audio_no_reason(codes, duration=2, num_steps=steps, randn=randn)
            assert wav_ref.shape == wav.shape and torch.equal(wav_ref, wav), f"n_codes {n_codes}: oracle != reference ({(wav_ref - wav).abs().max()})"
            out["cases"].append(dict(codes=codes, steps=steps, duration=2, draws=draws, wav=wav_ref))
            print(f"[ok] {n_codes} codes, {steps} steps: {len(draws)} noise draws, waveform {tuple(wav_ref.shape)} bit-exact")
    out["checksum"] = {k: float(v.double().sum()) for k, v in p.items()}
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
