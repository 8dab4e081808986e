"""Generate tests/golden/dit_golden.pt from the UNMODIFIED in-repo flow-matching decoder (transformer_1d_flow.py,
attention.py imported over oracle/diffusers_stub.py; class BASECFM executed from the unmodified source text of
AudioDiffusion1D.py, whose module-level imports - whisper, peft, fairseq ... - are not installable here) and assert that
oracle/dit_oracle.py is bit-identical to them on CPU.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_dit
"""
import ast
import os
import sys
from abc import ABC

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dit_oracle as DO  # noqa: E402
from oracle.diffusers_stub import install_diffusers_stub  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "dit_golden.pt")
THREADS = 4


def dit_cfgs():
    return {
        # 2 heads x 64 like the production head size; in = latent 8 + incontext 8 + cond 24
        "tiny": DO.DitCfg(num_attention_heads=2, attention_head_dim=64, in_channels=40, out_channels=8, num_layers=2,
                          num_positional_embeddings=64),
        # 4 heads, 3 layers, odd sequence length
        "mid": DO.DitCfg(num_attention_heads=4, attention_head_dim=64, in_channels=72, out_channels=12, num_layers=3,
                         num_positional_embeddings=96),
    }


def load_basecfm():
    """Execute the unmodified source of `class BASECFM` (AudioDiffusion1D.py:62-167) in a namespace holding only what its
    body needs."""
    path = os.path.join(REF_ROOT, "tools", "tokenizer", "ReasoningCodec_film", "models", "AudioDiffusion1D.py")
    src = open(path).read()
    node = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "BASECFM"][0]
    text = "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
    ns = {"torch": torch, "ABC": ABC, "tqdm": lambda it: it, "F": torch.nn.functional}
    exec(compile(text, path, "exec"), ns)
    return ns["BASECFM"]


def main():
    torch.set_num_threads(THREADS)
    install_diffusers_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    BASECFM = load_basecfm()
    out = {}
    with torch.no_grad():
        for name, cfg in dit_cfgs().items():
            sd = DO.random_state_dict(cfg, seed=909)
            ref = Transformer1DModel(**cfg.ctor_kwargs())
            full = ref.state_dict()
            assert set(full.keys()) == set(sd.keys()), sorted(set(full) ^ set(sd))
            assert torch.equal(full["pos_embed.pe"], sd["pos_embed.pe"]), "sinusoidal table restated differently"
            ref.load_state_dict(sd, strict=True)
            ref = ref.float().eval()
            orc = DO.DitOracle(cfg, sd)
            g = torch.Generator().manual_seed(17)
            kw = {"resolution": None, "aspect_ratio": None}
            cases = []
            for B, T in ((2, 10), (1, 33), (3, 17)):
                x = torch.randn(B, T, cfg.in_channels, generator=g)
                t = torch.rand(1, generator=g).repeat(B)
                y_ref = ref(x, timestep=t, added_cond_kwargs=kw).sample
                y = orc.forward(x, t)
                assert torch.equal(y_ref, y), f"{name}: estimator oracle != reference ({(y_ref - y).abs().max()})"
                cases.append(dict(x=x, t=t, y=y_ref))
            # Euler solver with classifier-free guidance (reason_tokenizer.py:273: guidance_scale 1.5).  Batch 1 only: the
            # reference repeats the timestep twice (AudioDiffusion1D.py:114), so a larger batch fails inside the estimator
            cfm = BASECFM(ref)
            lat = cfg.out_channels
            cond = cfg.in_channels - 2 * lat
            solves = []
            for B, T, ic, steps in ((1, 20, 6, 4), (1, 13, 0, 3)):
                z = torch.randn(B, T, lat, generator=g)
                incontext = torch.randn(B, T, lat, generator=g)
                incontext[:, ic:] = 0
                mu = torch.randn(B, T, cond, generator=g)
                t_span = torch.linspace(0, 1, steps + 1)
                # inference_codes builds these two tensors (AudioDiffusion1D.py:599-607); the estimator ignores them
                # (use_additional_conditions = False, transformer_1d_flow.py:248)
                akw = {"resolution": torch.tensor([T, 1]).repeat(B, 1).float(), "aspect_ratio": torch.tensor([T / 1500.0]).repeat(B, 1)}
                r = cfm.solve_euler(z.clone(), incontext, ic, t_span, mu, akw, 1.5)
                o = orc.solve_euler(z.clone(), incontext, ic, t_span, mu, 1.5)
                assert torch.equal(r, o), f"{name}: solver oracle != reference ({(r - o).abs().max()})"
                solves.append(dict(z=z, incontext=incontext, incontext_length=ic, mu=mu, steps=steps, guidance_scale=1.5, out=r))
            out[name] = dict(cases=cases, solves=solves)
            out[f"__checksum_{name}"] = {k: float(v.double().sum()) for k, v in sd.items()}
            print(f"[ok] {name}: {len(cases)} estimator calls and {len(solves)} Euler solves bit-exact")
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
