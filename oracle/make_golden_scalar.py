"""Pin oracle/scalar_oracle.py against the real ScalarModel (scalar24k.py) imported with stubs; write tests/golden/scalar_golden.pt.

    python -m oracle.make_golden_scalar
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scalar_oracle as SO  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "scalar_golden.pt")


def scalar_cfgs():
    return {
        "causal": SO.ScalarCfg(causal=True, downsample_factors=[2, 3, 4], downsample_kernel_sizes=[4, 6, 8], upsample_factors=[4, 3, 2],
                               upsample_kernel_sizes=[8, 6, 4], latent_hidden_dim=24, init_channel=8),
        "noncausal": SO.ScalarCfg(causal=False, downsample_factors=[2, 4], downsample_kernel_sizes=[4, 8], upsample_factors=[4, 2],
                                  upsample_kernel_sizes=[8, 4], latent_hidden_dim=136, init_channel=16),
    }


def import_reference():
    for name in ("pytorch_lightning", "omegaconf"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "pytorch_lightning":
                m.LightningModule = torch.nn.Module
            else:
                m.OmegaConf = type("OmegaConf", (), {})
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "ref_scalar24k", os.path.join(REF_ROOT, "tools", "tokenizer", "ReasoningCodec_film", "models", "scalar24k.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.set_num_threads(8)
    mod = import_reference()
    out = {}
    for name, cfg in scalar_cfgs().items():
        torch.manual_seed(77)
        ref = mod.ScalarModel(cfg.num_bands, cfg.sample_rate, cfg.causal, cfg.num_samples, cfg.downsample_factors,
                              cfg.downsample_kernel_sizes, cfg.upsample_factors, cfg.upsample_kernel_sizes, cfg.latent_hidden_dim,
                              cfg.default_kernel_size, cfg.delay_kernel_size, cfg.init_channel, cfg.res_kernel_size).eval()
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        g = torch.Generator().manual_seed(3)
        for k in sd:  # make weight_g / PReLU slopes non-trivial
            if k.endswith("weight_g"):
                sd[k] = sd[k] * (0.5 + torch.rand(sd[k].shape, generator=g))
            if "activation" in k:
                sd[k] = 0.1 + 0.3 * torch.rand(sd[k].shape, generator=g)
        # DownsampleLayer's `activation=nn.PReLU()` default argument (scalar24k.py:196) is ONE module instance shared by every
        # encoder block, so all `down_conv.activation.weight` entries alias the same parameter: keep them equal
        shared = [k for k in sd if k.endswith("down_conv.activation.weight")]
        for k in shared:
            sd[k] = sd[shared[0]].clone()
        ref.load_state_dict(sd)
        hop = 1
        for s in cfg.upsample_factors:
            hop *= s
        z = torch.rand(2, cfg.latent_hidden_dim, 13, generator=g) * 2 - 1
        wav = torch.randn(2, 1, hop * 11 + (0 if not cfg.causal else 0), generator=g) * 0.3
        with torch.no_grad():
            y_ref = ref.decode(z)
            e_ref = ref.encode(wav)
            y_o = SO.scalar_decode(z, sd, cfg)
            e_o = SO.scalar_encode(wav, sd, cfg)
        assert torch.equal(y_ref, y_o), f"{name}: decode mismatch {(y_ref - y_o).abs().max()}"
        assert torch.equal(e_ref, e_o), f"{name}: encode mismatch {(e_ref - e_o).abs().max()}"
        out[name] = dict(sd=sd, z=z, wav=wav, decoded=y_ref, encoded=e_ref)
        print(f"[ok] {name}: decode {tuple(y_ref.shape)} encode {tuple(e_ref.shape)} bit-exact vs scalar24k.ScalarModel")
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
