import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "llm_golden.pt"), weights_only=False)


def build_product_model_cpu(cfg):
    """The drop-in Model_stage3 for an oracle Stage3Cfg with its parameters on the CPU (no handle, no kernels)."""
    return _make_product_model(cfg, "cpu", None)


def build_product_model(cfg, sd, device, max_batch, max_seq=None):
    """Instantiate the product's Model_stage3 for an oracle Stage3Cfg and load the oracle's state dict."""
    m = _make_product_model(cfg, device, max_seq)
    m.load_state_dict(sd, strict=True)
    m.setup_caches(max_batch)
    return m


def _make_product_model(cfg, device, max_seq):
    from uniaudio2_b200.llm_models import config as pc
    from uniaudio2_b200.llm_models.model_new import Model_stage3, ModelArgs

    def d(c, name):
        return dict(name=name, n_layer=c.n_layer, n_embd=c.n_embd, n_head=c.n_head, n_query_groups=c.n_query_groups,
                    head_size=c.head_size, intermediate_size=c.intermediate_size, padded_vocab_size=c.padded_vocab_size,
                    norm_eps=c.norm_eps, rope_base=c.rope_base, rope_adjustments=c.rope_adjustments)

    pc.name_to_config["ua2-test-backbone"] = d(cfg.backbone, "ua2-test-backbone")
    pc.name_to_config["ua2-test-decoder"] = d(cfg.decoder, "ua2-test-decoder")
    saved = {k: pc.name_to_config[k] for k in ("Llama-3.2-Understanding", "Llama-3.2-Generation")}
    pc.name_to_config["Llama-3.2-Understanding"] = d(cfg.understanding, "Llama-3.2-Understanding")
    pc.name_to_config["Llama-3.2-Generation"] = d(cfg.generation, "Llama-3.2-Generation")
    try:
        args = ModelArgs("ua2-test-backbone", "ua2-test-decoder", "", "", "", cfg.audio_vocab - 7, 7, cfg.num_codebooks)
        m = Model_stage3(args, device=device, max_seq_length=max_seq or cfg.max_seq_length)
    finally:
        pc.name_to_config.update(saved)
    return m
