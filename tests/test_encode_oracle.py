"""CPU: the restated encode chain of ReasoningCodec_film (oracle/encode_oracle.py) reproduces the fixtures that
oracle/make_golden_encode.py produced by executing the UNMODIFIED source of AudioDiffusion1D.fetch_codes_batch / time_film
(AudioDiffusion1D.py:428-438, :492-551) on stand-in SSL features - codes bit-equal, features bit-equal."""
import os

import pytest
import torch

from oracle import encode_oracle as EO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    """Bit-exact comparisons run with the thread count oracle/make_golden_encode.py used (ATen partitions sums by it); restored after."""
    n = torch.get_num_threads()
    torch.set_num_threads(4)
    yield
    torch.set_num_threads(n)


def test_encode_chain_matches_reference_golden():
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "encode_golden.pt"), weights_only=False)
    p = EO.random_params(gold["param_seed"])
    for c in gold["cases"]:
        feats = EO.stand_in_features(c["feat_seed"], *c["shape"])
        with torch.no_grad():
            codes, merge = EO.fetch_codes_from_features(p, film_masks=[m.float() for m in c["film_masks"]], **feats)
        assert codes.shape == c["codes"].shape == (c["shape"][0], c["shape"][1] // 4, 8)
        assert torch.equal(codes, c["codes"])
        assert torch.equal(merge, c["merge"])


def test_residual_vq_restatement_properties():
    """The (unpinned) ResidualVQ restatement: indices are the Euclidean nearest neighbours of the running residual and the output is
    project_out of the summed code vectors (the decode-side identity that tests/test_detok_oracle.py pins from the other direction)."""
    p = EO.random_params(5)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 7, EO.CODEC_DIM, generator=g)
    q, idx = EO.residual_vq_forward(x, p, "vq_acoustic", 6)
    h = torch.nn.functional.linear(x, p["vq_acoustic.project_in.weight"], p["vq_acoustic.project_in.bias"])
    r, total = h, torch.zeros_like(h)
    for i in range(6):
        e = p["vq_acoustic.codebooks"][i]
        nn = torch.cdist(r, e.unsqueeze(0).expand(2, -1, -1)).argmin(-1)
        assert torch.equal(nn, idx[..., i])
        r, total = r - e[nn], total + e[nn]
    ref = torch.nn.functional.linear(total, p["vq_acoustic.project_out.weight"], p["vq_acoustic.project_out.bias"])
    assert torch.allclose(q, ref, atol=1e-6)
