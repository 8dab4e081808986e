"""CPU suite: oracle/film_oracle.py (time_film, feature_combine of AudioDiffusion1D.py:428-456) reproduces the fixtures that
oracle/make_golden_film.py wrote by executing the UNMODIFIED reference source of those methods."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import film_oracle as FO
from oracle.make_golden_film import THREADS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    n = torch.get_num_threads()
    torch.set_num_threads(THREADS)
    yield
    torch.set_num_threads(n)


@pytest.fixture(scope="module")
def film_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "film_golden.pt"), weights_only=False)


def test_time_film_matches_reference(film_golden):
    fx = film_golden["time_film"]
    out = FO.time_film(fx["params"], fx["features"], fx["zero_mask"], fx["gamma_scale"])
    assert torch.equal(out, fx["out"])
    # the draw is the reference's own: torch.rand(B, 1, 1) < 0.2 under the recorded seed
    torch.manual_seed(fx["seed"])
    assert torch.equal((torch.rand(fx["params"].shape[0], 1, 1) < 0.2).view(-1).to(torch.uint8), fx["zero_mask"])
    # zero-conditioned samples pass the features through unchanged
    z = fx["zero_mask"].bool()
    assert z.any() and not z.all()
    assert torch.equal(out[z], fx["features"][z])
    # the formula the GPU test of ua2_film_f32 checks against (tests/test_scalar_gpu.py) is this one
    dg, beta = fx["params"].chunk(2, dim=-1)
    mk = fx["zero_mask"].float().view(-1, 1, 1)
    assert torch.equal(((1.0 + 0.1 * dg.tanh()) * (1 - mk) + 1.0 * mk) * fx["features"] + (beta * (1 - mk) + 0.0 * mk), fx["out"])


def test_feature_combine_matches_reference(film_golden):
    fx = film_golden["feature_combine"]
    for c in fx["cases"]:
        out = FO.feature_combine(fx["weight"], fx["bias"], c["reasoning"], c["rec"])
        assert torch.equal(out, c["out"])
        # nearest x2.5: output frame t reads input frame floor(t / 2.5)  (what ua2_interp_nearest_f32 implements)
        r = F.linear(c["reasoning"], fx["weight"], fx["bias"])
        T = c["rec"].shape[1]
        idx = torch.clamp((torch.arange(T).float() / 2.5).floor().long(), max=r.shape[1] - 1)
        assert torch.equal(c["rec"] + r[:, idx], c["out"])
