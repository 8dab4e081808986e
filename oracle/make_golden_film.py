"""Generate tests/golden/film_golden.pt by executing the UNMODIFIED source of AudioDiffusion1D.time_film / feature_combine
(tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py:428-456) on a stand-in `self`, and assert that
oracle/film_oracle.py is bit-identical.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_film
"""
import ast
import os
import sys
import textwrap
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import film_oracle as FO  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "film_golden.pt")
THREADS = 4


def load_methods(names):
    path = os.path.join(REF_ROOT, "tools", "tokenizer", "ReasoningCodec_film", "models", "AudioDiffusion1D.py")
    src = open(path).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "AudioDiffusion1D"][0]
    ns = {"torch": torch, "F": F, "nn": nn}
    out = {}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in names:
            text = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
            exec(compile(text, path, "exec"), ns)
            out[fn.name] = ns[fn.name]
    assert set(out) == set(names), sorted(set(names) - set(out))
    return out


def main():
    torch.set_num_threads(THREADS)
    m = load_methods(["time_film", "feature_combine"])
    g = torch.Generator().manual_seed(23)
    out = {}
    with torch.no_grad():
        # ---- time_film: B = 6 so that the seeded 20 % draw hits both branches
        B, T, C, Kin = 6, 50, 64, 96
        layer = nn.Linear(Kin, 2 * C)
        cond = torch.randn(B, T, Kin, generator=g)
        feat = torch.randn(B, T, C, generator=g)
        self_ = types.SimpleNamespace(gamma=0.1)
        for seed in range(40):  # a seed whose draw zeroes some but not all samples
            torch.manual_seed(seed)
            mask = (torch.rand(B, 1, 1) < 0.2).float()
            if 0 < int(mask.sum()) < B:
                break
        torch.manual_seed(seed)
        ref = m["time_film"](self_, cond, feat, layer)
        params = layer(cond)
        got = FO.time_film(params, feat, mask.view(-1), 0.1)
        assert torch.equal(ref, got), "time_film restatement differs from the reference source"
        out["time_film"] = dict(params=params, features=feat, zero_mask=mask.view(-1).to(torch.uint8), gamma_scale=0.1, out=ref, seed=seed)
        print(f"[ok] time_film: zero-condition draw {mask.view(-1).int().tolist()} (seed {seed}) bit-exact")
        # ---- feature_combine: T not a multiple of 2.5 * T_q on either side (crop and exact fit)
        cases = []
        adaptor = nn.Linear(48, 48)
        self2 = types.SimpleNamespace(reason_adaptor=adaptor)
        for Tq, T in ((10, 25), (10, 23), (7, 17)):
            rf = torch.randn(2, Tq, 48, generator=g)
            rec = torch.randn(2, T, 48, generator=g)
            ref = m["feature_combine"](self2, rf, rec)
            got = FO.feature_combine(adaptor.weight, adaptor.bias, rf, rec)
            assert torch.equal(ref, got), "feature_combine restatement differs from the reference source"
            cases.append(dict(reasoning=rf, rec=rec, out=ref))
        out["feature_combine"] = dict(weight=adaptor.weight.detach().clone(), bias=adaptor.bias.detach().clone(), cases=cases)
        print(f"[ok] feature_combine: {len(cases)} cases bit-exact")
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e3, "KB")


if __name__ == "__main__":
    main()
