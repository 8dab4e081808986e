"""Wire / disk formats of the inference path: reference-format checkpoints load into the
drop-in Model_stage3 by the reference's own key names, `llm_config.yaml` maps to ModelArgs like multi_task_inference.py,
token files round-trip.  CPU only (parameters live on the CPU until `.to('cuda')` + `setup_caches`)."""
import argparse
import os
import sys
import time

import pytest
import torch

from oracle import llm_oracle as O
from oracle.cases import tiny_cfgs
from conftest import build_product_model_cpu
from uniaudio2_b200.llm_utils import train_utils as TU


def test_reference_format_checkpoint_loads(tmp_path):
    cfg = tiny_cfgs()["tiny"]
    sd = O.random_state_dict(cfg, seed=11)
    # what the reference's save_checkpoint writes on rank 0 (llm_utils/train_utils.py:179-195), with DDP-style prefixes
    ck = {"model": {"module." + k: v for k, v in sd.items()}, "optimizer": {}, "scheduler": {}, "reporter": {}}
    p1 = tmp_path / "ep1.checkpoint"
    torch.save(ck, p1)
    time.sleep(0.05)
    sd2 = {k: v + 1.0 for k, v in sd.items()}
    p2 = tmp_path / "ep2.checkpoint"
    torch.save({"model": sd2}, p2)
    m = build_product_model_cpu(cfg)
    used = TU.resume_for_inference(str(p1), None, m, "cpu")
    assert used == str(p1)
    got = m.state_dict()
    assert set(got.keys()) == set(sd.keys())
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # newest ep*.checkpoint of exp_dir when no path is given
    used = TU.resume_for_inference(None, str(tmp_path), m, "cpu")
    assert used == str(p2)
    assert all(torch.equal(m.state_dict()[k], sd2[k]) for k in sd)
    with pytest.raises(ValueError):
        TU.resume_for_inference(None, str(tmp_path / "nothing_here"), m, "cpu")
    # a checkpoint with a missing key is rejected (strict load, like the reference)
    bad = dict(sd)
    bad.pop(next(iter(bad)))
    torch.save({"model": bad}, tmp_path / "bad.checkpoint")
    with pytest.raises(RuntimeError):
        TU.resume_for_inference(str(tmp_path / "bad.checkpoint"), None, m, "cpu")


@pytest.mark.skipif(not os.path.isdir("/root/reference/llm_utils"), reason="reference tree not present")
def test_reference_loader_accepts_the_drop_in_model(tmp_path):
    """The UNMODIFIED reference function (llm_utils/train_utils.py:159) restores the drop-in model: same keys, strict."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_train_utils", "/root/reference/llm_utils/train_utils.py")
    ref = importlib.util.module_from_spec(spec)
    sys.path.insert(0, "/root/reference")  # its sibling imports (llm_utils.*) resolve against the reference tree
    saved = {k: v for k, v in sys.modules.items() if k == "llm_utils" or k.startswith("llm_utils.")}
    try:
        spec.loader.exec_module(ref)
    except Exception as e:  # optional third-party imports of the training utilities
        pytest.skip(f"reference train_utils not importable here: {e}")
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "llm_utils" or k.startswith("llm_utils.")]:
            if k not in saved:
                del sys.modules[k]
    cfg = tiny_cfgs()["tiny"]
    sd = O.random_state_dict(cfg, seed=12)
    torch.save({"model": {"module." + k: v for k, v in sd.items()}}, tmp_path / "ep3.checkpoint")
    m = build_product_model_cpu(cfg)
    ref.resume_for_inference(None, str(tmp_path), m, "cpu")
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)


def test_llm_config_yaml_to_model_args(tmp_path):
    y = tmp_path / "llm_config.yaml"
    y.write_text("local_model: Llama-3.2-300M\nllm_pretrained_model: ''\nllm_name: Llama-3.2-3B\naudio_semantic_card: 8200\n"
                 "audio_reason_card: 4100\nparallel_number: 9\naudio_embeddings_path: ''\naudio_understanding_expert_path: ''\n"
                 "text_pad_token: 128002\n")
    ta = TU.load_llm_config(str(y))
    assert isinstance(ta, argparse.Namespace) and ta.parallel_number == 9
    ma = TU.model_args_from_config(ta)
    assert (ma.llm_name, ma.decoder_name, ma.audio_num_codebooks) == ("Llama-3.2-3B", "Llama-3.2-300M", 8)
    assert ma.audio_semantic_vocab_size + ma.audio_reason_vocab_size == 12300


def test_token_files_round_trip(tmp_path):
    g = torch.Generator().manual_seed(0)
    reason = torch.randint(0, 4096, (8, 21), generator=g, dtype=torch.int32)
    semantic = torch.randint(0, 8192, (8, 51), generator=g)
    pr, ps = TU.save_token_files(str(tmp_path), "utt1", reason, semantic)
    assert os.path.basename(pr) == "utt1_reason.pt" and os.path.basename(ps) == "utt1_semantic.pt"
    r2, s2 = TU.load_token_files(str(tmp_path), "utt1")
    assert r2.dtype == torch.long and s2.dtype == torch.long
    assert torch.equal(r2, reason.long()) and torch.equal(s2, semantic)
    # exactly what multi_task_inference.py:143-144 writes can be read back
    torch.save(semantic, tmp_path / "utt2_semantic.pt")
    torch.save(reason.long(), tmp_path / "utt2_reason.pt")
    r3, s3 = TU.load_token_files(str(tmp_path), "utt2")
    assert torch.equal(r3, reason.long()) and torch.equal(s3, semantic)
    with pytest.raises(ValueError):
        TU.save_token_files(str(tmp_path), "bad", reason[0], semantic)
