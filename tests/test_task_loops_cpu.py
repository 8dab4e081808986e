"""CPU suite: the PRODUCT's task loops (uniaudio2_b200/evaluation: prompt packing, reason -> semantic phase switch, EOS stop,
feedback of the sampled frame, text-decode loop) against the reference's own loops.

tests/golden/tasks_golden.pt was written by oracle/make_golden_tasks.py, which executes the UNMODIFIED `class Generator` of the
reference's evaluation/{tts,musicgen,audiogen,songen,asr,audio_music_caption,lyric_asr}_task.py over the UNMODIFIED reference
Model_stage3 behind a scripted sampler (EOS frames are injected at scripted frame numbers, since random weights never emit
them).  Here the product's Generators run over the CPU oracle model (bit-identical to the reference model,
tests/test_oracle_golden.py) behind the same script: outputs AND the sequence of model calls (positions, forbid_prefix, ...)
must be identical.  No GPU, no kernels: this checks the host-side logic only."""
import os

import pytest
import torch

from oracle import llm_oracle as O
from oracle.cases import tiny_cfgs
from oracle.make_golden_tasks import GEN_CASES, SPECIALS, TEXT_CASES, THREADS, ScriptedModel, train_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    n = torch.get_num_threads()
    torch.set_num_threads(THREADS)
    yield
    torch.set_num_threads(n)


@pytest.fixture(scope="module")
def tasks_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "tasks_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def oracle_model():
    cfg = tiny_cfgs()["tiny"]
    return O.Stage3Oracle(cfg, O.random_state_dict(cfg, seed=1234))


@pytest.mark.parametrize("name", [c[0] for c in GEN_CASES])
def test_generation_loops_match_reference(tasks_golden, oracle_model, name):
    from uniaudio2_b200.evaluation import tts_task

    fx = tasks_golden[name]
    args = train_args()
    sm = ScriptedModel(oracle_model, fx["script"], args, oracle_api=True)
    gen = tts_task.Generator(sm, args, is_cfg=fx["is_cfg"], tag=fx["tag"])
    gen.special_token_dict = dict(SPECIALS)
    method = {"generate_tts": gen.generate_tts, "generate_audio": gen.generate_audio, "generate_LTS": gen.generate_LTS}[fx["method"]]
    torch.manual_seed(2024)
    r, s = method(fx["prompt"], name, text_token=fx["text"], pinned_staging=False, **fx["sampling"])
    assert torch.equal(r.cpu(), fx["reason"].to(torch.int64)) and torch.equal(s.cpu(), fx["semantic"].to(torch.int64))
    assert sm.calls == fx["calls"]  # same prefill shape, same (frame, position, maxp1, temperature, topk, forbid_prefix) sequence


@pytest.mark.parametrize("name", [c[0] for c in TEXT_CASES])
def test_text_decode_loops_match_reference(tasks_golden, oracle_model, name):
    from uniaudio2_b200.evaluation import asr_task

    fx = tasks_golden[name]
    args = train_args()
    sm = ScriptedModel(oracle_model, fx["script"], args, oracle_api=True)
    gen = asr_task.Generator(sm, args)
    torch.manual_seed(2025)
    ids = getattr(gen, fx["method"])(fx["prompt"], name, semantic_token=fx["semantic_in"], reason_token=fx["reason_in"], **fx["sampling"])
    assert ids == fx["ids"]
    assert sm.calls == fx["calls"]


def _cond_names():
    from oracle.make_golden_tasks import COND_CASES

    return [c[0] for c in COND_CASES]


@pytest.mark.parametrize("name", _cond_names())
def test_condition_sequence_tasks_match_reference(tasks_golden, oracle_model, name):
    """instruct TTS, audio understanding (incl. an audio_prompt entry), speech-to-text, speech-to-speech: the generic
    condition-sequence packer + the two loops (insturct_tts_task.py, audio_understanding.py, speech_s2t.py, speech_s2s.py)."""
    from uniaudio2_b200.evaluation import audio_understanding, tts_task

    fx = tasks_golden[name]
    args = train_args()
    sm = ScriptedModel(oracle_model, fx["script"], args, oracle_api=True)
    torch.manual_seed(2026)
    if fx["kind"] == "instruct":
        gen = tts_task.Generator(sm, args)
        gen.special_token_dict = dict(SPECIALS)
        r, s = gen.generate_instruct_tts(fx["prompt"], name, text_token=fx["text"], caption=fx["caption"], pinned_staging=False, **fx["sampling"])
        assert torch.equal(r.cpu(), fx["reason"].to(torch.int64)) and torch.equal(s.cpu(), fx["semantic"].to(torch.int64))
    elif fx["kind"] == "audio":
        gen = audio_understanding.AudioGenerator(sm, args)
        gen.special_token_dict = dict(SPECIALS)
        r, s = gen.generate_audio(fx["prompt"], name, d=fx["d"], keys=fx["keys"], types=fx["types"], pinned_staging=False, **fx["sampling"])
        assert torch.equal(r.cpu(), fx["reason"].to(torch.int64)) and torch.equal(s.cpu(), fx["semantic"].to(torch.int64))
    else:
        gen = audio_understanding.Generator(sm, args)
        gen.special_token_dict = dict(SPECIALS)
        res = gen.generate_answer(fx["prompt"], name, d=fx["d"], keys=fx["keys"], types=fx["types"], speech_s2t=fx["kind"] == "s2t", **fx["sampling"])
        if fx["kind"] == "s2t":
            assert res[0] == fx["ids"] and res[1] == fx["prompt_len"]
        else:
            assert res == fx["ids"]
    assert sm.calls == fx["calls"]


def test_speech_s2t_refuses_long_prompts(oracle_model):
    """speech_s2t.py:352-353: prompts of >= 1500 frames return (-1, -1) before touching the model."""
    from uniaudio2_b200.evaluation import audio_understanding

    args = train_args()
    sm = ScriptedModel(oracle_model, {}, args, oracle_api=True)
    gen = audio_understanding.Generator(sm, args)
    d = {"semantic_seq": torch.zeros(8, 1500, dtype=torch.long)}
    assert gen.generate_answer(torch.tensor([5, 6]), "s2t", d=d, keys=["semantic_seq"], types=["audio"], speech_s2t=True) == (-1, -1)
    assert sm.calls == []
