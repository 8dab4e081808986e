"""Generate tests/golden/tasks_golden.pt: the UNMODIFIED task loops of the reference (class Generator of evaluation/tts_task.py,
musicgen_task.py, audiogen_task.py, songen_task.py, asr_task.py, audio_music_caption_task.py, lyric_asr_task.py - executed
from their source text, because the modules import torchaudio / huggingface_hub / the codec stack) driving the UNMODIFIED
reference Model_stage3 (imported through oracle/ref_shims) on a tiny seeded configuration.  TEST INFRASTRUCTURE ONLY.

Random weights never emit an end-of-phase frame, so the model is wrapped by a SCRIPTED sampler: the wrapper calls the real
generate_frame and, at scripted frame numbers, overwrites the sampled row with the reason-EOS frame / the semantic-EOS frame /
the end-of-text token.  The fixture records prompts, scripts and the loops' outputs; tests/test_task_loops_cpu.py replays them
through the PRODUCT's Generator classes (uniaudio2_b200/evaluation) over the CPU oracle model wrapped by the same script and
requires identical tokens - the product's host-side loop logic (prompt packing, phase switch, EOS, CFG batch, feedback) is
thereby checked against the reference's own loops without a GPU.

    python -m oracle.make_golden_tasks
"""
import ast
import os
import sys
import textwrap
import types
from typing import List, Tuple

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llm_oracle as O  # noqa: E402
from oracle.cases import REASON_CARD, tiny_cfgs  # noqa: E402
from oracle.ref_shims import REF_ROOT, install_llm_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "tasks_golden.pt")
THREADS = 4
EOS_TEXT = 128001


def train_args(cfg_name="tiny"):
    """Token ids inside the tiny vocabularies (reason card 37+..., see oracle/cases.py); same fields as llm_config.yaml."""
    return types.SimpleNamespace(text_pad_token=3, semantic_pad_token=80, semantic_eos=81, semantic_bos=82, reason_eos=37, reason_bos=38,
                                 reason_pad_token=36, parallel_number=9, audio_reason_card=REASON_CARD[cfg_name], audio_semantic_card=90,
                                 audio_prompt_bos=83, audio_prompt_eos=84)


SPECIALS = {k: 900 + i for i, k in enumerate(['<think>', '</think>', '</answer>', '<transcription>', '</transcription>', '<lyric>', '</lyric>',
                                              '<caption>', '</caption>', '<answer>', '<reason_token>', '<semantic_token>'])}


class ScriptedModel:
    """Wraps a model with the reference's Model_stage3 call surface; overwrites the sampled frame at scripted frame numbers.
    script: {frame_index: 'reason_eos' | 'end' | 'eot'}.  Accepts input_pos as a tensor (reference loops) or an int (product loops)."""

    def __init__(self, model, script, args, oracle_api=False):
        self.m, self.script, self.a, self.oracle_api = model, dict(script), args, oracle_api
        self.frame = 0
        self.calls = []

    def parameters(self):
        if self.oracle_api:
            return iter([torch.zeros(1)])
        return self.m.parameters()

    def setup_caches(self, b):
        self.m.setup_caches(b)

    def reset_caches(self):
        self.frame = 0
        self.m.reset_caches()

    def forward_prefix(self, tokens, labels=None, tokens_mask=None, loss_mask=None, input_pos=None, input_pos_maxp1=None):
        self.calls.append(("prefix", tuple(tokens.shape)))
        if self.oracle_api:
            return self.m.forward_prefix(tokens, tokens_mask, input_pos)
        return self.m.forward_prefix(tokens, labels=labels, tokens_mask=tokens_mask, loss_mask=loss_mask, input_pos=input_pos)

    def generate_frame(self, tokens, tokens_mask, input_pos=None, input_pos_maxp1=None, temperature=1.0, topk=1, forbid_prefix=0,
                       cfg_scale=1.0, **kw):
        pos = input_pos if torch.is_tensor(input_pos) else torch.tensor([int(input_pos)])
        if self.oracle_api:
            s = self.m.generate_frame(tokens, tokens_mask, pos, input_pos_maxp1, temperature, topk, forbid_prefix, cfg_scale)
        else:
            s = self.m.generate_frame(tokens, tokens_mask, input_pos=pos, input_pos_maxp1=input_pos_maxp1, temperature=temperature,
                                      topk=topk, forbid_prefix=forbid_prefix, cfg_scale=cfg_scale)
        s = s.clone()
        what = self.script.get(self.frame)
        if what == "reason_eos":
            s[:, 1:] = self.a.reason_eos
        elif what == "end":
            s[:, 1:] = self.a.semantic_eos + self.a.audio_reason_card
        elif what == "eot":
            s[:, 0] = EOS_TEXT
        self.calls.append(("frame", self.frame, int(pos[0]), int(input_pos_maxp1), float(temperature), int(topk), int(forbid_prefix)))
        self.frame += 1
        return s


def load_reference_generator(task_file):
    """class Generator of evaluation/<task_file>.py, executed from its unmodified source text."""
    path = os.path.join(REF_ROOT, "evaluation", task_file + ".py")
    src = open(path).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "Generator"][0]
    text = textwrap.dedent("\n".join(src.splitlines()[cls.lineno - 1:cls.end_lineno]))
    ns = {"torch": torch, "Tuple": Tuple, "List": List, "Model": object, "Segment": object,
          "load_text_tokenizer": lambda p: None, "load_audio_tokenizer": lambda **k: None}
    exec(compile(text, path, "exec"), ns)
    return ns["Generator"]


# (case name, task file, method, tag of the product Generator, kind, is_cfg, script, sampling)
GEN_CASES = [
    ("tts_greedy", "tts_task", "generate_tts", "transcription", False, {3: "reason_eos", 9: "end"}, dict(temperature=1.0, topk=1)),
    ("tts_topk_seeded", "tts_task", "generate_tts", "transcription", False, {4: "reason_eos", 8: "end"}, dict(temperature=0.9, topk=5)),
    # (is_cfg=True is not exercised: the reference loop itself fails there - it keeps row 0 of the sample and then concatenates
    #  a (1, 8) mask with a (2, 1) one, tts_task.py:257-275 - and its CLI never enables it, multi_task_inference.py:486-493)
    ("ttm", "musicgen_task", "generate_audio", "caption", False, {3: "reason_eos", 8: "end"}, dict(temperature=1.0, topk=1)),
    ("tta", "audiogen_task", "generate_audio", "caption", False, {5: "reason_eos", 9: "end"}, dict(temperature=0.8, topk=3)),
    ("lts", "songen_task", "generate_LTS", "lyric", False, {3: "reason_eos", 7: "end"}, dict(temperature=1.0, topk=1)),
]
TEXT_CASES = [
    ("asr_greedy", "asr_task", "generate_asr", {6: "eot"}, dict(temperature=1.0, topk=1)),
    ("asr_topk_seeded", "asr_task", "generate_asr", {5: "eot"}, dict(temperature=0.9, topk=4)),
    ("caption", "audio_music_caption_task", "generate_audio_caption", {7: "eot"}, dict(temperature=1.0, topk=1)),
    ("lyric_asr", "lyric_asr_task", "generate_lyric_asr", {4: "eot"}, dict(temperature=1.0, topk=1)),
]


# generic condition-sequence tasks: (case name, task file, method, kind, script, sampling, keys, types)
COND_CASES = [
    ("instruct_tts", "insturct_tts_task", "generate_instruct_tts", "instruct", {3: "reason_eos", 8: "end"}, dict(temperature=1.0, topk=1), None, None),
    ("audio_understanding", "audio_understanding", "generate_answer", "text", {5: "eot"}, dict(temperature=1.0, topk=1),
     ["reason_seq", "semantic_seq", "text_seq"], ["audio", "audio", "text"]),
    ("audio_understanding_prompt", "audio_understanding", "generate_answer", "text", {4: "eot"}, dict(temperature=0.9, topk=3),
     ["semantic_seq_prompt", "transcription_seq", "reason_seq"], ["audio_prompt", "text", "audio"]),
    ("speech_s2t", "speech_s2t", "generate_answer", "s2t", {6: "eot"}, dict(temperature=1.0, topk=1),
     ["reason_seq", "semantic_seq", "text_seq"], ["audio", "audio", "text"]),
    ("speech_s2s", "speech_s2s", "generate_audio", "audio", {2: "reason_eos", 7: "end"}, dict(temperature=1.0, topk=1),
     ["reason_seq", "semantic_seq", "caption_seq"], ["audio", "audio", "text"]),
]


class IdTokenizer:
    """The text-decode loops end with self._text_tokenizer.decode(ids): keep the ids."""

    @staticmethod
    def decode(ids):
        return [int(i) for i in ids]


def main():
    torch.set_num_threads(THREADS)
    model_new = install_llm_shims()
    from oracle.make_golden import build_reference

    cfg = tiny_cfgs()["tiny"]
    sd = O.random_state_dict(cfg, seed=1234)
    ref = build_reference(model_new, cfg, sd)
    args = train_args()
    g = torch.Generator().manual_seed(77)
    out = {}
    for name, task_file, method, tag, is_cfg, script, samp in GEN_CASES:
        Gen = load_reference_generator(task_file)
        sm = ScriptedModel(ref, script, args)
        gen = Gen(sm, args, None, None, None, is_cfg)
        gen.special_token_dict = dict(SPECIALS)
        prompt = torch.randint(4, 800, (5,), generator=g)
        text = torch.randint(4, 800, (6,), generator=g)
        torch.manual_seed(2024)
        r, s = getattr(gen, method)(prompt, name, text_token=text, **samp)
        out[name] = dict(task_file=task_file, method=method, tag=tag, is_cfg=is_cfg, script=script, sampling=samp, prompt=prompt, text=text,
                         reason=r.clone(), semantic=s.clone(), calls=list(sm.calls))
        print(f"[ok] {name}: reason {tuple(r.shape)} semantic {tuple(s.shape)} over {sm.frame} frames")
    for name, task_file, method, script, samp in TEXT_CASES:
        Gen = load_reference_generator(task_file)
        sm = ScriptedModel(ref, script, args)
        gen = Gen(sm, args, None, None, None, False)
        gen.special_token_dict = dict(SPECIALS)
        gen._text_tokenizer = IdTokenizer()
        prompt = torch.randint(4, 800, (5,), generator=g)
        reason = torch.randint(0, 36, (4, 8), generator=g)
        sem = torch.randint(0, 80, (6, 8), generator=g)
        torch.manual_seed(2025)
        ids = getattr(gen, method)(prompt, name, semantic_token=sem, reason_token=reason, **samp)
        out[name] = dict(task_file=task_file, method=method, script=script, sampling=samp, prompt=prompt, reason_in=reason, semantic_in=sem,
                         ids=list(ids), calls=list(sm.calls))
        print(f"[ok] {name}: {len(ids)} text ids over {sm.frame} frames")
    for name, task_file, method, kind, script, samp, keys, types_ in COND_CASES:
        Gen = load_reference_generator(task_file)
        sm = ScriptedModel(ref, script, args)
        gen = Gen(sm, args, None, None, None, False)
        gen.special_token_dict = dict(SPECIALS)
        gen._text_tokenizer = IdTokenizer()
        prompt = torch.randint(4, 800, (5,), generator=g)
        fx = dict(task_file=task_file, method=method, kind=kind, script=script, sampling=samp, prompt=prompt, keys=keys, types=types_)
        torch.manual_seed(2026)
        if kind == "instruct":
            fx["text"], fx["caption"] = torch.randint(4, 800, (6,), generator=g), torch.randint(4, 800, (4,), generator=g)
            r, s_ = getattr(gen, method)(prompt, name, text_token=fx["text"], caption=fx["caption"], **samp)
            fx.update(reason=r.clone(), semantic=s_.clone())
        else:
            d = {}
            for k, tp in zip(keys, types_):
                if tp == "text":
                    d[k] = torch.randint(4, 800, (5,), generator=g)
                elif k.startswith("reason_seq"):
                    d[k] = torch.randint(0, 36, (8, 4), generator=g)   # (8, T): the loaders hand codes over codebook-major
                else:
                    d[k] = torch.randint(0, 80, (8, 6), generator=g)
            fx["d"] = d
            res = getattr(gen, method)(prompt, name, d=d, keys=keys, types=types_, **samp)
            if kind == "audio":
                fx.update(reason=res[0].clone(), semantic=res[1].clone())
            elif kind == "s2t":
                fx.update(ids=list(res[0]), prompt_len=int(res[1]))
            else:
                fx.update(ids=list(res))
        fx["calls"] = list(sm.calls)
        out[name] = fx
        print(f"[ok] {name}: {sm.frame} frames, prefill {sm.calls[0][1]}")
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e3, "KB")


if __name__ == "__main__":
    main()
