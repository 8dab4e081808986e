"""Task driver for audio-tokens -> text (ASR; audio / music caption and speech_s2t use the same loop).

Mirrors the `Generator` of the reference's evaluation/asr_task.py (:175-688; prepare_asr_task :299-326, generate_asr
:630-688) and evaluation/audio_music_caption_task.py (:201-258, identical loop): the prompt is
[task prompt text | reason_bos, reason codes, reason_eos | semantic_bos, semantic codes + reason_card, semantic_eos],
one forward_prefix call, then greedy / top-k text decoding with the audio slots fed back as zeros until token 128001.

The local decoder still runs every frame in `generate_frame` (the reference discards its output, asr_task.py:671-673);
so does this drop-in, to return bit-identical (B, 9) frames.
"""
from typing import Optional

import torch

from .tts_task import SPECIAL_TOKENS, default_train_args  # noqa: F401

EOS_TEXT = 128001  # Llama-3 <|end_of_text|>, asr_task.py:674


class Generator:
    def __init__(self, model, train_args, audio_tokenizer_config=None, audio_model_path=None, text_tokenizer_path=None,
                 is_cfg: bool = False, text_tokenizer=None):
        self._model = model
        self._model.setup_caches(2 if is_cfg else 1)  # asr_task.py:186-189
        self.is_cfg = is_cfg
        self._text_tokenizer = text_tokenizer
        self.device = next(model.parameters()).device
        self.sample_rate = 24000
        self.empty_token = 0
        self.semantic_eos = train_args.semantic_eos
        self.semantic_bos = train_args.semantic_bos
        self.reason_eos = train_args.reason_eos
        self.reason_bos = train_args.reason_bos
        self.parallel_number = train_args.parallel_number
        self.audio_reason_card = train_args.audio_reason_card

    def text_pad(self, x):
        seq = torch.ones((len(x), self.parallel_number)).to(torch.int64)
        seq[:, -1] = x
        seq[:, :-1] = seq[:, :-1] * self.empty_token
        return seq

    def audio_pad(self, x):
        seq = torch.ones((x.shape[0], self.parallel_number)).to(torch.int64) * self.empty_token
        seq[:, :-1] = x
        return seq

    def prepare_asr_task(self, task_prompt, reason_data, semantic_data):
        """asr_task.py:299-326.  reason_data (T_r, 8), semantic_data (T_s, 8) int64."""
        text = self.text_pad(task_prompt)
        tmask = torch.zeros((text.shape[0], self.parallel_number))
        tmask[:, -1] = True
        nq = reason_data.shape[1]
        reason = torch.cat([torch.ones(1, nq) * self.reason_bos, reason_data, torch.ones(1, nq) * self.reason_eos], dim=0)
        sem = torch.cat([torch.ones(1, nq) * self.semantic_bos, semantic_data, torch.ones(1, nq) * self.semantic_eos], dim=0)
        sem = sem + self.audio_reason_card
        audio = self.audio_pad(torch.cat([reason, sem], dim=0))
        amask = torch.zeros((audio.shape[0], self.parallel_number))
        amask[:, :-1] = True
        return torch.cat([text, audio], dim=0), torch.cat([tmask, amask], dim=0)

    @torch.inference_mode()
    def generate_asr(self, task_prompt, task_name=None, text_token=None, semantic_token=None, reason_token=None,
                     temperature: float = 0.9, topk: int = 200, cfg_scale=1.0, max_audio_frames: int = 500,
                     return_ids: Optional[bool] = None, _packed=None):
        """Returns the decoded text (reference behaviour) when a text tokenizer was supplied, else the list of text ids."""
        model, dev = self._model, self.device
        model.reset_caches()
        # _packed: (tokens, mask) from another task's prompt packer (audio understanding, speech-to-text)
        tokens, tokens_mask = _packed if _packed is not None else self.prepare_asr_task(task_prompt, reason_token, semantic_token)
        S = tokens.size(0)
        curr_tokens = tokens.unsqueeze(0).to(dev)
        curr_mask = tokens_mask.bool().unsqueeze(0).to(dev)
        pos = torch.arange(0, S, device=dev).unsqueeze(0)
        model.forward_prefix(curr_tokens[:, :-1], labels=curr_tokens[:, 1:, :-1], tokens_mask=curr_mask, loss_mask=curr_mask,
                             input_pos=pos[:, :-1], input_pos_maxp1=S - 1)
        curr_tokens, curr_mask = curr_tokens[:, -1:], curr_mask[:, -1:]
        nq = self.parallel_number - 1
        text_mask = torch.cat([torch.zeros(1, 1, nq, dtype=torch.bool), torch.ones(1, 1, 1, dtype=torch.bool)], dim=-1).to(dev)
        curr_pos, maxp1 = S - 1, S
        ids = []
        for _ in range(max_audio_frames):
            sample = model.generate_frame(curr_tokens, curr_mask, input_pos=curr_pos, input_pos_maxp1=maxp1, temperature=temperature,
                                          topk=topk, forbid_prefix=0)
            t = int(sample[0, 0].item())  # one 4-byte D2H per frame (the reference compares on the device and syncs too)
            if t == EOS_TEXT:
                break
            ids.append(t)
            nxt = torch.zeros(1, 1, nq + 1, dtype=torch.int64, device=dev)
            nxt[0, 0, -1] = t
            curr_tokens, curr_mask = nxt, text_mask
            curr_pos += 1
            maxp1 += 1
        if (return_ids is None and self._text_tokenizer is None) or return_ids:
            return ids
        return self._text_tokenizer.decode(torch.tensor(ids))

    generate_audio_caption = generate_asr  # audio_music_caption_task.py:201-258: same loop, caption prompt
    generate_lyric_asr = generate_asr      # lyric_asr_task.py:202-258: same loop and prompt packing (prepare_lyric_asr_task :175)
