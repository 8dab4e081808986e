"""Task driver for text -> audio-token generation (TTS; TTM/TTA use the same loop with <caption> tags).

Mirrors the `Generator` of the reference's evaluation/tts_task.py (:53-285) and evaluation/musicgen_task.py
(identical loop, `<caption>` instead of `<transcription>`): prompt packing into (S, 9) token/mask frames,
one forward_prefix call, then the per-frame AR loop with the reason -> semantic phase switch and EOS stop.

Differences that do not change results:
  * the per-frame EOS / phase checks read ONE (1, 9) int32 row back to the host instead of issuing several
    torch.all(...) device syncs (tts_task.py:261-263);
  * `max_audio_frames` (500 in the reference, :222) and a synthetic `fixed_schedule` are arguments, so random-weight
    benchmarks (which never emit EOS) can follow the TTS-10 s schedule of SURVEY.md section 8d.
"""
from types import SimpleNamespace
from typing import Optional, Tuple

import torch

SPECIAL_TOKENS = {  # tts_task.py:89-93
    '<think>': 128002, '</think>': 128003, '</answer>': 128005, '<transcription>': 128011, '</transcription>': 128012,
    '<lyric>': 128013, '</lyric>': 128014, '<caption>': 128015, '</caption>': 128016, '<answer>': 128017,
    '<reason_token>': 128018, '<semantic_token>': 128019,
}


def default_train_args(audio_reason_card=4100, audio_semantic_card=8200, parallel_number=9):
    """The llm_config.yaml keys the Generator reads (tts_task.py:76-86; defaults of llm_utils/arguments.py:27-52).
    The yaml ships only with the HF checkpoint, so the cards are ASSUMED."""
    return SimpleNamespace(text_pad_token=128002, semantic_pad_token=8192, semantic_eos=8193, semantic_bos=8194,
                           reason_eos=4097, reason_bos=4098, reason_pad_token=4096, parallel_number=parallel_number,
                           audio_reason_card=audio_reason_card, audio_semantic_card=audio_semantic_card)


class Generator:
    def __init__(self, model, train_args, audio_tokenizer_config=None, audio_model_path=None, text_tokenizer_path=None,
                 is_cfg: bool = False, text_tokenizer=None, tag: str = "transcription"):
        self._model = model
        self._model.setup_caches(2 if is_cfg else 1)  # tts_task.py:64-67
        self.is_cfg = is_cfg
        self._text_tokenizer = text_tokenizer
        self.device = next(model.parameters()).device
        self.sample_rate = 24000
        self.empty_token = 0
        self.text_pad_token = train_args.text_pad_token
        self.semantic_pad_token = train_args.semantic_pad_token
        self.semantic_eos = train_args.semantic_eos
        self.semantic_bos = train_args.semantic_bos
        self.reason_eos = train_args.reason_eos
        self.reason_bos = train_args.reason_bos
        self.reason_pad_token = train_args.reason_pad_token
        self.parallel_number = train_args.parallel_number
        self.audio_reason_card = train_args.audio_reason_card
        self.special_token_dict = dict(SPECIAL_TOKENS)
        self.tag = tag  # 'transcription' (TTS) or 'caption' (TTM / TTA, musicgen_task.py)

    # ---- prompt packing (tts_task.py:143-206)
    def text_pad(self, x):
        seq = torch.ones((len(x), self.parallel_number)).to(torch.int64)
        seq[:, -1] = x
        seq[:, :-1] = seq[:, :-1] * self.empty_token
        return seq

    def add_special_token(self, key, data):
        if key == 'text_seq':
            return data
        key = key.replace('_seq', '')
        bos = torch.ones(1) * self.special_token_dict['<' + key + '>']
        eos = torch.ones(1) * self.special_token_dict['</' + key + '>']
        return torch.cat([bos, data, eos], dim=0)

    def prepare_tts_task(self, task_prompt, text_seq) -> Tuple[torch.Tensor, torch.Tensor]:
        prompt = self.text_pad(task_prompt)
        pmask = torch.zeros((prompt.shape[0], self.parallel_number))
        pmask[:, -1] = True
        text_seq = self.add_special_token(self.tag + '_seq', text_seq)
        text = self.text_pad(text_seq)
        tmask = torch.zeros((text.shape[0], self.parallel_number))
        tmask[:, -1] = True
        return torch.cat([prompt, text], dim=0), torch.cat([pmask, tmask], dim=0)

    def prepare_tts_task_for_cfg(self, task_prompt, text_seq):
        task_prompt = torch.ones_like(task_prompt) * self.text_pad_token
        text_seq = self.add_special_token(self.tag + '_seq', text_seq)
        text_seq = torch.ones_like(text_seq) * self.text_pad_token
        data = torch.cat([self.text_pad(task_prompt), self.text_pad(text_seq)], dim=0)
        mask = torch.zeros((data.shape[0], self.parallel_number))
        mask[:, -1] = True
        return data, mask

    # ---- AR loop (tts_task.py:208-285)
    @torch.inference_mode()
    def generate_tts(self, task_prompt, task_name=None, text_token=None, semantic_token=None, reason_token=None,
                     temperature: float = 0.9, topk: int = 200, cfg_scale=1.0, max_audio_frames: int = 500,
                     fixed_schedule: Optional[Tuple[int, int]] = None, pinned_staging: bool = True, _packed=None,
                     device_loop: bool = False, sync_every: int = 16):
        """Returns (reason tokens (8, T_r), semantic tokens (8, T_s)) int64 on the model device.
        fixed_schedule=(n_reason, n_semantic): synthetic mode - switch phase after n_reason frames and stop after
        n_reason + n_semantic frames regardless of EOS (random weights never emit EOS).
        device_loop=True (B = 1): the sample feedback and the phase / EOS tests of :253-279 run on the device (Model_stage3.tts_frames);
        the host reads a 16-byte state record once per `sync_every` frames instead of the sampled row after every frame - at most
        sync_every - 1 frames are computed past the end frame and discarded.  Same tokens as the host loop."""
        model, dev = self._model, self.device
        model.reset_caches()
        # _packed: (tokens, mask, cfg_tokens, cfg_mask) from another task's prompt packer (instruct TTS, speech-to-speech)
        tokens, tokens_mask = _packed[:2] if _packed is not None else self.prepare_tts_task(task_prompt, text_token)
        S = tokens.size(0)
        if self.is_cfg:
            ctok, cmask = _packed[2:] if _packed is not None else self.prepare_tts_task_for_cfg(task_prompt, text_token)
            tokens = torch.stack([tokens, ctok])
            tokens_mask = torch.stack([tokens_mask, cmask])
            bs = 2
        else:
            tokens, tokens_mask = tokens.unsqueeze(0), tokens_mask.unsqueeze(0)
            bs = 1
        stage_t = tokens.pin_memory() if pinned_staging else tokens
        stage_m = tokens_mask.bool()
        stage_m = stage_m.pin_memory() if pinned_staging else stage_m
        curr_tokens = stage_t.to(dev, non_blocking=True)
        curr_mask = stage_m.to(dev, non_blocking=True)
        pos = torch.arange(0, S, device=dev).unsqueeze(0).repeat(bs, 1)
        self.h2d_bytes = stage_t.numel() * 8 + stage_m.numel()
        self.d2h_bytes = 0
        model.forward_prefix(curr_tokens[:, :-1], labels=curr_tokens[:, 1:, :-1], tokens_mask=curr_mask, loss_mask=curr_mask,
                             input_pos=pos[:, :-1], input_pos_maxp1=S - 1)
        curr_tokens, curr_mask = curr_tokens[:, -1:], curr_mask[:, -1:]
        curr_pos, maxp1 = S - 1, S
        is_reason, save_flag, forbid_prefix = True, True, 0
        pre_reason, pre_semantic = [], []
        nq = self.parallel_number - 1
        audio_mask = torch.cat([torch.ones(bs, 1, nq, dtype=torch.bool), torch.zeros(bs, 1, 1, dtype=torch.bool)], dim=-1)
        audio_mask = audio_mask.to(dev)
        end_tok = self.semantic_eos + self.audio_reason_card
        n_frames = 0
        if device_loop and not self.is_cfg:
            state = torch.zeros(4, dtype=torch.int32, device=dev)
            frames_buf = torch.zeros(max_audio_frames, nq + 1, dtype=torch.int32, device=dev)
            total = max_audio_frames if fixed_schedule is None else min(max_audio_frames, fixed_schedule[0] + fixed_schedule[1])
            total = min(total, getattr(model, "_max_seq_length", 1 << 30) - curr_pos)  # the KV caches end there
            launched, st = 0, None
            while launched < total:
                n = min(int(sync_every), total - launched)
                model.tts_frames(curr_tokens if launched == 0 else None, curr_mask if launched == 0 else None, curr_pos + launched, n, state,
                                 frames_buf, temperature, topk, self.reason_eos, end_tok, self.audio_reason_card,
                                 fixed_schedule[0] if fixed_schedule is not None else -1)
                launched += n
                st = state.cpu()  # ONE D2H of 16 bytes per chunk of frames
                self.d2h_bytes += 16
                if int(st[1]):
                    break
            n_rec, switch = int(st[2]), int(st[3])
            rows = frames_buf[:n_rec].cpu().to(torch.int64)
            self.d2h_bytes += rows.numel() * 4
            self.n_frames = n_rec + (1 if int(st[1]) else 0)
            # rows 1 .. switch - 1 are reason frames, row `switch` is the reason_eos frame (not kept, :263-271), the rest semantic
            reason_rows = rows[: (switch - 1 if switch else n_rec), 1:]
            sem_rows = rows[switch:, 1:] - self.audio_reason_card if switch else rows[:0, 1:]
            de_reason = reason_rows[1:].t().contiguous() if reason_rows.shape[0] > 1 else torch.zeros(nq, 0, dtype=torch.int64)
            de_sem = sem_rows[1:].t().contiguous() if sem_rows.shape[0] > 1 else torch.zeros(nq, 0, dtype=torch.int64)
            return de_reason.to(dev), de_sem.to(dev)
        for _ in range(max_audio_frames):
            sample = model.generate_frame(curr_tokens, curr_mask, input_pos=curr_pos, input_pos_maxp1=maxp1,
                                          temperature=temperature, topk=topk, forbid_prefix=forbid_prefix)
            row = sample[0].cpu()  # ONE D2H of 36 bytes per frame: EOS / phase logic runs on the host
            self.d2h_bytes += row.numel() * 4
            n_frames += 1
            audio = row[1:]
            if fixed_schedule is None:
                if bool((audio == end_tok).all()):
                    break
                switch = bool((audio == self.reason_eos).all())
            else:  # frames 1..n_reason run with forbid_prefix=0 (the last one plays the reason_eos frame)
                switch = n_frames == fixed_schedule[0]
            if switch:
                is_reason, save_flag, forbid_prefix = False, False, self.audio_reason_card
            if save_flag:
                (pre_reason if is_reason else pre_semantic).append(audio if is_reason else audio - self.audio_reason_card)
            else:
                save_flag = True
            # feed back: audio tokens in cols 0..7, text token in col 8 (tts_task.py:276-279); the sample is already on
            # the device, so no H2D is needed for the next frame
            s0 = sample[0:1].long()
            curr_tokens = torch.cat([s0[:, 1:], s0[:, 0:1]], dim=-1).unsqueeze(1)
            if self.is_cfg:
                curr_tokens = curr_tokens.repeat(2, 1, 1)
            curr_mask = audio_mask
            curr_pos += 1
            maxp1 += 1
            if fixed_schedule is not None and n_frames >= fixed_schedule[0] + fixed_schedule[1]:
                break
        self.n_frames = n_frames
        de_reason = torch.stack(pre_reason[1:]).transpose(0, 1).to(torch.int64) if len(pre_reason) > 1 else torch.zeros(nq, 0, dtype=torch.int64)
        de_sem = torch.stack(pre_semantic[1:]).transpose(0, 1).to(torch.int64) if len(pre_semantic) > 1 else torch.zeros(nq, 0, dtype=torch.int64)
        return de_reason.to(dev), de_sem.to(dev)

    # ---- instruct TTS (insturct_tts_task.py:170-298): [task prompt | <caption> style caption | <transcription> text], same loop
    def _text_block(self, key, seq, blank=False):
        seq = self.add_special_token(key, seq)
        if blank:
            seq = torch.ones_like(seq) * self.text_pad_token
        data = self.text_pad(seq)
        mask = torch.zeros((data.shape[0], self.parallel_number))
        mask[:, -1] = True
        return data, mask

    def prepare_instruct_tts_task(self, task_prompt, caption_seq, text_seq, blank=False):
        if blank:  # prepare_instruct_tts_task_for_cfg: every text position replaced by text_pad_token
            task_prompt = torch.ones_like(task_prompt) * self.text_pad_token
        blocks = [self._text_block('text_seq', task_prompt), self._text_block('caption_seq', caption_seq, blank),
                  self._text_block('transcription_seq', text_seq, blank)]
        return torch.cat([b[0] for b in blocks], dim=0), torch.cat([b[1] for b in blocks], dim=0)

    def generate_instruct_tts(self, task_prompt, task_name=None, text_token=None, caption=None, semantic_token=None, reason_token=None,
                              temperature: float = 0.9, topk: int = 200, cfg_scale=1.0, **kw):
        packed = self.prepare_instruct_tts_task(task_prompt, caption, text_token)
        if self.is_cfg:
            packed = packed + self.prepare_instruct_tts_task(task_prompt, caption, text_token, blank=True)
        else:
            packed = packed + (None, None)
        return self.generate_tts(task_prompt, task_name, text_token=text_token, temperature=temperature, topk=topk, cfg_scale=cfg_scale,
                                 _packed=packed, **kw)

    generate_audio = generate_tts  # musicgen_task.py:210 / audiogen_task.py:209 (TTM / TTA): same loop, construct with tag='caption'
    generate_LTS = generate_tts    # songen_task.py:210 (lyrics -> song): same loop, construct with tag='lyric'
