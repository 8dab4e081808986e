"""Task drivers with the generic condition-sequence packer: audio understanding (question answering over audio),
speech-to-text and speech-to-speech.

Mirrors the `Generator` classes of the reference's evaluation/audio_understanding.py (get_condition_seq :227-282,
generate_answer :284-338), evaluation/speech_s2t.py (generate_answer :328-390, which also returns the prompt length and
refuses prompts of >= 1500 frames) and evaluation/speech_s2s.py (generate_audio :283-360).  The prompt is the task prompt
followed by the entries of `d` in the order of `keys`, each packed by its type:
    'text'          <key> tokens </key> special tokens (none for 'text_seq*'), text stream
    'audio_prompt'  semantic codes + reason_card with semantic bos/eos, whose first and last frame are then replaced by
                    audio_prompt_bos / audio_prompt_eos
    anything else   reason codes with reason bos/eos frames ('reason_seq*') or semantic codes + reason_card with semantic bos/eos
The AR loops are the two of asr_task.py / tts_task.py (text decode until <|end_of_text|>; audio frames with the reason ->
semantic phase switch)."""
import torch

from . import asr_task, tts_task
from .tts_task import SPECIAL_TOKENS


class _ConditionPacker:
    """get_condition_seq of the three reference files; `codes_time_major`: audio entries arrive as (8, T) and are transposed
    (audio_understanding.py:245, speech_s2s.py) - speech_s2t.py:277-279 transposes only when the tensor looks like (8, T)."""

    audio_prompt_bos = None
    audio_prompt_eos = None

    def _init_packer(self, train_args):
        self.special_token_dict = dict(SPECIAL_TOKENS)
        self.semantic_bos, self.semantic_eos = train_args.semantic_bos, train_args.semantic_eos
        self.reason_bos, self.reason_eos = train_args.reason_bos, train_args.reason_eos
        self.audio_reason_card = train_args.audio_reason_card
        self.audio_prompt_bos = getattr(train_args, "audio_prompt_bos", None)
        self.audio_prompt_eos = getattr(train_args, "audio_prompt_eos", None)

    def add_special_token(self, key, data):
        if key.startswith('text_seq'):
            return data
        key = key.replace('_seq', '')
        bos = torch.ones(1) * self.special_token_dict['<' + key + '>']
        eos = torch.ones(1) * self.special_token_dict['</' + key + '>']
        return torch.cat([bos, data, eos], dim=0)

    def reason_seq_bos_eos(self, x):
        bos = torch.ones(1, x.shape[1]) * self.reason_bos
        eos = torch.ones(1, x.shape[1]) * self.reason_eos
        return torch.cat([bos, x, eos], dim=0)

    def semantic_seq_bos_eos(self, x):
        bos = torch.ones(1, x.shape[1]) * self.semantic_bos
        eos = torch.ones(1, x.shape[1]) * self.semantic_eos
        return torch.cat([bos, x, eos], dim=0) + self.audio_reason_card

    def audio_prompt_seq_bos_eos(self, x):
        bos = torch.ones(1, x.shape[1]) * self.audio_prompt_bos
        eos = torch.ones(1, x.shape[1]) * self.audio_prompt_eos
        return torch.cat([bos, x[1:-1, :], eos], dim=0)

    def get_condition_seq(self, d, keys, types, task_prompt_data, smart_transpose=False):
        data = self.text_pad(task_prompt_data)
        m = torch.zeros((data.shape[0], self.parallel_number))
        m[:, -1] = True
        sequence, mask = [data], [m]
        for key, tp in zip(keys, types):
            if tp == 'text':
                x = self.text_pad(self.add_special_token(key, d[key]))
                m = torch.zeros((x.shape[0], self.parallel_number))
                m[:, -1] = True
            else:
                x = d[key].long()
                if not smart_transpose or (x.dim() == 2 and x.shape[0] == 8 and x.shape[1] != 8):
                    x = x.transpose(0, 1)
                if tp == 'audio_prompt':
                    x = self.audio_prompt_seq_bos_eos(self.semantic_seq_bos_eos(x))
                elif key.startswith('reason_seq'):
                    x = self.reason_seq_bos_eos(x)
                else:
                    x = self.semantic_seq_bos_eos(x)
                x = self.audio_pad(x)
                m = torch.zeros((x.shape[0], self.parallel_number))
                m[:, :-1] = True
            sequence.append(x)
            mask.append(m)
        return torch.cat(sequence, dim=0).to(torch.int64), torch.cat(mask, dim=0)


class Generator(asr_task.Generator, _ConditionPacker):
    """audio_understanding.py / speech_s2t.py: condition sequence -> text."""

    def __init__(self, model, train_args, *a, **kw):
        super().__init__(model, train_args, *a, **kw)
        self._init_packer(train_args)

    def generate_answer(self, task_prompt, task_name=None, d=None, keys=None, types=None, temperature: float = 0.9, topk: int = 200,
                        cfg_scale=1.0, speech_s2t: bool = False, **kw):
        """audio_understanding.py:284-338.  speech_s2t=True: the speech_s2t.py variant - (8, T)-aware packing, `(-1, -1)` for
        prompts of >= 1500 frames, and the prompt length returned next to the text (:328-390)."""
        packed = self.get_condition_seq(d, keys, types, task_prompt, smart_transpose=speech_s2t)
        if speech_s2t and packed[0].shape[0] >= 1500:
            return -1, -1
        out = self.generate_asr(task_prompt, task_name, temperature=temperature, topk=topk, cfg_scale=cfg_scale, _packed=packed, **kw)
        return (out, 1) if speech_s2t else out  # speech_s2t.py:390 also returns curr_tokens.shape[1]: 1, the last fed-back frame


class AudioGenerator(tts_task.Generator, _ConditionPacker):
    """speech_s2s.py and speech_edit_ss.py (identical Generator classes): condition sequence -> (reason tokens, semantic tokens)."""

    def __init__(self, model, train_args, *a, **kw):
        super().__init__(model, train_args, *a, **kw)
        self._init_packer(train_args)

    def audio_pad(self, x):
        seq = torch.ones((x.shape[0], self.parallel_number)).to(torch.int64) * self.empty_token
        seq[:, :-1] = x
        return seq

    def generate_audio(self, task_prompt, task_name=None, d=None, keys=None, types=None, temperature: float = 0.9, topk: int = 200,
                       cfg_scale=1.0, **kw):
        packed = self.get_condition_seq(d, keys, types, task_prompt) + (None, None)
        return self.generate_tts(task_prompt, task_name, temperature=temperature, topk=topk, cfg_scale=cfg_scale, _packed=packed, **kw)
