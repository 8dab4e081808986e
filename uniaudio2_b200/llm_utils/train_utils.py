"""Checkpoint / token-file plumbing of the inference path, mirroring the reference's formats byte for byte:

  resume_for_inference(resume, exp_dir, model, device)   llm_utils/train_utils.py:159-177
      `torch.load(ckpt)['model']`, DDP/FSDP `module.` prefixes stripped, strict `load_state_dict` (same key names as the
      reference's Model_stage3), newest `ep*.checkpoint` of `exp_dir` when `resume` is None.
  load_llm_config(path) / model_args_from_config(cfg)    multi_task_inference.py:153-182
      `llm_config.yaml` -> argparse.Namespace -> ModelArgs (audio_num_codebooks = parallel_number - 1).
  save_token_files / load_token_files                    multi_task_inference.py:143-144, :523-524
      `{name}_reason.pt`, `{name}_semantic.pt`: `torch.save` of (8, T) int64 tensors.
Host-side only (no kernels): the drop-in takes the public checkpoints and writes the files the reference's later stages read.
"""
import argparse
import logging
import os
from pathlib import Path
from typing import Optional, Tuple

import torch


def _newest_checkpoint(exp_dir: str) -> str:
    """Path of the most recently created `ep*.checkpoint` under exp_dir (creation-time order, like the reference)."""
    found = sorted(Path(exp_dir).glob("ep*.checkpoint"), key=lambda p: os.stat(str(p)).st_ctime)
    if not found:
        raise ValueError("Model for resume is not provided and cannot be detected.")
    return str(found[-1])


def _strip_wrapper_prefix(key: str) -> str:
    # DDP / FSDP save parameters as 'module.<name>'; the reference keeps what follows the LAST 'module.' of such keys
    return key.split("module.")[-1] if key.startswith("module.") else key


def resume_for_inference(resume: Optional[str], exp_dir: Optional[str], model, device="cpu") -> str:
    """Restore `model` from an explicit checkpoint path, or from the newest one in `exp_dir`; returns the path used.
    The checkpoint is the dict written by the reference's trainer: weights under the key 'model'."""
    path = resume if resume is not None else _newest_checkpoint(exp_dir)
    logging.info("restoring weights from %s", path)
    weights = torch.load(path, map_location="cpu")["model"]
    model.load_state_dict({_strip_wrapper_prefix(k): v for k, v in weights.items()})  # strict: same keys as the reference
    return path


def load_llm_config(path: str) -> argparse.Namespace:
    import yaml

    with open(path, "r", encoding="utf-8") as f:
        return argparse.Namespace(**yaml.safe_load(f))


def model_args_from_config(train_args):
    from ..llm_models.model_new import ModelArgs

    return ModelArgs(
        decoder_name=train_args.local_model,
        llm_pretrained_model=train_args.llm_pretrained_model,
        llm_name=train_args.llm_name,
        audio_semantic_vocab_size=train_args.audio_semantic_card,
        audio_reason_vocab_size=train_args.audio_reason_card,
        audio_num_codebooks=train_args.parallel_number - 1,
        audio_embeddings_path=train_args.audio_embeddings_path,
        audio_understanding_expert_path=train_args.audio_understanding_expert_path,
    )


def save_token_files(out_dir: str, name: str, reason: torch.Tensor, semantic: torch.Tensor) -> Tuple[str, str]:
    """(8, T_r) / (8, T_s) integer codes -> `{name}_reason.pt`, `{name}_semantic.pt` (int64, CPU) like the tokenizer stage."""
    for t in (reason, semantic):
        if t.dim() != 2:
            raise ValueError("token tensors must be (num_codebooks, T)")
    os.makedirs(out_dir, exist_ok=True)
    pr, ps = os.path.join(out_dir, f"{name}_reason.pt"), os.path.join(out_dir, f"{name}_semantic.pt")
    torch.save(reason.detach().to("cpu", torch.long), pr)
    torch.save(semantic.detach().to("cpu", torch.long), ps)
    return pr, ps


def load_token_files(out_dir: str, name: str) -> Tuple[torch.Tensor, torch.Tensor]:
    reason = torch.load(os.path.join(out_dir, f"{name}_reason.pt"), map_location="cpu")
    semantic = torch.load(os.path.join(out_dir, f"{name}_semantic.pt"), map_location="cpu")
    return reason, semantic
