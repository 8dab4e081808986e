"""Multi-GPU plumbing: independent prompts / clips are sharded one replica per GPU, outputs are gathered once.

The reference has no collective on the inference path: multi-GPU inference is N independent processes selected with
`--rank` over a kaldi-style split of the file list (multi_task_inference.py:597,164-169; evaluation/asr_task.py:731-736).
The B200 equivalent keeps replica-per-GPU (the 4.86 B-parameter fp32 model is 19.5 GB, it fits many times over), shards item
i -> rank i mod W, and adds ONE collective: an all_gather of the (padded) generated tokens + lengths over NCCL/NVLink
(SURVEY.md section 8e).  Works with the gloo backend on CPU tensors too (used by the CPU tests).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Items handled by `rank`: i with i mod world == rank (mirrors --rank / JOB sharding)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return list(range(rank, n_items, world))


def gather_variable(local: Sequence[torch.Tensor], n_items: int, pad_value: int = 0, group=None) -> List[torch.Tensor]:
    """All-gather per-item results of different lengths.

    local: this rank's results for shard_indices(n_items, rank, world), each (C, T_i) with a common C and dtype.
    Returns the n_items tensors in original item order on every rank.  Exactly two collectives are issued
    (lengths, then payload padded to the global max length) regardless of the number of items."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_indices(n_items, rank, world)
    if len(local) != len(mine):
        raise ValueError(f"rank {rank} holds {len(local)} results for {len(mine)} items")
    per_rank = (n_items + world - 1) // world
    if world == 1:
        return list(local)
    ref = local[0] if len(local) else None
    # every rank needs C / dtype / device even when it holds nothing: exchange a small header
    dev = ref.device if ref is not None else torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    lens = torch.zeros(per_rank + 1, dtype=torch.int64, device=dev)
    for j, t in enumerate(local):
        lens[j] = t.shape[-1]
    lens[per_rank] = ref.shape[0] if ref is not None else 0
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    all_lens_t = torch.stack(all_lens)
    C = int(all_lens_t[:, per_rank].max())
    max_len = int(all_lens_t[:, :per_rank].max())
    dtype = ref.dtype if ref is not None else torch.int64
    buf = torch.full((per_rank, C, max(max_len, 1)), pad_value, dtype=dtype, device=dev)
    for j, t in enumerate(local):
        buf[j, :, : t.shape[-1]] = t
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    out: List[torch.Tensor] = []
    for i in range(n_items):
        r, j = i % world, i // world
        out.append(bufs[r][j, :, : int(all_lens_t[r, j])].clone())
    return out
