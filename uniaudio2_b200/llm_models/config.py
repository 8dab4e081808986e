"""Config registry for the GPT stacks on the UniAudio2 hot path.

Mirrors the interface of the reference's llm_models/config.py (`Config.from_name`, `name_to_config`) for
the entries the path uses (config.py:785-899): Llama-3.2-{1B,3B,300M,4Layer,Understanding,Generation}.
The rest of litgpt's model zoo is out of scope.
"""
from dataclasses import dataclass, field
from typing import Any, Optional

_LLAMA32_ROPE = dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_seq_len=8192)


@dataclass
class Config:
    name: str = ""
    block_size: int = 131072
    n_layer: int = 16
    n_embd: int = 2048
    vocab_size: int = 128000
    padded_vocab_size: int = 128256
    n_head: int = 32
    n_query_groups: Optional[int] = None
    head_size: Optional[int] = None
    intermediate_size: int = 8192
    norm_eps: float = 1e-5  # config.py:38 (the Llama-3.2 entries do not override it)
    rope_base: int = 500000
    rope_adjustments: Optional[dict] = field(default_factory=lambda: dict(_LLAMA32_ROPE))
    # fixed by every entry on the path; kept so callers can introspect like with litgpt's Config
    rotary_percentage: float = 1.0
    parallel_residual: bool = False
    bias: bool = False
    norm_class_name: str = "RMSNorm"
    mlp_class_name: str = "LLaMAMLP"

    def __post_init__(self):
        if self.head_size is None:
            assert self.n_embd % self.n_head == 0
            self.head_size = self.n_embd // self.n_head
        if self.n_query_groups is None:
            self.n_query_groups = self.n_head
        assert self.n_head % self.n_query_groups == 0
        self.rope_n_elem = int(self.rotary_percentage * self.head_size)
        if (self.rotary_percentage != 1.0 or self.parallel_residual or self.bias or self.norm_class_name != "RMSNorm"
                or self.mlp_class_name != "LLaMAMLP"):
            raise ValueError(f"config {self.name!r}: only bias-free RMSNorm/LLaMAMLP/full-RoPE stacks are on this path")

    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "Config":
        key = name.split("/")[-1]  # 'meta-llama/Llama-3.2-3B' -> 'Llama-3.2-3B' (config.py:137-150 org/name lookup)
        if key not in name_to_config and name not in name_to_config:
            raise ValueError(f"{name!r} is not a supported config name")
        d = dict(name_to_config.get(name, name_to_config.get(key)))
        d.update(kwargs)
        return cls(**d)


def _llama32(name, n_layer, n_embd, n_head):
    return dict(name=name, n_layer=n_layer, n_embd=n_embd, n_head=n_head, n_query_groups=8, intermediate_size=8192)


configs = [
    _llama32("Llama-3.2-1B", 16, 2048, 32),
    _llama32("Llama-3.2-3B", 28, 3072, 24),
    _llama32("Llama-3.2-300M", 4, 2048, 32),
    _llama32("Llama-3.2-4Layer", 4, 2048, 32),
    _llama32("Llama-3.2-Understanding", 3, 3072, 24),
    _llama32("Llama-3.2-Generation", 2, 3072, 24),
]
name_to_config = {c["name"]: c for c in configs}
