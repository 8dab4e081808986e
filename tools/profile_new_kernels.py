"""Profiling driver (run under ncu): a few launches of the kernels added late in round 1 - the streaming transformer's ring
attention, sample_token and the flow decoder's attention.  Not a benchmark.

Round-1 note: the one attempt to capture this under `ncu --set full` hit its 200 s limit before the first matching kernel (the
first `import torch` on a fresh box takes about a minute by itself, more under ncu's injection) and used up the round's GPU
budget - page torch in first (`python -c "import torch"` in the same gpurun command) and give the call a longer limit:

    python -c "import torch"; ncu --set full --clock-control none --import-source on -k regex:"ring_attn_kernel|sample_token_kernel|dit_attn_kernel" -s 6 -c 6 \
        -o gpurun_out/new_kernels python tools/profile_new_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.llm_modules.transformer import StreamingTransformer  # noqa: E402
from uniaudio2_b200.llm_utils.sampling import sample_token  # noqa: E402
from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    # one layer of the temporal-transformer shape with a full 3000-slot ring, batch 4: 12 splits x 32 heads x 4 rows
    m = StreamingTransformer(d_model=4096, num_heads=32, num_layers=1, dim_feedforward=16384, causal=True, context=3000,
                             positional_embedding="rope", norm="rms_norm_f32", gating="silu", device=dev)
    m.set_option("graph", 0)
    x = torch.randn(4, 500, 4096, device=dev)
    with m.streaming(4):
        for _ in range(6):
            m(x)          # fills the 3000-slot ring 500 rows at a time (many-row path)
        for _ in range(2):
            m(x[:, :1])   # decode steps against the full ring: the launches to look at
    lg = torch.randn(8, 2048, device=dev) * 2
    sample_token(lg, use_sampling=True, temp=0.8, top_k=250)
    lg2 = torch.randn(1, 128256, device=dev) * 2
    sample_token(lg2, use_sampling=True, temp=0.8, top_k=50)
    # two layers of the production DiT on the CFG batch of a 20 s window
    d = Transformer1DModel(num_attention_heads=24, attention_head_dim=64, in_channels=1040, out_channels=136, num_layers=2,
                           attention_bias=True, activation_fn="gelu-approximate", norm_type="ada_norm_single",
                           norm_elementwise_affine=False, norm_eps=1e-6, num_embeds_ada_norm=1000, device=dev)
    xx = torch.randn(2, 500, 1040, device=dev)
    t = torch.full((2,), 0.35, device=dev)
    d(xx, timestep=t)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
