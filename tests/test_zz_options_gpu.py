"""GPU tests of the kernels behind options (first run on a B200 in round 2: gpurun_out/r2_first, profiles/r2_first_call.md).

(1) GPU parity of the codes -> waveform caller: uniaudio2_b200's AudioDiffusion1D.inference_codes (code lookups, projections,
flow-matching solve through the C ABI) under the product's ReasoningTokenizer.token2audio_no_reason, against the fixtures that
oracle/make_golden_detok.py produced from the UNMODIFIED reference source (tests/golden/detok_golden.pt).  The SQ-codec decoder
is the fixture's stand-in (a transposed convolution with the real hop of 960, evaluated with torch in this TEST only - the
product's ScalarModel has its own parity suite, tests/test_scalar_gpu.py); every noise draw is replayed from the fixture.
Bar: waveform / latents within 1e-4 max-abs relative to the tensor's scale.

(2) The "bf16" option of the flow-matching decoder (ua2_dit_set_option): many-row linears on bf16 operands with fp32 accumulation
(CUTLASS tcgen05 kind::f16 collective, csrc/ua2_tcgemm_bf16.cu) against the fp32 oracle, bar 3e-2 relative to the output scale
(bf16 has 8 mantissa bits; the reference runs these linears under torch.autocast(bfloat16) itself).

(3) The global option "conv_tc": causal convolutions with Cin * K >= 1024 as im2col + tcgen05 3xTF32 GEMM (csrc/ua2_convtc.cu)
against the conv oracle (2e-5 relative, the bar of the SIMT core's own test) and, on the full codec geometry, bit-equal VQ indices.

(4) The fused SEANet residual block (ua2_resblock_f32, global option "resblock_fused") against the two-convolution oracle and on
the full codec geometry."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import detok_oracle as TO
from oracle import dit_oracle as DO
from oracle.make_golden_detok import CB_DIM, CB_SIZE, CODEC_DIM, DIT, VQS, random_params

pytestmark = [pytest.mark.gpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


@pytest.fixture(scope="module")
def detok_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "detok_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def parts():
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import AudioDiffusion1D
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    p = random_params(31)
    dit_sd = DO.random_state_dict(DIT, seed=32)
    m = AudioDiffusion1D(Transformer1DModel(**DIT.ctor_kwargs()), codec_dim=CODEC_DIM, codebook_size=CB_SIZE, codebook_dim=CB_DIM)
    sd = dict(m.state_dict())
    for name, nq in VQS:
        for i in range(nq):
            sd[f"{name}.layers.{i}._codebook.embed"] = p[f"{name}.codebooks"][i:i + 1].clone()
        sd[f"{name}.project_out.weight"] = p[f"{name}.project_out.weight"]
        sd[f"{name}.project_out.bias"] = p[f"{name}.project_out.bias"]
    sd["cond_feature_emb.weight"], sd["cond_feature_emb.bias"] = p["cond_feature_emb.weight"], p["cond_feature_emb.bias"]
    sd["zero_cond_embedding1"] = p["zero_cond_embedding1"]
    for k, v in dit_sd.items():
        sd["cfm_wrapper.estimator." + k] = v
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda:0")
    orc = TO.DetokOracle(p, DO.DitOracle(DIT, dit_sd), None)
    return p, m, orc


def test_inference_codes_matches_oracle(parts):
    """One window: 25 codes -> 50 latent frames, 7 in-context frames, 3 Euler steps."""
    p, m, orc = parts
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, CB_SIZE, (1, 8, 25), generator=g)
    true_latents = torch.randn(1, 50, 136, generator=g)
    noise = torch.randn(1, 50, 136, generator=g)
    with torch.no_grad():
        ref = orc.inference_codes(codes, true_latents, 50, 7, 1.5, 3, lambda shape: noise.clone())
    m.prepare_latents = lambda bs, n, dtype, device: noise.to(device)
    try:
        out = m.inference_codes([codes.cuda()], None, true_latents.cuda(), 50, 7, additional_feats=[], guidance_scale=1.5, num_steps=3,
                                scenario="other_seg")
    finally:
        del m.prepare_latents
    assert out.shape == ref.shape
    assert torch.equal(out[:, :7].cpu(), true_latents[:, :7])  # the in-context rows are the given latents
    assert _rel(out.cpu(), ref) < TOL


def test_token2audio_matches_reference_golden(detok_golden, parts):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.reason_tokenizer import ReasoningTokenizer

    p, m, _ = parts
    w = p["sq_decode.weight"].cuda()

    class SQStandIn:
        @staticmethod
        def decode(latent):
            return F.conv_transpose1d(latent.contiguous(), w, stride=960)

    tok = ReasoningTokenizer(m, SQStandIn(), device=torch.device("cuda:0"))
    assert tok.autocast_bf16  # default: the reference's torch.autocast(bfloat16) around the window loop (reason_tokenizer.py:265)
    try:
        # the fixtures come from the reference on the CPU, where its cuda autocast context is inert: fp32.  The fp32-class mode meets
        # 1e-4; the default bf16 mode is held to bf16's reach (8 mantissa bits through a 4-step flow solve)
        for autocast, tol in ((False, TOL), (True, 5e-2)):
            tok.autocast_bf16 = autocast
            for c in detok_golden["cases"]:
                it = iter(c["draws"])
                tok._randn = lambda *shape: next(it)
                m.prepare_latents = lambda bs, n, dtype, device: next(it).to(device)
                wav = tok.token2audio_no_reason(c["codes"], False, duration=c["duration"], num_steps=c["steps"], disable_progress=True)
                assert wav.shape == c["wav"].shape and wav.device.type == "cpu"
                assert _rel(wav, c["wav"]) < tol, f"{c['codes'].shape[-1]} codes, autocast {autocast}"
                assert next(it, None) is None  # every recorded draw was consumed, in the reference's order
    finally:
        m.cfm_wrapper.estimator.set_option("bf16", 0)
        if "prepare_latents" in m.__dict__:
            del m.prepare_latents


@pytest.mark.parametrize("heads,hd,T,B", [(3, 64, 150, 2), (2, 128, 70, 1)])
def test_dit_bf16_option_against_fp32_oracle(heads, hd, T, B):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    cfg = DO.DitCfg(num_attention_heads=heads, attention_head_dim=hd, in_channels=56, out_channels=12, num_layers=2,
                    num_positional_embeddings=160)
    sd = DO.random_state_dict(cfg, seed=heads * 10 + hd)
    m = Transformer1DModel(**cfg.ctor_kwargs())
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda:0")
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, cfg.in_channels, generator=g)
    t = torch.rand(B, generator=g)
    with torch.no_grad():
        ref = DO.DitOracle(cfg, sd).forward(x, t)
    y32 = m(x.cuda(), timestep=t.cuda()).sample.cpu()
    m.set_option("bf16", 1)
    y16 = m(x.cuda(), timestep=t.cuda()).sample.cpu()
    m.set_option("bf16", 0)
    y32b = m(x.cuda(), timestep=t.cuda()).sample.cpu()
    assert _rel(y32, ref) < 1e-4 and torch.equal(y32, y32b)  # the option leaves the default path untouched
    err = _rel(y16, ref)
    assert 1e-6 < err < 3e-2, err  # really a different arithmetic, and within bf16's reach


# (3) option "conv_tc": wide causal convolutions as im2col + tcgen05 3xTF32 GEMM (csrc/ua2_convtc.cu)
@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,elu,res,rep", [
    (2, 128, 256, 2501, 10, 5, 1, 0, 0),    # K_total 1280: one chunk, ragged length
    (1, 512, 1024, 640, 16, 8, 1, 0, 0),    # K_total 8192
    (1, 1024, 512, 333, 3, 1, 1, 1, 0),     # last encoder conv + residual
    (2, 512, 512, 400, 4, 2, 0, 0, 1),      # ConvDownsample1d: replicate pad, no bias
    (1, 64, 128, 4000, 8, 4, 1, 0, 0),      # K_total 512 < 1024: stays on the SIMT core (the option must not change it)
])
def test_conv_tc_option_matches_oracle(B, Cin, Cout, T, K, stride, elu, res, rep):
    import math

    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(Cin + Cout + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = None if rep else torch.randn(Cout, generator=g) * 0.1
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, dilation=1, pad_mode="replicate" if rep else "constant")
    r = torch.randn_like(ref) if res else None
    if res:
        ref = r + ref
    xd, wd = x.cuda(), w.contiguous().cuda()
    bd = b.cuda() if b is not None else None
    rd = r.cuda() if res else None
    outs = []
    try:
        for opt in (0, 1):
            _lib.check(L.ua2_set_global_option(b"conv_tc", opt))
            y = torch.full((B, Cout, ref.shape[-1]), float("nan"), device="cuda")
            _lib.check(L.ua2_conv1d_causal_gemm_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(rd), _lib.ptr(y), B, Cin, Cout, T, K,
                                                    stride, 1, elu, rep, None))
            torch.cuda.synchronize()
            outs.append(y.cpu())
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_tc", 0))
    # the SIMT core meets 2e-5; the tensor-core accumulator (TMEM, fp32) rounds toward zero at every k-step of 8, a bias of about
    # steps / 2 ulp of the running sum (measured on the B200: 2.4e-5 relative at Cin * K = 2048, i.e. 768 steps of the 3x longer
    # inner dimension) - bar 1.5 * steps * 2^-24
    tc_tol = max(2e-5, 1.5 * (3 * Cin * K / 8) * 2.0 ** -24)
    for y, tol in zip(outs, (2e-5, tc_tol)):
        assert float((y - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))
    if Cin * K < 1024:
        assert torch.equal(outs[0], outs[1])


def test_conv_tc_option_keeps_codec_indices():
    """Full mimi_config.yaml geometry: VQ indices with the wide convolutions on the tensor cores equal the oracle's."""
    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    cfg = CO.MimiCfg()
    sd = CO.random_mimi_state_dict(cfg, seed=7)
    m = MimiCodec(sample_rate=cfg.sample_rate, n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, compress=cfg.compress,
                  latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim, rvq_layers=cfg.rvq_layers,
                  num_heads=cfg.num_heads, num_layers=cfg.num_layers, layer_scale=cfg.layer_scale, context=cfg.context, device="cuda")
    full = m.state_dict()
    full.update({k: v.cuda() for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    wav = torch.randn(4, 1, 3 * 24000 + 311, generator=torch.Generator().manual_seed(5)) * 0.2
    with torch.no_grad():
        ref_codes = CO.MimiOracle(cfg, sd).encode(wav)
    L = _lib.lib()
    try:
        _lib.check(L.ua2_set_global_option(b"conv_tc", 1))
        codes = m.encode(wav.cuda())
        torch.cuda.synchronize()
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_tc", 0))
    assert torch.equal(codes.cpu(), ref_codes), f"VQ index agreement {float((codes.cpu() == ref_codes).float().mean()):.4f}"


@pytest.mark.parametrize("B,Cin,Cout,T,stride", [(1, 1024, 512, 40, 8), (2, 512, 256, 177, 6), (1, 256, 128, 700, 5), (2, 128, 64, 3001, 4),
                                                 (1, 64, 32, 500, 4)])  # the last one (2 * Cin < 256) stays on the SIMT phase GEMMs
def test_convtr_tc_option_matches_oracle(B, Cin, Cout, T, stride):
    import math

    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(Cin + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g) / math.sqrt(Cin * 2)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.convtr1d_causal(F.elu(x), w, b, stride)
    xd, wt, bd = x.cuda(), w.contiguous().cuda(), b.cuda()
    wp = torch.empty(stride * Cout * Cin * 2, device="cuda")
    _lib.check(L.ua2_convtr1d_repack_phase_f32(_lib.ptr(wt), _lib.ptr(wp), Cin, Cout, stride, None))
    outs = []
    try:
        for opt in (0, 1):
            _lib.check(L.ua2_set_global_option(b"conv_tc", opt))
            y = torch.full((B, Cout, T * stride), float("nan"), device="cuda")
            _lib.check(L.ua2_convtr1d_causal_gemm_f32(_lib.ptr(xd), _lib.ptr(wp), _lib.ptr(bd), _lib.ptr(y), B, Cin, Cout, T, stride, 1, None))
            torch.cuda.synchronize()
            outs.append(y.cpu())
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_tc", 0))
    for y in outs:
        assert y.shape == ref.shape
        assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    if 2 * Cin < 256:
        assert torch.equal(outs[0], outs[1])


def test_conv_tc_option_keeps_decoded_waveform():
    """Full mimi_config.yaml geometry, decode: waveform within 1e-4 max-abs of the oracle with the wide (transposed) convolutions
    on the tensor cores."""
    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    cfg = CO.MimiCfg()
    sd = CO.random_mimi_state_dict(cfg, seed=7)
    m = MimiCodec(sample_rate=cfg.sample_rate, n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, compress=cfg.compress,
                  latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim, rvq_layers=cfg.rvq_layers,
                  num_heads=cfg.num_heads, num_layers=cfg.num_layers, layer_scale=cfg.layer_scale, context=cfg.context, device="cuda")
    full = m.state_dict()
    full.update({k: v.cuda() for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    codes = torch.randint(0, cfg.codebook_size, (4, cfg.rvq_layers, 40), generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref = CO.MimiOracle(cfg, sd).decode(codes)
    L = _lib.lib()
    try:
        _lib.check(L.ua2_set_global_option(b"conv_tc", 1))
        out = m.decode(codes.cuda())
        torch.cuda.synchronize()
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_tc", 0))
    assert float((out.cpu() - ref).abs().max()) < 1e-4


# (4) fused SEANet residual block (csrc/ua2_resblock.cu, global option "resblock_fused")
@pytest.mark.parametrize("B,T", [(1, 128), (2, 1000), (3, 4099), (1, 5)])
def test_fused_resblock_matches_oracle(B, T):
    import math

    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    C, H = 64, 32
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, C, T, generator=g)
    w1 = torch.randn(H, C, 3, generator=g) / math.sqrt(C * 3)
    b1 = torch.randn(H, generator=g) * 0.1
    w2 = torch.randn(C, H, 1, generator=g) / math.sqrt(H)
    b2 = torch.randn(C, generator=g) * 0.1
    hid = CO.conv1d_causal(F.elu(x), w1, b1)
    ref = x + CO.conv1d_causal(F.elu(hid), w2, b2)
    # device operands are named: a temporary `w.cuda()` inside the call would be freed (and its block reused by the next
    # temporary) before the kernel runs - that, not the kernel, failed the first hardware run of this test
    xd, w1d, b1d, w2d, b2d = x.cuda(), w1.cuda(), b1.cuda(), w2.contiguous().cuda(), b2.cuda()
    y = torch.full_like(xd, float("nan"))
    _lib.check(L.ua2_resblock_f32(_lib.ptr(xd), _lib.ptr(w1d), _lib.ptr(b1d), _lib.ptr(w2d), _lib.ptr(b2d), _lib.ptr(y), B, C, H, T, None))
    torch.cuda.synchronize()
    assert float((y.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    with pytest.raises(ValueError):
        _lib.check(L.ua2_resblock_f32(_lib.ptr(xd), _lib.ptr(w1d), _lib.ptr(b1d), _lib.ptr(w2d), _lib.ptr(b2d), _lib.ptr(y), B, 128, 64, T, None))


def test_resblock_fused_option_keeps_codec_results():
    """Full mimi_config.yaml geometry with the 64-channel residual blocks fused: indices equal the oracle's, decoded waveform
    within 1e-4."""
    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    cfg = CO.MimiCfg()
    sd = CO.random_mimi_state_dict(cfg, seed=7)
    m = MimiCodec(sample_rate=cfg.sample_rate, n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, compress=cfg.compress,
                  latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim, rvq_layers=cfg.rvq_layers,
                  num_heads=cfg.num_heads, num_layers=cfg.num_layers, layer_scale=cfg.layer_scale, context=cfg.context, device="cuda")
    full = m.state_dict()
    full.update({k: v.cuda() for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    wav = torch.randn(2, 1, 2 * 24000 + 311, generator=torch.Generator().manual_seed(5)) * 0.2
    orc = CO.MimiOracle(cfg, sd)
    with torch.no_grad():
        ref_codes = orc.encode(wav)
        ref_wav = orc.decode(ref_codes)
    L = _lib.lib()
    try:
        _lib.check(L.ua2_set_global_option(b"resblock_fused", 1))
        codes = m.encode(wav.cuda())
        out = m.decode(ref_codes.cuda())
        torch.cuda.synchronize()
    finally:
        _lib.check(L.ua2_set_global_option(b"resblock_fused", 0))
    assert torch.equal(codes.cpu(), ref_codes)
    assert float((out.cpu() - ref_wav).abs().max()) < 1e-4


# ---- (5) option "attn_ring": persistent K/V chunk ring for long batched contexts (csrc/ua2_attn.cu)
@pytest.mark.parametrize("hs,n_head,G,M,S_max", [(128, 24, 8, 32, 1024), (64, 8, 8, 16, 1024), (32, 4, 2, 40, 800), (128, 24, 8, 5, 1024), (128, 24, 8, 32, 2048)])
def test_attn_ring_option_matches_split_kernel_and_reference(hs, n_head, G, M, S_max):
    """ua2_attn_f32 with the option on: bit-equal to the one-shot split kernel (same item arithmetic) and within 2e-5 of an fp64
    reference; ragged positions (chunk edges, single key, last slot), permuted cache rows.  The launcher takes the ring from 592 work
    items and up to 16 chunks per (row, group): the fourth case (5 x 8 x 16 = 640 items) sits just above the threshold, the last one
    (32 chunks) stays on the one-shot kernel either way.  (The Moshi context window is not an argument of this operator; the windowed form of
    both kernels is compared on the CPU shim, tests/test_kernels_on_cpu_shim.py.)"""
    from uniaudio2_b200 import _lib

    L, P = _lib.lib(), _lib.ptr
    g = torch.Generator().manual_seed(hs + M)
    kc, vc = torch.randn(M, G, S_max, hs, generator=g), torch.randn(M, G, S_max, hs, generator=g)
    q = torch.randn(M, n_head * hs, generator=g)
    pos = torch.randint(0, S_max, (M,), generator=g).to(torch.int32)
    pos[0], pos[1], pos[2] = S_max - 1, 0, 63
    if M > 3:
        pos[3] = 64
    bidx = torch.randperm(M, generator=g).to(torch.int32)
    qpk = n_head // G
    ref = torch.empty(M, n_head * hs)
    for m in range(M):
        n = int(pos[m]) + 1
        for h in range(n_head):
            k, v = kc[int(bidx[m]), h // qpk, :n].double(), vc[int(bidx[m]), h // qpk, :n].double()
            w = torch.softmax(k @ q[m, h * hs:(h + 1) * hs].double() / hs ** 0.5, 0)
            ref[m, h * hs:(h + 1) * hs] = (w @ v).float()
    d = [t.cuda() for t in (q, kc, vc, pos, bidx)]
    ws = torch.empty(L.ua2_attn_workspace_floats(M, n_head, hs, S_max), device="cuda")
    outs = []
    try:
        for ring in (0, 1):
            _lib.check(L.ua2_set_global_option(b"attn_ring", ring))
            y = torch.full((M, n_head * hs), float("nan"), device="cuda")
            ws.fill_(float("nan"))
            _lib.check(L.ua2_attn_f32(P(d[0]), P(d[1]), P(d[2]), P(d[3]), P(d[4]), P(y), P(ws), M, n_head, G, hs, S_max, None))
            torch.cuda.synchronize()
            assert _rel(y.cpu(), ref) < 2e-5, ring
            outs.append(y.cpu())
    finally:
        _lib.check(L.ua2_set_global_option(b"attn_ring", 1))  # library default
    assert torch.equal(outs[0], outs[1])


def test_attn_ring_option_keeps_llm_golden_ids():
    """The batch-32 decode case of the LLM suite with the option OFF (the default is on): same token ids."""
    import subprocess
    import sys

    env = dict(os.environ, UA2_OPTIONS="attn_ring=0")  # applied by uniaudio2_b200/_lib.py at load
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_llm_gpu.py"), "-q", "-x", "-m", "gpu", "-k", "batch"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]


# ---- (6) option "conv_umma": causal convolutions as an implicit GEMM on tcgen05 straight from (B, C, T) (csrc/ua2_convumma.cu)
@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,elu,res", [
    (2, 64, 32, 1500, 3, 1, 1, 0),      # residual block, first convolution (weights resident, two accumulators)
    (1, 128, 256, 700, 1, 1, 1, 1),     # pointwise with residual, 256 output channels (single accumulator)
    (2, 64, 128, 2051, 8, 4, 1, 0),     # strided down-sampling, ragged length
    (1, 128, 256, 1285, 10, 5, 1, 0),   # 40 k-blocks: weights streamed through the ring
    (3, 32, 48, 333, 7, 1, 0, 0),       # Cout not a multiple of 32, no activation
    (1, 64, 16, 40, 3, 1, 1, 0),        # too few positions: stays on the SIMT core (the option must not change it)
])
def test_conv_umma_option_matches_oracle(B, Cin, Cout, T, K, stride, elu, res):
    import math

    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(Cin + Cout + T + K)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, dilation=1)
    r = torch.randn_like(ref) if res else None
    if res:
        ref = r + ref
    xd, wd, bd = x.cuda(), w.contiguous().cuda(), b.cuda()
    rd = r.cuda() if res else None
    outs = []
    try:
        for opt in (0, 1):
            _lib.check(L.ua2_set_global_option(b"conv_umma", opt))
            y = torch.full((B, Cout, ref.shape[-1]), float("nan"), device="cuda")
            _lib.check(L.ua2_conv1d_causal_gemm_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(rd), _lib.ptr(y), B, Cin, Cout, T, K,
                                                    stride, 1, elu, 0, None))
            torch.cuda.synchronize()
            outs.append(y.cpu())
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_umma", 1))
    # fp32-class: 3xTF32 + the TMEM accumulator's round-toward-zero per k-step (see test_conv_tc_option_matches_oracle) + the SFU exponential
    # of the fused ELU (abs 1e-7)
    tol = max(2e-5, 1.5 * (3 * Cin * K / 8) * 2.0 ** -24)
    for y in outs:
        assert bool(torch.isfinite(y).all())
        assert float((y - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))


# ---- (6b) the k = 1 streaming kernel of the narrow residual blocks (csrc/ua2_sgemm.cu: conv1d_pointwise_kernel), option "conv_pointwise"
@pytest.mark.parametrize("B,Cin,Cout,T,elu,res", [(2, 32, 64, 3001, 1, 1), (1, 64, 128, 700, 1, 1), (3, 96, 64, 255, 0, 0), (1, 32, 128, 1, 1, 1)])
def test_conv_pointwise_matches_oracle(B, Cin, Cout, T, elu, res):
    import math

    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(Cin + Cout + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, 1, generator=g) / math.sqrt(Cin)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=1, dilation=1)
    r = torch.randn_like(ref) if res else None
    if res:
        ref = r + ref
    xd, wd, bd = x.cuda(), w.contiguous().cuda(), b.cuda()
    rd = r.cuda() if res else None
    outs = []
    try:
        for opt in (1, 0):
            _lib.check(L.ua2_set_global_option(b"conv_pointwise", opt))
            y = torch.full((B, Cout, T), float("nan"), device="cuda")
            _lib.check(L.ua2_conv1d_causal_gemm_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(rd), _lib.ptr(y), B, Cin, Cout, T, 1, 1, 1, elu,
                                                    0, None))
            torch.cuda.synchronize()
            outs.append(y.cpu())
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_pointwise", 1))
    for y in outs:  # fp32 FMA chain over <= 96 terms: 2e-5 of the output scale (the bar of tests/test_codec_gpu.py)
        assert bool(torch.isfinite(y).all())
        assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Cin,Cout,T,stride", [(2, 128, 64, 1000, 4), (1, 64, 32, 777, 5), (1, 128, 64, 600, 8)])
def test_convtr_umma_option_matches_oracle(B, Cin, Cout, T, stride):
    """Transposed convolution (kernel = 2 x stride, causal trim) with all phases as columns of one implicit GEMM over the input grid."""
    import math

    from oracle import codec_oracle as CO
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(Cin + T + stride)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g) / math.sqrt(Cin * 2)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.convtr1d_causal(F.elu(x), w, b, stride=stride)
    xd, wd, bd = x.cuda(), w.contiguous().cuda(), b.cuda()
    wp = torch.empty(stride, Cout, Cin, 2, device="cuda")
    _lib.check(L.ua2_convtr1d_repack_phase_f32(_lib.ptr(wd), _lib.ptr(wp), Cin, Cout, stride, None))
    outs = []
    try:
        for opt in (0, 1):
            _lib.check(L.ua2_set_global_option(b"conv_umma", opt))
            y = torch.full((B, Cout, T * stride), float("nan"), device="cuda")
            _lib.check(L.ua2_convtr1d_causal_gemm_f32(_lib.ptr(xd), _lib.ptr(wp), _lib.ptr(bd), _lib.ptr(y), B, Cin, Cout, T, stride, 1, None))
            torch.cuda.synchronize()
            outs.append(y.cpu())
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_umma", 1))
    tol = max(2e-5, 1.5 * (3 * 2 * Cin / 8) * 2.0 ** -24)
    for y in outs:
        assert y.shape == ref.shape and bool(torch.isfinite(y).all())
        assert float((y - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))


# ---- (7) ScalarModel convolutions on the tensor cores: scalar nn.PReLU after the bias (activation(conv(x)), scalar24k.py:139-150; the skip
#      add comes after it), dilation, symmetric padding, channel counts that
#      are multiples of 16 (48, 96: zero-padded k-blocks), more than 256 output channels (column chunks / im2col GEMM), k = 1 + skip
@pytest.mark.parametrize("B,Cin,Cout,T,K,dil,causal,res", [
    (1, 96, 96, 3000, 7, 9, 1, 0),     # conv_umma, dilation 9, 96 -> 128-column tile
    (1, 96, 96, 3000, 1, 1, 1, 1),     # pointwise<48> with PReLU + skip
    (2, 48, 48, 2500, 7, 3, 1, 0),     # Cin 48: the second channel block of every tap is half zero padding
    (1, 48, 48, 5000, 1, 1, 1, 1),     # pointwise<48>, Cin % 32 != 0
    (2, 192, 192, 700, 7, 5, 0, 0),    # symmetric padding (non-causal configuration)
    (1, 384, 384, 9600, 7, 7, 1, 0),   # 384 output channels: two column chunks of the implicit GEMM
    (1, 768, 768, 300, 7, 1, 1, 0),    # few positions, long reduction: materialised im2col + stream-K GEMM
    (1, 384, 384, 3000, 1, 1, 1, 1),   # pointwise with 384 channels + skip
    (1, 136, 1536, 500, 5, 1, 0, 0),   # the delay convolution of the decoder (Cin 136: not a multiple of 16 -> fp32 SIMT core)
    (3, 64, 32, 1000, 3, 1, 1, 1),     # several batch elements, tile tails (1000 = 7 x 128 + 104), residual
    (1, 32, 64, 516, 7, 9, 0, 0),      # widest halo (54 samples) on both sides of a short sequence
    (2, 64, 64, 1028, 2, 1, 1, 0),     # two taps: every splitter group owns exactly one tap per stage
])
def test_scalar_model_convs_on_tensor_cores(B, Cin, Cout, T, K, dil, causal, res):
    import math

    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(Cin + Cout + T + K + dil)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g) * 0.1
    slope = torch.tensor([0.25])
    if causal:  # scalar24k.py:47-50: left pad dilation * (k - 1)
        pl, pr = dil * (K - 1), 0
    else:       # get_padding, scalar24k.py:18-19
        pl = pr = (K * dil - dil) // 2
    ref = F.prelu(F.conv1d(F.pad(x, (pl, pr)), w, b, dilation=dil), slope)
    r = torch.randn_like(ref) if res else None
    if res:
        ref = ref + r
    xd, wd, bd, sd = x.cuda(), w.contiguous().cuda(), b.cuda(), slope.cuda()
    rd = r.cuda() if res else None
    tol = max(2e-5, 1.5 * (3 * Cin * K / 8) * 2.0 ** -24)  # fp32 class: 3xTF32 + the TMEM accumulator's round-toward-zero per k-step
    try:
        for staged in (1, 0):  # activations through the TMA-staged shared-memory ring (default) / gathered into registers
            _lib.check(L.ua2_set_global_option(b"conv_umma_staged", staged))
            y = torch.full(tuple(ref.shape), float("nan"), device="cuda")
            _lib.check(L.ua2_conv1d_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(sd), _lib.ptr(rd), _lib.ptr(y), B, Cin, Cout, T, K, 1, dil, pl,
                                        pr, None))
            torch.cuda.synchronize()
            assert bool(torch.isfinite(y).all())
            err, scale = float((y.cpu() - ref).abs().max()), max(1.0, float(ref.abs().max()))
            assert err < tol * scale, (staged, err, tol, scale)
    finally:
        _lib.check(L.ua2_set_global_option(b"conv_umma_staged", 1))
