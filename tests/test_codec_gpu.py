"""GPU parity of the codec path (SEANet causal convs, Moshi-family transformer, residual VQ) through the C ABI against
  (a) golden vectors produced by the UNMODIFIED reference codec (tools/tokenizer/MimiCodec, tests/golden/codec_golden.pt),
  (b) the CPU oracle (oracle/codec_oracle.py) on fresh seeded inputs.
Bar (BASELINE.json): bit-exact VQ indices, waveform within 1e-4 max-abs (fp32)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import codec_oracle as CO
from oracle.make_golden_codec import codec_cfgs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from uniaudio2_b200 import _lib

    assert torch.cuda.is_available()
    return _lib.lib()


def _p(t):
    from uniaudio2_b200._lib import ptr

    return ptr(t)


def _chk(rc):
    from uniaudio2_b200._lib import check

    check(rc)


def _repack(w, transposed=False):
    # torch conv (Cout, Cin, K) / convtr (Cin, Cout, K) -> (Cin, K, Cout)
    return (w.permute(0, 2, 1) if transposed else w.permute(1, 2, 0)).contiguous()


@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,dil,elu,res,rep", [
    (2, 1, 64, 1000, 7, 1, 1, 0, 0, 0),     # SEANet first conv
    (1, 64, 32, 777, 3, 1, 1, 1, 0, 0),     # resblock k3
    (1, 32, 64, 777, 1, 1, 1, 1, 1, 0),     # resblock k1 + skip
    (2, 64, 128, 1003, 8, 4, 1, 1, 0, 0),   # strided down conv (ragged length -> extra right pad)
    (1, 128, 256, 501, 10, 5, 1, 1, 0, 0),
    (1, 96, 200, 333, 12, 6, 1, 1, 0, 0),   # channel counts that are not multiples of the tile
    (1, 512, 1024, 64, 16, 8, 1, 1, 0, 0),
    (1, 1024, 512, 33, 3, 1, 1, 1, 0, 0),   # last encoder conv
    (2, 128, 128, 57, 4, 2, 1, 0, 0, 1),    # ConvDownsample1d: replicate pad, no bias
    (1, 16, 24, 200, 3, 1, 2, 1, 0, 0),     # dilation
    (1, 64, 1, 999, 3, 1, 1, 1, 0, 0),      # last decoder conv (Cout = 1)
])
def test_conv1d_causal(L, B, Cin, Cout, T, K, stride, dil, elu, res, rep):
    g = torch.Generator().manual_seed(Cin + Cout + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = None if rep else torch.randn(Cout, generator=g) * 0.1
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, dilation=dil, pad_mode="replicate" if rep else "constant")
    r = torch.randn_like(ref) if res else None
    if res:
        ref = r + ref
    xd, wd = x.cuda(), _repack(w).cuda()
    bd = b.cuda() if b is not None else None
    rd = r.cuda() if res else None
    y = torch.empty(B, Cout, ref.shape[-1], device="cuda")
    _chk(L.ua2_conv1d_causal_f32(_p(xd), _p(wd), _p(bd), _p(rd), _p(y), B, Cin, Cout, T, K, stride, dil, elu, rep, None))
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert float((y.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    # implicit-GEMM variant (what the codec handle runs), weights in the reference layout
    y2 = torch.full_like(y, float("nan"))
    wt = w.contiguous().cuda()
    _chk(L.ua2_conv1d_causal_gemm_f32(_p(xd), _p(wt), _p(bd), _p(rd), _p(y2), B, Cin, Cout, T, K, stride, dil, elu, rep, None))
    torch.cuda.synchronize()
    assert float((y2.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Cin,Cout,T,stride", [(1, 1024, 512, 20, 8), (2, 512, 256, 77, 6), (1, 256, 128, 300, 5),
                                                 (1, 128, 64, 1000, 4), (1, 48, 20, 65, 3), (1, 32, 32, 40, 2)])
def test_convtr1d_causal(L, B, Cin, Cout, T, stride):
    g = torch.Generator().manual_seed(Cin + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g) / math.sqrt(Cin * 2)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.convtr1d_causal(F.elu(x), w, b, stride)
    y = torch.empty(B, Cout, T * stride, device="cuda")
    xd, wd, bd = x.cuda(), _repack(w, True).cuda(), b.cuda()
    _chk(L.ua2_convtr1d_causal_f32(_p(xd), _p(wd), _p(bd), _p(y), B, Cin, Cout, T, stride, 1, None))
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert float((y.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    # phase-GEMM variant (what the codec handle runs)
    wt = w.contiguous().cuda()
    wp = torch.empty(stride * Cout * Cin * 2, device="cuda")
    _chk(L.ua2_convtr1d_repack_phase_f32(_p(wt), _p(wp), Cin, Cout, stride, None))
    y2 = torch.full_like(y, float("nan"))
    _chk(L.ua2_convtr1d_causal_gemm_f32(_p(xd), _p(wp), _p(bd), _p(y2), B, Cin, Cout, T, stride, 1, None))
    torch.cuda.synchronize()
    assert float((y2.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def test_convtr1d_depthwise(L):
    g = torch.Generator().manual_seed(3)
    B, C, T, s = 2, 512, 37, 2
    x = torch.randn(B, C, T, generator=g)
    w = torch.randn(C, 1, 2 * s, generator=g)
    ref = CO.convtr1d_causal(x, w, None, s, groups=C)
    y = torch.empty(B, C, T * s, device="cuda")
    xd, wd = x.cuda(), w.cuda()
    _chk(L.ua2_convtr1d_depthwise_f32(_p(xd), _p(wd), _p(y), B, C, T, s, None))
    torch.cuda.synchronize()
    assert float((y.cpu() - ref).abs().max()) < 1e-6


@pytest.mark.parametrize("B,D,T,K,n_q", [(1, 256, 125, 2048, 8), (3, 32, 50, 64, 6), (2, 64, 33, 300, 5), (1, 256, 7, 2048, 31)])
def test_rvq_encode_decode(L, B, D, T, K, n_q):
    """argmin indices equal the reference algorithm's (cdist + argmin, core_vq.py:179-185) for every quantizer, incl. the
    residual chain; decode equals the sum of the selected codebook rows."""
    g = torch.Generator().manual_seed(D + K)
    x = torch.randn(B, D, T, generator=g)
    emb = torch.randn(n_q, K, D, generator=g)
    # oracle on CPU (same loop as CO.rvq_encode without the projection)
    residual, ref_codes = x.clone(), []
    for q in range(n_q):
        flat = residual.transpose(1, 2).reshape(-1, D)
        codes = torch.cdist(flat[None], emb[q][None], p=2)[0].argmin(-1).view(B, T)
        residual = residual - F.embedding(codes, emb[q]).transpose(1, 2)
        ref_codes.append(codes)
    ref_codes = torch.stack(ref_codes, 1)
    xd, ed = x.cuda(), emb.cuda()
    sq = (ed * ed).sum(-1).contiguous()
    codes = torch.full((B, n_q + 2, T), -1, dtype=torch.int64, device="cuda")
    _chk(L.ua2_rvq_encode_f32(_p(xd), _p(ed), _p(sq), _p(codes), B, D, T, K, n_q, n_q + 2, 1, None))
    torch.cuda.synchronize()
    assert torch.equal(codes[:, 1:1 + n_q].cpu(), ref_codes)
    assert int(codes[:, 0].max()) == -1 and int(codes[:, -1].max()) == -1  # rows outside [q_off, q_off+n_q) untouched
    # many-frame variant: frame-major residual, tiled-GEMM scores + argmin/update kernels
    if D % 8 == 0 and K % 2 == 0:
        r_md = x.transpose(1, 2).reshape(B * T, D).contiguous().cuda()
        S = torch.empty(B * T, K, device="cuda")
        codes2 = torch.full((B, n_q + 2, T), -1, dtype=torch.int64, device="cuda")
        _chk(L.ua2_rvq_encode_gemm_f32(_p(r_md), _p(ed), _p(sq), _p(S), _p(codes2), B, D, T, K, n_q, n_q + 2, 1, None))
        torch.cuda.synchronize()
        assert torch.equal(codes2[:, 1:1 + n_q].cpu(), ref_codes)
        assert float((r_md.cpu() - residual.transpose(1, 2).reshape(B * T, D)).abs().max()) < 1e-4
    out = torch.empty(B, D, T, device="cuda")
    _chk(L.ua2_rvq_decode_f32(_p(codes), _p(ed), _p(out), B, D, T, K, n_q, n_q + 2, 1, None))
    torch.cuda.synchronize()
    ref = sum(F.embedding(ref_codes[:, q], emb[q]).transpose(1, 2) for q in range(n_q))
    assert float((out.cpu() - ref).abs().max()) < 1e-5


def _build(cfg, sd):
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    m = MimiCodec(n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size,
                  codebook_dim=cfg.codebook_dim, rvq_layers=cfg.rvq_layers, num_heads=cfg.num_heads, num_layers=cfg.num_layers,
                  layer_scale=cfg.layer_scale, context=cfg.context, device="cuda")
    full = m.state_dict()
    full.update({k: v.cuda() for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    return m


@pytest.fixture(scope="module")
def codec_golden():
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return torch.load(os.path.join(root, "tests", "golden", "codec_golden.pt"), weights_only=False)


@pytest.mark.parametrize("cname", ["tiny", "mid"])
def test_codec_matches_reference_golden(codec_golden, cname):
    cfg = codec_cfgs()[cname]
    sd = CO.random_mimi_state_dict(cfg, seed=4321)
    assert {k: float(v.double().sum()) for k, v in sd.items()} == codec_golden[f"__checksum_{cname}"]
    m = _build(cfg, sd)
    for key, fx in codec_golden.items():
        if key.startswith("__") or fx["cfg_name"] != cname:
            continue
        codes = m.encode(fx["wav"].cuda())
        assert torch.equal(codes.cpu(), fx["codes"]), f"{key}: VQ indices differ from the reference"
        wav = m.decode(fx["codes"].cuda())
        torch.cuda.synchronize()
        assert wav.shape == fx["recon"].shape
        err = float((wav.cpu() - fx["recon"]).abs().max())
        assert err < 1e-4, f"{key}: waveform max-abs error {err:.2e} (bar 1e-4)"


def test_codec_full_config_vs_oracle():
    """mimi_config.yaml geometry (rates [8,6,5,4], 32 x 2048 x 256 RVQ, 8-layer d=512 transformer), 1 s clip + ragged tail."""
    cfg = CO.MimiCfg()
    sd = CO.random_mimi_state_dict(cfg, seed=7)
    m = _build(cfg, sd)
    orc = CO.MimiOracle(cfg, sd)
    g = torch.Generator().manual_seed(5)
    wav = torch.randn(4, 1, 3 * 24000 + 311, generator=g) * 0.2  # 4 x 79 frames @25 Hz (tiled GEMM transformer), 4 x 40 code frames (GEMM RVQ)
    with torch.no_grad():
        ref_codes = orc.encode(wav)
        ref_wav = orc.decode(ref_codes)
    codes = m.encode(wav.cuda())
    assert codes.shape == ref_codes.shape
    match = float((codes.cpu() == ref_codes).float().mean())
    assert torch.equal(codes.cpu(), ref_codes), f"VQ index agreement {match:.4f}"
    out = m.decode(ref_codes.cuda())
    torch.cuda.synchronize()
    assert float((out.cpu() - ref_wav).abs().max()) < 1e-4


def test_codec_error_paths():
    cfg = codec_cfgs()["tiny"]
    m = _build(cfg, CO.random_mimi_state_dict(cfg, seed=4321))
    with pytest.raises(ValueError):
        m.encode(torch.zeros(2, 2, 100))
    with pytest.raises(ValueError):
        m.decode(torch.zeros(1, 3, 5, dtype=torch.long))
