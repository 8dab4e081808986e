"""GPU parity of the stand-alone C-ABI operators against the CPU oracle (oracle/llm_oracle.py primitives)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import llm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from uniaudio2_b200 import _lib

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.lib()


def _p(t):
    from uniaudio2_b200._lib import ptr

    return ptr(t)


def _chk(rc):
    from uniaudio2_b200._lib import check

    check(rc)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


# fp32 accumulation in a different order than the CPU oracle: relative error bound for K <= 8192 dot products
TOL = 2e-5


@pytest.mark.parametrize("M,N,K", [(1, 3072, 3072), (1, 5120, 3072), (1, 2048, 8192), (2, 512, 384), (3, 130, 256),
                                   (4, 1024, 2048), (5, 258, 640), (8, 384, 1280), (11, 96, 132), (1, 12300, 2048),
                                   (37, 768, 512), (200, 1536, 512), (333, 130, 256), (128, 3072, 1024)])
@pytest.mark.parametrize("norm,res", [(False, False), (True, False), (False, True), (True, True)])
def test_linear(L, M, N, K, norm, res):
    g = torch.Generator().manual_seed(M * 1000 + N + K)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    nw = 1 + 0.1 * torch.randn(K, generator=g) if norm else None
    r = torch.randn(M, N, generator=g) if res else None
    ref = F.linear(O.rms_norm(x, nw, 1e-5) if norm else x, W)
    if res:
        ref = ref + r
    xd, Wd = x.cuda(), W.cuda()
    nwd = nw.cuda() if norm else None
    rd = r.cuda() if res else None
    y = torch.empty(M, N, device="cuda")
    _chk(L.ua2_linear_f32(_p(xd), _p(Wd), _p(nwd), 1e-5, _p(rd), _p(y), M, N, K, None))
    torch.cuda.synchronize()
    assert _rel(y.cpu(), ref) < TOL


def test_linear_inplace_residual(L):
    g = torch.Generator().manual_seed(5)
    M, N, K = 2, 384, 512
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    r = torch.randn(M, N, generator=g)
    ref = F.linear(x, W) + r
    y = r.clone().cuda()
    xd, Wd = x.cuda(), W.cuda()
    _chk(L.ua2_linear_f32(_p(xd), _p(Wd), None, 0.0, _p(y), _p(y), M, N, K, None))
    torch.cuda.synchronize()
    assert _rel(y.cpu(), ref) < TOL


@pytest.mark.parametrize("M,N,K", [(1, 8192, 3072), (1, 8192, 2048), (2, 384, 256), (4, 513, 384), (9, 640, 768), (150, 640, 768), (257, 129, 512)])
@pytest.mark.parametrize("norm", [False, True])
def test_swiglu(L, M, N, K, norm):
    g = torch.Generator().manual_seed(N + K)
    x = torch.randn(M, K, generator=g)
    W1 = torch.randn(N, K, generator=g) / math.sqrt(K)
    W2 = torch.randn(N, K, generator=g) / math.sqrt(K)
    nw = 1 + 0.1 * torch.randn(K, generator=g) if norm else None
    xn = O.rms_norm(x, nw, 1e-5) if norm else x
    ref = F.silu(F.linear(xn, W1)) * F.linear(xn, W2)
    y = torch.empty(M, N, device="cuda")
    xd, W1d, W2d = x.cuda(), W1.cuda(), W2.cuda()
    nwd = nw.cuda() if norm else None
    _chk(L.ua2_swiglu_f32(_p(xd), _p(W1d), _p(W2d), _p(nwd), 1e-5, _p(y), M, N, K, None))
    torch.cuda.synchronize()
    assert _rel(y.cpu(), ref) < TOL


@pytest.mark.parametrize("hs,n_head,G,D,B,T", [(128, 24, 8, 3072, 1, 1), (64, 32, 8, 2048, 2, 1), (64, 6, 2, 384, 2, 5),
                                               (128, 6, 2, 768, 3, 7), (32, 8, 2, 256, 1, 9), (64, 8, 2, 512, 4, 40)])
def test_qkv_rope_and_attention(L, hs, n_head, G, D, B, T):
    """Kernel A (RMSNorm->QKV->RoPE->cache append) then kernel B (attention) vs lit_model.py:424-532 restated."""
    S_max, past = max(160, 131 + T + 8), 131  # > ATTN_CHUNK so the split path is exercised
    g = torch.Generator().manual_seed(hs + n_head + D)
    cfg = O.GPTCfg(n_layer=1, n_embd=D, n_head=n_head, n_query_groups=G, intermediate_size=4 * D, head_size=hs)
    x = torch.randn(B, T, D, generator=g)
    Wqkv = torch.randn((n_head + 2 * G) * hs, D, generator=g) / math.sqrt(D)
    nw = 1 + 0.1 * torch.randn(D, generator=g)
    cos, sin = O.build_rope_cache(S_max, hs, cfg.rope_base, cfg.rope_adjustments)
    kv = O.KV(B, G, S_max, hs)
    kv.k[:, :, :past] = torch.randn(B, G, past, hs, generator=g)
    kv.v[:, :, :past] = torch.randn(B, G, past, hs, generator=g)
    k0, v0 = kv.k.clone(), kv.v.clone()
    # per-row positions: row b starts at past - b (ragged across the batch)
    pos = torch.stack([torch.arange(T) + past - b for b in range(B)])  # (B,T)
    # ---- oracle
    xn = O.rms_norm(x, nw, 1e-5)
    qkv = F.linear(xn, Wqkv)
    q, k, v = qkv.split((n_head * hs, G * hs, G * hs), dim=-1)
    q = q.view(B, T, n_head, hs).transpose(1, 2)
    k = k.view(B, T, G, hs).transpose(1, 2)
    v = v.view(B, T, G, hs).transpose(1, 2)
    q = O.apply_rope(q, cos[pos], sin[pos])
    k = O.apply_rope(k, cos[pos], sin[pos])
    kf, vf = kv.write(pos, k, v)
    mask = torch.tril(torch.ones(S_max, S_max, dtype=torch.bool))[pos].unsqueeze(1)  # (B,1,T,S)
    rep = n_head // G
    yref = F.scaled_dot_product_attention(q, kf.repeat_interleave(rep, 1), vf.repeat_interleave(rep, 1), attn_mask=mask,
                                          scale=1.0 / math.sqrt(hs)).transpose(1, 2).reshape(B * T, n_head * hs)
    # ---- device
    M = B * T
    xd = x.reshape(M, D).cuda()
    Wd, nwd, cosd, sind = Wqkv.cuda(), nw.cuda(), cos.cuda(), sin.cuda()
    posd = pos.reshape(-1).to(torch.int32).cuda()
    bidx = torch.arange(B).repeat_interleave(T).to(torch.int32).cuda()
    kc, vc = k0.cuda(), v0.cuda()
    qd = torch.empty(M, n_head * hs, device="cuda")
    _chk(L.ua2_qkv_rope_f32(_p(xd), _p(Wd), _p(nwd), 1e-5, _p(posd), _p(bidx), _p(cosd), _p(sind), _p(qd), _p(kc), _p(vc),
                            M, D, n_head, G, hs, S_max, None))
    torch.cuda.synchronize()
    assert _rel(qd.cpu(), q.transpose(1, 2).reshape(M, -1)) < TOL
    assert _rel(kc.cpu(), kv.k) < TOL and _rel(vc.cpu(), kv.v) < TOL
    # untouched slots stay bit-identical
    untouched = torch.ones(S_max, dtype=torch.bool)
    untouched[pos.reshape(-1).unique()] = False
    assert torch.equal(kc.cpu()[:, :, untouched], k0[:, :, untouched])
    ws = torch.empty(L.ua2_attn_workspace_floats(M, n_head, hs, S_max), device="cuda")
    y = torch.empty(M, n_head * hs, device="cuda")
    _chk(L.ua2_attn_f32(_p(qd), _p(kc), _p(vc), _p(posd), _p(bidx), _p(y), _p(ws), M, n_head, G, hs, S_max, None))
    torch.cuda.synchronize()
    assert _rel(y.cpu(), yref) < 5e-5


def _noise(R, V, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.empty(R, V).exponential_(1, generator=g)


@pytest.mark.parametrize("V", [130, 800, 12300, 16384, 20000, 128256])
@pytest.mark.parametrize("topk,temp,forbid", [(1, 1.0, 0), (1, 0.7, 41), (5, 0.9, 0), (50, 0.9, 40), (200, 1.3, 0)])
def test_sampler_bit_exact(L, V, topk, temp, forbid):
    """Token ids equal the oracle's (model_new.py:146-187) for identical logits and identical Exp(1) draws."""
    R = 3
    if topk > V - forbid:
        pytest.skip("topk larger than effective vocab")
    g = torch.Generator().manual_seed(V + topk)
    logits = torch.randn(R, V, generator=g) * 3.0
    q = _noise(R, V, 7 + V)
    ref = O.audio_sample_topk(logits, topk, temp, forbid, noise=q).squeeze(1)
    out = torch.empty(R, dtype=torch.int32, device="cuda")
    ld, qd = logits.cuda(), q.cuda()
    _chk(L.ua2_sample_topk_f32(_p(ld), R, V, temp, topk, forbid, 1.0, _p(qd), 0, 0, _p(out), None))
    torch.cuda.synchronize()
    assert out.cpu().tolist() == ref.tolist()


def test_sampler_ties_at_threshold(L):
    """Ties at the k-th value are all kept (`logits < kth` removed, model_new.py:150) and the first index wins
    an exact p/q tie (torch.argmax)."""
    V = 4096
    logits = torch.full((1, V), -2.0)
    logits[0, [5, 900, 901, 3000]] = 1.5  # 4 equal maxima, topk=2 -> all 4 kept
    q = torch.ones(1, V)
    q[0, 3000] = 0.25  # smallest noise wins
    ref = O.audio_sample_topk(logits, 2, 1.0, 0, noise=q)
    out = torch.empty(1, dtype=torch.int32, device="cuda")
    ld, qd = logits.cuda(), q.cuda()
    _chk(L.ua2_sample_topk_f32(_p(ld), 1, V, 1.0, 2, 0, 1.0, _p(qd), 0, 0, _p(out), None))
    assert out.item() == ref.item() == 3000
    q2 = torch.ones(1, V)  # exact tie in p/q between the four -> first index
    ref2 = O.audio_sample_topk(logits, 2, 1.0, 0, noise=q2)
    qd2 = q2.cuda()
    _chk(L.ua2_sample_topk_f32(_p(ld), 1, V, 1.0, 2, 0, 1.0, _p(qd2), 0, 0, _p(out), None))
    assert out.item() == ref2.item() == 5


@pytest.mark.parametrize("V", [12300, 128256])
@pytest.mark.parametrize("topk", [3, 128, 129])
def test_sampler_candidate_path_edges(L, V, topk):
    """The candidate-list fast path (k <= 128) and its overflow fallback: (a) a plateau of equal logits far longer than
    the list (every element ties with the k-th value -> all kept, model_new.py:150), (b) a row whose top values sit in
    one 4-thread group, (c) the largest k served by the fast path and the first k that is not."""
    g = torch.Generator().manual_seed(V * 7 + topk)
    rows = []
    plateau = torch.full((V,), 0.25)
    plateau[torch.randperm(V, generator=g)[:5]] = -3.0
    rows.append(plateau)
    clustered = torch.randn(V, generator=g)
    clustered[1000:1000 + 4 * 140:4] += 20.0  # 140 large values at stride 4: few thread groups own all of the top k
    rows.append(clustered)
    rows.append(torch.randn(V, generator=g) * 4)
    logits = torch.stack(rows)
    R = logits.shape[0]
    q = _noise(R, V, 99 + topk)
    ref = O.audio_sample_topk(logits, topk, 0.9, 0, noise=q).squeeze(1)
    out = torch.empty(R, dtype=torch.int32, device="cuda")
    ld, qd = logits.cuda(), q.cuda()
    _chk(L.ua2_sample_topk_f32(_p(ld), R, V, 0.9, topk, 0, 1.0, _p(qd), 0, 0, _p(out), None))
    torch.cuda.synchronize()
    assert out.cpu().tolist() == ref.tolist()


@pytest.mark.parametrize("V", [800, 128256])
def test_sampler_cfg(L, V):
    """CFG mix of model_new.py:618-622: u + (c-u)*scale with row 0 = cond, row 1 = uncond; one sample, repeated."""
    g = torch.Generator().manual_seed(V)
    logits = torch.randn(2, V, generator=g) * 2
    q = _noise(1, V, 3)
    mixed = logits[1:] + (logits[0:1] - logits[1:]) * 1.5
    ref = O.audio_sample_topk(mixed, 20, 0.8, 10, noise=q)
    out = torch.full((2,), -1, dtype=torch.int32, device="cuda")
    ld, qd = logits.cuda(), q.cuda()
    _chk(L.ua2_sample_topk_f32(_p(ld), 1, V, 0.8, 20, 10, 1.5, _p(qd), 0, 0, _p(out), None))
    assert out.cpu().tolist() == [ref.item(), ref.item()]


def test_sampler_internal_philox_statistics(L):
    """Without a noise tensor the kernel draws Exp(1) from Philox itself: frequencies follow softmax(logits)
    (same style of KAT as llm_utils/sampling.py:156-174, bound 1.5e-2)."""
    ps = torch.tensor([5.0, 2, 12, 6, 8, 1, 0, 4])
    p = ps / ps.sum()
    R = 4000
    logits = torch.log(p.clamp_min(1e-30)).unsqueeze(0).repeat(R, 1)
    logits[:, 6] = -float("inf")
    out = torch.empty(R, dtype=torch.int32, device="cuda")
    ld = logits.cuda()
    counts = torch.zeros(8)
    for it in range(5):
        _chk(L.ua2_sample_topk_f32(_p(ld), R, 8, 1.0, 7, 0, 1.0, None, 1234, it, _p(out), None))
        counts += torch.bincount(out.cpu().long(), minlength=8).float()
    assert (counts / counts.sum() - p).abs().max() < 1.5e-2
