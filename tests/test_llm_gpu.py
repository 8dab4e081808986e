"""GPU parity of the drop-in Model_stage3 (uniaudio2_b200) through the C ABI against
  (a) the committed golden fixtures produced by the UNMODIFIED reference, and
  (b) the CPU oracle run on fresh seeded inputs.
Bar (BASELINE.json north_star): bit-exact token ids (greedy and shared-noise top-k); hidden states / logits
within 1e-4 relative (fp32, different summation order)."""
import pytest
import torch

from conftest import build_product_model
from oracle import llm_oracle as O
from oracle.cases import CASES, REASON_CARD, run_case, tiny_cfgs

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


TC_DEFAULT = 1  # library default of the global option tc_gemm
PF_DEFAULT = (0, 0)  # library defaults of gemv3_prefetch_mb / gemv3_prefetch_idle_mb


@pytest.fixture(scope="module")
def models():
    out = {}
    for cname, cfg in tiny_cfgs().items():
        sd = O.random_state_dict(cfg, seed=1234)
        out[cname] = (cfg, sd, build_product_model(cfg, sd, "cuda", 3))
    return out


@pytest.mark.parametrize("mode", ["eager", "graph", "graph_direct_local_attn"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_frames_match_reference_golden(models, golden, case, mode):
    """eager: one launch per kernel; graph: CUDA-graph replay with PDL edges; graph_direct_local_attn: the local decoder's
    <= 8-key attention computed inside the proj kernel's prologue instead of its own split-softmax launch (option
    attn_direct = 1), with the tail-prefetch planner enabled (a no-op unless built with UA2_GEMV3_TAIL_PREFETCH)."""
    name, cname, kind, B, S, nf, topk, temp, cfg_scale = case
    cfg, sd, m = models[cname]
    fx = golden[name]
    from uniaudio2_b200 import _lib

    m.set_option("graph", 0 if mode.startswith("eager") else 1)
    m.set_option("attn_direct", 1 if mode == "graph_direct_local_attn" else 0)
    pf = 48 if mode == "graph_direct_local_attn" else 0
    _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_prefetch_mb", pf))
    _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_prefetch_idle_mb", pf))
    try:
        r = run_case(m, kind, cfg, B, S, nf, topk, temp, cfg_scale, REASON_CARD[cname], 42, True, device="cuda",
                     explicit_noise=True)
        torch.cuda.synchronize()
    finally:
        m.set_option("attn_direct", 0)
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_prefetch_mb", PF_DEFAULT[0]))
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_prefetch_idle_mb", PF_DEFAULT[1]))
    got = r["frames"].cpu()
    assert torch.equal(got, fx["ref_frames"]), (
        f"{name}: token ids differ from the reference (min top-1/top-2 margins: text {fx['margin_text']:.2e}, "
        f"audio {fx['margin_audio']:.2e})\n{got.tolist()}\nvs\n{fx['ref_frames'].tolist()}")
    # hidden-state parity through the KV caches and the last frame's logits
    k, _ = m.kv_cache(0, cfg.backbone.n_layer - 1)
    assert _rel(k[:B, :, : S + nf].cpu(), fx["last_backbone_k"]) < REL_TOL
    _, v = m.kv_cache(3, cfg.generation.n_layer - 1)
    assert _rel(v[:B, :, : S + nf].cpu(), fx["last_gen_v"]) < REL_TOL
    assert _rel(m.debug_buffer("h_final", B).cpu(), fx["h_final"][-1][:, -1]) < REL_TOL
    assert _rel(m.debug_buffer("text_logits", B).cpu(), fx["text_logits"][-1]) < REL_TOL
    assert _rel(m.debug_buffer("audio_logits", B).cpu(), fx["ci_logits"][-1]) < REL_TOL


def test_fresh_inputs_vs_oracle_long_context(models):
    """Fresh seeded prompt that crosses the 128-key split boundary of the attention kernel (S = 140 > ATTN_CHUNK)
    with a chunk-crossing prefill; compared with the oracle run in-process."""
    cfg0, sd, _ = models["tiny"]
    import copy

    cfg = copy.deepcopy(cfg0)
    cfg.max_seq_length = 320
    m = build_product_model(cfg, sd, "cuda", 2, max_seq=320)
    m.set_option("prefill_chunk_rows", 256)  # the 300-row prompt below is prefilled in two passes
    orc = O.Stage3Oracle(cfg, sd)
    orc.setup_caches(2)
    for (kind, B, S, nf, topk, temp) in (("mixed", 2, 140, 6, 1, 1.0), ("text", 1, 300, 4, 8, 0.9)):
        o = run_case(orc, kind, cfg, B, S, nf, topk, temp, 1.0, REASON_CARD["tiny"], 7, False, explicit_noise=True)
        r = run_case(m, kind, cfg, B, S, nf, topk, temp, 1.0, REASON_CARD["tiny"], 7, True, device="cuda", explicit_noise=True)
        assert torch.equal(r["frames"].cpu(), o["frames"])
        k, v = m.kv_cache(0, cfg.backbone.n_layer - 1)
        assert _rel(k[:B, :, : S + nf].cpu(), orc.backbone.kv[-1].k[:B, :, : S + nf]) < REL_TOL
        assert _rel(m.debug_buffer("text_logits", B).cpu(), o["text_logits"][-1]) < REL_TOL


def test_reset_caches_zero_fills(models):
    cfg, sd, m = models["tiny"]
    run_case(m, "text", cfg, 1, 8, 2, 1, 1.0, 1.0, REASON_CARD["tiny"], 3, True, device="cuda")
    k, v = m.kv_cache(0, 0)
    assert float(k.abs().sum()) > 0
    m.reset_caches()
    torch.cuda.synchronize()
    for which in range(4):
        k, v = m.kv_cache(which, 0)
        assert float(k.abs().sum()) == 0.0 and float(v.abs().sum()) == 0.0


def test_weights_loaded_after_setup_caches_are_used():
    """The handle keeps raw parameter pointers and a transposed audio_head copy: load_state_dict / .to() after setup_caches
    must not leave it stale (round-1 advisor finding).  Both drop the handle; after a new setup_caches the ids follow the NEW weights."""
    cfg = tiny_cfgs()["tiny"]
    sd_a, sd_b = O.random_state_dict(cfg, seed=1), O.random_state_dict(cfg, seed=2)
    m = build_product_model(cfg, sd_a, "cuda", 1)
    m.load_state_dict(sd_b, strict=True)
    with pytest.raises(TypeError):
        m.reset_caches()
    m.setup_caches(1)
    orc = O.Stage3Oracle(cfg, sd_b)
    orc.setup_caches(1)
    got = run_case(m, "text", cfg, 1, 8, 3, 1, 1.0, 1.0, REASON_CARD["tiny"], 11, True, device="cuda")
    ref = run_case(orc, "text", cfg, 1, 8, 3, 1, 1.0, 1.0, REASON_CARD["tiny"], 11, False)
    assert torch.equal(got["frames"].cpu(), ref["frames"])
    m.float()  # _apply: the tensors may have moved
    with pytest.raises(TypeError):
        m.reset_caches()


def test_error_behaviour(models):
    """ValueError cases of model_new.py:165-180 and the position range check of lit_model.py:143-144."""
    cfg, sd, m = models["tiny"]
    tok = torch.zeros(1, 1, 9, dtype=torch.long, device="cuda")
    msk = torch.zeros(1, 1, 9, dtype=torch.bool, device="cuda")
    msk[..., -1] = True
    pos = torch.tensor([3], device="cuda")
    for kw in (dict(temperature=0.0, topk=1), dict(temperature=1.0, topk=0), dict(temperature=1.0, topk=1, forbid_prefix=-1),
               dict(temperature=1.0, topk=1, forbid_prefix=cfg.audio_vocab), dict(temperature=1.0, topk=100, forbid_prefix=40)):
        with pytest.raises(ValueError):
            m.generate_frame(tok, msk, pos, 4, **kw)
    with pytest.raises(ValueError):
        m.generate_frame(tok, msk, torch.tensor([cfg.max_seq_length]), None, temperature=1.0, topk=1)
    with pytest.raises(ValueError):
        m.generate_frame(torch.zeros(4, 1, 9, dtype=torch.long), torch.zeros(4, 1, 9, dtype=torch.bool), pos, 4, temperature=1.0, topk=1)


def test_rng_modes_run(models):
    """'torch' mode draws like model_new.py:141-143 on the device; 'philox' draws inside the kernel."""
    cfg, sd, m = models["tiny"]
    for mode in ("torch", "philox"):
        m.rng_mode = mode
        r = run_case(m, "text", cfg, 1, 8, 4, 10, 0.9, 1.0, REASON_CARD["tiny"], 5, True, device="cuda")
        f = r["frames"].cpu()
        assert f.shape == (4, 1, 9) and int(f[:, :, 1:].min()) >= 0 and int(f[:, :, 1:].max()) < cfg.audio_vocab
        # second half of the frames runs with forbid_prefix = reason_card
        assert int(f[2:, :, 1:].min()) >= REASON_CARD["tiny"]
    m.rng_mode = "torch"
    assert m.last_launch_count() >= 19


def test_task_generators(models):
    """The mirrored task drivers (evaluation/tts_task.py, evaluation/asr_task.py) against the oracle driven by the same
    loops: prompt packing, phase switch, fixed synthetic schedule, greedy text decode."""
    from types import SimpleNamespace

    from uniaudio2_b200.evaluation import asr_task, tts_task

    cfg, sd, m = models["tiny"]
    args = SimpleNamespace(text_pad_token=3, semantic_pad_token=80, semantic_eos=81, semantic_bos=82, reason_eos=37, reason_bos=38,
                           reason_pad_token=36, parallel_number=9, audio_reason_card=REASON_CARD["tiny"], audio_semantic_card=90)
    g = torch.Generator().manual_seed(9)
    prompt = torch.randint(0, 1000, (5,), generator=g)
    text = torch.randint(0, 1000, (7,), generator=g)
    # ---- TTS: 3 reason-phase + 4 semantic-phase frames, greedy
    gen = tts_task.Generator(m, args)
    gen.special_token_dict = {k: 1000 + i for i, k in enumerate(tts_task.SPECIAL_TOKENS)}  # ids inside the tiny vocab
    r, s = gen.generate_tts(prompt, "TTS", text_token=text, temperature=1.0, topk=1, fixed_schedule=(3, 4))
    assert r.shape == (8, 1) and s.shape == (8, 3) and gen.n_frames == 7
    orc = O.Stage3Oracle(cfg, sd)
    orc.setup_caches(1)
    tokens, mask = gen.prepare_tts_task(prompt, text)
    S = tokens.size(0)
    tokens, mask = tokens.unsqueeze(0), mask.bool().unsqueeze(0)
    orc.reset_caches()
    orc.forward_prefix(tokens[:, :-1], mask, torch.arange(S).unsqueeze(0)[:, :-1])
    ct, cm, frames = tokens[:, -1:], mask[:, -1:], []
    for f in range(7):
        smp = orc.generate_frame(ct, cm, torch.tensor([S - 1 + f]), S + f, 1.0, 1, 0 if f < 3 else REASON_CARD["tiny"])
        frames.append(smp[0, 1:].long())
        ct = torch.cat([smp[:, 1:], smp[:, 0:1]], -1).long().unsqueeze(1)
        cm = torch.cat([torch.ones(1, 1, 8, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1)
    assert torch.equal(r.cpu(), torch.stack(frames[1:2]).t())  # frames 1..2 saved minus the first; frame 3 is the switch
    assert torch.equal(s.cpu(), torch.stack(frames[4:7]).t() - REASON_CARD["tiny"])
    # ---- ASR-style text decode, greedy, 6 frames
    agen = asr_task.Generator(m, args)
    reason = torch.randint(0, 36, (4, 8), generator=g)
    sem = torch.randint(0, 80, (6, 8), generator=g)
    ids = agen.generate_asr(prompt, "ASR", semantic_token=sem, reason_token=reason, temperature=1.0, topk=1, max_audio_frames=6)
    tokens, mask = agen.prepare_asr_task(prompt, reason, sem)
    S = tokens.size(0)
    assert S == 5 + 6 + 8
    tokens, mask = tokens.unsqueeze(0), mask.bool().unsqueeze(0)
    orc.reset_caches()
    orc.forward_prefix(tokens[:, :-1], mask, torch.arange(S).unsqueeze(0)[:, :-1])
    ct, cm, ref_ids = tokens[:, -1:], mask[:, -1:], []
    for f in range(6):
        smp = orc.generate_frame(ct, cm, torch.tensor([S - 1 + f]), S + f, 1.0, 1, 0)
        t = int(smp[0, 0])
        if t == asr_task.EOS_TEXT:
            break
        ref_ids.append(t)
        ct = torch.zeros(1, 1, 9, dtype=torch.long)
        ct[0, 0, -1] = t
        cm = torch.cat([torch.zeros(1, 1, 8, dtype=torch.bool), torch.ones(1, 1, 1, dtype=torch.bool)], -1)
    assert ids == ref_ids and len(ids) == 6


def test_device_side_tts_loop_matches_host_loop(models):
    """Generator.generate_tts with the sample feedback and the phase / EOS state machine on the device (ua2_llm_tts_frames, one 16-byte
    D2H per chunk of frames) returns the tokens of the per-frame host loop - fixed schedule (random weights never emit EOS), greedy and
    sampled (torch noise and in-kernel Philox)."""
    from types import SimpleNamespace

    from uniaudio2_b200.evaluation import tts_task

    cfg, sd, m = models["tiny"]
    args = SimpleNamespace(text_pad_token=3, semantic_pad_token=80, semantic_eos=81, semantic_bos=82, reason_eos=37, reason_bos=38,
                           reason_pad_token=36, parallel_number=9, audio_reason_card=REASON_CARD["tiny"], audio_semantic_card=90)
    gen = tts_task.Generator(m, args)
    gen.special_token_dict = {k: 1000 + i for i, k in enumerate(tts_task.SPECIAL_TOKENS)}  # ids inside the tiny vocab
    g = torch.Generator().manual_seed(3)
    prompt, text = torch.randint(0, 100, (5,), generator=g), torch.randint(0, 100, (6,), generator=g)
    for rng_mode, topk in (("torch", 1), ("torch", 20), ("philox", 20)):
        m.rng_mode = rng_mode
        outs = []
        for dev_loop in (False, True):
            torch.manual_seed(11)
            m.seed = 888
            r, s = gen.generate_tts(prompt, "TTS", text_token=text, temperature=0.9, topk=topk, fixed_schedule=(5, 9), device_loop=dev_loop,
                                    sync_every=4)
            outs.append((r.cpu(), s.cpu(), gen.n_frames))
        if rng_mode == "torch":  # Philox draws are keyed by the library's frame counter, which keeps running between the two runs
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (rng_mode, topk)
        assert outs[0][2] == outs[1][2] == 14
        assert outs[1][0].shape == outs[0][0].shape == (8, 3) and outs[1][1].shape == outs[0][1].shape == (8, 8)
    m.rng_mode = "torch"


def test_out_of_range_token_id_is_reported():
    """nn.Embedding raises IndexError for an id outside its table; the kernel reads row 0 instead of foreign memory and the NEXT call
    of the handle raises ValueError (the flag travels through mapped host memory, no synchronisation on the hot path)."""
    cfg = tiny_cfgs()["tiny"]
    sd = O.random_state_dict(cfg, seed=5)
    m = build_product_model(cfg, sd, "cuda", 1)
    tok = torch.zeros(1, 1, 9, dtype=torch.long, device="cuda")
    msk = torch.ones(1, 1, 9, dtype=torch.bool, device="cuda")
    tok[0, 0, 2] = cfg.audio_vocab  # one past the last audio id
    m.generate_frame(tok, msk, torch.tensor([0]), 1, temperature=1.0, topk=1)
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="out of range"):
        m.reset_caches()
    m.reset_caches()  # reported once, then cleared
    tok[0, 0, 2], tok[0, 0, 8] = 0, cfg.backbone.padded_vocab_size + 5
    m.generate_frame(tok, msk, torch.tensor([0]), 1, temperature=1.0, topk=1)
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="out of range"):
        m.generate_frame(tok, msk, torch.tensor([1]), 2, temperature=1.0, topk=1)


def test_device_side_tts_state_machine_follows_the_reference_rules(models):
    """The EOS / phase rules of tts_task.py:253-279 on scripted rows: frames_out / state after every call of the device kernel equal a
    restatement of the reference's loop body (break on all == end_tok before recording; all == reason_eos switches the phase)."""
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    nq, reason_eos, end_tok, card = 8, 5, 100 + 7, 100
    script = [[1] * 9, [0] + [3] * 8, [0] + [reason_eos] * 8, [2] + [150] * 8, [9] + [end_tok] * 7 + [101], [4] + [end_tok] * 8, [1] + [120] * 8]
    state = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = torch.zeros(16, nq + 1, dtype=torch.int32, device="cuda")
    ref_rows, forbid, done, switch = [], 0, False, 0
    for row in script:
        smp = torch.tensor(row, dtype=torch.int32, device="cuda")
        _lib.check(L.ua2_tts_state_step(_lib.ptr(smp), nq, _lib.ptr(state), _lib.ptr(out), 16, reason_eos, end_tok, card, -1, None))
        if not done:
            if all(v == end_tok for v in row[1:]):
                done = True
            else:
                ref_rows.append(row)
                if all(v == reason_eos for v in row[1:]):
                    forbid, switch = card, len(ref_rows)
        st = state.cpu().tolist()
        assert st == [forbid, int(done), len(ref_rows), switch], (row, st)
    assert out[: len(ref_rows)].cpu().tolist() == ref_rows and len(ref_rows) == 5 and done


def test_batch32_caption_config():
    """SURVEY section 8(d) config 3 shape: 32 equal-length mixed prompts, batched prefill (B*S rows go through the tiled
    GEMM path) + greedy frames at B = 32 (four M tiles of 8 rows per linear), against the oracle run in-process.
    Also a ragged batch (B = 11: tiles of 8 + 3)."""
    import copy

    cfg = copy.deepcopy(tiny_cfgs()["tiny"])
    sd = O.random_state_dict(cfg, seed=77)
    m = build_product_model(cfg, sd, "cuda", 32)
    orc = O.Stage3Oracle(cfg, sd)
    orc.setup_caches(32)
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    # (tc_gemm, tc_min_rows): skinny kernels; the tcgen05 mainloop (fp32 weights split on chip, csrc/ua2_umma.cu) for frames of >= 8 rows
    for (tc, min_rows) in ((0, 128), (1, 8)):
        _lib.check(L.ua2_set_global_option(b"tc_gemm", tc))
        _lib.check(L.ua2_set_global_option(b"tc_min_rows", min_rows))
        try:
            for (B, S, nf) in ((32, 21, 4), (11, 13, 3)):
                o = run_case(orc, "mixed", cfg, B, S, nf, 1, 1.0, 1.0, REASON_CARD["tiny"], 11, False, explicit_noise=True)
                r = run_case(m, "mixed", cfg, B, S, nf, 1, 1.0, 1.0, REASON_CARD["tiny"], 11, True, device="cuda", explicit_noise=True)
                assert torch.equal(r["frames"].cpu(), o["frames"]), (tc, min_rows, B)
                assert _rel(m.debug_buffer("text_logits", B).cpu(), o["text_logits"][-1]) < REL_TOL
        finally:
            _lib.check(L.ua2_set_global_option(b"tc_gemm", TC_DEFAULT))
            _lib.check(L.ua2_set_global_option(b"tc_min_rows", 32))


def test_cache_filled_to_the_last_slot():
    """Maximum size: prefill + frames up to position max_seq_length - 1 (the last KV slot, lit_model.py:120-124 allows
    input_pos < max_seq_length), ids equal to the oracle; the next position is rejected like the reference does."""
    import copy

    cfg = copy.deepcopy(tiny_cfgs()["tiny"])
    cfg.max_seq_length = 72  # crosses the 64-key attention split at the very end
    sd = O.random_state_dict(cfg, seed=3)
    m = build_product_model(cfg, sd, "cuda", 1, max_seq=72)
    orc = O.Stage3Oracle(cfg, sd)
    orc.setup_caches(1)
    S, nf = 66, 7  # frames at positions 65 .. 71
    o = run_case(orc, "text", cfg, 1, S, nf, 4, 0.8, 1.0, REASON_CARD["tiny"], 13, False, explicit_noise=True)
    r = run_case(m, "text", cfg, 1, S, nf, 4, 0.8, 1.0, REASON_CARD["tiny"], 13, True, device="cuda", explicit_noise=True)
    assert torch.equal(r["frames"].cpu(), o["frames"])
    k, _ = m.kv_cache(0, cfg.backbone.n_layer - 1)
    assert _rel(k[:1, :, :72].cpu(), orc.backbone.kv[-1].k[:1, :, :72]) < REL_TOL
    tok = torch.zeros(1, 1, 9, dtype=torch.long, device="cuda")
    msk = torch.ones(1, 1, 9, dtype=torch.bool, device="cuda")
    with pytest.raises(ValueError):
        m.generate_frame(tok, msk, torch.tensor([72], device="cuda"), None, temperature=1.0, topk=1)
