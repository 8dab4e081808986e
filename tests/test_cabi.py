"""CPU suite: the C-ABI library loads, exports every symbol include/ua2_b200.h declares, and its argument
validation (the reference's ValueError paths) works without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from uniaudio2_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from uniaudio2_b200.build import build

        build()
    return _lib.lib()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ua2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ua2_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(L):
    from uniaudio2_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_lib.SYMBOLS.keys()) == declared
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ua2_b200.h but not exported by the .so"


def test_library_matches_the_sources():
    """The in-tree .so was built from the sources, headers and flags the tree holds now (object stamps of uniaudio2_b200/build.py):
    a failed or forgotten rebuild must not let a stale library stand in for the code under review."""
    from uniaudio2_b200 import build as B

    if not os.path.isdir(B.OBJ):
        pytest.skip("library was not built by uniaudio2_b200.build in this tree (prebuilt .so only)")
    assert B.stale_sources() == [], "run `python -m uniaudio2_b200.build` (and read its output)"


def test_version_and_error_string(L):
    assert b"sm_100a" in L.ua2_version()
    assert isinstance(L.ua2_last_error(), bytes)


def test_create_validates_config(L):
    from uniaudio2_b200 import _lib

    good = _lib.GptCfg(2, 256, 4, 2, 64, 512, 1e-5)
    bad_hs = _lib.GptCfg(2, 256, 4, 2, 48, 512, 1e-5)
    h = C.c_void_p()
    cfg = _lib.LlmCfg(good, good, good, good, 1024, 128, 8, 64)
    assert L.ua2_llm_create(C.byref(cfg), C.byref(h)) == 0
    # setup before weights are registered must fail loudly, not crash
    with pytest.raises(ValueError):
        _lib.check(L.ua2_llm_setup_caches(h, 1, None))
    # unknown key
    shape = (C.c_int64 * 2)(4, 4)
    with pytest.raises(ValueError):
        _lib.check(L.ua2_llm_load_weight(h, b"no.such.weight", C.c_void_p(16), shape, 2))
    # wrong shape for a known key
    with pytest.raises(ValueError):
        _lib.check(L.ua2_llm_load_weight(h, b"backbone.transformer.h.0.attn.qkv.weight", C.c_void_p(16), shape, 2))
    assert L.ua2_llm_destroy(h) == 0
    cfg2 = _lib.LlmCfg(bad_hs, good, good, good, 1024, 128, 8, 64)
    with pytest.raises(ValueError):
        _lib.check(L.ua2_llm_create(C.byref(cfg2), C.byref(h)))
    assert b"head_size" in L.ua2_last_error()


def test_tokenize_handles_validate_config_and_keys_without_gpu(L):
    """ua2_wavlm_* / ua2_thinking_* / the front-end operators: configuration, key, shape and size errors are rejected before any CUDA call."""
    from uniaudio2_b200 import _lib

    arr = lambda v: (C.c_int32 * 8)(*(list(v) + [0] * (8 - len(v))))
    dims, ker, strd = arr([512] * 7), arr([10, 3, 3, 3, 3, 2, 2]), arr([5, 2, 2, 2, 2, 2, 2])
    h = C.c_void_p()
    good = _lib.WavLMCfg(768, 12, 3072, 12, 7, dims, ker, strd, 0, 128, 16, 320, 800, 1e-5)
    assert L.ua2_wavlm_create(C.byref(good), C.byref(h)) == 0
    assert L.ua2_wavlm_frames(h, 480160) == 1500 and L.ua2_wavlm_frames(h, 300) == 0  # 30 s + 160 zeros -> 1500 frames; too short -> 0
    with pytest.raises(ValueError):  # forward before finalize
        _lib.check(L.ua2_wavlm_forward(h, C.c_void_p(16), 480160, 1, 480160, 6, 10, C.c_void_p(16), None, None))
    shape3 = (C.c_int64 * 3)(512, 1, 10)
    assert L.ua2_wavlm_load_weight(h, b"feature_extractor.conv_layers.0.conv.weight", C.c_void_p(16), shape3, 3) == 0
    with pytest.raises(ValueError):  # conv_bias = False in this configuration
        _lib.check(L.ua2_wavlm_load_weight(h, b"feature_extractor.conv_layers.0.conv.bias", C.c_void_p(16), (C.c_int64 * 1)(512), 1))
    with pytest.raises(ValueError):  # wrong shape
        _lib.check(L.ua2_wavlm_load_weight(h, b"encoder.layers.0.attention.rel_attn_embed.weight", C.c_void_p(16), (C.c_int64 * 2)(320, 8), 2))
    with pytest.raises(ValueError):  # only layer 0 carries the bucket embedding
        _lib.check(L.ua2_wavlm_load_weight(h, b"encoder.layers.1.attention.rel_attn_embed.weight", C.c_void_p(16), (C.c_int64 * 2)(320, 12), 2))
    with pytest.raises(ValueError):  # missing parameters
        _lib.check(L.ua2_wavlm_finalize(h, None))
    with pytest.raises(ValueError):
        _lib.check(L.ua2_wavlm_set_option(h, b"no_such_option", 1))
    assert L.ua2_wavlm_destroy(h) == 0
    for bad in (_lib.WavLMCfg(768, 12, 3072, 12, 7, dims, ker, strd, 0, 128, 8, 320, 800, 1e-5),     # 96 channels per positional-convolution group
                _lib.WavLMCfg(768, 10, 3072, 12, 7, dims, ker, strd, 0, 128, 16, 320, 800, 1e-5),    # hidden / heads not an integer
                _lib.WavLMCfg(768, 12, 3072, 12, 7, dims, arr([20, 3, 3, 3, 3, 2, 2]), strd, 0, 128, 16, 320, 800, 1e-5)):  # first kernel > 16
        with pytest.raises(ValueError):
            _lib.check(L.ua2_wavlm_create(C.byref(bad), C.byref(h)))
    tc = _lib.ThinkingCfg(768, 128, 5, 5, 1024, 1024, 4)
    assert L.ua2_thinking_create(C.byref(tc), C.byref(h)) == 0
    assert L.ua2_thinking_rows(h, 1500, 750) == 900 and L.ua2_thinking_rows(h, 1500, 751) == 900 and L.ua2_thinking_rows(h, 1500, 748) == 0
    with pytest.raises(ValueError):
        _lib.check(L.ua2_thinking_load_weight(h, b"encoder_transformers.5.ff_scale.scale", C.c_void_p(16), (C.c_int64 * 1)(768), 1))  # 5 blocks: 0..4
    with pytest.raises(ValueError):
        _lib.check(L.ua2_thinking_encode(h, C.c_void_p(16), C.c_void_p(16), 1, 1500, 750, C.c_void_p(16), None))  # before finalize
    assert L.ua2_thinking_destroy(h) == 0
    with pytest.raises(ValueError):
        _lib.check(L.ua2_thinking_create(C.byref(_lib.ThinkingCfg(768, 64, 5, 5, 1024, 1024, 4)), C.byref(h)))  # dim_heads is 128 in the reference
    p = C.c_void_p(16)
    with pytest.raises(ValueError):  # n_valid beyond ceil(new * L / orig)
        _lib.check(L.ua2_resample_f32(p, 300, p, p, 400, 1, 300, 300, 400, 3, 2, 10, None))
    with pytest.raises(ValueError):  # more frames than 1 + L / hop
        _lib.check(L.ua2_whisper_logmel_f32(p, 16000, p, p, p, 1, 16000, 400, 160, 80, 102, None))
    with pytest.raises(ValueError):  # odd n_fft
        _lib.check(L.ua2_whisper_logmel_f32(p, 16000, p, p, p, 1, 16000, 401, 160, 80, 100, None))


def test_sampler_argument_errors_without_gpu(L):
    """model_new.py:165-180 error cases are rejected before any CUDA call."""
    from uniaudio2_b200 import _lib

    p = C.c_void_p(16)
    for (temp, topk, forbid) in ((0.0, 1, 0), (1.0, 1, -1), (1.0, 1, 64), (1.0, 0, 0), (1.0, 33, 32)):
        with pytest.raises(ValueError):
            _lib.check(L.ua2_sample_topk_f32(p, 1, 64, temp, topk, forbid, 1.0, None, 0, 0, p, None))


def test_product_has_no_cpu_fallback():
    import torch

    from uniaudio2_b200 import _lib
    from uniaudio2_b200.llm_models import config as pc
    from uniaudio2_b200.llm_models.model_new import Model_stage3, ModelArgs

    pc.name_to_config["t-bb"] = dict(name="t-bb", n_layer=1, n_embd=128, n_head=2, n_query_groups=1, intermediate_size=256, padded_vocab_size=64)
    saved = {k: pc.name_to_config[k] for k in ("Llama-3.2-Understanding", "Llama-3.2-Generation")}
    pc.name_to_config["Llama-3.2-Understanding"] = dict(pc.name_to_config["t-bb"])
    pc.name_to_config["Llama-3.2-Generation"] = dict(pc.name_to_config["t-bb"])
    try:
        m = Model_stage3(ModelArgs("t-bb", "t-bb", "", "", "", 30, 2, 8))
    finally:
        pc.name_to_config.update(saved)
    with pytest.raises(_lib.Ua2Error):
        m.setup_caches(1)  # parameters on the CPU -> refuse
    with pytest.raises(TypeError):
        m.reset_caches()  # lit_model.py:134-135 'You need to call set_kv_cache'


def test_product_does_not_import_oracle():
    """The product package must not reference oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "uniaudio2_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} imports the oracle"
    # dev tools measure / profile the product: they must not pull the oracle in either
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "tools", f)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, f"tools/{f} imports the oracle"
    # bench.py: only inside the CPU-baseline legs (cpu_baseline_sample, the `if cpu:` blocks of bench_codec / bench_flow_decoder)
    import ast

    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        imports = [n for n in ast.walk(fn) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
        if imports:
            assert fn.name in ("cpu_baseline_sample", "bench_codec", "bench_flow_decoder", "bench_whisper_encoder", "bench_tokenize_frontends"), \
                f"bench.py::{fn.name} imports the oracle"
            for imp in imports:  # ... and there only under the `if cpu:` guard of the baseline leg
                guards = [n for n in ast.walk(fn) if isinstance(n, ast.If) and imp in list(ast.walk(n))]
                assert fn.name == "cpu_baseline_sample" or any(isinstance(g.test, ast.Name) and g.test.id == "cpu" for g in guards), fn.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n)]
    assert not top, "bench.py imports the oracle at module level"


def test_ua2_options_environment_is_applied_at_load():
    """UA2_OPTIONS="name=value,..." (uniaudio2_b200/_lib.py): applied through ua2_set_global_option when the library loads; an
    unknown option or a malformed item fails the import instead of being ignored."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "from uniaudio2_b200 import _lib; _lib.lib(); print('loaded')"
    ok = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, env=dict(os.environ, UA2_OPTIONS="attn_ring=1, conv_tc=0"))
    assert ok.returncode == 0 and "loaded" in ok.stdout, ok.stderr[-2000:]
    bad = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, env=dict(os.environ, UA2_OPTIONS="no_such_option=1"))
    assert bad.returncode != 0 and "no_such_option" in bad.stderr
    bad = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, env=dict(os.environ, UA2_OPTIONS="attn_ring"))
    assert bad.returncode != 0
