import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "llm_golden.pt"), weights_only=False)


def build_product_model_cpu(cfg):
    """The drop-in Model_stage3 for an oracle Stage3Cfg with its parameters on the CPU (no handle, no kernels)."""
    return _make_product_model(cfg, "cpu", None)


def build_product_model(cfg, sd, device, max_batch, max_seq=None):
    """Instantiate the product's Model_stage3 for an oracle Stage3Cfg and load the oracle's state dict."""
    m = _make_product_model(cfg, device, max_seq)
    m.load_state_dict(sd, strict=True)
    m.setup_caches(max_batch)
    return m


def _make_product_model(cfg, device, max_seq):
    from uniaudio2_b200.llm_models import config as pc
    from uniaudio2_b200.llm_models.model_new import Model_stage3, ModelArgs

    def d(c, name):
        return dict(name=name, n_layer=c.n_layer, n_embd=c.n_embd, n_head=c.n_head, n_query_groups=c.n_query_groups,
                    head_size=c.head_size, intermediate_size=c.intermediate_size, padded_vocab_size=c.padded_vocab_size,
                    norm_eps=c.norm_eps, rope_base=c.rope_base, rope_adjustments=c.rope_adjustments)

    pc.name_to_config["ua2-test-backbone"] = d(cfg.backbone, "ua2-test-backbone")
    pc.name_to_config["ua2-test-decoder"] = d(cfg.decoder, "ua2-test-decoder")
    saved = {k: pc.name_to_config[k] for k in ("Llama-3.2-Understanding", "Llama-3.2-Generation")}
    pc.name_to_config["Llama-3.2-Understanding"] = d(cfg.understanding, "Llama-3.2-Understanding")
    pc.name_to_config["Llama-3.2-Generation"] = d(cfg.generation, "Llama-3.2-Generation")
    try:
        args = ModelArgs("ua2-test-backbone", "ua2-test-decoder", "", "", "", cfg.audio_vocab - 7, 7, cfg.num_codebooks)
        m = Model_stage3(args, device=device, max_seq_length=max_seq or cfg.max_seq_length)
    finally:
        pc.name_to_config.update(saved)
    return m


@pytest.fixture(scope="session")
def encoder_handles_shim(tmp_path_factory):
    """TEST INFRASTRUCTURE: whole translation units of the product with the real headers (-DUA2_CPU_SHIM), like the third build of
    tests/test_kernels_on_cpu_shim.py - csrc/ua2_wavlm.cu and csrc/ua2_thinking.cu (kernels AND handles) over the real linear launchers
    (ua2_gemv.cu, ua2_gemv3.cu, ua2_sgemm.cu, ua2_attn.cu, ua2_misc.cu); the tcgen05 GEMM is a CPU GEMM (cpu_shim/stubs_real_headers.cpp),
    the dense attention comes from the kernel part of ua2_dit.cu.  Built once per session (40 s of g++)."""
    import ctypes as C
    import re
    import subprocess

    shim, csrc = os.path.join(ROOT, "tests", "cpu_shim"), os.path.join(ROOT, "uniaudio2_b200", "csrc")
    gxx = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread"] + (["-fsanitize=address", "-fno-omit-frame-pointer", "-g"]
                                                                          if os.environ.get("UA2_SHIM_ASAN") == "1" else [])
    strip = r'#include ["<](\.\./\.\./include/ua2_b200\.h|ua2_kernels\.cuh|ua2_umma\.cuh|ua2_enc_dev\.cuh|ua2_philox\.cuh|cuda_bf16\.h)[">]'
    d = str(tmp_path_factory.mktemp("shim_encoders"))
    hdr = os.path.join(ROOT, "include", "ua2_b200.h")
    srcs = []
    for name in ("ua2_codec", "ua2_sgemm", "ua2_convtc", "ua2_resblock", "ua2_attn", "ua2_gemv3", "ua2_gemv", "ua2_misc", "ua2_codec_model", "ua2_wavlm",
                 "ua2_thinking"):
        src = open(os.path.join(csrc, name + ".cu")).read()
        src = re.sub(r"extern __shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = static_cast<\1*>(shim::g_dyn_smem);", src)
        src = src.replace('#include "../../include/ua2_b200.h"', f'#include "{hdr}"')
        open(os.path.join(d, name + ".cpp"), "w").write(src)
        srcs.append(os.path.join(d, name + ".cpp"))
    dit = open(os.path.join(csrc, "ua2_dit.cu")).read()
    dit = re.sub(strip, "", re.sub(r"extern __shared__[^;]*;", "", dit[:dit.index("\nusing namespace ua2;")]))
    open(os.path.join(d, "ua2_dit_kernels.inc"), "w").write(dit)
    for stub in ("stubs_real_headers.cpp", "stubs_wavlm.cpp"):
        text = open(os.path.join(shim, stub)).read().replace('#include "../../include/ua2_b200.h"', f'#include "{hdr}"')
        open(os.path.join(d, stub), "w").write(text)
        srcs.append(os.path.join(d, stub))
    so = os.path.join(d, "libshim_encoders.so")
    r = subprocess.run(gxx + ["-DUA2_CPU_SHIM", "-DUA2_ATTN_RING_MIN_ITEMS=64", "-I", d, "-I", csrc, "-I", os.path.join(shim, "rt"), "-Wl,--no-undefined"] + srcs +
                       ["-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    lib = C.CDLL(so)
    lib.ua2_last_error.restype = C.c_char_p
    lib.shim_set_sm_count(4)  # persistent kernels size their grids by the SM count: keep the OS-thread emulation small
    lib.ua2_wavlm_frames.restype = C.c_longlong
    lib.ua2_wavlm_frames.argtypes = [C.c_void_p, C.c_longlong]
    lib.ua2_wavlm_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ua2_wavlm_load_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    lib.ua2_thinking_rows.restype = C.c_longlong
    lib.ua2_thinking_load_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    lib.ua2_thinking_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib
