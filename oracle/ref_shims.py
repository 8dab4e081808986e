"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Import shims that let the *unmodified* reference at /root/reference be imported in the
build container (it does not exist on the GPU box).  Used only by oracle/make_golden.py to
(a) validate the oracle restatement against the real reference code and (b) generate the
committed fixtures under tests/golden/.

Shims:
  litgpt.config  -> llm_models.config      (llm_models/lit_model.py:18)
  litgpt.model   -> llm_models.lit_model   (llm_models/config.py:178-205 resolves LLaMAMLP/RMSNorm there)
  litgpt.scripts.convert_hf_checkpoint.qkv_reassemble -> stub (legacy ckpt only, lit_model.py:556-565)
  torchtune      -> empty module           (model_new.py:18, unused)
  modules / utils.compile -> llm_modules / llm_utils.compile (Moshi family, llm_modules/__init__.py)
"""
import os
import sys
import types

REF_ROOT = os.environ.get("UA2_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "llm_models"))


def install_llm_shims():
    """Make `from llm_models.model_new import Model_stage3` importable."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    os.environ.setdefault("NO_TORCH_COMPILE", "1")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib

    litgpt = types.ModuleType("litgpt")
    litgpt.__path__ = []
    sys.modules["litgpt"] = litgpt
    cfg = importlib.import_module("llm_models.config")
    sys.modules["litgpt.config"] = cfg
    litgpt.config = cfg
    scripts = types.ModuleType("litgpt.scripts")
    scripts.__path__ = []
    conv = types.ModuleType("litgpt.scripts.convert_hf_checkpoint")

    def qkv_reassemble(*a, **k):  # legacy-checkpoint helper, never hit on this path
        raise NotImplementedError("qkv_reassemble stub")

    conv.qkv_reassemble = qkv_reassemble
    sys.modules["litgpt.scripts"] = scripts
    sys.modules["litgpt.scripts.convert_hf_checkpoint"] = conv
    sys.modules.setdefault("torchtune", types.ModuleType("torchtune"))
    lit_model = importlib.import_module("llm_models.lit_model")
    sys.modules["litgpt.model"] = lit_model
    litgpt.model = lit_model
    return importlib.import_module("llm_models.model_new")


def install_mimi_shims():
    """Make tools/tokenizer/MimiCodec importable (needs only einops)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    os.environ.setdefault("NO_TORCH_COMPILE", "1")
    p = os.path.join(REF_ROOT, "tools", "tokenizer", "MimiCodec")
    if p not in sys.path:
        sys.path.insert(0, p)


def install_moshi_shims():
    """Make the named Moshi-family modules importable as the reference itself addresses them (`from modules.gating import
    ...`, `from utils.compile import ...`, llm_modules/transformer.py:21-24): `modules` -> llm_modules (by __path__, so the
    package __init__ with its dead imports never runs), `utils.compile` -> llm_utils.compile.  Returns
    (modules.transformer, llm_utils.sampling)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    os.environ.setdefault("NO_TORCH_COMPILE", "1")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib

    mods = types.ModuleType("modules")
    mods.__path__ = [os.path.join(REF_ROOT, "llm_modules")]
    sys.modules["modules"] = mods
    utils = types.ModuleType("utils")
    utils.__path__ = []
    sys.modules["utils"] = utils
    comp = importlib.import_module("llm_utils.compile")
    sys.modules["utils.compile"] = comp
    utils.compile = comp
    tr = importlib.import_module("modules.transformer")
    samp = importlib.import_module("llm_utils.sampling")
    return tr, samp
