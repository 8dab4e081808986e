"""ctypes binding of libua2_b200.so (the C ABI declared in include/ua2_b200.h).

There is NO CPU fallback: if the library is missing or a call fails this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libua2_b200.so")

_lib = None


class Ua2Error(RuntimeError):
    pass


class GptCfg(C.Structure):
    _fields_ = [("n_layer", C.c_int32), ("n_embd", C.c_int32), ("n_head", C.c_int32), ("n_query_groups", C.c_int32),
                ("head_size", C.c_int32), ("intermediate_size", C.c_int32), ("norm_eps", C.c_float)]


class LlmCfg(C.Structure):
    _fields_ = [("backbone", GptCfg), ("decoder", GptCfg), ("understanding", GptCfg), ("generation", GptCfg),
                ("text_vocab", C.c_int32), ("audio_vocab", C.c_int32), ("num_codebooks", C.c_int32),
                ("max_seq_length", C.c_int32)]


class CodecCfg(C.Structure):
    _fields_ = [("n_filters", C.c_int32), ("ratios", C.c_int32 * 8), ("n_ratios", C.c_int32), ("latent_dim", C.c_int32),
                ("codebook_size", C.c_int32), ("codebook_dim", C.c_int32), ("rvq_layers", C.c_int32), ("num_heads", C.c_int32),
                ("num_layers", C.c_int32), ("context", C.c_int32), ("dim_feedforward", C.c_int32),
                ("resample_stride", C.c_int32), ("max_period", C.c_float)]


class StxCfg(C.Structure):
    _fields_ = [("d_model", C.c_int32), ("num_heads", C.c_int32), ("num_layers", C.c_int32), ("causal", C.c_int32),
                ("context", C.c_int32), ("positional_embedding", C.c_int32), ("norm", C.c_int32), ("gating", C.c_int32),
                ("weights_per_step", C.c_int32), ("layer_scale", C.c_int32), ("dim_feedforward", C.c_int32 * 64),
                ("max_period", C.c_float), ("positional_scale", C.c_float)]


class DitCfg(C.Structure):
    _fields_ = [("num_attention_heads", C.c_int32), ("attention_head_dim", C.c_int32), ("in_channels", C.c_int32),
                ("out_channels", C.c_int32), ("num_layers", C.c_int32), ("num_positional_embeddings", C.c_int32),
                ("flow_t_size", C.c_int32), ("norm_eps", C.c_float)]


class WhisperCfg(C.Structure):
    _fields_ = [("d_model", C.c_int32), ("encoder_attention_heads", C.c_int32), ("encoder_ffn_dim", C.c_int32),
                ("encoder_layers", C.c_int32), ("max_source_positions", C.c_int32), ("num_mel_bins", C.c_int32)]


class WavLMCfg(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_attention_heads", C.c_int32), ("intermediate_size", C.c_int32),
                ("num_hidden_layers", C.c_int32), ("num_feat_extract_layers", C.c_int32), ("conv_dim", C.c_int32 * 8),
                ("conv_kernel", C.c_int32 * 8), ("conv_stride", C.c_int32 * 8), ("conv_bias", C.c_int32),
                ("num_conv_pos_embeddings", C.c_int32), ("num_conv_pos_embedding_groups", C.c_int32), ("num_buckets", C.c_int32),
                ("max_bucket_distance", C.c_int32), ("layer_norm_eps", C.c_float)]


class ThinkingCfg(C.Structure):
    _fields_ = [("dim", C.c_int32), ("dim_heads", C.c_int32), ("depth", C.c_int32), ("interval", C.c_int32), ("whisper_dim", C.c_int32),
                ("mu_dim", C.c_int32), ("ff_mult", C.c_int32)]


# every symbol include/ua2_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ua2_last_error": (C.c_char_p, []),
    "ua2_device_sm_count": (C.c_int, []),
    "ua2_version": (C.c_char_p, []),
    "ua2_set_global_option": (C.c_int, [C.c_char_p, C.c_int]),
    "ua2_llm_create": (C.c_int, [C.POINTER(LlmCfg), C.POINTER(_P)]),
    "ua2_llm_destroy": (C.c_int, [_P]),
    "ua2_llm_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_llm_setup_caches": (C.c_int, [_P, C.c_int, _P]),
    "ua2_llm_reset_caches": (C.c_int, [_P, _P]),
    "ua2_llm_prefill": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, _P]),
    "ua2_llm_generate_frame": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, C.c_float, C.c_int, C.c_int, C.c_float, _P,
                                         C.c_uint64, _P, _P]),
    "ua2_llm_get_kv": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P)]),
    "ua2_llm_get_buffer": (C.c_int, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "ua2_llm_tts_frames": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_float, C.c_int, _P, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                     C.c_int, _P, _P, C.c_int, _P]),
    "ua2_llm_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ua2_llm_last_launch_count": (C.c_int, [_P]),
    "ua2_tts_state_step": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_linear_f32": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_tc_linear_f32": (C.c_int, [_P, _P, _P, _P, C.c_float, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_flash_attn_bf16": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_swiglu_f32": (C.c_int, [_P, _P, _P, _P, C.c_float, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_qkv_rope_f32": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, _P]),
    "ua2_attn_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_attn_workspace_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ua2_conv1d_causal_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, _P]),
    "ua2_conv1d_causal_gemm_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, _P]),
    "ua2_resblock_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_convtr1d_causal_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_convtr1d_repack_phase_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_convtr1d_causal_gemm_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_conv1d_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, _P]),
    "ua2_convtr1d_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_elementwise_f32": (C.c_int, [_P, _P, C.c_longlong, C.c_int, C.c_float, _P]),
    "ua2_film_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "ua2_interp_nearest_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "ua2_linear_bias_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_convtr1d_depthwise_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_rvq_encode_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_rvq_encode_gemm_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_rvq_decode_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_codec_create": (C.c_int, [C.POINTER(CodecCfg), C.POINTER(_P)]),
    "ua2_codec_destroy": (C.c_int, [_P]),
    "ua2_codec_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_codec_finalize": (C.c_int, [_P, _P]),
    "ua2_codec_frames": (C.c_int64, [_P, C.c_int64]),
    "ua2_codec_encode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "ua2_codec_decode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "ua2_sample_topk_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, _P, C.c_uint64,
                                      C.c_uint64, _P, _P]),
    "ua2_rope_ring_append_f32": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_ring_attn_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                    C.c_int, _P]),
    "ua2_sample_token_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, _P, C.c_uint64,
                                       C.c_uint64, _P, _P]),
    "ua2_dit_create": (C.c_int, [C.POINTER(DitCfg), C.POINTER(_P)]),
    "ua2_dit_destroy": (C.c_int, [_P]),
    "ua2_dit_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_dit_finalize": (C.c_int, [_P, _P]),
    "ua2_dit_forward": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "ua2_dit_solve_euler": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_float), C.c_int, _P, C.c_int, C.c_float, C.c_float, _P]),
    "ua2_dit_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ua2_dit_last_launch_count": (C.c_int, [_P]),
    "ua2_whisper_create": (C.c_int, [C.POINTER(WhisperCfg), C.POINTER(_P)]),
    "ua2_whisper_destroy": (C.c_int, [_P]),
    "ua2_whisper_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_whisper_finalize": (C.c_int, [_P, _P]),
    "ua2_whisper_forward": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "ua2_whisper_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ua2_whisper_last_launch_count": (C.c_int, [_P]),
    "ua2_resample_f32": (C.c_int, [_P, C.c_longlong, _P, _P, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_whisper_logmel_f32": (C.c_int, [_P, C.c_longlong, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_wavlm_create": (C.c_int, [C.POINTER(WavLMCfg), C.POINTER(_P)]),
    "ua2_wavlm_destroy": (C.c_int, [_P]),
    "ua2_wavlm_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_wavlm_finalize": (C.c_int, [_P, _P]),
    "ua2_wavlm_frames": (C.c_longlong, [_P, C.c_longlong]),
    "ua2_wavlm_forward": (C.c_int, [_P, _P, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "ua2_wavlm_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ua2_wavlm_last_launch_count": (C.c_int, [_P]),
    "ua2_wavlm_rel_bucket_table": (C.c_int, [C.c_int, C.c_int, C.c_int, _P]),
    "ua2_wavlm_ops_f32": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ua2_thinking_create": (C.c_int, [C.POINTER(ThinkingCfg), C.POINTER(_P)]),
    "ua2_thinking_destroy": (C.c_int, [_P]),
    "ua2_thinking_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_thinking_finalize": (C.c_int, [_P, _P]),
    "ua2_thinking_rows": (C.c_longlong, [_P, C.c_int, C.c_int]),
    "ua2_thinking_encode": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ua2_thinking_last_launch_count": (C.c_int, [_P]),
    "ua2_stx_create": (C.c_int, [C.POINTER(StxCfg), C.POINTER(_P)]),
    "ua2_stx_destroy": (C.c_int, [_P]),
    "ua2_stx_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "ua2_stx_finalize": (C.c_int, [_P]),
    "ua2_stx_start_streaming": (C.c_int, [_P, C.c_int, _P]),
    "ua2_stx_stop_streaming": (C.c_int, [_P]),
    "ua2_stx_reset_streaming": (C.c_int, [_P]),
    "ua2_stx_forward": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    "ua2_stx_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ua2_stx_last_launch_count": (C.c_int, [_P]),
    "ua2_stx_get_kv": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
}


def lib():
    """Load (once) and return the ctypes library with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ua2Error(f"{LIB_PATH} not found - build it with `python -m uniaudio2_b200.build` "
                           "(there is no CPU fallback for this path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        # UA2_OPTIONS="attn_ring=1,conv_tc=1": global options applied at load, for A/B runs of whole programs (bench.py, the test
        # suite) without editing them; an unknown name or a refused value raises - and leaves the module unloaded, so that a caller
        # that swallows the exception cannot go on with half of the options applied
        for item in filter(None, (x.strip() for x in os.environ.get("UA2_OPTIONS", "").split(","))):
            name, _, value = item.partition("=")
            rc = l.ua2_set_global_option(name.strip().encode(), int(value))
            if rc != 0:
                raise ValueError(f"UA2_OPTIONS {item}: " + l.ua2_last_error().decode("utf-8", "replace"))
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().ua2_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}" if what else msg)  # mirrors the reference's ValueError cases
        if rc == -3:
            raise TypeError(f"{what}: {msg}" if what else msg)  # lit_model.py:134-135 'You need to call set_kv_cache'
        raise Ua2Error(f"{what}: {msg}" if what else msg)


def ptr(t):
    """Device (or host) address of a torch tensor as c_void_p; None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
