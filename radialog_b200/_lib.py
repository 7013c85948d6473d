"""ctypes binding of ``libradialog_b200.so`` (C-ABI in ``include/radialog_b200.h``).

There is no CPU / PyTorch fallback: if the shared library is missing the import of any compute entry point raises,
and every call checks the status code and raises ``RuntimeError`` with ``rd_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libradialog_b200.so")

RD_F16, RD_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SWIGLU = 0, 1, 2, 3
ALGO_AUTO, ALGO_GEMV, ALGO_TC, ALGO_SIMT = 0, 1, 2, 3

# weight slots (include/radialog_b200.h)
W_EMBED, W_FINAL_NORM, W_LM_HEAD, W_IMG_PROJ_W, W_IMG_PROJ_B, W_ROPE_COS, W_ROPE_SIN = 0, 1, 2, 3, 4, 5, 6
W_QKV, W_O, W_GATE_UP, W_DOWN, W_LN1, W_LN2, W_LORA_A, W_LORA_B = 10, 11, 12, 13, 14, 15, 16, 17

PROFILE_CLASSES = ("rmsnorm", "qkv", "rope", "attn", "o", "gate_up", "down", "lm_head", "argmax", "embed")


class Epilogue(C.Structure):
    _fields_ = [("bias_dev", C.c_void_p), ("residual_dev", C.c_void_p), ("ld_res", C.c_int64), ("res_mode", C.c_int),
                ("act", C.c_int), ("lora_t_dev", C.c_void_p), ("lora_b_dev", C.c_void_p), ("lora_r", C.c_int),
                ("lora_scale", C.c_float)]


class LlmConfig(C.Structure):
    _fields_ = [("vocab", C.c_int), ("hidden", C.c_int), ("inter", C.c_int), ("layers", C.c_int), ("heads", C.c_int),
                ("max_pos", C.c_int), ("rms_eps", C.c_float), ("dtype", C.c_int), ("lora_r", C.c_int),
                ("lora_scale", C.c_float), ("qformer_hidden", C.c_int), ("max_batch", C.c_int), ("max_ctx", C.c_int),
                ("pad_id", C.c_int), ("eos_id", C.c_int), ("img_id", C.c_int)]


class VisionConfig(C.Structure):
    _fields_ = [("image_size", C.c_int), ("layers", C.c_int * 4), ("width", C.c_int), ("backbone_to_vit", C.c_int),
                ("joint", C.c_int), ("num_query", C.c_int), ("q_hidden", C.c_int), ("q_heads", C.c_int),
                ("q_layers", C.c_int), ("q_inter", C.c_int), ("cross_freq", C.c_int), ("ln_vision_eps", C.c_float),
                ("q_ln_eps", C.c_float), ("dtype", C.c_int), ("max_batch", C.c_int), ("pooler_blocks", C.c_int), ("pooler_heads", C.c_int),
                ("pooler_hidden", C.c_int), ("pooler_ln_eps", C.c_float)]


_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol declared in include/radialog_b200.h must appear here
# (tests/test_abi.py checks the header against the loaded library).
SIGNATURES = {
    "rd_last_error": (C.c_char_p, []),
    "rd_version": (_i, []),
    "rd_device_ok": (_i, [_i]),
    "rd_set_pdl": (_i, [_i]),
    "rd_linear": (_i, [_p, _i64, _p, _i64, _p, _i64, _i, _i, _i, C.POINTER(Epilogue), _i, _i, _p, _i64, _p]),
    "rd_linear_workspace_bytes": (_i64, [_i, _i, _i]),
    "rd_linear_force_splits": (_i, [_i]),
    "rd_linear_splitk_mode": (_i, [_i]),
    "rd_linear_tmem_staging": (_i, [_i]),
    "rd_linear_wide_persistent": (_i, [_i]),
    "rd_linear_wide_min_tiles": (_i, [_i]),
    "rd_linear_wide_force_nt": (_i, [_i]),
    "rd_linear_wide_force_stages": (_i, [_i]),
    "rd_linear_wide_pair": (_i, [_i]),
    "rd_conv_nhwc_implicit": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(Epilogue), _i, _p]),
    "rd_conv_set_implicit": (_i, [_i]),
    "rd_im2col_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "rd_rmsnorm": (_i, [_p, _p, _p, _i, _i, _f, _p, _i, _p, _i, _p]),
    "rd_layernorm": (_i, [_p, _p, _p, _p, _i, _i, _f, _i, _p]),
    "rd_rope_kv_store": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _f, _i, _p]),
    "rd_attention": (_i, [_p, _i64, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "rd_attention_set_tensor_core": (_i, [_i]),
    "rd_attention_decode": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _f, _i, _p]),
    "rd_attention_decode_partials": (_i, [_p, _i, C.c_longlong, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _f, _i, _p]),
    "rd_embed_splice": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "rd_llm_create": (_i, [C.POINTER(LlmConfig), C.POINTER(_p)]),
    "rd_llm_destroy": (None, [_p]),
    "rd_llm_set_weight": (_i, [_p, _i, _i, _p]),
    "rd_llm_set_algo": (_i, [_p, _i]),
    "rd_llm_set_qkv_partials": (_i, [_p, _i]),
    "rd_llm_set_od_partials": (_i, [_p, _i]),
    "rd_llm_set_l2_prefetch": (_i, [_p, C.c_longlong, C.c_longlong, C.c_longlong]),
    "rd_llm_prefill": (_i, [_p, _p, _p, _i, _i, _p, _i, _p]),
    "rd_llm_truncate": (_i, [_p, _i, _p, _p]),
    "rd_llm_extend": (_i, [_p, _p, _i, _i, _i, _p]),
    "rd_llm_decode_step": (_i, [_p, _p]),
    "rd_llm_note_replayed_steps": (_i, [_p, _i]),
    "rd_llm_state": (_i, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_i)]),
    "rd_llm_reorder_cache": (_i, [_p, _p, _p]),
    "rd_llm_force_tokens": (_i, [_p, _p, _p]),
    "rd_llm_done_flag": (_i, [_p, C.POINTER(_p)]),
    "rd_llm_profile": (_i, [_p, _i]),
    "rd_llm_profile_read": (_i, [_p, C.POINTER(_f), C.POINTER(_i), _i]),
    "rd_llm_launch_count": (_i64, [_p]),
    "rd_vision_create": (_i, [C.POINTER(VisionConfig), C.POINTER(_p)]),
    "rd_vision_destroy": (None, [_p]),
    "rd_vision_set_weight": (_i, [_p, C.c_char_p, _p]),
    "rd_vision_forward": (_i, [_p, _p, _i, _p, _p, _p]),
    "rd_vision_forward_temporal": (_i, [_p, _p, _p, _i, _p, _p, _p]),
    "rd_vision_launch_count": (_i64, [_p]),
    "rd_preproc_create": (_i, [_i, _i, _i, _i, C.POINTER(_p)]),
    "rd_preproc_destroy": (None, [_p]),
    "rd_preproc_run": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "rd_preproc_launch_count": (_i64, [_p]),
}

_lib = None


def lib_path() -> str:
    return LIB_PATH


def load(build_if_missing: bool = True):
    """Loads the shared library (building it with nvcc first if it is absent).  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        # build() is a no-op when lib/build.stamp matches the digest of csrc/ + the header: a stale .so (sources edited after
        # the last build) is rebuilt instead of being loaded silently
        from . import build as _build
        _build.build()
    elif not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found; run `python -m radialog_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().rd_last_error()
        raise RuntimeError(f"libradialog_b200 {what} failed ({status}): {msg.decode() if msg else '?'}")


def dtype_code(torch_dtype) -> int:
    import torch
    if torch_dtype == torch.float16:
        return RD_F16
    if torch_dtype == torch.bfloat16:
        return RD_BF16
    raise ValueError(f"unsupported dtype {torch_dtype}: the B200 path computes in float16 or bfloat16")


def ptr(t) -> int:
    """Device pointer of a CUDA tensor (or 0 for None).  Refuses CPU tensors: there is no CPU path."""
    if t is None:
        return 0
    if not t.is_cuda:
        raise RuntimeError("radialog_b200: expected a CUDA tensor (the product path has no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("radialog_b200: expected a contiguous tensor")
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
