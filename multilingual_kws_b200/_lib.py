"""ctypes binding of libkws_b200.so (the C ABI declared in include/kws_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
this module raises at import-of-symbol time with the build command; if no CUDA device is present
every compute entry returns KWS_ERR_CUDA and is surfaced as a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KWS_LIB_PATH") or os.path.join(_HERE, "csrc", "libkws_b200.so")   # override: A/B builds

c_void_p, c_int, c_float, c_size_t, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_int64

# name -> (restype, argtypes); kept in one place so tests can check the export list against the header.
SIGNATURES = {
    "kws_last_error": (ctypes.c_char_p, []),
    "kws_abi_version": (c_int, []),
    "kws_host_alloc": (c_int, [ctypes.POINTER(c_void_p), c_size_t, c_int]),
    "kws_host_free": (c_int, [c_void_p]),
    "kws_frontend_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                    c_float, c_float, c_float, c_int, c_float, c_float, c_int, c_int, c_int]),
    "kws_frontend_destroy": (None, [c_void_p]),
    "kws_frontend_num_frames": (c_int, [c_void_p, c_int]),
    "kws_frontend_tables": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_size_t)]),
    "kws_frontend_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "kws_frontend_stream_num_windows": (c_int64, [c_void_p, c_int64, c_int, c_int]),
    "kws_frontend_stream_scratch_bytes": (c_size_t, [c_void_p, c_int64]),
    "kws_embed_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_int]),
    "kws_embed_destroy": (None, [c_void_p]),
    "kws_embed_info": (c_int, [c_void_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                               ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double)]),
    "kws_embed_op_name": (c_int, [c_void_p, c_int, ctypes.c_char_p, c_size_t, ctypes.POINTER(c_int64)]),
    "kws_embed_op_info": (c_int, [c_void_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                  ctypes.POINTER(c_int)]),
    "kws_embed_forward_timed": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "kws_embed_launches": (c_int, [c_void_p, c_int]),
    "kws_embed_set_graph": (c_int, [c_void_p, c_int]),
    "kws_embed_set_fuse": (c_int, [c_void_p, c_int]),
    "kws_embed_set_chunk_late": (c_int, [c_void_p, c_int]),
    "kws_embed_set_chunk": (c_int, [c_void_p, c_int]),
    "kws_embed_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "kws_embed_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kws_embed_forward_budget": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "kws_embed_forward_until": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "kws_embed_forward_tap": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "kws_gemm_h16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                             c_int, c_int, c_void_p]),
    "kws_head_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_float, c_float, c_float]),
    "kws_head_destroy": (None, [c_void_p]),
    "kws_head_flat_size": (c_size_t, [c_void_p]),
    "kws_head_num_params": (c_int, [c_void_p]),
    "kws_head_step_count": (ctypes.c_longlong, [c_void_p]),
    "kws_head_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "kws_head_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "kws_head_apply_adam": (c_int, [c_void_p, c_void_p, c_float, c_void_p]),
    "kws_head_get_params": (c_int, [c_void_p, c_void_p]),
    "kws_head_reset_optimizer": (c_int, [c_void_p]),
    "kws_head_input_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "kws_train_transpose_h16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "kws_train_swish_fwd": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "kws_train_gap_swish_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "kws_train_gap_swish_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "kws_train_dw_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "kws_train_dw_bwd_scratch_floats": (c_size_t, [c_int, c_int, c_int]),
    "kws_train_dw_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kws_train_gate_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "kws_train_gate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kws_train_act_bwd": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_float, c_void_p, c_void_p]),
    "kws_train_colsum": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "kws_train_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_float, c_float,
                               ctypes.c_longlong, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "kws_train_lr_step": (c_int, [c_void_p, c_float, c_float, c_float, c_void_p, c_void_p]),
    "kws_head_apply_adam_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "kws_head_advance_step_count": (c_int, [c_void_p, ctypes.c_longlong]),
    "kws_stream_detect": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, ctypes.c_double, ctypes.c_double, c_int, c_void_p,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "kws_augment_pcm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                c_void_p]),
    "kws_spec_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "kws_frontend_stream": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_int64, c_float, c_void_p,
                                    c_void_p, c_int, c_void_p]),
}

_lib = None


class KwsError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with: python -c 'import __graft_entry__ as g; g.build()' "
                "(nvcc, sm_100a). multilingual_kws_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().kws_last_error().decode("utf-8", "replace")
        raise KwsError(f"{what or 'libkws_b200'} failed (code {rc}): {msg}")


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)
