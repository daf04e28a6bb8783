"""ctypes binding of the C-ABI library (include/deco_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DECO_B200_LIB") or os.path.join(_HERE, "_C", "libdeco_b200.so")   # override: instrumented builds

_vp, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float

# name -> (restype, argtypes); mirrors include/deco_b200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "deco_last_error": (C.c_char_p, []),
    "deco_abi_version": (_i, []),
    "deco_gemm_bf16": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _ll, _vp, _ll, _i, _i, _vp]),
    "deco_gemm_tile_plan": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "deco_gemm_bf16_tn": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _vp]),
    "deco_gemm_bf16_tn_deint16": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _vp]),
    "deco_gemm_bf16_f32_splitk": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp]),
    "deco_gemm_stream_parts": (_i, [_i, _i]),
    "deco_gemm_stream": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _i, _vp, _vp, _ll, _vp, _ll,
                              _vp, _vp]),
    "deco_gemm_norm_qkv": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _i, _i, _f, _vp, _ll, _i, _i, _vp, _vp, _vp,
                                _i, _vp, _i, _f, _i, _vp]),
    "deco_gemm_norm_swiglu": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _i, _i, _f, _vp, _ll, _vp]),
    "deco_gemm_set_tuning": (_i, [_i, _i]),
    "deco_gemm_reserve_sms": (_i, [_i]),
    "deco_patchify": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "deco_timestep_freq": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "deco_cond_combine": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "deco_rmsnorm_modulate": (_i, [_vp, _i, _vp, _vp, _vp, _ll, _i, _vp, _ll, _i, _f, _vp]),
    "deco_qknorm_rope": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _vp]),
    "deco_headnorm_rope": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _vp]),
    "deco_headnorm_rope_to": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _vp]),
    "deco_rmsnorm_addpos": (_i, [_vp, _vp, _vp, _i, _vp, _ll, _i, _f, _vp]),
    "deco_cast_f32_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "deco_attention_fwd": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _f, _vp]),
    "deco_attention_fwd_pitched": (_i, [_vp, _ll, _i, _vp, _vp, _ll, _i, _i, _vp, _vp, _ll, _i, _i, _vp, _ll, _i, _i, _i, _i, _f, _vp]),
    "deco_silu_add_rows": (_i, [_vp, _i, _vp, _vp, _ll, _i, _i, _vp]),
    "deco_decoder_blob_bytes": (_i, [_i]),
    "deco_pixel_decoder": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "deco_cfg_step": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _ll, _vp]),
    "deco_fp2uint8": (_i, [_vp, _vp, _ll, _vp]),
    "deco_sampler_advance": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp]),
    "deco_cfg_step_dev": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp]),
    "deco_cfg_step_ex": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _vp,
                              _vp, _vp, _vp, _vp, _ll, _vp]),
    "deco_heun_sde_step": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _f, _f, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _vp,
                                _ll, _vp]),
    "deco_layernorm_modulate": (_i, [_vp, _vp, _vp, _ll, _i, _vp, _ll, _i, _f, _vp]),
    "deco_unpatchify": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "deco_center_rows": (_i, [_vp, _vp, _ll, _i, _vp]),
    "deco_opt_chunk_elems": (_i, []),
    "deco_adamw_ema_step": (_i, [_vp, _vp, _i, _f, _f, _f, _f, _f, _f, _f, _f, _vp]),
    "deco_dct_scratch_doubles": (_i, []),
    "deco_dct_fm_loss": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "deco_transpose_cast": (_i, [_vp, _i, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "deco_colsum": (_i, [_vp, _i, _ll, _vp, _ll, _i, _vp]),
    "deco_gate_residual": (_i, [_vp, _vp, _vp, _ll, _vp, _i, _ll, _i, _vp]),
    "deco_gate_residual_norm": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _i, _vp, _ll, _i, _f, _vp]),
    "deco_gate_bwd": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _i, _ll, _i, _vp]),
    "deco_silu_add_rows_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _vp]),
    "deco_swiglu_fwd": (_i, [_vp, _vp, _ll, _i, _vp]),
    "deco_swiglu_bwd": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "deco_rmsnorm_modulate_bwd": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _ll, _i, _f, _vp]),
    "deco_rmsnorm_modulate_bwd_gate": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _ll, _i, _f,
                                            _vp, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp]),
    "deco_headnorm_rope_bwd": (_i, [_vp, _ll, _vp, _ll, _i, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _vp]),
    "deco_cond_combine_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "deco_silu_bwd": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "deco_attention_fwd_lse": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _vp, _ll, _vp, _i, _i, _i, _i, _f, _vp]),
    "deco_attention_bwd": (_i, [_vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp,
                                _i, _i, _i, _i, _i, _i, _f, _vp]),
    "deco_decoder_tc_blob_bytes": (_i, [_i]),
    "deco_pixel_decoder_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp,
                                   _vp]),
    "deco_nerf_decoder_blob_bytes": (_i, [_i]),
    "deco_nerf_decoder": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "deco_train_timesteps": (_i, [_vp, _vp, _vp, _f, _i, _vp, _vp, _i, _vp]),
    "deco_flow_pair": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _vp]),
    "deco_label_dropout": (_i, [_vp, _vp, _vp, _f, _vp, _i, _vp]),
    "deco_attention_bwd_tc": (_i, [_vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp,
                                   _i, _i, _i, _i, _i, _f, _vp]),
    "deco_decoder_train_blob_floats": (_i, [_i]),
    "deco_pixel_decoder_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "deco_decoder_bwd_blob_bytes": (_i, [_i]),
    "deco_pixel_decoder_bwd_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None
_lock = threading.Lock()
launch_count = 0  # number of kernel-launching C-ABI calls made through `call` (bench.py reports it)


class DecoLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libdeco_b200.so (building nothing: use `python -m deco_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DecoLibraryError(
                f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m deco_b200.build`). "
                "deco_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on a non-zero status."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.deco_last_error()
        raise DecoLibraryError(f"{name} failed with status {rc}: {msg.decode() if msg else ''}")
    launch_count += 1


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_of(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
