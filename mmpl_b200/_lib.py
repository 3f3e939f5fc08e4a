"""ctypes binding of the C ABI declared in include/mmpl_b200.h.

`load()` returns the library with argtypes set for every exported symbol; `check(status)` turns a
negative status into a Python exception carrying mmpl_last_error(). Nothing here computes anything:
if the shared library is missing (and cannot be built) loading raises, there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import _build

_LIB = None
ABI_VERSION = 4  # == MMPL_ABI_VERSION in include/mmpl_b200.h (tests/test_abi.py checks the header)

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float
c_int_p = C.POINTER(C.c_int)
c_void_pp = C.POINTER(C.c_void_p)


class ModelConfig(C.Structure):
    _fields_ = [
        ("dim", c_int), ("ffn_dim", c_int), ("num_heads", c_int), ("num_layers", c_int),
        ("freq_dim", c_int), ("text_dim", c_int), ("text_len", c_int), ("in_dim", c_int), ("out_dim", c_int),
        ("eps", c_float), ("max_tokens", c_int),
    ]


class UniPCCoeffs(C.Structure):
    """mmpl_unipc_coeffs."""
    _fields_ = [
        ("guidance", c_float), ("sigma", c_float), ("corr_order", c_int),
        ("corr_a", c_float), ("corr_b", c_float), ("corr_c", c_float), ("corr_rk", c_float),
        ("corr_rho0", c_float), ("corr_rho1", c_float), ("pred_order", c_int),
        ("pred_a", c_float), ("pred_b", c_float), ("pred_c", c_float), ("pred_rk", c_float), ("true_division", c_int),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("latents", c_void_p), ("lat_stride_f", c_int64), ("lat_stride_c", c_int64),
        ("n_frames", c_int), ("lat_h", c_int), ("lat_w", c_int),
        ("timesteps", c_void_p), ("context", c_void_p),
        ("kv_k", c_void_pp), ("kv_v", c_void_pp), ("cache_rows", c_int64),
        ("frame_pos", c_int_p), ("kv_row", c_int_p), ("kv_to_tail", c_int),
        ("n_seg", c_int), ("seg_start", c_int_p), ("seg_rows", c_int_p),
        ("cross_k", c_void_pp), ("cross_v", c_void_pp), ("cross_init", c_int),
        ("flow", c_void_p), ("x0", c_void_p), ("sigma", c_void_p),
    ]


# name -> (restype, argtypes); must list every function declared in include/mmpl_b200.h
SIGNATURES = {
    "mmpl_abi_version": (c_int, []),
    "mmpl_build_id": (C.c_char_p, []),
    "mmpl_last_error": (C.c_char_p, []),
    "mmpl_unipc_cfg_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, C.POINTER(UniPCCoeffs), c_void_p]),
    "mmpl_gemm_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                               c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "mmpl_conv3d_cl": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_void_p]),
    "mmpl_vae_norm_act": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "mmpl_vae_upsample2x": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmpl_vae_pick_odd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmpl_softmax_rows": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p]),
    "mmpl_anchor_broadcast": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "mmpl_flash_attn": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                c_int64, c_int, c_int, c_int_p, c_int_p, c_int_p, c_void_p, c_int64, c_float, c_void_p]),
    "mmpl_gemm_set_streamk": (c_int, [c_int]),
    "mmpl_attn_set_split": (c_int, [c_int]),
    "mmpl_attn_set_ctas": (c_int, [c_int]),
    "mmpl_attn_plan": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int]),
    "mmpl_ln_modulate": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p,
                                 c_int64, c_int, c_void_p]),
    "mmpl_ln_affine": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "mmpl_rmsnorm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_float, c_void_p]),
    "mmpl_qk_norm_rope_kv": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int_p,
                                     c_int_p, c_float, c_void_p]),
    "mmpl_modulation_add": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mmpl_sinusoid_embedding": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mmpl_skinny_linear": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p]),
    "mmpl_patchify": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmpl_unpatchify_x0": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_int, c_void_p]),
    "mmpl_add_noise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "mmpl_ctx_create": (c_int, [C.POINTER(ModelConfig), C.POINTER(c_void_p)]),
    "mmpl_ctx_destroy": (None, [c_void_p]),
    "mmpl_bind_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int64]),
    "mmpl_bind_rope_table": (c_int, [c_void_p, c_void_p]),
    "mmpl_launch_count": (c_int64, [c_void_p, c_int]),
    "mmpl_forward": (c_int, [c_void_p, C.POINTER(ForwardArgs), c_void_p]),
    "mmpl_total_launches": (c_int64, [c_int]),
    "mmpl_profile_enable": (c_int, [c_void_p, c_int]),
    "mmpl_profile_mask": (c_int, [c_void_p]),
    "mmpl_launch_credit": (c_int64, [c_void_p, c_int64]),
    "mmpl_workspace_generation": (c_int64, []),
    "mmpl_profile_read": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_int64), c_int]),
    "mmpl_profile_read_sites": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_int64), c_int]),
}


class MmplError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"mmpl_b200 error {status}: {message}")
        self.status = status


def lib_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load libmmpl_b200.so, building it with nvcc first if it does not match the sources in the tree (content hash,
    mmpl_b200/_build.py). The build is serialised across processes; a failed build raises, and a library whose compiled-in
    build id or ABI version differs from this tree is refused rather than loaded."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    override = os.environ.get("MMPL_B200_LIB")  # A/B runs of tools/bench_kernels.py against another build
    if override:
        path, build_if_missing = Path(override), False
    if build_if_missing and _build.is_stale():
        _build.build_library()
    if not path.exists():
        raise ImportError(f"{path} is missing: build it with `python -m mmpl_b200._build` (needs nvcc); "
                          "mmpl_b200 has no CPU fallback")
    lib = C.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{path} does not export {name}: the library is older than include/mmpl_b200.h, rebuild it "
                              "(python -m mmpl_b200._build --force)") from e
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mmpl_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.mmpl_abi_version()}, this binding expects {ABI_VERSION}")
    if not override:
        built, want = lib.mmpl_build_id().decode(), _build.source_id()
        if built != want:
            raise ImportError(f"{path} was built from other sources (build id {built}, tree {want}): rebuild it with "
                              "`python -m mmpl_b200._build --force`")
    _LIB = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().mmpl_last_error()
        raise MmplError(status, msg.decode() if msg else "")
