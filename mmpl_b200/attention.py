"""`flash_attention` / `attention` with the reference call-site signature (wan/modules/attention.py:32-185),
backed by the tcgen05 flash-attention kernel (csrc/attention_tcgen05.cu) instead of the flash-attn wheel.

Supported: what the MMPL hot path uses — non-causal, no dropout, no window, head_dim 128, q/k/v [B, L, N, 128]
with optional per-sample q_lens / k_lens. Anything else raises; there is no SDPA or CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["flash_attention", "attention"]

FLASH_ATTN_2_AVAILABLE = False  # the flash-attn wheel is not used
FLASH_ATTN_3_AVAILABLE = False


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                    window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """q: [B, Lq, Nq, C]; k: [B, Lk, Nk, C]; v: [B, Lk, Nk, C]. Returns [B, Lq, Nq, C] in q's dtype.
    Rows of q beyond q_lens[b] are returned as zeros (the reference drops them by packing)."""
    half_dtypes = (torch.float16, torch.bfloat16)
    assert dtype in half_dtypes
    assert q.device.type == "cuda" and q.size(-1) <= 256
    if causal or dropout_p != 0. or tuple(window_size) != (-1, -1):
        raise NotImplementedError("mmpl_b200.flash_attention: causal / dropout / sliding-window are not on the MMPL "
                                  "hot path (block causality comes from which rows are in the KV cache)")
    if q.size(-1) != 128 or k.size(2) != q.size(2):
        raise NotImplementedError("mmpl_b200.flash_attention: head_dim must be 128 and Nq == Nk")
    b, lq, lk, out_dtype = q.size(0), q.size(1), k.size(1), q.dtype

    def bf16(x):
        return x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)

    q, k, v = bf16(q), bf16(k), bf16(v)
    if q_scale is not None:
        q = q * q_scale
    out = torch.zeros_like(q) if q_lens is not None else torch.empty_like(q)
    for i in range(b):
        nq = int(q_lens[i]) if q_lens is not None else lq
        nk = int(k_lens[i]) if k_lens is not None else lk
        qi = q[i, :nq]
        qi = qi if qi.stride(-1) == 1 and qi.stride(1) == 128 else qi.contiguous()
        ki = k[i] if k[i].stride(-1) == 1 and k[i].stride(1) == 128 else k[i].contiguous()
        vi = v[i] if v[i].stride(-1) == 1 and v[i].stride(1) == 128 and v[i].stride(0) == ki.stride(0) else v[i].contiguous()
        if vi.stride(0) != ki.stride(0):
            ki = ki.contiguous()
        oi = out[i, :nq]
        if oi.is_contiguous():
            ops.flash_attn(qi, ki, vi, segments=[(0, nk)], softmax_scale=softmax_scale, out=oi)
        else:
            oi.copy_(ops.flash_attn(qi, ki, vi, segments=[(0, nk)], softmax_scale=softmax_scale))
    return out.type(out_dtype)


def attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
              window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, fa_version=None):
    """wan/modules/attention.py:139-185 — always the native kernel here."""
    return flash_attention(q=q, k=k, v=v, q_lens=q_lens, k_lens=k_lens, dropout_p=dropout_p,
                           softmax_scale=softmax_scale, q_scale=q_scale, causal=causal, window_size=window_size,
                           deterministic=deterministic, dtype=dtype, version=fa_version)
