"""ORACLE tooling — the CPU timing arm. Runs the oracle restatement of the reference pipeline with the same
native bf16 torch CPU operators the reference's own CPU path executes (nn.Linear on bf16 tensors -> oneDNN bf16 GEMM,
F.scaled_dot_product_attention, wan/modules/attention.py:170-185), instead of the oracle's fp32-accumulating
restatements, so that the reported CPU baseline reflects the reference's cost on the host cores. Only bench.py's
cpu_baseline / --impl reference leg uses this ("kind": "port": /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import causal_wan_oracle as O


def _native_linear(x, w, b):
    return F.linear(x, w, b)


def _native_attention(q, k, v):
    # [L, H, hd] -> [1, H, L, hd], as attention() does on the SDPA branch
    o = F.scaled_dot_product_attention(q.transpose(0, 1).unsqueeze(0), k.transpose(0, 1).unsqueeze(0),
                                       v.transpose(0, 1).unsqueeze(0))
    return o[0].transpose(0, 1).contiguous()


@contextlib.contextmanager
def native_ops():
    saved = (O.linear, O.attention)
    O.linear, O.attention = _native_linear, _native_attention
    try:
        yield
    finally:
        O.linear, O.attention = saved


@torch.no_grad()
def causal_inference_native(cfg, weights, noise, prompt, **kw):
    """noise [F,C,H,W] bf16, prompt [text_len, text_dim] bf16 -> latents; same schedule as O.causal_inference."""
    fs = (noise.shape[2] // 2) * (noise.shape[3] // 2)
    with native_ops():
        return O.causal_inference(cfg, weights, noise, prompt, cache_rows=noise.shape[0] * fs, **kw)
