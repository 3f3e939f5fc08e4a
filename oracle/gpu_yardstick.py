"""ORACLE tooling (never imported by the product, by bench.py or by tests that gate anything): the same-box GPU yardstick
SURVEY.md §2.2 asks for - what the REFERENCE'S OWN STACK delivers on a B200 for the benchmarked workload. The reference's
GPU path is eager PyTorch: nn.Linear -> cuBLASLt, flash_attn_varlen_func of the flash-attn 2 wheel (mma.sync kernels, no
sm_100 tensor-core path; wan/modules/attention.py:104-133), and ~25 ATen element-wise launches per block around them. This
script runs the oracle's restatement of the pipeline (same schedule, same operators at the same rounding points) with exactly
those library calls on the device, and times the cfg2 rollout with CUDA events:

    python -m oracle.gpu_yardstick [--workload cfg2] [--steps 2] [--attention flash_attn|sdpa]

It prints one JSON line. /root/reference does not exist on the GPU box, so this is the closest thing to "the unmodified
reference on this GPU" that can run there; results are committed under profiles/ next to the bench lines they sit beside.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import causal_wan_oracle as O  # noqa: E402

WORKLOADS = {
    "cfg2": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30), 21, 60, 104),
    "cfg1": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30), 3, 30, 52),
    "tiny": (dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32), 6, 8, 12),
}


def _attention_fn(kind: str):
    if kind == "flash_attn":
        from flash_attn import flash_attn_func   # the wheel the reference calls (varlen form with one sequence = this call)

        def attn(q, k, v):
            return flash_attn_func(q[None], k[None], v[None])[0]
        return attn, "flash-attn 2 wheel (flash_attn_func)"

    def attn(q, k, v):   # [L, H, hd] -> [1, H, L, hd]: torch picks cuDNN / flash / mem-efficient
        o = F.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1)[None], v.transpose(0, 1)[None])
        return o[0].transpose(0, 1).contiguous()
    return attn, "torch F.scaled_dot_product_attention"


@contextlib.contextmanager
def library_ops(attn):
    saved = (O.linear, O.attention)
    O.linear, O.attention = (lambda x, w, b: F.linear(x, w, b)), attn
    try:
        yield
    finally:
        O.linear, O.attention = saved


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--attention", default="flash_attn", choices=["flash_attn", "sdpa"])
    a = ap.parse_args()
    assert torch.cuda.is_available(), "needs a GPU"
    dev = "cuda:0"
    dims, frames, lh, lw = WORKLOADS[a.workload]
    cfg = O.WanConfig(**dims)
    w = {k: v.to(dev) for k, v in O.make_weights(cfg, seed=0).items()}
    noise = torch.randn(frames, 16, lh, lw, generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).to(dev)
    prompt = torch.randn(cfg.text_len, cfg.text_dim, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    try:
        attn, attn_name = _attention_fn(a.attention)
    except Exception as e:  # noqa: BLE001
        attn, attn_name = _attention_fn("sdpa")
        attn_name += f" (flash-attn wheel unavailable: {type(e).__name__})"
    fs = (lh // 2) * (lw // 2)
    times = []
    with library_ops(attn), torch.no_grad():
        for i in range(a.warmup + a.steps):
            torch.manual_seed(1234)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = O.causal_inference(cfg, w, noise, prompt, cache_rows=frames * fs)
            e1.record()
            torch.cuda.synchronize()
            if i >= a.warmup:
                times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    print(json.dumps({
        "what": "same-box yardstick: the reference's stack (eager PyTorch ops, cuBLASLt Linear, " + attn_name + ") running the "
                "oracle's restatement of CausalInferencePipeline.inference on one B200",
        "workload": a.workload, "metric": "denoised_latent_frames_per_s", "value": frames / (ms / 1e3), "ms_per_step": ms,
        "steps": len(times), "warmup": a.warmup, "finite": bool(torch.isfinite(out.float()).all()),
        "note": "cache allocation (zeros) is inside the timed call, as in the reference's first inference(); RoPE in complex128 "
                "eager ops like the reference"}))


if __name__ == "__main__":
    main()
