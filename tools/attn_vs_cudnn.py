#!/usr/bin/env python
"""One launch of our attention kernel and one of cuDNN's fused attention (torch SDPA, cuDNN backend) per shape, for an
ncu capture that compares the two:  ncu --set full --clock-control none -o gpurun_out/attn_cmp python tools/attn_vs_cudnn.py"""
import os
import sys

import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmpl_b200 import ops  # noqa: E402

shapes = [(4680, 4680, 12), (4680, 32760, 12), (4680, 512, 12)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (Lq, Lk, H) in shapes:
    q = torch.randn(Lq, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(Lk, H, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(Lk, H, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(q)
    q4, k4, v4 = (t.transpose(0, 1)[None] for t in (q, k, v))
    for _ in range(2):
        ops.flash_attn(q, k, v, out=out)
        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            torch.nn.functional.scaled_dot_product_attention(q4, k4, v4)
    torch.cuda.synchronize()
print("done")
