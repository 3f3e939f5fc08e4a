#!/usr/bin/env python
"""One launch of our GEMM and one of cuBLAS (F.linear) per shape, for an ncu capture that compares the two kernels
(tile / cluster shape in the cuBLAS kernel name, tensor-pipe activity, L2 and DRAM bytes):
  ncu --set full --clock-control none -o gpurun_out/gemm_cmp python tools/gemm_vs_cublas.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmpl_b200 import ops  # noqa: E402

shapes = [(4680, 1536, 8960), (4680, 1536, 1536), (10920, 5120, 5120)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):  # the second launch of each is the warm one
        ops.linear(x, w, b, out=out, tile_n=512)
        torch.nn.functional.linear(x, w, b)
    torch.cuda.synchronize()
print("done")
