#!/usr/bin/env python
"""Runs a few backbone forwards of the cfg2 workload (Wan-1.3B dims, 3 frames at 60x104, KV length (chunk+1)*4680)
for ncu captures:  ncu ... python tools/profile_forward.py --chunk 3 --forwards 2"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmpl_b200.causal_model import CausalWanModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chunk", type=int, default=3)
ap.add_argument("--forwards", type=int, default=2)
ap.add_argument("--layers", type=int, default=30)
ap.add_argument("--model", default="1.3B", choices=["1.3B", "14B"])
a = ap.parse_args()
dims = dict(dim=1536, ffn_dim=8960, num_heads=12) if a.model == "1.3B" else dict(dim=5120, ffn_dim=13824, num_heads=40)
dev = "cuda:0"
torch.manual_seed(0)
with torch.device(dev):
    model = CausalWanModel(num_layers=a.layers, **dims)
model = model.to(torch.bfloat16).eval()
H = dims["num_heads"]
rows = 32760
kv = [{"k": torch.randn(1, rows, H, 128, device=dev, dtype=torch.bfloat16), "v": torch.randn(1, rows, H, 128, device=dev, dtype=torch.bfloat16),
       "global_end_index": torch.tensor([a.chunk * 4680], device=dev), "local_end_index": torch.tensor([a.chunk * 4680], device=dev)}
      for _ in range(a.layers)]
cross = [{"k": None, "v": None, "is_init": False} for _ in range(a.layers)]
x = torch.randn(1, 16, 3, 60, 104, device=dev, dtype=torch.bfloat16)
ctx = torch.randn(1, 512, 4096, device=dev, dtype=torch.bfloat16)
t = torch.full((1, 3), 937.5, device=dev)
for i in range(a.forwards):
    if i == a.forwards - 1:  # ncu --profile-from-start off captures only the last forward
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    model(x, t=t, context=ctx, seq_len=32760, kv_cache=kv, crossattn_cache=cross, current_start=a.chunk * 4680)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", model.launch_count())
