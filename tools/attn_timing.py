#!/usr/bin/env python
"""Per-phase clock64() breakdown of flash_attn_kernel (development tool).

Needs a library built with -DMMPL_ATTN_TIMING=1 (exports mmpl_attn_debug_read), selected with MMPL_B200_LIB.
Prints, for CTA 0: per KV tile, the cycles each softmax warp spends waiting for S, computing P and handing P over,
and the cycles the MMA warp spends waiting for P / K / V."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmpl_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
raw = C.CDLL(os.environ["MMPL_B200_LIB"])
raw.mmpl_attn_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 32)()
for (Lq, Lk, H) in [(4680, 18720, 12), (4680, 32760, 12), (10920, 14040, 40)]:
    q = torch.randn(Lq, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(Lk, H, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(Lk, H, 128, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(q)
    for _ in range(3):
        ops.flash_attn(q, k, v, out=out)
    raw.mmpl_attn_debug_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.flash_attn(q, k, v, out=out)
    e1.record()
    raw.mmpl_attn_debug_read(buf, 1)
    d = list(buf)
    us = e0.elapsed_time(e1) * 1e3
    tiles = max(d[13], 1)
    print(f"Lq={Lq} Lk={Lk} H={H}: {us:.0f} us, CTA0 ran {tiles} KV tiles; MMA warp {d[12] / tiles:.0f} clk/tile "
          f"(waiting: P0 {d[8] / tiles:.0f}, P1 {d[9] / tiles:.0f}, V {d[10] / tiles:.0f}, K {d[11] / tiles:.0f})")
    for qt in range(2):
        n = max(d[qt * 4 + 3], 1)
        print(f"   softmax warp of q tile {qt}: per tile wait-S {d[qt * 4] / n:.0f}  compute-P {d[qt * 4 + 1] / n:.0f}  "
              f"store+arrive {d[qt * 4 + 2] / n:.0f} clk  ({n} tiles)")
