#!/usr/bin/env python
"""Micro-benchmarks of the hot kernels at the cfg2 / 14B shapes (CUDA events, L2-cold rotation of buffers)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmpl_b200 import ops  # noqa: E402

dev = "cuda"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timeit_graph(fn, iters=20, reps=5):
    """Same, but the `iters` launches are captured into a CUDA graph first: no CPU launch cost between kernels,
    which is how they run inside mmpl_forward (back-to-back launches from C with programmatic dependent launch)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps)


if os.environ.get("BENCH_GRAPH", "1") != "0":
    timeit = timeit_graph  # noqa: F811


def bench_gemm(shapes, tiles):
    for (M, N, K) in shapes:
        nbuf = 4
        xs = [torch.randn(M, K, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(nbuf)]
        b = torch.randn(N, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        line = f"gemm M={M} N={N} K={K}:"
        for t in tiles:
            if t in (128, 256) and N % t:
                continue
            if t == 0 and os.environ.get('GEMM_NO_AUTO'):
                continue
            i = [0]

            def run():
                i[0] = (i[0] + 1) % nbuf
                ops.linear(xs[i[0]], ws[i[0]], b, out=out, tile_n=t)
            ms = timeit(run)
            line += f"  tile{t}: {ms * 1e3:7.1f} us {2 * M * N * K / ms / 1e9:7.1f} TF/s |"
        i = [0]

        def run_t():
            i[0] = (i[0] + 1) % nbuf
            torch.nn.functional.linear(xs[i[0]], ws[i[0]], b)
        ms = timeit(run_t)
        line += f"  cuBLAS: {ms * 1e3:7.1f} us {2 * M * N * K / ms / 1e9:7.1f} TF/s"
        print(line, flush=True)


SPLITS = [int(x) for x in os.environ.get('ATTN_SPLITS', '0,1,3').split(',')]  # 0 = cost model


def bench_attn(cases):
    for (Lq, Lk, H) in cases:
        q = torch.randn(Lq, H, 128, device=dev, dtype=torch.bfloat16)
        k = torch.randn(Lk, H, 128, device=dev, dtype=torch.bfloat16)
        v = torch.randn(Lk, H, 128, device=dev, dtype=torch.bfloat16)
        out = torch.empty_like(q)
        from mmpl_b200 import _lib
        lib = _lib.load()
        line = f"attn Lq={Lq} Lk={Lk} H={H}:"
        for sp in SPLITS:
            lib.mmpl_attn_set_split(sp)
            ms = timeit(lambda: ops.flash_attn(q, k, v, out=out), iters=10)
            line += f" split{sp}: {ms * 1e3:7.1f} us {4 * Lq * Lk * H * 128 / ms / 1e9:6.1f} |"
            print(f"  [{Lq}x{Lk}x{H}] split{sp}: {ms * 1e3:7.1f} us", flush=True)
        lib.mmpl_attn_set_split(0)
        if os.environ.get("ATTN_CUDNN", "1") != "0":  # cuDNN's Blackwell fused attention through torch SDPA, as a yardstick
            try:
                from torch.nn.attention import SDPBackend, sdpa_kernel
                q4, k4, v4 = (t.transpose(0, 1)[None] for t in (q, k, v))  # [1, H, L, 128] views
                with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                    ms3 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4), iters=10)
                line += f" | cuDNN SDPA: {ms3 * 1e3:8.1f} us {4 * Lq * Lk * H * 128 / ms3 / 1e9:7.1f} TF/s"
            except Exception as e:  # noqa: BLE001
                line += f" | cuDNN SDPA unavailable ({type(e).__name__}: {str(e)[:80]})"
        if os.environ.get("ATTN_NO_FA2"):
            print(line, flush=True)
            continue
        try:
            from flash_attn import flash_attn_func
            ms2 = timeit(lambda: flash_attn_func(q[None], k[None], v[None]), iters=10)
            line += f" | flash-attn 2 wheel: {ms2 * 1e3:8.1f} us {4 * Lq * Lk * H * 128 / ms2 / 1e9:7.1f} TF/s"
        except Exception as e:  # noqa: BLE001
            line += f" | flash-attn unavailable ({type(e).__name__})"
        print(line, flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,attn")
    a = ap.parse_args()
    if "gemm" in a.what:
        gemm_shapes = [(4680, 4608, 1536), (4680, 1536, 1536), (4680, 8960, 1536), (4680, 1536, 8960),
                       (10920, 15360, 5120), (10920, 5120, 5120), (10920, 13824, 5120), (10920, 5120, 13824)]
        if os.environ.get("GEMM_SHAPES"):  # e.g. "4680x1536x1536"
            gemm_shapes = [tuple(int(v) for v in sh.split("x")) for sh in os.environ["GEMM_SHAPES"].split(",")]
        bench_gemm(gemm_shapes,
                   tiles=[int(t) for t in os.environ.get('GEMM_TILES', '128,256,512').split(',')])
    if "attn" in a.what:
        shapes = [(4680, 4680, 12), (4680, 18720, 12), (4680, 32760, 12), (4680, 512, 12), (10920, 14040, 40)]
        if os.environ.get("ATTN_SHAPES"):  # e.g. "4680x9360x12,4680x14040x12"
            shapes = [tuple(int(v) for v in sh.split("x")) for sh in os.environ["ATTN_SHAPES"].split(",")]
        bench_attn(shapes)
