#!/usr/bin/env python
"""Times the VAE segment connect (mmpl_b200.vae.WanVAEWrapper.segment_connect) on one B200 at the
reference's geometry: anchors [1, 8, 16, 60, 104] -> 4 latent frames decoded to 13 pixel frames at 480x832, frames 8..12
re-encoded to 2 latents. CUDA events on the launching stream, warm-up first; seeded random-init VAE (checkpoint absent).

    python tools/bench_vae_connect.py [--iters 5] [--warmup 2] [--latent-h 60 --latent-w 104] [--i2v]

Prints JSON lines: first the connect - ms per connect, algorithmic TFLOP (2 x positions x taps x Cin x Cout over every convolution + the
middle attention) and the fraction of the measured sustained bf16 peak (MEASURED_PEAKS.json), plus the per-call launch count.
Then the final 21 -> 81 frame decode of a segment and the i2v image encode (skip with --skip-extras).
Run under `ncu --set full -k regex:gemm_bf16_kernel` for the tap-GEMM capture."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def connect_flops(vae, h, w):
    """Algorithmic FLOPs of segment_connect at latent size h x w, from the bound weights and the layer programs."""
    total = 0.0

    def conv(name, frames, hh, ww):
        kt, kh, kw = vae._kernel[name]
        cout8, taps, cin64 = vae._w[name + ".weight"].shape
        cout, cin = vae._shape[name]
        return 2.0 * frames * hh * ww * kt * kh * kw * cin * cout

    def run(prog, t, hh, ww, c):
        nonlocal total
        for kind, p in prog:
            if kind == "res":
                cout = vae._res_out[p]
                total += conv(p + ".residual.2", t, hh, ww) + conv(p + ".residual.6", t, hh, ww)
                if (p + ".shortcut") in vae._kernel:
                    total += conv(p + ".shortcut", t, hh, ww)
                c = cout
            elif kind == "attn":
                n = hh * ww
                total += t * (2.0 * n * c * 3 * c + 4.0 * n * n * c + 2.0 * n * c * c)
            elif kind in ("up2d", "up3d"):
                if kind == "up3d" and t > 1:
                    total += conv(p + ".time_conv", t - 1, hh, ww)
                    t = 1 + 2 * (t - 1)
                hh, ww = 2 * hh, 2 * ww
                total += conv(p + ".resample.1", t, hh, ww)
                c //= 2
            else:
                hh, ww = hh // 2, ww // 2
                total += conv(p + ".resample.1", t, hh, ww)
                if kind == "down3d" and t > 1:
                    total += conv(p + ".time_conv", (t - 1) // 2, hh, ww)
                    t = 1 + (t - 1) // 2
        return t, hh, ww, c

    top = vae.dim * vae._dim_mult[-1]
    total += conv("conv2", 4, h, w) + conv("decoder.conv1", 4, h, w)
    t, hh, ww, c = run(vae._dec, 4, h, w, top)
    total += conv("decoder.head.2", t, hh, ww)
    total += conv("encoder.conv1", 5, 8 * h, 8 * w)
    t, hh, ww, c = run(vae._enc, 5, 8 * h, 8 * w, vae.dim)
    total += conv("encoder.head.2", t, hh, ww) + conv("conv1", t, hh, ww)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--latent-h", type=int, default=60)
    ap.add_argument("--latent-w", type=int, default=104)
    ap.add_argument("--i2v", action="store_true", help="3 anchors (frames 0, 19, 20) instead of the t2v payload of 8")
    ap.add_argument("--skip-extras", action="store_true", help="only the connect (no final decode / image encode timing)")
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the mmpl_b200 path has no CPU fallback")
    from mmpl_b200 import _lib
    from mmpl_b200.vae import WanVAEWrapper
    lib = _lib.load()
    dev = "cuda:0"
    vae = WanVAEWrapper()
    sd = vae.init_random_weights(seed=0, device=dev)
    vae._shape = {k[:-len(".weight")]: (v.shape[0], v.shape[1]) for k, v in sd.items() if k.endswith(".weight") and v.dim() >= 4}
    anchors = torch.randn(1, 3 if a.i2v else 8, 16, a.latent_h, a.latent_w, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    for _ in range(max(1, a.warmup)):
        out = vae.segment_connect(anchors)
    torch.cuda.synchronize()
    lib.mmpl_total_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = vae.segment_connect(anchors)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    launches = int(lib.mmpl_total_launches(0)) // a.iters
    flops = connect_flops(vae, a.latent_h, a.latent_w)
    peak = 1400.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", peak))
    except Exception:
        pass
    tf = flops / (ms * 1e-3) / 1e12
    record = lambda d: print(json.dumps(d), flush=True)  # noqa: E731
    record({"what": "vae segment connect", "ms": ms, "algorithmic_tflop": flops / 1e12, "tflops": tf, "peak_tflops": peak,
                      "frac_of_sustained_bf16_peak": tf / peak, "gpu_launches_per_connect": launches,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "finite": bool(torch.isfinite(out.float()).all()),
                      "geometry": f"latent {a.latent_h}x{a.latent_w}, 4 latent -> 13 pixel frames decoded, 5 pixel frames encoded"})
    if a.skip_extras:
        return

    def timed(fn, iters=2):
        fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            y = fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / iters, y

    # the decode that ends a segment (pipeline/casual_fps_inference.py:445: vae.decode_to_pixel(output), 21 latent -> 81 pixel
    # frames at 480x832) and the i2v image encode (MMPL_i2v/Wan_fps_inference_parallel_4gpu_20s.py:190-194: one frame)
    g = torch.Generator().manual_seed(2)
    lat = torch.randn(1, 21, 16, a.latent_h, a.latent_w, generator=g).to(torch.bfloat16).to(dev)
    torch.cuda.reset_peak_memory_stats()
    ms_dec, video = timed(lambda: vae.decode_to_pixel(lat))
    record({"what": "final decode of one segment: 21 latent frames -> 81 pixel frames", "ms": ms_dec, "shape": list(video.shape),
            "finite": bool(torch.isfinite(video).all()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
            "pixel_frames_per_s": video.shape[1] / (ms_dec / 1e3)})
    del video
    image = (torch.rand(1, 3, 1, 8 * a.latent_h, 8 * a.latent_w, generator=g) * 2 - 1).to(torch.bfloat16).to(dev)
    ms_enc, z = timed(lambda: vae.encode_to_latent(image), iters=5)
    record({"what": "i2v image encode: 1 pixel frame -> 1 latent frame", "ms": ms_enc, "shape": list(z.shape),
            "finite": bool(torch.isfinite(z).all())})


if __name__ == "__main__":
    main()
