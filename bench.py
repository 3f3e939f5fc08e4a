#!/usr/bin/env python
"""bench.py — denoised latent frames/s of the chunk-wise causal denoising hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|tiny]

One "step" = one pass of the hot path over one batch of synthetic input: a full CausalInferencePipeline.inference()
rollout of the workload (cfg2: Wan2.1-T2V-1.3B dims, 21 latent frames at 60x104, 3-frame chunks, 4 denoising steps +
1 clean-context pass per chunk = 35 backbone forwards, KV length 4680 -> 32760), random-init weights, synthetic
latents and prompt embeddings. Inputs (63 MB of activations per forward, 6 GB of KV cache) exceed the 126 MB L2, so
there is no L2 flush between steps.

`value`  : whole-job latent frames/s with inputs resident in HBM, timed with CUDA events (max over ranks).
`e2e`    : the same metric through the public API with HOST buffers: pinned-host noise + prompt embeddings copied to
           the device and the denoised latents read back inside the timed region, every step.
`roofline`: the dominant kernel (tcgen05 self-attention over the KV cache): algorithmic FLOPs / CUDA-event time of its
           launches during the timed region, against the measured bf16 peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the reference algorithm on the host cores. The reference is pure Python/PyTorch and
           /root/reference does not exist on the GPU box, so this is the oracle port (oracle/causal_wan_oracle.py) run
           with native bf16 torch CPU ops (what the reference's CPU path executes), on a bounded sample of the SAME
           workload: the first of cfg2's seven chunks (3 latent frames at 60x104, 5 forwards, KV length 4680 -- the
           cheapest chunk: 16.2 TFLOP per forward against 28.3 on average over the rollout, so the CPU figure is an
           upper bound for the whole video). `--cpu-sample cfg1` times BASELINE config 0 (30x52 frames) instead.
N > 1 (torchrun, one rank per GPU): the few-step CausalInferencePipeline path is strictly sequential over chunks, so
ranks are independent replicas (weak scaling, no data-path collective); see DESIGN.md §6.
The MMPL segment-parallel mode (BASELINE configs 3-5: Wan2.1-14B, anchors over NCCL, optional CFG-pair lanes and VAE
segment connect) has the same one-JSON-line driver at tools/run_segment_parallel.py (launched with torchrun the same
way; measured 1 / 2 / 4 / 8 B200 lines in profiles/).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints to stdout) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "denoised_latent_frames_per_s"
UNIT = "latent frames/s"

WORKLOADS = {
    # name: (model dims, frames, lat_h, lat_w)
    "cfg2": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30), 21, 60, 104),
    "cfg1": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30), 3, 30, 52),
    # the same chunk-wise rollout with the Wan2.1-14B backbone (BASELINE metric: "Wan-1.3B/14B causal"): 28 GB of weights,
    # 26.8 GB of KV cache; not the default bench line
    "cfg2_14b": (dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40), 21, 60, 104),
    "tiny": (dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32), 6, 8, 12),
    # CPU-arm sample: the first chunk of the cfg2 rollout
    "cfg2_chunk0": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30), 3, 60, 104),
}
WORKLOAD_DESC = {
    "cfg2": "Wan2.1-T2V-1.3B causal T2V, 21 latent frames 60x104, 3-frame chunks, 4 steps + context pass (35 forwards), KV 4680..32760, bf16",
    "cfg1": "Wan2.1-T2V-1.3B causal T2V, 1 chunk of 3 latent frames 30x52, 4 steps + context pass (5 forwards), bf16",
    "cfg2_14b": "Wan2.1-14B causal T2V, 21 latent frames 60x104, 3-frame chunks, 4 steps + context pass (35 forwards), KV 4680..32760, bf16",
    "tiny": "2-block dim-256 test model, 6 latent frames 8x12",
    "cfg2_chunk0": "first chunk of the cfg2 rollout: Wan2.1-T2V-1.3B causal T2V, 3 latent frames 60x104, 4 steps + context pass "
                   "(5 forwards, S = KV = 4680; the cheapest of the 7 chunks), bf16",
}


def forward_flops(dims, S, Lkv, frames, text_len=512, with_text=False):
    """Algorithmic FLOPs of one backbone forward."""
    D, Fd, L = dims["dim"], dims["ffn_dim"], dims["num_layers"]
    f = L * (12 * S * D * D + 4 * S * D * Fd + 4 * S * D * (Lkv + text_len)) + 2 * S * 64 * D + 2 * S * D * 64
    f += frames * (2 * 256 * D + 14 * D * D)
    if with_text:
        f += 2 * text_len * dims.get("text_dim", 4096) * D + 2 * text_len * D * D + 4 * text_len * D * D * L
    return f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1432.6), d.get("hbm_gbs", 6460.5), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(steps: int, warmup: int, sample: str = "cfg2_chunk0"):
    """Times the oracle port of the reference pipeline on the host cores (native bf16 torch CPU ops)."""
    from oracle import causal_wan_oracle as O
    from oracle import cpu_port
    dims, frames, lh, lw = WORKLOADS[sample]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.WanConfig(**dims)
    w = O.make_weights(cfg, seed=0)
    g0, g1 = torch.Generator().manual_seed(0), torch.Generator().manual_seed(1)
    noise = torch.randn(frames, cfg.in_dim, lh, lw, generator=g0).to(torch.bfloat16)
    prompt = torch.randn(cfg.text_len, cfg.text_dim, generator=g1).to(torch.bfloat16)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_port.causal_inference_native(cfg, w, noise, prompt)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return frames * len(times) / total, total / len(times) * 1e3, cores, WORKLOAD_DESC[sample]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one step is ~40 s on 16 host cores: a warm-up pass only for short runs, so that K = 5 still ends within minutes
    steps = max(1, args.steps)
    warmup = max(0, min(args.warmup, 1)) if steps <= 2 else 0
    value, ms, cores, sample = cpu_reference_run(steps, warmup, args.cpu_sample)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "note": "CPU arm runs a bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------ our arm
def build_pipeline(dims, device, text_len=512, text_dim=4096):
    from mmpl_b200.causal_model import CausalWanModel
    from mmpl_b200.pipeline import CausalInferencePipeline
    from mmpl_b200.wan_wrapper import WanDiffusionWrapper
    kw = dict(text_len=text_len, text_dim=text_dim)
    kw.update(dims)
    torch.manual_seed(0)
    with torch.device(device):
        model = CausalWanModel(**kw)
    model = model.to(dtype=torch.bfloat16)
    with torch.no_grad():  # random-init weights of the named architecture; the reference zeroes the head
        torch.nn.init.normal_(model.head.head.weight, std=0.02)
    model.eval().requires_grad_(False)
    gen = WanDiffusionWrapper(model=model, timestep_shift=5.0)
    holder = {}

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": holder["prompt"]}

    class VAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

    a = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                              independent_first_frame=False, context_noise=0, num_frame_per_block=3, model_kwargs={})
    pipe = CausalInferencePipeline(a, torch.device(device), generator=gen, text_encoder=Text(), vae=VAE())
    return pipe, model, holder


def run_ours(args):
    import torch.distributed as dist
    from mmpl_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mmpl_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
    lib = _lib.load()
    dims, frames, lh, lw = WORKLOADS[args.workload]
    text_len, text_dim = dims.get("text_len", 512), dims.get("text_dim", 4096)
    mdims = {k: v for k, v in dims.items() if k not in ("text_len", "text_dim")}
    pipe, model, holder = build_pipeline(mdims, device, text_len, text_dim)

    g = torch.Generator().manual_seed(1000 + rank)
    noise_host = torch.randn(1, frames, 16, lh, lw, generator=g).to(torch.bfloat16).pin_memory()
    prompt_host = torch.randn(1, text_len, text_dim, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).pin_memory()
    noise_dev = noise_host.to(device)
    holder["prompt"] = prompt_host.to(device)
    out_host = torch.empty(1, frames, 16, lh, lw, dtype=torch.bfloat16).pin_memory()

    def step_resident():
        return pipe.inference(noise=noise_dev, text_prompts=["synthetic"], return_latents=True)[1]

    def step_e2e():
        holder["prompt"] = prompt_host.to(device, non_blocking=True)
        lat = pipe.inference(noise=noise_host.to(device, non_blocking=True), text_prompts=["synthetic"], return_latents=True)[1]
        out_host.copy_(lat, non_blocking=True)
        return out_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    K, W = max(1, args.steps), max(3, args.warmup)
    for _ in range(W):
        step_resident()
    torch.cuda.synchronize()

    # timed region: K steps, inputs resident in HBM; self-attention launches bracketed by CUDA events
    ctx = model._ctx
    _lib.check(lib.mmpl_profile_enable(ctx, 1 << 0))
    import ctypes as C
    ms_arr, work_arr, n_arr = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_int64 * 4)()
    _lib.check(lib.mmpl_profile_read(ctx, ms_arr, work_arr, n_arr, 1))
    lib.mmpl_total_launches(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, K)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.mmpl_total_launches(0))
    if world > 1:
        lt = torch.tensor([launches], device=device, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    _lib.check(lib.mmpl_profile_read(ctx, ms_arr, work_arr, n_arr, 1))
    attn_ms, attn_flops, attn_n = ms_arr[0], work_arr[0], n_arr[0]
    _lib.check(lib.mmpl_profile_enable(ctx, 0))

    # end-to-end through the public API with host buffers
    step_e2e()
    ms_e2e = timed(step_e2e, K)

    # one extra instrumented step: time share per kernel category (not part of any reported throughput)
    _lib.check(lib.mmpl_profile_enable(ctx, 0xF))
    _lib.check(lib.mmpl_profile_read(ctx, ms_arr, work_arr, n_arr, 1))
    s_ms, s_work, s_n = (C.c_double * 14)(), (C.c_double * 14)(), (C.c_int64 * 14)()
    _lib.check(lib.mmpl_profile_read_sites(ctx, s_ms, s_work, s_n, 1))
    ms_prof = timed(step_resident, 1)
    _lib.check(lib.mmpl_profile_read(ctx, ms_arr, work_arr, n_arr, 1))
    _lib.check(lib.mmpl_profile_read_sites(ctx, s_ms, s_work, s_n, 1))
    _lib.check(lib.mmpl_profile_enable(ctx, 0))
    cats = ["self_attn", "cross_attn", "gemm", "pointwise"]
    breakdown = {c: {"ms": round(ms_arr[i], 3), "launches": int(n_arr[i]),
                     ("tflops" if i < 3 else "gbs"): round(work_arr[i] / max(ms_arr[i], 1e-9) / (1e9 if i < 3 else 1e6), 1)}
                 for i, c in enumerate(cats)}
    breakdown["step_ms_instrumented"] = round(ms_prof, 3)
    # the same spans by call site inside the block: [ms per step, launches, TFLOP/s or GB/s]
    site_names = ["self_attn", "cross_attn", "gemm_qkv", "gemm_o", "gemm_cross_q", "gemm_cross_o", "gemm_ffn0", "gemm_ffn2",
                  "ln_modulate", "ln_affine", "qk_norm_rope_kv", "rmsnorm", "modulation_add", "other"]
    breakdown["sites"] = {nm: [round(s_ms[i], 3), int(s_n[i]),
                               round(s_work[i] / max(s_ms[i], 1e-9) / (1e9 if i < 8 else 1e6), 1)]
                          for i, nm in enumerate(site_names)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fs = (lh // 2) * (lw // 2)
    chunks = frames // 3
    step_flops = sum(forward_flops(mdims, 3 * fs, (c + 1) * 3 * fs, 3, text_len) * 5 for c in range(chunks))
    peak_tf, peak_gbs, peak_src = measured_peaks()
    value = world * frames * K / (ms_total / 1e3)
    e2e_value = world * frames * K / (ms_e2e / 1e3)
    achieved_tf = attn_flops / max(attn_ms, 1e-9) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "batch": 1, "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                   "l2": ("inputs larger than L2 (no flush): 6 GB KV cache + 2.8 GB weights per step" if args.workload != "cfg2_14b"
                          else "inputs larger than L2 (no flush): 26.8 GB KV cache + 28 GB weights per step"),
                   "step_tflop": round(step_flops / 1e12, 1),
                   "model_tflops_per_gpu": round(step_flops * K / (ms_total / 1e3) / 1e12, 1)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": noise_host.numel() * 2 + prompt_host.numel() * 2,
                "d2h_bytes_per_step": out_host.numel() * 2},
        "gpu_launches": launches,
        "roofline": {"kernel": "flash_attn_kernel (tcgen05 self-attention over the KV cache)", "bound": "tensor",
                     "achieved": round(achieved_tf, 1), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(achieved_tf / peak_tf, 4),
                     # dram__bytes_read.sum + dram__bytes_write.sum of the L_kv = 32760 launch in
                     # profiles/r01_ncu_attention_hybrid_vs_cudnn.txt (ncu --set full): 350.5 MB + 24.5 MB; its algorithmic
                     # K/V + Q + O bytes are 230 MB, the rest is K/V read a second time by the ranged units of the hybrid
                     # schedule (5 of 12 heads) and their partials (the uniform split it replaced: 231.9 + 63.1 MB and a
                     # merge kernel with 88 MB more)
                     "traffic": 375.0e6 if args.workload == "cfg2" else None, "traffic_launch": "L_kv=32760, S=4680, 12 heads",
                     "flops_per_launch_avg": round(attn_flops / max(attn_n, 1)), "peak_source": peak_src,
                     "launches": int(attn_n), "ms_in_timed_region": round(attn_ms, 2),
                     "share_of_step": round(attn_ms / ms_total, 4)},
        "breakdown": breakdown,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = cpu_reference_run(steps=1, warmup=0, sample=args.cpu_sample)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms": ms}
        if args.cpu_sample != "cfg1":
            # SURVEY.md §8(d) quotes the CPU baseline on BASELINE config 0 (30x52 latent frames: a quarter of the tokens
            # per frame); reported next to the same-workload sample, never used as the headline
            v1, ms1, _, sample1 = cpu_reference_run(steps=1, warmup=0, sample="cfg1")
            out["cpu_baseline"]["config0"] = {"value": v1, "unit": UNIT, "sample": sample1, "ms": ms1}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", default="cfg2_chunk0", choices=["cfg2_chunk0", "cfg1", "tiny"],
                    help="what the CPU arm (cpu_baseline / --impl reference) times")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
