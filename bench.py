#!/usr/bin/env python
"""bench.py — denoised latent frames/s of the chunk-wise causal denoising hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|tiny] [--no-chain]

One "step" = one pass of the hot path over one batch of synthetic input: a full CausalInferencePipeline.inference()
rollout of the workload (cfg2: Wan2.1-T2V-1.3B dims, 21 latent frames at 60x104, 3-frame chunks, 4 denoising steps +
1 clean-context pass per chunk = 35 backbone forwards, KV length 4680 -> 32760), random-init weights, synthetic
latents and prompt embeddings. Inputs (63 MB of activations per forward, 6 GB of KV cache) exceed the 126 MB L2, so
there is no L2 flush between steps.

`value`  : whole-job latent frames/s with inputs resident in HBM, timed with CUDA events (max over ranks).
`e2e`    : the same metric through the public API with HOST buffers: pinned-host noise + prompt embeddings copied to
           the device and the denoised latents read back inside the timed region, every step.
`roofline`: the dominant kernel (tcgen05 self-attention over the KV cache): algorithmic FLOPs / CUDA-event time of its
           launches during the timed region, against the measured bf16 peak in MEASURED_PEAKS.json; `traffic` comes
           from the committed ncu capture named in profiles/ncu_traffic.json, and only while the attention kernel's
           sources still hash to what that capture was taken from.
`parity` : (N = 1) the first chunk of the same rollout, same weight / noise / prompt seeds and the same re-noising draws,
           on the GPU path against the CPU arm's output in the same run: max-abs, cosine, finite; plus resident and
           host-buffer steps giving identical latents.
`chain`  : the MMPL segment-parallel mode (BASELINE configs 3-5) measured in the same process at every N:
           Wan2.1-14B dims, CausalFPSWanModel, stages [2,7,6,6], 50 UniPC steps x CFG (fused CFG + UniPC kernel), anchors
           over NCCL, CFG-pair lanes from 2 GPUs up, VAE segment connect on every hand-off (tools/chain_bench.py). One
           record per layout (chains x slots x lanes) with frames/s, T_anchor/T_segment, per-rank finish times, anchor
           bytes, NCCL counts, connect ms and the limiter. The top-level `value` stays the cfg2 replica line so that rounds
           compare.
`cpu_baseline` / `--impl reference`: the reference algorithm on the host cores. The reference is pure Python/PyTorch and
           /root/reference does not exist on the GPU box, so this is the oracle port (oracle/cpu_port.py: the oracle
           restatement with native bf16 torch CPU ops, what the reference's CPU path executes; pinned against the
           reference golden by tests/test_oracle_golden.py). `--impl reference` times the WHOLE cfg2 rollout (same config
           as the GPU arm; as many steps as fit a time budget, at least one, true count reported). The in-line
           `cpu_baseline` of the default run is the bounded sample the contract asks for: the first of cfg2's seven
           chunks (3 latent frames at 60x104, 5 forwards, KV 4680 - the cheapest chunk, so an upper bound).
N > 1 (torchrun, one rank per GPU): the few-step CausalInferencePipeline path is strictly sequential over chunks, so for
the top-level line ranks are independent replicas (weak scaling, no data-path collective); the `chain`
record is where the ranks cooperate.
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
# stdout must carry exactly one JSON line, whatever libraries print (NCCL's version banner goes to fd 1 when NCCL_DEBUG is
# VERSION or higher and no NCCL_DEBUG_FILE is set): fd 1 is pointed at stderr for the whole run and the line is written to
# the saved descriptor by emit().
os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(record: dict) -> None:
    os.write(_STDOUT_FD, (json.dumps(record) + "\n").encode())


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
RNG_SEED = 1234   # re-noising draws (torch.randn_like in the reference, causal_inference.py:208): a CPU generator on both arms


def synthetic_inputs(cfg_dims, frames, lh, lw, noise_seed=0):
    """Noise [F, C, H, W] and prompt embeddings [text_len, text_dim], bf16: seed 0 / seed 1 on both arms."""
    text_len, text_dim = cfg_dims.get("text_len", 512), cfg_dims.get("text_dim", 4096)
    noise = torch.randn(frames, 16, lh, lw, generator=torch.Generator().manual_seed(noise_seed)).to(torch.bfloat16)
    prompt = torch.randn(text_len, text_dim, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    return noise, prompt


def cpu_reference_run(steps: int, warmup: int, sample: str = "cfg2_chunk0", budget_s: float = 0.0, weights=None):
    """Times the oracle port of the reference pipeline on the host cores (native bf16 torch CPU ops). With `budget_s`,
    stops early once another step would not fit (at least one step runs). Returns the last step's latents too.
    `weights`: a state dict to run with (the in-line leg passes the GPU arm's own random-init weights, so that the two arms
    of one run compute the same function); default: the oracle's seeded random init of the same architecture."""
    from oracle import causal_wan_oracle as O
    from oracle import cpu_port
    dims, frames, lh, lw = WORKLOADS[sample]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.WanConfig(**dims)
    w = weights if weights is not None else O.make_weights(cfg, seed=0)
    noise, prompt = synthetic_inputs(dims, frames, lh, lw)
    times, latents, t_begin = [], None, time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        latents = cpu_port.causal_inference_native(cfg, w, noise, prompt, rng=torch.Generator().manual_seed(RNG_SEED))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s and (time.perf_counter() - t_begin) + dt > budget_s:
            break
    total = sum(times)
    return dict(value=frames * len(times) / total, ms=total / len(times) * 1e3, cores=cores, sample=WORKLOAD_DESC[sample],
                steps=len(times), latents=latents)


def run_reference(args):
    """`--impl reference`: the whole workload (same config as the GPU arm) on the host cores. One cfg2 step is ~5 min on 16
    cores, so there is no warm-up pass and the step count is what fits `--cpu-budget` seconds (at least 1, reported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.workload if args.cpu_sample == "same" else args.cpu_sample
    r = cpu_reference_run(max(1, args.steps), 0, sample, budget_s=args.cpu_budget)
    same = sample == args.workload
    emit({
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": 0, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "same_config": same,
                   "note": ("whole workload on the host cores; steps = what fits --cpu-budget "
                            f"({args.cpu_budget:.0f} s), requested {args.steps}") if same else
                           "CPU arm runs a bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "steps": r["steps"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


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


def ncu_traffic(workload: str):
    """`roofline.traffic`: DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture that
    profiles/ncu_traffic.json names, valid only while the kernel's sources still hash to what was captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if workload != "cfg2" or not os.path.exists(path):
        return None, {"traffic_source": None}
    with open(path) as f:
        rec = json.load(f)
    import hashlib
    h = hashlib.sha256()
    for name in rec["sources"]:
        with open(os.path.join(ROOT, name), "rb") as f:
            h.update(f.read())
    current = h.hexdigest()[:16]
    info = {"traffic_source": rec["capture"], "traffic_launch": rec["launch"], "traffic_algorithmic": rec.get("algorithmic_bytes"),
            "traffic_kernel_sha": rec["sources_sha"], "traffic_kernel_sha_now": current}
    return (rec["dram_bytes"] if current == rec["sources_sha"] else None), info


def chunk0_parity(pipe, holder, device, cpu_latents):
    """The first cfg2 chunk on the GPU path with the CPU arm's seeds and re-noising draws, against the CPU arm's latents."""
    dims, frames, lh, lw = WORKLOADS["cfg2_chunk0"]
    noise, prompt = synthetic_inputs(dims, frames, lh, lw)
    holder["prompt"] = prompt[None].to(device)
    g = torch.Generator().manual_seed(RNG_SEED)
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: torch.randn(x.shape, generator=g, dtype=x.dtype).to(x.device)
    try:
        lat = pipe.inference(noise=noise[None].to(device), text_prompts=["synthetic"], return_latents=True)[1]
    finally:
        torch.randn_like = orig
    got, ref = lat[0].float().cpu(), cpu_latents.float()
    return {"what": "cfg2 chunk 0 (3 frames, 5 forwards, 30 blocks): GPU path vs the CPU arm of this run, same seeds and draws",
            "max_abs": (got - ref).abs().max().item(),
            "cos": torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item(),
            "finite": bool(torch.isfinite(got).all())}


def run_chain(args, rank, world, device, deadline):
    """The `chain` record (see the module docstring). Every rank takes part; rank 0 returns the record."""
    import torch.distributed as dist
    from tools.chain_bench import ChainBench, limiter_of
    cb = ChainBench(device, model="14B", layers=args.chain_layers, sampling_steps=args.chain_steps, vae_connect=not args.chain_no_vae)
    t0 = time.perf_counter()
    # untimed pass over every rank: communicators, tensor maps, KV caches, VAE workspaces (2-step segments, one more than
    # there are slots, so that every slot - slot 0 included - has received anchors and run the VAE connect once)
    cb.run(1, world // cb.lanes + 1, sampling_steps=2, warm=True)
    torch.cuda.synchronize()
    warm_s = time.perf_counter() - t0
    records, skipped = [], []
    est_forward_s = 0.30 * cb.dims["num_layers"] / 40   # prior: one 14B forward of ~6 frames on a B200; replaced by measurements
    for v in cb.variants():
        if args.chain_layouts and str(v["chains"]) not in args.chain_layouts.split(","):
            continue
        segs = 2 if v["slots"] == 1 else (3 * v["slots"] if v["chains"] > 1 else 4)
        # predicted wall time: segments run back to back on a slot; a chain emits one segment per anchor stage (~0.37 T_seg)
        steps = args.chain_steps
        t_seg = est_forward_s * (8 * steps + 8) / v["lanes"]
        pred = t_seg * (1 + (segs - 1) * max(0.37, 1.0 / v["slots"])) + 5
        left = deadline - time.perf_counter()
        if pred > left:
            steps = max(0, int(steps * (left - 10) / pred))
        flag = torch.tensor([steps], device=device)
        if world > 1:
            dist.broadcast(flag, src=0)   # rank 0's clock decides for everyone
        steps = int(flag.item())
        if steps < 8:
            skipped.append({"layout": f"{v['chains']} chain(s) x {v['slots']} slot(s) x {v['lanes']} lane(s)",
                            "why": f"{left:.0f} s of the time budget left, {pred:.0f} s needed at {args.chain_steps} steps"})
            continue
        rec = cb.run(v["chains"], segs, sampling_steps=steps)
        t_seg_ms = torch.tensor([rec["t_segment_ms"] if rec else 0.0], device=device)
        if world > 1:
            dist.broadcast(t_seg_ms, src=0)
        if t_seg_ms.item() > 0:
            est_forward_s = t_seg_ms.item() / 1e3 * v["lanes"] / (8 * steps + 8)
        if rec is not None:
            rec["limiter"] = limiter_of(rec)
            records.append(rec)
    vae_ends = cb.time_final_decode() if (deadline - time.perf_counter()) > 30 else None
    if rank != 0:
        return None
    best = max(records, key=lambda r: r["value"]) if records else None
    return {
        "config": {"workload": f"Wan2.1-14B MMPL T2V segment-parallel (BASELINE configs[2]): CausalFPSWanModel, {cb.dims['num_layers']} blocks, "
                               f"segments of 21 latent frames 60x104, stages [2,7,6,6], {args.chain_steps} UniPC steps x CFG 5.0 "
                               "(fused CFG + UniPC kernel), anchors over NCCL, VAE segment connect on every hand-off",
                   "n_gpus": world, "lanes": cb.lanes, "data": "synthetic, random-init weights (model and VAE)"},
        "value": best["value"] if best else None, "unit": UNIT, "best_layout": best["layout"] if best else None,
        "variants": records, "skipped": skipped, "vae_outside_the_bracket": vae_ends,
        "model_build_s": round(cb.build_s, 2), "warmup_s": round(warm_s, 2),
    }


def run_ours(args):
    import torch.distributed as dist
    from mmpl_b200 import _lib
    t_start = time.perf_counter()
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

    # the same seeds as the CPU arm on rank 0 (noise 0, prompt 1); other replicas get their own noise
    noise, prompt = synthetic_inputs(dims, frames, lh, lw, noise_seed=rank)
    noise_host, prompt_host = noise[None].pin_memory(), prompt[None].pin_memory()
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

    # timed region: K steps, inputs resident in HBM; self-attention launches bracketed by CUDA events - except for workloads
    # whose forwards are small enough to be replayed from CUDA graphs (launch-bound sizes, mmpl_b200/causal_model.py): events
    # cannot live inside a graph, so there the dominant kernel is timed in the instrumented step that follows instead
    ctx = model._ctx
    graphed = 3 * (lh // 2) * (lw // 2) <= model.graph_max_tokens and os.environ.get("MMPL_CUDA_GRAPHS", "1") != "0"
    _lib.check(lib.mmpl_profile_enable(ctx, 0 if graphed else 1 << 0))
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

    # resident and host-buffer steps are the same computation: with the re-noising generator reset they must agree bit for bit
    torch.manual_seed(RNG_SEED)
    lat_resident = step_resident().clone()
    torch.manual_seed(RNG_SEED)
    step_e2e()
    torch.cuda.synchronize()
    e2e_equal = bool(torch.equal(lat_resident.cpu(), out_host))
    finite = bool(torch.isfinite(lat_resident.float()).all())

    # one extra instrumented step: time share per kernel category (not part of any reported throughput)
    _lib.check(lib.mmpl_profile_enable(ctx, 0xF))
    _lib.check(lib.mmpl_profile_read(ctx, ms_arr, work_arr, n_arr, 1))
    s_ms, s_work, s_n = (C.c_double * 14)(), (C.c_double * 14)(), (C.c_int64 * 14)()
    _lib.check(lib.mmpl_profile_read_sites(ctx, s_ms, s_work, s_n, 1))
    ms_prof = timed(step_resident, 1)
    _lib.check(lib.mmpl_profile_read(ctx, ms_arr, work_arr, n_arr, 1))
    _lib.check(lib.mmpl_profile_read_sites(ctx, s_ms, s_work, s_n, 1))
    _lib.check(lib.mmpl_profile_enable(ctx, 0))
    if graphed:
        attn_ms, attn_flops, attn_n = ms_arr[0], work_arr[0], n_arr[0]
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

    out = None
    if rank == 0:
        fs = (lh // 2) * (lw // 2)
        chunks = frames // 3
        step_flops = sum(forward_flops(mdims, 3 * fs, (c + 1) * 3 * fs, 3, text_len) * 5 for c in range(chunks))
        peak_tf, peak_gbs, peak_src = measured_peaks()
        value = world * frames * K / (ms_total / 1e3)
        e2e_value = world * frames * K / (ms_e2e / 1e3)
        achieved_tf = attn_flops / max(attn_ms, 1e-9) / 1e9
        traffic, traffic_info = ncu_traffic(args.workload)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "batch": 1, "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "l2": ("inputs larger than L2 (no flush): 6 GB KV cache + 2.8 GB weights per step" if args.workload != "cfg2_14b"
                              else "inputs larger than L2 (no flush): 26.8 GB KV cache + 28 GB weights per step"),
                       "step_tflop": round(step_flops / 1e12, 1),
                       "model_tflops_per_gpu": round(step_flops * K / (ms_total / 1e3) / 1e12, 1),
                       "seeds": {"weights": 0, "noise": "rank", "prompt": 1}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": noise_host.numel() * 2 + prompt_host.numel() * 2,
                    "d2h_bytes_per_step": out_host.numel() * 2},
            "gpu_launches": launches,
            "roofline": dict({"kernel": "flash_attn_kernel (tcgen05 self-attention over the KV cache)", "bound": "tensor",
                              "achieved": round(achieved_tf, 1), "peak": peak_tf, "unit": "TFLOP/s",
                              "frac": round(achieved_tf / peak_tf, 4), "traffic": traffic,
                              "flops_per_launch_avg": round(attn_flops / max(attn_n, 1)), "peak_source": peak_src,
                              "launches": int(attn_n), "ms_in_timed_region": round(attn_ms, 2),
                              "share_of_step": round(attn_ms / (ms_prof if graphed else ms_total), 4),
                              "timed_in": ("one instrumented step with eager launches (the timed region replays CUDA graphs, "
                                           "which cannot carry per-launch events)") if graphed else "the timed region"},
                             **traffic_info),
            "parity": {"finite": finite, "resident_equals_e2e": e2e_equal},
            "breakdown": breakdown,
        }

    # CPU arm (rank 0 of a 1-GPU run): bounded sample of the same workload + parity of the GPU path against its output
    if world == 1 and not args.no_cpu_baseline:
        sample = "cfg2_chunk0" if args.cpu_sample == "same" else args.cpu_sample
        same_arch = WORKLOADS[sample][0] == dims
        r = cpu_reference_run(steps=1, warmup=0, sample=sample,
                              weights={k: v.detach().cpu() for k, v in model.state_dict().items()} if same_arch else None)
        out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"], "ms": r["ms"]}
        if sample == "cfg2_chunk0" and args.workload == "cfg2":
            out["parity"].update(chunk0_parity(pipe, holder, device, r["latents"]))
        if sample != "cfg1":
            # SURVEY.md §8(d) quotes the CPU baseline on BASELINE config 0 (30x52 latent frames: a quarter of the tokens
            # per frame); reported next to the same-workload sample, never used as the headline
            r1 = cpu_reference_run(steps=1, warmup=0, sample="cfg1")
            out["cpu_baseline"]["config0"] = {"value": r1["value"], "unit": UNIT, "sample": r1["sample"], "ms": r1["ms"]}

    # MMPL segment-parallel chain (configs 3-5), same process, every N
    if not args.no_chain and args.workload == "cfg2":
        del pipe, model
        holder.clear()
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        # a failure in the chain measurement must not cost the cfg2 line: it is recorded in the line instead
        try:
            chain = run_chain(args, rank, world, device, deadline=t_start + args.time_budget)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            chain, chain_failed = {"error": f"{type(e).__name__}: {e}"}, True
        else:
            chain_failed = False
        if rank == 0:
            out["chain"] = chain
    else:
        chain_failed = False
    if rank == 0:
        out["wall_s"] = round(time.perf_counter() - t_start, 1)
        emit(out)
    if chain_failed:
        os._exit(0)   # peers may be stuck in a collective of the failed measurement: leave without the group's shutdown handshake
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
    ap.add_argument("--cpu-sample", default="same", choices=["same", "cfg2_chunk0", "cfg1", "tiny"],
                    help="what the CPU arm times: `same` = the whole workload for --impl reference and its first chunk for the "
                         "in-line cpu_baseline")
    ap.add_argument("--cpu-budget", type=float, default=500.0, help="--impl reference: stop adding steps beyond this many seconds")
    ap.add_argument("--no-chain", action="store_true", help="skip the MMPL segment-parallel `chain` record")
    ap.add_argument("--chain-steps", type=int, default=50, help="UniPC steps of the chain record (the reference's 50)")
    ap.add_argument("--chain-layers", type=int, default=0, help="override the 14B model's 40 blocks (quick checks only)")
    ap.add_argument("--chain-layouts", default="", help="comma-separated chain counts to measure (default: every layout of the box)")
    ap.add_argument("--chain-no-vae", action="store_true", help="pass-through connect instead of the VAE (labelled in the record)")
    ap.add_argument("--time-budget", type=float, default=780.0,
                    help="wall seconds the whole run may take: chain layouts that would not fit run with fewer UniPC steps "
                         "(stated per variant) or are skipped (listed)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
