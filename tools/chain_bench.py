"""MMPL segment-parallel chain measurement (BASELINE.json configs 3-5) for bench.py's `chain` record and for
tools/run_segment_parallel.py: Wan2.1-14B dimensions, `CausalFPSWanModel`, stages [2,7,6,6] (t2v) or [1,1,7,6,6] (i2v),
UniPC steps x CFG, anchors over NCCL, CFG-pair lanes when there are at least two ranks, and the VAE segment connect on every
hand-off. Random-init weights and synthetic inputs (there are no checkpoints on the box).

A *variant* is one way of laying a box of N ranks out:  chains x slots x lanes = N
    lanes   1, or 2 = CFG-pair split (conditional / unconditional forwards on two ranks, one flow all-gather per step)
    slots   segment slots of one chain: slot s runs segments s, s + slots, ... of one video; anchors go slot -> slot + 1
    chains  independent videos on disjoint rank groups (no exchange between them)
Every variant generates `segments_per_chain` segments of 21 latent frames per chain and reports denoised latent frames/s
per box = chains * segments * 21 / device time of the slowest rank (CUDA events), plus what explains it:
T_anchor / T_segment (the chain emits at most one segment per anchor stage), per-rank finish times, anchor
bytes, point-to-point and all-gather counts, VAE-connect time per boundary.
"""
from __future__ import annotations

import os
import time
import types
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from mmpl_b200.causal_model import CausalFPSWanModel
from mmpl_b200.pipeline import CausalFPSInferencePipeline
from mmpl_b200.segment_parallel import (AnchorChannel, I2V_ANCHOR_SHAPE, SegmentParallelRunner, T2V_ANCHOR_SHAPE,
                                        passthrough_connect, vae_segment_connect)
from mmpl_b200.wan_wrapper import MODEL_CONFIGS, WanFPSWrapper

LAT_H, LAT_W, FRAMES = 60, 104, 21


def forward_flops(dims: dict, S: int, Lkv: int, frames: int, text_len: int = 512) -> float:
    """Algorithmic FLOPs of one backbone forward."""
    D, Fd, L = dims["dim"], dims["ffn_dim"], dims["num_layers"]
    return (L * (12 * S * D * D + 4 * S * D * Fd + 4 * S * D * (Lkv + text_len)) + 2 * S * 64 * D + 2 * S * D * 64
            + frames * (2 * 256 * D + 14 * D * D))


def segment_flops(dims: dict, steps: int, i2v: bool = False, first: bool = True) -> float:
    """Algorithmic FLOPs of one 21-frame t2v segment: per stage (2*steps + 2) forwards of n frames against the frames
    visible at that stage (136 421 TFLOP at 14B / 50 steps). Segments after the first replace stage 0 by
    two t=0 prefill forwards."""
    fs = (LAT_H // 2) * (LAT_W // 2)
    stages = [(2, 2), (7, 9), (6, 13), (6, 21)]          # (frames of the stage, frames visible to it)
    total = 0.0
    for i, (n, vis) in enumerate(stages):
        calls = 2 * steps + 2 if (first or i > 0) else 2
        total += calls * forward_flops(dims, n * fs, vis * fs, n)
    return total


def chain_layouts(world: int) -> List[dict]:
    """Every layout chains x slots x lanes = world worth measuring, most informative first: lanes = 2 (CFG pair) from two
    ranks up; chain counts are the powers of two that leave every chain at least one whole slot; two chains first on a box
    that has them (the layout that fills 8 GPUs), then the single long video, then the narrower ones."""
    lanes = 2 if world >= 2 else 1
    slots_total = world // lanes
    out, chains = [], 1
    while chains <= slots_total and slots_total % chains == 0 and world % chains == 0:
        out.append(dict(chains=chains, slots=slots_total // chains, lanes=lanes))
        chains *= 2
    order = {2: 0, 1: 1}
    return sorted(out, key=lambda v: order.get(v["chains"], v["chains"]))


class ChainBench:
    """Everything that is built once per process: the 14B replica, the pipeline (with its CFG-pair group), the VAE and every
    rank group the variants need."""

    def __init__(self, device: str, model: str = "14B", layers: int = 0, sampling_steps: int = 50, i2v: bool = False,
                 vae_connect: bool = True):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device, self.i2v, self.sampling_steps = device, i2v, sampling_steps
        self.lanes = 2 if self.world >= 2 else 1
        dims = dict(MODEL_CONFIGS["Wan2.1-T2V-14B" if model == "14B" else "Wan2.1-T2V-1.3B"])
        if layers:
            dims["num_layers"] = layers
        self.dims, self.model_name = dims, model
        t0 = time.perf_counter()
        torch.manual_seed(0)  # identical random-init replica on every rank
        with torch.device(device):
            self.model = CausalFPSWanModel(**dims)
        self.model = self.model.to(torch.bfloat16).eval().requires_grad_(False)
        self.build_s = time.perf_counter() - t0
        gen = WanFPSWrapper(model=self.model, timestep_shift=5.0)
        # rank groups, created collectively and in the same order everywhere: CFG pairs, then the chain splits
        self.pair_group = None
        if self.lanes == 2:
            for s in range(self.world // 2):
                g = dist.new_group([2 * s, 2 * s + 1])
                if self.rank // 2 == s:
                    self.pair_group = g
        self.chain_groups: Dict[int, Optional[dist.ProcessGroup]] = {1: None}
        chains = 2
        while self.world // chains >= self.lanes and self.world % chains == 0:
            per = self.world // chains
            for c in range(chains):
                g = dist.new_group(list(range(c * per, (c + 1) * per)))
                if self.rank // per == c:
                    self.chain_groups[chains] = g
            chains *= 2
        self.prompts: Dict[str, torch.Tensor] = {}
        negative = self._embed(2)

        class Text(torch.nn.Module):
            def forward(inner, text_prompts):
                key = text_prompts[0]
                return {"prompt_embeds": negative if key == "__negative__" else self.prompts[key]}

        class PassVAE(torch.nn.Module):  # the final decode is not part of the denoise bracket
            def decode_to_pixel(inner, latents, use_cache=False):
                return latents

        args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0,
                                     negative_prompt="__negative__", independent_first_frame=False,
                                     sampling_steps=sampling_steps, model_kwargs={}, i2v=i2v)
        torch.manual_seed(1234)  # constructor randint + re-noising draws: the same stream on both lanes of a pair
        self.pipe = CausalFPSInferencePipeline(args, device, generator=gen, text_encoder=Text(), vae=PassVAE(),
                                               device_cond=device, device_uncond=device, cfg_group=self.pair_group)
        self.vae = None
        if vae_connect:
            from mmpl_b200.vae import WanVAEWrapper
            self.vae = WanVAEWrapper()
            self.vae.init_random_weights(seed=0, device=device)
        self.first_frame = torch.randn(1, 1, 16, LAT_H, LAT_W, generator=torch.Generator().manual_seed(7)).to(torch.bfloat16).to(device)

    def _embed(self, seed: int) -> torch.Tensor:
        return torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16).to(self.device)

    def variants(self) -> List[dict]:
        """Layouts worth measuring on this world size, most informative first."""
        return [v for v in chain_layouts(self.world) if v["chains"] in self.chain_groups]

    # -------------------------------------------------------------------------------------------------------------- run
    def run(self, chains: int, segments_per_chain: int, sampling_steps: Optional[int] = None, warm: bool = False) -> Optional[dict]:
        """One variant. Returns the record on rank 0 (None elsewhere)."""
        dev, world, rank = self.device, self.world, self.rank
        per = world // chains
        chain, crank = rank // per, rank % per
        group = self.chain_groups[chains]
        self.pipe.sampling_steps = sampling_steps or self.sampling_steps
        steps = self.pipe.sampling_steps
        key = f"synthetic prompt {chain}"
        self.prompts[key] = self._embed(1 + 10 * chain)
        channel = AnchorChannel(group=group, lanes=self.lanes)
        slots = per // self.lanes
        if slots > 1:  # create the point-to-point connections outside the timed region
            buf = torch.zeros(8, device=dev)
            nxt = (crank + self.lanes) % per
            prv = (crank - self.lanes) % per
            gl = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
            ops = [dist.P2POp(dist.isend, buf, gl(nxt), group), dist.P2POp(dist.irecv, torch.empty_like(buf), gl(prv), group)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        connect = vae_segment_connect(self.vae) if self.vae is not None else passthrough_connect
        # timing marks (CUDA events on the compute stream): per segment its start, the connect, the hand-off stage, its end
        marks: Dict[int, dict] = {}
        current = [None]

        def observer(event, seg):
            current[0] = seg
            marks.setdefault(seg, {})[event] = self._mark()

        def on_stage(index, rec, latents):
            if rec.handoff is not None:
                marks[current[0]]["anchors"] = self._mark()

        runner = SegmentParallelRunner(self.pipe, channel, anchor_shape=I2V_ANCHOR_SHAPE if self.i2v else T2V_ANCHOR_SHAPE,
                                       connect=connect, first_initial=self.first_frame if self.i2v else None, observer=observer)

        def make_noise(seg):
            g = torch.Generator().manual_seed(100 + seg + 1000 * chain)
            return torch.randn(1, FRAMES, 16, LAT_H, LAT_W, generator=g).to(torch.bfloat16).to(dev)

        self.pipe.on_stage = on_stage
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        self.model.launch_count(reset=True)
        e0 = self._mark()
        try:
            outs = runner.run(make_noise, [key], segments_per_chain)
        finally:
            self.pipe.on_stage = None
        e1 = self._mark()
        torch.cuda.synchronize()
        my_ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        worst = torch.tensor([my_ms], device=dev)
        if world > 1:
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        seg_times, connect_ms = [], []
        for seg in sorted(marks):
            m = marks[seg]
            t0 = m.get("connect_end", m["segment_start"])      # the denoise bracket of this segment starts after the connect
            seg_times.append(dict(segment=seg, start_ms=e0.elapsed_time(m["segment_start"]),
                                  anchors_ms=t0.elapsed_time(m["anchors"]) if "anchors" in m else None,
                                  total_ms=t0.elapsed_time(m["segment_end"])))
            if "connect_start" in m:
                connect_ms.append(m["connect_start"].elapsed_time(m["connect_end"]))
        info = dict(rank=rank, chain=chain, slot=crank // self.lanes, lane=crank % self.lanes, segments=sorted(outs),
                    finish_ms=my_ms, launches=self.model.launch_count(), anchor_bytes_sent=channel.bytes_sent,
                    p2p=sum(1 for op in runner.log if op[0] in ("send", "recv") and op[2] != crank),
                    cfg_allgathers=(4 * steps * len(outs) if self.lanes == 2 else 0),
                    cfg_bytes_exchanged_last_segment=self.pipe.cfg_bytes_exchanged,
                    connect_ms=connect_ms, segment_times=seg_times,
                    finite=all(torch.isfinite(v.float()).all().item() for v in outs.values()),
                    checksum={k: float(v.float().abs().sum()) for k, v in outs.items()})
        gathered = [info]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, info)
        if rank != 0 or warm:
            return None
        ms = worst.item()
        frames = chains * segments_per_chain * FRAMES
        full = [t for r in gathered for t in r["segment_times"] if t["anchors_ms"] and t["total_ms"]]
        t_anchor = sum(t["anchors_ms"] for t in full) / max(1, len(full))
        t_segment = sum(t["total_ms"] for t in full) / max(1, len(full))
        flops = chains * (segment_flops(self.dims, steps, self.i2v, True)
                          + (segments_per_chain - 1) * segment_flops(self.dims, steps, self.i2v, False))
        connects = [c for r in gathered for c in r["connect_ms"]]
        lanes_ok = all(a["checksum"] == b["checksum"] for a, b in zip(gathered[0::2], gathered[1::2])) if self.lanes == 2 else True
        return {
            "layout": f"{chains} chain(s) x {slots} slot(s) x {self.lanes} lane(s)", "chains": chains, "slots": slots, "lanes": self.lanes,
            "segments_per_chain": segments_per_chain, "latent_frames": frames, "sampling_steps": steps,
            "value": frames / (ms / 1e3), "unit": "latent frames/s", "ms_total": ms,
            "model_tflops_per_gpu": flops / (ms / 1e3) / 1e12 / world,
            "t_anchor_ms": t_anchor, "t_segment_ms": t_segment,
            "t_anchor_over_t_segment": (t_anchor / t_segment) if t_segment else None,
            "rank_finish_ms": [round(r["finish_ms"], 1) for r in gathered],
            "anchor_bytes": sum(r["anchor_bytes_sent"] for r in gathered),
            "nccl_p2p": sum(r["p2p"] for r in gathered), "cfg_allgathers_per_rank": gathered[0]["cfg_allgathers"],
            "vae_connect_ms": [round(c, 2) for c in connects],
            "connect": "VAE decode -> pixel frames 8:13 -> encode (random-init Wan VAE)" if self.vae is not None
                       else "pass-through of the last two anchors (benchmarking shortcut, NOT the reference transform)",
            "launches_per_rank": gathered[0]["launches"], "finite": all(r["finite"] for r in gathered),
            "cfg_lanes_bit_identical": lanes_ok,
            "segment_times_rank0": gathered[0]["segment_times"],
        }

    def time_final_decode(self) -> Optional[dict]:
        """The decode that ends a segment in the reference (pipeline/casual_fps_inference.py:445: all 21 latent frames -> 81
        pixel frames at 480x832) and the i2v image encode, on this rank's GPU. Outside the denoise bracket the metric is
        defined on, so reported next to the chain numbers, not inside them."""
        if self.vae is None:
            return None
        lat = torch.randn(1, FRAMES, 16, LAT_H, LAT_W, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).to(self.device)
        image = (torch.rand(1, 3, 1, 8 * LAT_H, 8 * LAT_W, generator=torch.Generator().manual_seed(4)) * 2 - 1).to(torch.bfloat16).to(self.device)
        out = {}
        for name, fn in (("final_decode_ms", lambda: self.vae.decode_to_pixel(lat)), ("image_encode_ms", lambda: self.vae.encode_to_latent(image))):
            fn()
            torch.cuda.synchronize()
            e0, e1 = self._mark(), None
            y = fn()
            e1 = self._mark()
            torch.cuda.synchronize()
            out[name] = round(e0.elapsed_time(e1), 2)
            out[name.replace("_ms", "_finite")] = bool(torch.isfinite(y).all())
            del y
        out["note"] = "21 latent -> 81 pixel frames at 480x832, and one image frame -> one latent; outside the denoise bracket"
        return out

    @staticmethod
    def _mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev


def limiter_of(rec: dict) -> str:
    """Names what bounds a variant, from its own timings."""
    r = rec.get("t_anchor_over_t_segment") or 0.0
    if rec["slots"] == 1:
        return "compute (one slot: segments run back to back)" + ("; CFG-pair exchange per step" if rec["lanes"] == 2 else "")
    if rec["slots"] * r > 1.0:
        return (f"anchor-stage dependency: T_anchor/T_segment = {r:.2f}, so a chain feeds at most {1 / r:.1f} slots and this "
                f"layout has {rec['slots']}")
    return f"pipeline fill/drain: {rec['segments_per_chain']} segments on {rec['slots']} slots (T_anchor/T_segment = {r:.2f})"
