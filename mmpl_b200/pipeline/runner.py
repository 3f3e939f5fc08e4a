"""Executes a RolloutPlan (plan.py) against a generator, its caches and a sampler. One loop serves the three reference
pipelines; what differs between them is data: the plan, the guidance branches and the sampler.

  Branch    one guidance branch: a conditioning dict plus the KV / cross-attention caches it owns. The few-step pipeline
            has one, the CFG pipelines two (conditional, unconditional); under the CFG-pair split a rank holds one of the
            two and exchanges flow predictions with its partner.
  Sampler   turns noise into latents by calling `forward(latents, timestep)`:
              FewStepSampler  x0 prediction -> re-noise to the next (warped) timestep   (causal_inference.py:190-216)
              UniPCSampler    CFG combine + UniPC multistep update, one fused launch     (casual_fps_inference.py:337-374)
  Rollout   walks the plan: prefill records, then per stage the re-noising / visibility edits, the sampler run, the
            write-back, the anchor hand-off and the clean-context pass that rewrites the stage's K/V.

RNG: the global torch generator is consumed exactly where the reference consumes it (`randn_like` of the flattened x0 per
non-final few-step call; `randn_like` of the first, then the last frame of a re-noised MMPL stage), which the CPU golden
tests check bit for bit against the reference pipelines.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .plan import Denoise, Prefill, RolloutPlan


@dataclass
class Branch:
    conditioning: dict
    kv_cache: Optional[List[dict]]       # None: the partner rank of a CFG pair owns this branch
    crossattn_cache: Optional[List[dict]]


class FewStepSampler:
    """Few-step distilled rollout: every call predicts x0, all but the last re-noise it to the next timestep."""
    wants = "x0"

    def __init__(self, timesteps: torch.Tensor, scheduler, context_noise):
        self.timesteps, self.scheduler, self.context_noise = timesteps, scheduler, context_noise

    def run(self, latents: torch.Tensor, forward: Callable, on_step: Optional[Callable] = None):
        batch, frames = latents.shape[:2]
        last = len(self.timesteps) - 1
        for i, t in enumerate(self.timesteps):
            timestep = torch.ones([batch, frames], device=latents.device, dtype=torch.int64) * t
            x0 = forward(latents, timestep)[0]
            if on_step is not None:
                on_step(i, x0)
            if i < last:
                flat = x0.flatten(0, 1)
                to = self.timesteps[i + 1] * torch.ones([batch * frames], device=latents.device, dtype=torch.long)
                latents = self.scheduler.add_noise(flat, torch.randn_like(flat), to).unflatten(0, (batch, frames))
        return x0, torch.ones_like(timestep) * self.context_noise


class UniPCSampler:
    """Many-step guided rollout. `stepper(like)` returns an object with `step(flow_cond, flow_uncond, sample)`; the default
    is the fused kernel (mmpl_b200.unipc.FusedUniPC)."""
    wants = "flow"

    def __init__(self, steps: int, shift: float, guidance: float, num_train_timesteps: int = 1000,
                 stepper: Optional[Callable] = None):
        from ..unipc import FusedUniPC, unipc_table
        self.steps, self.shift, self.guidance, self.num_train_timesteps = steps, shift, guidance, num_train_timesteps
        if stepper is None:
            def stepper(like):
                return FusedUniPC(unipc_table(steps, float(shift), float(guidance), num_train_timesteps, "cuda", str(like.device)), like)
        self._stepper = stepper
        self.timesteps = unipc_table(steps, float(shift), float(guidance), num_train_timesteps, "cpu").timesteps

    def run(self, latents: torch.Tensor, forward: Callable, on_step: Optional[Callable] = None):
        batch, frames = latents.shape[:2]
        state = self._stepper(latents)
        # [steps, B, F] float32: the per-frame timestep tensor of every call, built once per stage
        grid = self.timesteps.to(device=latents.device, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, batch, frames).contiguous()
        latents = latents.contiguous()
        for i in range(self.steps):
            flow_cond, flow_uncond = forward(latents, grid[i])
            latents = state.step(flow_cond, flow_uncond, latents)
            if on_step is not None:
                on_step(i, latents)
        return latents, grid[-1] * 0


class Rollout:
    def __init__(self, plan: RolloutPlan, generator, branches: Sequence[Branch], sampler, frame_tokens: int,
                 prefill_dtype: torch.dtype, scheduler=None, renoise_timestep: Optional[torch.Tensor] = None,
                 anchor_sink: Optional[Callable] = None, pair_group=None, on_stage: Optional[Callable] = None):
        self.plan, self.generator, self.branches, self.sampler = plan, generator, list(branches), sampler
        self.fs, self.prefill_dtype, self.scheduler = frame_tokens, prefill_dtype, scheduler
        self.renoise_timestep, self.anchor_sink, self.pair_group, self.on_stage = renoise_timestep, anchor_sink, pair_group, on_stage
        self.bytes_exchanged = 0
        self._pair_buf = {}

    # ------------------------------------------------------------------------------------------------------ model calls
    def _positions(self, rec):
        def tok(v):
            return None if v is None else ([f * self.fs for f in v] if isinstance(v, tuple) else v * self.fs)
        kw = dict(current_start=tok(rec.temporal))
        if rec.slot is not None:
            kw["cache_start"] = tok(rec.slot)
        return kw

    def _forward(self, rec, latents, timestep, want: Optional[str]):
        """One generator call per local branch. Returns per-branch x0 / flow predictions (None for a branch held by the
        partner rank until the exchange fills it in)."""
        pos = self._positions(rec)
        outs = []
        for br in self.branches:
            if br.kv_cache is None:
                outs.append(None)
                continue
            flow, x0 = self.generator(noisy_image_or_video=latents, conditional_dict=br.conditioning, timestep=timestep,
                                      kv_cache=br.kv_cache, crossattn_cache=br.crossattn_cache, **pos)
            outs.append(x0 if want == "x0" else flow)
        if want == "flow" and self.pair_group is not None:
            outs = self._exchange(outs)
        return outs

    def _exchange(self, outs):
        """CFG-pair split: each rank computed one branch's flow; one all-gather gives both ranks both."""
        mine = next(o for o in outs if o is not None).contiguous()
        key = (tuple(mine.shape), mine.device)
        if key not in self._pair_buf:
            self._pair_buf[key] = [torch.empty_like(mine), torch.empty_like(mine)]
        pair = self._pair_buf[key]
        dist.all_gather(pair, mine, group=self.pair_group)
        self.bytes_exchanged += mine.numel() * mine.element_size()
        return pair

    # ------------------------------------------------------------------------------------------------------------ stages
    def _edit_visibility(self, rec: Denoise):
        for br in self.branches:
            for block in br.kv_cache or []:
                vis = block["attention_vis_index"]
                for f in rec.hide:
                    if f * self.fs in vis:
                        vis.remove(f * self.fs)
                for f in rec.show:
                    if f * self.fs not in vis:
                        vis.append(f * self.fs)

    def _renoise(self, rec: Denoise, latents, output):
        batch = latents.shape[0]
        t = self.renoise_timestep.flatten(0, 1).to(latents.device)
        for at, src in rec.renoise:
            eps = torch.randn_like(latents[:, at:at + 1]).flatten(0, 1)
            latents[:, at:at + 1] = self.scheduler.add_noise(output[:, src:src + 1].flatten(0, 1), eps, t).unflatten(0, (batch, 1))

    def run(self, noise: torch.Tensor, initial_latent: Optional[torch.Tensor]) -> torch.Tensor:
        batch, _, channels, height, width = noise.shape
        output = torch.zeros([batch, self.plan.out_frames, channels, height, width], device=noise.device, dtype=noise.dtype)
        for index, rec in enumerate(self.plan.records):
            if isinstance(rec, Prefill):
                clean = initial_latent[:, rec.source[0]:rec.source[1]]
                output[:, list(rec.out)] = clean
                zero = torch.zeros([batch, rec.t_len], device=noise.device, dtype=self.prefill_dtype)
                self._forward(rec, clean, zero, None)
                continue
            latents = noise[:, list(rec.noise)]          # a copy: stages may overwrite frames of it
            if rec.renoise:
                self._renoise(rec, latents, output)
            if rec.hide or rec.show:
                self._edit_visibility(rec)
            latents, clean_t = self.sampler.run(latents, lambda x, t, rec=rec: self._forward(rec, x, t, self.sampler.wants))
            output[:, list(rec.out)] = latents
            if rec.handoff is not None and self.anchor_sink is not None:
                frames, with_stage = rec.handoff
                parts = [output[:, list(frames)]] + ([latents] if with_stage else [])
                self.anchor_sink(torch.cat(parts, dim=1))
            self._forward(rec, latents, clean_t, None)   # rewrite the stage's K/V from the clean latents
            if self.on_stage is not None:
                self.on_stage(index, rec, latents)
        return output
