"""`CausalFPSInferencePipeline` with the reference's constructor and `inference()` signature
(pipeline/casual_fps_inference.py:34-524; i2v variant MMPL_i2v/pipeline/casual_fps_inference.py): the MMPL
macro-from-micro schedule of one 21-frame segment on the frame-slot model.

A front over guided.GuidedPipeline: the stage map, the re-noised boundary frames, the visibility edits and the anchor
hand-off are records of plan.plan_mmpl; the hot loop (50 UniPC steps x {cond, uncond} + CFG combine + scheduler step, then
two t=0 forwards) is runner.UniPCSampler with one fused kernel per step. Two things the reference does not have:

  anchor_sink   after the anchor stage the reference `torch.save`s the hand-off payload for the next segment's thread to
                poll (:380-383). Here it goes to `anchor_sink(payload)` - an NCCL send in segment-parallel runs
                (mmpl_b200/segment_parallel.py), nothing otherwise. No file, no polling.
  cfg_group     CFG-pair split (the reference's device_cond / device_uncond hooks, :42-43,346-367, taken
                to one process per GPU): a process group of two ranks. Rank 0 of the group runs the conditional forwards
                with kv_cache_pos, rank 1 the unconditional ones with kv_cache_neg; the flow predictions ([B, n, 16, H,
                W] bf16, <= 1.4 MB) are exchanged with one all-gather per step and both ranks run the same fused update,
                so their latents stay identical. Both ranks must be constructed and called with the same seeds.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from ..wan_wrapper import WanFPSWrapper
from . import caches
from .guided import GuidedPipeline
from .plan import plan_mmpl


class CausalFPSInferencePipeline(GuidedPipeline):
    visibility_lists = True
    prefill_dtype = torch.float32

    def __init__(self, args, device, generator=None, text_encoder=None, vae=None, device_cond="cuda:0",
                 device_uncond="cuda:0", save="latents_chunk1.pt", anchor_sink: Optional[Callable] = None,
                 cfg_group: Optional["dist.ProcessGroup"] = None):
        if torch.device(device_cond) != torch.device(device_uncond):
            raise NotImplementedError("cond/uncond on different devices: every reference driver passes the same device; "
                                      "use cfg_group to split the branches over two ranks")
        if generator is None:
            generator = WanFPSWrapper(**getattr(args, "model_kwargs", {}), is_causal=True)
        super().__init__(args, generator.to(device_cond), text_encoder, vae)
        self.generator_cond = self.generator                 # the reference's name for it
        self.device_cond, self.device_uncond = device_cond, device_uncond
        self.save, self.need_wait = save, False              # reference attributes; nothing is written to `save`
        self.anchor_sink, self.cfg_group = anchor_sink, cfg_group
        self.cfg_role = None
        if cfg_group is not None:
            if dist.get_world_size(cfg_group) != 2:
                raise ValueError("cfg_group must contain exactly two ranks (conditional, unconditional)")
            self.cfg_role = dist.get_rank(cfg_group)
        self.generator.model.num_frame_per_block = 1
        self.num_frame_per_block, self.local_attn_size = 1, -1
        self.variant = "i2v" if getattr(args, "i2v", False) else "t2v"
        self.cfg_bytes_exchanged = 0
        # noise level of the re-noised stage-boundary frames: ONE draw from the global generator at construction, looked
        # up in the few-step schedule's timestep table and offset by 1000 (:93-108) - so add_noise() resolves it to the
        # table's first entry
        self.ddpm_scheduler = self.generator.get_scheduler()
        self.ddpm_index = torch.randint(980, self.num_train_timesteps, [1, 1], device=self.device_cond, dtype=torch.long)
        self.ddmp_timestep = self.ddpm_scheduler.timesteps.to(self.ddpm_index.device)[self.ddpm_index] + 1000

    def make_plan(self, num_frames, num_input_frames, start_frame_index):
        if self.independent_first_frame:
            raise NotImplementedError("independent_first_frame has no stage in the MMPL stage map")
        return plan_mmpl(self.variant, num_frames, num_input_frames)

    def cache_rows(self):
        return caches.MMPL_SLOTS * self.frame_seq_length

    def holds(self, branch):
        return self.cfg_role in (None, branch)               # under the CFG-pair split a rank holds only its own branch

    def branch_device(self, branch, default):
        return self.device_uncond if branch else self.device_cond

    def rollout_options(self):
        return dict(scheduler=self.ddpm_scheduler, renoise_timestep=self.ddmp_timestep, anchor_sink=self.anchor_sink,
                    pair_group=self.cfg_group)

    def finished(self, rollout):
        self.cfg_bytes_exchanged = rollout.bytes_exchanged
