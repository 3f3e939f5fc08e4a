"""`CausalFPSInferencePipeline` with the reference's constructor and `inference()` signature
(pipeline/casual_fps_inference.py:34-524; i2v variant MMPL_i2v/pipeline/casual_fps_inference.py): the MMPL
macro-from-micro schedule of one 21-frame segment on the frame-slot model.

A front over plan.plan_mmpl + runner.Rollout: the stage map, the re-noised boundary frames, the visibility edits and the
anchor hand-off are records of the plan; the hot loop (50 UniPC steps x {cond, uncond} + CFG combine + scheduler step,
then two t=0 forwards) is runner.UniPCSampler with one fused kernel per step. Two things the reference does not have:

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

from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from ..wan_wrapper import WanFPSWrapper
from . import caches
from .plan import plan_mmpl
from .runner import Branch, Rollout, UniPCSampler


class CausalFPSInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None, device_cond="cuda:0",
                 device_uncond="cuda:0", save="latents_chunk1.pt", anchor_sink: Optional[Callable] = None,
                 cfg_group: Optional["dist.ProcessGroup"] = None):
        super().__init__()
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected (outside the denoising hot path)")
        if torch.device(device_cond) != torch.device(device_uncond):
            raise NotImplementedError("cond/uncond on different devices: every reference driver passes the same device; "
                                      "use cfg_group to split the branches over two ranks")
        self.device_cond, self.device_uncond = device_cond, device_uncond
        self.save, self.need_wait = save, False          # reference attributes; nothing is written to `save`
        self.anchor_sink, self.cfg_group = anchor_sink, cfg_group
        self.cfg_role = None
        if cfg_group is not None:
            if dist.get_world_size(cfg_group) != 2:
                raise ValueError("cfg_group must contain exactly two ranks (conditional, unconditional)")
            self.cfg_role = dist.get_rank(cfg_group)
        self.generator_cond = generator if generator is not None else \
            WanFPSWrapper(**getattr(args, "model_kwargs", {}), is_causal=True)
        self.generator_cond.requires_grad_(False)
        self.generator_cond = self.generator_cond.to(self.device_cond)
        model = self.generator_cond.model
        model.num_frame_per_block = 1
        self.text_encoder, self.vae, self.args = text_encoder, vae, args
        self.num_train_timesteps = args.num_train_timestep
        self.sampling_steps = getattr(args, "sampling_steps", 50)   # reference literal
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        self.num_transformer_blocks = model.num_layers
        self.num_frame_per_block = 1
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = -1
        self.frame_seq_length = 1560
        self.variant = "i2v" if getattr(args, "i2v", False) else "t2v"
        self.kv_cache_pos = self.kv_cache_neg = self.crossattn_cache_pos = self.crossattn_cache_neg = None
        self.unipc_stepper = None   # None: the fused kernel; tests inject an eager stand-in on the CPU
        self.on_stage = None
        self.timesteps = None
        self.cfg_bytes_exchanged = 0
        # noise level of the re-noised stage-boundary frames: ONE draw from the global generator at construction, looked
        # up in the few-step schedule's timestep table and offset by 1000 (:93-108) - so add_noise() resolves it to the
        # table's first entry
        self.ddpm_scheduler = self.generator_cond.get_scheduler()
        self.ddpm_index = torch.randint(980, self.num_train_timesteps, [1, 1], device=self.device_cond, dtype=torch.long)
        table = self.ddpm_scheduler.timesteps.to(self.ddpm_index.device)
        self.ddmp_timestep = table[self.ddpm_index] + 1000

    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, start_frame_index: Optional[int] = 0) -> torch.Tensor:
        batch_size, num_frames, _, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        if self.independent_first_frame:
            raise NotImplementedError("independent_first_frame has no stage in the MMPL stage map")
        plan = plan_mmpl(self.variant, num_frames, 0 if initial_latent is None else initial_latent.shape[1])
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        unconditional_dict = self.text_encoder(text_prompts=[self.args.negative_prompt] * len(text_prompts))
        own = self.kv_cache_neg if self.cfg_role == 1 else self.kv_cache_pos
        if caches.batch_of(own) != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            caches.rewind(self.kv_cache_pos, self.crossattn_cache_pos, self.device_cond)
            caches.rewind(self.kv_cache_neg, self.crossattn_cache_neg, self.device_uncond)
        sampler = UniPCSampler(self.sampling_steps, self.shift, self.args.guidance_scale, self.num_train_timesteps,
                               stepper=self.unipc_stepper)
        self.timesteps = sampler.timesteps.to(noise.device)
        branches = [Branch(conditional_dict, self.kv_cache_pos, self.crossattn_cache_pos),
                    Branch(unconditional_dict, self.kv_cache_neg, self.crossattn_cache_neg)]
        rollout = Rollout(plan, self.generator_cond, branches, sampler, self.frame_seq_length, prefill_dtype=torch.float32,
                          scheduler=self.ddpm_scheduler, renoise_timestep=self.ddmp_timestep, anchor_sink=self.anchor_sink,
                          pair_group=self.cfg_group, on_stage=self.on_stage)
        output = rollout.run(noise, initial_latent)
        self.cfg_bytes_exchanged = rollout.bytes_exchanged
        video = (self.vae.decode_to_pixel(output) * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video

    def _initialize_kv_cache(self, batch_size, dtype, device):
        """15 frame slots per branch; under the CFG-pair split a rank holds only its own branch's cache."""
        model, rows = self.generator_cond.model, caches.MMPL_SLOTS * self.frame_seq_length
        self.kv_cache_pos = caches.new_kv_cache(model, batch_size, rows, dtype, self.device_cond, visibility=True) \
            if self.cfg_role in (None, 0) else None
        self.kv_cache_neg = caches.new_kv_cache(model, batch_size, rows, dtype, self.device_uncond, visibility=True) \
            if self.cfg_role in (None, 1) else None

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        model = self.generator_cond.model
        self.crossattn_cache_pos = caches.new_cross_cache(model, batch_size, dtype, self.device_cond) \
            if self.cfg_role in (None, 0) else None
        self.crossattn_cache_neg = caches.new_cross_cache(model, batch_size, dtype, self.device_uncond) \
            if self.cfg_role in (None, 1) else None
