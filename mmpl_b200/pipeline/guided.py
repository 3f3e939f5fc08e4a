"""What the two classifier-free-guided pipelines share: a generator, one KV / cross-attention cache pair per guidance
branch, the UniPC sampler, and an `inference()` that builds a plan, binds the branches and runs the rollout. The fronts
(`CausalDiffusionInferencePipeline`, `CausalFPSInferencePipeline`) say which plan, how large a cache, and what happens at
stage boundaries."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import caches
from .plan import RolloutPlan
from .runner import Branch, Rollout, UniPCSampler


class GuidedPipeline(torch.nn.Module):
    visibility_lists = False        # frame-slot caches carry "attention_vis_index"
    prefill_dtype = torch.int64     # dtype of the all-zero timestep tensor of a prefill call

    def __init__(self, args, generator, text_encoder, vae):
        super().__init__()
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected (outside the denoising hot path)")
        self.generator = generator
        self.generator.requires_grad_(False)
        self.text_encoder, self.vae, self.args = text_encoder, vae, args
        # attribute names the reference drivers read or set
        self.num_train_timesteps = args.num_train_timestep
        self.sampling_steps = getattr(args, "sampling_steps", 50)   # reference literal
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        self.num_transformer_blocks = generator.model.num_layers
        self.independent_first_frame = args.independent_first_frame
        self.frame_seq_length = 1560                                 # re-derived from the latent shape on every call
        self.kv_cache_pos = self.kv_cache_neg = self.crossattn_cache_pos = self.crossattn_cache_neg = None
        self.unipc_stepper = None   # None: the fused kernel; tests inject an eager stand-in on the CPU
        self.on_stage = None        # optional hook(record index, record, latents) after every generated stage
        self.timesteps = None

    # ---- what a front decides ------------------------------------------------------------------------------------
    def make_plan(self, num_frames: int, num_input_frames: int, start_frame_index: int) -> RolloutPlan:
        raise NotImplementedError

    def cache_rows(self) -> int:
        raise NotImplementedError

    def holds(self, branch: int) -> bool:
        """Whether this process owns guidance branch 0 (conditional) / 1 (unconditional)."""
        return True

    def branch_device(self, branch: int, default):
        return default

    def rollout_options(self) -> dict:
        return {}

    # ---- shared machinery ----------------------------------------------------------------------------------------
    def _caches(self):
        return ((self.kv_cache_pos, self.crossattn_cache_pos), (self.kv_cache_neg, self.crossattn_cache_neg))

    def _initialize_kv_cache(self, batch_size, dtype, device):
        made = [caches.new_kv_cache(self.generator.model, batch_size, self.cache_rows(), dtype, self.branch_device(b, device),
                                    visibility=self.visibility_lists) if self.holds(b) else None for b in (0, 1)]
        self.kv_cache_pos, self.kv_cache_neg = made

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        made = [caches.new_cross_cache(self.generator.model, batch_size, dtype, self.branch_device(b, device))
                if self.holds(b) else None for b in (0, 1)]
        self.crossattn_cache_pos, self.crossattn_cache_neg = made

    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, start_frame_index: Optional[int] = 0) -> torch.Tensor:
        batch_size, num_frames, _, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        plan = self.make_plan(num_frames, 0 if initial_latent is None else initial_latent.shape[1], start_frame_index)
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        unconditional_dict = self.text_encoder(text_prompts=[self.args.negative_prompt] * len(text_prompts))
        own = next(kv for b, (kv, _) in enumerate(self._caches()) if self.holds(b))
        if caches.batch_of(own) != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            for b, (kv, cross) in enumerate(self._caches()):
                caches.rewind(kv, cross, self.branch_device(b, noise.device))
        sampler = UniPCSampler(self.sampling_steps, self.shift, self.args.guidance_scale, self.num_train_timesteps,
                               stepper=self.unipc_stepper)
        self.timesteps = sampler.timesteps.to(noise.device)
        branches = [Branch(cond, kv, cross) for cond, (kv, cross) in zip((conditional_dict, unconditional_dict), self._caches())]
        rollout = Rollout(plan, self.generator, branches, sampler, self.frame_seq_length, prefill_dtype=self.prefill_dtype,
                          on_stage=self.on_stage, **self.rollout_options())
        output = rollout.run(noise, initial_latent)
        self.finished(rollout)
        video = (self.vae.decode_to_pixel(output) * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video

    def finished(self, rollout: Rollout) -> None:
        pass
