"""`CausalDiffusionInferencePipeline` with the reference's constructor and `inference()` signature
(pipeline/causal_diffusion_inference.py:11-378): many-step (UniPC), classifier-free-guided chunk-wise
rollout on the contiguous-cache `CausalWanModel`, with one KV / cross-attention cache pair per guidance branch
(`kv_cache_pos/neg`, `crossattn_cache_pos/neg`).

A front over plan.plan_contiguous + runner.Rollout with two branches and the UniPCSampler (one fused kernel per step for
CFG combine + UniPC update). Sizes the reference hard-codes for Wan-14B come from the injected model and the latent
shape; device shuffling of the text encoder / VAE is dropped.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..wan_wrapper import WanDiffusionWrapper
from . import caches
from .plan import plan_contiguous
from .runner import Branch, Rollout, UniPCSampler


class CausalDiffusionInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected (outside the denoising hot path)")
        self.generator = generator if generator is not None else \
            WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True)
        self.generator.requires_grad_(False)
        self.text_encoder, self.vae, self.args = text_encoder, vae, args
        self.num_train_timesteps = args.num_train_timestep
        self.sampling_steps = getattr(args, "sampling_steps", 50)   # reference literal
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        model = self.generator.model
        self.num_transformer_blocks = model.num_layers
        self.local_attn_size = model.local_attn_size
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 3)
        self.independent_first_frame = args.independent_first_frame
        if self.num_frame_per_block > 1:
            model.num_frame_per_block = self.num_frame_per_block
        self.frame_seq_length = 1560
        self.kv_cache_pos = self.kv_cache_neg = self.crossattn_cache_pos = self.crossattn_cache_neg = None
        self.unipc_stepper = None   # None: the fused kernel; tests inject an eager stand-in on the CPU
        self.on_stage = None
        self.timesteps = None

    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, start_frame_index: Optional[int] = 0) -> torch.Tensor:
        batch_size, num_frames, _, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        plan = plan_contiguous(num_frames, 0 if initial_latent is None else initial_latent.shape[1],
                               self.num_frame_per_block, self.independent_first_frame, sampler="unipc",
                               start_frame=start_frame_index, with_slot=True)
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        unconditional_dict = self.text_encoder(text_prompts=[self.args.negative_prompt] * len(text_prompts))
        if caches.batch_of(self.kv_cache_pos) != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            caches.rewind(self.kv_cache_pos, self.crossattn_cache_pos, noise.device)
            caches.rewind(self.kv_cache_neg, self.crossattn_cache_neg, noise.device)
        sampler = UniPCSampler(self.sampling_steps, self.shift, self.args.guidance_scale, self.num_train_timesteps,
                               stepper=self.unipc_stepper)
        self.timesteps = sampler.timesteps.to(noise.device)
        branches = [Branch(conditional_dict, self.kv_cache_pos, self.crossattn_cache_pos),
                    Branch(unconditional_dict, self.kv_cache_neg, self.crossattn_cache_neg)]
        output = Rollout(plan, self.generator, branches, sampler, self.frame_seq_length, prefill_dtype=torch.int64,
                         on_stage=self.on_stage).run(noise, initial_latent)
        video = (self.vae.decode_to_pixel(output) * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video

    def _initialize_kv_cache(self, batch_size, dtype, device):
        rows = caches.CONTIGUOUS_ROWS if self.local_attn_size == -1 else self.local_attn_size * self.frame_seq_length
        model = self.generator.model
        self.kv_cache_pos = caches.new_kv_cache(model, batch_size, rows, dtype, device)
        self.kv_cache_neg = caches.new_kv_cache(model, batch_size, rows, dtype, device)

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        model = self.generator.model
        self.crossattn_cache_pos = caches.new_cross_cache(model, batch_size, dtype, device)
        self.crossattn_cache_neg = caches.new_cross_cache(model, batch_size, dtype, device)
