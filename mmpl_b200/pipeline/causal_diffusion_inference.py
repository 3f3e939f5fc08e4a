"""CausalDiffusionInferencePipeline mirror (pipeline/causal_diffusion_inference.py:11-378): the
many-step (UniPC, 50 steps) classifier-free-guided chunk-wise rollout on the contiguous-cache `CausalWanModel`, with
separate cond / uncond KV and cross-attention caches (`kv_cache_pos/neg`, `crossattn_cache_pos/neg`).

Same constructor / `inference()` signature, cache dict layouts, call order (cond then uncond, clean-context pass for
both caches after every chunk). Sizes the reference hard-codes for Wan-14B (40 blocks, 40 heads, 1560 tokens per
frame) come from the injected model and the latent shape; device shuffling of the text encoder / VAE is dropped.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..unipc import FlowUniPCMultistepScheduler
from ..wan_wrapper import WanDiffusionWrapper


class CausalDiffusionInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        self.generator = WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True) \
            if generator is None else generator
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected (outside the denoising hot path)")
        self.text_encoder = text_encoder
        self.vae = vae
        self.generator.requires_grad_(False)
        self.num_train_timesteps = args.num_train_timestep
        self.sampling_steps = getattr(args, "sampling_steps", 50)  # reference literal: 50
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        self.num_transformer_blocks = self.generator.model.num_layers  # reference literal: 40
        self.frame_seq_length = 1560
        self.kv_cache_pos = None
        self.kv_cache_neg = None
        self.crossattn_cache_pos = None
        self.crossattn_cache_neg = None
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 3)
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = self.generator.model.local_attn_size
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block

    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, start_frame_index: Optional[int] = 0) -> torch.Tensor:
        batch_size, num_frames, num_channels, height, width = noise.shape
        fs = self.frame_seq_length = (height // 2) * (width // 2)
        if not self.independent_first_frame or (self.independent_first_frame and initial_latent is not None):
            assert num_frames % self.num_frame_per_block == 0
            num_blocks = num_frames // self.num_frame_per_block
        else:
            assert (num_frames - 1) % self.num_frame_per_block == 0
            num_blocks = (num_frames - 1) // self.num_frame_per_block
        num_input_frames = initial_latent.shape[1] if initial_latent is not None else 0
        num_output_frames = num_frames + num_input_frames
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        unconditional_dict = self.text_encoder(text_prompts=[self.args.negative_prompt] * len(text_prompts))
        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)

        # Step 1: caches (:112-138)
        if self.kv_cache_pos is None or self.kv_cache_pos[0]["k"].shape[0] != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            for block_index in range(self.num_transformer_blocks):
                self.crossattn_cache_pos[block_index]["is_init"] = False
                self.crossattn_cache_neg[block_index]["is_init"] = False
            for cache in (self.kv_cache_pos, self.kv_cache_neg):
                for block in cache:
                    block["global_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)
                    block["local_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)

        def both(latents, timestep, current_start_frame, cache_start_frame):
            kw = dict(noisy_image_or_video=latents, timestep=timestep, current_start=current_start_frame * fs,
                      cache_start=cache_start_frame * fs)
            flow_c, _ = self.generator(conditional_dict=conditional_dict, kv_cache=self.kv_cache_pos,
                                       crossattn_cache=self.crossattn_cache_pos, **kw)
            flow_u, _ = self.generator(conditional_dict=unconditional_dict, kv_cache=self.kv_cache_neg,
                                       crossattn_cache=self.crossattn_cache_neg, **kw)
            return flow_c, flow_u

        # Step 2: cache the context frames (:141-208)
        current_start_frame = start_frame_index
        cache_start_frame = 0
        if initial_latent is not None:
            timestep = torch.ones([batch_size, 1], device=noise.device, dtype=torch.int64) * 0
            if self.independent_first_frame:
                assert (num_input_frames - 1) % self.num_frame_per_block == 0
                num_input_blocks = (num_input_frames - 1) // self.num_frame_per_block
                output[:, :1] = initial_latent[:, :1]
                both(initial_latent[:, :1], timestep * 0, current_start_frame, cache_start_frame)
                current_start_frame += 1
                cache_start_frame += 1
            else:
                assert num_input_frames % self.num_frame_per_block == 0
                num_input_blocks = num_input_frames // self.num_frame_per_block
            for _ in range(num_input_blocks):
                ref = initial_latent[:, cache_start_frame:cache_start_frame + self.num_frame_per_block]
                output[:, cache_start_frame:cache_start_frame + self.num_frame_per_block] = ref
                both(ref, timestep * 0, current_start_frame, cache_start_frame)
                current_start_frame += self.num_frame_per_block
                cache_start_frame += self.num_frame_per_block

        # Step 3: temporal denoising loop (:216-297)
        all_num_frames = [self.num_frame_per_block] * num_blocks
        if self.independent_first_frame and initial_latent is None:
            all_num_frames = [1] + all_num_frames
        for current_num_frames in all_num_frames:
            latents = noise[:, cache_start_frame - num_input_frames:
                            cache_start_frame + current_num_frames - num_input_frames]
            sample_scheduler = self._initialize_sample_scheduler(noise)
            for t in sample_scheduler.timesteps:
                timestep = t * torch.ones([batch_size, current_num_frames], device=noise.device, dtype=torch.float32)
                flow_pred_cond, flow_pred_uncond = both(latents, timestep, current_start_frame, cache_start_frame)
                flow_pred = flow_pred_uncond + self.args.guidance_scale * (flow_pred_cond - flow_pred_uncond)
                latents = sample_scheduler.step(flow_pred, t, latents, return_dict=False)[0]
            output[:, cache_start_frame:cache_start_frame + current_num_frames] = latents
            both(latents, timestep * 0, current_start_frame, cache_start_frame)  # clean-context pass (:272-290)
            current_start_frame += current_num_frames
            cache_start_frame += current_num_frames

        video = self.vae.decode_to_pixel(output)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        if return_latents:
            return video, output
        return video

    def _initialize_kv_cache(self, batch_size, dtype, device):
        """:309-343 — reference layout, heads and head_dim from the model."""
        model = self.generator.model
        kv_cache_size = self.local_attn_size * self.frame_seq_length if self.local_attn_size != -1 else 32760
        heads, hd = model.num_heads, model.dim // model.num_heads

        def make():
            return [{
                "k": torch.zeros([batch_size, kv_cache_size, heads, hd], dtype=dtype, device=device),
                "v": torch.zeros([batch_size, kv_cache_size, heads, hd], dtype=dtype, device=device),
                "global_end_index": torch.tensor([0], dtype=torch.long, device=device),
                "local_end_index": torch.tensor([0], dtype=torch.long, device=device),
            } for _ in range(self.num_transformer_blocks)]

        self.kv_cache_pos, self.kv_cache_neg = make(), make()

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        """:345-365."""
        model = self.generator.model
        heads, hd = model.num_heads, model.dim // model.num_heads

        def make():
            return [{
                "k": torch.zeros([batch_size, model.text_len, heads, hd], dtype=dtype, device=device),
                "v": torch.zeros([batch_size, model.text_len, heads, hd], dtype=dtype, device=device),
                "is_init": False,
            } for _ in range(self.num_transformer_blocks)]

        self.crossattn_cache_pos, self.crossattn_cache_neg = make(), make()

    def _initialize_sample_scheduler(self, noise):
        """:367-378 (unipc branch)."""
        sample_scheduler = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1,
                                                       use_dynamic_shifting=False)
        sample_scheduler.set_timesteps(self.sampling_steps, device=noise.device, shift=self.shift)
        self.timesteps = sample_scheduler.timesteps
        return sample_scheduler
