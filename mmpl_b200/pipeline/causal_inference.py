"""CausalInferencePipeline mirror (pipeline/causal_inference.py:9-312): chunk-wise few-step rollout that owns
`kv_cache1` / `crossattn_cache` with the reference dict layout and keeps the reference `inference()` signature,
call order and RNG consumption (`torch.randn_like` once per non-final denoising step, :208).

Differences from the reference, all host-side: the generator / text encoder / VAE must be injected or built
from `args.model_kwargs` (no checkpoint loading here); the sizes the reference hard-codes for Wan-1.3B
(30 blocks, 12 heads, 1560 tokens per frame; :33-34, :292) are taken from the injected model and the latent
shape instead; per-step prints are behind `verbose`.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..wan_wrapper import WanDiffusionWrapper


class CausalInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        self.generator = WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True) \
            if generator is None else generator
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected: the umT5 encoder and the Wan VAE are outside the "
                             "denoising hot path (SURVEY.md §2)")
        self.text_encoder = text_encoder
        self.vae = vae

        self.scheduler = self.generator.get_scheduler()
        self.denoising_step_list = torch.tensor(args.denoising_step_list, dtype=torch.long)
        if args.warp_denoising_step:
            timesteps = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
            self.denoising_step_list = timesteps[1000 - self.denoising_step_list]

        self.num_transformer_blocks = self.generator.model.num_layers   # reference literal: 30
        self.frame_seq_length = 1560                                     # reference literal; re-derived per call
        self.kv_cache1 = None
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = self.generator.model.local_attn_size
        self.verbose = getattr(args, "verbose", False)
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block
        self.last_profile = None

    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, profile: bool = False, low_memory: bool = False) -> torch.Tensor:
        """pipeline/causal_inference.py:47-276. noise [B, F, C, H, W]; returns video (or (video, latents))."""
        batch_size, num_frames, num_channels, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        if not self.independent_first_frame or (self.independent_first_frame and initial_latent is not None):
            assert num_frames % self.num_frame_per_block == 0
            num_blocks = num_frames // self.num_frame_per_block
        else:
            assert (num_frames - 1) % self.num_frame_per_block == 0
            num_blocks = (num_frames - 1) // self.num_frame_per_block
        num_input_frames = initial_latent.shape[1] if initial_latent is not None else 0
        num_output_frames = num_frames + num_input_frames
        conditional_dict = self.text_encoder(text_prompts=text_prompts)

        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width],
                             device=noise.device, dtype=noise.dtype)
        if profile:
            ev = {k: torch.cuda.Event(enable_timing=True) for k in
                  ("init_start", "init_end", "diffusion_start", "diffusion_end", "vae_start", "vae_end")}
            block_events = []
            ev["init_start"].record()

        # Step 1: KV caches (:112-132)
        if self.kv_cache1 is None or self.kv_cache1[0]["k"].shape[0] != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            for block_index in range(self.num_transformer_blocks):
                self.crossattn_cache[block_index]["is_init"] = False
            for block_index in range(len(self.kv_cache1)):
                self.kv_cache1[block_index]["global_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)
                self.kv_cache1[block_index]["local_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)

        # Step 2: cache the context frames (:134-169)
        current_start_frame = 0
        if initial_latent is not None:
            timestep = torch.ones([batch_size, 1], device=noise.device, dtype=torch.int64) * 0
            if self.independent_first_frame:
                assert (num_input_frames - 1) % self.num_frame_per_block == 0
                num_input_blocks = (num_input_frames - 1) // self.num_frame_per_block
                output[:, :1] = initial_latent[:, :1]
                self.generator(noisy_image_or_video=initial_latent[:, :1], conditional_dict=conditional_dict,
                               timestep=timestep * 0, kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache,
                               current_start=current_start_frame * self.frame_seq_length)
                current_start_frame += 1
            else:
                assert num_input_frames % self.num_frame_per_block == 0
                num_input_blocks = num_input_frames // self.num_frame_per_block
            for _ in range(num_input_blocks):
                current_ref_latents = initial_latent[:, current_start_frame:current_start_frame + self.num_frame_per_block]
                output[:, current_start_frame:current_start_frame + self.num_frame_per_block] = current_ref_latents
                self.generator(noisy_image_or_video=current_ref_latents, conditional_dict=conditional_dict,
                               timestep=timestep * 0, kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache,
                               current_start=current_start_frame * self.frame_seq_length)
                current_start_frame += self.num_frame_per_block

        if profile:
            ev["init_end"].record()
            ev["diffusion_start"].record()

        # Step 3: temporal denoising loop (:177-244)
        all_num_frames = [self.num_frame_per_block] * num_blocks
        if self.independent_first_frame and initial_latent is None:
            all_num_frames = [1] + all_num_frames
        for current_num_frames in all_num_frames:
            if profile:
                bs, be = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                bs.record()
            noisy_input = noise[:, current_start_frame - num_input_frames:
                                current_start_frame + current_num_frames - num_input_frames]
            for index, current_timestep in enumerate(self.denoising_step_list):
                if self.verbose:
                    print(f"current_timestep: {current_timestep}")
                timestep = torch.ones([batch_size, current_num_frames], device=noise.device,
                                      dtype=torch.int64) * current_timestep
                _, denoised_pred = self.generator(
                    noisy_image_or_video=noisy_input, conditional_dict=conditional_dict, timestep=timestep,
                    kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache,
                    current_start=current_start_frame * self.frame_seq_length)
                if index < len(self.denoising_step_list) - 1:
                    next_timestep = self.denoising_step_list[index + 1]
                    noisy_input = self.scheduler.add_noise(
                        denoised_pred.flatten(0, 1), torch.randn_like(denoised_pred.flatten(0, 1)),
                        next_timestep * torch.ones([batch_size * current_num_frames], device=noise.device, dtype=torch.long)
                    ).unflatten(0, denoised_pred.shape[:2])
            # Step 3.2: record the model's output
            output[:, current_start_frame:current_start_frame + current_num_frames] = denoised_pred
            # Step 3.3: rerun at the context timestep to rewrite this chunk's K/V from the clean latent (:227-235)
            context_timestep = torch.ones_like(timestep) * self.args.context_noise
            self.generator(noisy_image_or_video=denoised_pred, conditional_dict=conditional_dict,
                           timestep=context_timestep, kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache,
                           current_start=current_start_frame * self.frame_seq_length)
            if profile:
                be.record()
                block_events.append((bs, be))
            current_start_frame += current_num_frames

        if profile:
            ev["diffusion_end"].record()
            ev["vae_start"].record()

        # Step 4: decode (:255-256)
        video = self.vae.decode_to_pixel(output, use_cache=False)
        video = (video * 0.5 + 0.5).clamp(0, 1)

        if profile:
            ev["vae_end"].record()
            torch.cuda.synchronize()
            self.last_profile = {
                "init_ms": ev["init_start"].elapsed_time(ev["init_end"]),
                "diffusion_ms": ev["diffusion_start"].elapsed_time(ev["diffusion_end"]),
                "vae_ms": ev["vae_start"].elapsed_time(ev["vae_end"]),
                "block_ms": [a.elapsed_time(b) for a, b in block_events],
            }
            if self.verbose:
                print("Profiling results:", self.last_profile)
        if return_latents:
            return video, output
        return video

    def _initialize_kv_cache(self, batch_size, dtype, device):
        """Per-GPU KV cache, reference dict layout (:278-297): list[num_layers] of
        {"k","v": [B, rows, heads, 128], "global_end_index","local_end_index": int64[1]}."""
        model = self.generator.model
        kv_cache_size = self.local_attn_size * self.frame_seq_length if self.local_attn_size != -1 else 32760
        heads, hd = model.num_heads, model.dim // model.num_heads
        self.kv_cache1 = [{
            "k": torch.zeros([batch_size, kv_cache_size, heads, hd], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, kv_cache_size, heads, hd], dtype=dtype, device=device),
            "global_end_index": torch.tensor([0], dtype=torch.long, device=device),
            "local_end_index": torch.tensor([0], dtype=torch.long, device=device),
        } for _ in range(self.num_transformer_blocks)]

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        """Reference layout (:299-312): {"k","v": [B, text_len, heads, 128], "is_init": False}."""
        model = self.generator.model
        heads, hd = model.num_heads, model.dim // model.num_heads
        self.crossattn_cache = [{
            "k": torch.zeros([batch_size, model.text_len, heads, hd], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, model.text_len, heads, hd], dtype=dtype, device=device),
            "is_init": False,
        } for _ in range(self.num_transformer_blocks)]
