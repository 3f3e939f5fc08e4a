"""`CausalInferencePipeline` with the reference's constructor and `inference()` signature
(pipeline/causal_inference.py:9-312): the few-step chunk-wise rollout that owns `kv_cache1` / `crossattn_cache`.

The class is a front: `inference()` turns the request into a plan (plan.plan_contiguous), binds the caches to one
guidance branch and lets runner.Rollout execute it with the FewStepSampler. Differences from the reference, all
host-side: generator / text encoder / VAE are injected (no checkpoint loading on this path); the sizes the reference
hard-codes for Wan-1.3B (30 blocks, 12 heads, 1560 tokens per frame) come from the injected model and the latent shape.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..wan_wrapper import WanDiffusionWrapper
from . import caches
from .plan import plan_contiguous
from .runner import Branch, FewStepSampler, Rollout


def few_step_timesteps(scheduler, step_list, warp: bool) -> torch.Tensor:
    """The denoising timesteps of the distilled model: the configured list, or - warped - the schedule's own timestep at
    each listed index counted from the end of the 1000-entry table extended by a final 0 (causal_inference.py:26-31)."""
    steps = torch.tensor(step_list, dtype=torch.long)
    if not warp:
        return steps
    table = torch.cat((scheduler.timesteps.cpu(), torch.zeros(1, dtype=torch.float32)))
    return table[len(table) - 1 - steps]


class CausalInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected: the umT5 encoder and the Wan VAE are outside the "
                             "denoising hot path (SURVEY.md §2)")
        self.generator = generator if generator is not None else \
            WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True)
        self.text_encoder, self.vae, self.args = text_encoder, vae, args
        self.scheduler = self.generator.get_scheduler()
        self.denoising_step_list = few_step_timesteps(self.scheduler, args.denoising_step_list, args.warp_denoising_step)
        model = self.generator.model
        self.num_transformer_blocks = model.num_layers
        self.local_attn_size = model.local_attn_size
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = args.independent_first_frame
        if self.num_frame_per_block > 1:
            model.num_frame_per_block = self.num_frame_per_block
        self.frame_seq_length = 1560          # re-derived from the latent shape on every call
        self.kv_cache1 = self.crossattn_cache = None
        self.verbose = getattr(args, "verbose", False)
        self.last_profile = None
        self.on_call = None                   # optional hook(step index, x0) after every denoising call (tests, tracing)

    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, profile: bool = False, low_memory: bool = False) -> torch.Tensor:
        """noise [B, F, C, H, W] -> video (or (video, latents)); `initial_latent` frames are prefilled and prepended."""
        batch_size, num_frames, _, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        plan = plan_contiguous(num_frames, 0 if initial_latent is None else initial_latent.shape[1],
                               self.num_frame_per_block, self.independent_first_frame, sampler="fewstep")
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        marks = _Marks(profile)
        marks("init_start")
        if caches.batch_of(self.kv_cache1) != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            caches.rewind(self.kv_cache1, self.crossattn_cache, noise.device)
        sampler = FewStepSampler(self.denoising_step_list, self.scheduler, self.args.context_noise)
        if self.on_call is not None or self.verbose:
            run = sampler.run
            sampler.run = lambda latents, forward: run(latents, forward, self._step_hook)
        rollout = Rollout(plan, self.generator, [Branch(conditional_dict, self.kv_cache1, self.crossattn_cache)], sampler,
                          self.frame_seq_length, prefill_dtype=torch.int64, scheduler=self.scheduler,
                          on_stage=marks.stage if profile else None)
        marks("init_end")
        marks("diffusion_start")
        output = rollout.run(noise, initial_latent)
        marks("diffusion_end")
        marks("vae_start")
        video = (self.vae.decode_to_pixel(output, use_cache=False) * 0.5 + 0.5).clamp(0, 1)
        marks("vae_end")
        if profile:
            self.last_profile = marks.report()
            if self.verbose:
                print("Profiling results:", self.last_profile)
        return (video, output) if return_latents else video

    def _step_hook(self, index, x0):
        if self.verbose:
            print(f"current_timestep: {self.denoising_step_list[index]}")
        if self.on_call is not None:
            self.on_call(index, x0)

    def _initialize_kv_cache(self, batch_size, dtype, device):
        rows = caches.CONTIGUOUS_ROWS if self.local_attn_size == -1 else self.local_attn_size * self.frame_seq_length
        self.kv_cache1 = caches.new_kv_cache(self.generator.model, batch_size, rows, dtype, device)

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        self.crossattn_cache = caches.new_cross_cache(self.generator.model, batch_size, dtype, device)


class _Marks:
    """CUDA-event brackets of `inference(profile=True)`: cache set-up, the denoise loop (the metric's bracket, SURVEY.md
    §8d), the VAE, and one span per generated chunk (causal_inference.py:99-109,171-174,237-271)."""

    def __init__(self, enabled: bool):
        self.enabled, self.events, self.chunk_ends = enabled, {}, []

    def __call__(self, name: str):
        if self.enabled:
            self.events[name] = torch.cuda.Event(enable_timing=True)
            self.events[name].record()

    def stage(self, index, record, latents):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.chunk_ends.append(ev)

    def report(self) -> dict:
        torch.cuda.synchronize()
        ms = lambda a, b: self.events[a].elapsed_time(self.events[b])
        edges = [self.events["diffusion_start"]] + self.chunk_ends
        return {"init_ms": ms("init_start", "init_end"), "diffusion_ms": ms("diffusion_start", "diffusion_end"),
                "vae_ms": ms("vae_start", "vae_end"),
                "block_ms": [a.elapsed_time(b) for a, b in zip(edges[:-1], edges[1:])]}
