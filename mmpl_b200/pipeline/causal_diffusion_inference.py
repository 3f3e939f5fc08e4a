"""`CausalDiffusionInferencePipeline` with the reference's constructor and `inference()` signature
(pipeline/causal_diffusion_inference.py:11-378): many-step (UniPC), classifier-free-guided chunk-wise
rollout on the contiguous-cache `CausalWanModel`, one KV / cross-attention cache pair per guidance branch
(`kv_cache_pos/neg`, `crossattn_cache_pos/neg`).

A front over guided.GuidedPipeline: the plan is plan.plan_contiguous (chunks of `num_frame_per_block` frames, optional
prefill, `start_frame_index` offsetting the temporal positions), the cache the reference's 32760 rows. Sizes the reference
hard-codes for Wan-14B come from the injected model and the latent shape; device shuffling of the text encoder / VAE is
dropped.
"""
from __future__ import annotations

from ..wan_wrapper import WanDiffusionWrapper
from . import caches
from .guided import GuidedPipeline
from .plan import plan_contiguous


class CausalDiffusionInferencePipeline(GuidedPipeline):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        if generator is None:
            generator = WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True)
        super().__init__(args, generator, text_encoder, vae)
        self.local_attn_size = generator.model.local_attn_size
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 3)
        if self.num_frame_per_block > 1:
            generator.model.num_frame_per_block = self.num_frame_per_block

    def make_plan(self, num_frames, num_input_frames, start_frame_index):
        return plan_contiguous(num_frames, num_input_frames, self.num_frame_per_block, self.independent_first_frame,
                               sampler="unipc", start_frame=start_frame_index, with_slot=True)

    def cache_rows(self):
        return caches.CONTIGUOUS_ROWS if self.local_attn_size == -1 else self.local_attn_size * self.frame_seq_length
