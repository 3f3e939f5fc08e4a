"""CausalFPSInferencePipeline mirror (pipeline/casual_fps_inference.py:34-524; i2v variant
MMPL_i2v/pipeline/casual_fps_inference.py): the MMPL macro-from-micro schedule for one 21-frame segment.

  stage map  clean_steps = [0,0,1,1,2,2,2,2,2,2,1,1,1,3,3,3,3,3,3,1,1]  ->  stages
             [0,1] | [2,3,10,11,12,19,20] (anchors) | [4..9] | [13..18]           (t2v, :250-252)
  per stage  50 UniPC steps x {cond forward, uncond forward, CFG combine, scheduler step}, then two t=0 forwards that
             rewrite the stage's K/V from the clean latents; stage 2/3 re-noise their first/last frame from
             already generated neighbours and edit `attention_vis_index` (:281-325).
  hand-off   after the anchor stage the reference `torch.save`s cat(output[:, :1], latents) for the next segment
             (:380-383). Here the payload goes to `anchor_sink(payload)` — an NCCL send in segment-parallel runs
             (mmpl_b200/segment_parallel.py), a no-op / capture otherwise. No file, no polling.

Same constructor and `inference()` signature, same cache dict layouts (`kv_cache_pos/neg` with
"attention_vis_index", `crossattn_cache_pos/neg`), same RNG consumption order. Sizes the reference hard-codes for
Wan-14B at 480x832 (40 blocks, 40 heads, 1560 tokens per frame, 15-slot cache) are derived from the injected model
and the latent shape.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from ..unipc import FlowUniPCMultistepScheduler
from ..wan_wrapper import WanFPSWrapper

T2V_CLEAN_STEPS = [0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 1, 1, 1, 3, 3, 3, 3, 3, 3, 1, 1]
T2V_STAGE_FRAMES = [2, 7, 6, 6]
# MMPL_i2v/pipeline/casual_fps_inference.py:253-255
I2V_CLEAN_STEPS = [0, 1, 2, 2, 3, 3, 3, 3, 3, 3, 2, 2, 2, 4, 4, 4, 4, 4, 4, 2, 2]
I2V_STAGE_FRAMES = [1, 1, 7, 6, 6]


class CausalFPSInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None, device_cond="cuda:0",
                 device_uncond="cuda:0", save="latents_chunk1.pt", anchor_sink: Optional[Callable] = None,
                 cfg_group: Optional["dist.ProcessGroup"] = None):
        super().__init__()
        self.need_wait = False
        self.save = save
        self.anchor_sink = anchor_sink
        # CFG-pair split (the reference's device_cond / device_uncond hooks, :42-43,346-367, taken to one
        # process per GPU): `cfg_group` is a process group of exactly two ranks. Rank 0 of the group runs the
        # conditional forwards with kv_cache_pos, rank 1 the unconditional ones with kv_cache_neg; after every
        # denoising forward the two flow predictions ([B, n, 16, H, W] bf16, <= 1.4 MB) are exchanged with one
        # all-gather and both ranks apply the same CFG combine and UniPC update, so their latents stay identical.
        # Both ranks must be constructed and called with the same seeds (the pipeline draws noise with torch.randn_like).
        self.cfg_group = cfg_group
        self.cfg_role = None
        if cfg_group is not None:
            if dist.get_world_size(cfg_group) != 2:
                raise ValueError("cfg_group must contain exactly two ranks (conditional, unconditional)")
            self.cfg_role = dist.get_rank(cfg_group)
        self.device_cond = device_cond
        self.device_uncond = device_uncond
        if torch.device(device_cond) != torch.device(device_uncond):
            raise NotImplementedError("cond/uncond on different devices: every reference driver passes the same device")
        self.generator_cond = WanFPSWrapper(**getattr(args, "model_kwargs", {}), is_causal=True) \
            if generator is None else generator
        self.generator_cond.requires_grad_(False)
        self.generator_cond = self.generator_cond.to(self.device_cond)
        self.generator_cond.model.num_frame_per_block = 1
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be injected (outside the denoising hot path)")
        self.text_encoder = text_encoder
        self.vae = vae

        self.num_train_timesteps = args.num_train_timestep
        self.sampling_steps = getattr(args, "sampling_steps", 50)          # reference literal: 50
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        self.num_transformer_blocks = self.generator_cond.model.num_layers  # reference literal: 40
        self.frame_seq_length = 1560
        self.kv_cache_pos = None
        self.kv_cache_neg = None
        self.crossattn_cache_pos = None
        self.crossattn_cache_neg = None
        self.args = args
        self.num_frame_per_block = 1
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = -1
        self.verbose = getattr(args, "verbose", False)
        # "t2v" = MMPL_t2v schedule; "i2v" = MMPL_i2v schedule (first frame / segment-connect frames prefilled at t=0,
        # no anchor re-noising, hand-off payload = frames 0, 19, 20)
        self.variant = "i2v" if getattr(args, "i2v", False) else "t2v"

        # re-noising timestep for the stage-boundary frames (:93-108): one torch.randint draw at construction
        self.ddpm_scheduler = self.generator_cond.get_scheduler()
        self.image_or_video_shape = [1, 1, 16, 60, 104]
        self.ddpm_index = self._get_timestep(980, self.num_train_timesteps, self.image_or_video_shape[0],
                                             self.image_or_video_shape[1], 1, uniform_timestep=False)
        self.ddpm_batch_size, self.ddpm_num_frame = self.image_or_video_shape[:2]
        self.ddpm_scheduler.timesteps = self.ddpm_scheduler.timesteps.to(self.ddpm_index.device)
        self.ddmp_timestep = self.ddpm_scheduler.timesteps[self.ddpm_index]
        self.ddmp_timestep = self.ddmp_timestep + 1000

    def _get_timestep(self, min_timestep, max_timestep, batch_size, num_frame, num_frame_per_block,
                      uniform_timestep=False):
        """:111-153."""
        if uniform_timestep:
            return torch.randint(min_timestep, max_timestep, [batch_size, 1], device=self.device_cond,
                                 dtype=torch.long).repeat(1, num_frame)
        timestep = torch.randint(min_timestep, max_timestep, [batch_size, num_frame], device=self.device_cond,
                                 dtype=torch.long)
        if not self.independent_first_frame:
            timestep = timestep.reshape(timestep.shape[0], -1, num_frame_per_block)
            timestep[:, :, 1:] = timestep[:, :, 0:1]
            timestep = timestep.reshape(timestep.shape[0], -1)
        return timestep

    # ------------------------------------------------------------------------------------------------ inference
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, start_frame_index: Optional[int] = 0) -> torch.Tensor:
        batch_size, num_frames, num_channels, height, width = noise.shape
        fs = self.frame_seq_length = (height // 2) * (width // 2)
        assert num_frames % self.num_frame_per_block == 0
        num_output_frames = num_frames
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        unconditional_dict = self.text_encoder(text_prompts=[self.args.negative_prompt] * len(text_prompts))
        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)

        self.cfg_bytes_exchanged = 0
        own = self.kv_cache_neg if self.cfg_role == 1 else self.kv_cache_pos
        if own is None or own[0]["k"].shape[0] != batch_size:
            self._initialize_kv_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
            self._initialize_crossattn_cache(batch_size=batch_size, dtype=noise.dtype, device=noise.device)
        else:
            for cross in (self.crossattn_cache_pos, self.crossattn_cache_neg):
                for block in cross or []:
                    block["is_init"] = False
            for cache, dev in ((self.kv_cache_pos, self.device_cond), (self.kv_cache_neg, self.device_uncond)):
                for block in cache or []:
                    block["global_end_index"] = torch.tensor([0], dtype=torch.long, device=dev)
                    block["local_end_index"] = torch.tensor([0], dtype=torch.long, device=dev)
                    block["attention_vis_index"] = []

        i2v = self.variant == "i2v"
        clean_steps = I2V_CLEAN_STEPS if i2v else T2V_CLEAN_STEPS
        all_num_frames = list(I2V_STAGE_FRAMES if i2v else T2V_STAGE_FRAMES)
        result = [[i for i, v in enumerate(clean_steps) if v == target] for target in range(len(all_num_frames))]
        global_chunk_index = 0
        if self.independent_first_frame and initial_latent is None:
            all_num_frames = [1] + all_num_frames

        def both(latents, timestep, frames, need_flow=True):
            cur = [i * fs for i in frames]
            if self.cfg_group is not None:
                cond = self.cfg_role == 0
                flow, _ = self.generator_cond(
                    noisy_image_or_video=latents, conditional_dict=conditional_dict if cond else unconditional_dict,
                    timestep=timestep, kv_cache=self.kv_cache_pos if cond else self.kv_cache_neg,
                    crossattn_cache=self.crossattn_cache_pos if cond else self.crossattn_cache_neg,
                    current_start=cur, cache_start=cur)
                if not need_flow:
                    return None, None  # clean-context pass: only the K/V written into this rank's cache matter
                pair = [torch.empty_like(flow), torch.empty_like(flow)]
                dist.all_gather(pair, flow.contiguous(), group=self.cfg_group)
                self.cfg_bytes_exchanged += flow.numel() * flow.element_size()
                return pair[0], pair[1]
            flow_c, _ = self.generator_cond(noisy_image_or_video=latents, conditional_dict=conditional_dict,
                                            timestep=timestep, kv_cache=self.kv_cache_pos,
                                            crossattn_cache=self.crossattn_cache_pos, current_start=cur, cache_start=cur)
            flow_u, _ = self.generator_cond(noisy_image_or_video=latents, conditional_dict=unconditional_dict,
                                            timestep=timestep, kv_cache=self.kv_cache_neg,
                                            crossattn_cache=self.crossattn_cache_neg, current_start=cur, cache_start=cur)
            return flow_c, flow_u

        for current_num_frames in all_num_frames:
            if (initial_latent is None) or (global_chunk_index != 0):
                current_start_frame = result[global_chunk_index]
                latents = noise[:, result[global_chunk_index]]
                if i2v and current_num_frames != latents.shape[1]:
                    continue  # MMPL_i2v :277 (the entry skipped after a two-frame prefill)
                target_values = [20 * fs, 19 * fs]  # reference literals 31200, 29640
                if not i2v and global_chunk_index in (2, 3):
                    # re-noise the stage's first / last frame from generated neighbours (:284-296, :306-318)
                    first_src, last_src = (3, 10) if global_chunk_index == 2 else (12, 19)
                    latents[:, 0:1] = self.ddpm_scheduler.add_noise(
                        output[:, first_src:first_src + 1].flatten(0, 1),
                        torch.randn_like(latents[:, 0:1]).flatten(0, 1),
                        self.ddmp_timestep.flatten(0, 1).to(latents.device)).unflatten(0, (batch_size, 1))
                    latents[:, -1:] = self.ddpm_scheduler.add_noise(
                        output[:, last_src:last_src + 1].flatten(0, 1),
                        torch.randn_like(latents[:, -1:]).flatten(0, 1),
                        self.ddmp_timestep.flatten(0, 1).to(latents.device)).unflatten(0, (batch_size, 1))
                    for cache in (self.kv_cache_pos, self.kv_cache_neg):
                        for block in cache or []:
                            for val in target_values:
                                present = val in block["attention_vis_index"]
                                if global_chunk_index == 2 and present:
                                    block["attention_vis_index"].remove(val)      # hide the far anchors (:298-303)
                                elif global_chunk_index == 3 and not present:
                                    block["attention_vis_index"].append(val)      # show them again (:320-325)

                # Step 3.1: spatial denoising loop (:337-374)
                sample_scheduler = self._initialize_sample_scheduler(noise)
                for t in sample_scheduler.timesteps:
                    timestep = t * torch.ones([batch_size, current_num_frames], device=noise.device, dtype=torch.float32)
                    flow_pred_cond, flow_pred_uncond = both(latents, timestep, current_start_frame)
                    flow_pred = flow_pred_uncond + self.args.guidance_scale * (flow_pred_cond - flow_pred_uncond)
                    latents = sample_scheduler.step(flow_pred, t, latents, return_dict=False)[0]

                # Step 3.2: record the stage output (:377-378)
                output[:, result[global_chunk_index]] = latents
                if global_chunk_index == (2 if i2v else 1):
                    # anchors for the next segment (:380-383; MMPL_i2v :340-342)
                    save_latents = torch.cat([output[:, :1], output[:, -2:]], dim=1) if i2v \
                        else torch.cat([output[:, :1], latents], dim=1)
                    if self.anchor_sink is not None:
                        self.anchor_sink(save_latents)
                # Step 3.3: clean-context pass at t=0 for both caches (:386-403)
                both(latents, timestep * 0, current_start_frame, need_flow=False)
                global_chunk_index += 1
            else:
                # prefill with the given first frame(s) instead of generating stage 0 (:407-439; MMPL_i2v :368-435)
                timestep = 0 * torch.ones([batch_size, current_num_frames], device=noise.device, dtype=torch.float32)
                if i2v and initial_latent.shape[1] > 1:
                    for step in range(2):  # "segment connect": two frames, one slot each
                        both(initial_latent[:, step:step + 1], timestep * 0, result[step], need_flow=False)
                        output[:, result[step]] = initial_latent[:, step:step + 1]
                    global_chunk_index = 2
                elif i2v:
                    both(initial_latent[:, 0:1], timestep * 0, result[0], need_flow=False)
                    output[:, result[0]] = initial_latent
                    global_chunk_index += 1
                else:
                    both(initial_latent, timestep * 0, result[0], need_flow=False)
                    output[:, result[global_chunk_index]] = initial_latent
                    global_chunk_index += 1

        video = self.vae.decode_to_pixel(output)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        if return_latents:
            return video, output
        return video

    # ------------------------------------------------------------------------------------------------- caches
    def _initialize_kv_cache(self, batch_size, dtype, device):
        """:453-482 — 15 frame slots (reference literal 32760 - 6*1560 rows), plus "attention_vis_index"."""
        model = self.generator_cond.model
        kv_cache_size = 15 * self.frame_seq_length
        heads, hd = model.num_heads, model.dim // model.num_heads

        def make(dev):
            return [{
                "k": torch.zeros([batch_size, kv_cache_size, heads, hd], dtype=dtype, device=dev),
                "v": torch.zeros([batch_size, kv_cache_size, heads, hd], dtype=dtype, device=dev),
                "global_end_index": torch.tensor([0], dtype=torch.long, device=dev),
                "local_end_index": torch.tensor([0], dtype=torch.long, device=dev),
                "attention_vis_index": [],
            } for _ in range(self.num_transformer_blocks)]

        # under the CFG-pair split a rank holds only the cache of its own branch
        self.kv_cache_pos = make(self.device_cond) if self.cfg_role in (None, 0) else None
        self.kv_cache_neg = make(self.device_uncond) if self.cfg_role in (None, 1) else None

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        """:484-501."""
        model = self.generator_cond.model
        heads, hd = model.num_heads, model.dim // model.num_heads

        def make(dev):
            return [{
                "k": torch.zeros([batch_size, model.text_len, heads, hd], dtype=dtype, device=dev),
                "v": torch.zeros([batch_size, model.text_len, heads, hd], dtype=dtype, device=dev),
                "is_init": False,
            } for _ in range(self.num_transformer_blocks)]

        self.crossattn_cache_pos = make(self.device_cond) if self.cfg_role in (None, 0) else None
        self.crossattn_cache_neg = make(self.device_uncond) if self.cfg_role in (None, 1) else None

    def _initialize_sample_scheduler(self, noise):
        """:503-512 (unipc branch)."""
        sample_scheduler = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1,
                                                       use_dynamic_shifting=False)
        sample_scheduler.set_timesteps(self.sampling_steps, device=noise.device, shift=self.shift)
        self.timesteps = sample_scheduler.timesteps
        return sample_scheduler
