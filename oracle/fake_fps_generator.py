"""Test infrastructure: a deterministic stand-in for WanFPSWrapper / WanDiffusionWrapper that lets the MMPL macro-from-micro *pipeline*
(pipeline/casual_fps_inference.py in the reference, mmpl_b200/pipeline/causal_fps_inference.py here) run on the CPU at the
full 60x104 latent size in seconds. It records every call (branch, timesteps, frame positions, visibility list before /
after), reproduces the model's bookkeeping of `attention_vis_index` (wan/modules/causal_fps_model.py:209-264: union with
the call's frame starts unless the call is the last stage, which contains frame 15), and returns a flow prediction that
depends on the input latents, the timestep, the branch, the frame positions and the visible set - so a pipeline that
schedules, combines or re-noises anything differently from the reference produces different latents.

With an integer `current_start` (the contiguous-cache pipelines: pipeline/causal_diffusion_inference.py,
pipeline/causal_inference.py) it reproduces the model's index recurrence instead (wan/modules/causal_model.py:203-226:
local_end = local_end_prev + current_start + S - global_end_prev; global_end = current_start + S) on the cache dicts.

Used by oracle/make_golden_fps_pipeline.py (around the *reference* pipeline, to record tests/golden/fps_pipeline_*.pt)
and by tests/test_fps_pipeline_golden.py (around the mirror). Never imported by the product."""
from __future__ import annotations

import hashlib

import torch


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().view(torch.uint8).numpy().tobytes()).hexdigest()[:16]


class _FakeBackbone(torch.nn.Module):
    def __init__(self, num_layers=2, num_heads=2, dim=256, text_len=32):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.num_layers, self.num_heads, self.dim, self.text_len = num_layers, num_heads, dim, text_len
        self.num_frame_per_block = 1
        self.local_attn_size = -1


class FakeFPSGenerator(torch.nn.Module):
    def __init__(self, scheduler, frame_tokens: int = 1560):
        super().__init__()
        self.model = _FakeBackbone()
        self.scheduler = scheduler
        self.fs = frame_tokens
        self.calls = []

    def get_scheduler(self):
        return self.scheduler

    def to(self, *args, **kwargs):  # the reference pipelines move the generator to "cuda" / "cpu" by name
        return self

    def _forward_contiguous(self, x, conditional_dict, timestep, kv_cache, crossattn_cache, current_start, cache_start):
        branch = float(conditional_dict["prompt_embeds"].flatten()[0])
        S = x.shape[1] * self.fs
        before = (int(kv_cache[0]["global_end_index"]), int(kv_cache[0]["local_end_index"]))
        for blk in kv_cache:  # causal_model.py:203-226
            g_prev, l_prev = int(blk["global_end_index"]), int(blk["local_end_index"])
            local_end = l_prev + int(current_start) + S - g_prev
            blk["global_end_index"].fill_(int(current_start) + S)
            blk["local_end_index"].fill_(local_end)
        for blk in crossattn_cache:
            blk["is_init"] = True
        after = (int(kv_cache[0]["global_end_index"]), int(kv_cache[0]["local_end_index"]))
        self.calls.append(dict(branch=branch, timestep=[round(float(v), 4) for v in timestep.flatten()],
                               current_start=int(current_start), cache_start=None if cache_start is None else int(cache_start),
                               end_before=before, end_after=after, frames=int(x.shape[1]), x=digest(x)))
        t = timestep.float().reshape(x.shape[0], -1, 1, 1, 1) / 1000.0
        flow = (0.35 + 0.1 * branch) * x.float() * torch.cos(1.3 * t) + 0.05 * branch * torch.sin(torch.tensor(0.37 * after[1] / self.fs) + t) \
            + 0.002 * (after[1] // self.fs) - 0.1 * t
        flow = flow.to(x.dtype)
        # x0 = xt - sigma * flow the way WanDiffusionWrapper._convert_flow_pred_to_x0 does (fp64, utils/wan_wrapper.py:172-196)
        sched = self.scheduler
        tid = torch.argmin((sched.timesteps.double().unsqueeze(0) - timestep.double().flatten().unsqueeze(1)).abs(), dim=1)
        sigma = sched.sigmas.double()[tid].reshape(x.shape[0], -1, 1, 1, 1)
        x0 = (x.double() - sigma * flow.double()).to(x.dtype)
        return flow, x0

    def forward(self, noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache, current_start=None,
                cache_start=None):
        x = noisy_image_or_video
        if not isinstance(current_start, (list, tuple)):
            return self._forward_contiguous(x, conditional_dict, timestep, kv_cache, crossattn_cache, current_start, cache_start)
        branch = float(conditional_dict["prompt_embeds"].flatten()[0])  # +1 conditional, -1 unconditional
        cur = list(current_start)
        vis_before = sorted(kv_cache[0]["attention_vis_index"])
        for blk in kv_cache:  # causal_fps_model.py:209-264
            if 15 * self.fs not in cur:
                blk["attention_vis_index"] = list(set(blk["attention_vis_index"] + cur))
            else:
                blk["attention_vis_index"] = list(set(blk["attention_vis_index"]))
        for blk in crossattn_cache:
            blk["is_init"] = True
        vis_after = sorted(kv_cache[0]["attention_vis_index"])
        self.calls.append(dict(branch=branch, timestep=[round(float(v), 4) for v in timestep.flatten()],
                               current_start=cur, cache_start=list(cache_start) if cache_start is not None else None,
                               vis_before=vis_before, vis_after=vis_after, frames=int(x.shape[1]), x=digest(x)))
        t = timestep.float().reshape(x.shape[0], x.shape[1], 1, 1, 1) / 1000.0
        pos = torch.tensor([c / self.fs for c in cur], dtype=torch.float32).reshape(1, -1, 1, 1, 1)
        flow = (0.35 + 0.1 * branch) * x.float() * torch.cos(1.3 * t) + 0.05 * branch * torch.sin(0.37 * pos + t) \
            + 0.002 * len(vis_after) - 0.1 * t
        return flow.to(x.dtype), None
