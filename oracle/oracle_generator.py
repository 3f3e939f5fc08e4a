"""ORACLE (test infrastructure, never imported by the product): the oracle's backbone restatements behind the generator
call signature of the reference's WanDiffusionWrapper (utils/wan_wrapper.py:221-292), so that a *pipeline* can be driven
by the oracle instead of by the CUDA model: `OracleGenerator(cfg, weights, fps=False)(noisy_image_or_video, conditional_dict,
timestep, kv_cache, crossattn_cache, current_start[, cache_start]) -> (flow, x0)`.

It speaks the reference's cache-dict protocol (pipeline/causal_inference.py:278-312, casual_fps_inference.py:453-501): K/V
tensors are written in place, `global_end_index` / `local_end_index` are filled, `attention_vis_index` and `is_init` are
replaced - by mapping each dict onto the oracle's KVCache / CrossCache for the duration of a call. Runs on whatever device
the tensors live on (plain torch ops), batch size 1. Used by the GPU parity tests of the pipelines: the same host schedule
once around the CUDA model and once around this object must give the same latents within the stated tolerance."""
from __future__ import annotations

import types

import torch

from . import causal_wan_oracle as O


class OracleScheduler:
    """The few-step schedule the pipelines ask the generator for: tables + add_noise (utils/scheduler.py:106-176)."""

    def __init__(self, shift: float = 5.0):
        self._s = O.FlowMatchSchedule(shift)
        self.sigmas, self.timesteps = self._s.sigmas, self._s.timesteps

    def add_noise(self, original_samples, noise, timestep):
        return self._s.add_noise(original_samples, noise, timestep.float().flatten().to(noise.device))


class OracleGenerator(torch.nn.Module):
    def __init__(self, cfg: O.WanConfig, weights: dict, fps: bool = False, shift: float = 5.0):
        super().__init__()
        self.cfg, self.w, self.fps = cfg, weights, fps
        self.model = types.SimpleNamespace(num_layers=cfg.num_layers, num_heads=cfg.num_heads, dim=cfg.dim, text_len=cfg.text_len,
                                           local_attn_size=-1, num_frame_per_block=1)
        self.scheduler = OracleScheduler(shift)
        self._sched = O.FlowMatchSchedule(shift)
        self._freqs = None
        self.calls = 0

    def get_scheduler(self):
        return self.scheduler

    def to(self, *args, **kwargs):
        return self

    @torch.no_grad()
    def forward(self, noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache, current_start=None,
                cache_start=None):
        x = noisy_image_or_video
        assert x.shape[0] == 1, "the oracle generator handles batch size 1"
        dev = x.device
        if self._freqs is None or self._freqs.device != dev:
            self._freqs = O.rope_freqs(self.cfg.head_dim).to(dev)
            self.w = {k: v.to(dev) for k, v in self.w.items()}
        kv = [O.KVCache(d["k"][0], d["v"][0], int(d["global_end_index"]), int(d["local_end_index"]),
                        list(d.get("attention_vis_index", []))) for d in kv_cache]
        cross = [O.CrossCache(c["k"][0] if c["is_init"] else None, c["v"][0] if c["is_init"] else None, bool(c["is_init"]))
                 for c in crossattn_cache]
        frames = x.shape[1]
        t = timestep.reshape(1, -1).expand(1, frames)[0].to(dev)   # [B, 1] prefill timesteps broadcast over the frames
        context = conditional_dict["prompt_embeds"][0].to(dev)
        lat = x[0].permute(1, 0, 2, 3)
        if self.fps:
            flow = O.fps_model_forward(self.cfg, self.w, lat, t, context, kv, cross, [int(s) for s in current_start], self._freqs)
        else:
            flow = O.model_forward(self.cfg, self.w, lat, t, context, kv, cross, int(current_start), self._freqs)
        flow = flow.permute(1, 0, 2, 3)
        x0 = self._sched.flow_to_x0(flow, x[0], t)
        for d, c in zip(kv_cache, kv):
            d["global_end_index"].fill_(c.global_end_index)
            d["local_end_index"].fill_(c.local_end_index)
            if "attention_vis_index" in d:
                d["attention_vis_index"] = c.visible
        for d, c in zip(crossattn_cache, cross):
            if not d["is_init"]:
                d["k"], d["v"], d["is_init"] = c.k[None], c.v[None], True
        self.calls += 1
        return flow[None], x0[None]
