"""WanDiffusionWrapper mirror (utils/wan_wrapper.py:116-306) for the KV-cache path.

forward(noisy_image_or_video [B,F,C,H,W], conditional_dict, timestep [B,F], kv_cache, crossattn_cache,
current_start, cache_start) -> (flow_pred, pred_x0), both [B,F,C,H,W] bf16. The flow->x0 conversion
(_convert_flow_pred_to_x0, float64) is fused into the backbone's last kernel; the sigma lookup (argmin over the
1000-entry timestep table) stays in torch on the device, exactly as the reference computes it.
"""
from __future__ import annotations

import types
from typing import List, Optional

import torch
from torch import nn

from .causal_model import CausalFPSWanModel, CausalWanModel
from .scheduler import FlowMatchScheduler

# model_name -> constructor kwargs (wan/configs/wan_t2v_1_3B.py:15-24, wan_t2v_14B.py:15-24)
MODEL_CONFIGS = {
    "Wan2.1-T2V-1.3B": dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30),
    "Wan2.1-T2V-14B": dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40),
}


class WanDiffusionWrapper(nn.Module):
    """The reference loads weights with diffusers' from_pretrained; diffusers and the checkpoints are not part
    of this path, so the backbone is either passed in (`model=`) or built with random init from `model_name`
    and filled by `load_state_dict` (parameter names match the reference)."""

    model_cls = CausalWanModel

    def __init__(self, model_name="Wan2.1-T2V-14B", timestep_shift=8.0, is_causal=True, local_attn_size=-1, sink_size=0,
                 model: Optional[nn.Module] = None):
        super().__init__()
        if not is_causal:
            raise NotImplementedError("only the causal (KV-cache) backbone is implemented")
        if model is None:
            model = self.model_cls(local_attn_size=local_attn_size, sink_size=sink_size, **MODEL_CONFIGS[model_name])
        self.model = model
        self.model.eval()
        self.uniform_timestep = not is_causal
        self.scheduler = FlowMatchScheduler(shift=timestep_shift, sigma_min=0.0, extra_one_step=True)
        self.scheduler.set_timesteps(1000, training=True)
        self.seq_len = 32760
        self._tables64 = None
        self.post_init()

    def _sigma_of(self, timestep: torch.Tensor) -> torch.Tensor:
        """float64 sigma of the schedule entry nearest to each timestep (utils/wan_wrapper.py:186-194); the float64
        tables live on the device after the first call."""
        key = (timestep.device, id(self.scheduler.sigmas))
        if self._tables64 is None or self._tables64[0] != key:
            self._tables64 = (key, self.scheduler.sigmas.double().to(timestep.device),
                              self.scheduler.timesteps.double().to(timestep.device))
        _, sigmas, timesteps = self._tables64
        nearest = (timesteps[None, :] - timestep.reshape(-1, 1)).abs().argmin(dim=1)
        return sigmas[nearest].reshape(timestep.shape)

    def forward(self, noisy_image_or_video: torch.Tensor, conditional_dict: dict, timestep: torch.Tensor,
                kv_cache: Optional[List[dict]] = None, crossattn_cache: Optional[List[dict]] = None,
                current_start=None, classify_mode=False, concat_time_embeddings=False, clean_x=None, aug_t=None,
                cache_start=None):
        if kv_cache is None or classify_mode or clean_x is not None:
            raise NotImplementedError("only the KV-cache inference call is implemented (training paths are out of scope)")
        prompt_embeds = conditional_dict["prompt_embeds"]
        flow_pred = self.model(
            noisy_image_or_video.permute(0, 2, 1, 3, 4), t=timestep, context=prompt_embeds, seq_len=self.seq_len,
            kv_cache=kv_cache, crossattn_cache=crossattn_cache, current_start=current_start, cache_start=cache_start,
            sigma=self._sigma_of(timestep)).permute(0, 2, 1, 3, 4)
        return flow_pred, self.model.last_x0

    def get_scheduler(self):
        return self.scheduler

    def post_init(self):
        self.get_scheduler()


class WanFPSWrapper(WanDiffusionWrapper):
    """utils/wan_wrapper.py:317-493 — same wrapper around CausalFPSWanModel."""
    model_cls = CausalFPSWanModel
