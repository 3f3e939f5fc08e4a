"""Flow-matching noise schedule of the few-step pipeline: the reference's `FlowMatchScheduler` interface
(utils/scheduler.py:106-176) reduced to what inference uses — the shifted sigma table, the timestep -> sigma lookup
and `add_noise` — with `add_noise` on CUDA bf16 latents running the fused `mmpl_add_noise` kernel.

The table is a pure function of (steps, shift, sigma range): `sigma_table()` evaluates the reference's expression with
the same fp32 torch operators, so `sigmas` / `timesteps` carry the reference's bits. Training-only members of the
reference class (loss weights, `training_target`, inverse / reversed schedules) are not part of the inference path and
are rejected rather than carried along.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


def sigma_table(steps: int, shift: float, sigma_min: float, sigma_max: float, extra_one_step: bool,
                denoising_strength: float = 1.0) -> torch.Tensor:
    """sigma_k, k = 0..steps-1, descending from sigma_max: a uniform grid warped by s -> shift*s / (1 + (shift-1)*s)
    (utils/scheduler.py:118-128). `extra_one_step` drops the end point of a (steps+1)-point grid."""
    top = sigma_min + (sigma_max - sigma_min) * denoising_strength
    grid = torch.linspace(top, sigma_min, steps + 1)[:-1] if extra_one_step else torch.linspace(top, sigma_min, steps)
    return shift * grid / (1 + (shift - 1) * grid)


class FlowMatchScheduler:
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, inverse_timesteps=False, extra_one_step=False, reverse_sigmas=False):
        if inverse_timesteps or reverse_sigmas:
            raise NotImplementedError("inverse / reversed sigma schedules are not used by the causal inference path")
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.sigma_max, self.sigma_min, self.extra_one_step = sigma_max, sigma_min, extra_one_step
        self._on: Dict[torch.device, tuple] = {}
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False):
        """`training=True` is accepted (the wrapper passes it, utils/wan_wrapper.py:141) but only asks the reference for
        loss weights, which inference never reads."""
        self.sigmas = sigma_table(num_inference_steps, self.shift, self.sigma_min, self.sigma_max, self.extra_one_step,
                                  denoising_strength)
        self.timesteps = self.sigmas * self.num_train_timesteps
        self._on.clear()

    def tables_on(self, device) -> tuple:
        """(sigmas, timesteps) resident on `device`, uploaded once."""
        device = torch.device(device)
        if self.sigmas.device == device:
            return self.sigmas, self.timesteps
        if device not in self._on:
            self._on[device] = (self.sigmas.to(device), self.timesteps.to(device))
        return self._on[device]

    def sigma_at(self, timestep: torch.Tensor) -> torch.Tensor:
        """sigma of the table entry nearest to each timestep (the reference's argmin lookup, utils/scheduler.py:166-168)."""
        sigmas, timesteps = self.tables_on(timestep.device)
        return sigmas[(timesteps[None, :] - timestep.reshape(-1, 1)).abs().argmin(dim=1)]

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        """bf16((1 - sigma) * x0 + sigma * noise), fp32 inside (utils/scheduler.py:159-176).
        original_samples / noise: [N, C, H, W]; timestep: [N] (or [B, T], flattened)."""
        if not (noise.is_cuda and noise.dtype == torch.bfloat16 and original_samples.dtype == torch.bfloat16):
            raise RuntimeError("mmpl_b200.FlowMatchScheduler.add_noise needs CUDA bfloat16 latents (no CPU fallback)")
        sigma = self.sigma_at(timestep.to(noise.device)).float().contiguous()
        return ops.add_noise(original_samples.contiguous(), noise.contiguous(), sigma)
