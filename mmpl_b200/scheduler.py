"""FlowMatchScheduler mirror (utils/scheduler.py:106-194): shifted sigma schedule, add_noise, step.

The schedule itself is host-side torch arithmetic identical to the reference (same ops, same dtype, so the
sigma/timestep tables are bit-identical); `add_noise` on CUDA bf16 latents runs the fused mmpl_add_noise kernel.
"""
from __future__ import annotations

import torch

from . import ops


class FlowMatchScheduler:
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, inverse_timesteps=False, extra_one_step=False, reverse_sigmas=False):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        self.sigma_max = sigma_max
        self.sigma_min = sigma_min
        self.inverse_timesteps = inverse_timesteps
        self.extra_one_step = extra_one_step
        self.reverse_sigmas = reverse_sigmas
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False):
        """utils/scheduler.py:118-142."""
        sigma_start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps)
        if self.inverse_timesteps:
            self.sigmas = torch.flip(self.sigmas, dims=[0])
        self.sigmas = self.shift * self.sigmas / (1 + (self.shift - 1) * self.sigmas)
        if self.reverse_sigmas:
            self.sigmas = 1 - self.sigmas
        self.timesteps = self.sigmas * self.num_train_timesteps
        if training:
            x = self.timesteps
            y = torch.exp(-2 * ((x - num_inference_steps / 2) / num_inference_steps) ** 2)
            y_shifted = y - y.min()
            self.linear_timesteps_weights = y_shifted * (num_inference_steps / y_shifted.sum())

    def _to(self, device):
        self.sigmas = self.sigmas.to(device)
        self.timesteps = self.timesteps.to(device)

    def timestep_id(self, timestep):
        return torch.argmin((self.timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)

    def step(self, model_output, timestep, sample, to_final=False):
        """utils/scheduler.py:144-157."""
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        self._to(model_output.device)
        timestep_id = self.timestep_id(timestep)
        sigma = self.sigmas[timestep_id].reshape(-1, 1, 1, 1)
        if to_final or (timestep_id + 1 >= len(self.timesteps)).any():
            sigma_ = 1 if (self.inverse_timesteps or self.reverse_sigmas) else 0
        else:
            sigma_ = self.sigmas[timestep_id + 1].reshape(-1, 1, 1, 1)
        return sample + model_output * (sigma_ - sigma)

    def add_noise(self, original_samples, noise, timestep):
        """utils/scheduler.py:159-176: (1 - sigma) * x0 + sigma * noise in fp32, cast to noise.dtype.
        original_samples / noise: [B*T, C, H, W]; timestep: [B*T]."""
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        self._to(noise.device)
        sigma = self.sigmas[self.timestep_id(timestep)]
        if noise.is_cuda and noise.dtype == torch.bfloat16 and original_samples.dtype == torch.bfloat16:
            return ops.add_noise(original_samples.contiguous(), noise.contiguous(), sigma.float().contiguous())
        raise RuntimeError("mmpl_b200.FlowMatchScheduler.add_noise needs CUDA bfloat16 latents (no CPU fallback)")

    def training_target(self, sample, noise, timestep):
        return noise - sample
