"""ORACLE (test infrastructure, never imported by the product): eager restatement of the reference's flow-matching UniPC
sampler, wan/utils/fm_solvers_unipc.py:20-739, for the only configuration the reference pipelines instantiate
(pipeline/casual_fps_inference.py:503-512): solver_order 2, solver_type "bh2", predict_x0, flow_prediction,
lower_order_final, no thresholding, final sigma 0, `shift=1` at construction and the real shift passed to `set_timesteps`.

It is the same sequence of element-wise torch operators on the latent dtype as the reference (bf16 latents stay bf16: the
sigma scalars are 0-dim fp32 CPU tensors and do not promote them), so on any device it computes what the reference
computes on that device. Pinned bit-exact against a 50-step trajectory of the unmodified reference class
(tests/golden/unipc_50.pt, tests/test_fps_host.py). The product's sampler is mmpl_b200/unipc.py (coefficient table + one
fused kernel per step); the tests check it against this file, and the CPU host-logic tests of the pipelines inject this
class where the product would launch its kernel.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import numpy as np
import torch


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, shift: Optional[float] = 1.0,
                 use_dynamic_shifting: bool = False, lower_order_final: bool = True, disable_corrector: List[int] = ()):
        if use_dynamic_shifting:
            raise NotImplementedError("dynamic shifting is not used by the MMPL pipeline")
        if solver_order not in (1, 2):
            raise NotImplementedError("orders above 2 need the linear solve of fm_solvers_unipc.py:457-458")
        self.num_train_timesteps = num_train_timesteps
        self.solver_order = solver_order
        self.shift = shift
        self.lower_order_final = lower_order_final
        self.disable_corrector = list(disable_corrector)
        self.num_inference_steps = None
        # fm_solvers_unipc.py:107-131
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas.to("cpu")
        self.timesteps = sigmas * num_train_timesteps
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self._reset_state()

    def _reset_state(self):
        self.model_outputs = [None] * self.solver_order
        self.timestep_list = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = 1
        self._step_index = None
        self._begin_index = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps: int, device=None, shift: Optional[float] = None):
        """fm_solvers_unipc.py:160-228: linspace(sigma_max, sigma_min, n+1)[:-1], shifted, final sigma 0; integer
        (truncated) timesteps."""
        sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas).to("cpu")
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._reset_state()

    # ---------------------------------------------------------------------------------------------- pieces
    def _lambda(self, sigma):
        return torch.log(1 - sigma) - torch.log(sigma)

    def _coeffs(self, sigma_t, sigma_s0):
        """h, h*phi_1(h) and B(h) for predict_x0 / bh2 (fm_solvers_unipc.py:408-447)."""
        h = self._lambda(sigma_t) - self._lambda(sigma_s0)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        return h, h_phi_1, torch.expm1(hh)

    def convert_model_output(self, model_output, sample):
        """flow prediction -> x0 (:318-321)."""
        return sample - self.sigmas[self.step_index] * model_output

    def _predict(self, sample, order):
        """multistep_uni_p_bh_update (:350-484)."""
        m0 = self.model_outputs[-1]
        x = sample
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t = 1 - sigma_t
        h, h_phi_1, B_h = self._coeffs(sigma_t, sigma_s0)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if order == 2:
            mi = self.model_outputs[-2]
            rk = (self._lambda(self.sigmas[self.step_index - 1]) - self._lambda(sigma_s0)) / h
            D1 = (mi - m0) / rk
            rhos_p = torch.tensor([0.5], dtype=x.dtype, device=x.device)
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, torch.stack([D1], dim=1))
        else:
            pred_res = 0
        x_t = x_t_ - alpha_t * B_h * pred_res
        return x_t.to(x.dtype)

    def _correct(self, this_model_output, last_sample, this_sample, order):
        """multistep_uni_c_bh_update (:486-626)."""
        m0 = self.model_outputs[-1]
        x = last_sample
        model_t = this_model_output
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t = 1 - sigma_t
        h, h_phi_1, B_h = self._coeffs(sigma_t, sigma_s0)
        device = this_sample.device
        hh = -h
        if order == 1:
            rhos_c = torch.tensor([0.5], dtype=x.dtype, device=device)
            D1s = None
        else:
            mi = self.model_outputs[-2]
            rk = (self._lambda(self.sigmas[self.step_index - 2]) - self._lambda(sigma_s0)) / h
            D1s = torch.stack([(mi - m0) / rk], dim=1)
            rks = torch.tensor([rk, 1.0], device=device)
            # R, b of :594-603 for order 2
            h_phi_k = h_phi_1 / hh - 1
            b0 = h_phi_k * 1 / B_h
            h_phi_k = h_phi_k / hh - 1 / 2
            b1 = h_phi_k * 2 / B_h
            R = torch.stack([torch.pow(rks, 0), torch.pow(rks, 1)])
            b = torch.tensor([b0, b1], device=device)
            rhos_c = torch.linalg.solve(R, b).to(device).to(x.dtype)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], D1s) if D1s is not None else 0
        D1_t = model_t - m0
        x_t = x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)
        return x_t.to(x.dtype)

    def _init_step_index(self, timestep):
        """:628-653 — second match if the timestep is duplicated (integer truncation can duplicate)."""
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        indices = (self.timesteps == timestep).nonzero()
        pos = 1 if len(indices) > 1 else 0
        self._step_index = indices[pos].item()

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             return_dict: bool = True, generator=None) -> Tuple[torch.Tensor]:
        """:655-739."""
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        if self.step_index is None:
            self._init_step_index(timestep)
        use_corrector = (self.step_index > 0 and self.step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        model_output_convert = self.convert_model_output(model_output, sample)
        if use_corrector:
            sample = self._correct(model_output_convert, self.last_sample, sample, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
            self.timestep_list[i] = self.timestep_list[i + 1]
        self.model_outputs[-1] = model_output_convert
        self.timestep_list[-1] = timestep
        if self.lower_order_final:
            this_order = min(self.solver_order, len(self.timesteps) - self.step_index)
        else:
            this_order = self.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev_sample = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return (prev_sample,)
