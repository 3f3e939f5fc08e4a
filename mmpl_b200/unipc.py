"""The UniPC flow sampler of the MMPL hot loop as a coefficient table plus one fused kernel launch per step
(SURVEY.md §8f-1).

What the reference does between two backbone forwards (pipeline/casual_fps_inference.py:366-374 and
wan/utils/fm_solvers_unipc.py:655-739) is ~25 element-wise torch operators whose only non-tensor inputs are scalars that
depend on the step index alone: sigma_i, sigma_{i+1}/sigma_i, alpha*h*phi_1(h), alpha*B(h), the order-2 ratios r_k and
the corrector weights rho. Here those scalars are computed ONCE per (steps, shift) into a `UniPCTable`
(host fp32 arithmetic with the reference's own formulas, so every scalar has the reference's bits), and a step is

    mmpl_unipc_cfg_step(flow_cond, flow_uncond, x, m1, m2, last) -> (next x, x0, corrected x)      (csrc/sampler.cu)

with the three state tensors of the multistep method (the two previous x0 predictions and the last corrected sample)
held in a ring of preallocated buffers. No host synchronisation, no per-step scalar tensors, no diffusers dependency.

Only the configuration the reference pipelines instantiate is supported (casual_fps_inference.py:503-512,
causal_diffusion_inference.py:367-378): solver_order 2, "bh2", predict_x0, flow_prediction, lower_order_final,
final sigma 0, no thresholding, no dynamic shifting.

`semantics` selects which eager behaviour of torch the table reproduces (include/mmpl_b200.h, mmpl_unipc_cfg_step):
"cuda" (default) is what the reference computes when it runs on a GPU; "cpu" is what it computes on the host, which is
how tests/golden/unipc_50.pt was recorded.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def flow_sigmas(steps: int, shift: float, num_train_timesteps: int = 1000) -> Tuple[torch.Tensor, torch.Tensor]:
    """(sigmas fp32 [steps + 1] ending in 0, integer timesteps [steps]) of FlowUniPCMultistepScheduler(shift=1)
    .set_timesteps(steps, shift=shift) (fm_solvers_unipc.py:107-131 for the end points, :160-216 for the grid)."""
    n = num_train_timesteps
    unshifted = torch.from_numpy(1.0 - np.linspace(1, 1 / n, n)[::-1].copy()).to(torch.float32)
    hi, lo = unshifted[0].item(), unshifted[-1].item()
    grid = np.linspace(hi, lo, steps + 1).copy()[:-1]
    grid = shift * grid / (1 + (shift - 1) * grid)
    timesteps = torch.from_numpy(grid * n).to(torch.int64)
    sigmas = torch.from_numpy(np.concatenate([grid, [0]]).astype(np.float32))
    return sigmas, timesteps


@dataclass(frozen=True)
class StepCoeffs:
    """The scalars of one step; field for field `mmpl_unipc_coeffs`."""
    guidance: float
    sigma: float
    corr_order: int
    corr_a: float
    corr_b: float
    corr_c: float
    corr_rk: float
    corr_rho0: float
    corr_rho1: float
    pred_order: int
    pred_a: float
    pred_b: float
    pred_c: float
    pred_rk: float
    true_division: int

    def as_struct(self) -> "_lib.UniPCCoeffs":
        return _lib.UniPCCoeffs(*[getattr(self, f) for f, _ in _lib.UniPCCoeffs._fields_])


def _bf16(v: float) -> float:
    return torch.tensor(v, dtype=torch.float32).to(torch.bfloat16).item()


class UniPCTable:
    """Every scalar of an n-step sampler run. `coeffs[i]` drives the launch that consumes the flow predicted at
    `timesteps[i]`."""

    def __init__(self, steps: int, shift: float, guidance: float = 1.0, num_train_timesteps: int = 1000,
                 semantics: str = "cuda", solve_device: Optional[torch.device] = None):
        if semantics not in ("cuda", "cpu"):
            raise ValueError("semantics must be 'cuda' or 'cpu'")
        self.steps, self.shift, self.guidance, self.semantics = steps, shift, guidance, semantics
        self.sigmas, self.timesteps = flow_sigmas(steps, shift, num_train_timesteps)
        first = (self.timesteps == self.timesteps[0]).nonzero()
        if len(first) > 1:  # the reference would start at the second match (:628-653) and then run off the table
            raise ValueError(f"{steps} steps at shift {shift} repeat the first integer timestep")
        # a 0-dim CPU tensor that multiplies a tensor from the left is rounded to bf16 first by the CPU kernels and used as
        # fp32 by the CUDA ones; a python number (the guidance scale) is used as fp32 by both
        scal = _bf16 if semantics == "cpu" else float
        solve_on = torch.device("cpu") if semantics == "cpu" or solve_device is None else solve_device
        s = self.sigmas
        lam = torch.log(1 - s) - torch.log(s)          # lambda_i = log(alpha_i) - log(sigma_i), fp32 like the reference
        self.coeffs: List[StepCoeffs] = []
        orders = []                                     # this_order chosen at each step (:715-722)
        for i in range(steps):
            orders.append(min(2, steps - i, len(orders) + 1))
        for i in range(steps):
            kw = dict(guidance=float(guidance), sigma=scal(s[i].item()), true_division=int(semantics == "cpu"),
                      corr_order=0, corr_a=0.0, corr_b=0.0, corr_c=0.0, corr_rk=1.0, corr_rho0=0.0, corr_rho1=0.0)
            if i > 0:
                # corrector of the step (i-1 -> i) at the order the previous predictor used (:486-626)
                order = orders[i - 1]
                a, b, c, hh, h = self._abc(s[i], s[i - 1], lam[i], lam[i - 1])
                kw.update(corr_order=order, corr_a=scal(a), corr_b=scal(b), corr_c=scal(c))
                if order == 1:
                    kw.update(corr_rho1=0.5)
                else:
                    rk = (lam[i - 2] - lam[i - 1]) / h
                    kw.update(corr_rk=self._divisor(rk))
                    phi = torch.expm1(hh) / hh - 1
                    b0 = phi * 1 / torch.expm1(hh)
                    b1 = (phi / hh - 1 / 2) * 2 / torch.expm1(hh)
                    rks = torch.tensor([rk, 1.0], device=solve_on)
                    R = torch.stack([torch.pow(rks, 0), torch.pow(rks, 1)])
                    rho = torch.linalg.solve(R, torch.tensor([b0, b1], device=solve_on)).to(torch.bfloat16).float().cpu()
                    kw.update(corr_rho0=rho[0].item(), corr_rho1=rho[1].item())
            # predictor of the step (i -> i+1) (:350-484)
            a, b, c, hh, h = self._abc(s[i + 1], s[i], lam[i + 1], lam[i])
            kw.update(pred_order=orders[i], pred_a=scal(a), pred_b=scal(b), pred_c=scal(c), pred_rk=1.0)
            if orders[i] == 2:
                kw.update(pred_rk=self._divisor((lam[i - 1] - lam[i]) / h))
            self.coeffs.append(StepCoeffs(**kw))

    @staticmethod
    def _abc(sigma_t, sigma_s0, lam_t, lam_s0):
        """sigma_t/sigma_s0, alpha_t*h*phi_1(h), alpha_t*B(h) for predict_x0 / bh2 (:408-447, :464-468)."""
        h = lam_t - lam_s0
        hh = -h
        e = torch.expm1(hh)
        alpha_t = 1 - sigma_t
        return (sigma_t / sigma_s0).item(), (alpha_t * e).item(), (alpha_t * e).item(), hh, h

    def _divisor(self, rk: torch.Tensor) -> float:
        """`D / rk` with rk a CPU scalar: a division on the CPU, a multiplication by float(1 / double(rk)) on CUDA."""
        if self.semantics == "cpu":
            return rk.item()
        return torch.tensor(1.0 / rk.double().item(), dtype=torch.float64).float().item()


@lru_cache(maxsize=16)
def unipc_table(steps: int, shift: float, guidance: float, num_train_timesteps: int = 1000, semantics: str = "cuda",
                solve_device: Optional[str] = None) -> UniPCTable:
    return UniPCTable(steps, shift, guidance, num_train_timesteps, semantics,
                      torch.device(solve_device) if solve_device else None)


class FusedUniPC:
    """Multistep state for one denoising run of `shape` bf16 latents on a CUDA device: `step()` once per timestep."""

    def __init__(self, table: UniPCTable, like: torch.Tensor):
        if not (like.is_cuda and like.dtype == torch.bfloat16):
            raise RuntimeError("mmpl_b200.unipc.FusedUniPC needs CUDA bfloat16 latents (there is no CPU fallback)")
        self.table = table
        self._lib = _lib.load()
        self._structs = [c.as_struct() for c in table.coeffs]
        # x0 ring (3: the step writes the oldest while reading the other two), corrected-sample and next-sample buffers
        self._x0 = [torch.empty_like(like, memory_format=torch.contiguous_format) for _ in range(3)]
        self._last = torch.empty_like(self._x0[0])
        self._next = [torch.empty_like(self._x0[0]) for _ in range(2)]
        self.index = 0

    @property
    def timesteps(self) -> torch.Tensor:
        return self.table.timesteps

    @property
    def last_x0(self) -> torch.Tensor:
        """x0 prediction of the most recent step (model_outputs[-1])."""
        return self._x0[(self.index - 1) % 3]

    def step(self, flow_cond: torch.Tensor, flow_uncond: Optional[torch.Tensor], sample: torch.Tensor) -> torch.Tensor:
        """flow_uncond None: `flow_cond` already is the guided flow. Returns the sample for the next forward (a buffer
        owned by this object, valid until the second next call)."""
        i = self.index
        if i >= self.table.steps:
            raise RuntimeError("UniPC run is over: construct a new FusedUniPC for the next stage")
        for t in (flow_cond, flow_uncond, sample):
            if t is not None and not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()
                                      and t.numel() == self._last.numel()):
                raise ValueError("FusedUniPC.step: tensors must be contiguous CUDA bfloat16 of the run's shape")
        m1, m2, out0 = self._x0[(i - 1) % 3], self._x0[(i - 2) % 3], self._x0[i % 3]
        nxt = self._next[i % 2]
        with torch.cuda.device(sample.device):
            _lib.check(self._lib.mmpl_unipc_cfg_step(
                flow_cond.data_ptr(), flow_uncond.data_ptr() if flow_uncond is not None else None, sample.data_ptr(),
                m1.data_ptr(), m2.data_ptr(), self._last.data_ptr(), nxt.data_ptr(), out0.data_ptr(), self._last.data_ptr(),
                sample.numel(), C.byref(self._structs[i]), torch.cuda.current_stream().cuda_stream))
        self.index = i + 1
        return nxt.view(sample.shape)


class FlowUniPCMultistepScheduler:
    """Reference-shaped front (fm_solvers_unipc.py:20-739) over the table and the fused kernel, for callers that drive the
    sampler themselves: `set_timesteps(n, device=, shift=)`, `.timesteps`, `.sigmas`, `step(model_output, timestep, sample,
    return_dict=False)`. `model_output` is the (already guided) flow prediction."""
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, shift: Optional[float] = 1.0,
                 use_dynamic_shifting: bool = False, **unsupported):
        if use_dynamic_shifting or solver_order != 2 or unsupported:
            raise NotImplementedError("only the configuration the MMPL pipelines instantiate: order 2, bh2, predict_x0, "
                                      f"flow_prediction, static shift (got {dict(unsupported, solver_order=solver_order)})")
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.num_inference_steps = None
        self._run: Optional[FusedUniPC] = None

    def set_timesteps(self, num_inference_steps: int, device=None, shift: Optional[float] = None):
        self.table = unipc_table(num_inference_steps, float(self.shift if shift is None else shift), 1.0,
                                 self.num_train_timesteps, "cuda", str(device) if device is not None and torch.device(device).type == "cuda" else None)
        self.sigmas = self.table.sigmas
        self.timesteps = self.table.timesteps.to(device)
        self.num_inference_steps = num_inference_steps
        self._run = None

    @property
    def step_index(self):
        return None if self._run is None else self._run.index

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, generator=None):
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        if self._run is None:
            self._run = FusedUniPC(self.table, sample)
        out = self._run.step(model_output.contiguous(), None, sample.contiguous()).clone()
        return (out,)
