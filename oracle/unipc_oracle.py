"""ORACLE (test infrastructure, never imported by the product): the flow-matching UniPC sampler of the reference,
wan/utils/fm_solvers_unipc.py, restated as pure functions over an explicit run state, for the one configuration the reference
pipelines instantiate (pipeline/casual_fps_inference.py:503-512): order 2, "bh2", x0 prediction from a flow output,
lower_order_final, no thresholding, final sigma 0, shift 1 at construction and the real shift given with the step count.

What is restated is the *operator sequence on the latent tensors*: which torch operator is applied to which operands in
which order, with 0-dim fp32 CPU tensors as the step scalars (they do not promote bf16 latents). That sequence is what
decides the roundings, so on any device these functions compute what the reference computes on that device. Pinned
bit-exact against a 50-step trajectory recorded from the unmodified reference class (tests/golden/unipc_50.pt,
tests/test_fps_host.py). The product's sampler is mmpl_b200/unipc.py (a per-step coefficient table + one fused kernel);
its tests compare it with this file, and the CPU host-logic tests of the pipelines run this where the product launches
its kernel.

Layout: `sigma_schedule` (the two sigma tables), `StepTerms` / `step_terms` (everything of one update that depends on
the step index alone), `predict` (UniP), `corrector_weights` + `correct` (UniC), `advance` (one scheduler step over a
`Run`), and `FlowUniPCMultistepScheduler`, a facade with the reference's call shape for the tests that drive it like
the reference class.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

ORDER = 2   # solver_order of the reference pipelines; order 3 would need the general linear solve of :457-458


# ------------------------------------------------------------------------------------------------------ sigma tables
def training_sigmas(num_train: int, shift: float) -> torch.Tensor:
    """The table the constructor builds (fm_solvers_unipc.py:107-131); only its two ends are used afterwards."""
    s = torch.from_numpy(1.0 - np.linspace(1, 1 / num_train, num_train)[::-1].copy()).to(torch.float32)
    return (shift * s / (1 + (shift - 1) * s)).to("cpu")


def sigma_schedule(steps: int, shift: float, sigma_max: float, sigma_min: float, num_train: int, device=None):
    """`set_timesteps` (:160-228): `steps` sigmas from sigma_max down towards sigma_min (end point dropped), shifted, with a
    final 0 appended; timesteps = sigma * num_train truncated to integers. Returns (sigmas fp32 CPU [steps + 1], timesteps
    int64 on `device` [steps])."""
    s = np.linspace(sigma_max, sigma_min, steps + 1).copy()[:-1]
    s = shift * s / (1 + (shift - 1) * s)
    timesteps = torch.from_numpy(s * num_train).to(device=device, dtype=torch.int64)
    sigmas = torch.from_numpy(np.concatenate([s, [0]]).astype(np.float32)).to("cpu")
    return sigmas, timesteps


# -------------------------------------------------------------------------------------------- per-step scalar terms
def log_snr(sigma: torch.Tensor) -> torch.Tensor:
    """lambda = log(alpha) - log(sigma) with alpha = 1 - sigma (:323-325 and the lambda lines of both updates)."""
    return torch.log(1 - sigma) - torch.log(sigma)


@dataclass
class StepTerms:
    """Scalars of one update from sigma_from to sigma_to (0-dim fp32 CPU tensors)."""
    carry: torch.Tensor        # sigma_to / sigma_from: weight of the sample the update starts from
    alpha: torch.Tensor        # 1 - sigma_to
    h: torch.Tensor            # lambda_to - lambda_from
    em1: torch.Tensor          # expm1(-h): both h*phi_1(h) and B(h) of the "bh2" variant under x0 prediction
    r_prev: Optional[torch.Tensor]   # (lambda_prev - lambda_from) / h for the second-order difference, None at order 1


def step_terms(sigmas: torch.Tensor, i_from: int, i_to: int, i_prev: Optional[int]) -> StepTerms:
    s_to, s_from = sigmas[i_to], sigmas[i_from]
    lam_from = log_snr(s_from)
    h = log_snr(s_to) - lam_from
    r_prev = None if i_prev is None else (log_snr(sigmas[i_prev]) - lam_from) / h
    return StepTerms(carry=s_to / s_from, alpha=1 - s_to, h=h, em1=torch.expm1(-h), r_prev=r_prev)


def _weighted(rhos: torch.Tensor, diffs: Sequence[torch.Tensor]):
    """sum_k rhos[k] * diffs[k] as the reference forms it: one einsum over the stacked differences (:470-471, :612-613)."""
    return torch.einsum("k,bkc...->bc...", rhos, torch.stack(list(diffs), dim=1))


# ------------------------------------------------------------------------------------------------------ the updates
def flow_to_x0(sample: torch.Tensor, flow: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """convert_model_output for flow prediction (:318-321)."""
    return sample - sigma * flow


def predict(sample: torch.Tensor, x0_now: torch.Tensor, x0_before: Optional[torch.Tensor], tm: StepTerms) -> torch.Tensor:
    """UniP (multistep_uni_p_bh_update, :350-484): the next sample from the current one and the last one or two x0
    estimates. Second order when `x0_before` is given (then tm.r_prev is too); its single weight is the constant 1/2
    (:459-461)."""
    base = tm.carry * sample - tm.alpha * tm.em1 * x0_now
    if x0_before is None:
        second = 0
    else:
        half = torch.tensor([0.5], dtype=sample.dtype, device=sample.device)
        second = _weighted(half, [(x0_before - x0_now) / tm.r_prev])
    return (base - tm.alpha * tm.em1 * second).to(sample.dtype)


def corrector_weights(tm: StepTerms, order: int, dtype, device) -> torch.Tensor:
    """The rho_c of UniC (:560-610). Order 1: the constant 1/2. Order 2: the solution of the 2 x 2 system R rho = b with
    R = [[1, 1], [r_prev, 1]] and b_k = k! * phi_{k+1}(-h) * (-h) / B(h), solved with torch.linalg.solve on `device` like
    the reference does."""
    if order == 1:
        return torch.tensor([0.5], dtype=dtype, device=device)
    hh = -tm.h
    phi = tm.em1 / hh - 1
    b_first = phi * 1 / tm.em1
    phi = phi / hh - 1 / 2
    b_second = phi * 2 / tm.em1
    nodes = torch.tensor([tm.r_prev, 1.0], device=device)
    system = torch.stack([torch.pow(nodes, 0), torch.pow(nodes, 1)])
    rhs = torch.tensor([b_first, b_second], device=device)
    return torch.linalg.solve(system, rhs).to(device).to(dtype)


def correct(start: torch.Tensor, x0_last: torch.Tensor, x0_before: Optional[torch.Tensor], x0_new: torch.Tensor,
            tm: StepTerms, like: torch.Tensor) -> torch.Tensor:
    """UniC (multistep_uni_c_bh_update, :486-626): redo the previous step's update - from `start`, the sample that step
    started from - now that the x0 estimate at its end point (`x0_new`) is known. `x0_last` / `x0_before` are the one or
    two estimates the previous step had; `like` (the predicted sample) only supplies the device of the small tensors."""
    order = 1 if x0_before is None else 2
    rhos = corrector_weights(tm, order, start.dtype, like.device)
    base = tm.carry * start - tm.alpha * tm.em1 * x0_last
    history = 0 if x0_before is None else _weighted(rhos[:-1], [(x0_before - x0_last) / tm.r_prev])
    newest = x0_new - x0_last
    return (base - tm.alpha * tm.em1 * (history + rhos[-1] * newest)).to(start.dtype)


# ------------------------------------------------------------------------------------------------- one scheduler step
@dataclass
class Run:
    """State of one sampling run (what the reference class keeps in attributes between `step` calls)."""
    sigmas: torch.Tensor
    timesteps: torch.Tensor
    index: Optional[int] = None                 # step_index; found from the first timestep handed in
    x0: List[torch.Tensor] = field(default_factory=list)   # the last (at most ORDER) x0 estimates, oldest first
    start: Optional[torch.Tensor] = None        # the sample the previous step started from (last_sample)
    order_used: int = 1                         # order of the previous prediction (this_order)
    warm: int = 0                               # completed steps, capped at ORDER (lower_order_nums)
    skip_corrector: Tuple[int, ...] = ()


def first_index(timesteps: torch.Tensor, timestep) -> int:
    """_init_step_index (:628-653): the position of `timestep` in the table - the SECOND one if it occurs twice
    (truncation to integers can repeat a value)."""
    if isinstance(timestep, torch.Tensor):
        timestep = timestep.to(timesteps.device)
    hits = (timesteps == timestep).nonzero()
    return hits[1 if len(hits) > 1 else 0].item()


def advance(run: Run, flow: torch.Tensor, timestep, sample: torch.Tensor) -> torch.Tensor:
    """`step` (:655-739): x0 from the flow; correct the incoming sample with it (every step but the first); predict the
    next sample at the order the history and the number of remaining steps allow."""
    if run.index is None:
        run.index = first_index(run.timesteps, timestep)
    i = run.index
    x0_new = flow_to_x0(sample, flow, run.sigmas[i])
    if i > 0 and (i - 1) not in run.skip_corrector and run.start is not None:
        second = run.order_used == 2
        tm = step_terms(run.sigmas, i - 1, i, i - 2 if second else None)
        sample = correct(run.start, run.x0[-1], run.x0[-2] if second else None, x0_new, tm, sample)
    run.x0 = (run.x0 + [x0_new])[-ORDER:]
    order = min(ORDER, len(run.timesteps) - i, run.warm + 1)     # lower_order_final, then the warm-up cap (:715-723)
    run.order_used, run.start = order, sample
    second = order == 2
    tm = step_terms(run.sigmas, i, i + 1, i - 1 if second else None)
    nxt = predict(sample, run.x0[-1], run.x0[-2] if second else None, tm)
    run.warm = min(ORDER, run.warm + 1)
    run.index = i + 1
    return nxt


# ------------------------------------------------------------------------------------- the reference's call shape
class FlowUniPCMultistepScheduler:
    """The functions above behind the reference class's constructor / `set_timesteps` / `step` calls."""

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = ORDER, shift: Optional[float] = 1.0,
                 use_dynamic_shifting: bool = False, lower_order_final: bool = True, disable_corrector: Sequence[int] = ()):
        if use_dynamic_shifting or solver_order != ORDER or not lower_order_final:
            raise NotImplementedError("the MMPL pipeline uses order 2, lower_order_final and a static shift")
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.skip_corrector = tuple(disable_corrector)
        table = training_sigmas(num_train_timesteps, shift)
        self.sigma_max, self.sigma_min = table[0].item(), table[-1].item()
        self.run = Run(sigmas=table, timesteps=table * num_train_timesteps)
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int, device=None, shift: Optional[float] = None):
        sigmas, timesteps = sigma_schedule(num_inference_steps, self.shift if shift is None else shift, self.sigma_max,
                                           self.sigma_min, self.num_train_timesteps, device)
        self.run = Run(sigmas=sigmas, timesteps=timesteps, skip_corrector=self.skip_corrector)
        self.num_inference_steps = len(timesteps)

    sigmas = property(lambda self: self.run.sigmas)
    timesteps = property(lambda self: self.run.timesteps)
    step_index = property(lambda self: self.run.index)
    model_outputs = property(lambda self: self.run.x0)       # x0 estimates, newest last

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, generator=None):
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        return (advance(self.run, model_output, timestep, sample),)
