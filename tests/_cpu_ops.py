"""Test helpers: CPU stand-ins, built from the ORACLE's arithmetic, for the two places where the product launches a CUDA
kernel from host code — `FlowMatchScheduler.add_noise` and the fused UniPC step. The host-logic tests of the pipelines
run on the CPU around a fake generator; they inject these so that what is compared with the reference pipelines' golden
traces is scheduling, cache bookkeeping and RNG order (the kernels have their own GPU parity tests)."""
import torch

from mmpl_b200.scheduler import FlowMatchScheduler
from oracle import causal_wan_oracle as O
from oracle.unipc_oracle import FlowUniPCMultistepScheduler as OracleUniPC


class CpuScheduler(FlowMatchScheduler):
    """The product's schedule tables with add_noise evaluated by the oracle restatement (utils/scheduler.py:159-176)."""

    def add_noise(self, original_samples, noise, timestep):
        sched = O.FlowMatchSchedule.__new__(O.FlowMatchSchedule)
        sched.sigmas, sched.timesteps = self.sigmas, self.timesteps
        return sched.add_noise(original_samples, noise, timestep.float().flatten())


def cpu_scheduler(shift=5.0):
    s = CpuScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
    s.set_timesteps(1000, training=True)
    return s


class EagerUniPC:
    """What the reference evaluates between two forwards, operator by operator (casual_fps_inference.py:366-374 +
    the oracle's UniPC restatement)."""

    def __init__(self, steps, shift, guidance, like, num_train_timesteps=1000):
        self.s = OracleUniPC(num_train_timesteps=num_train_timesteps, shift=1, use_dynamic_shifting=False)
        self.s.set_timesteps(steps, device=like.device, shift=shift)
        self.guidance, self.i = guidance, 0

    def step(self, flow_cond, flow_uncond, sample):
        flow = flow_cond if flow_uncond is None else flow_uncond + self.guidance * (flow_cond - flow_uncond)
        t = self.s.timesteps[self.i]
        self.i += 1
        return self.s.step(flow, t, sample, return_dict=False)[0]


def eager_unipc_factory(pipe):
    """`pipe.unipc_stepper = eager_unipc_factory(pipe)` makes a CFG pipeline run its sampler with eager torch operators."""
    return lambda like: EagerUniPC(pipe.sampling_steps, pipe.shift, pipe.args.guidance_scale, like, pipe.num_train_timesteps)
