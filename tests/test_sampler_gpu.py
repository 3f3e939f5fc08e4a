"""mmpl_unipc_cfg_step (csrc/sampler.cu) on the GPU, through the C ABI:

  * against the plain-torch emulation of its element program (tests/_emulate.py), both scalar semantics, bit-exact;
  * "cuda" semantics against the eager operator sequence of the reference run ON THE DEVICE (the oracle's restatement of
    fm_solvers_unipc.py + the pipeline's CFG combine) - what the reference computes on a GPU - bit-exact;
  * "cpu" semantics against the 50-step trajectory recorded from the unmodified reference scheduler on the CPU
    (tests/golden/unipc_50.pt), bit-exact.
"""
import ctypes as C
from pathlib import Path

import pytest
import torch

from _emulate import emulate_unipc_step
from mmpl_b200 import _lib
from mmpl_b200.unipc import FlowUniPCMultistepScheduler, FusedUniPC, UniPCTable
from oracle.unipc_oracle import FlowUniPCMultistepScheduler as OracleUniPC

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
DEV = "cuda"


def _rand(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16)


@pytest.mark.parametrize("semantics", ["cuda", "cpu"])
@pytest.mark.parametrize("n", [7 * 16 * 60 * 104, 8 * 1000 + 5, 3])
def test_kernel_equals_its_element_program(semantics, n):
    lib = _lib.load()
    table = UniPCTable(6, 5.0, guidance=5.0, semantics=semantics)
    for i, k in enumerate(table.coeffs):
        c, u, x, m1, m2, last = (_rand((n,), 10 * i + j) for j in range(6))
        for combine in (True, False):
            ref = emulate_unipc_step(k, c, u if combine else None, x, m1, m2, last)
            d = [t.to(DEV) for t in (c, u, x, m1, m2, last)]
            out = [torch.empty(n, dtype=torch.bfloat16, device=DEV) for _ in range(3)]
            ks = k.as_struct()
            _lib.check(lib.mmpl_unipc_cfg_step(d[0].data_ptr(), d[1].data_ptr() if combine else None, d[2].data_ptr(),
                                               d[3].data_ptr(), d[4].data_ptr(), d[5].data_ptr(), out[0].data_ptr(),
                                               out[1].data_ptr(), out[2].data_ptr(), n, C.byref(ks),
                                               torch.cuda.current_stream().cuda_stream))
            for name, got, want in zip(("next", "x0", "corrected"), out, ref):
                assert torch.equal(got.cpu(), want), f"step {i} combine={combine}: {name} differs from the element program"


@pytest.mark.parametrize("steps,shift,scale", [(3, 5.0, 5.0), (8, 3.0, 7.5), (50, 5.0, 5.0)])
def test_fused_run_equals_eager_torch_on_the_device(steps, shift, scale):
    """The whole multistep run: CFG combine + scheduler.step as eager torch operators on cuda (what the reference executes
    on a GPU) against one fused launch per step. Same flows for both; bit-exact at every step."""
    shape = (1, 7, 16, 12, 20)
    x0 = _rand(shape, 1).to(DEV)
    fc = [_rand(shape, 100 + i).to(DEV) for i in range(steps)]
    fu = [_rand(shape, 500 + i).to(DEV) for i in range(steps)]
    eager = OracleUniPC(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    eager.set_timesteps(steps, device=DEV, shift=shift)
    fused = FusedUniPC(UniPCTable(steps, shift, guidance=scale, semantics="cuda", solve_device=torch.device(DEV)), x0)
    assert torch.equal(fused.timesteps, eager.timesteps.cpu())
    ref, got = x0.clone(), x0.clone()
    for i, t in enumerate(eager.timesteps):
        flow = fu[i] + scale * (fc[i] - fu[i])
        ref = eager.step(flow, t, ref, return_dict=False)[0]
        got = fused.step(fc[i], fu[i], got)
        bad = (got != ref).sum().item()
        assert bad == 0, f"step {i}: {bad} of {ref.numel()} elements differ from eager torch on the device"
        assert torch.equal(fused.last_x0, eager.model_outputs[-1])


def test_cpu_semantics_reproduce_the_reference_trajectory():
    fix = torch.load(GOLDEN / "unipc_50.pt", weights_only=False)
    table = UniPCTable(fix["steps"], fix["shift"], guidance=1.0, semantics="cpu")
    x = fix["x_init"].to(DEV)
    run = FusedUniPC(table, x)
    for i in range(fix["steps"]):
        x = run.step(fix["flows"][i].to(DEV).contiguous(), None, x.contiguous())
        assert torch.equal(x.cpu(), fix["outs"][i]), f"UniPC step {i} differs from the reference trajectory"
    with pytest.raises(RuntimeError):
        run.step(fix["flows"][0].to(DEV), None, x)


def test_reference_shaped_scheduler_front():
    """FlowUniPCMultistepScheduler(...).set_timesteps / .step with the reference's call shape drives the same kernel."""
    steps, shift = 5, 5.0
    shape = (1, 2, 16, 8, 12)
    eager = OracleUniPC(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    eager.set_timesteps(steps, device=DEV, shift=shift)
    mine = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    mine.set_timesteps(steps, device=DEV, shift=shift)
    assert torch.equal(mine.timesteps, eager.timesteps) and torch.equal(mine.sigmas, eager.sigmas)
    a = b = _rand(shape, 3).to(DEV)
    for i, t in enumerate(eager.timesteps):
        flow = _rand(shape, 40 + i).to(DEV)
        a = eager.step(flow, t, a, return_dict=False)[0]
        b = mine.step(flow, t, b, return_dict=False)[0]
        assert torch.equal(a, b), i
