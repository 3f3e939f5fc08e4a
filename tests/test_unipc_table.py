"""The product's UniPC sampler (mmpl_b200/unipc.py: coefficient table + element program of mmpl_unipc_cfg_step) on the CPU:
the table with "cpu" semantics, run through the plain-torch emulation of the kernel's element program, must reproduce the
50-step trajectory recorded from the UNMODIFIED reference scheduler (tests/golden/unipc_50.pt) bit for bit, and must
agree with the oracle's eager restatement when a CFG combine precedes each step. The kernel itself is checked against the
same emulation and against eager torch on the device in tests/test_sampler_gpu.py."""
from pathlib import Path

import torch

from mmpl_b200.unipc import UniPCTable, flow_sigmas
from oracle.unipc_oracle import FlowUniPCMultistepScheduler as OracleUniPC
from _emulate import emulate_unipc_run

GOLDEN = Path(__file__).parent / "golden"


def test_sigma_grid_matches_the_reference_scheduler():
    fix = torch.load(GOLDEN / "unipc_50.pt", weights_only=False)
    sigmas, timesteps = flow_sigmas(fix["steps"], fix["shift"])
    assert torch.equal(sigmas, fix["sigmas"]) and torch.equal(timesteps, fix["timesteps"])


def test_table_reproduces_the_reference_trajectory_bit_exact():
    fix = torch.load(GOLDEN / "unipc_50.pt", weights_only=False)
    table = UniPCTable(fix["steps"], fix["shift"], guidance=1.0, semantics="cpu")
    assert [c.pred_order for c in table.coeffs] == [1] + [2] * (fix["steps"] - 2) + [1]
    assert [c.corr_order for c in table.coeffs] == [0, 1] + [2] * (fix["steps"] - 2)
    outs = emulate_unipc_run(table, fix["flows"], None, fix["x_init"].clone())
    for i, (got, ref) in enumerate(zip(outs, fix["outs"])):
        assert torch.equal(got, ref), f"UniPC step {i} differs from the reference trajectory"


def test_cfg_combine_and_short_runs_against_the_oracle():
    """3 / 4 / 7-step runs with a CFG combine in front of every step (pipeline/casual_fps_inference.py:366-374), against the
    eager operator sequence."""
    g = torch.Generator().manual_seed(5)
    for steps, shift, scale in ((3, 5.0, 5.0), (4, 3.0, 7.5), (7, 8.0, 3.3)):
        x = torch.randn(1, 2, 16, 6, 10, generator=g).to(torch.bfloat16)
        fc = [torch.randn(x.shape, generator=g).to(torch.bfloat16) for _ in range(steps)]
        fu = [torch.randn(x.shape, generator=g).to(torch.bfloat16) for _ in range(steps)]
        table = UniPCTable(steps, shift, guidance=scale, semantics="cpu")
        outs = emulate_unipc_run(table, fc, fu, x.clone())
        s = OracleUniPC(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(steps, device="cpu", shift=shift)
        assert torch.equal(s.timesteps, table.timesteps)
        ref = x.clone()
        for i, t in enumerate(s.timesteps):
            flow = fu[i] + scale * (fc[i] - fu[i])
            ref = s.step(flow, t, ref, return_dict=False)[0]
            assert torch.equal(outs[i], ref), f"{steps}-step run, step {i}"
