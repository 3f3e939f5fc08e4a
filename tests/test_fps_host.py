"""Host-side logic of the MMPL (FPS) path against golden data from the reference:
UniPC sampler bit-exact vs the reference FlowUniPCMultistepScheduler; frame-slot plan vs the rows the reference
CausalFPSWanModel actually wrote and the visibility lists it kept (tests/golden/fps_model_tiny.pt)."""
from pathlib import Path

import torch

from mmpl_b200.cache_plan import plan_fps
from oracle.unipc_oracle import FlowUniPCMultistepScheduler

GOLDEN = Path(__file__).parent / "golden"


def test_unipc_bit_exact_against_reference_trajectory():
    fix = torch.load(GOLDEN / "unipc_50.pt", weights_only=False)
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(fix["steps"], device="cpu", shift=fix["shift"])
    assert torch.equal(s.timesteps, fix["timesteps"]) and torch.equal(s.sigmas, fix["sigmas"])
    x = fix["x_init"].clone()
    for i, t in enumerate(s.timesteps):
        x = s.step(fix["flows"][i], t, x, return_dict=False)[0]
        assert torch.equal(x, fix["outs"][i]), f"UniPC step {i} differs"


def test_fps_plan_matches_rows_written_by_the_reference_model():
    fix = torch.load(GOLDEN / "fps_model_tiny.pt", weights_only=False)
    fs, rows = 1560, 15 * 1560
    vis = []
    for i, call in enumerate(fix["calls"]):
        # the pipeline's visibility edits between stages (pipeline/casual_fps_inference.py:297-325)
        if i == 3:
            for v in (31200, 29640):
                if v in vis:
                    vis.remove(v)
        if i == 4:
            for v in (31200, 29640):
                if v not in vis:
                    vis.append(v)
        p = plan_fps(vis, [f * fs for f in call["frames"]], fs, rows)
        assert sorted(vis) == call["vis"], f"call {i}: visibility list"
        written = [] if p.kv_to_tail else sorted(r // fs for r in p.kv_row)
        assert written == call["changed_slots"], f"call {i}: written slots {written} vs reference {call['changed_slots']}"
        assert p.frame_pos == call["frames"]
        attended = sorted(s + j * fs for s, n in p.segments for j in range(n // fs))
        expect = sorted(v - 6 * fs if v >= 19 * fs else v for v in call["vis"])
        assert attended == expect
        assert call["end_indices"] == (0, 0)  # the FPS model never touches the end indices
