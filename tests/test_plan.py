"""Rollout plans (mmpl_b200/pipeline/plan.py) are pure host data: the records they contain are checked here against the
reference's stage literals and loop arithmetic (pipeline/casual_fps_inference.py:250-325,380-439; MMPL_i2v :253-435;
pipeline/causal_inference.py:84-98,134-185; pipeline/causal_diffusion_inference.py:141-230). That executing the plans
reproduces the reference pipelines call for call is tests/test_*_pipelines_golden.py."""
import pytest

from mmpl_b200.pipeline.plan import Denoise, Prefill, mmpl_stages, plan_contiguous, plan_mmpl


def test_mmpl_stage_maps_are_the_reference_literals():
    assert mmpl_stages("t2v") == [(0, 1), (2, 3, 10, 11, 12, 19, 20), (4, 5, 6, 7, 8, 9), (13, 14, 15, 16, 17, 18)]
    assert mmpl_stages("i2v") == [(0,), (1,), (2, 3, 10, 11, 12, 19, 20), (4, 5, 6, 7, 8, 9), (13, 14, 15, 16, 17, 18)]


def test_t2v_segment_plan():
    p = plan_mmpl("t2v", 21, 0)
    assert [type(r) for r in p.records] == [Denoise] * 4 and p.out_frames == 21 and p.sampler == "unipc"
    s0, s1, s2, s3 = p.records
    assert s1.handoff == ((0,), True) and s0.handoff is s2.handoff is s3.handoff is None
    assert s2.renoise == ((0, 3), (5, 10)) and s2.hide == (20, 19) and s2.show == ()
    assert s3.renoise == ((0, 12), (5, 19)) and s3.show == (20, 19) and s3.hide == ()
    assert all(r.temporal == r.slot == r.noise == r.out for r in p.records)
    # continuing a previous segment: the two connect frames replace stage 0
    q = plan_mmpl("t2v", 21, 2)
    assert isinstance(q.records[0], Prefill) and q.records[0].out == (0, 1) and q.records[0].t_len == 2
    assert q.records[1:] == p.records[1:]
    with pytest.raises(AssertionError):
        plan_mmpl("t2v", 21, 1)
    with pytest.raises(AssertionError):
        plan_mmpl("t2v", 20, 0)


def test_i2v_segment_plans():
    image = plan_mmpl("i2v", 21, 1)
    assert [type(r) for r in image.records] == [Prefill] + [Denoise] * 4
    assert image.records[0].out == (0,) and image.records[1].out == (1,)
    assert image.records[2].handoff == ((0, 19, 20), False)
    assert all(not r.renoise and not r.hide and not r.show for r in image.stages)   # no re-noising in the i2v schedule
    connect = plan_mmpl("i2v", 21, 2)
    assert [type(r) for r in connect.records] == [Prefill, Prefill] + [Denoise] * 3
    assert [r.out for r in connect.records[:2]] == [(0,), (1,)] and all(r.t_len == 1 for r in connect.records[:2])
    assert connect.records[2:] == image.records[2:]


def test_contiguous_plans():
    p = plan_contiguous(21, 0, 3, False, "fewstep")          # BASELINE config 1: seven 3-frame chunks
    assert len(p.records) == 7 and all(isinstance(r, Denoise) and r.slot is None for r in p.records)
    assert [r.temporal for r in p.records] == [0, 3, 6, 9, 12, 15, 18] and p.records[3].noise == (9, 10, 11)
    ext = plan_contiguous(6, 3, 3, False, "fewstep")          # video extension: 3 clean frames, then 6 generated
    assert isinstance(ext.records[0], Prefill) and ext.records[0].source == (0, 3) and ext.out_frames == 9
    assert ext.records[1].noise == (0, 1, 2) and ext.records[1].out == (3, 4, 5) and ext.records[1].temporal == 3
    first = plan_contiguous(7, 0, 3, True, "fewstep")         # independent_first_frame: 1 + 3 + 3
    assert [len(r.out) for r in first.records] == [1, 3, 3]
    i2v = plan_contiguous(6, 4, 3, True, "fewstep")           # first frame alone, then one block of clean frames
    assert [(type(r).__name__, len(r.out)) for r in i2v.records] == [("Prefill", 1), ("Prefill", 3), ("Denoise", 3), ("Denoise", 3)]
    cfg = plan_contiguous(6, 0, 3, False, "unipc", start_frame=5, with_slot=True)
    assert [(r.temporal, r.slot) for r in cfg.records] == [(5, 0), (8, 3)]
    with pytest.raises(AssertionError):
        plan_contiguous(7, 0, 3, False, "fewstep")
