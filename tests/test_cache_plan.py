"""Bit-exact integer behaviour of the KV-cache bookkeeping (mmpl_b200/cache_plan.py) against the reference's own
traces (golden fixtures) and against the oracle's restatement, plus the properties the reference relies on."""
from pathlib import Path

import pytest
import torch

from mmpl_b200.cache_plan import merge_rows, plan_contiguous, plan_fps
from oracle import causal_wan_oracle as O

GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("fixture", ["causal_tiny.pt", "causal_cfg1.pt"])
def test_contiguous_plan_reproduces_reference_trace(fixture):
    fix = torch.load(GOLDEN / fixture, weights_only=False)
    fs = (fix["lat_h"] // 2) * (fix["lat_w"] // 2)
    g = l = 0
    for call in fix["trace"]:
        p = plan_contiguous(l, g, call["current_start"], fix["nfpb"], fs, fix["cache_rows"])
        assert (p.global_end, p.local_end) == (call["global_end"], call["local_end"])
        assert p.kv_row == [p.local_start + i * fs for i in range(fix["nfpb"])]
        assert p.frame_pos == [call["current_start"] // fs + i for i in range(fix["nfpb"])]
        g, l = p.global_end, p.local_end


def test_contiguous_plan_equals_oracle_for_cfg2_schedule():
    """21 frames, 3-frame chunks, 5 calls per chunk, 32760-row cache, window 32760."""
    fs, g, l = 1560, 0, 0
    for chunk in range(7):
        for call in range(5):
            p = plan_contiguous(l, g, chunk * 3 * fs, 3, fs, 32760)
            o = O.contiguous_cache_plan(l, g, chunk * 3 * fs, 3 * fs)
            assert (p.local_start, p.local_end, p.segments[0][0]) == (o["local_start"], o["local_end"], o["win_start"])
            assert p.segments == [(0, (chunk + 1) * 4680)]
            # re-running the chunk rewrites the same rows
            assert p.local_start == chunk * 4680
            g, l = p.global_end, p.local_end
    assert l == g == 32760


def test_contiguous_plan_window_and_overflow():
    p = plan_contiguous(local_end_prev=9360, global_end_prev=9360, current_start=9360, num_frames=3, frame_seqlen=1560,
                        cache_rows=32760, max_attention_size=6 * 1560)
    assert p.segments == [(14040 - 9360, 9360)] and p.local_start == 9360
    with pytest.raises(IndexError):
        plan_contiguous(32760, 32760, 32760, 3, 1560, 32760)


def test_merge_rows():
    assert merge_rows([0, 10, 20, 50, 40], 10) == [(0, 30), (40, 20)]
    assert merge_rows([5, 5], 3) == [(5, 3)]


def test_fps_plan_follows_the_t2v_stage_schedule():
    """Stages [[0,1],[2,3,10,11,12,19,20],[4..9],[13..18]] (pipeline/casual_fps_inference.py:250-266); the slot map
    {0..12, 19->13, 20->14}, stage-3 no-write and stage-2 visibility edit are the probe results of SURVEY.md §8c."""
    fs, rows = 1560, 15 * 1560
    vis = []
    p0 = plan_fps(vis, [0, fs], fs, rows)
    assert p0.kv_row == [0, fs] and p0.segments == [(0, 2 * fs)] and not p0.kv_to_tail and p0.frame_pos == [0, 1]
    stage1 = [f * fs for f in (2, 3, 10, 11, 12, 19, 20)]
    p1 = plan_fps(vis, stage1, fs, rows)
    assert p1.kv_row == [2 * fs, 3 * fs, 10 * fs, 11 * fs, 12 * fs, 13 * fs, 14 * fs]
    assert p1.frame_pos == [2, 3, 10, 11, 12, 19, 20]
    assert p1.segments == [(0, 4 * fs), (10 * fs, 5 * fs)]
    assert sorted(vis) == sorted([0, fs] + stage1)
    # a second call of the same stage (next UniPC step) changes nothing
    p1b = plan_fps(vis, stage1, fs, rows)
    assert (p1b.kv_row, p1b.segments) == (p1.kv_row, p1.segments) and len(vis) == 9
    # stage 2: the pipeline hides frames 19, 20 (casual_fps_inference.py:297-325)
    vis.remove(19 * fs); vis.remove(20 * fs)
    p2 = plan_fps(vis, [f * fs for f in range(4, 10)], fs, rows)
    assert p2.kv_row == [f * fs for f in range(4, 10)] and p2.segments == [(0, 13 * fs)]
    # stage 3: frames 19, 20 visible again; no cache write, new K/V attended as a tail
    vis += [19 * fs, 20 * fs]
    before = list(vis)
    p3 = plan_fps(vis, [f * fs for f in range(13, 19)], fs, rows)
    assert p3.kv_to_tail and p3.kv_row == [i * fs for i in range(6)] and vis == before
    assert p3.segments == [(0, 15 * fs)] and p3.frame_pos == list(range(13, 19))


def test_fps_plan_rejects_rows_outside_the_cache():
    with pytest.raises(IndexError):
        plan_fps([], [14 * 1560, 16 * 1560], 1560, 15 * 1560)
