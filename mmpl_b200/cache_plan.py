"""Integer bookkeeping of the KV caches (pure Python, bit-exact with the reference, testable without a GPU).

The reference does this arithmetic inside every attention layer with two `.item()` host syncs per layer
(wan/modules/causal_model.py:193-226) and, for the MMPL model, with Python lists of token offsets
(wan/modules/causal_fps_model.py:192-264). Here it runs once per forward on host integers and the result is
handed to the CUDA path as (rows to write, RoPE positions, row segments to attend).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple


@dataclass
class AttendPlan:
    """What one forward does to a layer's KV cache."""
    frame_pos: List[int]                 # temporal RoPE position per frame of this call
    kv_row: List[int]                    # cache row that receives each frame's first token
    segments: List[Tuple[int, int]]      # (start_row, n_rows) attended after the write, in order
    kv_to_tail: bool = False             # K/V of this call are not stored, only attended (last MMPL stage)
    global_end: int = 0                  # value written to kv_cache["global_end_index"]
    local_end: int = 0                   # value written to kv_cache["local_end_index"]
    local_start: int = 0


def plan_contiguous(local_end_prev: int, global_end_prev: int, current_start: int, num_frames: int,
                    frame_seqlen: int, cache_rows: int, max_attention_size: int = 32760) -> AttendPlan:
    """CausalWanSelfAttention KV branch (causal_model.py:193-226).

    local_end = local_end_prev + (current_start + S) - global_end_prev; rows [local_end - S, local_end) are
    overwritten and rows [max(0, local_end - max_attention_size), local_end) are attended. Re-running a chunk
    at another timestep therefore rewrites the same rows; a new chunk advances both indices by S."""
    num_new = num_frames * frame_seqlen
    current_end = current_start + num_new
    local_end = local_end_prev + current_end - global_end_prev
    local_start = local_end - num_new
    if local_start < 0 or local_end > cache_rows:
        # the reference would fail with a shape mismatch in the slice assignment (causal_model.py:216)
        raise IndexError(f"KV cache overflow: rows [{local_start}, {local_end}) outside a {cache_rows}-row cache")
    win_start = max(0, local_end - max_attention_size)
    start_frame = current_start // frame_seqlen
    return AttendPlan(
        frame_pos=[start_frame + f for f in range(num_frames)],
        kv_row=[local_start + f * frame_seqlen for f in range(num_frames)],
        segments=[(win_start, local_end - win_start)],
        global_end=current_end, local_end=local_end, local_start=local_start)


# ---------------------------------------------------------------------------------------------------------
# MMPL frame-slot cache (CausalFPSWanModel)
# ---------------------------------------------------------------------------------------------------------
FPS_REMAP_FROM = 19   # frames >= 19 live 6 slots lower (causal_fps_model.py:213-216: x - 6*1560 if x >= 19*1560)
FPS_REMAP_SHIFT = 6
FPS_LAST_STAGE_FRAME = 15  # a call containing frame 15 is the last stage: no cache write (:254-264)


def fps_slot(frame: int) -> int:
    return frame - FPS_REMAP_SHIFT if frame >= FPS_REMAP_FROM else frame


def merge_rows(starts: Sequence[int], length: int) -> List[Tuple[int, int]]:
    """Sorted, merged (start, n_rows) runs covering rows [s, s+length) for s in starts."""
    runs: List[Tuple[int, int]] = []
    for s in sorted(set(starts)):
        if runs and runs[-1][0] + runs[-1][1] == s:
            runs[-1] = (runs[-1][0], runs[-1][1] + length)
        else:
            runs.append((s, length))
    return runs


def plan_fps(visible: List[int], current_start: Sequence[int], frame_seqlen: int, cache_rows: int) -> AttendPlan:
    """CausalFPSWanModel self-attention (causal_fps_model.py:192-264).

    `current_start` is the list of per-frame token offsets (frame * frame_seqlen) of this call; `visible` is
    the layer's kv_cache["attention_vis_index"] list, updated in place like the reference does. Three cases,
    keyed exactly as the reference keys them (its literal 1560 is `frame_seqlen` here):
      * frame 15 present (last stage, :254-264): nothing is written, `visible` is only de-duplicated, and the
        new K/V are attended as an extra trailing segment;
      * frame 19 present (anchor stage, :228-252): every frame is written at its own offset except list
        positions 5 and 6, which are written 6 slots lower;
      * otherwise (:209-227): every frame is written at its own offset.
    In the two writing cases `visible` becomes the union with `current_start`, and the attended rows are the
    slots of all visible offsets with x -> x - 6*frame_seqlen for x >= 19*frame_seqlen. The reference gathers
    them in `list(set(...))` order; attention is invariant to key order, so sorted merged runs are returned."""
    cur = [int(s) for s in current_start]
    fs = frame_seqlen
    frames = [s // fs for s in cur]
    last_stage = FPS_LAST_STAGE_FRAME * fs in cur
    if last_stage:
        dedup = list(dict.fromkeys(visible))
        visible[:] = dedup
        kv_row = [i * fs for i in range(len(cur))]  # rows of the tail buffer
    else:
        if FPS_REMAP_FROM * fs in cur:
            kv_row = [s - FPS_REMAP_SHIFT * fs if i in (5, 6) else s for i, s in enumerate(cur)]
        else:
            kv_row = list(cur)
        seen = set(visible)
        for s in cur:
            if s not in seen:
                visible.append(s)
                seen.add(s)
        for r in kv_row:
            if r < 0 or r + fs > cache_rows:
                raise IndexError(f"KV rows [{r}, {r + fs}) outside a {cache_rows}-row cache")
    starts = [v - FPS_REMAP_SHIFT * fs if v >= FPS_REMAP_FROM * fs else v for v in visible]
    for r in starts:
        if r < 0 or r + fs > cache_rows:
            raise IndexError(f"visible rows [{r}, {r + fs}) outside a {cache_rows}-row cache")
    return AttendPlan(frame_pos=frames, kv_row=kv_row, segments=merge_rows(starts, fs), kv_to_tail=last_stage)
