"""Rollout plans: what a pipeline will do, decided before anything runs.

The three reference pipelines (pipeline/causal_inference.py, pipeline/causal_diffusion_inference.py,
pipeline/casual_fps_inference.py) interleave *deciding* what the next generator call is (frame ranges, cache positions,
which frames to re-noise, which cache frames to hide) with *making* the call, inside nested loops that carry a dozen
mutable indices. Here the decisions are a pure function of the request's shape: `plan_*()` returns a flat tuple of
records, and `runner.Rollout` executes any plan with one loop. Nothing in this file touches a tensor.

Records (frames are latent-frame indices; `temporal` is the position the model's RoPE / cache bookkeeping sees, as an int
for the contiguous-cache model and a tuple of frame slots for the frame-slot (MMPL) model):

  Prefill  clean frames of `initial_latent` pushed through the model at t = 0 to write their K/V, copied to the output
  Denoise  frames generated from noise by the plan's sampler, followed by the clean-context pass that rewrites their K/V

plus, for the MMPL schedule, what happens at stage boundaries: frames of the stage re-noised from frames already
generated, far-anchor frames hidden from / returned to the attention visibility list, and the anchor hand-off.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple, Union

Temporal = Union[int, Tuple[int, ...]]


@dataclass(frozen=True)
class Prefill:
    source: Tuple[int, int]            # [lo, hi) frames of initial_latent
    out: Tuple[int, ...]               # output frames that receive them
    temporal: Temporal                 # -> current_start
    slot: Optional[Temporal] = None    # -> cache_start (None: the call has no such argument)
    t_len: int = 1                     # entries per sample of the (all-zero) timestep tensor the call receives


@dataclass(frozen=True)
class Denoise:
    noise: Tuple[int, ...]             # frames of the noise tensor the stage starts from
    out: Tuple[int, ...]               # output frames it produces
    temporal: Temporal
    slot: Optional[Temporal] = None
    renoise: Tuple[Tuple[int, int], ...] = ()   # (index inside the stage, output frame whose re-noised copy replaces it)
    hide: Tuple[int, ...] = ()         # frames removed from every block's visibility list before the stage
    show: Tuple[int, ...] = ()         # frames put back
    handoff: Optional[Tuple[Tuple[int, ...], bool]] = None  # (output frames, append the stage's own latents) -> anchor sink


@dataclass(frozen=True)
class RolloutPlan:
    records: Tuple[Union[Prefill, Denoise], ...]
    out_frames: int
    sampler: str                       # "fewstep" | "unipc"
    description: str = ""
    calls_per_denoise: int = field(default=0, compare=False)

    @property
    def stages(self):
        return [r for r in self.records if isinstance(r, Denoise)]


def _span(lo: int, n: int) -> Tuple[int, ...]:
    return tuple(range(lo, lo + n))


def _chunks(num_frames: int, per_block: int, first_alone: bool, what: str):
    """Chunk sizes of a contiguous rollout: `per_block` frames each, optionally preceded by a single frame
    (independent_first_frame). Raises on a frame count the schedule cannot tile."""
    body = num_frames - 1 if first_alone else num_frames
    if body < 0 or body % per_block:
        raise AssertionError(f"{what}: {num_frames} frames do not split into "
                             f"{'1 + ' if first_alone else ''}blocks of {per_block}")
    return ([1] if first_alone else []) + [per_block] * (body // per_block)


def plan_contiguous(num_frames: int, num_input_frames: int, per_block: int, independent_first_frame: bool,
                    sampler: str, start_frame: int = 0, with_slot: bool = False) -> RolloutPlan:
    """Chunk-wise rollout on the contiguous KV cache: CausalInferencePipeline (few-step, causal_inference.py:47-276) and
    CausalDiffusionInferencePipeline (UniPC + CFG, causal_diffusion_inference.py:54-307).

    `num_input_frames` clean frames (image-to-video / video extension) are prefilled, then `num_frames` are generated.
    The model sees temporal positions offset by `start_frame` (the CFG pipeline's start_frame_index) while the cache and
    the output are indexed from 0; `with_slot` passes the latter as cache_start."""
    has_input = num_input_frames > 0
    gen = _chunks(num_frames, per_block, independent_first_frame and not has_input, "noise")
    pre = _chunks(num_input_frames, per_block, independent_first_frame, "initial_latent") if has_input else []
    records, at = [], 0
    for n in pre:
        records.append(Prefill(source=(at, at + n), out=_span(at, n), temporal=start_frame + at,
                               slot=at if with_slot else None))
        at += n
    for n in gen:
        records.append(Denoise(noise=_span(at - num_input_frames, n), out=_span(at, n), temporal=start_frame + at,
                               slot=at if with_slot else None))
        at += n
    return RolloutPlan(tuple(records), out_frames=num_frames + num_input_frames, sampler=sampler,
                       description=f"contiguous: {len(pre)} prefill + {len(gen)} generated chunks of {per_block}")


# ---------------------------------------------------------------------------------------------------------------------
# MMPL macro-from-micro schedule of one 21-frame segment (pipeline/casual_fps_inference.py:250-252 and :266-439;
# MMPL_i2v/pipeline/casual_fps_inference.py:253-255 and :266-435). A stage is the set of frames with the same rank:
#   t2v  rank 0: the two opening frames | 1: the anchors spread over the segment | 2: first gap | 3: second gap
#   i2v  rank 0: the image frame | 1: its successor | 2: anchors | 3, 4: the gaps
T2V_RANK = (0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 1, 1, 1, 3, 3, 3, 3, 3, 3, 1, 1)
I2V_RANK = (0, 1, 2, 2, 3, 3, 3, 3, 3, 3, 2, 2, 2, 4, 4, 4, 4, 4, 4, 2, 2)
FAR_ANCHORS = (20, 19)   # hidden while the first gap is filled (reference literals 31200 = 20*1560, 29640 = 19*1560)


def mmpl_stages(variant: str):
    rank = I2V_RANK if variant == "i2v" else T2V_RANK
    return [tuple(i for i, r in enumerate(rank) if r == s) for s in range(max(rank) + 1)]


def plan_mmpl(variant: str, num_frames: int, num_input_frames: int) -> RolloutPlan:
    """One MMPL segment. t2v: stage 0 is generated, or replaced by a t = 0 prefill when the segment continues a previous
    one (`initial_latent` = the two connect frames). i2v: a one-frame `initial_latent` (the encoded image) replaces stage
    0; a two-frame one (segment connect) replaces stages 0 and 1, one prefill call per frame."""
    if variant not in ("t2v", "i2v"):
        raise ValueError(variant)
    stages = mmpl_stages(variant)
    if num_frames != len(T2V_RANK):
        raise AssertionError(f"an MMPL segment is {len(T2V_RANK)} latent frames, got {num_frames}")
    anchor_stage = 2 if variant == "i2v" else 1
    records = []
    skip = 0
    if num_input_frames:
        if variant == "t2v":
            if num_input_frames != len(stages[0]):
                raise AssertionError(f"t2v prefill replaces stage 0 = {len(stages[0])} frames")
            records.append(Prefill(source=(0, num_input_frames), out=stages[0], temporal=stages[0], slot=stages[0],
                                   t_len=len(stages[0])))
            skip = 1
        else:
            if num_input_frames > 2:
                raise AssertionError("i2v prefill is the image frame or the two connect frames")
            for f in range(num_input_frames):
                records.append(Prefill(source=(f, f + 1), out=stages[f], temporal=stages[f], slot=stages[f], t_len=1))
            skip = num_input_frames
    for s in range(skip, len(stages)):
        frames = stages[s]
        kw = {}
        if variant == "t2v" and s == 2:     # first gap: ends re-noised from frames 3 and 10, far anchors hidden (:281-303)
            kw = dict(renoise=((0, 3), (len(frames) - 1, 10)), hide=FAR_ANCHORS)
        elif variant == "t2v" and s == 3:   # second gap: ends from frames 12 and 19, far anchors visible again (:306-325)
            kw = dict(renoise=((0, 12), (len(frames) - 1, 19)), show=FAR_ANCHORS)
        if s == anchor_stage:               # next segment's anchors (:380-383; MMPL_i2v :340-342)
            kw["handoff"] = ((0, num_frames - 2, num_frames - 1), False) if variant == "i2v" else ((0,), True)
        records.append(Denoise(noise=frames, out=frames, temporal=frames, slot=frames, **kw))
    return RolloutPlan(tuple(records), out_frames=num_frames, sampler="unipc",
                       description=f"MMPL {variant}: stages {[len(s) for s in stages]}, first generated stage {skip}")
