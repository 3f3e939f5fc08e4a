"""Long-lived segment scheduler for one box: what the reference server does per request with
one Python thread per chunk, `.pt` files and one-second polls (fastapi_parallel_t2v_server.py:509-653), as a set of
resident ranks — one process per GPU, weights and caches loaded once — that take video jobs from a queue and place them
on the box so that it stays full.

Why placement matters: one video's chain emits at most one segment per anchor stage and saturates at T_segment / T_anchor
~ 3-4 segment slots (measured 1.81 / 3.88 / 7.09 / 10.02 latent frames/s at 1 / 2 / 4 / 8 B200 with CFG-pair
lanes). So with several videos queued the box is split into independent chains on disjoint rank groups
(narrow chains are the efficient ones); with one video queued it gets the whole box.

Scope: scheduling only — no HTTP, storage, upload or callbacks (the reference's FastAPI / S3 / callback code is its
control plane, out of scope). The unit of work is `SegmentParallelRunner.run`; all exchanges are the ones it already has
plus one object broadcast per round (the plan) and one object gather (job results / timings).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, Iterator, List, Optional

import torch
import torch.distributed as dist

from .segment_parallel import AnchorChannel, SegmentParallelRunner


@dataclass
class VideoJob:
    """One long video: `num_segments` segments of 21 latent frames for one prompt; `seed` fixes its noise."""
    job_id: str
    prompts: List[str]
    num_segments: int
    seed: int = 0


@dataclass
class RoundPlan:
    """What the box does next: `chains` equal rank groups, group c runs jobs[c] (groups beyond len(jobs) idle)."""
    chains: int
    jobs: List[VideoJob] = field(default_factory=list)


def _pow2_floor(n: int) -> int:
    p = 1
    while 2 * p <= n:
        p *= 2
    return p


def plan_round(pending: List[VideoJob], world: int, lanes: int = 1, min_slots: int = 1) -> Optional[RoundPlan]:
    """Number of concurrent chains for the next round and the jobs they run (FIFO).

    Narrow chains are the efficient ones (measured on 8 x B200, DESIGN.md §6: one chain on 8 GPUs 10.0 latent frames/s, two
    chains on 4 GPUs each 2 x 7.1, four on 2 GPUs each 4 x 3.9), wide chains the fast ones for a single video. So the box
    is cut into as many chains as there are videos waiting -- down to `min_slots` segment slots per chain, the latency
    floor the operator wants -- and a lone video gets the whole box. Chain counts are powers of two that divide the
    number of slots, so every rank group the planner can ask for is created once, up front."""
    if not pending:
        return None
    if world % lanes:
        raise ValueError(f"world size {world} is not a multiple of lanes={lanes}")
    slots_total = world // lanes
    limit = max(1, slots_total // max(1, min_slots))
    chains = min(_pow2_floor(len(pending)), _pow2_floor(limit))
    while slots_total % chains:
        chains //= 2
    return RoundPlan(chains=chains, jobs=list(pending[:chains]))


class SegmentService:
    """Resident ranks serving VideoJobs. Every rank of the default process group constructs one and calls `serve`.

    `pipeline`: this rank's CausalFPSInferencePipeline (weights resident). `make_noise(job, segment)` returns a segment's
    noise on this rank's device. `connect`: the anchor transform - `vae_segment_connect(vae)` is the reference's,
    `passthrough_connect` a labelled shortcut (required, as for SegmentParallelRunner)."""

    def __init__(self, pipeline, make_noise: Callable[[VideoJob, int], torch.Tensor], anchor_shape,
                 connect: Callable[[torch.Tensor], torch.Tensor], lanes: int = 1, min_slots: int = 1,
                 first_initial: Optional[torch.Tensor] = None):
        self.pipeline, self.make_noise, self.anchor_shape = pipeline, make_noise, tuple(anchor_shape)
        self.lanes, self.min_slots, self.connect, self.first_initial = lanes, min_slots, connect, first_initial
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world % lanes:
            raise ValueError(f"world size {self.world} is not a multiple of lanes={lanes}")
        # every split the planner can choose, created once and collectively (new_group is collective and ordered)
        self._groups: Dict[int, Optional[dist.ProcessGroup]] = {}
        chains = 1
        while chains <= self.world // lanes and (self.world // lanes) % chains == 0:
            per = self.world // chains
            mine = None  # chains == 1: the default group
            if chains > 1:
                for c in range(chains):
                    g = dist.new_group(list(range(c * per, (c + 1) * per)))
                    if self.rank // per == c:
                        mine = g
            self._groups[chains] = mine
            chains *= 2
        self._runners: Dict[int, SegmentParallelRunner] = {}
        self.history: List[dict] = []   # rank 0: one record per finished job
        self._seen_ids = set()

    def _runner(self, chains: int) -> SegmentParallelRunner:
        if chains not in self._runners:
            channel = AnchorChannel(group=self._groups[chains], lanes=self.lanes)
            self._runners[chains] = SegmentParallelRunner(self.pipeline, channel, anchor_shape=self.anchor_shape,
                                                          connect=self.connect, first_initial=self.first_initial)
        return self._runners[chains]

    def _pull(self, source: Optional[Iterator[VideoJob]], pending: List[VideoJob], want: int) -> Optional[Iterator[VideoJob]]:
        """Rank 0: tops `pending` up to `want` jobs from the front end's iterator, one `next()` at a time - so a generator
        that yields requests as they arrive is consulted between rounds, never drained up front. A front end with nothing
        ready yields `None` (or raises StopIteration when it is closed); neither blocks the box."""
        while source is not None and len(pending) < want:
            try:
                job = next(source)
            except StopIteration:
                return None
            if job is None:   # nothing ready right now
                break
            if job.job_id in self._seen_ids:
                raise ValueError(f"duplicate job id {job.job_id!r}: results and history are keyed by it")
            self._seen_ids.add(job.job_id)
            pending.append(job)
        return source

    def serve(self, jobs: Optional[Iterable[VideoJob]] = None) -> Dict[str, Dict[int, torch.Tensor]]:
        """Runs until the front end is exhausted and the queue is empty. `jobs` is read on rank 0 only: a list, or a generator
        fed by whatever front end owns the requests (pulled lazily between rounds, see `_pull`). Returns
        {job_id: {segment: latents}} for the segments this rank produced."""
        source = iter(jobs) if (self.rank == 0 and jobs is not None) else None
        pending: List[VideoJob] = []
        mine: Dict[str, Dict[int, torch.Tensor]] = {}
        max_chains = max(self._groups)
        while True:
            if self.rank == 0:
                source = self._pull(source, pending, max_chains)
                if not pending and source is not None:   # the front end is open but idle: ask again instead of quitting
                    time.sleep(0.001)
                    source = self._pull(source, pending, max_chains)
            box = [(plan_round(pending, self.world, self.lanes, self.min_slots), source is not None) if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            plan, front_end_open = box[0]
            if plan is None:
                if front_end_open:
                    continue
                return mine
            per = self.world // plan.chains
            chain = self.rank // per
            job = plan.jobs[chain] if chain < len(plan.jobs) else None
            record = None
            if job is not None:
                runner = self._runner(plan.chains)
                runner.log.clear()
                t0 = time.perf_counter()
                outs = runner.run(lambda seg, job=job: self.make_noise(job, seg), job.prompts, job.num_segments)
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                mine.setdefault(job.job_id, {}).update(outs)
                record = {"job_id": job.job_id, "rank": self.rank, "chain": chain, "chains": plan.chains,
                          "segments": sorted(outs), "seconds": time.perf_counter() - t0}
            # results travel inside the chain's own group to its first rank (a chain that finishes early does not wait for
            # the others here); the round boundary itself is the plan broadcast above
            group = self._groups[plan.chains]
            lead = chain * per
            parts = [None] * per if self.rank == lead else None
            dist.gather_object(record, parts, dst=lead, group=group)
            summary = None
            if self.rank == lead and job is not None:
                summary = {"job_id": job.job_id, "chains": plan.chains, "ranks": sorted(r["rank"] for r in parts if r),
                           "seconds": max(r["seconds"] for r in parts if r), "latent_frames": 21 * job.num_segments}
            leads = [None] * self.world if self.rank == 0 else None
            dist.gather_object(summary, leads, dst=0)
            if self.rank == 0:
                self.history.extend(x for x in leads if x is not None)
                pending = pending[len(plan.jobs):]
