"""CPU (gloo, world size 4): the resident segment scheduler (mmpl_b200/segment_service.py) places
queued videos on disjoint rank groups, reuses the groups across rounds, and every video comes out exactly as a
single-process run of the same job produces it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmpl_b200.segment_parallel import AnchorChannel, SegmentParallelRunner, passthrough_connect
from mmpl_b200.segment_service import SegmentService, VideoJob, plan_round

SHAPE = (1, 21, 4, 2, 2)
ANCHOR = (1, 8, 4, 2, 2)


class FakeFPSPipeline:
    """Stand-in for CausalFPSInferencePipeline: anchors after 'stage 1', output depends on noise, prompt and initial latent."""

    def __init__(self):
        self.anchor_sink = None

    def inference(self, noise, text_prompts, initial_latent=None, return_latents=True):
        base = noise.clone() + float(len(text_prompts[0]))
        if initial_latent is not None:
            base[:, :2] = initial_latent * 0.5
        out = base + 1.0
        self.anchor_sink(torch.cat([out[:, :1], out[:, [2, 3, 10, 11, 12, 19, 20]]], dim=1))
        return out, out


def make_noise(job, seg):
    return torch.randn(*SHAPE, generator=torch.Generator().manual_seed(1000 * job.seed + seg))


JOBS = [VideoJob("a", ["first"], 3, seed=1), VideoJob("b", ["second prompt"], 2, seed=2), VideoJob("c", ["third"], 4, seed=3),
        VideoJob("d", ["x"], 1, seed=4)]


LATE = [VideoJob("f", ["late one"], 3, seed=6), VideoJob("g", ["late two"], 3, seed=7)]


def test_plan_round_policy():
    j = JOBS
    assert plan_round([], 8) is None
    assert plan_round(j[:1], 8, lanes=2).chains == 1                       # a lone video gets the box
    p = plan_round(j[:3], 8, lanes=2)
    assert p.chains == 2 and [x.job_id for x in p.jobs] == ["a", "b"]      # FIFO, power-of-two split
    assert plan_round(j * 3, 8, lanes=2).chains == 4                       # one pair per video, never narrower than a pair
    assert plan_round(j * 3, 8, lanes=1).chains == 8
    assert plan_round(j * 3, 8, lanes=2, min_slots=2).chains == 2          # latency floor: at least 2 slots per chain
    assert plan_round(j, 6, lanes=1).chains == 2                           # 6 ranks: 4 does not divide, 2 does
    with pytest.raises(ValueError):
        plan_round(j, 7, lanes=2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        svc = SegmentService(FakeFPSPipeline(), make_noise, ANCHOR, connect=passthrough_connect)
        mine = svc.serve(JOBS if rank == 0 else None)
        # a second batch on the same resident service: groups and runners are reused
        mine2 = svc.serve([VideoJob("e", ["again"], 2, seed=5)] if rank == 0 else None)
        # two videos waiting: two chains of two ranks, anchors cross ranks inside each sub-group
        mine3 = svc.serve(LATE if rank == 0 else None)
        q.put((rank, {k: {s: v.float().numpy() for s, v in d.items()} for k, d in {**mine, **mine2, **mine3}.items()}, svc.history))
    finally:
        dist.destroy_process_group()


def test_service_world4_gloo():
    world, port = 4, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = {r: o for r, o, _ in got}
    history = next(h for r, _, h in got if r == 0)
    # placement: 4 jobs waiting -> 4 chains of one rank; then the late job alone on the whole box
    assert [(h["job_id"], h["chains"], h["ranks"]) for h in history] == [
        ("a", 4, [0]), ("b", 4, [1]), ("c", 4, [2]), ("d", 4, [3]), ("e", 1, [0, 1, 2, 3]), ("f", 2, [0, 1]), ("g", 2, [2, 3])]
    assert history[4]["latent_frames"] == 42
    # every video equals the single-process run of the same job
    for job in JOBS + [VideoJob("e", ["again"], 2, seed=5)] + LATE:
        want = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(), anchor_shape=ANCHOR, connect=passthrough_connect).run(
            lambda seg, job=job: make_noise(job, seg), job.prompts, job.num_segments)
        have = {}
        for r in range(world):
            have.update(by_rank[r].get(job.job_id, {}))
        assert sorted(have) == sorted(want) == list(range(job.num_segments)), job.job_id
        for seg in want:
            assert torch.equal(torch.from_numpy(have[seg]), want[seg].float()), (job.job_id, seg)


def _worker_lanes(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        svc = SegmentService(FakeFPSPipeline(), make_noise, ANCHOR, connect=passthrough_connect, lanes=2)
        mine = svc.serve([JOBS[2]] if rank == 0 else None)          # one video: 2 slots x 2 lanes, anchors go lane to lane
        mine2 = svc.serve(LATE if rank == 0 else None)              # two videos: one pair each
        q.put((rank, {k: {s: v.float().numpy() for s, v in d.items()} for k, d in {**mine, **mine2}.items()}, svc.history))
    finally:
        dist.destroy_process_group()


def test_service_with_cfg_pair_lanes_world4_gloo():
    """lanes=2: a segment occupies two consecutive ranks (conditional / unconditional branch) that hold the same latents."""
    world, port = 4, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_lanes, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = {r: o for r, o, _ in got}
    history = next(h for r, _, h in got if r == 0)
    assert [(h["job_id"], h["chains"], h["ranks"]) for h in history] == [("c", 1, [0, 1, 2, 3]), ("f", 2, [0, 1]), ("g", 2, [2, 3])]
    for job in [JOBS[2]] + LATE:
        want = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(), anchor_shape=ANCHOR, connect=passthrough_connect).run(
            lambda seg, job=job: make_noise(job, seg), job.prompts, job.num_segments)
        for lane in (0, 1):   # both lanes of every pair end with the same latents
            have = {}
            for r in range(lane, world, 2):
                have.update(by_rank[r].get(job.job_id, {}))
            assert sorted(have) == list(range(job.num_segments)), (job.job_id, lane)
            for seg in want:
                assert torch.equal(torch.from_numpy(have[seg]), want[seg].float()), (job.job_id, lane, seg)


def _worker_lazy(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        svc = SegmentService(FakeFPSPipeline(), make_noise, ANCHOR, connect=passthrough_connect)
        pulled = []

        def front_end():
            """Requests arrive over time: one job, then 'nothing ready' twice, then two jobs, then the front end closes."""
            for item in (JOBS[0], None, None, JOBS[1], JOBS[3]):
                pulled.append((None if item is None else item.job_id, len(svc.history)))
                yield item

        svc.serve(front_end() if rank == 0 else None)
        dup = None
        if rank == 0:
            try:
                svc._pull(iter([VideoJob("a", ["again"], 1)]), [], 1)
            except ValueError as e:
                dup = str(e)
        q.put((rank, svc.history, pulled, dup))
    finally:
        dist.destroy_process_group()


def test_jobs_are_pulled_lazily_between_rounds_world2_gloo():
    """The job source is consulted between rounds, never drained up front (a generator fed by a front end), an idle front
    end does not stop the service, and duplicate job ids are refused."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_lazy, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: rest for r, *rest in [q.get(timeout=120) for _ in range(world)]}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    history, pulled, dup = got[0]
    # job "a" ran alone on the whole box BEFORE the front end was asked for "b": the generator was not drained up front
    assert [(h["job_id"], h["chains"]) for h in history] == [("a", 1), ("b", 2), ("d", 2)]
    assert pulled[0] == ("a", 0) and ("b", 1) in pulled and ("d", 1) in pulled
    assert dup is not None and "duplicate job id" in dup
