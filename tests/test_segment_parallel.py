"""Segment-parallel scheduling on CPU with the gloo backend, world_size 2 (host-side logic of SURVEY.md §8e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmpl_b200.segment_parallel import (AnchorChannel, SegmentParallelRunner, default_segment_connect, producer_of,
                                        segments_of_rank)


def test_round_robin_placement():
    assert segments_of_rank(0, 4, 12) == [0, 4, 8] and segments_of_rank(3, 4, 12) == [3, 7, 11]
    assert segments_of_rank(1, 2, 3) == [1] and segments_of_rank(0, 8, 4) == [0] and segments_of_rank(5, 8, 4) == []
    assert [producer_of(s, 4) for s in range(6)] == [0, 1, 2, 3, 0, 1]


class FakeFPSPipeline:
    """Stands in for CausalFPSInferencePipeline: emits anchors after 'stage 1' and returns latents that encode the
    segment's dependency chain, so the test can check who received what."""

    def __init__(self):
        self.anchor_sink = None

    def inference(self, noise, text_prompts, initial_latent=None, return_latents=True):
        base = noise.clone()
        if initial_latent is not None:
            base[:, :2] = initial_latent          # prefilled first two frames (stage 0 replaced)
        out = base + 1.0
        anchors = torch.cat([out[:, :1], out[:, [2, 3, 10, 11, 12, 19, 20]]], dim=1)
        self.anchor_sink(anchors)
        return out, out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_segments, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        runner = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(), anchor_shape=(1, 8, 4, 2, 2))
        outs = runner.run(lambda seg: torch.full((1, 21, 4, 2, 2), float(seg * 100)), ["p"], num_segments)
        q.put((rank, {k: v[0, :, 0, 0, 0].tolist() for k, v in outs.items()}, runner.log, runner.channel.bytes_sent))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_segments", [3, 4])
def test_anchor_handoff_world2_gloo(num_segments):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_segments, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs, logs, sent = {}, {}, 0
    for rank, o, log, nbytes in res:
        assert sorted(o) == segments_of_rank(rank, 2, num_segments)
        outs.update(o)
        logs[rank] = log
        sent += nbytes
    # sequential reference: segment k's first two frames are frames 19, 20 of segment k-1's output
    prev = None
    for seg in range(num_segments):
        base = [float(seg * 100)] * 21
        if prev is not None:
            base[0], base[1] = prev[19], prev[20]
        expect = [b + 1.0 for b in base]
        assert outs[seg] == expect, (seg, outs[seg], expect)
        prev = expect
    assert sent == (num_segments - 1) * 8 * 4 * 2 * 2 * 4  # one anchor payload per boundary
    assert ("send", 0, 1) in logs[0] and ("recv", 1, 0) in logs[1]


def test_single_rank_runs_all_segments_locally():
    runner = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(), anchor_shape=(1, 8, 4, 2, 2))
    outs = runner.run(lambda seg: torch.zeros(1, 21, 4, 2, 2), ["p"], 3)
    assert sorted(outs) == [0, 1, 2] and outs[2][0, 0, 0, 0, 0].item() == 2.0 and outs[2][0, 5, 0, 0, 0].item() == 1.0
    a = torch.arange(8.).view(1, 8, 1, 1, 1)
    assert default_segment_connect(a).flatten().tolist() == [6.0, 7.0]
