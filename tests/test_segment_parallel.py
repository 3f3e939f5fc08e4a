"""Segment-parallel scheduling on CPU with the gloo backend, world_size 2 (host-side logic of SURVEY.md §8e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmpl_b200.segment_parallel import (AnchorChannel, SegmentParallelRunner, passthrough_connect, make_chain_groups,
                                        producer_of, segments_of_rank)


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
        runner = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(), anchor_shape=(1, 8, 4, 2, 2), connect=passthrough_connect)
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


def _chain_worker(rank, world, port, chains, num_segments, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chain, group, ranks = make_chain_groups(chains)
        runner = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(group=group), anchor_shape=(1, 8, 4, 2, 2), connect=passthrough_connect)
        outs = runner.run(lambda seg: torch.full((1, 21, 4, 2, 2), float(chain * 1000 + seg * 100)), ["p"], num_segments)
        q.put((rank, chain, ranks, {k: v[0, :, 0, 0, 0].tolist() for k, v in outs.items()}, runner.log))
    finally:
        dist.destroy_process_group()


def test_independent_chains_world4_gloo():
    """Two independent chains on disjoint rank groups of a 4-rank world (2 segment slots each): every chain follows the
    same group-relative schedule, its anchors stay inside its group, and its result equals the sequential chain."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    num_segments = 3
    procs = [ctx.Process(target=_chain_worker, args=(r, 4, port, 2, num_segments, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs = {0: {}, 1: {}}
    for rank, chain, ranks, o, log in res:
        assert chain == rank // 2 and ranks == [2 * chain, 2 * chain + 1]
        assert sorted(o) == segments_of_rank(rank % 2, 2, num_segments)
        assert all(peer in (0, 1) for _, _, peer in log)  # group-relative peers only
        outs[chain].update(o)
    for chain in (0, 1):
        prev = None
        for seg in range(num_segments):
            base = [float(chain * 1000 + seg * 100)] * 21
            if prev is not None:
                base[0], base[1] = prev[19], prev[20]
            expect = [b + 1.0 for b in base]
            assert outs[chain][seg] == expect, (chain, seg)
            prev = expect


def test_single_rank_runs_all_segments_locally():
    runner = SegmentParallelRunner(FakeFPSPipeline(), AnchorChannel(), anchor_shape=(1, 8, 4, 2, 2), connect=passthrough_connect)
    outs = runner.run(lambda seg: torch.zeros(1, 21, 4, 2, 2), ["p"], 3)
    assert sorted(outs) == [0, 1, 2] and outs[2][0, 0, 0, 0, 0].item() == 2.0 and outs[2][0, 5, 0, 0, 0].item() == 1.0
    a = torch.arange(8.).view(1, 8, 1, 1, 1)
    assert passthrough_connect(a).flatten().tolist() == [6.0, 7.0]


# ------------------------------------------------------------------------------------------------------------------
# CFG-pair split: conditional / unconditional branch on two ranks, flow all-gather per step
def _fps_pipeline(cfg_group=None):
    import types

    from mmpl_b200.pipeline import CausalFPSInferencePipeline
    from _cpu_ops import cpu_scheduler, eager_unipc_factory

    class FakeGenerator(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = types.SimpleNamespace(num_layers=3, local_attn_size=-1, num_heads=2, dim=256, text_len=32,
                                               num_frame_per_block=1)
            self.scheduler = cpu_scheduler()

        def get_scheduler(self):
            return self.scheduler

    gen = FakeGenerator()
    tags = []

    def fwd(noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache, current_start, cache_start):
        tags.append(conditional_dict["tag"])
        assert kv_cache is not None and crossattn_cache is not None
        sign = 1.0 if conditional_dict["tag"] == "pos" else -0.5
        return noisy_image_or_video * 0.1 + sign * 0.01 * float(timestep.flatten()[0]) / 1000.0, None

    gen.forward = fwd
    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="neg",
                                 independent_first_frame=False, sampling_steps=3, model_kwargs={})
    text = lambda text_prompts: {"tag": "neg" if text_prompts[0] == "neg" else "pos"}  # noqa: E731
    vae = types.SimpleNamespace(decode_to_pixel=lambda latents, use_cache=False: latents)
    torch.manual_seed(11)  # constructor randint + the re-noising randn_like draws: same stream on both lanes
    pipe = CausalFPSInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=text, vae=vae,
                                      device_cond="cpu", device_uncond="cpu", cfg_group=cfg_group)
    pipe.unipc_stepper = eager_unipc_factory(pipe)
    return pipe, tags


def _fps_run(pipe):
    noise = torch.randn(1, 21, 16, 8, 12, generator=torch.Generator().manual_seed(3))
    sent = []
    pipe.anchor_sink = sent.append
    _, latents = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    return latents, sent


def _cfg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipe, tags = _fps_pipeline(cfg_group=dist.group.WORLD)
        latents, sent = _fps_run(pipe)
        # numpy arrays travel through the queue by value; torch tensors would go through shared-memory handles that die
        # with this process (a race with the parent's q.get)
        q.put((rank, latents.float().numpy(), sent[0].float().numpy(), sorted(set(tags)), len(tags), pipe.kv_cache_pos is None,
               pipe.kv_cache_neg is None, pipe.cfg_bytes_exchanged))
    finally:
        dist.destroy_process_group()


def test_cfg_pair_split_world2_gloo():
    """Two ranks, one per CFG branch, must reproduce the single-process pipeline bit for bit; each rank runs only
    its branch's forwards and holds only its branch's caches; one flow exchange per denoising step."""
    pipe, tags = _fps_pipeline()
    ref_latents, ref_sent = _fps_run(pipe)
    n_forwards = len(tags)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cfg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, latents, anchors, seen, n, pos_none, neg_none, nbytes in res:
        assert torch.equal(torch.from_numpy(latents), ref_latents.float()), f"rank {rank}: latents differ from the single-process run"
        assert torch.equal(torch.from_numpy(anchors), ref_sent[0].float())
        assert seen == (["pos"] if rank == 0 else ["neg"]) and n == n_forwards // 2
        assert (pos_none, neg_none) == ((False, True) if rank == 0 else (True, False))
        # 4 stages x 3 steps, flow of n frames [1, n, 16, 8, 12] fp32 in this CPU test
        assert nbytes == 3 * (2 + 7 + 6 + 6) * 16 * 8 * 12 * 4


def test_lane_placement():
    # 8 ranks, 2 lanes: 4 segment slots; lane l of slot s is rank 2s + l and hands over to lane l of the next slot
    assert segments_of_rank(0, 8, 6, lanes=2) == [0, 4] and segments_of_rank(1, 8, 6, lanes=2) == [0, 4]
    assert segments_of_rank(7, 8, 6, lanes=2) == [3] and segments_of_rank(5, 8, 6, lanes=2) == [2]
    assert [producer_of(s, 8, 2, 1) for s in range(5)] == [1, 3, 5, 7, 1]


# ------------------------------------------------------------------------------------------------------------------
# one checkpoint load per box: rank 0's weights reach every rank over the communicator
def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmpl_b200.segment_parallel import broadcast_weights
        torch.manual_seed(100 + rank)   # every rank starts from different weights
        m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32), torch.nn.Linear(32, 8)).to(torch.bfloat16)
        m.register_buffer("table", torch.randn(5, 3, dtype=torch.float64))
        n = broadcast_weights(m, src=0, bucket_bytes=600, direct_bytes=1024)   # small buckets + one tensor sent in place, mixed dtypes
        # raw bytes travel through the queue by value (torch tensors would go through shared-memory handles that die with
        # this process: a race with the parent's q.get)
        q.put((rank, n, {k: (str(v.dtype), tuple(v.shape), v.contiguous().view(torch.uint8).numpy().tobytes())
                         for k, v in m.state_dict().items()}))
    finally:
        dist.destroy_process_group()


def test_broadcast_weights_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (n, sd)) for r, n, sd in [q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32), torch.nn.Linear(32, 8)).to(torch.bfloat16)
    ref.register_buffer("table", torch.randn(5, 3, dtype=torch.float64))
    want = ref.state_dict()
    assert got[0][0] == got[1][0] == sum(v.numel() * v.element_size() for v in want.values())
    for r in (0, 1):
        assert sorted(got[r][1]) == sorted(want)
        for k in want:
            dtype, shape, raw = got[r][1][k]
            assert dtype == str(want[k].dtype) and shape == tuple(want[k].shape)
            assert raw == want[k].contiguous().view(torch.uint8).numpy().tobytes(), (r, k)
