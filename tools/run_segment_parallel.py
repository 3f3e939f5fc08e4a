#!/usr/bin/env python
"""MMPL segment-parallel long-video generation (BASELINE.json configs 3-5) with synthetic weights and inputs.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \\
      tools/run_segment_parallel.py --model 14B --segments 4 --sampling-steps 50

One process per GPU; segment k runs on rank k % G; anchors go rank -> rank over NCCL (mmpl_b200.segment_parallel).
Prints one JSON line with denoised latent frames/s per box = segments * 21 / wall time of the whole chain (device
time, max over ranks), the per-segment times and the anchor bytes moved. `--sampling-steps` below the reference's 50
shortens every stage proportionally (same schedule, same kernels, fewer UniPC steps) for quick scaling checks.
`--chains C` runs C independent videos on disjoint groups of world/C ranks (box throughput beyond the point where one
chain saturates); `--cfg-pair` splits every segment over two ranks (conditional / unconditional branch).
"""
import argparse
import json
import os
import sys
import time
import types

os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
# stdout carries JSON lines only: fd 1 is pointed at stderr for libraries (NCCL's version banner), emit() writes to the saved fd
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(record):
    os.write(_STDOUT_FD, (json.dumps(record) + "\n").encode())


import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmpl_b200.causal_model import CausalFPSWanModel  # noqa: E402
from mmpl_b200.pipeline import CausalFPSInferencePipeline  # noqa: E402
from mmpl_b200.segment_parallel import (AnchorChannel, I2V_ANCHOR_SHAPE, SegmentParallelRunner,  # noqa: E402
                                        T2V_ANCHOR_SHAPE, make_chain_groups, passthrough_connect)
from mmpl_b200.wan_wrapper import MODEL_CONFIGS, WanFPSWrapper  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="14B", choices=["14B", "1.3B"])
    ap.add_argument("--segments", type=int, default=4)
    ap.add_argument("--sweep", default="", help="comma-separated segment counts run back to back on one model build, one "
                                                 "JSON line each (e.g. 1,2,4,8,12 = 5-60 s videos); overrides --segments")
    ap.add_argument("--videos", type=int, default=0,
                    help="serve this many queued videos (of --segments segments each, own prompt and noise) through the resident "
                         "scheduler (mmpl_b200.segment_service): the box is cut into as many chains as videos are waiting")
    ap.add_argument("--min-slots", type=int, default=1, help="--videos: never cut a chain narrower than this many segment slots")
    ap.add_argument("--sampling-steps", type=int, default=50)
    ap.add_argument("--layers", type=int, default=0, help="override the number of blocks (0 = the model's own)")
    ap.add_argument("--i2v", action="store_true")
    ap.add_argument("--broadcast-weights", action="store_true",
                    help="rank 0's weights are sent to every rank over NCCL (one checkpoint load per box) instead of every rank "
                         "initialising its own identical replica; the time and bytes go to stderr")
    ap.add_argument("--vae-connect", action="store_true",
                    help="run the reference's VAE segment connect (decode anchors, frames 8:13, re-encode) on the hand-off "
                         "with a seeded random-init VAE (the checkpoint is absent) instead of passing the last two anchors through")
    ap.add_argument("--chains", type=int, default=1,
                    help="independent videos (own prompt / noise) on disjoint rank groups of world/chains ranks each: box throughput")
    ap.add_argument("--cfg-pair", action="store_true",
                    help="CFG-pair split: two ranks per segment (conditional / unconditional branch), world must be even")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    dims = dict(MODEL_CONFIGS["Wan2.1-T2V-14B" if a.model == "14B" else "Wan2.1-T2V-1.3B"])
    if a.layers:
        dims["num_layers"] = a.layers
    torch.manual_seed(0)  # identical random-init replica on every rank
    t0 = time.perf_counter()
    with torch.device(dev):
        model = CausalFPSWanModel(**dims)
    model = model.to(torch.bfloat16).eval().requires_grad_(False)
    if a.broadcast_weights and world > 1:
        from mmpl_b200.segment_parallel import broadcast_weights
        dist.broadcast(torch.zeros(8, device=dev), src=0)   # communicator set-up is not part of the transfer
        torch.cuda.synchronize()
        dist.barrier()
        tb = time.perf_counter()
        nbytes = broadcast_weights(model, src=0)
        torch.cuda.synchronize()
        if rank == 0:
            dt = time.perf_counter() - tb
            emit({"what": "broadcast_weights: rank 0's parameters to every rank over NCCL (one checkpoint load per box)",
                  "n_gpus": world, "gigabytes": nbytes / 1e9, "seconds": dt, "gb_per_s": nbytes / 1e9 / dt})
    gen = WanFPSWrapper(model=model, timestep_shift=5.0)
    build_s = time.perf_counter() - t0
    chain, chain_group, chain_ranks = (0, None, list(range(world)))
    if world > 1:
        chain, chain_group, chain_ranks = make_chain_groups(a.chains)
    elif a.chains != 1:
        raise SystemExit("--chains needs torchrun with a multiple of that many ranks")
    cworld, crank = len(chain_ranks), rank - chain_ranks[0]
    prompt = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(1 + 10 * chain)).to(torch.bfloat16).to(dev)
    negative = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(dev)

    prompt_table = {}  # --videos: one embedding per video's prompt

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            if text_prompts[0] in prompt_table:
                return {"prompt_embeds": prompt_table[text_prompts[0]]}
            return {"prompt_embeds": negative if text_prompts[0] == "__negative__" else prompt}

    class VAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="__negative__",
                                 independent_first_frame=False, sampling_steps=a.sampling_steps, model_kwargs={}, i2v=a.i2v)
    lanes = 2 if a.cfg_pair else 1
    cfg_group = None
    if a.cfg_pair:
        assert cworld % 2 == 0, "--cfg-pair needs an even number of ranks per chain"
        for s in range(world // 2):  # every rank creates every pair group, in the same order
            grp = dist.new_group([2 * s, 2 * s + 1])
            if rank // 2 == s:
                cfg_group = grp
    torch.manual_seed(1234)  # the pipeline draws its re-noising noise with torch.randn_like: same stream on both lanes
    pipe = CausalFPSInferencePipeline(args, dev, generator=gen, text_encoder=Text(), vae=VAE(), device_cond=dev, device_uncond=dev,
                                      cfg_group=cfg_group)
    channel = AnchorChannel(group=chain_group, lanes=lanes)
    if cfg_group is not None:  # create the pair's communicator outside the timed region
        warm = [torch.zeros(8, device=dev), torch.zeros(8, device=dev)]
        dist.all_gather(warm, torch.zeros(8, device=dev), group=cfg_group)
    if cworld > lanes:  # create the NCCL point-to-point connections outside the timed region
        buf = torch.zeros(8, device=dev)
        nxt, prv = chain_ranks[(crank + lanes) % cworld], chain_ranks[(crank - lanes) % cworld]
        ops = [dist.P2POp(dist.isend, buf, nxt, chain_group), dist.P2POp(dist.irecv, torch.empty_like(buf), prv, chain_group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    connect, connect_name = passthrough_connect, "pass-through of the last two anchors (benchmarking shortcut, NOT the reference transform)"
    if a.vae_connect:
        from mmpl_b200.segment_parallel import vae_segment_connect
        from mmpl_b200.vae import WanVAEWrapper
        vae = WanVAEWrapper()
        vae.init_random_weights(seed=0, device=dev)
        connect, connect_name = vae_segment_connect(vae), "VAE decode -> pixel frames 8:13 -> encode (random-init Wan VAE)"
    # the i2v schedule needs a first frame for segment 0 (VAE-encoded image in the reference)
    first = torch.randn(1, 1, 16, 60, 104, generator=torch.Generator().manual_seed(7)).to(torch.bfloat16).to(dev) if a.i2v else None
    runner = SegmentParallelRunner(pipe, channel, anchor_shape=I2V_ANCHOR_SHAPE if a.i2v else T2V_ANCHOR_SHAPE, connect=connect,
                                   first_initial=first)

    def make_noise(seg):
        g = torch.Generator().manual_seed(100 + seg + 1000 * chain)
        return torch.randn(1, 21, 16, 60, 104, generator=g).to(torch.bfloat16).to(dev)

    if a.videos:
        from mmpl_b200.segment_service import SegmentService, VideoJob
        assert a.chains == 1, "--videos places the chains itself"
        jobs = [VideoJob(f"video{i}", [f"synthetic prompt {i}"], a.segments, seed=i) for i in range(a.videos)]
        for j in jobs:
            prompt_table[j.prompts[0]] = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(10 + j.seed)).to(torch.bfloat16).to(dev)

        def job_noise(job, seg):
            g = torch.Generator().manual_seed(100 + seg + 1000 * job.seed)
            return torch.randn(1, 21, 16, 60, 104, generator=g).to(torch.bfloat16).to(dev)

        svc = SegmentService(pipe, job_noise, I2V_ANCHOR_SHAPE if a.i2v else T2V_ANCHOR_SHAPE, connect=connect, lanes=lanes,
                             min_slots=a.min_slots, first_initial=first)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = svc.serve(jobs if rank == 0 else None)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        finite = all(torch.isfinite(v.float()).all().item() for d in outs.values() for v in d.values())
        flags = [finite]
        if world > 1:
            flags = [None] * world
            dist.all_gather_object(flags, finite)
        if rank == 0:
            emit({
                "metric": "denoised_latent_frames_per_s", "unit": "latent frames/s", "n_gpus": world,
                "value": a.videos * a.segments * 21 / (ms.item() / 1e3), "ms_total": ms.item(),
                "config": {"workload": f"Wan2.1-{a.model} MMPL {'I2V' if a.i2v else 'T2V'}: {a.videos} queued videos x {a.segments} segments x 21 latent "
                                       f"frames 60x104 through the resident scheduler, {a.sampling_steps} UniPC steps x CFG, {dims['num_layers']} blocks",
                           "parallelism": f"chains chosen per round, {lanes} lane(s) per segment, min {a.min_slots} slot(s) per chain"},
                "finite": all(flags), "history": svc.history, "connect": connect_name, "model_build_s": build_s})
        if world > 1:
            dist.destroy_process_group()
        return

    def run_once(nseg):
        runner.log.clear()
        channel.bytes_sent = 0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        e0.record()
        model.launch_count(reset=True)
        outs = runner.run(make_noise, ["synthetic prompt"], nseg)
        e1.record()
        torch.cuda.synchronize()
        my_ms = e0.elapsed_time(e1)
        wall = time.perf_counter() - wall0
        if world > 1:
            dist.barrier()
        total_wall = time.perf_counter() - wall0
        ms = torch.tensor([my_ms], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        finite = all(torch.isfinite(v.float()).all().item() for v in outs.values())
        info = dict(rank=rank, chain=chain, segments=sorted(outs), ms=my_ms, launches=model.launch_count(), finite=finite,
                    anchor_bytes_sent=channel.bytes_sent, cfg_bytes_exchanged=getattr(pipe, "cfg_bytes_exchanged", 0), log=runner.log,
                    checksum={k: float(v.float().abs().sum()) for k, v in outs.items()})
        gathered = [info]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, info)
        if rank == 0:
            emit({
                "metric": "denoised_latent_frames_per_s", "unit": "latent frames/s", "n_gpus": world,
                "value": a.chains * nseg * 21 / (ms.item() / 1e3), "ms_total": ms.item(), "wall_s_incl_barrier": total_wall,
                "config": {"workload": f"Wan2.1-{a.model} MMPL {'I2V' if a.i2v else 'T2V'} segment-parallel, {a.chains} chain(s) x {nseg} segments x 21 latent frames 60x104, "
                                       f"stages {'[1,1,7,6,6]' if a.i2v else '[2,7,6,6]'}, {a.sampling_steps} UniPC steps x CFG, {dims['num_layers']} blocks",
                           "parallelism": ((f"{a.chains} independent chains, each " if a.chains > 1 else "") + f"segment-parallel x{cworld // lanes} slots" + (" x 2 CFG lanes (flow all-gather per step)" if a.cfg_pair else "") +
                                           ", anchors over NCCL send/recv")},
                "connect": connect_name, "model_build_s": build_s, "ranks": gathered})

    # --sweep: several chain lengths against one model build (BASELINE config 5: 5-60 s videos)
    for nseg in ([int(x) for x in a.sweep.split(',')] if a.sweep else [a.segments]):
        run_once(nseg)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
