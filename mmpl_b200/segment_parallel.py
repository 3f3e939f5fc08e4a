"""MMPL segment-parallel long-video generation over torch.distributed.

The reference runs one pipeline replica per GPU in Python threads and hands the anchor latents of segment k to segment
k+1 through a `.pt` file that the consumer polls every second (Wan_fps_inference_parallel_4gpu_20s.py:176-261;
pipeline/casual_fps_inference.py:380-383). Here every GPU is one process (torchrun), segment k runs on rank k % G, and
the hand-off is a point-to-point message on the communicator (NCCL over NVLink on GPUs: the send is enqueued on the
compute stream right after the anchor stage, so the producer continues with its next stage immediately; gloo in the CPU
tests). The path has no other exchange: each rank holds a full replica and its own caches.

With the CFG-pair split (`lanes=2`; CausalFPSInferencePipeline(cfg_group=...)) a segment runs on two consecutive
ranks -- lane 0 the conditional branch, lane 1 the unconditional one -- which hold identical latents, so each lane
hands its copy of the anchors to the same lane of the next segment's pair. One prompt's chain saturates at
T_segment / T_anchor ~ 3 segment slots; the pair split halves both times, which is what lets one
chain use 8 GPUs (4 slots x 2 lanes).

Box throughput beyond one chain's saturation point comes from independent chains (different prompts / seeds) on
disjoint rank groups (`make_chain_groups`): an AnchorChannel built on a chain's group addresses ranks relative to that
group, so a chain's schedule does not depend on where in the box it runs, and chains never exchange data.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

T2V_ANCHOR_SHAPE = (1, 8, 16, 60, 104)   # frame 0 + stage-1 frames [2,3,10,11,12,19,20]  (1.6 MB bf16)
I2V_ANCHOR_SHAPE = (1, 3, 16, 60, 104)   # frames 0, 19, 20                               (0.6 MB bf16)


def segments_of_rank(rank: int, world: int, num_segments: int, lanes: int = 1) -> List[int]:
    """Round-robin placement: slot s = rank // lanes runs segments s, s+G, s+2G, ... with G = world // lanes slots
    (the 5-60s driver's rotation, :231-238)."""
    return list(range(rank // lanes, num_segments, world // lanes))


def producer_of(segment: int, world: int, lanes: int = 1, lane: int = 0) -> int:
    """Rank that runs `segment` (its lane `lane` under the CFG-pair split)."""
    return (segment % (world // lanes)) * lanes + lane


def make_chain_groups(chains: int, world: Optional[int] = None, rank: Optional[int] = None):
    """Splits the world into `chains` contiguous rank groups of equal size (one independent video each). Every rank
    must call this (torch.distributed.new_group is collective); returns (chain index of this rank, its group, ranks of
    the group). With chains == 1 the group is None (= the default group)."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    if chains < 1 or world % chains:
        raise ValueError(f"world size {world} is not a multiple of chains={chains}")
    per = world // chains
    mine = rank // per
    if chains == 1:
        return 0, None, list(range(world))
    group = None
    for c in range(chains):  # same order on every rank
        g = dist.new_group(list(range(c * per, (c + 1) * per)))
        if c == mine:
            group = g
    return mine, group, list(range(mine * per, (mine + 1) * per))


@torch.no_grad()
def broadcast_weights(module: torch.nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None,
                      bucket_bytes: int = 1 << 28, direct_bytes: int = 1 << 24) -> int:
    """Loads a checkpoint once per box: rank `src` holds the weights (load_state_dict from disk), every other rank of the
    group receives them over the communicator (NCCL over NVLink: 28 GB at 14B) instead of reading the file eight times as
    the reference's per-GPU `from_pretrained` does (Wan_fps_inference_parallel_4gpu_20s.py:66-87).
    Contiguous tensors of `direct_bytes` or more (the projection matrices: 52-141 MB each at 14B, 99 % of the bytes) are
    broadcast in place, with no staging copy; the small ones (biases, norms, modulation) are flattened into buckets of
    `bucket_bytes` per dtype so that the ~1 000 of them cost a handful of collectives. Returns the bytes broadcast."""
    root = src if group is None else dist.get_global_rank(group, src)
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.numel()]
    nbytes = lambda t: t.numel() * t.element_size()
    small = []
    total = 0
    for t in tensors:
        if t.is_contiguous() and nbytes(t) >= direct_bytes:
            dist.broadcast(t, src=root, group=group)
            total += nbytes(t)
        else:
            small.append(t)
    i = 0
    while i < len(small):
        bucket, size = [small[i]], nbytes(small[i])
        i += 1
        while (i < len(small) and small[i].dtype == bucket[0].dtype and small[i].device == bucket[0].device
               and size + nbytes(small[i]) <= bucket_bytes):
            bucket.append(small[i])
            size += nbytes(small[i])
            i += 1
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.broadcast(flat, src=root, group=group)
        off = 0
        for t in bucket:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
        total += size
    return total


def passthrough_connect(anchors: torch.Tensor) -> torch.Tensor:
    """BENCHMARKING SHORTCUT, not the reference's transform: the last two anchor latents (frames 19, 20 of the previous
    segment) become the next segment's first two frames as they are. The reference decodes the anchors with the VAE, takes
    pixel frames 8:13 and re-encodes them (Wan_fps_inference_parallel_4gpu_20s.py:191-205) - that is `vae_segment_connect`.
    Results produced with this function must say so."""
    return anchors[:, -2:].contiguous()


def vae_segment_connect(vae) -> Callable[[torch.Tensor], torch.Tensor]:
    """The reference's segment connect itself (decode the anchors, pixel frames 8:13, re-encode, first 2 latents) on the
    hand-written VAE path: `vae` is a mmpl_b200.vae.WanVAEWrapper with weights bound. Pass as `connect=`."""
    return lambda anchors: vae.segment_connect(anchors)


class AnchorChannel:
    """Point-to-point hand-off of one segment's anchor latents to the rank that runs the next segment."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, lanes: int = 1):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world % lanes:
            raise ValueError(f"world size {self.world} is not a multiple of lanes={lanes}")
        self.lanes = lanes
        self.lane = self.rank % lanes
        # ranks below are relative to `group`; point-to-point calls take global ranks
        self._global = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        self._local: Dict[int, torch.Tensor] = {}
        self._pending = []
        self.bytes_sent = 0

    def send(self, anchors: torch.Tensor, segment: int) -> None:
        dst = producer_of(segment + 1, self.world, self.lanes, self.lane)
        payload = anchors.contiguous()
        self.bytes_sent += payload.numel() * payload.element_size()
        if dst == self.rank:
            self._local[segment + 1] = payload.clone()
            return
        # batched form (one-op batch): on NCCL an unbatched send on an eagerly initialised group is ordered like a collective
        # against everything else on that group; a batch is an independent point-to-point transfer
        works = dist.batch_isend_irecv([dist.P2POp(dist.isend, payload, self._global(dst), self.group, segment + 1)])
        self._pending.extend((w, payload) for w in works)

    def recv(self, segment: int, shape: Sequence[int], dtype: torch.dtype, device) -> torch.Tensor:
        src = producer_of(segment - 1, self.world, self.lanes, self.lane)
        if src == self.rank:
            return self._local.pop(segment)
        buf = torch.empty(tuple(shape), dtype=dtype, device=device)
        for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, buf, self._global(src), self.group, segment)]):
            w.wait()   # NCCL: orders the current stream after the transfer, the host does not block
        return buf

    def flush(self) -> None:
        for work, _ in self._pending:
            work.wait()
        self._pending.clear()


class SegmentParallelRunner:
    """Runs `num_segments` 21-frame segments of one long video across the ranks of the group.

    `pipeline` is a CausalFPSInferencePipeline (or anything with `inference(noise, text_prompts, initial_latent=...,
    return_latents=True)` and a settable `anchor_sink`). `connect` turns the anchors received from the previous segment
    into this segment's `initial_latent`: `vae_segment_connect(vae)` is what the reference does, `passthrough_connect` a
    labelled shortcut; there is no default, so that no result is produced with the shortcut by accident.
    `first_initial`: `initial_latent` of segment 0 (the VAE-encoded image of the i2v schedule). `observer(event, segment)`
    is called at "segment_start", "connect_start", "connect_end" and "segment_end" (timing hooks)."""

    def __init__(self, pipeline, channel: Optional[AnchorChannel], anchor_shape: Sequence[int],
                 connect: Callable[[torch.Tensor], torch.Tensor], first_initial: Optional[torch.Tensor] = None,
                 observer: Optional[Callable[[str, int], None]] = None):
        self.pipeline = pipeline
        self.channel = channel or AnchorChannel()
        self.anchor_shape = tuple(anchor_shape)
        self.connect = connect
        self.first_initial = first_initial
        self.observer = observer or (lambda event, segment: None)
        self.log: List[tuple] = []

    def run(self, make_noise: Callable[[int], torch.Tensor], text_prompts: List[str], num_segments: int) -> Dict[int, torch.Tensor]:
        ch = self.channel
        outputs: Dict[int, torch.Tensor] = {}
        for seg in segments_of_rank(ch.rank, ch.world, num_segments, ch.lanes):
            noise = make_noise(seg)
            self.observer("segment_start", seg)
            initial = self.first_initial
            if seg > 0:
                anchors = ch.recv(seg, (noise.shape[0],) + self.anchor_shape[1:3] + tuple(noise.shape[3:]), noise.dtype, noise.device)
                self.log.append(("recv", seg, producer_of(seg - 1, ch.world, ch.lanes, ch.lane)))
                self.observer("connect_start", seg)
                initial = self.connect(anchors)
                self.observer("connect_end", seg)
            has_next = seg + 1 < num_segments

            def sink(payload, seg=seg, has_next=has_next):
                if has_next:
                    ch.send(payload, seg)
                    self.log.append(("send", seg, producer_of(seg + 1, ch.world, ch.lanes, ch.lane)))

            self.pipeline.anchor_sink = sink
            _, latents = self.pipeline.inference(noise=noise, text_prompts=text_prompts, initial_latent=initial,
                                                 return_latents=True)
            self.observer("segment_end", seg)
            outputs[seg] = latents
        ch.flush()
        return outputs
