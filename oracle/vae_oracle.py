"""ORACLE (test infrastructure, never imported by the product) — CPU restatement of the Wan2.1 causal video VAE as MMPL
uses it either side of the anchor hand-off: SURVEY.md §8(f) row 2, the "segment connect"
(Wan_fps_inference_parallel_4gpu_20s.py:191-205) and the wrapper calls around it (utils/wan_wrapper.py:74-113).

Everything is a pure function over the reference's own state-dict (`WanVAE_.state_dict()` names: `encoder.conv1.weight`,
`decoder.upsamples.3.resample.1.weight`, ...), so a real `Wan2.1_VAE.pth` loads unchanged. Two formulations:

* `encode` / `decode`  — the reference's *streaming* schedule restated call for call: the encoder sees the pixel frames
  as chunks of 1, 4, 4, ... and the decoder one latent frame at a time, every causal convolution carrying its last two
  input frames between chunks (wan/modules/vae.py:16-36,189-213,513-566). Same torch operators in the same order as the
  reference, so on the CPU it is bit-identical to it (tests/test_vae_oracle.py against goldens recorded from the
  unmodified reference by oracle/make_golden_vae.py).
* `decode_whole` / `encode_whole` — the same mathematics as single passes over the whole frame axis (what a B200 kernel
  sequence would run: one launch per layer instead of one per layer per frame). The carried frames are exactly the causal
  left context, with two quirks that the whole-sequence form has to keep: the first latent frame is never temporally
  up-sampled and the temporal up-sampling convolution of the *second* frame sees zeros, not frame 0, as its history
  (the 'Rep' marker, vae.py:102-131); symmetrically the encoder's temporal down-sampling passes frame 0 through and
  then strides over [previous chunk's last frame, chunk] (vae.py:139-155).

`segment_connect` restates the driver's hand-off transform and `segment_connect_causal` the reduced form the B200 path
is designed around: because both halves are causal, pixel frames 8..12 depend on the first 4 latent frames only and
latents 0..1 on the first 5 pixel frames only, so 4 of 21 decoder steps and 2 of 21 encoder steps give the same bits.

Parity: PINNED (goldens from the unmodified reference VAE, random init, fp32 and bf16; the reference has no tests of its
own for this path).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# utils/wan_wrapper.py:52-63 (== wan/modules/vae.py:645-652)
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


@dataclass
class VaeConfig:
    """vae.py:603-611 (`_video_vae` defaults)."""
    dim: int = 96
    z_dim: int = 16
    dim_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temporal_downsample: Tuple[bool, ...] = (False, True, True)


# ------------------------------------------------------------------------------------------------ layer programs
# A "program" is the flat list of layers the reference's nn.Sequential containers hold, as (kind, state-dict prefix, extra).

def encoder_program(cfg: VaeConfig) -> List[tuple]:
    """vae.py:284-310: per level `num_res_blocks` residual blocks, then a down-sampler except after the last level."""
    dims = [cfg.dim * u for u in (1,) + tuple(cfg.dim_mult)]
    prog, i = [], 0
    for lvl, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg.num_res_blocks):
            prog.append(("res", f"encoder.downsamples.{i}", (cin, cout)))
            cin = cout
            i += 1
        if lvl != len(cfg.dim_mult) - 1:
            prog.append(("down3d" if cfg.temporal_downsample[lvl] else "down2d", f"encoder.downsamples.{i}", cout))
            i += 1
    prog += [("res", "encoder.middle.0", (cout, cout)), ("attn", "encoder.middle.1", cout), ("res", "encoder.middle.2", (cout, cout))]
    return prog


def decoder_program(cfg: VaeConfig) -> List[tuple]:
    """vae.py:384-415: `num_res_blocks + 1` residual blocks per level; levels 1..3 start from half the channels because
    the up-sampler before them halves them."""
    dims = [cfg.dim * u for u in (cfg.dim_mult[-1],) + tuple(cfg.dim_mult[::-1])]
    t_up = tuple(cfg.temporal_downsample[::-1])
    prog = [("res", "decoder.middle.0", (dims[0], dims[0])), ("attn", "decoder.middle.1", dims[0]),
            ("res", "decoder.middle.2", (dims[0], dims[0]))]
    i = 0
    for lvl, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if lvl in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg.num_res_blocks + 1):
            prog.append(("res", f"decoder.upsamples.{i}", (cin, cout)))
            cin = cout
            i += 1
        if lvl != len(cfg.dim_mult) - 1:
            prog.append(("up3d" if t_up[lvl] else "up2d", f"decoder.upsamples.{i}", cout))
            i += 1
    return prog


def make_weights(cfg: VaeConfig = VaeConfig(), seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded random weights with the reference's state-dict names and shapes (the checkpoint is absent: SURVEY.md §8d).
    Fan-in-scaled normals keep activations O(1) through the ~60 convolutions; gammas near 1; the attention output
    projection is NOT zeroed (vae.py:239 zeroes it at init, a trained checkpoint does not) so the attention block counts."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, Tensor] = {}

    def conv(name, cout, cin, *k):
        fan = cin
        for s in k:
            fan *= s
        w[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * fan ** -0.5
        w[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def gamma(name, c, nd):
        w[name] = 1.0 + 0.1 * torch.randn(c, *([1] * nd), generator=g)

    def res(p, cin, cout):
        gamma(p + ".residual.0.gamma", cin, 3)
        conv(p + ".residual.2", cout, cin, 3, 3, 3)
        gamma(p + ".residual.3.gamma", cout, 3)
        conv(p + ".residual.6", cout, cout, 3, 3, 3)
        if cin != cout:
            conv(p + ".shortcut", cout, cin, 1, 1, 1)

    def program(prog):
        for kind, p, x in prog:
            if kind == "res":
                res(p, *x)
            elif kind == "attn":
                gamma(p + ".norm.gamma", x, 2)
                conv(p + ".to_qkv", 3 * x, x, 1, 1)
                conv(p + ".proj", x, x, 1, 1)
            elif kind in ("down2d", "down3d"):
                conv(p + ".resample.1", x, x, 3, 3)
                if kind == "down3d":
                    conv(p + ".time_conv", x, x, 3, 1, 1)
            else:
                conv(p + ".resample.1", x // 2, x, 3, 3)
                if kind == "up3d":
                    conv(p + ".time_conv", 2 * x, x, 3, 1, 1)

    top = cfg.dim * cfg.dim_mult[-1]
    conv("encoder.conv1", cfg.dim, 3, 3, 3, 3)
    program(encoder_program(cfg))
    gamma("encoder.head.0.gamma", top, 3)
    conv("encoder.head.2", 2 * cfg.z_dim, top, 3, 3, 3)
    conv("conv1", 2 * cfg.z_dim, 2 * cfg.z_dim, 1, 1, 1)
    conv("conv2", cfg.z_dim, cfg.z_dim, 1, 1, 1)
    conv("decoder.conv1", top, cfg.z_dim, 3, 3, 3)
    program(decoder_program(cfg))
    gamma("decoder.head.0.gamma", cfg.dim, 3)
    conv("decoder.head.2", 3, cfg.dim, 3, 3, 3)
    return {k: v.to(dtype) for k, v in w.items()}


# ------------------------------------------------------------------------------------------------------ primitives

def causal_conv3d(x: Tensor, w: Tensor, b: Tensor, history: Optional[Tensor] = None, stride=(1, 1, 1)) -> Tensor:
    """vae.py:16-36. x [B,C,T,H,W]. All temporal padding goes in front (2*pad_t = kt-1 frames); `history` frames carried
    from the previous chunk replace that many zero frames. Spatial padding is symmetric ((k-1)/2)."""
    kt, kh, kw = w.shape[2:]
    pt = kt - 1 if stride[0] == 1 else 0  # the stride-2 temporal conv is built with padding 0 (vae.py:92-93)
    if history is not None and pt > 0:
        x = torch.cat([history.to(x.device), x], dim=2)
        pt -= history.shape[2]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pt, 0))
    return F.conv3d(x, w, b, stride=stride)


def rms_norm(x: Tensor, gamma: Tensor, dim: int = 1) -> Tensor:
    """vae.py:39-55: L2-normalise over channels (eps 1e-12 inside F.normalize), times sqrt(C), times gamma."""
    return F.normalize(x, dim=dim) * (x.shape[dim] ** 0.5) * gamma


def per_frame(fn, x: Tensor) -> Tensor:
    """'b c t h w -> (b t) c h w', fn, and back (vae.py:133-136, 245-262)."""
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def attention_block(W: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """vae.py:241-266: single-head attention over the h*w positions of each frame, residual."""

    def one(f):
        n, c, h, w = f.shape
        f = rms_norm(f, W[p + ".norm.gamma"])
        qkv = F.conv2d(f, W[p + ".to_qkv.weight"], W[p + ".to_qkv.bias"])
        q, k, v = qkv.reshape(n, 1, 3 * c, h * w).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.squeeze(1).permute(0, 2, 1).reshape(n, c, h, w)
        return F.conv2d(o, W[p + ".proj.weight"], W[p + ".proj.bias"])

    return per_frame(one, x) + x


def upsample2x_conv(W, p, x):
    """vae.py:58-64,75-78: nearest 2x in fp32, cast back, 3x3 conv to half the channels."""
    def one(f):
        f = F.interpolate(f.float(), scale_factor=(2.0, 2.0), mode="nearest").type_as(f)
        return F.conv2d(f, W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], padding=1)
    return per_frame(one, x)


def downsample2x_conv(W, p, x):
    """vae.py:85-88: zero-pad right/bottom by one, 3x3 conv stride 2."""
    def one(f):
        return F.conv2d(F.pad(f, (0, 1, 0, 1)), W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], stride=2)
    return per_frame(one, x)


def interleave_time(x: Tensor) -> Tensor:
    """vae.py:127-130: the 2C channels of the temporal up-sampling conv are two frames: [B,2,C,T,H,W] -> [B,C,2T,H,W]."""
    b, c2, t, h, w = x.shape
    x = x.reshape(b, 2, c2 // 2, t, h, w)
    return torch.stack((x[:, 0], x[:, 1]), 3).reshape(b, c2 // 2, 2 * t, h, w)


# ------------------------------------------------------------------------------------------- streaming formulation

REP = "Rep"  # vae.py:104: marks "first chunk seen, nothing to carry" in the temporal up-sampler's slot


@dataclass
class StreamState:
    """One slot per CausalConv3d in module order (vae.py:582-590); `i` is the running slot index of the current chunk."""
    slots: List[object]
    i: int = 0

    @staticmethod
    def new(n: int) -> "StreamState":
        return StreamState([None] * n)


def _carried_conv(W, name, x, st: StreamState) -> Tensor:
    """The pattern repeated at vae.py:197-209, 317-331, 351-364: remember the last two input frames (topped up with
    the previous chunk's last frame when the chunk has one frame), convolve with the previous chunk's frames as history."""
    prev = st.slots[st.i]
    keep = x[:, :, -2:].clone()
    if keep.shape[2] < 2 and prev is not None:
        keep = torch.cat([prev[:, :, -1:].to(keep.device), keep], dim=2)
    y = causal_conv3d(x, W[name + ".weight"], W[name + ".bias"], prev)
    st.slots[st.i] = keep
    st.i += 1
    return y


def _res_block(W, p, x, st: StreamState) -> Tensor:
    """vae.py:173-213. The 1x1x1 shortcut carries nothing."""
    h = causal_conv3d(x, W[p + ".shortcut.weight"], W[p + ".shortcut.bias"]) if (p + ".shortcut.weight") in W else x
    x = F.silu(rms_norm(x, W[p + ".residual.0.gamma"]))
    x = _carried_conv(W, p + ".residual.2", x, st)
    x = F.silu(rms_norm(x, W[p + ".residual.3.gamma"]))
    x = _carried_conv(W, p + ".residual.6", x, st)
    return x + h


def _up(W, p, x, st: StreamState, temporal: bool) -> Tensor:
    """vae.py:98-136 ('upsample2d' / 'upsample3d')."""
    if temporal:
        prev = st.slots[st.i]
        if prev is None:
            st.slots[st.i] = REP  # first chunk: no temporal up-sampling at all
        else:
            keep = x[:, :, -2:].clone()
            if keep.shape[2] < 2:
                head = torch.zeros_like(keep) if isinstance(prev, str) else prev[:, :, -1:].to(keep.device)
                keep = torch.cat([head, keep], dim=2)
            x = causal_conv3d(x, W[p + ".time_conv.weight"], W[p + ".time_conv.bias"], None if isinstance(prev, str) else prev)
            st.slots[st.i] = keep
            x = interleave_time(x)
        st.i += 1
    return upsample2x_conv(W, p, x)


def _down(W, p, x, st: StreamState, temporal: bool) -> Tensor:
    """vae.py:133-155 ('downsample2d' / 'downsample3d')."""
    x = downsample2x_conv(W, p, x)
    if temporal:
        prev = st.slots[st.i]
        if prev is None:
            st.slots[st.i] = x.clone()  # first chunk passes through
        else:
            keep = x[:, :, -1:].clone()
            x = causal_conv3d(torch.cat([prev[:, :, -1:], x], 2), W[p + ".time_conv.weight"], W[p + ".time_conv.bias"],
                              stride=(2, 1, 1))
            st.slots[st.i] = keep
        st.i += 1
    return x


def _run(W, prog, x, st):
    for kind, p, _ in prog:
        if kind == "res":
            x = _res_block(W, p, x, st)
        elif kind == "attn":
            x = attention_block(W, p, x)
        elif kind in ("up2d", "up3d"):
            x = _up(W, p, x, st, kind == "up3d")
        else:
            x = _down(W, p, x, st, kind == "down3d")
    return x


def n_slots(prog) -> int:
    """count_conv3d (vae.py:441-446): every CausalConv3d instance, 1x1x1 shortcuts included (their slots stay unused)."""
    n = 2  # conv1 + head conv
    for kind, _, x in prog:
        if kind == "res":
            n += 2 + (1 if x[0] != x[1] else 0)
        elif kind in ("up3d", "down3d"):
            n += 1
    return n


def encoder_chunk(W, cfg: VaeConfig, x: Tensor, st: StreamState) -> Tensor:
    """Encoder3d.forward with a feature cache (vae.py:313-366)."""
    st.i = 0
    x = _carried_conv(W, "encoder.conv1", x, st)
    x = _run(W, encoder_program(cfg), x, st)
    x = F.silu(rms_norm(x, W["encoder.head.0.gamma"]))
    return _carried_conv(W, "encoder.head.2", x, st)


def decoder_chunk(W, cfg: VaeConfig, x: Tensor, st: StreamState) -> Tensor:
    """Decoder3d.forward with a feature cache (vae.py:417-438... same file, decoder half)."""
    st.i = 0
    x = _carried_conv(W, "decoder.conv1", x, st)
    x = _run(W, decoder_program(cfg), x, st)
    x = F.silu(rms_norm(x, W["decoder.head.0.gamma"]))
    return _carried_conv(W, "decoder.head.2", x, st)


def _scale(cfg, ref: Tensor):
    mean = torch.tensor(LATENT_MEAN[:cfg.z_dim], dtype=torch.float32).to(device=ref.device, dtype=ref.dtype)
    inv_std = 1.0 / torch.tensor(LATENT_STD[:cfg.z_dim], dtype=torch.float32).to(device=ref.device, dtype=ref.dtype)
    return mean.view(1, -1, 1, 1, 1), inv_std.view(1, -1, 1, 1, 1)


@torch.no_grad()
def encode(W, cfg: VaeConfig, pixels: Tensor, max_chunks: Optional[int] = None) -> Tensor:
    """WanVAE_.encode (vae.py:501-528). pixels [B,3,T,H,W] in [-1,1], T = 1 + 4k -> normalised mu [B,z,1+k,H/8,W/8].
    `max_chunks` stops after that many chunks (chunk 0 is one frame, the others four)."""
    st = StreamState.new(n_slots(encoder_program(cfg)))
    n = 1 + (pixels.shape[2] - 1) // 4
    if max_chunks is not None:
        n = min(n, max_chunks)
    outs = []
    for i in range(n):
        chunk = pixels[:, :, :1] if i == 0 else pixels[:, :, 1 + 4 * (i - 1):1 + 4 * i]
        outs.append(encoder_chunk(W, cfg, chunk, st))
    out = torch.cat(outs, 2)
    mu, _ = causal_conv3d(out, W["conv1.weight"], W["conv1.bias"]).chunk(2, dim=1)
    mean, inv_std = _scale(cfg, mu)
    return (mu - mean) * inv_std


@torch.no_grad()
def decode(W, cfg: VaeConfig, z: Tensor, max_frames: Optional[int] = None) -> Tensor:
    """WanVAE_.decode (vae.py:530-552). z [B,z,T,h,w] normalised latents -> pixels [B,3,1+4(T-1),8h,8w] (not clamped)."""
    mean, inv_std = _scale(cfg, z)
    z = z / inv_std + mean
    x = causal_conv3d(z, W["conv2.weight"], W["conv2.bias"])
    st = StreamState.new(n_slots(decoder_program(cfg)))
    n = x.shape[2] if max_frames is None else min(x.shape[2], max_frames)
    return torch.cat([decoder_chunk(W, cfg, x[:, :, i:i + 1], st) for i in range(n)], 2)


def encode_to_latent(W, cfg, pixel: Tensor, **kw) -> Tensor:
    """WanVAEWrapper.encode_to_latent (utils/wan_wrapper.py:74-89): [B,3,T,H,W] -> fp32 [B,T',z,h,w]."""
    return torch.stack([encode(W, cfg, u.unsqueeze(0), **kw).float().squeeze(0) for u in pixel]).permute(0, 2, 1, 3, 4)


def decode_to_pixel(W, cfg, latent: Tensor, **kw) -> Tensor:
    """WanVAEWrapper.decode_to_pixel (utils/wan_wrapper.py:91-113): [B,T,z,h,w] -> fp32 [B,T',3,H,W] clamped to [-1,1]."""
    zs = latent.permute(0, 2, 1, 3, 4)
    out = [decode(W, cfg, u.unsqueeze(0), **kw).float().clamp_(-1, 1).squeeze(0) for u in zs]
    return torch.stack(out).permute(0, 2, 1, 3, 4)


# ---------------------------------------------------------------------------------------- whole-sequence formulation

def _res_block_whole(W, p, x):
    h = causal_conv3d(x, W[p + ".shortcut.weight"], W[p + ".shortcut.bias"]) if (p + ".shortcut.weight") in W else x
    x = F.silu(rms_norm(x, W[p + ".residual.0.gamma"]))
    x = causal_conv3d(x, W[p + ".residual.2.weight"], W[p + ".residual.2.bias"])
    x = F.silu(rms_norm(x, W[p + ".residual.3.gamma"]))
    x = causal_conv3d(x, W[p + ".residual.6.weight"], W[p + ".residual.6.bias"])
    return x + h


@torch.no_grad()
def decode_whole(W, cfg: VaeConfig, z: Tensor) -> Tensor:
    """`decode` as one pass per layer over all T latent frames. Temporal up-sampling: frame 0 is kept as is, frames
    1.. go through the (3,1,1) convolution with ZERO history in front of frame 1 (the 'Rep' quirk) and each becomes two."""
    mean, inv_std = _scale(cfg, z)
    x = causal_conv3d(z / inv_std + mean, W["conv2.weight"], W["conv2.bias"])
    x = causal_conv3d(x, W["decoder.conv1.weight"], W["decoder.conv1.bias"])
    for kind, p, _ in decoder_program(cfg):
        if kind == "res":
            x = _res_block_whole(W, p, x)
        elif kind == "attn":
            x = attention_block(W, p, x)
        else:
            if kind == "up3d" and x.shape[2] > 1:
                rest = interleave_time(causal_conv3d(x[:, :, 1:], W[p + ".time_conv.weight"], W[p + ".time_conv.bias"]))
                x = torch.cat([x[:, :, :1], rest], 2)
            x = upsample2x_conv(W, p, x)
    x = F.silu(rms_norm(x, W["decoder.head.0.gamma"]))
    return causal_conv3d(x, W["decoder.head.2.weight"], W["decoder.head.2.bias"])


@torch.no_grad()
def encode_whole(W, cfg: VaeConfig, pixels: Tensor) -> Tensor:
    """`encode` as one pass per layer over all T = 1 + 4k pixel frames. Temporal down-sampling: frame 0 passes through,
    the rest is a stride-2 (3,1,1) convolution over [frame 0 .. ] without padding: windows (0,1,2), (2,3,4), ..."""
    x = causal_conv3d(pixels, W["encoder.conv1.weight"], W["encoder.conv1.bias"])
    for kind, p, _ in encoder_program(cfg):
        if kind == "res":
            x = _res_block_whole(W, p, x)
        elif kind == "attn":
            x = attention_block(W, p, x)
        else:
            x = downsample2x_conv(W, p, x)
            if kind == "down3d" and x.shape[2] > 1:
                rest = causal_conv3d(x, W[p + ".time_conv.weight"], W[p + ".time_conv.bias"], stride=(2, 1, 1))
                x = torch.cat([x[:, :, :1], rest], 2)
    x = F.silu(rms_norm(x, W["encoder.head.0.gamma"]))
    x = causal_conv3d(x, W["encoder.head.2.weight"], W["encoder.head.2.bias"])
    mu, _ = causal_conv3d(x, W["conv1.weight"], W["conv1.bias"]).chunk(2, dim=1)
    mean, inv_std = _scale(cfg, mu)
    return (mu - mean) * inv_std


# ------------------------------------------------------------------------------------------------- segment connect

def _connect(W, cfg, anchors: Tensor, num_frames: int, dec_kw: dict, enc_kw: dict, n_pix: Optional[int] = None) -> Tensor:
    """Wan_fps_inference_parallel_4gpu_20s.py:191-205, operation for operation (bf16 casts included)."""
    a = anchors.to(torch.bfloat16)
    masked = torch.zeros(a.shape[0], num_frames, *a.shape[2:], dtype=torch.bfloat16)
    masked[:, 0:1] = a[:, 0:1]
    masked[:, 1:2] = a[:, -2:-1]
    masked[:, 2:4] = a[:, -2:]
    vid = decode_to_pixel(W, cfg, masked, **dec_kw).to(torch.bfloat16)
    vid = (vid * 0.5 + 0.5).clamp(0, 1).to(torch.bfloat16)
    n_pix = 1 + 4 * (num_frames - 1) if n_pix is None else n_pix
    test = torch.zeros(vid.shape[0], n_pix, *vid.shape[2:], dtype=torch.bfloat16)
    test[:, 0:5] = vid[:, 8:13]
    test = (test * 2.0 - 1.0).permute(0, 2, 1, 3, 4)
    lat = encode_to_latent(W, cfg, test, **enc_kw)
    return lat[:, :2].to(torch.bfloat16)


def segment_connect(W, cfg, anchors: Tensor, num_frames: int = 21) -> Tensor:
    """The driver's transform as written: decode all `num_frames` latents (17 of them zero), re-encode all 81 frames.
    anchors [B,A,z,h,w] (t2v: A = 8) -> the next segment's `initial_latent` [B,2,z,h,w] bf16."""
    return _connect(W, cfg, anchors, num_frames, {}, {})


def segment_connect_causal(W, cfg, anchors: Tensor, num_frames: int = 21) -> Tensor:
    """The same bits from 4 decoder steps and 2 encoder chunks: pixel frames 8..12 are produced by latent frames 2 and 3
    (frames 5-8, 9-12) with causal history 0..1, and latents 0..1 by pixel frames 0..4."""
    return _connect(W, cfg, anchors, num_frames, {"max_frames": 4}, {"max_chunks": 2}, n_pix=5)
