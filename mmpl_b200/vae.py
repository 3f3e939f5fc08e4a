"""Host side of the VAE "segment connect" (SURVEY.md §8(f) row 2): the transform the reference driver applies to a
segment's anchor latents before the next segment may start (Wan_fps_inference_parallel_4gpu_20s.py:191-205) —
decode 21 latent frames of which 4 are non-zero, keep pixel frames 8..12, re-encode, keep latents 0..1 — through the
reference's wrapper interface (`WanVAEWrapper.decode_to_pixel` / `encode_to_latent`, utils/wan_wrapper.py:74-113).

B200 design:
* activations never leave one layout: a zero-haloed channels-last grid `[T, H + 2, W + 2, C]` bf16 (the halo = the spatial
  padding; the causal padding frames are not stored: the convolution addresses them with negative row coordinates, which
  TMA fills with zeros). Every convolution of the network — 3x3x3 causal, (3,1,1) temporal, per-frame 3x3, 1x1 — is one
  launch of the tap-GEMM tcgen05 kernel (`mmpl_conv3d_cl`) reading that grid in place; bias and the residual `x + h` ride
  in its epilogue, and every kernel writes the halo of its output as zeros, so no buffer is ever cleared;
* one pass per layer over ALL frames instead of the reference's per-frame / per-chunk passes with carried frames
  (`feat_cache`): the carried frames are exactly the causal left context. The two places where the chunked schedule
  is not a plain causal convolution are kept as the reference has them: the temporal up-sampler skips frame 0 and
  starts the remaining frames from zero history, the temporal down-sampler passes frame 0 through
  (vae.py:98-131,139-155; proven equal to the streaming form by tests/test_vae_oracle.py);
* causality also bounds the work: pixel frames 8..12 need latent frames 0..3 only and latents 0..1 need pixel frames
  0..4 only, so the connect decodes 4 frames and encodes 5 (bit-identical on the oracle) instead of 21 and 81;
* stride-2 convolutions run as stride-1 "same" convolutions followed by a pick of the odd positions / every second
  frame (three cheap 9-tap and two 3-tap layers); the middle attention (one head of 384 channels, outside the flash
  kernel's head_dim) is QK^T and PV on the GEMM kernel with a row-softmax kernel between them.

torch is used for device memory and for moving whole frames / channel slices between grids (copies, no arithmetic apart
from the 16-channel latent (de)normalisation and the final clamp of the wrapper). No CPU path: every op raises without
the sm_100a library.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

# utils/wan_wrapper.py:52-63
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


def _program(kind: str, dim: int, dim_mult, num_res_blocks: int, temporal) -> List[tuple]:
    """Flat layer list of Encoder3d.downsamples + middle / Decoder3d.middle + upsamples with the reference's state-dict
    prefixes (vae.py:284-310, 384-415)."""
    prog: List[tuple] = []
    if kind == "encoder":
        dims = [dim * u for u in [1] + list(dim_mult)]
        i = 0
        for lvl, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                prog.append(("res", f"encoder.downsamples.{i}"))
                i += 1
            if lvl != len(dim_mult) - 1:
                prog.append(("down3d" if temporal[lvl] else "down2d", f"encoder.downsamples.{i}"))
                i += 1
        prog += [("res", "encoder.middle.0"), ("attn", "encoder.middle.1"), ("res", "encoder.middle.2")]
    else:
        t_up = list(temporal)[::-1]
        prog += [("res", "decoder.middle.0"), ("attn", "decoder.middle.1"), ("res", "decoder.middle.2")]
        i = 0
        for lvl in range(len(dim_mult)):
            for _ in range(num_res_blocks + 1):
                prog.append(("res", f"decoder.upsamples.{i}"))
                i += 1
            if lvl != len(dim_mult) - 1:
                prog.append(("up3d" if t_up[lvl] else "up2d", f"decoder.upsamples.{i}"))
                i += 1
    return prog


class WanVAEWrapper(torch.nn.Module):
    """Drop-in for utils/wan_wrapper.py:WanVAEWrapper on the segment-connect path. Weights are bound with
    `load_state_dict`-compatible names (`WanVAE_.state_dict()`, i.e. the contents of Wan2.1_VAE.pth) through
    `load_vae_state_dict`; convolution weights are re-laid out once for the tap-GEMM."""

    def __init__(self, dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                 temperal_downsample=(False, True, True)):
        super().__init__()
        self.dim, self.z_dim = dim, z_dim
        self.mean = torch.tensor(LATENT_MEAN[:z_dim], dtype=torch.float32)
        self.std = torch.tensor(LATENT_STD[:z_dim], dtype=torch.float32)
        self._dim_mult = tuple(dim_mult)
        self._enc = _program("encoder", dim, dim_mult, num_res_blocks, temperal_downsample)
        self._dec = _program("decoder", dim, dim_mult, num_res_blocks, temperal_downsample)
        # output channels of every residual block (vae.py:284-300, 384-405)
        self._res_out: Dict[str, int] = {}
        i = 0
        for lvl, m in enumerate(dim_mult):
            for _ in range(num_res_blocks):
                self._res_out[f"encoder.downsamples.{i}"] = dim * m
                i += 1
            i += 1 if lvl != len(dim_mult) - 1 else 0
        i = 0
        for lvl, m in enumerate(list(dim_mult)[::-1]):
            for _ in range(num_res_blocks + 1):
                self._res_out[f"decoder.upsamples.{i}"] = dim * m
                i += 1
            i += 1 if lvl != len(dim_mult) - 1 else 0
        for side in ("encoder", "decoder"):
            for j in (0, 2):
                self._res_out[f"{side}.middle.{j}"] = dim * dim_mult[-1]
        self._w: Dict[str, torch.Tensor] = {}      # packed conv weights / raw vectors, by state-dict name
        self._kernel: Dict[str, tuple] = {}        # conv name -> (kt, kh, kw)

    # ------------------------------------------------------------------------------------------------ weights
    def load_vae_state_dict(self, sd: Dict[str, torch.Tensor], device="cuda") -> None:
        """Binds `WanVAE_.state_dict()`. 5-D / 4-D `.weight` tensors are convolutions (packed tap-major for
        mmpl_conv3d_cl); the 1x1 convolutions of the attention block stay [out, in] matrices for the GEMM."""
        self._w.clear()
        self._kernel.clear()
        for name, t in sd.items():
            t = t.detach().to(device)
            if name.endswith(".weight") and t.dim() >= 4:
                base = name[:-len(".weight")]
                if base.endswith(".to_qkv") or base.endswith(".proj"):
                    self._w[name] = t.reshape(t.shape[0], t.shape[1]).to(torch.bfloat16).contiguous()
                else:
                    k = tuple(t.shape[2:]) if t.dim() == 5 else (1,) + tuple(t.shape[2:])
                    self._kernel[base] = k
                    self._w[name] = ops.pack_conv_weight(t)
            else:
                self._w[name] = t.reshape(-1).to(torch.bfloat16).contiguous()

    def init_random_weights(self, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
        """Synthetic stand-in for the absent Wan2.1_VAE.pth: seeded fan-in-scaled normals under the
        reference's state-dict names and shapes, bound like a checkpoint. Returns the state dict."""
        g = torch.Generator().manual_seed(seed)
        sd: Dict[str, torch.Tensor] = {}

        def conv(name, cout, cin, *k):
            fan = cin
            for s_ in k:
                fan *= s_
            sd[name + ".weight"] = (torch.randn(cout, cin, *k, generator=g) * fan ** -0.5).to(torch.bfloat16)
            sd[name + ".bias"] = (torch.randn(cout, generator=g) * 0.05).to(torch.bfloat16)

        def gamma(name, c, nd):
            sd[name] = (1.0 + 0.1 * torch.randn(c, *([1] * nd), generator=g)).to(torch.bfloat16)

        def walk(prog, c):
            """Channel count through the layer list: a residual block changes it where the reference does (the encoder
            doubles at levels 1 and 2; the decoder halves after each up-sampler and doubles back at level 1)."""
            for kind, p in prog:
                if kind == "res":
                    cout = self._res_out[p]
                    gamma(p + ".residual.0.gamma", c, 3)
                    conv(p + ".residual.2", cout, c, 3, 3, 3)
                    gamma(p + ".residual.3.gamma", cout, 3)
                    conv(p + ".residual.6", cout, cout, 3, 3, 3)
                    if c != cout:
                        conv(p + ".shortcut", cout, c, 1, 1, 1)
                    c = cout
                elif kind == "attn":
                    gamma(p + ".norm.gamma", c, 2)
                    conv(p + ".to_qkv", 3 * c, c, 1, 1)
                    conv(p + ".proj", c, c, 1, 1)
                elif kind in ("down2d", "down3d"):
                    conv(p + ".resample.1", c, c, 3, 3)
                    if kind == "down3d":
                        conv(p + ".time_conv", c, c, 3, 1, 1)
                else:
                    if kind == "up3d":
                        conv(p + ".time_conv", 2 * c, c, 3, 1, 1)
                    conv(p + ".resample.1", c // 2, c, 3, 3)
                    c = c // 2
            return c

        top = self.dim * self._dim_mult[-1]
        conv("encoder.conv1", self.dim, 3, 3, 3, 3)
        walk(self._enc, self.dim)
        gamma("encoder.head.0.gamma", top, 3)
        conv("encoder.head.2", 2 * self.z_dim, top, 3, 3, 3)
        conv("conv1", 2 * self.z_dim, 2 * self.z_dim, 1, 1, 1)
        conv("conv2", self.z_dim, self.z_dim, 1, 1, 1)
        conv("decoder.conv1", top, self.z_dim, 3, 3, 3)
        last = walk(self._dec, top)
        gamma("decoder.head.0.gamma", last, 3)
        conv("decoder.head.2", 3, last, 3, 3, 3)
        self.load_vae_state_dict(sd, device=device)
        return sd

    # ----------------------------------------------------------------------------------------------- primitives
    def _conv(self, name: str, grid: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.conv3d_causal_cl(grid, self._w[name + ".weight"], self._w.get(name + ".bias"), self._kernel[name],
                                    history=0, residual=residual)

    def _res_block(self, p: str, x: torch.Tensor) -> torch.Tensor:
        """ResidualBlock.forward (vae.py:189-213): shortcut(x) + conv(silu(norm(conv(silu(norm(x))))))."""
        h = self._conv(p + ".shortcut", x) if (p + ".shortcut") in self._kernel else x
        y = ops.vae_norm_act(x, self._w[p + ".residual.0.gamma"], silu=True)
        y = self._conv(p + ".residual.2", y)
        y = ops.vae_norm_act(y, self._w[p + ".residual.3.gamma"], silu=True)
        return self._conv(p + ".residual.6", y, residual=h)

    def _attention(self, p: str, x: torch.Tensor) -> torch.Tensor:
        """AttentionBlock.forward (vae.py:241-266), frame by frame: one head over the h*w positions, C channels."""
        frames, hp, wp, c = x.shape
        n = (hp - 2) * (wp - 2)
        xn = ops.vae_norm_act(x, self._w[p + ".norm.gamma"], silu=False)
        out = torch.zeros_like(x)
        for f in range(frames):
            rows = xn[f, 1:-1, 1:-1].reshape(n, c)                      # compact copy of the interior
            ident = x[f, 1:-1, 1:-1].reshape(n, c)
            qkv = ops.linear(rows, self._w[p + ".to_qkv.weight"], self._w[p + ".to_qkv.bias"])
            q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
            s = ops.linear(q, k)                                        # [n, n] raw scores, K = C
            prob = ops.softmax_rows(s, c ** -0.5)
            o = ops.linear(prob, v.t().contiguous())                    # [n, C], K = n
            y = ops.linear(o, self._w[p + ".proj.weight"], self._w[p + ".proj.bias"], epilogue=ops.EPI_BIAS_RES,
                           residual=ident)
            out[f, 1:-1, 1:-1] = y.reshape(hp - 2, wp - 2, c)
        return out

    def _up(self, p: str, x: torch.Tensor, temporal: bool) -> torch.Tensor:
        """Resample 'upsample2d' / 'upsample3d' (vae.py:98-136) over the whole frame axis."""
        t, hp, wp, c = x.shape
        if temporal and t > 1:
            # frames 1.. through the (3,1,1) convolution with zero history in front of frame 1 (a view: the zeros are
            # implicit), two frames out per frame in, interleaved behind frame 0
            y = self._conv(p + ".time_conv", x[1:])                     # [t - 1, .., 2c]
            wide = torch.empty((1 + 2 * (t - 1), hp, wp, c), dtype=x.dtype, device=x.device)
            wide[0] = x[0]
            wide[1::2] = y[..., :c]
            wide[2::2] = y[..., c:]
            x = wide
        x = ops.vae_upsample2x(x)
        return self._conv(p + ".resample.1", x)

    def _down(self, p: str, x: torch.Tensor, temporal: bool) -> torch.Tensor:
        """Resample 'downsample2d' / 'downsample3d' (vae.py:133-155) over the whole frame axis."""
        x = ops.vae_pick_odd(self._conv(p + ".resample.1", x))
        t = x.shape[0]
        if temporal and t > 1:
            # stride-2 windows (0,1,2), (2,3,4), ... of the un-padded sequence = the causal convolution at frames 2, 4, ...;
            # frame 0 passes through
            y = self._conv(p + ".time_conv", x)
            x = torch.cat([x[:1], y[2::2]], dim=0)
        return x

    def _run(self, prog, x):
        for kind, p in prog:
            if kind == "res":
                x = self._res_block(p, x)
            elif kind == "attn":
                x = self._attention(p, x)
            elif kind in ("up2d", "up3d"):
                x = self._up(p, x, kind == "up3d")
            else:
                x = self._down(p, x, kind == "down3d")
        return x

    # ------------------------------------------------------------------------------------------- encode / decode
    @torch.no_grad()
    def _decode_one(self, z: torch.Tensor) -> torch.Tensor:
        """WanVAE_.decode (vae.py:530-552). z [z_dim, T, h, w] bf16 -> [3, 1 + 4(T-1), 8h, 8w] bf16."""
        mean = self.mean.to(device=z.device, dtype=z.dtype).view(-1, 1, 1, 1)
        inv_std = (1.0 / self.std.to(device=z.device, dtype=z.dtype)).view(-1, 1, 1, 1)
        x = ops.to_haloed(z / inv_std + mean)
        x = self._conv("conv2", x)
        x = self._conv("decoder.conv1", x)
        x = self._run(self._dec, x)
        x = ops.vae_norm_act(x, self._w["decoder.head.0.gamma"], silu=True)
        x = self._conv("decoder.head.2", x)
        return ops.from_haloed(x, 3)

    @torch.no_grad()
    def _encode_one(self, pixels: torch.Tensor) -> torch.Tensor:
        """WanVAE_.encode (vae.py:501-528). pixels [3, 1 + 4k, H, W] bf16 -> mu [z_dim, 1 + k, H/8, W/8] bf16."""
        x = ops.to_haloed(pixels)
        x = self._conv("encoder.conv1", x)
        x = self._run(self._enc, x)
        x = ops.vae_norm_act(x, self._w["encoder.head.0.gamma"], silu=True)
        x = self._conv("encoder.head.2", x)
        x = self._conv("conv1", x)
        mu = ops.from_haloed(x, self.z_dim)                       # .chunk(2, dim=1)[0]
        mean = self.mean.to(device=mu.device, dtype=mu.dtype).view(-1, 1, 1, 1)
        inv_std = (1.0 / self.std.to(device=mu.device, dtype=mu.dtype)).view(-1, 1, 1, 1)
        return (mu - mean) * inv_std

    def encode_to_latent(self, pixel: torch.Tensor) -> torch.Tensor:
        """utils/wan_wrapper.py:74-89: [B, 3, T, H, W] -> fp32 [B, T', z_dim, h, w]."""
        out = torch.stack([self._encode_one(u.to(torch.bfloat16)).float() for u in pixel])
        return out.permute(0, 2, 1, 3, 4)

    def decode_to_pixel(self, latent: torch.Tensor, use_cache: bool = False) -> torch.Tensor:
        """utils/wan_wrapper.py:91-113: [B, T, z_dim, h, w] -> fp32 [B, T', 3, H, W] clamped to [-1, 1]."""
        if use_cache:
            raise NotImplementedError("cached_decode (feature cache kept across calls, vae.py:554-577) is not on the segment-"
                                      "connect path: every call here decodes its whole frame axis in one pass")
        zs = latent.permute(0, 2, 1, 3, 4)
        out = torch.stack([self._decode_one(u.to(torch.bfloat16)).float().clamp_(-1, 1) for u in zs])
        return out.permute(0, 2, 1, 3, 4)

    # ------------------------------------------------------------------------------------------ segment connect
    @torch.no_grad()
    def segment_connect(self, anchors: torch.Tensor) -> torch.Tensor:
        """Wan_fps_inference_parallel_4gpu_20s.py:191-205 on the causal support of its result. anchors [B, A, z, h, w]
        (t2v: frame 0 + the 7 stage-1 frames) -> the next segment's `initial_latent` [B, 2, z, h, w] bf16."""
        a = anchors.to(torch.bfloat16)
        masked = torch.cat([a[:, 0:1], a[:, -2:-1], a[:, -2:]], dim=1)                   # latent frames 0..3 of the 21
        vid = self.decode_to_pixel(masked).to(torch.bfloat16)                           # 13 pixel frames
        vid = (vid * 0.5 + 0.5).clamp(0, 1).to(torch.bfloat16)
        test = (vid[:, 8:13] * 2.0 - 1.0).permute(0, 2, 1, 3, 4)                         # 5 frames -> latents 0..1
        return self.encode_to_latent(test)[:, :2].to(torch.bfloat16)
