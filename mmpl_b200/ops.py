"""Tensor-level wrappers over the per-kernel C-ABI entry points (include/mmpl_b200.h).

torch is used only to own device memory and to name the current CUDA stream; every function launches
one hand-written sm_100a kernel through libmmpl_b200.so and raises if that fails. Inputs must be CUDA
bf16 tensors whose last dimension is contiguous.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional, Sequence

import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_SILU, EPI_BIAS_RES, EPI_BIAS_GATE_RES = range(5)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Runs the wrapped op with the CUDA device of its first tensor argument current, so that the library's launch, the
    stream handed to it and its per-device caches all belong to the device that owns the pointers (a process may hold
    models on several GPUs, as the reference's thread-per-GPU drivers do)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: mmpl_b200 has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() > 0 and t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")


def _ints(v: Sequence[int]):
    return (C.c_int * len(v))(*[int(i) for i in v])


@_on_device
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = EPI_BIAS,
           residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_frame: int = 0,
           out: Optional[torch.Tensor] = None, tile_n: int = 0) -> torch.Tensor:
    """epilogue(x[M,K] @ weight[N,K]^T + bias). gate: [frames, N] (row pitch = gate.stride(0))."""
    lib = _lib.load()
    _req(x, "x"); _req(weight, "weight")
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.mmpl_gemm_bf16(
        x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0), _p(bias), out.data_ptr(), out.stride(0),
        M, N, K, epilogue, _p(residual), 0 if residual is None else residual.stride(0),
        _p(gate), 0 if gate is None else gate.stride(0), rows_per_frame, tile_n, _stream()))
    return out


@_on_device
def flash_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, segments: Optional[Sequence[tuple]] = None,
               k_tail: Optional[torch.Tensor] = None, v_tail: Optional[torch.Tensor] = None,
               softmax_scale: Optional[float] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q: [Lq, H, 128]; k, v: [rows, H, 128]. segments: list of (start, rows[, src]) over k/v rows
    (src 1 = k_tail/v_tail); default = all rows of k."""
    lib = _lib.load()
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, n)
    Lq, H, hd = q.shape
    if hd != 128 or k.shape[1:] != (H, 128) or v.shape != k.shape:
        raise ValueError("flash_attn: head_dim must be 128 and k/v must be [rows, H, 128]")
    if q.stride(1) != 128 or k.stride(1) != 128 or v.stride(1) != 128 or k.stride(0) != v.stride(0):
        raise ValueError("flash_attn: heads must be packed (stride 128) and k/v share a row pitch")
    if segments is None:
        segments = [(0, k.shape[0], 0)]
    segs = [(s[0], s[1], s[2] if len(s) > 2 else 0) for s in segments]
    if out is None:
        out = torch.empty((Lq, H, 128), dtype=torch.bfloat16, device=q.device)
    scale = float(softmax_scale) if softmax_scale is not None else 128 ** -0.5
    has_tail = k_tail is not None
    _lib.check(lib.mmpl_flash_attn(
        q.data_ptr(), q.stride(0), Lq, H, k.data_ptr(), v.data_ptr(), k.stride(0), k.shape[0],
        _p(k_tail), _p(v_tail), k_tail.stride(0) if has_tail else 0, k_tail.shape[0] if has_tail else 0,
        len(segs), _ints([s[0] for s in segs]), _ints([s[1] for s in segs]), _ints([s[2] for s in segs]),
        out.data_ptr(), out.stride(0), scale, _stream()))
    return out


@_on_device
def ln_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, rows_per_frame: int, eps: float = 1e-6,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [S, D]; shift/scale: [frames, D] views sharing a row pitch."""
    lib = _lib.load()
    _req(x, "x"); _req(shift, "shift"); _req(scale, "scale")
    S, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    assert shift.stride(0) == scale.stride(0) or shift.shape[0] == 1
    _lib.check(lib.mmpl_ln_modulate(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), S, D, eps,
                                    shift.data_ptr(), scale.data_ptr(), shift.stride(0), rows_per_frame, _stream()))
    return out


@_on_device
def ln_affine(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    lib = _lib.load()
    _req(x, "x"); _req(weight, "weight"); _req(bias, "bias")
    out = torch.empty_like(x)
    _lib.check(lib.mmpl_ln_affine(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1], eps,
                                  weight.data_ptr(), bias.data_ptr(), _stream()))
    return out


@_on_device
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, inplace: bool = False) -> torch.Tensor:
    lib = _lib.load()
    _req(x, "x"); _req(weight, "weight")
    out = x if inplace else torch.empty_like(x)
    _lib.check(lib.mmpl_rmsnorm(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1],
                                weight.data_ptr(), eps, _stream()))
    return out


@_on_device
def qk_norm_rope_kv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, norm_q_w: torch.Tensor, norm_k_w: torch.Tensor,
                    rope_table: torch.Tensor, k_dst: torch.Tensor, v_dst: torch.Tensor, grid_hw: tuple,
                    frame_pos: Sequence[int], kv_row: Sequence[int], eps: float = 1e-6,
                    q_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q,k,v: [S, D] views with one row pitch; k_dst/v_dst: [rows, D]; rope_table: float64 [1024,64,2] on device."""
    lib = _lib.load()
    for n, t in (("q", q), ("k", k), ("v", v), ("k_dst", k_dst), ("v_dst", v_dst)):
        _req(t, n)
    _req(rope_table, "rope_table", torch.float64)
    S, D = q.shape
    assert q.stride(0) == k.stride(0) == v.stride(0) and k_dst.stride(0) == v_dst.stride(0)
    if q_out is None:
        q_out = torch.empty((S, D), dtype=torch.bfloat16, device=q.device)
    _lib.check(lib.mmpl_qk_norm_rope_kv(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), norm_q_w.data_ptr(), norm_k_w.data_ptr(),
        rope_table.data_ptr(), q_out.data_ptr(), q_out.stride(0), k_dst.data_ptr(), v_dst.data_ptr(), k_dst.stride(0),
        S, D, grid_hw[0], grid_hw[1], len(frame_pos), _ints(frame_pos), _ints(kv_row), eps, _stream()))
    return q_out


@_on_device
def modulation_add(mod: torch.Tensor, src: torch.Tensor, src_fstride: int, src_jstride: int, F: int) -> torch.Tensor:
    """out[f, j, :] = bf16(mod[j, :] + src[f*src_fstride + j*src_jstride + :]); mod: [J, D]."""
    lib = _lib.load()
    _req(mod, "mod"); _req(src, "src")
    J, D = mod.shape
    out = torch.empty((F, J, D), dtype=torch.bfloat16, device=mod.device)
    _lib.check(lib.mmpl_modulation_add(mod.data_ptr(), src.data_ptr(), src_fstride, src_jstride, out.data_ptr(),
                                       F, J, D, _stream()))
    return out


@_on_device
def sinusoid_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    lib = _lib.load()
    _req(t, "t", torch.float64)
    out = torch.empty((t.numel(), dim), dtype=torch.bfloat16, device=t.device)
    _lib.check(lib.mmpl_sinusoid_embedding(t.data_ptr(), out.data_ptr(), t.numel(), dim, _stream()))
    return out


@_on_device
def skinny_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], silu_in: bool = False,
                  silu_out: bool = False) -> torch.Tensor:
    lib = _lib.load()
    _req(x, "x"); _req(weight, "weight")
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.mmpl_skinny_linear(x.data_ptr(), x.stride(0), weight.data_ptr(), _p(bias), out.data_ptr(),
                                      out.stride(0), M, N, K, int(silu_in), int(silu_out), _stream()))
    return out


@_on_device
def patchify(x: torch.Tensor) -> torch.Tensor:
    """x: [F, C, H, W] (any frame/channel strides, contiguous H*W planes) -> [F*(H/2)*(W/2), C*4]."""
    lib = _lib.load()
    _req(x, "x")
    F, Cc, H, W = x.shape
    assert x.stride(3) == 1 and x.stride(2) == W
    out = torch.empty((F * (H // 2) * (W // 2), Cc * 4), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.mmpl_patchify(x.data_ptr(), x.stride(0), x.stride(1), out.data_ptr(), F, Cc, H, W, _stream()))
    return out


@_on_device
def unpatchify_x0(head: torch.Tensor, shape: tuple, xt: Optional[torch.Tensor] = None,
                  sigma: Optional[torch.Tensor] = None):
    """head: [F*(H/2)*(W/2), 4*C] -> flow [F, C, H, W] (and x0 when xt [F,C,H,W] and sigma float64 [F] are given)."""
    lib = _lib.load()
    _req(head, "head")
    F, Cc, H, W = shape
    flow = torch.empty(shape, dtype=torch.bfloat16, device=head.device)
    x0 = torch.empty_like(flow) if xt is not None else None
    if xt is not None:
        _req(xt, "xt"); _req(sigma, "sigma", torch.float64)
    _lib.check(lib.mmpl_unpatchify_x0(head.data_ptr(), head.stride(0), _p(xt), 0 if xt is None else xt.stride(0),
                                      0 if xt is None else xt.stride(1), _p(sigma), flow.data_ptr(), _p(x0),
                                      F, Cc, H, W, _stream()))
    return flow, x0


@_on_device
def add_noise(x0: torch.Tensor, noise: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """x0, noise: [N, ...] contiguous bf16; sigma: float32 [N]."""
    lib = _lib.load()
    _req(x0, "x0"); _req(noise, "noise"); _req(sigma, "sigma", torch.float32)
    assert x0.is_contiguous() and noise.is_contiguous()
    out = torch.empty_like(noise)
    n = x0.shape[0]
    _lib.check(lib.mmpl_add_noise(x0.data_ptr(), noise.data_ptr(), sigma.data_ptr(), out.data_ptr(), n,
                                  x0.numel() // n, _stream()))
    return out


# ------------------------------------------------------------------------------------ VAE segment connect

def pack_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """Reference conv parameter [Cout, Cin, KT, KH, KW] (or [Cout, Cin, KH, KW] of a per-frame Conv2d) -> the layout
    mmpl_conv3d_cl reads, bf16, Cout padded to a multiple of 8 and Cin to a multiple of 8:
      KW = 1: [Cout8, KT*KH, Cin64]              one K span of Cin (zero-padded to 64) per tap
      KW = 3: [Cout8, KT*KH, pad64(3 * Cin8)]    one K span per (dt, dh): the three dw taps side by side, (dw, c) order,
                                                 because they are one contiguous run of the channels-last grid
    One-time layout transform of the weights (no arithmetic)."""
    if weight.dim() == 4:
        weight = weight.unsqueeze(2)
    cout, cin, kt, kh, kw = weight.shape
    cin8, cout8 = -(-cin // 8) * 8, -(-cout // 8) * 8
    w = weight.permute(0, 2, 3, 4, 1).to(torch.bfloat16)                      # [cout, kt, kh, kw, cin]
    if kw == 3:
        span = -(-3 * cin8 // 64) * 64
        packed = torch.zeros((cout8, kt * kh, span), dtype=torch.bfloat16, device=weight.device)
        rows = torch.zeros((cout, kt * kh, 3, cin8), dtype=torch.bfloat16, device=weight.device)
        rows[..., :cin] = w.reshape(cout, kt * kh, 3, cin)
        packed[:cout, :, :3 * cin8] = rows.reshape(cout, kt * kh, 3 * cin8)
        return packed
    cin64 = -(-cin // 64) * 64
    packed = torch.zeros((cout8, kt * kh * kw, cin64), dtype=torch.bfloat16, device=weight.device)
    packed[:cout, :, :cin] = w.reshape(cout, kt * kh * kw, cin)
    return packed


def to_haloed(x: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    """[C, T, H, W] -> channels-last grid [T, H + 2, W + 2, C8] with a zero halo (the layout the VAE kernels keep their
    activations in; C padded to a multiple of 8 or to `channels`). The causal padding frames are not stored: the
    convolution reads them as zeros (see mmpl_conv3d_cl `history`)."""
    c, t, h, w = x.shape
    c8 = channels if channels is not None else -(-c // 8) * 8
    g = torch.zeros((t, h + 2, w + 2, c8), dtype=torch.bfloat16, device=x.device)
    g[:, 1:-1, 1:-1, :c] = x.permute(1, 2, 3, 0).to(torch.bfloat16)
    return g


def from_haloed(g: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    """Inverse of to_haloed: [T, H + 2, W + 2, C8] -> [C, T, H, W]."""
    c = g.shape[3] if channels is None else channels
    return g[:, 1:-1, 1:-1, :c].permute(3, 0, 1, 2).contiguous()


@_on_device
def conv3d_causal_cl(grid: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], kernel: Sequence[int],
                     history: int = 0, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CausalConv3d.forward (wan/modules/vae.py:16-36), stride 1, on a haloed channels-last grid (see to_haloed).
    `grid` is [history + T, H + 2, W + 2, Cin]: `history` (0..KT-1) carried frames in front of the T frames to convolve,
    zeros standing in for the rest of the causal padding. Returns [T, H + 2, W + 2, Cout] (halo written as zeros by the
    kernel). `residual`: a grid like the result, added in the epilogue."""
    lib = _lib.load()
    _req(grid, "grid"); _req(w_packed, "w_packed")
    kt, kh, kw = (int(k) for k in kernel)
    frames, hp, wp, cin = grid.shape
    t = frames - history
    cout = w_packed.shape[0]
    if not 0 <= history <= kt - 1 or t <= 0:
        raise ValueError(f"history={history} frames in a grid of {frames}; the kernel takes 0..{kt - 1}")
    span = -(-3 * cin // 64) * 64 if kw == 3 else -(-cin // 64) * 64
    if tuple(w_packed.shape[1:]) != (kt * kh * (1 if kw == 3 else kw), span):
        raise ValueError("w_packed does not match the grid's channels / the kernel size (see pack_conv_weight)")
    if not grid.is_contiguous():
        raise ValueError("grid must be contiguous")
    if out is None:
        out = torch.empty((t, hp, wp, cout), dtype=torch.bfloat16, device=grid.device)
    if bias is not None and bias.numel() != cout:
        bias = torch.cat([bias.to(torch.bfloat16), bias.new_zeros(cout - bias.numel(), dtype=torch.bfloat16)])
    if bias is not None:
        bias = bias.to(torch.bfloat16).contiguous()
    _lib.check(lib.mmpl_conv3d_cl(grid.data_ptr(), w_packed.data_ptr(), _p(bias), out.data_ptr(), _p(residual),
                                  t, hp - 2, wp - 2, cin, cout, kt, kh, kw, history, _stream()))
    return out


@_on_device
def vae_norm_act(grid: torch.Tensor, gamma: torch.Tensor, silu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """RMS_norm (+ SiLU) over the channels of every grid position (wan/modules/vae.py:39-55). gamma: any shape with C elements."""
    lib = _lib.load()
    _req(grid, "grid")
    if not grid.is_contiguous():
        raise ValueError("grid must be contiguous")
    c = grid.shape[-1]
    g = gamma.reshape(-1).to(torch.bfloat16)
    if g.numel() != c:  # channel-padded layouts: padded channels are zero and stay zero
        g = torch.cat([g, g.new_ones(c - g.numel())])
    if out is None:
        out = torch.empty_like(grid)
    _lib.check(lib.mmpl_vae_norm_act(grid.data_ptr(), out.data_ptr(), grid.numel() // c, c, g.data_ptr(), int(silu), _stream()))
    return out


@_on_device
def vae_upsample2x(grid: torch.Tensor) -> torch.Tensor:
    """Nearest-neighbour 2x up-sampling of every frame of a haloed grid (vae.py:58-64)."""
    lib = _lib.load()
    _req(grid, "grid")
    frames, hp, wp, c = grid.shape
    out = torch.empty((frames, 2 * (hp - 2) + 2, 2 * (wp - 2) + 2, c), dtype=torch.bfloat16, device=grid.device)
    _lib.check(lib.mmpl_vae_upsample2x(grid.data_ptr(), out.data_ptr(), frames, hp - 2, wp - 2, c, _stream()))
    return out


@_on_device
def vae_pick_odd(grid: torch.Tensor) -> torch.Tensor:
    """out interior (i, j) = in interior (2i+1, 2j+1): stride-1 "same" conv -> ZeroPad2d((0,1,0,1)) + stride-2 conv (vae.py:85-88)."""
    lib = _lib.load()
    _req(grid, "grid")
    frames, hp, wp, c = grid.shape
    out = torch.empty((frames, (hp - 2) // 2 + 2, (wp - 2) // 2 + 2, c), dtype=torch.bfloat16, device=grid.device)
    _lib.check(lib.mmpl_vae_pick_odd(grid.data_ptr(), out.data_ptr(), frames, hp - 2, wp - 2, c, _stream()))
    return out


@_on_device
def softmax_rows(s: torch.Tensor, scale: float) -> torch.Tensor:
    """softmax(scale * s) over the last dimension of a 2-D bf16 tensor, fp32 inside."""
    lib = _lib.load()
    _req(s, "s")
    p = torch.empty_like(s)
    _lib.check(lib.mmpl_softmax_rows(s.data_ptr(), s.stride(0), p.data_ptr(), p.stride(0), s.shape[0], s.shape[1],
                                     float(scale), _stream()))
    return p
