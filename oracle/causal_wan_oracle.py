"""ORACLE — test infrastructure only. A CPU restatement of the reference's chunk-wise causal denoising
path (Tele-AI/MMPL, MMPL_t2v/), written from the maths of the files cited below, in plain PyTorch ops on
whatever device the inputs live on (CPU in the test-suite). It is the checker for the CUDA path, never
the thing shipped or measured: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import it.

Pinning: the reference ships no tests or golden vectors for this path, so this
restatement is pinned against outputs of the reference code itself, generated in the build container by
oracle/make_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py checks them, bit-exact
on CPU). Model activations are bf16 and every op rounds to bf16 where the reference does (SURVEY.md
appendix A); `dtype=torch.float32` gives the un-rounded fp32 restatement used to calibrate tolerances.

Functions cite `file:line` under /root/reference/MMPL_t2v/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# configuration and synthetic weights
# --------------------------------------------------------------------------------------------------
@dataclass
class WanConfig:
    """Constructor arguments of CausalWanModel (wan/modules/causal_model.py:371-387)."""
    dim: int = 1536
    ffn_dim: int = 8960
    num_heads: int = 12
    num_layers: int = 30
    freq_dim: int = 256
    text_dim: int = 4096
    text_len: int = 512
    in_dim: int = 16
    out_dim: int = 16
    eps: float = 1e-6
    patch_size: tuple = (1, 2, 2)

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


WAN_1_3B = WanConfig()  # wan/configs/wan_t2v_1_3B.py:15-24
WAN_14B = WanConfig(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)  # wan_t2v_14B.py:15-24


def param_shapes(cfg: WanConfig) -> Dict[str, tuple]:
    """State-dict names and shapes of CausalWanModel (causal_model.py:445-476, 86-118, 234-271, 329-344)."""
    D, Fd = cfg.dim, cfg.ffn_dim
    s: Dict[str, tuple] = {
        "patch_embedding.weight": (D, cfg.in_dim, *cfg.patch_size), "patch_embedding.bias": (D,),
        "text_embedding.0.weight": (D, cfg.text_dim), "text_embedding.0.bias": (D,),
        "text_embedding.2.weight": (D, D), "text_embedding.2.bias": (D,),
        "time_embedding.0.weight": (D, cfg.freq_dim), "time_embedding.0.bias": (D,),
        "time_embedding.2.weight": (D, D), "time_embedding.2.bias": (D,),
        "time_projection.1.weight": (6 * D, D), "time_projection.1.bias": (6 * D,),
        "head.head.weight": (cfg.out_dim * math.prod(cfg.patch_size), D),
        "head.head.bias": (cfg.out_dim * math.prod(cfg.patch_size),),
        "head.modulation": (1, 2, D),
    }
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        for a in ("self_attn", "cross_attn"):
            for p in ("q", "k", "v", "o"):
                s[f"{b}{a}.{p}.weight"] = (D, D)
                s[f"{b}{a}.{p}.bias"] = (D,)
            s[f"{b}{a}.norm_q.weight"] = (D,)
            s[f"{b}{a}.norm_k.weight"] = (D,)
        s[f"{b}norm3.weight"] = (D,)
        s[f"{b}norm3.bias"] = (D,)
        s[f"{b}ffn.0.weight"] = (Fd, D)
        s[f"{b}ffn.0.bias"] = (Fd,)
        s[f"{b}ffn.2.weight"] = (D, Fd)
        s[f"{b}ffn.2.bias"] = (D,)
        s[f"{b}modulation"] = (1, 6, D)
    return s


def make_weights(cfg: WanConfig, seed: int = 0, dtype=torch.bfloat16, device="cpu") -> Dict[str, Tensor]:
    """Seeded synthetic weights with the distributions of CausalWanModel.init_weights (causal_model.py:1119-1141):
    xavier-uniform Linear weights, normal(0, .02) text/time embeddings, modulation ~ randn/sqrt(D). Unlike the
    reference init, biases, norm weights and the head are randomised too (the reference zeroes
    head.head.weight, which makes every output exactly 0 and parity vacuous)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("modulation"):
            w = torch.randn(shape, generator=g) / cfg.dim ** 0.5
        elif "norm" in name and name.endswith("weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            w = 0.02 * torch.randn(shape, generator=g)
        elif name.startswith(("text_embedding", "time_embedding")) or name == "head.head.weight":
            w = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_out, fan_in = shape[0], math.prod(shape[1:])
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            w = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[name] = w.to(dtype).to(device)
    return out


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: Tensor) -> Tensor:
    """wan/modules/model.py:15-25 — float64 outer product, cat(cos, sin)."""
    half = dim // 2
    pos = position.to(torch.float64)
    inv = torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=pos.device) / half)
    ang = pos[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1)


def rope_freqs(head_dim: int = 128, max_len: int = 1024, theta: float = 10000.0) -> Tensor:
    """complex128 table [max_len, head_dim/2]: 22 temporal + 21 height + 21 width pairs for head_dim 128
    (model.py:29-36 rope_params; concatenation at causal_model.py:473-478)."""
    def params(dim):
        f = torch.outer(torch.arange(max_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64) / dim))
        return torch.polar(torch.ones_like(f), f)
    d = head_dim
    return torch.cat([params(d - 4 * (d // 6)), params(2 * (d // 6)), params(2 * (d // 6))], dim=1)


def rope_table_real(freqs: Tensor) -> Tensor:
    """[max_len, 64, 2] float64 (cos, sin) view of the complex table — the layout the CUDA kernel reads."""
    return torch.view_as_real(freqs).contiguous()


def rope_apply(x: Tensor, grid: Sequence[int], freqs: Tensor, frame_pos: Sequence[int]) -> Tensor:
    """causal_rope_apply (causal_model.py:27-55) / causal_fps_rope_apply (causal_fps_model.py:27-55) for one
    sample: x [S, H, hd]; token (f,h,w) is rotated by freqs[frame_pos[f]] (temporal), freqs[h], freqs[w].
    Complex multiply in float64, result cast back to x.dtype."""
    f, h, w = grid
    S, n, hd = x.shape
    c = hd // 2
    split = [c - 2 * (c // 3), c // 3, c // 3]
    ft, fh, fw = freqs.to(x.device).split(split, dim=1)
    pos_t = torch.as_tensor(list(frame_pos), dtype=torch.long, device=x.device)
    fr = torch.cat([
        ft[pos_t].view(f, 1, 1, -1).expand(f, h, w, -1),
        fh[:h].view(1, h, 1, -1).expand(f, h, w, -1),
        fw[:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(S, 1, c)
    xc = torch.view_as_complex(x.to(torch.float64).reshape(S, n, c, 2))
    return torch.view_as_real(xc * fr).flatten(2).to(x.dtype)


def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """WanRMSNorm (model.py:70-86): fp32 normalisation over the full last dim, cast, then * weight."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return y.to(x.dtype) * weight


def layer_norm(x: Tensor, eps: float, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None) -> Tensor:
    """WanLayerNorm (model.py:89-99): fp32 statistics, one rounding to x.dtype."""
    xf = x.float()
    mu = xf.mean(dim=-1, keepdim=True)
    var = (xf - mu).pow(2).mean(dim=-1, keepdim=True)
    y = (xf - mu) * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * weight.float() + bias.float()
    return y.to(x.dtype)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Linear with fp32 accumulation and a single rounding of (acc + bias) to the activation dtype."""
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    return y.to(x.dtype)


def gelu_tanh(x: Tensor) -> Tensor:
    return F.gelu(x.float(), approximate="tanh").to(x.dtype)


def silu(x: Tensor) -> Tensor:
    return F.silu(x.float()).to(x.dtype)


def attention(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """softmax(q k^T / sqrt(hd)) v without mask (attention.py:119-136 / :177-185): q [Lq,H,hd], k/v [Lk,H,hd];
    fp32 scores, softmax and accumulation, output rounded to q.dtype. Chunked over queries to bound memory."""
    Lq, H, hd = q.shape
    scale = 1.0 / math.sqrt(hd)
    kf = k.float().permute(1, 2, 0)  # H, hd, Lk
    vf = v.float().permute(1, 0, 2)  # H, Lk, hd
    out = torch.empty((H, Lq, hd), dtype=torch.float32, device=q.device)
    step = max(1, min(Lq, (1 << 26) // max(1, k.shape[0] * H)))
    for s in range(0, Lq, step):
        qs = q[s:s + step].float().permute(1, 0, 2)  # H, l, hd
        p = torch.softmax(torch.bmm(qs, kf) * scale, dim=-1)
        out[:, s:s + step] = torch.bmm(p, vf)
    return out.permute(1, 0, 2).to(q.dtype)


def modulate(x_norm: Tensor, shift: Tensor, scale: Tensor, frame_seqlen: int) -> Tensor:
    """norm(x).unflatten(F, fs) * (1 + scale) + shift, every op rounding in x.dtype (causal_model.py:305,318)."""
    S, D = x_norm.shape
    nf = S // frame_seqlen
    y = x_norm.view(nf, frame_seqlen, D) * (1 + scale.view(nf, 1, D)) + shift.view(nf, 1, D)
    return y.view(S, D)


def gate_residual(x: Tensor, y: Tensor, gate: Tensor, frame_seqlen: int) -> Tensor:
    """x + (y.unflatten(F, fs) * gate).flatten (causal_model.py:310,322)."""
    S, D = x.shape
    nf = S // frame_seqlen
    return x + (y.view(nf, frame_seqlen, D) * gate.view(nf, 1, D)).view(S, D)


# --------------------------------------------------------------------------------------------------
# KV-cache bookkeeping (integer, bit-exact)
# --------------------------------------------------------------------------------------------------
def contiguous_cache_plan(local_end_prev: int, global_end_prev: int, current_start: int, num_new: int,
                          max_attention_size: int = 32760):
    """causal_model.py:203-226: rows written [local_start, local_end), rows attended [win_start, local_end),
    and the new (global_end, local_end) values."""
    current_end = current_start + num_new
    local_end = local_end_prev + current_end - global_end_prev
    local_start = local_end - num_new
    win_start = max(0, local_end - max_attention_size)
    return dict(local_start=local_start, local_end=local_end, win_start=win_start,
                global_end=current_end, local_end_index=local_end)


def fps_slot_of_frame(frame: int) -> int:
    """Frame -> cache slot of CausalFPSWanModel (causal_fps_model.py:213-247): frames >= 19 are stored 6 slots
    lower (19 -> 13, 20 -> 14)."""
    return frame - 6 if frame >= 19 else frame


# --------------------------------------------------------------------------------------------------
# model forward
# --------------------------------------------------------------------------------------------------
@dataclass
class KVCache:
    """Per-layer cache dict of CausalInferencePipeline._initialize_kv_cache (pipeline/causal_inference.py:278-297)
    for batch size 1: k, v [rows, H, hd]; indices as python ints."""
    k: Tensor
    v: Tensor
    global_end_index: int = 0
    local_end_index: int = 0
    visible: List[int] = field(default_factory=list)  # FPS model: token offsets of visible frames


@dataclass
class CrossCache:
    """pipeline/causal_inference.py:299-312."""
    k: Optional[Tensor] = None
    v: Optional[Tensor] = None
    is_init: bool = False


def new_caches(cfg: WanConfig, rows: int, dtype=torch.bfloat16, device="cpu"):
    kv = [KVCache(torch.zeros(rows, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
                  torch.zeros(rows, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device))
          for _ in range(cfg.num_layers)]
    cross = [CrossCache() for _ in range(cfg.num_layers)]
    return kv, cross


def self_attention(cfg: WanConfig, w: Dict[str, Tensor], pre: str, x: Tensor, grid, freqs: Tensor, kv: KVCache,
                   current_start: int, trace: Optional[list] = None) -> Tensor:
    """CausalWanSelfAttention.forward, KV-cache branch (causal_model.py:86-231)."""
    S, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    q = rms_norm(linear(x, w[pre + "q.weight"], w[pre + "q.bias"]), w[pre + "norm_q.weight"], cfg.eps).view(S, H, hd)
    k = rms_norm(linear(x, w[pre + "k.weight"], w[pre + "k.bias"]), w[pre + "norm_k.weight"], cfg.eps).view(S, H, hd)
    v = linear(x, w[pre + "v.weight"], w[pre + "v.bias"]).view(S, H, hd)
    frame_seqlen = grid[1] * grid[2]
    start_frame = current_start // frame_seqlen
    pos = [start_frame + i for i in range(grid[0])]
    rq = rope_apply(q, grid, freqs, pos)
    rk = rope_apply(k, grid, freqs, pos)
    plan = contiguous_cache_plan(kv.local_end_index, kv.global_end_index, current_start, S)
    ls, le, ws = plan["local_start"], plan["local_end"], plan["win_start"]
    kv.k[ls:le] = rk
    kv.v[ls:le] = v
    if trace is not None:
        trace.append((ls, le, ws))
    out = attention(rq, kv.k[ws:le], kv.v[ws:le])
    kv.global_end_index = plan["global_end"]
    kv.local_end_index = plan["local_end_index"]
    return linear(out.reshape(S, D), w[pre + "o.weight"], w[pre + "o.bias"])


def fps_self_attention(cfg: WanConfig, w: Dict[str, Tensor], pre: str, x: Tensor, grid, freqs: Tensor, kv: KVCache,
                       current_start: Sequence[int], trace: Optional[list] = None) -> Tensor:
    """CausalWanSelfAttention.forward of the frame-slot (MMPL) model, KV-cache branch (causal_fps_model.py:192-264):
    `current_start` lists one token offset per frame; frame f's K/V go to slot f (frames 19, 20 to slots 13, 14), the
    visible set grows by the call's frames and attention reads the visible slots; a call that contains frame 15 (the last
    stage) writes nothing and attends cat(visible slots, its own K/V). RoPE positions are the frames themselves
    (causal_fps_rope_apply :27-55: fp32 x times the complex128 table = the same complex128 product as the contiguous model).
    The reference hard-codes 1560 tokens per frame; here it is the grid's h*w (equal at 60x104)."""
    S, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    fs = grid[1] * grid[2]
    q = rms_norm(linear(x, w[pre + "q.weight"], w[pre + "q.bias"]), w[pre + "norm_q.weight"], cfg.eps).view(S, H, hd)
    k = rms_norm(linear(x, w[pre + "k.weight"], w[pre + "k.bias"]), w[pre + "norm_k.weight"], cfg.eps).view(S, H, hd)
    v = linear(x, w[pre + "v.weight"], w[pre + "v.bias"]).view(S, H, hd)
    starts = [int(s) for s in current_start]
    pos = [s // fs for s in starts]
    rq = rope_apply(q, grid, freqs, pos)
    rk = rope_apply(k, grid, freqs, pos)
    slot_row = lambda s: fps_slot_of_frame(s // fs) * fs
    last_stage = 15 * fs in starts
    if not last_stage:
        for i, s in enumerate(starts):
            kv.k[slot_row(s):slot_row(s) + fs] = rk[i * fs:(i + 1) * fs]
            kv.v[slot_row(s):slot_row(s) + fs] = v[i * fs:(i + 1) * fs]
        kv.visible = list(set(kv.visible + starts))
    else:
        kv.visible = list(set(kv.visible))
    rows = torch.cat([torch.arange(slot_row(s), slot_row(s) + fs, device=x.device) for s in kv.visible])
    if trace is not None:
        trace.append((sorted(kv.visible), last_stage))
    if last_stage:
        out = attention(rq, torch.cat([kv.k[rows], rk]), torch.cat([kv.v[rows], v]))
    else:
        out = attention(rq, kv.k[rows], kv.v[rows])
    return linear(out.reshape(S, D), w[pre + "o.weight"], w[pre + "o.bias"])


def fps_model_forward(cfg: WanConfig, w: Dict[str, Tensor], x: Tensor, t: Tensor, context: Tensor, kv_cache: List[KVCache],
                      cross_cache: List[CrossCache], current_start: Sequence[int], freqs: Optional[Tensor] = None,
                      trace: Optional[list] = None) -> Tensor:
    """CausalFPSWanModel._forward_inference (causal_fps_model.py:780-900: the contiguous model's forward with the frame-slot
    self-attention) for one sample. x [C,F,H,W], t [F], current_start: F token offsets -> flow [C,F,H,W]."""
    return model_forward(cfg, w, x, t, context, kv_cache, cross_cache, list(current_start), freqs, trace, fps_self_attention)


def cross_attention(cfg: WanConfig, w: Dict[str, Tensor], pre: str, x: Tensor, context: Tensor, cc: CrossCache) -> Tensor:
    """WanT2VCrossAttention.forward (model.py:159-194) with the crossattn_cache protocol."""
    S, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    q = rms_norm(linear(x, w[pre + "q.weight"], w[pre + "q.bias"]), w[pre + "norm_q.weight"], cfg.eps).view(S, H, hd)
    if not cc.is_init:
        cc.is_init = True
        cc.k = rms_norm(linear(context, w[pre + "k.weight"], w[pre + "k.bias"]), w[pre + "norm_k.weight"], cfg.eps).view(-1, H, hd)
        cc.v = linear(context, w[pre + "v.weight"], w[pre + "v.bias"]).view(-1, H, hd)
    out = attention(q, cc.k, cc.v)
    return linear(out.reshape(S, D), w[pre + "o.weight"], w[pre + "o.bias"])


def block_forward(cfg: WanConfig, w: Dict[str, Tensor], i: int, x: Tensor, e0: Tensor, grid, freqs: Tensor, context: Tensor,
                  kv: KVCache, cc: CrossCache, current_start, trace: Optional[list] = None, self_attn=None) -> Tensor:
    """CausalWanAttentionBlock.forward (causal_model.py:274-326). x [S,D]; e0 [F,6,D]. `self_attn` selects the
    self-attention restatement (default: the contiguous-cache one; `fps_self_attention` for the frame-slot model)."""
    self_attn = self_attn or self_attention
    b = f"blocks.{i}."
    fs = grid[1] * grid[2]
    e = (w[b + "modulation"].view(1, 6, -1) + e0).unbind(dim=1)  # six [F, D]
    y = self_attn(cfg, w, b + "self_attn.", modulate(layer_norm(x, cfg.eps), e[0], e[1], fs), grid, freqs, kv,
                  current_start, trace)
    x = gate_residual(x, y, e[2], fs)
    x = x + cross_attention(cfg, w, b + "cross_attn.", layer_norm(x, cfg.eps, w[b + "norm3.weight"], w[b + "norm3.bias"]),
                            context, cc)
    h = gelu_tanh(linear(modulate(layer_norm(x, cfg.eps), e[3], e[4], fs), w[b + "ffn.0.weight"], w[b + "ffn.0.bias"]))
    y = linear(h, w[b + "ffn.2.weight"], w[b + "ffn.2.bias"])
    return gate_residual(x, y, e[5], fs)


def patch_embed(cfg: WanConfig, w: Dict[str, Tensor], x: Tensor) -> Tensor:
    """patch_embedding Conv3d with kernel = stride = (1,2,2) (causal_model.py:445-446, 812-816): x [C,F,H,W] ->
    tokens [F*(H/2)*(W/2), D], ordered (f, h, w)."""
    y = F.conv3d(x.unsqueeze(0).float(), w["patch_embedding.weight"].float(), w["patch_embedding.bias"].float(),
                 stride=cfg.patch_size)
    return y.to(x.dtype).flatten(2).transpose(1, 2)[0]


def unpatchify(cfg: WanConfig, x: Tensor, grid) -> Tensor:
    """causal_model.py:1094-1117: [S, out_dim*4] -> [C, F, H, W]."""
    c = cfg.out_dim
    u = x.view(*grid, *cfg.patch_size, c)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(c, *[i * j for i, j in zip(grid, cfg.patch_size)])


def model_forward(cfg: WanConfig, w: Dict[str, Tensor], x: Tensor, t: Tensor, context: Tensor, kv_cache: List[KVCache],
                  cross_cache: List[CrossCache], current_start, freqs: Optional[Tensor] = None,
                  trace: Optional[list] = None, self_attn=None) -> Tensor:
    """CausalWanModel._forward_inference (causal_model.py:763-892) for one sample.
    x [C,F,H,W] latent chunk, t [F] timesteps, context [text_len, text_dim]  ->  flow [C,F,H,W]."""
    if freqs is None:
        freqs = rope_freqs(cfg.head_dim)
    dt = x.dtype
    C, Fn, Hh, Ww = x.shape
    grid = (Fn // cfg.patch_size[0], Hh // cfg.patch_size[1], Ww // cfg.patch_size[2])
    tok = patch_embed(cfg, w, x)
    emb = sinusoidal_embedding_1d(cfg.freq_dim, t.flatten()).to(dt)
    e = linear(silu(linear(emb, w["time_embedding.0.weight"], w["time_embedding.0.bias"])),
               w["time_embedding.2.weight"], w["time_embedding.2.bias"])                      # [F, D]
    e0 = linear(silu(e), w["time_projection.1.weight"], w["time_projection.1.bias"]).view(Fn, 6, cfg.dim)
    ctx = linear(gelu_tanh(linear(context, w["text_embedding.0.weight"], w["text_embedding.0.bias"])),
                 w["text_embedding.2.weight"], w["text_embedding.2.bias"])
    for i in range(cfg.num_layers):
        tok = block_forward(cfg, w, i, tok, e0, grid, freqs, ctx, kv_cache[i], cross_cache[i], current_start, trace, self_attn)
    # CausalHead (causal_model.py:346-357)
    fs = grid[1] * grid[2]
    eh = (w["head.modulation"].view(1, 2, -1) + e.view(Fn, 1, -1)).unbind(dim=1)
    out = linear(modulate(layer_norm(tok, cfg.eps), eh[0], eh[1], fs), w["head.head.weight"], w["head.head.bias"])
    return unpatchify(cfg, out, grid)


# --------------------------------------------------------------------------------------------------
# scheduler / wrapper arithmetic and the chunk-wise pipeline
# --------------------------------------------------------------------------------------------------
class FlowMatchSchedule:
    """FlowMatchScheduler(shift, sigma_min=0, extra_one_step=True).set_timesteps(1000, training=True)
    (utils/scheduler.py:106-133; built at utils/wan_wrapper.py:138-141)."""

    def __init__(self, shift: float = 5.0, num_train_timesteps: int = 1000, sigma_min: float = 0.0, sigma_max: float = 1.0):
        s = torch.linspace(sigma_max, sigma_min, num_train_timesteps + 1)[:-1]
        self.sigmas = shift * s / (1 + (shift - 1) * s)
        self.timesteps = self.sigmas * num_train_timesteps

    def timestep_id(self, timestep: Tensor) -> Tensor:
        ts = self.timesteps.to(timestep.device)
        return torch.argmin((ts.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)

    def add_noise(self, x0: Tensor, noise: Tensor, timestep: Tensor) -> Tensor:
        """utils/scheduler.py:159-176 — evaluated in fp32 because sigma is an fp32 tensor."""
        sigma = self.sigmas.to(noise.device)[self.timestep_id(timestep)].reshape(-1, 1, 1, 1)
        return ((1 - sigma) * x0 + sigma * noise).type_as(noise)

    def flow_to_x0(self, flow: Tensor, xt: Tensor, timestep: Tensor) -> Tensor:
        """WanDiffusionWrapper._convert_flow_pred_to_x0 (utils/wan_wrapper.py:172-196), float64."""
        sig = self.sigmas.double().to(flow.device)
        ts = self.timesteps.double().to(flow.device)
        tid = torch.argmin((ts.unsqueeze(0) - timestep.double().unsqueeze(1)).abs(), dim=1)
        return (xt.double() - sig[tid].reshape(-1, 1, 1, 1) * flow.double()).to(flow.dtype)

    def warped_steps(self, denoising_step_list: Sequence[int]) -> Tensor:
        """pipeline/causal_inference.py:27-31 (warp_denoising_step)."""
        ts = torch.cat((self.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
        return ts[1000 - torch.tensor(list(denoising_step_list), dtype=torch.long)]


def generator_forward(cfg, w, sched: FlowMatchSchedule, noisy: Tensor, context: Tensor, timestep: Tensor, kv, cross,
                      current_start: int, freqs=None, trace=None):
    """WanDiffusionWrapper.forward, KV-cache path (utils/wan_wrapper.py:221-292), batch 1:
    noisy [F,C,H,W], timestep [F] -> (flow [F,C,H,W], x0 [F,C,H,W])."""
    flow = model_forward(cfg, w, noisy.permute(1, 0, 2, 3), timestep, context, kv, cross, current_start, freqs, trace)
    flow = flow.permute(1, 0, 2, 3)
    return flow, sched.flow_to_x0(flow, noisy, timestep)


def causal_inference(cfg: WanConfig, w: Dict[str, Tensor], noise: Tensor, context: Tensor,
                     denoising_step_list=(1000, 750, 500, 250), num_frame_per_block: int = 3, shift: float = 5.0,
                     context_noise: int = 0, cache_rows: Optional[int] = None, rng: Optional[torch.Generator] = None,
                     record: Optional[dict] = None) -> Tensor:
    """CausalInferencePipeline.inference, T2V branch without initial_latent (pipeline/causal_inference.py:47-276),
    batch 1: noise [F,C,H,W] -> denoised latents [F,C,H,W]. `rng` replaces the global torch generator used by
    randn_like (:208). `record` collects the per-call trace: (current_start, timestep) and cache row windows."""
    sched = FlowMatchSchedule(shift)
    steps = sched.warped_steps(denoising_step_list)
    nF, C, Hh, Ww = noise.shape
    fs = (Hh // 2) * (Ww // 2)
    rows = cache_rows if cache_rows is not None else max(32760, nF * fs)
    kv, cross = new_caches(cfg, rows, noise.dtype, noise.device)
    freqs = rope_freqs(cfg.head_dim)
    out = torch.zeros_like(noise)
    calls = [] if record is not None else None
    start_frame = 0
    for _ in range(nF // num_frame_per_block):
        nb = num_frame_per_block
        noisy = noise[start_frame:start_frame + nb]
        for idx, cur_t in enumerate(steps):
            timestep = torch.ones(nb, dtype=torch.int64, device=noise.device) * cur_t
            tr = [] if record is not None else None
            _, x0 = generator_forward(cfg, w, sched, noisy, context, timestep, kv, cross, start_frame * fs, freqs, tr)
            if calls is not None:
                calls.append(dict(current_start=start_frame * fs, timestep=float(cur_t), rows=tr[0], x0=x0.clone()))
            if idx < len(steps) - 1:
                eps = torch.randn(x0.shape, generator=rng, dtype=x0.dtype, device=x0.device) if rng is not None \
                    else torch.randn_like(x0)
                nxt = steps[idx + 1] * torch.ones(nb, device=noise.device, dtype=torch.long)
                noisy = sched.add_noise(x0, eps, nxt)
        out[start_frame:start_frame + nb] = x0
        ctx_t = torch.ones(nb, dtype=torch.int64, device=noise.device) * context_noise
        tr = [] if record is not None else None
        generator_forward(cfg, w, sched, x0, context, ctx_t, kv, cross, start_frame * fs, freqs, tr)
        if calls is not None:
            calls.append(dict(current_start=start_frame * fs, timestep=float(context_noise), rows=tr[0], x0=None))
        start_frame += nb
    if record is not None:
        record["calls"] = calls
        record["kv"] = kv
        record["cross"] = cross
    return out
