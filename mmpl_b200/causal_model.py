"""Drop-in mirror of the reference backbones `CausalWanModel` (wan/modules/causal_model.py:360-1142) and
`CausalFPSWanModel` (wan/modules/causal_fps_model.py:398-1053) for the KV-cache inference path.

Same constructor arguments, same parameter / state-dict names (so reference checkpoints load), same
`forward(x, t, context, seq_len, kv_cache=..., crossattn_cache=..., current_start=..., cache_start=...)` call
and the same in-place protocol on the cache dicts. The torch modules below are parameter containers only;
the arithmetic runs in libmmpl_b200.so (hand-written sm_100a kernels) through `mmpl_forward`. There is no
CPU or eager fallback: calling forward without a CUDA bf16 model raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .cache_plan import AttendPlan, plan_contiguous, plan_fps

__all__ = ["CausalWanModel", "CausalFPSWanModel", "rope_params", "sinusoidal_embedding_1d"]


def sinusoidal_embedding_1d(dim, position):
    """wan/modules/model.py:15-25 (host helper kept for API parity; the forward uses the CUDA kernel)."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """wan/modules/model.py:29-36: complex128 rotation table."""
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


class _RMSNormParams(nn.Module):
    """Parameter holder for WanRMSNorm (model.py:70-86)."""

    def __init__(self, dim, eps):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))


class _AttentionParams(nn.Module):
    """q/k/v/o Linear + norm_q/norm_k, names as in CausalWanSelfAttention / WanT2VCrossAttention."""

    def __init__(self, dim, num_heads, eps):
        super().__init__()
        self.dim, self.num_heads, self.head_dim, self.eps = dim, num_heads, dim // num_heads, eps
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = _RMSNormParams(dim, eps)
        self.norm_k = _RMSNormParams(dim, eps)


class _BlockParams(nn.Module):
    """CausalWanAttentionBlock parameters (causal_model.py:234-271)."""

    def __init__(self, dim, ffn_dim, num_heads, local_attn_size, eps):
        super().__init__()
        self.self_attn = _AttentionParams(dim, num_heads, eps)
        self.self_attn.local_attn_size = local_attn_size
        self.self_attn.max_attention_size = 32760 if local_attn_size == -1 else local_attn_size * 1560
        self.norm3 = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)
        self.cross_attn = _AttentionParams(dim, num_heads, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)


class _HeadParams(nn.Module):
    """CausalHead parameters (causal_model.py:329-344)."""

    def __init__(self, dim, out_dim, patch_size, eps):
        super().__init__()
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)


class CausalWanModel(nn.Module):
    """B200-native CausalWanModel. Constructor mirrors wan/modules/causal_model.py:371-387."""

    is_fps_model = False

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, local_attn_size=-1,
                 sink_size=0, qk_norm=True, cross_attn_norm=True, eps=1e-6):
        super().__init__()
        if model_type != "t2v":
            raise NotImplementedError("only the t2v cross-attention is on the MMPL hot path (SURVEY.md §2: MMPL I2V "
                                      "prefills the first frame through the T2V backbone)")
        if tuple(patch_size) != (1, 2, 2) or dim // num_heads != 128 or not qk_norm or not cross_attn_norm:
            raise NotImplementedError("kernels are specialised for patch (1,2,2), head_dim 128, qk_norm and cross_attn_norm")
        self.model_type, self.patch_size, self.text_len, self.in_dim = model_type, tuple(patch_size), text_len, in_dim
        self.dim, self.ffn_dim, self.freq_dim, self.text_dim, self.out_dim = dim, ffn_dim, freq_dim, text_dim, out_dim
        self.num_heads, self.num_layers, self.local_attn_size, self.sink_size = num_heads, num_layers, local_attn_size, sink_size
        self.qk_norm, self.cross_attn_norm, self.eps = qk_norm, cross_attn_norm, eps

        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([_BlockParams(dim, ffn_dim, num_heads, local_attn_size, eps) for _ in range(num_layers)])
        self.head = _HeadParams(dim, out_dim, patch_size, eps)

        d = dim // num_heads
        # not a buffer, like the reference, so .to(dtype) leaves it complex128 (causal_model.py:470-478)
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1)
        self.init_weights()
        self.gradient_checkpointing = False
        self.block_mask = None
        self.num_frame_per_block = 1
        self.independent_first_frame = False

        # native state
        self._ctx = None
        self._ctx_tokens = 0
        self._bound_sig = None
        self._fused: List[torch.Tensor] = []
        self._rope_table = None
        self._index_mirror: Dict[int, dict] = {}
        self._param_list = None
        self._graphs: Dict[tuple, object] = {}
        self._graph_generation = -1
        self.last_x0 = None

    # ------------------------------------------------------------------------------------------- init
    def init_weights(self):
        """causal_model.py:1119-1141 (note: zeroes the output head)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for m in self.text_embedding.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=.02)
        for m in self.time_embedding.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=.02)
        nn.init.zeros_(self.head.head.weight)

    # ------------------------------------------------------------------------------------- native ctx
    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.load().mmpl_ctx_destroy(self._ctx)
        except Exception:
            pass

    def _param_signature(self):
        """Cheap per-forward fingerprint of the bound parameters: storage pointer (a `.to()` / `.cuda()` moves them) and
        version counter (an in-place update such as load_state_dict or `p.copy_()` bumps it) of every parameter. The q/k/v
        projections are copied into a fused [3D, D] matrix and non-contiguous parameters into contiguous ones, so any
        change must re-bind. Contract: the parameter list is re-read after `_apply` (.to / .cuda / .half), `load_state_dict`
        and `invalidate()`; code that swaps an `nn.Parameter` object or writes through `p.data` calls `invalidate()`."""
        if self._param_list is None:
            self._param_list = list(self.parameters())
        return tuple((p.data_ptr(), p._version) for p in self._param_list)

    def invalidate(self):
        """Forget every host mirror (bound weight pointers, fused q/k/v copies, cache end-index values): the next forward
        re-reads the parameters and the cache dicts. Call after replacing parameters or editing index tensors behind the
        model's back."""
        self._param_list = None
        self._bound_sig = None
        self._index_mirror.clear()
        self._graphs.clear()

    def _apply(self, fn, *args, **kwargs):
        self._param_list = None  # parameters may be replaced by .to() / .cuda()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()  # assign=True replaces the Parameter objects; a plain load bumps their versions
        return out

    def _ensure_ctx(self, device: torch.device, tokens: int):
        lib = _lib.load()
        p0 = self.patch_embedding.weight
        if not p0.is_cuda or p0.dtype != torch.bfloat16:
            raise RuntimeError("mmpl_b200.CausalWanModel runs only as a CUDA bfloat16 model "
                               "(call .to(device='cuda', dtype=torch.bfloat16)); there is no CPU fallback")
        if self._ctx is not None and tokens > self._ctx_tokens:
            lib.mmpl_ctx_destroy(self._ctx)
            self._ctx, self._bound_sig, self._rope_table = None, None, None
            self._graphs.clear()
        if self._ctx is None:
            cap = max(tokens, 3 * 1560)
            cfg = _lib.ModelConfig(self.dim, self.ffn_dim, self.num_heads, self.num_layers, self.freq_dim, self.text_dim,
                                   self.text_len, self.in_dim, self.out_dim, float(self.eps), cap)
            ctx = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(lib.mmpl_ctx_create(C.byref(cfg), C.byref(ctx)))
            self._ctx, self._ctx_tokens = ctx, cap
        sig = self._param_signature()
        if sig != self._bound_sig:
            self._bind_weights(lib)
            self._bound_sig = sig
        if self._rope_table is None or self._rope_table.device != device:
            self._rope_table = torch.view_as_real(self.freqs).contiguous().to(device)
            _lib.check(lib.mmpl_bind_rope_table(self._ctx, self._rope_table.data_ptr()))

    def _bind_weights(self, lib):
        self._graphs.clear()   # captured launch sequences hold the old weight pointers
        self._fused = []
        sd = dict(self.named_parameters())

        def bind(name, t):
            if not t.is_contiguous():
                t = t.contiguous()
                self._fused.append(t)
            _lib.check(lib.mmpl_bind_weight(self._ctx, name.encode(), t.data_ptr(), t.numel()))

        with torch.no_grad():
            for name, p in sd.items():
                if ".self_attn.q." in name or ".self_attn.k." in name or ".self_attn.v." in name:
                    continue
                bind(name, p.data)
            for i, blk in enumerate(self.blocks):
                a = blk.self_attn
                w = torch.cat([a.q.weight.data, a.k.weight.data, a.v.weight.data], dim=0).contiguous()
                b = torch.cat([a.q.bias.data, a.k.bias.data, a.v.bias.data], dim=0).contiguous()
                self._fused += [w, b]
                bind(f"blocks.{i}.self_attn.qkv.weight", w)
                bind(f"blocks.{i}.self_attn.qkv.bias", b)

    def launch_count(self, reset: bool = False) -> int:
        """Kernels launched by this model's context since the last reset."""
        if self._ctx is None:
            return 0
        return int(_lib.load().mmpl_launch_count(self._ctx, int(reset)))

    # ------------------------------------------------------------------------------ cache bookkeeping
    def _read_indices(self, kv_cache: Sequence[dict]):
        """Host mirror of global_end_index / local_end_index. The reference reads both with .item() in every
        layer (causal_model.py:210); here the device values are read only when the pipeline has replaced the
        tensors or written them in place (cache reset), in one batched copy."""
        key = id(kv_cache)
        mir = self._index_mirror.get(key)
        if mir is not None and len(mir["g"]) == len(kv_cache) and all(
                d["global_end_index"] is g and d["local_end_index"] is l and g._version == vg and l._version == vl
                for d, g, l, vg, vl in zip(kv_cache, mir["g"], mir["l"], mir["vg"], mir["vl"])):
            return mir
        g = [d["global_end_index"] for d in kv_cache]
        l = [d["local_end_index"] for d in kv_cache]
        vals = torch.stack([t.reshape(-1)[0] for t in g + l]).tolist()
        n = len(kv_cache)
        mir = {"g": g, "l": l, "gv": [int(v) for v in vals[:n]], "lv": [int(v) for v in vals[n:]],
               "vg": [t._version for t in g], "vl": [t._version for t in l]}
        if len(self._index_mirror) > 8:
            self._index_mirror.clear()
        self._index_mirror[key] = mir
        return mir

    def _plan(self, kv_cache: Sequence[dict], current_start, num_frames: int, frame_seqlen: int) -> AttendPlan:
        cache_rows = kv_cache[0]["k"].shape[1]
        mir = self._read_indices(kv_cache)
        if len(set(mir["gv"])) != 1 or len(set(mir["lv"])) != 1:
            raise NotImplementedError("per-layer KV caches with different end indices are not supported")
        max_att = self.blocks[0].self_attn.max_attention_size
        plan = plan_contiguous(mir["lv"][0], mir["gv"][0], int(current_start), num_frames, frame_seqlen, cache_rows, max_att)
        dg, dl = plan.global_end - mir["gv"][0], plan.local_end - mir["lv"][0]
        # kv_cache["global_end_index"].fill_(current_end); ["local_end_index"].fill_(local_end)  (:225-226)
        if dg:
            torch._foreach_add_(mir["g"], dg)
        if dl:
            torch._foreach_add_(mir["l"], dl)
        n = len(kv_cache)
        mir["gv"], mir["lv"] = [plan.global_end] * n, [plan.local_end] * n
        # someone else writing the index tensors in place (e.g. a reset with .zero_()) shows up as a version mismatch
        mir["vg"], mir["vl"] = [t._version for t in mir["g"]], [t._version for t in mir["l"]]
        return plan

    # ---------------------------------------------------------------------------------------- forward
    def forward(self, *args, **kwargs):
        if kwargs.get("kv_cache", None) is None:
            raise NotImplementedError("only the KV-cache inference path (_forward_inference) is implemented; "
                                      "the training path (_forward_train, flex-attention masks) is out of scope")
        x = args[0] if args else kwargs["x"]
        dev = x.device if torch.is_tensor(x) else x[0].device
        if dev.type == "cuda" and dev.index != torch.cuda.current_device():
            # everything below (context creation, the launches, the stream they go to, the library's per-device caches)
            # belongs to the device that owns the tensors, whatever device is current in the calling thread
            with torch.cuda.device(dev):
                return self._forward_inference(*args, **kwargs)
        return self._forward_inference(*args, **kwargs)

    @torch.no_grad()
    def _forward_inference(self, x, t, context, seq_len=None, clip_fea=None, y=None, kv_cache=None,
                           crossattn_cache=None, current_start=0, cache_start=None, sigma=None, **unused):
        """causal_model.py:763-892. x [B,C,F,H,W] (any strides over B/C/F), t [B,F], context [B,L,text_dim].
        `sigma` (float64 [B,F], optional, not in the reference signature) additionally produces
        x0 = x - sigma*flow in `self.last_x0` ([B,F,C,H,W]) inside the same launch sequence."""
        lib = _lib.load()
        if clip_fea is not None or y is not None:
            raise NotImplementedError("i2v conditioning inputs are not used by MMPL (SURVEY.md §2)")
        if not torch.is_tensor(x):
            x = torch.stack(list(x))
        B, Cc, Fn, Hh, Ww = x.shape
        gh, gw = Hh // 2, Ww // 2
        fs = gh * gw
        S = Fn * fs
        dev = x.device
        if not x.is_cuda or x.dtype != torch.bfloat16:
            raise RuntimeError("input latents must be CUDA bfloat16 tensors (the reference path is bf16-only, SURVEY.md F7)")
        self._ensure_ctx(dev, S)
        if x.stride(4) != 1 or x.stride(3) != Ww:
            x = x.contiguous()
        if not torch.is_tensor(context):
            context = torch.stack([torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context])
        need_cross = not all(c["is_init"] for c in crossattn_cache)
        if need_cross:
            if context.shape[1] < self.text_len:
                context = torch.cat([context, context.new_zeros(B, self.text_len - context.shape[1], context.shape[2])], 1)
            context = context.to(torch.bfloat16).contiguous()
            H = self.num_heads
            for c in crossattn_cache:
                for n in ("k", "v"):
                    tsr = c.get(n)
                    if not (torch.is_tensor(tsr) and tsr.shape == (B, self.text_len, H, 128) and tsr.is_cuda
                            and tsr.dtype == torch.bfloat16 and tsr.is_contiguous()):
                        c[n] = torch.empty((B, self.text_len, H, 128), dtype=torch.bfloat16, device=dev)
        # t is [B, F], or [B, 1] broadcast over the frames (the pipelines' t = 0 prefill calls, causal_inference.py:137-169;
        # the reference broadcasts e0 [B, 1, 6, D] over the frame axis, causal_model.py:297-305)
        t64 = t.reshape(B, -1).to(device=dev, dtype=torch.float64).expand(B, Fn).contiguous()
        if sigma is not None:
            sigma = sigma.reshape(B, -1).to(device=dev, dtype=torch.float64).expand(B, Fn).contiguous()
            x0 = torch.empty((B, Fn, Cc, Hh, Ww), dtype=torch.bfloat16, device=dev)
        flow = torch.empty((B, Fn, self.out_dim, Hh, Ww), dtype=torch.bfloat16, device=dev)

        plan = self._make_plan(kv_cache, current_start, Fn, fs)
        for d in kv_cache:
            if d["k"].dtype != torch.bfloat16 or not d["k"].is_contiguous() or not d["v"].is_contiguous():
                raise RuntimeError("kv_cache tensors must be contiguous bfloat16 [B, rows, heads, 128]")
        ptrs = tuple(t.data_ptr() for d in kv_cache for t in (d["k"], d["v"])) + \
            tuple(t.data_ptr() for c in crossattn_cache for t in (c["k"], c["v"]))
        io = dict(x=x, t64=t64, sigma=sigma, flow=flow, x0=x0 if sigma is not None else None)
        if self._graph_ok(S, B, need_cross):
            key = (B, Cc, Fn, Hh, Ww, tuple(plan.frame_pos), tuple(plan.kv_row), tuple(plan.segments), bool(plan.kv_to_tail),
                   sigma is not None, ptrs)
            io = self._run_graphed(key, io, plan, kv_cache, crossattn_cache, dev)
        else:
            self._launch(io, context, plan, kv_cache, crossattn_cache, need_cross, dev)
        if need_cross:
            for c in crossattn_cache:
                c["is_init"] = True  # model.py:175
        self.last_x0 = io["x0"]
        return io["flow"].permute(0, 2, 1, 3, 4)

    def _launch(self, io, context, plan, kv_cache, crossattn_cache, need_cross, dev):
        """One mmpl_forward per sample on the current stream of `dev`."""
        lib = _lib.load()
        x, t64, sigma, flow, x0 = io["x"], io["t64"], io["sigma"], io["flow"], io["x0"]
        B, _, Fn, Hh, Ww = x.shape
        L = self.num_layers
        cache_rows = kv_cache[0]["k"].shape[1]
        stream = torch.cuda.current_stream(dev).cuda_stream
        n_seg = len(plan.segments)
        seg_start = (C.c_int * max(1, n_seg))(*[s for s, _ in plan.segments])
        seg_rows = (C.c_int * max(1, n_seg))(*[r for _, r in plan.segments])
        frame_pos = (C.c_int * Fn)(*plan.frame_pos)
        kv_row = (C.c_int * Fn)(*plan.kv_row)
        for b in range(B):
            kv_k = (C.c_void_p * L)(*[d["k"][b].data_ptr() for d in kv_cache])
            kv_v = (C.c_void_p * L)(*[d["v"][b].data_ptr() for d in kv_cache])
            ck = (C.c_void_p * L)(*[c["k"][b].data_ptr() for c in crossattn_cache])
            cv = (C.c_void_p * L)(*[c["v"][b].data_ptr() for c in crossattn_cache])
            args = _lib.ForwardArgs(
                latents=x[b].data_ptr(), lat_stride_f=x.stride(2), lat_stride_c=x.stride(1),
                n_frames=Fn, lat_h=Hh, lat_w=Ww, timesteps=t64[b].data_ptr(),
                context=context[b].data_ptr() if need_cross else None,
                kv_k=kv_k, kv_v=kv_v, cache_rows=cache_rows, frame_pos=frame_pos, kv_row=kv_row,
                kv_to_tail=int(plan.kv_to_tail), n_seg=n_seg, seg_start=seg_start, seg_rows=seg_rows,
                cross_k=ck, cross_v=cv, cross_init=0 if need_cross else 1,
                flow=flow[b].data_ptr(), x0=x0[b].data_ptr() if sigma is not None else None,
                sigma=sigma[b].data_ptr() if sigma is not None else None)
            _lib.check(lib.mmpl_forward(self._ctx, C.byref(args), stream))

    # --------------------------------------------------------------------------------------------- CUDA graphs
    # A forward is ~13 launches per block. At the benchmark's sizes (S = 4680, >= 16 ms of kernels per forward) the host
    # stays far ahead of the GPU; at small ones (BASELINE config 0: S = 1170, ~3 ms of kernels behind ~400 launches) the
    # launches themselves are the time. There the whole launch sequence of a forward is captured once per call shape -
    # (latent shape, RoPE positions, cache rows written and attended, cache tensors) - into a CUDA graph over static
    # input / output buffers and replayed: same kernels, same order, same arguments, one launch from the host.
    graph_max_tokens = 2048     # forwards with more new tokens than this are not launch-bound: no capture
    _graph_cap = 64

    def _graph_ok(self, S: int, B: int, need_cross: bool) -> bool:
        if S > self.graph_max_tokens or need_cross or os.environ.get("MMPL_CUDA_GRAPHS", "1") == "0":
            return False
        if torch.cuda.is_current_stream_capturing():
            return False       # the caller is capturing a graph of its own: just launch into it
        return _lib.load().mmpl_profile_mask(self._ctx) == 0    # per-launch timing events cannot live inside a graph

    def _run_graphed(self, key, io, plan, kv_cache, crossattn_cache, dev):
        lib = _lib.load()
        generation = lib.mmpl_workspace_generation()
        if generation != self._graph_generation:   # a workspace of the library moved: every captured pointer to it is stale
            self._graphs.clear()
            self._graph_generation = generation
        entry = self._graphs.get(key)
        if entry is None:
            # first sighting: run eagerly (lazy initialisation inside the library - workspaces, tensor maps, function
            # attributes - must not happen during a capture), remember the shape
            if len(self._graphs) >= self._graph_cap:
                self._graphs.clear()
            self._graphs[key] = False
            self._launch(io, None, plan, kv_cache, crossattn_cache, False, dev)
            return io
        if entry is False:
            static = {k: (None if v is None else torch.empty_like(v, memory_format=torch.contiguous_format)) for k, v in io.items()}
            for k in ("x", "t64", "sigma"):
                if io[k] is not None:
                    static[k].copy_(io[k])
            graph = torch.cuda.CUDAGraph()
            before = lib.mmpl_launch_count(self._ctx, 0)
            # thread-local capture mode: other threads of the process (NCCL's watchdog, a second pipeline of a threaded
            # driver) keep making CUDA calls while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._launch(static, None, plan, kv_cache, crossattn_cache, False, dev)
            entry = self._graphs[key] = dict(graph=graph, static=static, launches=lib.mmpl_launch_count(self._ctx, 0) - before)
            lib.mmpl_launch_credit(self._ctx, -entry["launches"])   # counted at capture, executed only by the replay below
        static = entry["static"]
        for k in ("x", "t64", "sigma"):
            if io[k] is not None:
                static[k].copy_(io[k])
        entry["graph"].replay()
        lib.mmpl_launch_credit(self._ctx, entry["launches"])
        # the static outputs are overwritten by the next replay of this shape: hand out copies
        return dict(io, flow=static["flow"].clone(), x0=None if static["x0"] is None else static["x0"].clone())

    def _make_plan(self, kv_cache, current_start, num_frames, frame_seqlen) -> AttendPlan:
        if isinstance(current_start, (list, tuple)) or (torch.is_tensor(current_start) and current_start.dim() > 0):
            raise TypeError("CausalWanModel takes an integer current_start; per-frame lists belong to CausalFPSWanModel")
        return self._plan(kv_cache, int(current_start), num_frames, frame_seqlen)


class CausalFPSWanModel(CausalWanModel):
    """MMPL anchor / macro-from-micro backbone (wan/modules/causal_fps_model.py:398): same blocks, but
    `current_start` is a list of per-frame token offsets, K/V are written by frame slot, attention reads the
    slots listed in kv_cache[i]["attention_vis_index"] (no gather copy: row segments), the last stage attends
    cache + new K/V without writing, and the end indices are never touched."""

    is_fps_model = True

    def init_weights(self):
        """causal_fps_model.py:1032-1053: as CausalWanModel.init_weights but the head is not zeroed."""
        super().init_weights()
        nn.init.xavier_uniform_(self.head.head.weight)

    def _make_plan(self, kv_cache, current_start, num_frames, frame_seqlen) -> AttendPlan:
        cur = current_start.tolist() if hasattr(current_start, "tolist") else list(current_start)
        if len(cur) != num_frames:
            raise ValueError(f"current_start lists {len(cur)} frames but the latent chunk has {num_frames}")
        cache_rows = kv_cache[0]["k"].shape[1]
        plans = []
        for d in kv_cache:
            vis = d["attention_vis_index"]
            plans.append(plan_fps(vis, cur, frame_seqlen, cache_rows))
        p0 = plans[0]
        if any(p.segments != p0.segments or p.kv_row != p0.kv_row for p in plans[1:]):
            raise NotImplementedError("per-layer attention_vis_index lists that differ are not supported")
        if len(p0.segments) + int(p0.kv_to_tail) > 8:
            raise NotImplementedError("more than 8 visible row runs")
        return p0
