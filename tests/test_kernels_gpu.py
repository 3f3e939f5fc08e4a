"""Per-kernel parity on the GPU: every C-ABI kernel entry point against the oracle primitive it replaces
(oracle/causal_wan_oracle.py, evaluated with torch ops on the same device for speed)."""
import math
import types

import pytest
import torch

from oracle import causal_wan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from mmpl_b200 import ops
    return ops


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _report(name, got, ref, atol, rtol, outlier_frac=0.0, outlier_abs=0.0):
    """Element-wise |got-ref| <= atol + rtol*|ref|. `outlier_frac` of the elements may instead be within
    `outlier_abs`: a 1-ulp flip of a bf16 intermediate (fp32 reduction order) feeding a cancelling add/rotation gives a
    result that is off by one ulp of the *intermediate*, which is unbounded relative to a small result."""
    got_f, ref_f = got.float(), ref.float()
    err = (got_f - ref_f).abs()
    tol = atol + rtol * ref_f.abs()
    bad = (err > tol).sum().item()
    cos = torch.nn.functional.cosine_similarity(got_f.flatten(), ref_f.flatten(), dim=0).item()
    print(f"{name}: max_abs={err.max().item():.4g} mean_abs={err.mean().item():.4g} cos={cos:.6f} bad={bad}/{err.numel()}")
    assert torch.isfinite(got_f).all(), f"{name}: non-finite output"
    assert bad <= outlier_frac * err.numel() and (bad == 0 or err.max().item() <= outlier_abs), \
        f"{name}: {bad} elements out of tolerance (max_abs={err.max().item():.4g}, cos={cos:.6f})"


@pytest.mark.parametrize("M,N,K,tile", [
    (128, 128, 64, 128), (128, 256, 64, 256), (128, 64, 64, 64),
    (300, 256, 128, 0), (300, 256, 192, 64), (300, 256, 192, 128), (300, 512, 192, 256),
    (1170, 1536, 1536, 0), (4680, 1536, 1536, 0), (4680, 4608, 1536, 0), (4680, 8960, 1536, 0), (4680, 1536, 8960, 0),
    (4680, 64, 1536, 0), (4680, 1536, 64, 0), (512, 1536, 4096, 0),
    # tile 512 = the cta_group::2 kernel (256 x 256 per CTA pair); 256 = force the single-CTA kernel
    (256, 256, 64, 512), (300, 512, 192, 512), (4680, 1536, 1536, 512), (4680, 4608, 1536, 512), (4680, 8960, 1536, 512),
    (4680, 1536, 8960, 512), (10920, 5120, 5120, 512), (4680, 1536, 1536, 256), (4680, 1536, 1536, 128),
    # tile 448 = the cta_group::2 kernel with 256 x 224 tiles (N tail: 1536 = 6 x 224 + 192, 512 = 2 x 224 + 64)
    (256, 224, 64, 448), (300, 512, 192, 448), (4680, 1536, 1536, 448), (4680, 1536, 8960, 448), (4680, 4608, 1536, 448),
    (1170, 1536, 1536, 448), (10920, 5120, 5120, 448),
    # tile 192 = single-CTA 128 x 192 tiles (N tail: 512 = 2 x 192 + 128)
    (128, 192, 64, 192), (300, 512, 192, 192), (4680, 1536, 1536, 192), (1170, 1536, 1536, 192),
    # tiles 1128 / 1192 = clusters of two 128 x 128 / 128 x 192 tiles sharing the A tile by TMA multicast (M tails of 44 and
    # 72 rows: the second half of the A tile partly or wholly out of bounds; N tail: 640 = 3 x 192 + 64)
    (128, 256, 64, 1128), (300, 512, 192, 1128), (4680, 1536, 1536, 1128), (1170, 1536, 1536, 1128), (10920, 5120, 5120, 1128),
    (128, 384, 64, 1192), (300, 640, 192, 1192), (4680, 1536, 1536, 1192), (1170, 1536, 1536, 1192), (300, 768, 1024, 1192),
])
def test_gemm_bias(M, N, K, tile):
    ops = _ops()
    x, w, b = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, scale=0.1, seed=3)
    got = ops.linear(x, w, b, tile_n=tile)
    _report(f"gemm {M}x{N}x{K} tile={tile}", got, O.linear(x, w, b), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("M,N,K,tile", [(1170, 1536, 1536, 0), (4680, 1536, 1536, 0), (4680, 1536, 1536, 128), (1170, 1536, 1536, 512),
                                        (4680, 1536, 1536, 448), (4680, 1536, 8960, 0), (1170, 1536, 1536, 448),
                                        (4680, 1536, 1536, 192), (4680, 1536, 1536, 1128), (4680, 1536, 1536, 1192)])
def test_gemm_epilogues(M, N, K, tile):
    import functools
    ops = _ops()
    ops = types.SimpleNamespace(**{k: getattr(ops, k) for k in dir(ops) if not k.startswith("__")})
    ops.linear = functools.partial(ops.linear, tile_n=tile)
    frames, fs = 3, M // 3
    x, w, b = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, scale=0.1, seed=3)
    res, gate = _rand(M, N, seed=4), _rand(frames, 6, N, seed=5)
    y = O.linear(x, w, b)
    _report("gelu", ops.linear(x, w, b, epilogue=ops.EPI_BIAS_GELU), O.gelu_tanh(y), 2e-2, 2e-2)
    _report("silu", ops.linear(x, w, b, epilogue=ops.EPI_BIAS_SILU), O.silu(y), 2e-2, 2e-2)
    _report("res", ops.linear(x, w, b, epilogue=ops.EPI_BIAS_RES, residual=res), res + y, 3e-2, 2e-2)
    g = gate[:, 2]
    ref = O.gate_residual(res, y, g, fs)
    _report("gate_res", ops.linear(x, w, b, epilogue=ops.EPI_BIAS_GATE_RES, residual=res, gate=g, rows_per_frame=fs), ref, 4e-2, 2e-2)
    # in place on the residual stream, as the forward uses it
    xres = res.clone()
    ops.linear(x, w, b, epilogue=ops.EPI_BIAS_GATE_RES, residual=xres, gate=g, rows_per_frame=fs, out=xres)
    _report("gate_res_inplace", xres, ref, 4e-2, 2e-2)


@pytest.mark.parametrize("M,N,K,tile", [(4680, 1536, 1536, 512), (4680, 1536, 8960, 512), (4680, 4608, 1536, 512),
                                        (10920, 5120, 5120, 512), (1170, 1536, 1536, 512), (300, 512, 1024, 512),
                                        (1998, 1280, 512, 512), (4680, 1536, 8960, 448), (1170, 1536, 1536, 448)])
def test_gemm_streamk_schedule(M, N, K, tile):
    """The stream-K tail schedule of the cta_group::2 kernel (gemm_tcgen05.cu: PairSched) against the whole-tile
    schedule and the oracle: forced on (mode 1) for shapes of 1 ... 12 waves incl. fewer tiles than CTA pairs; repeated
    launches and a CUDA-graph replay check that the hand-over flags are re-armed; results must be deterministic."""
    from mmpl_b200 import _lib
    ops, lib = _ops(), _lib.load()
    x, w, b = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, scale=0.1, seed=3)
    frames, fs = 3, M // 3
    res, gate = _rand(M, N, seed=4), _rand(frames, N, seed=5)
    ref = O.linear(x, w, b)
    ref_gr = O.gate_residual(res, ref, gate, fs)
    try:
        lib.mmpl_gemm_set_streamk(0)
        dp = ops.linear(x, w, b, tile_n=tile)
        lib.mmpl_gemm_set_streamk(1)
        outs = [ops.linear(x, w, b, tile_n=tile) for _ in range(3)]
        gr = ops.linear(x, w, b, epilogue=ops.EPI_BIAS_GATE_RES, residual=res, gate=gate, rows_per_frame=fs, tile_n=tile)
        gelu = ops.linear(x, w, b, epilogue=ops.EPI_BIAS_GELU, tile_n=tile)
        out_g = torch.empty_like(outs[0])
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ops.linear(x, w, b, out=out_g, tile_n=tile)
            ops.linear(x, w, b, out=out_g, tile_n=tile)
        for _ in range(3):
            out_g.zero_()
            graph.replay()
        torch.cuda.synchronize()
    finally:
        lib.mmpl_gemm_set_streamk(0)  # the library default
    _report(f"streamk {M}x{N}x{K}", outs[0], ref, atol=2e-2, rtol=2e-2)
    _report("streamk vs whole tiles", outs[0], dp, atol=2e-2, rtol=1e-2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "stream-K result changes between launches"
    assert torch.equal(outs[0], out_g), "stream-K result differs under CUDA-graph replay"
    _report("streamk gate_res", gr, ref_gr, 4e-2, 2e-2)
    _report("streamk gelu", gelu, O.gelu_tanh(ref), 2e-2, 2e-2)


@pytest.mark.parametrize("Lq,Lk,H", [(128, 128, 1), (256, 128, 2), (200, 300, 2), (390, 512, 2), (1170, 1170, 12),
                                     (4680, 4680, 12), (4680, 512, 12), (1560, 9360, 4)])
def test_flash_attn(Lq, Lk, H):
    ops = _ops()
    q, k, v = _rand(Lq, H, 128, seed=1), _rand(Lk, H, 128, seed=2), _rand(Lk, H, 128, seed=3)
    got = ops.flash_attn(q, k, v)
    _report(f"attn {Lq}x{Lk}x{H}", got, O.attention(q, k, v), atol=1e-2, rtol=2e-2)


def test_flash_attn_strided_window_and_segments():
    ops = _ops()
    H, D = 2, 256
    qkv = _rand(500, 3 * D, seed=1)
    q = qkv[:, :D].view(500, H, 128)  # row pitch 3D, as the forward passes it
    cache_k, cache_v = _rand(2000, H, 128, seed=2), _rand(2000, H, 128, seed=3)
    got = ops.flash_attn(q, cache_k, cache_v, segments=[(100, 700)])
    _report("window", got, O.attention(q, cache_k[100:800], cache_v[100:800]), 1e-2, 2e-2)
    segs = [(0, 390), (780, 200), (1500, 130)]
    idx = torch.cat([torch.arange(s, s + n) for s, n in segs]).to(DEV)
    got = ops.flash_attn(q, cache_k, cache_v, segments=segs)
    _report("segments", got, O.attention(q, cache_k[idx], cache_v[idx]), 1e-2, 2e-2)
    tk, tv = _rand(260, H, 128, seed=4), _rand(260, H, 128, seed=5)
    got = ops.flash_attn(q, cache_k, cache_v, segments=[(0, 390), (0, 260, 1)], k_tail=tk, v_tail=tv)
    ref = O.attention(q, torch.cat([cache_k[:390], tk]), torch.cat([cache_v[:390], tv]))
    _report("tail", got, ref, 1e-2, 2e-2)


def test_flash_attention_call_site_wrapper():
    """mmpl_b200.attention.flash_attention / attention - the reference call site (wan/modules/attention.py:32-185) - on the
    GPU: batch of 2, per-sample k_lens and q_lens (rows beyond q_lens come back as zeros, as the reference's packing leaves
    them), strided q/k/v views (slices of a fused projection and of a larger cache), q_scale / softmax_scale, fp16 inputs
    returned in q's dtype, and the hot path's plain (q, k, v) call. Each sample against the oracle's attention."""
    from mmpl_b200.attention import attention, flash_attention
    B, Lq, Lk, H = 2, 300, 700, 3
    qkv = _rand(B, Lq, 3 * H * 128, seed=1)
    q = qkv[..., :H * 128].view(B, Lq, H, 128)                      # row pitch 3*H*128
    cache = _rand(B, 1000, 2, H, 128, seed=2)
    k, v = cache[:, 100:100 + Lk, 0], cache[:, 100:100 + Lk, 1]     # strided in the row dimension
    out = flash_attention(q, k, v)
    assert out.shape == q.shape and out.dtype == torch.bfloat16
    for b in range(B):
        _report(f"call site sample {b}", out[b], O.attention(q[b], k[b], v[b]), 1e-2, 2e-2)
    q_lens, k_lens = torch.tensor([300, 170]), torch.tensor([700, 333])
    out = attention(q, k, v, q_lens=q_lens, k_lens=k_lens)
    for b in range(B):
        nq, nk = int(q_lens[b]), int(k_lens[b])
        _report(f"call site lens sample {b}", out[b, :nq], O.attention(q[b, :nq], k[b, :nk], v[b, :nk]), 1e-2, 2e-2)
        assert (out[b, nq:] == 0).all()
    # cross-attention's call: k_lens=None over all context rows (model.py:189)
    ctx_k, ctx_v = _rand(B, 512, H, 128, seed=3), _rand(B, 512, H, 128, seed=4)
    out = flash_attention(q, ctx_k, ctx_v, k_lens=None)
    _report("call site cross", out[1], O.attention(q[1], ctx_k[1], ctx_v[1]), 1e-2, 2e-2)
    # q_scale and softmax_scale
    out = flash_attention(q, k, v, q_scale=0.5, softmax_scale=0.5 / math.sqrt(128))
    _report("call site scales", out[0], O.attention(q[0] * 0.25, k[0], v[0]), 1e-2, 2e-2)   # powers of two: exact in bf16
    # half-precision inputs of the other kind come back in their own dtype
    out16 = flash_attention(q.to(torch.float16), k.to(torch.float16), v.to(torch.float16))
    assert out16.dtype == torch.float16
    _report("call site fp16", out16[0], O.attention(q[0], k[0], v[0]), 1e-2, 2e-2)
    with pytest.raises(NotImplementedError):
        flash_attention(q, k, v, causal=True)
    with pytest.raises(NotImplementedError):
        flash_attention(q, k, v, window_size=(128, 0))


@pytest.mark.parametrize("split", [1, 2, 3, 5])
def test_flash_attn_forced_kv_split(split):
    """Every unit cut into `split` KV chunks and merged by the combine kernel must equal the unsplit result."""
    from mmpl_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    q, k, v = _rand(700, 3, 128, seed=1), _rand(5000, 3, 128, seed=2), _rand(5000, 3, 128, seed=3)
    try:
        lib.mmpl_attn_set_split(split)
        got = ops.flash_attn(q, k, v, segments=[(0, 1300), (2000, 2900)])
    finally:
        lib.mmpl_attn_set_split(0)
    idx = torch.cat([torch.arange(0, 1300), torch.arange(2000, 4900)]).to(DEV)
    _report(f"attn split={split}", got, O.attention(q, k[idx], v[idx]), atol=1e-2, rtol=2e-2)


@pytest.mark.parametrize("Lq,Lk,H,hg", [(700, 5000, 3, 1), (700, 5000, 3, 2), (700, 5000, 3, 3), (4680, 4680, 12, 12),
                                        (4680, 9360, 12, 5), (1560, 20000, 4, 1), (300, 2100, 2, 2), (3120, 14040, 8, 3)])
def test_flash_attn_range_schedule(Lq, Lk, H, hg):
    """The range schedule (attention_tcgen05.cu: PieceIter with hg > 0): every CTA takes an equal contiguous range of
    each head group's (unit, KV tile) sequence; units that straddle a range boundary are merged by the combine
    kernel. Forced with a negative split (= heads per group), incl. group sizes that do not divide the head count,
    against the oracle and against the whole-unit schedule; also over row segments."""
    from mmpl_b200 import _lib
    ops, lib = _ops(), _lib.load()
    q, k, v = _rand(Lq, H, 128, seed=1), _rand(Lk, H, 128, seed=2), _rand(Lk, H, 128, seed=3)
    segs = [(0, Lk // 3 + 17), (Lk // 2, Lk // 2 - 5)]
    idx = torch.cat([torch.arange(a, a + n) for a, n in segs]).to(DEV)
    try:
        lib.mmpl_attn_set_split(1)
        whole = ops.flash_attn(q, k, v)
        lib.mmpl_attn_set_split(-hg)
        got = ops.flash_attn(q, k, v)
        got2 = ops.flash_attn(q, k, v)
        got_seg = ops.flash_attn(q, k, v, segments=segs)
    finally:
        lib.mmpl_attn_set_split(0)
    _report(f"attn ranges hg={hg} {Lq}x{Lk}x{H}", got, O.attention(q, k, v), atol=1e-2, rtol=2e-2)
    _report("ranges vs whole units", got, whole, atol=1e-2, rtol=2e-2)
    assert torch.equal(got, got2), "range schedule is not deterministic"
    _report("ranges over segments", got_seg, O.attention(q, k[idx], v[idx]), atol=1e-2, rtol=2e-2)


@pytest.mark.parametrize("Lq,Lk,H,ctas", [(700, 5000, 3, 4), (700, 5000, 3, 7), (1300, 2100, 5, 6), (4680, 4680, 12, 0),
                                          (4680, 18720, 12, 0), (3120, 6000, 40, 0), (600, 40000, 2, 5)])
def test_flash_attn_hybrid_schedule(Lq, Lk, H, ctas):
    """The hybrid schedule (attention_tcgen05.cu: u_base > 0): floor(U/G) rounds of whole units, the other U mod G
    units cut into equal ranges and merged inside the kernel. Forced with split <= -1000; `ctas` shrinks the persistent
    grid so that small problems have more units than CTAs (0 = one CTA per SM: the cfg2 / 14B-width shapes, where the
    cost model picks this schedule by itself). Checked against the oracle, the whole-unit schedule, for determinism
    and over row segments."""
    from mmpl_b200 import _lib
    ops, lib = _ops(), _lib.load()
    q, k, v = _rand(Lq, H, 128, seed=1), _rand(Lk, H, 128, seed=2), _rand(Lk, H, 128, seed=3)
    segs = [(0, Lk // 3 + 17), (Lk // 2, Lk // 2 - 5)]
    idx = torch.cat([torch.arange(a, a + n) for a, n in segs]).to(DEV)
    try:
        lib.mmpl_attn_set_ctas(ctas)
        lib.mmpl_attn_set_split(1)
        whole = ops.flash_attn(q, k, v)
        lib.mmpl_attn_set_split(-1000)
        got = ops.flash_attn(q, k, v)
        got2 = ops.flash_attn(q, k, v)
        got_seg = ops.flash_attn(q, k, v, segments=segs)
        lib.mmpl_attn_set_split(0)
        auto = ops.flash_attn(q, k, v)
    finally:
        lib.mmpl_attn_set_split(0)
        lib.mmpl_attn_set_ctas(0)
    ref = O.attention(q, k, v)
    _report(f"attn hybrid {Lq}x{Lk}x{H} ctas={ctas}", got, ref, atol=1e-2, rtol=2e-2)
    _report("hybrid vs whole units", got, whole, atol=1e-2, rtol=2e-2)
    _report("cost-model schedule", auto, ref, atol=1e-2, rtol=2e-2)
    assert torch.equal(got, got2), "hybrid schedule is not deterministic"
    _report("hybrid over segments", got_seg, O.attention(q, k[idx], v[idx]), atol=1e-2, rtol=2e-2)


@pytest.mark.parametrize("half", [0, 1])
def test_flash_attn_both_pipelines_variant(half):
    """The library holds two builds of flash_attn_kernel (attention_tcgen05.cu: MMPL_ATTN_SPLIT_S): the whole-tile S
    hand-over (the default) and the half-tile S pipeline (opt-in).
    MMPL_ATTN_HALF=0|1 forces one for every call (read once per process, hence the subprocess): every attention test of
    this file - all shapes, forced splits, range and hybrid schedules, segments, lazy rescale - must pass with either."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, MMPL_ATTN_HALF=str(half))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "flash_attn and not variant and not merge_kernel_fallback"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_flash_attn_merge_kernel_fallback():
    """Partial pieces are merged inside flash_attn_kernel (default for the range schedule) or by attn_combine_kernel
    (default for the uniform split); MMPL_ATTN_MERGE=inline|kernel forces one for both schedules (read once per process,
    hence the subprocess). Both must give the same result for both schedules."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from mmpl_b200 import ops, _lib\n"
        "g = torch.Generator().manual_seed(0)\n"
        "q, k, v = (torch.randn(n, 3, 128, generator=g).to(torch.bfloat16).cuda() for n in (700, 5000, 5000))\n"
        "lib = _lib.load(); outs = []\n"
        "for sp in (3, -3):\n"
        "    lib.mmpl_attn_set_split(sp); outs.append(ops.flash_attn(q, k, v).float().cpu())\n"
        "torch.save(outs, sys.argv[1])\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = {}
    for mode in ("inline", "kernel"):
        path = f"/tmp/mmpl_merge_{mode}_{os.getpid()}.pt"
        env = dict(os.environ, MMPL_ATTN_MERGE=mode)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        res[mode] = torch.load(path)
        os.remove(path)
    for a, b in zip(res["inline"], res["kernel"]):
        assert torch.isfinite(a).all() and (a - b).abs().max().item() <= 1e-2, "in-kernel merge and attn_combine_kernel disagree"


def test_flash_attn_large_logits():
    """Rows whose running max grows a lot between KV tiles exercise the lazy O rescale."""
    ops = _ops()
    q, k, v = _rand(256, 1, 128, seed=1), _rand(1024, 1, 128, seed=2), _rand(1024, 1, 128, seed=3)
    k[512:] *= 6.0
    got = ops.flash_attn(q, k, v)
    _report("attn large logits", got, O.attention(q, k, v), 1e-2, 2e-2)


@pytest.mark.parametrize("S,D,frames", [(1170, 1536, 3), (4680, 1536, 3), (780, 256, 2), (3120, 5120, 2)])
def test_ln_modulate_and_affine(S, D, frames):
    ops = _ops()
    x = _rand(S, D, seed=1) * 3 + 0.5
    e = _rand(frames, 6, D, scale=0.5, seed=2)
    fs = S // frames
    got = ops.ln_modulate(x, e[:, 0], e[:, 1], fs)
    ref = O.modulate(O.layer_norm(x, 1e-6), e[:, 0], e[:, 1], fs)
    # bf16(bf16(n)*s)+b: a 1-ulp flip of an intermediate (fp32 reduction order) can move the result by 2 ulp
    _report("ln_modulate", got, ref, atol=1e-6, rtol=1.6e-2, outlier_frac=1e-5, outlier_abs=0.0625)
    assert (got == ref).float().mean().item() > 0.9999
    w, b = _rand(D, seed=3), _rand(D, seed=4)
    _report("ln_affine", ops.ln_affine(x, w, b), O.layer_norm(x, 1e-6, w, b), atol=1e-6, rtol=8e-3)


@pytest.mark.parametrize("S,D", [(1170, 1536), (512, 1536), (300, 256), (512, 5120)])
def test_rmsnorm(S, D):
    ops = _ops()
    x, w = _rand(S, D, seed=1) * 2, _rand(D, seed=2)
    _report("rmsnorm", ops.rmsnorm(x, w), O.rms_norm(x, w, 1e-6), atol=1e-6, rtol=8e-3)


@pytest.mark.parametrize("D,H,grid,frame_pos,kv_row", [
    (256, 2, (3, 6, 5), [0, 1, 2], [0, 30, 60]),
    (1536, 12, (3, 15, 26), [3, 4, 5], [1170, 1560, 1950]),
    (1536, 12, (2, 30, 52), [19, 20], [13 * 1560, 14 * 1560]),
])
def test_qk_norm_rope_kv(D, H, grid, frame_pos, kv_row):
    ops = _ops()
    f, gh, gw = grid
    S = f * gh * gw
    qkv = _rand(S, 3 * D, seed=1)
    wq, wk = _rand(D, seed=2), _rand(D, seed=3)
    freqs = O.rope_freqs(128)
    table = O.rope_table_real(freqs).to(DEV)
    rows = max(kv_row) + gh * gw + 7
    kc = torch.zeros(rows, D, dtype=torch.bfloat16, device=DEV)
    vc = torch.zeros_like(kc)
    q_in, k_in, v_in = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    q_ref = O.rope_apply(O.rms_norm(q_in, wq, 1e-6).view(S, H, 128), grid, freqs, frame_pos).reshape(S, D)
    k_ref = O.rope_apply(O.rms_norm(k_in, wk, 1e-6).view(S, H, 128), grid, freqs, frame_pos).reshape(S, D)
    v_ref = v_in.clone()
    q_out = ops.qk_norm_rope_kv(q_in, k_in, v_in, wq, wk, table, kc, vc, (gh, gw), frame_pos, kv_row, q_out=q_in)
    assert q_out.data_ptr() == qkv.data_ptr()
    _report("roped q", q_out, q_ref, atol=1e-6, rtol=1.6e-2, outlier_frac=1e-5, outlier_abs=0.0625)
    fs = gh * gw
    written = torch.zeros(rows, dtype=torch.bool, device=DEV)
    for i, r0 in enumerate(kv_row):
        _report(f"cache k frame {i}", kc[r0:r0 + fs], k_ref[i * fs:(i + 1) * fs], atol=1e-6, rtol=1.6e-2, outlier_frac=1e-5, outlier_abs=0.0625)
        assert torch.equal(vc[r0:r0 + fs], v_ref[i * fs:(i + 1) * fs]), "v rows must be copied bit-exactly"
        written[r0:r0 + fs] = True
    assert (kc[~written] == 0).all() and (vc[~written] == 0).all(), "rows outside the write window were touched"
    # the fraction of bit-identical elements should be overwhelming (only fp32 reduction order differs)
    same = (q_out == q_ref).float().mean().item()
    print("bit-identical roped q fraction:", same)
    assert same > 0.9999


def test_time_embedding_pieces():
    ops = _ops()
    D = 1536
    t = torch.tensor([1000.0, 937.5, 625.0], dtype=torch.float64, device=DEV)
    emb = ops.sinusoid_embedding(t, 256)
    _report("sinusoid", emb, O.sinusoidal_embedding_1d(256, t).to(torch.bfloat16), atol=1e-6, rtol=8e-3)
    w0, b0 = _rand(D, 256, scale=0.02, seed=1), _rand(D, scale=0.02, seed=2)
    w1, b1 = _rand(6 * D, D, scale=0.02, seed=3), _rand(6 * D, scale=0.02, seed=4)
    h = ops.skinny_linear(emb, w0, b0, silu_out=True)
    _report("time0+silu", h, O.silu(O.linear(emb, w0, b0)), atol=1e-3, rtol=1e-2)
    e0 = ops.skinny_linear(h, w1, b1, silu_in=True)
    _report("silu+proj", e0, O.linear(O.silu(h), w1, b1), atol=1e-3, rtol=1e-2)
    m = ops.skinny_linear(_rand(7, D, seed=5), w1, b1)
    _report("skinny M=7", m, O.linear(_rand(7, D, seed=5), w1, b1), atol=1e-3, rtol=1e-2)


def test_modulation_add():
    ops = _ops()
    D, Fn = 1536, 3
    mod, e0, e = _rand(6, D, seed=1), _rand(Fn, 6, D, seed=2), _rand(Fn, D, seed=3)
    got = ops.modulation_add(mod, e0, 6 * D, D, Fn)
    assert torch.equal(got, mod.view(1, 6, D) + e0)
    hm = _rand(2, D, seed=4)
    got = ops.modulation_add(hm, e, D, 0, Fn)
    assert torch.equal(got, hm.view(1, 2, D) + e.view(Fn, 1, D))


def test_patchify_unpatchify_x0_add_noise():
    ops = _ops()
    cfg = O.WanConfig(dim=256, num_heads=2)
    Fn, C, H, W = 3, 16, 12, 20
    lat = _rand(Fn, C, H, W, seed=1)
    w = _rand(256, C, 1, 2, 2, scale=0.1, seed=2)
    b = _rand(256, scale=0.1, seed=3)
    a = ops.patchify(lat)
    tok = ops.linear(a, w.flatten(1), b)
    ref = O.patch_embed(cfg, {"patch_embedding.weight": w, "patch_embedding.bias": b}, lat.permute(1, 0, 2, 3))
    _report("patch embed", tok, ref, atol=1e-2, rtol=1e-2)
    # strided input view (the pipeline passes slices of [B, F, C, H, W])
    big = _rand(5, C, H, W, seed=4)
    assert torch.equal(ops.patchify(big[1:4]), ops.patchify(big[1:4].contiguous()))
    head = _rand(Fn * (H // 2) * (W // 2), 64, seed=5)
    sched = O.FlowMatchSchedule(5.0)
    tsteps = torch.tensor([1000.0, 833.3333, 625.0], device=DEV)
    sig = sched.sigmas.double().to(DEV)[sched.timestep_id(tsteps)]
    flow, x0 = ops.unpatchify_x0(head, (Fn, C, H, W), lat, sig)
    flow_ref = O.unpatchify(cfg, head, (Fn, H // 2, W // 2)).permute(1, 0, 2, 3)
    assert torch.equal(flow, flow_ref)
    assert torch.equal(x0, sched.flow_to_x0(flow_ref, lat, tsteps))
    noise = _rand(Fn, C, H, W, seed=6)
    sig32 = sched.sigmas.to(DEV)[sched.timestep_id(tsteps)]
    assert torch.equal(ops.add_noise(x0, noise, sig32), sched.add_noise(x0, noise, tsteps))
