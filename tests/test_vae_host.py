"""CPU: the host orchestration of the VAE segment connect (mmpl_b200/vae.py) with every kernel replaced by a torch
emulation of its C-ABI contract (include/mmpl_b200.h: same layouts, same bf16 rounding points), against the goldens
recorded from the unmodified reference VAE. This pins everything the host side decides — layer order and state-dict
names, the haloed-grid bookkeeping, the whole-sequence handling of the temporal up/down-samplers ('Rep' quirk, frame-0
pass-through), the stride-2 picks, the attention plumbing and the causally reduced connect — independently of the GPU;
the kernels themselves are checked against the oracle in tests/test_vae_gpu.py. The product never runs these emulations."""
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import vae_oracle as V  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "vae_small.pt"
BF = torch.bfloat16


def _r(x):  # one bf16 rounding
    return x.to(BF).float()


# ---------------------------------------------------------------------------------- kernel contracts, emulated
def emu_conv3d_causal_cl(grid, w_packed, bias, kernel, history=0, residual=None, out=None):
    """mmpl_conv3d_cl (include/mmpl_b200.h): every grid position is a GEMM row; a K span is a row shift of the grid read as
    one contiguous run (Cin elements for KW = 1, the three dw taps = 3*Cin elements starting one row earlier for KW = 3),
    zeros outside the tensor; the causal padding frames that are not stored (`history` < KT-1) are rows with negative
    coordinates; interior positions get the result, halo positions zeros."""
    kt, kh, kw = kernel
    frames, hp, wp, cin = grid.shape
    t, cout = frames - history, w_packed.shape[0]
    flat = grid.reshape(-1).float()
    rows_in, m = frames * hp * wp, t * hp * wp
    span = 3 * cin if kw == 3 else cin
    acc = torch.zeros(m, cout)
    r = torch.arange(m)
    i = 0
    for dt in range(kt):
        for dh in range(kh):
            frame_row = ((dt - (kt - 1) + history) * hp + (dh - kh // 2)) * wp
            for dw in ([None] if kw == 3 else range(kw)):
                first = r + frame_row + (-1 if kw == 3 else dw - kw // 2)          # first grid row of the span
                limit = rows_in - 2 if kw == 3 else rows_in                          # rows the tensor map covers
                ok = (first >= 0) & (first < limit)
                a = torch.zeros(m, span)
                gather = first[ok, None] * cin + torch.arange(span)[None, :]
                a[ok] = flat[gather]
                acc += a @ w_packed[:, i, :span].float().T
                i += 1
    if bias is not None:
        b = bias.float().reshape(-1)
        acc[:, :b.numel()] += b
    y = _r(acc)
    if residual is not None:
        y = _r(residual.reshape(-1, cout).float() + y)
    interior = ((r % wp) >= 1) & ((r % wp) < wp - 1) & (((r // wp) % hp) >= 1) & (((r // wp) % hp) < hp - 1)
    y[~interior] = 0
    return y.to(BF).reshape(t, hp, wp, cout)


def emu_vae_norm_act(grid, gamma, silu=True, out=None):
    x = grid.float()
    c = x.shape[-1]
    g = gamma.reshape(-1).float()
    if g.numel() != c:
        g = torch.cat([g, torch.ones(c - g.numel())])
    n = _r(x.pow(2).sum(-1, keepdim=True).sqrt()).clamp_min(1e-12)
    y = _r(_r(_r(x / n) * torch.tensor(float(c)).sqrt()) * g)
    if silu:
        y = y / (1 + torch.exp(-y))
    return y.to(BF)


def emu_vae_upsample2x(grid):
    frames, hp, wp, c = grid.shape
    out = torch.zeros(frames, 2 * (hp - 2) + 2, 2 * (wp - 2) + 2, c, dtype=BF)
    out[:, 1:-1, 1:-1] = grid[:, 1:-1, 1:-1].repeat_interleave(2, 1).repeat_interleave(2, 2)
    return out


def emu_vae_pick_odd(grid):
    frames, hp, wp, c = grid.shape
    h, w = (hp - 2) // 2, (wp - 2) // 2
    out = torch.zeros(frames, h + 2, w + 2, c, dtype=BF)
    out[:, 1:-1, 1:-1] = grid[:, 2:2 * h + 1:2, 2:2 * w + 1:2]
    return out


def emu_softmax_rows(s, scale):
    return torch.softmax(s.float() * scale, dim=-1).to(BF)


def emu_linear(x, weight, bias=None, *, epilogue=0, residual=None, **kw):
    y = x.float() @ weight.float().T
    if bias is not None:
        y = y + bias.float()
    y = _r(y)
    if epilogue == 3:   # EPI_BIAS_RES
        y = _r(residual.float() + y)
    else:
        assert epilogue == 0
    return y.to(BF)


@pytest.fixture()
def vae(monkeypatch):
    from mmpl_b200 import ops
    from mmpl_b200.vae import WanVAEWrapper
    for name, fn in (("conv3d_causal_cl", emu_conv3d_causal_cl), ("vae_norm_act", emu_vae_norm_act),
                     ("vae_upsample2x", emu_vae_upsample2x), ("vae_pick_odd", emu_vae_pick_odd),
                     ("softmax_rows", emu_softmax_rows), ("linear", emu_linear)):
        monkeypatch.setattr(ops, name, fn)
    m = WanVAEWrapper()
    m.load_vae_state_dict(V.make_weights(V.VaeConfig(), 0, BF), device="cpu")
    return m


def inputs():  # oracle/make_golden_vae.py:inputs(bf16)
    g = torch.Generator().manual_seed(7)
    pixels = (torch.rand(1, 3, 9, 32, 48, generator=g) * 2 - 1).to(BF)
    latents = torch.randn(1, 4, 16, 4, 6, generator=g).to(BF)
    anchors = torch.randn(1, 8, 16, 4, 6, generator=g).to(BF)
    return pixels, latents, anchors


def _close(name, got, want, atol):
    got, want = got.float(), want.float()
    err = float((got - want).abs().max())
    cos = float(F.cosine_similarity(got.flatten(), want.flatten(), dim=0))
    print(f"{name}: max_abs={err:.4g} cos={cos:.6f}")
    assert got.shape == want.shape and err <= atol and cos >= 0.9995, f"{name}: max_abs {err}, cos {cos}"


def test_decode_matches_reference_vae(vae):
    g = torch.load(GOLDEN)["bf16"]
    _close("decode", vae.decode_to_pixel(inputs()[1]), g["decode"], atol=0.06)


def test_encode_matches_reference_vae(vae):
    g = torch.load(GOLDEN)["bf16"]
    _close("encode", vae.encode_to_latent(inputs()[0]), g["encode"], atol=0.06)


def test_segment_connect_matches_reference_driver(vae):
    g = torch.load(GOLDEN)["bf16"]
    out = vae.segment_connect(inputs()[2])
    assert out.dtype == BF
    _close("connect", out, g["connect"], atol=0.08)


def test_no_cpu_path():
    """Without the emulations the wrapper goes to the sm_100a library and refuses CPU tensors."""
    from mmpl_b200.vae import WanVAEWrapper
    m = WanVAEWrapper()
    m.load_vae_state_dict({k: v for k, v in V.make_weights(V.VaeConfig(), 0, BF).items() if k.startswith(("conv2", "decoder.conv1"))},
                          device="cpu")
    with pytest.raises((ValueError, RuntimeError)):
        m.decode_to_pixel(inputs()[1])


def test_random_init_has_the_reference_inventory():
    """init_random_weights (the synthetic stand-in for Wan2.1_VAE.pth) produces exactly the names and shapes the
    reference's WanVAE_ loads with strict=True (checked against the reference by oracle/make_golden_vae.py)."""
    from mmpl_b200.vae import WanVAEWrapper
    sd = WanVAEWrapper().init_random_weights(seed=3, device="cpu")
    want = V.make_weights(V.VaeConfig(), 0, BF)
    assert sorted(sd) == sorted(want) == torch.load(GOLDEN)["bf16"]["state_dict_keys"]
    assert all(sd[k].shape == want[k].shape for k in want)


def test_connect_hook_in_the_segment_runner(vae):
    """SegmentParallelRunner(connect=vae_segment_connect(vae)): the next segment starts from the VAE-connected anchors
    (Wan_fps_inference_parallel_4gpu_20s.py:191-211), same result as applying the oracle's transform to them."""
    from mmpl_b200.segment_parallel import AnchorChannel, SegmentParallelRunner, vae_segment_connect
    seen = {}

    class Pipe:
        anchor_sink = None

        def inference(self, noise, text_prompts, initial_latent=None, return_latents=True):
            seen[len(seen)] = initial_latent
            out = noise.clone()
            self.anchor_sink(torch.cat([out[:, :1], out[:, [2, 3, 10, 11, 12, 19, 20]]], dim=1))
            return out, out

    g = torch.Generator().manual_seed(5)
    noise = [torch.randn(1, 21, 16, 4, 6, generator=g).to(BF) for _ in range(2)]
    runner = SegmentParallelRunner(Pipe(), AnchorChannel(), anchor_shape=(1, 8, 16, 4, 6), connect=vae_segment_connect(vae))
    runner.run(lambda seg: noise[seg], ["p"], 2)
    assert seen[0] is None and seen[1].shape == (1, 2, 16, 4, 6) and seen[1].dtype == BF
    anchors = torch.cat([noise[0][:, :1], noise[0][:, [2, 3, 10, 11, 12, 19, 20]]], dim=1)
    want = V.segment_connect_causal(V.make_weights(V.VaeConfig(), 0, BF), V.VaeConfig(), anchors)
    _close("runner connect", seen[1], want, atol=0.08)


def test_i2v_image_encode_and_three_anchor_connect(vae):
    """MMPL_i2v: the conditioning image is one pixel frame through encode_to_latent
    (MMPL_i2v/Wan_fps_inference_parallel_4gpu_20s.py:193) and the hand-off payload has 3 anchors (frames 0, 19, 20;
    lines 212-226: the same transform). Checked against the oracle's streaming restatement."""
    W, cfg = V.make_weights(V.VaeConfig(), 0, BF), V.VaeConfig()
    g = torch.Generator().manual_seed(9)
    image = (torch.rand(1, 3, 1, 32, 48, generator=g) * 2 - 1).to(BF)
    _close("image encode", vae.encode_to_latent(image), V.encode_to_latent(W, cfg, image), atol=0.06)
    anchors = torch.randn(1, 3, 16, 4, 6, generator=g).to(BF)
    _close("i2v connect", vae.segment_connect(anchors), V.segment_connect_causal(W, cfg, anchors), atol=0.08)


def test_smoke_vae_leg_runs(vae):
    """__graft_entry__._smoke_vae_connect (the VAE leg of smoke()) on CPU with the kernel emulations: the smoke code itself."""
    import __graft_entry__ as ge
    msg = ge._smoke_vae_connect("cpu")
    assert msg.startswith("vae connect max_abs=")


@pytest.mark.parametrize("T", [1, 2, 6])
def test_decode_any_length(vae, T):
    """The whole-sequence decoder for other frame counts (1 = a single image, 6 -> 21 pixel frames: both temporal
    up-samplers see several frames) against the oracle's streaming restatement of WanVAE_.decode."""
    W, cfg = V.make_weights(V.VaeConfig(), 0, BF), V.VaeConfig()
    lat = torch.randn(1, T, 16, 4, 6, generator=torch.Generator().manual_seed(20 + T)).to(BF)
    got = vae.decode_to_pixel(lat)
    assert got.shape == (1, 1 + 4 * (T - 1), 3, 32, 48)
    _close(f"decode T={T}", got, V.decode_to_pixel(W, cfg, lat), atol=0.06)


def test_encode_thirteen_frames(vae):
    """13 pixel frames -> 4 latents: both temporal down-samplers take more than one stride-2 window."""
    W, cfg = V.make_weights(V.VaeConfig(), 0, BF), V.VaeConfig()
    px = (torch.rand(1, 3, 13, 32, 48, generator=torch.Generator().manual_seed(31)) * 2 - 1).to(BF)
    got = vae.encode_to_latent(px)
    assert got.shape == (1, 4, 16, 4, 6)
    _close("encode 13 frames", got, V.encode_to_latent(W, cfg, px), atol=0.06)
