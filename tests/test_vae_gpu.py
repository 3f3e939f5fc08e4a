"""GPU parity of the VAE segment-connect kernels against the VAE oracle, through the C ABI:
mmpl_conv3d_cl (tap-GEMM causal convolution on a zero-haloed channels-last grid) vs oracle.vae_oracle.causal_conv3d
(the restatement of CausalConv3d.forward, wan/modules/vae.py:16-36, pinned bit-exact against the reference VAE on CPU).
Tolerance: bf16 output of an fp32-accumulated contraction, |got - ref| <= 2e-2 + 2e-2 |ref| (as for the Linear GEMMs)."""
import pytest
import torch

from oracle import vae_oracle as V

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _check(name, got, ref):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bad = int((err > 2e-2 + 2e-2 * ref.abs()).sum())
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    print(f"{name}: max_abs={err.max().item():.4g} cos={cos:.6f} bad={bad}/{err.numel()}")
    assert torch.isfinite(got).all() and bad == 0 and cos > 0.9999, f"{name}: {bad} elements out of tolerance, cos {cos}"


CASES = [
    # cin, cout, kernel, T, H, W, history frames carried in front, residual
    (16, 384, (3, 3, 3), 3, 6, 10, 0, False),     # decoder.conv1 (z -> 384)
    (96, 96, (3, 3, 3), 2, 30, 52, 0, True),      # residual block at the top level, x + h in the epilogue
    (384, 384, (3, 3, 3), 1, 15, 26, 2, True),    # one streamed frame with two carried frames (feat_cache)
    (192, 384, (1, 1, 1), 2, 9, 7, 0, False),     # shortcut (1x1x1)
    (384, 768, (3, 1, 1), 2, 8, 12, 1, False),    # temporal up-sampling conv, one carried frame
    (192, 96, (1, 3, 3), 3, 16, 24, 0, False),    # per-frame Conv2d 3x3 of the up-sampler
    (96, 3, (3, 3, 3), 2, 16, 24, 0, False),      # decoder head: 3 output channels in an 8-channel layout
    (3, 96, (3, 3, 3), 2, 16, 24, 0, False),      # encoder conv1: 3 input channels in an 8-channel layout
]


@pytest.mark.parametrize("cin,cout,kernel,T,H,W,hist,with_res", CASES)
def test_conv3d_tap_gemm(cin, cout, kernel, T, H, W, hist, with_res):
    from mmpl_b200 import ops
    kt, kh, kw = kernel
    fan = cin * kt * kh * kw
    x = _rand(cin, hist + T, H, W, seed=1)                    # [C, frames, H, W]; the first `hist` frames are history
    w = _rand(cout, cin, kt, kh, kw, scale=fan ** -0.5, seed=2)
    b = _rand(cout, scale=0.1, seed=3)
    res = _rand(cout, T, H, W, seed=4) if with_res else None

    # oracle: history frames replace that many zero frames of the causal padding (vae.py:28-33)
    xf = x.float().unsqueeze(0).cpu()     # reference on the CPU: the oracle as pinned, no cuDNN algorithm choice involved
    ref = V.causal_conv3d(xf[:, :, hist:], w.float().cpu(), b.float().cpu(), xf[:, :, :hist] if hist else None)
    ref = ref.to(torch.bfloat16).to(DEV)
    if with_res:
        ref = (ref + res.unsqueeze(0)).to(torch.bfloat16)    # ResidualBlock: x + h in bf16 (vae.py:213)
    ref = ref[0]

    lead = 2
    grid = ops.to_haloed(x[:, hist:], lead=lead)
    if hist:
        grid[lead - hist:lead] = ops.to_haloed(x[:, :hist], lead=0)
    res_grid = ops.to_haloed(res, lead=lead) if with_res else None
    out = ops.conv3d_causal_cl(grid, ops.pack_conv_weight(w), b, kernel, lead=lead, residual=res_grid)
    torch.cuda.synchronize()
    _check(f"conv {cin}->{cout} k{kernel} T{T} {H}x{W} hist{hist}", ops.from_haloed(out, lead, cout), ref)
    # the halo and the leading frames are never written
    assert float(out[:lead].abs().max()) == 0 and float(out[:, 0].abs().max()) == 0 and float(out[:, -1].abs().max()) == 0
    assert float(out[:, :, 0].abs().max()) == 0 and float(out[:, :, -1].abs().max()) == 0
    assert float(out[..., cout:].abs().max() if out.shape[-1] > cout else 0.0) == 0


def test_conv3d_chain_matches_whole_sequence():
    """Two stacked convolutions on the haloed grid without leaving the layout (the halo written by nobody stays zero, so
    the second convolution sees correct spatial padding): equals the oracle's whole-sequence causal convolutions."""
    from mmpl_b200 import ops
    x = _rand(96, 4, 12, 20, seed=5)
    w1, b1 = _rand(192, 96, 3, 3, 3, scale=(96 * 27) ** -0.5, seed=6), _rand(192, scale=0.1, seed=7)
    w2, b2 = _rand(96, 192, 3, 3, 3, scale=(192 * 27) ** -0.5, seed=8), _rand(96, scale=0.1, seed=9)
    h = V.causal_conv3d(x.float().unsqueeze(0).cpu(), w1.float().cpu(), b1.float().cpu()).to(torch.bfloat16)
    ref = V.causal_conv3d(h.float(), w2.float().cpu(), b2.float().cpu()).to(torch.bfloat16)[0].to(DEV)
    g = ops.to_haloed(x)
    g = ops.conv3d_causal_cl(g, ops.pack_conv_weight(w1), b1, (3, 3, 3))
    g = ops.conv3d_causal_cl(g, ops.pack_conv_weight(w2), b2, (3, 3, 3))
    torch.cuda.synchronize()
    _check("conv chain 96->192->96", ops.from_haloed(g), ref)
