"""CPU: the VAE oracle (the segment connect) against goldens recorded from the unmodified reference
VAE by oracle/make_golden_vae.py (wan/modules/vae.py, utils/wan_wrapper.py:74-113,
Wan_fps_inference_parallel_4gpu_20s.py:191-205). The oracle issues the same torch operators in the same order as the
reference, so on the recording machine every comparison is bit-exact; the asserted bounds leave room for a different
oneDNN kernel choice on another host (fp32 summation order; one bf16 ulp of the output range)."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import vae_oracle as V  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "vae_small.pt"
TOL = {"fp32": 2e-5, "bf16": 2 ** -6}
DT = {"fp32": torch.float32, "bf16": torch.bfloat16}


def inputs(dtype):  # same recipe as oracle/make_golden_vae.py:inputs
    g = torch.Generator().manual_seed(7)
    pixels = (torch.rand(1, 3, 9, 32, 48, generator=g) * 2 - 1).to(dtype)
    latents = torch.randn(1, 4, 16, 4, 6, generator=g).to(dtype)
    anchors = torch.randn(1, 8, 16, 4, 6, generator=g).to(torch.bfloat16)
    return pixels, latents, anchors


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN)


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def case(request, golden):
    name = request.param
    cfg = V.VaeConfig()
    return name, cfg, V.make_weights(cfg, 0, DT[name]), inputs(DT[name]), golden[name]


def test_state_dict_inventory(case):
    """Names and shapes are the reference's: the golden script loads these weights into WanVAE_ with strict=True."""
    name, cfg, W, _, g = case
    assert sorted(W) == g["state_dict_keys"]
    assert W["decoder.upsamples.3.time_conv.weight"].shape == (768, 384, 3, 1, 1)   # vae.py:80-81
    assert W["encoder.downsamples.5.time_conv.weight"].shape == (192, 192, 3, 1, 1)  # vae.py:92-93
    assert V.n_slots(V.decoder_program(cfg)) == 33 and V.n_slots(V.encoder_program(cfg)) == 26  # count_conv3d of the reference model: 33 / 26


def test_encode_matches_reference(case):
    name, cfg, W, (pixels, _, _), g = case
    out = V.encode_to_latent(W, cfg, pixels)
    assert out.shape == g["encode"].shape == (1, 3, 16, 4, 6) and out.dtype == torch.float32
    assert float((out - g["encode"]).abs().max()) <= TOL[name]


def test_decode_matches_reference(case):
    name, cfg, W, (_, latents, _), g = case
    out = V.decode_to_pixel(W, cfg, latents)
    assert out.shape == g["decode"].shape == (1, 13, 3, 32, 48)
    assert float((out - g["decode"]).abs().max()) <= TOL[name]
    assert float(out.abs().max()) <= 1.0  # wan_wrapper.py:108 clamps


def test_whole_sequence_form_equals_streaming(case):
    """One pass per layer over all frames == the reference's chunked schedule with carried frames, including the
    'Rep' quirk of the temporal up-sampler and the pass-through of frame 0 in the temporal down-sampler."""
    name, cfg, W, (pixels, latents, _), g = case
    enc = V.encode_whole(W, cfg, pixels).float().permute(0, 2, 1, 3, 4)
    dec = V.decode_whole(W, cfg, latents.permute(0, 2, 1, 3, 4)).float().clamp(-1, 1).permute(0, 2, 1, 3, 4)
    assert float((enc - g["encode"]).abs().max()) <= TOL[name]
    assert float((dec - g["decode"]).abs().max()) <= TOL[name]


def test_truncation_is_causal(case):
    """Frames already produced never change when more input follows (what makes the reduced segment connect exact)."""
    name, cfg, W, (pixels, latents, _), g = case
    d2 = V.decode_to_pixel(W, cfg, latents, max_frames=2)
    assert d2.shape[1] == 5 and float((d2 - g["decode"][:, :5]).abs().max()) <= TOL[name]
    e2 = V.encode_to_latent(W, cfg, pixels, max_chunks=2)
    assert e2.shape[1] == 2 and float((e2 - g["encode"][:, :2]).abs().max()) <= TOL[name]


def test_segment_connect_matches_reference_driver(golden):
    """The hand-off transform on 21 latent / 81 pixel frames, and the same result from 4 decoder steps + 2 encoder chunks."""
    cfg = V.VaeConfig()
    W = V.make_weights(cfg, 0, torch.bfloat16)
    anchors = inputs(torch.bfloat16)[2]
    want = golden["bf16"]["connect"]
    full = V.segment_connect(W, cfg, anchors)
    reduced = V.segment_connect_causal(W, cfg, anchors)
    assert full.shape == want.shape == (1, 2, 16, 4, 6) and full.dtype == torch.bfloat16
    assert float((full.float() - want.float()).abs().max()) <= TOL["bf16"]
    assert torch.equal(reduced, full)  # same operators on the same data: exact on any host
