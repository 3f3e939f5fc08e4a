#!/usr/bin/env python
"""Records tests/golden/vae_small.pt from the *unmodified reference* VAE (MMPL_t2v/wan/modules/vae.py: WanVAE_ with the
`_video_vae` configuration, dim 96, z 16) and the reference wrapper arithmetic (utils/wan_wrapper.py:74-113) on the CPU:

    python oracle/make_golden_vae.py

The checkpoint is absent, so the weights are oracle.vae_oracle.make_weights(seed 0) loaded through the reference's own
`load_state_dict` (strict: proves the name/shape inventory). Recorded at 32x48 pixels (latent 4x6), fp32 and bf16:
encode of 9 frames, decode of 4 latent frames, and the driver's segment-connect transform
(Wan_fps_inference_parallel_4gpu_20s.py:191-205) on 21 latent / 81 pixel frames. Only outputs and the seeded inputs'
recipe are stored (the weights are regenerated from the seed). Runs only in the build container."""
import importlib.util
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import vae_oracle as V  # noqa: E402

REF = Path("/root/reference/MMPL_t2v")
GOLDEN = ROOT / "tests" / "golden" / "vae_small.pt"


def load_reference_vae_module():
    """wan/modules/vae.py imported by path (its package __init__ pulls absent dependencies; the file itself needs
    torch + einops only)."""
    spec = importlib.util.spec_from_file_location("mmpl_ref_vae", REF / "wan" / "modules" / "vae.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs(dtype):
    g = torch.Generator().manual_seed(7)
    pixels = (torch.rand(1, 3, 9, 32, 48, generator=g) * 2 - 1).to(dtype)
    latents = torch.randn(1, 4, 16, 4, 6, generator=g).to(dtype)          # [B,T,z,h,w] as the wrapper takes them
    anchors = torch.randn(1, 8, 16, 4, 6, generator=g).to(torch.bfloat16)  # t2v anchor payload layout
    return pixels, latents, anchors


class RefWrapper:
    """WanVAEWrapper's two methods (utils/wan_wrapper.py:74-113) around a given WanVAE_ — the class itself cannot be
    constructed without the checkpoint file."""

    def __init__(self, model):
        self.model = model
        self.mean = torch.tensor(V.LATENT_MEAN, dtype=torch.float32)
        self.std = torch.tensor(V.LATENT_STD, dtype=torch.float32)


def main():
    ref = load_reference_vae_module()
    sys.path.insert(0, str(REF))
    # utils/wan_wrapper.py imports the whole model zoo; bind its two VAE methods to our stand-in instead of importing it
    src = (REF / "utils" / "wan_wrapper.py").read_text()
    ns: dict = {"torch": torch}
    start = src.index("    def encode_to_latent")
    end = src.index("class WanDiffusionWrapper")
    exec("class _M:\n" + src[start:end], ns)  # the reference's own method bodies, unmodified
    RefWrapper.encode_to_latent = ns["_M"].encode_to_latent
    RefWrapper.decode_to_pixel = ns["_M"].decode_to_pixel

    cfg = V.VaeConfig()
    out = {"recipe": "oracle.vae_oracle.make_weights(VaeConfig(), seed=0); inputs(): see oracle/make_golden_vae.py"}
    for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        W = V.make_weights(cfg, seed=0, dtype=dtype)
        with torch.device("meta"):
            model = ref.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                                temperal_downsample=[False, True, True], dropout=0.0)
        missing = model.load_state_dict(W, assign=True, strict=True)
        model = model.eval().requires_grad_(False)
        wrap = RefWrapper(model)
        pixels, latents, anchors = inputs(dtype)
        with torch.no_grad():
            enc = wrap.encode_to_latent(pixels)
            dec = wrap.decode_to_pixel(latents)
            rec = {"encode": enc.clone(), "decode": dec.clone(), "state_dict_keys": sorted(W)}
            if dtype == torch.bfloat16:
                # the driver's hand-off transform, lines 191-205, verbatim semantics on CPU
                lat = anchors.to(torch.bfloat16)
                masked = torch.zeros(1, 21, 16, 4, 6).to(torch.bfloat16)
                masked[:, 0:1] = lat[:, 0:1]
                masked[:, 1:2] = lat[:, -2:-1]
                masked[:, 2:4] = lat[:, -2:]
                vid = wrap.decode_to_pixel(masked).to(torch.bfloat16)
                vid = (vid * 0.5 + 0.5).clamp(0, 1).to(torch.bfloat16)
                test = torch.zeros_like(vid).to(torch.bfloat16)
                test[:, 0:5] = vid[:, 8:13]
                test = (test * 2.0 - 1.0).permute(0, 2, 1, 3, 4)   # rearrange "b t c h w -> b c t h w"
                rec["connect"] = wrap.encode_to_latent(test)[:, :2].to(torch.bfloat16).clone()
                rec["connect_vid_8_13"] = vid[:, 8:13].clone()
        out[name] = rec
        print(name, "encode", tuple(enc.shape), "decode", tuple(dec.shape), float(enc.abs().mean()), float(dec.abs().mean()))
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, GOLDEN.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
