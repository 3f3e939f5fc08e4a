#!/usr/bin/env python
"""Records tests/golden/fps_pipeline_{t2v,i2v}.pt by running the *unmodified reference* CausalFPSInferencePipeline
(MMPL_t2v/pipeline/casual_fps_inference.py:34-524, MMPL_i2v/pipeline/casual_fps_inference.py) on the CPU around
oracle/fake_fps_generator.FakeFPSGenerator:

    python oracle/make_golden_fps_pipeline.py t2v
    python oracle/make_golden_fps_pipeline.py i2v      (separate processes: both trees use the same module names)

What is pinned: the stage schedule, every generator call (branch, timesteps, frame positions, visibility list), the
pipeline's edits of `attention_vis_index`, CFG combine, the reference UniPC update, the re-noising of stage-boundary
frames (torch RNG order included), the anchor hand-off payload and the final latents - everything in the pipeline that is
not the backbone. Only runs in the build container (/root/reference); the fixtures travel."""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_shim  # noqa: E402
from oracle.fake_fps_generator import FakeFPSGenerator, digest  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
SAMPLING_STEPS = 3


def inputs(variant: str):
    g = torch.Generator().manual_seed(21)
    noise = torch.randn(1, 21, 16, 60, 104, generator=g).to(torch.bfloat16)
    first = torch.randn(1, 1, 16, 60, 104, generator=g).to(torch.bfloat16)     # i2v: image latent
    connect = torch.randn(1, 2, 16, 60, 104, generator=g).to(torch.bfloat16)   # i2v: "segment connect" frames
    return noise, first, connect


def main(variant: str):
    import importlib
    ref_shim.REF_ROOT = Path("/root/reference/MMPL_i2v" if variant == "i2v" else "/root/reference/MMPL_t2v")
    ref = ref_shim.load()
    ref_shim.load_unipc()
    # the DPM++ solver module is imported by the pipeline but not used by its unipc branch; it needs more of diffusers
    stub = types.ModuleType("wan.utils.fm_solvers")
    stub.FlowDPMSolverMultistepScheduler = stub.get_sampling_sigmas = stub.retrieve_timesteps = None
    sys.modules["wan.utils.fm_solvers"] = stub
    mod = importlib.import_module("pipeline.casual_fps_inference")

    sched = ref.scheduler.FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    sched.set_timesteps(1000, training=True)
    gen = FakeFPSGenerator(sched)

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            sign = -1.0 if text_prompts[0] == "__negative__" else 1.0
            return {"prompt_embeds": torch.full((1, 32, 64), sign, dtype=torch.bfloat16)}

    class VAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="__negative__",
                                 independent_first_frame=False, model_kwargs={})
    save_path = f"/tmp/fpsgold/anchors_{variant}.pt"
    torch.manual_seed(11)  # constructor: one torch.randint draw
    pipe = mod.CausalFPSInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=Text(), vae=VAE(),
                                          device_cond="cpu", device_uncond="cpu", save=save_path)
    pipe.sampling_steps = SAMPLING_STEPS
    pipe.num_transformer_blocks = 2

    # caches: the reference allocates 40 blocks x [B, 23400, 40, 128] x 4 tensors; the fake generator never touches K/V
    def small_caches():
        kv = lambda: [{"k": torch.zeros(1, 1, 1, 1), "v": torch.zeros(1, 1, 1, 1),
                       "global_end_index": torch.tensor([0]), "local_end_index": torch.tensor([0]),
                       "attention_vis_index": []} for _ in range(2)]
        cr = lambda: [{"k": torch.zeros(1, 1, 1, 1), "v": torch.zeros(1, 1, 1, 1), "is_init": False} for _ in range(2)]
        pipe.kv_cache_pos, pipe.kv_cache_neg = kv(), kv()
        pipe.crossattn_cache_pos, pipe.crossattn_cache_neg = cr(), cr()

    pipe._initialize_kv_cache = lambda **k: small_caches()
    pipe._initialize_crossattn_cache = lambda **k: None
    noise, first, connect = inputs(variant)
    runs = {}
    cases = [("plain", None)] if variant == "t2v" else [("image", first), ("connect", connect)]
    if variant == "t2v":
        cases.append(("extend", connect))  # stage 0 prefilled with two given frames (:405-439)
    for name, initial in cases:
        gen.calls = []
        pipe.kv_cache_pos = None
        torch.manual_seed(77)  # inference: torch.randn_like draws of the re-noising steps
        _, latents = pipe.inference(noise=noise.clone(), text_prompts=["p"], initial_latent=initial, return_latents=True)
        anchors = torch.load(save_path)
        runs[name] = dict(calls=gen.calls, latents_sub=latents[:, :, :, ::4, ::4].clone(), latents_sha=digest(latents),
                          anchors_sha=digest(anchors), anchors_shape=tuple(anchors.shape),
                          vis_end=sorted(pipe.kv_cache_pos[0]["attention_vis_index"]))
        print(variant, name, len(gen.calls), "calls", runs[name]["latents_sha"], runs[name]["anchors_shape"])
    fix = dict(kind="fps_pipeline", variant=variant, sampling_steps=SAMPLING_STEPS, ctor_seed=11, run_seed=77, input_seed=21,
               ddmp_timestep=pipe.ddmp_timestep.clone(), runs=runs)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save(fix, GOLDEN / f"fps_pipeline_{variant}.pt")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "t2v")
