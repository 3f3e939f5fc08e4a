#!/usr/bin/env python
"""Records tests/golden/contig_pipelines.pt by running the *unmodified reference* contiguous-cache pipelines
(MMPL_t2v/pipeline/causal_inference.py:9-312 - the few-step CausalInferencePipeline of BASELINE configs 1-2 - and
pipeline/causal_diffusion_inference.py:11-378 - the 50-step CFG CausalDiffusionInferencePipeline) on the CPU around
oracle/fake_fps_generator.FakeFPSGenerator at the full 60x104 latent size, including the branches the model-level
goldens (causal_tiny / causal_cfg1) do not reach: image-to-video and video-extension prefill (`initial_latent`),
`independent_first_frame`, unwarped step lists, `context_noise`.

    python oracle/make_golden_contig_pipelines.py

Pinned: chunk schedule, every generator call (branch, per-frame timesteps, current_start, cache indices before / after,
input latents), re-noising between steps (torch RNG order), CFG combine + reference UniPC, the final latents. Only runs
in the build container (/root/reference); the fixture travels."""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_shim  # noqa: E402
from oracle.fake_fps_generator import FakeFPSGenerator, digest  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

# name -> (pipeline, args overrides, noise frames, initial_latent frames)
CASES = {
    "causal_plain":        ("causal", dict(), 6, 0),
    "causal_image":        ("causal", dict(), 6, 3),                       # prefill one block, then generate two
    "causal_first_frame":  ("causal", dict(independent_first_frame=True), 7, 0),
    "causal_first_image":  ("causal", dict(independent_first_frame=True), 6, 1),
    "causal_unwarped_ctx": ("causal", dict(warp_denoising_step=False, context_noise=25, num_frame_per_block=2), 4, 0),
    "diffusion_plain":     ("diffusion", dict(), 6, 0),
    "diffusion_extend":    ("diffusion", dict(), 3, 3),
    "diffusion_first_img": ("diffusion", dict(independent_first_frame=True), 3, 1),
}


def case_inputs(noise_frames: int, init_frames: int):
    g = torch.Generator().manual_seed(31)
    noise = torch.randn(1, noise_frames, 16, 60, 104, generator=g).to(torch.bfloat16)
    initial = torch.randn(1, init_frames, 16, 60, 104, generator=g).to(torch.bfloat16) if init_frames else None
    return noise, initial


def base_args(kind: str, over: dict):
    if kind == "causal":
        a = dict(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True, independent_first_frame=False,
                 context_noise=0, num_frame_per_block=3, model_kwargs={})
    else:
        a = dict(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="__negative__",
                 independent_first_frame=False, num_frame_per_block=3, model_kwargs={})
    a.update(over)
    return types.SimpleNamespace(**a)


def small_caches(n_blocks=2):
    return [{"k": torch.zeros(1, 1, 1, 1), "v": torch.zeros(1, 1, 1, 1), "global_end_index": torch.tensor([0]),
             "local_end_index": torch.tensor([0])} for _ in range(n_blocks)], \
           [{"k": torch.zeros(1, 1, 1, 1), "v": torch.zeros(1, 1, 1, 1), "is_init": False} for _ in range(n_blocks)]


def main():
    import importlib
    ref = ref_shim.load()
    ref_shim.load_unipc()
    stub = types.ModuleType("wan.utils.fm_solvers")  # DPM++ module: imported, not used by the unipc branch
    stub.FlowDPMSolverMultistepScheduler = stub.get_sampling_sigmas = stub.retrieve_timesteps = None
    sys.modules["wan.utils.fm_solvers"] = stub
    diffusion_mod = importlib.import_module("pipeline.causal_diffusion_inference")

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            sign = -1.0 if text_prompts[0] == "__negative__" else 1.0
            return {"prompt_embeds": torch.full((1, 32, 64), sign, dtype=torch.bfloat16)}

        def to(self, *a, **k):
            return self

    class VAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

        def to(self, *a, **k):
            return self

    runs = {}
    for name, (kind, over, nf, ni) in CASES.items():
        sched = ref.scheduler.FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
        sched.set_timesteps(1000, training=True)
        gen = FakeFPSGenerator(sched)
        args = base_args(kind, over)
        if kind == "causal":
            pipe = ref.causal_inference.CausalInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=Text(), vae=VAE())
            pipe.num_transformer_blocks = 2

            def init_kv(**k):
                pipe.kv_cache1, _ = small_caches()

            def init_cross(**k):
                _, pipe.crossattn_cache = small_caches()
        else:
            pipe = diffusion_mod.CausalDiffusionInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=Text(), vae=VAE())
            pipe.sampling_steps = 3
            pipe.num_transformer_blocks = 2

            def init_kv(**k):
                pipe.kv_cache_pos, _ = small_caches()
                pipe.kv_cache_neg, _ = small_caches()

            def init_cross(**k):
                _, pipe.crossattn_cache_pos = small_caches()
                _, pipe.crossattn_cache_neg = small_caches()
        pipe._initialize_kv_cache = init_kv
        pipe._initialize_crossattn_cache = init_cross
        noise, initial = case_inputs(nf, ni)
        torch.manual_seed(55)
        _, latents = pipe.inference(noise=noise.clone(), text_prompts=["p"], initial_latent=initial, return_latents=True)
        kv = pipe.kv_cache1 if kind == "causal" else pipe.kv_cache_pos
        runs[name] = dict(kind=kind, over=over, noise_frames=nf, init_frames=ni, calls=gen.calls,
                          latents_sub=latents[:, :, :, ::4, ::4].clone(), latents_sha=digest(latents), shape=tuple(latents.shape),
                          end=(int(kv[0]["global_end_index"]), int(kv[0]["local_end_index"])))
        print(name, len(gen.calls), "calls", runs[name]["shape"], runs[name]["latents_sha"], runs[name]["end"])
    torch.save(dict(kind="contig_pipelines", run_seed=55, input_seed=31, runs=runs), GOLDEN / "contig_pipelines.pt")


if __name__ == "__main__":
    main()
