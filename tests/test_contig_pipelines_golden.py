"""CausalInferencePipeline / CausalDiffusionInferencePipeline mirrors against the *reference pipelines themselves*
(SURVEY.md §8 rows a1, f-3): tests/golden/contig_pipelines.pt was recorded by running the unmodified reference
pipeline/causal_inference.py and pipeline/causal_diffusion_inference.py around oracle/fake_fps_generator.FakeFPSGenerator
(oracle/make_golden_contig_pipelines.py) at the full 60x104 latent size. The mirrors run around the same fake generator on
the CPU and must reproduce bit for bit every generator call (branch, per-frame timesteps, current_start, cache indices
before/after, input latents) and the final latents - including the image-to-video / video-extension prefill,
independent_first_frame, unwarped step lists and context_noise branches."""
import types
from pathlib import Path

import pytest
import torch

from mmpl_b200.pipeline import CausalDiffusionInferencePipeline, CausalInferencePipeline
from _cpu_ops import cpu_scheduler, eager_unipc_factory
from oracle.fake_fps_generator import FakeFPSGenerator, digest

GOLDEN = Path(__file__).parent / "golden"
FIX = torch.load(GOLDEN / "contig_pipelines.pt", weights_only=False)


def small_caches(n_blocks=2):
    kv = [{"k": torch.zeros(1, 1, 1, 1), "v": torch.zeros(1, 1, 1, 1), "global_end_index": torch.tensor([0]),
           "local_end_index": torch.tensor([0])} for _ in range(n_blocks)]
    cross = [{"k": torch.zeros(1, 1, 1, 1), "v": torch.zeros(1, 1, 1, 1), "is_init": False} for _ in range(n_blocks)]
    return kv, cross


@pytest.mark.parametrize("name", sorted(FIX["runs"]))
def test_mirror_reproduces_the_reference_pipeline(name):
    ref = FIX["runs"][name]
    gen = FakeFPSGenerator(cpu_scheduler())
    text = lambda text_prompts: {"prompt_embeds": torch.full((1, 32, 64), -1.0 if text_prompts[0] == "__negative__" else 1.0,
                                                             dtype=torch.bfloat16)}
    vae = types.SimpleNamespace(decode_to_pixel=lambda latents, use_cache=False: latents)
    if ref["kind"] == "causal":
        a = dict(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True, independent_first_frame=False,
                 context_noise=0, num_frame_per_block=3, model_kwargs={})
        a.update(ref["over"])
        pipe = CausalInferencePipeline(types.SimpleNamespace(**a), torch.device("cpu"), generator=gen, text_encoder=text, vae=vae)

        def init_kv(**k):
            pipe.kv_cache1, _ = small_caches()

        def init_cross(**k):
            _, pipe.crossattn_cache = small_caches()
    else:
        a = dict(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="__negative__",
                 independent_first_frame=False, num_frame_per_block=3, model_kwargs={}, sampling_steps=3)
        a.update(ref["over"])
        pipe = CausalDiffusionInferencePipeline(types.SimpleNamespace(**a), torch.device("cpu"), generator=gen, text_encoder=text, vae=vae)
        pipe.unipc_stepper = eager_unipc_factory(pipe)   # CPU: the fused UniPC kernel's place is taken by the oracle's eager operators

        def init_kv(**k):
            pipe.kv_cache_pos, _ = small_caches()
            pipe.kv_cache_neg, _ = small_caches()

        def init_cross(**k):
            _, pipe.crossattn_cache_pos = small_caches()
            _, pipe.crossattn_cache_neg = small_caches()
    pipe._initialize_kv_cache = init_kv
    pipe._initialize_crossattn_cache = init_cross
    g = torch.Generator().manual_seed(FIX["input_seed"])
    noise = torch.randn(1, ref["noise_frames"], 16, 60, 104, generator=g).to(torch.bfloat16)
    initial = torch.randn(1, ref["init_frames"], 16, 60, 104, generator=g).to(torch.bfloat16) if ref["init_frames"] else None
    torch.manual_seed(FIX["run_seed"])
    _, latents = pipe.inference(noise=noise.clone(), text_prompts=["p"], initial_latent=initial, return_latents=True)

    assert len(gen.calls) == len(ref["calls"]), (len(gen.calls), len(ref["calls"]))
    for i, (mine, theirs) in enumerate(zip(gen.calls, ref["calls"])):
        for key in ("branch", "timestep", "current_start", "end_before", "end_after", "frames"):
            assert mine[key] == theirs[key], f"call {i}: {key} {mine[key]} != {theirs[key]}"
        assert mine["x"] == theirs["x"], f"call {i}: input latents differ from the reference pipeline's"
    assert tuple(latents.shape) == ref["shape"]
    assert torch.equal(latents[:, :, :, ::4, ::4], ref["latents_sub"]) and digest(latents) == ref["latents_sha"]
