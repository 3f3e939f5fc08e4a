"""CausalFPSInferencePipeline mirror against the *reference pipeline itself* (SURVEY.md §8 rows a13 / f-1): the goldens in
tests/golden/fps_pipeline_{t2v,i2v}.pt were recorded by running the unmodified reference
pipeline/casual_fps_inference.py (MMPL_t2v and MMPL_i2v) around oracle/fake_fps_generator.FakeFPSGenerator
(oracle/make_golden_fps_pipeline.py). The mirror runs around the same fake generator on the CPU and must reproduce, bit
for bit: every generator call (branch, timesteps, frame positions, visibility list before/after, input latents), the
re-noising timestep drawn at construction, the anchor hand-off payload, the final visibility list and the final latents -
i.e. stage schedule, CFG combine, UniPC update, re-noising and RNG order."""
import types
from pathlib import Path

import pytest
import torch

from mmpl_b200.pipeline import CausalFPSInferencePipeline
from _cpu_ops import cpu_scheduler, eager_unipc_factory
from oracle.fake_fps_generator import FakeFPSGenerator, digest

GOLDEN = Path(__file__).parent / "golden"


def build(variant):
    gen = FakeFPSGenerator(cpu_scheduler())
    text = lambda text_prompts: {"prompt_embeds": torch.full((1, 32, 64), -1.0 if text_prompts[0] == "__negative__" else 1.0,
                                                             dtype=torch.bfloat16)}
    vae = types.SimpleNamespace(decode_to_pixel=lambda latents, use_cache=False: latents)
    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="__negative__",
                                 independent_first_frame=False, model_kwargs={}, sampling_steps=3, i2v=variant == "i2v")
    return gen, args, text, vae


def inputs():
    g = torch.Generator().manual_seed(21)
    noise = torch.randn(1, 21, 16, 60, 104, generator=g).to(torch.bfloat16)
    first = torch.randn(1, 1, 16, 60, 104, generator=g).to(torch.bfloat16)
    connect = torch.randn(1, 2, 16, 60, 104, generator=g).to(torch.bfloat16)
    return noise, first, connect


@pytest.mark.parametrize("variant,case", [("t2v", "plain"), ("t2v", "extend"), ("i2v", "image"), ("i2v", "connect")])
def test_pipeline_mirror_reproduces_the_reference_pipeline(variant, case):
    fix = torch.load(GOLDEN / f"fps_pipeline_{variant}.pt", weights_only=False)
    ref = fix["runs"][case]
    gen, args, text, vae = build(variant)
    assert args.sampling_steps == fix["sampling_steps"]
    torch.manual_seed(fix["ctor_seed"])
    sent = []
    pipe = CausalFPSInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=text, vae=vae, device_cond="cpu",
                                      device_uncond="cpu", anchor_sink=sent.append)
    pipe.unipc_stepper = eager_unipc_factory(pipe)   # CPU: the oracle's eager operators stand in for the fused UniPC kernel
    assert torch.equal(pipe.ddmp_timestep, fix["ddmp_timestep"]), "re-noising timestep (constructor randint) differs"
    noise, first, connect = inputs()
    initial = {"plain": None, "extend": connect, "image": first, "connect": connect}[case]
    torch.manual_seed(fix["run_seed"])
    _, latents = pipe.inference(noise=noise.clone(), text_prompts=["p"], initial_latent=initial, return_latents=True)

    assert len(gen.calls) == len(ref["calls"]), (len(gen.calls), len(ref["calls"]))
    for i, (mine, theirs) in enumerate(zip(gen.calls, ref["calls"])):
        for key in ("branch", "timestep", "current_start", "cache_start", "vis_before", "vis_after", "frames"):
            assert mine[key] == theirs[key], f"call {i}: {key} {mine[key]} != {theirs[key]}"
        assert mine["x"] == theirs["x"], f"call {i}: input latents differ from the reference pipeline's"
    assert sorted(pipe.kv_cache_pos[0]["attention_vis_index"]) == ref["vis_end"]
    assert len(sent) == 1 and tuple(sent[0].shape) == ref["anchors_shape"] and digest(sent[0]) == ref["anchors_sha"]
    assert torch.equal(latents[:, :, :, ::4, ::4], ref["latents_sub"])
    assert digest(latents) == ref["latents_sha"]
