"""Host logic of the pipeline mirrors on CPU with a fake generator: call schedule, timestep list, cache dict layout
and RNG consumption order of CausalInferencePipeline (pipeline/causal_inference.py) — no kernels involved."""
import types

import torch

from mmpl_b200.pipeline import CausalInferencePipeline
from _cpu_ops import cpu_scheduler, eager_unipc_factory
from oracle import causal_wan_oracle as O


class FakeGenerator(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model = types.SimpleNamespace(num_layers=3, local_attn_size=-1, num_heads=2, dim=256, text_len=32,
                                           num_frame_per_block=1)
        self.scheduler = cpu_scheduler()
        self.calls = []

    def get_scheduler(self):
        return self.scheduler

    def forward(self, noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache, current_start):
        self.calls.append((int(current_start), float(timestep.flatten()[0]), tuple(noisy_image_or_video.shape)))
        assert len(kv_cache) == 3 and set(kv_cache[0]) == {"k", "v", "global_end_index", "local_end_index"}
        assert kv_cache[0]["k"].shape == (1, 32760, 2, 128) and kv_cache[0]["global_end_index"].dtype == torch.long
        assert set(crossattn_cache[0]) == {"k", "v", "is_init"} and crossattn_cache[0]["k"].shape == (1, 32, 2, 128)
        return noisy_image_or_video * 0.1, noisy_image_or_video * 0.5


def _pipe():
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 independent_first_frame=False, context_noise=0, num_frame_per_block=3, model_kwargs={})
    text = lambda text_prompts: {"prompt_embeds": torch.zeros(1, 32, 64)}  # noqa: E731
    vae = types.SimpleNamespace(decode_to_pixel=lambda latents, use_cache=False: latents)
    gen = FakeGenerator()
    return CausalInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=text, vae=vae), gen


def test_chunk_and_step_schedule_matches_reference_structure():
    pipe, gen = _pipe()
    assert [round(float(t), 2) for t in pipe.denoising_step_list] == [1000.0, 937.5, 833.33, 625.0]
    assert torch.equal(pipe.denoising_step_list, O.FlowMatchSchedule(5.0).warped_steps([1000, 750, 500, 250]))
    noise = torch.randn(1, 6, 16, 8, 12, generator=torch.Generator().manual_seed(0))
    calls_to_randn = []
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: (calls_to_randn.append(tuple(x.shape)), orig(x, *a, **k))[1]
    try:
        video, latents = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    finally:
        torch.randn_like = orig
    fs = 4 * 6
    expect = []
    for chunk in range(2):
        expect += [(chunk * 3 * fs, t) for t in (1000.0, 937.5, 833.3333129882812, 625.0, 0.0)]
    assert [(c[0], c[1]) for c in gen.calls] == expect
    assert all(c[2] == (1, 3, 16, 8, 12) for c in gen.calls)
    assert calls_to_randn == [(3, 16, 8, 12)] * 6          # one draw per non-final step, per chunk
    assert latents.shape == noise.shape and video.min() >= 0 and video.max() <= 1
    # second call resets the caches in place (indices replaced by fresh zero tensors, is_init cleared)
    pipe.kv_cache1[0]["global_end_index"].fill_(123)
    pipe.crossattn_cache[0]["is_init"] = True
    k_ptr = pipe.kv_cache1[0]["k"].data_ptr()
    pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    assert pipe.kv_cache1[0]["k"].data_ptr() == k_ptr and len(gen.calls) == 20


def test_initial_latent_prefill_calls():
    """I2V / video-extension branch (pipeline/causal_inference.py:136-169): prefill at timestep 0, output shifted."""
    pipe, gen = _pipe()
    noise = torch.randn(1, 3, 16, 8, 12)
    init = torch.randn(1, 3, 16, 8, 12)
    _, latents = pipe.inference(noise=noise, text_prompts=["p"], initial_latent=init, return_latents=True)
    assert latents.shape[1] == 6 and torch.equal(latents[:, :3], init)
    assert gen.calls[0][:2] == (0, 0.0) and gen.calls[1][:2] == (3 * 24, 1000.0) and len(gen.calls) == 6


def test_cfg_diffusion_pipeline_schedule():
    """CausalDiffusionInferencePipeline (pipeline/causal_diffusion_inference.py): per chunk, `steps` x (cond, uncond)
    forwards followed by the clean-context pass on both caches; contiguous current_start; separate pos/neg caches."""
    from mmpl_b200.pipeline import CausalDiffusionInferencePipeline
    gen = FakeGenerator()
    calls = []

    def fwd(noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache, current_start, cache_start):
        calls.append((conditional_dict["tag"], int(current_start), float(timestep.flatten()[0]), id(kv_cache)))
        return noisy_image_or_video * 0.1, None

    gen.forward = fwd
    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="neg",
                                 independent_first_frame=False, num_frame_per_block=3, sampling_steps=3, model_kwargs={})
    text = lambda text_prompts: {"tag": "neg" if text_prompts[0] == "neg" else "pos"}  # noqa: E731
    vae = types.SimpleNamespace(decode_to_pixel=lambda latents: latents)
    pipe = CausalDiffusionInferencePipeline(args, torch.device("cpu"), generator=gen, text_encoder=text, vae=vae)
    pipe.unipc_stepper = eager_unipc_factory(pipe)
    noise = torch.randn(1, 6, 16, 8, 12, generator=torch.Generator().manual_seed(0))
    _, lat = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    fs = 24
    steps = [float(t) for t in pipe.timesteps]
    assert len(steps) == 3 and steps == sorted(steps, reverse=True)
    expect = []
    for chunk in range(2):
        for t in steps + [0.0]:
            expect += [("pos", chunk * 3 * fs, t), ("neg", chunk * 3 * fs, t)]
    assert [c[:3] for c in calls] == expect
    assert len({c[3] for c in calls}) == 2 and lat.shape == noise.shape and torch.isfinite(lat).all()
