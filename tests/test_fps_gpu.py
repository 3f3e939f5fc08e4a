"""MMPL (FPS) model and pipeline on the GPU: CausalFPSWanModel mirror against golden outputs of the reference
CausalFPSWanModel (tests/golden/fps_model_tiny.pt), and the macro-from-micro pipeline mirror end to end."""
import types
from pathlib import Path

import pytest
import torch

from oracle import causal_wan_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
DEV = "cuda"


def _model(cfg, weights):
    from mmpl_b200.causal_model import CausalFPSWanModel
    m = CausalFPSWanModel(text_len=cfg.text_len, in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim,
                          text_dim=cfg.text_dim, out_dim=cfg.out_dim, num_heads=cfg.num_heads, num_layers=cfg.num_layers)
    m.load_state_dict(weights, strict=True)
    return m.to(DEV, torch.bfloat16).eval().requires_grad_(False)


def test_fps_model_matches_reference_golden():
    """Five calls through the t2v stage schedule (frame slots, 19/20 -> 13/14 remap, visibility edits, last stage
    attending cache + new K/V without writing). Tolerance: max-abs 0.0625, cosine 0.9999 on the (sub-sampled) flow."""
    fix = torch.load(GOLDEN / "fps_model_tiny.pt", weights_only=False)
    cfg = O.WanConfig(**fix["cfg"])
    w = O.make_weights(cfg, fix["weight_seed"])
    model = _model(cfg, w)
    g = torch.Generator().manual_seed(fix["input_seed"])
    noise = torch.randn(1, 21, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompt = torch.randn(1, cfg.text_len, cfg.text_dim, generator=g).to(torch.bfloat16).to(DEV)
    fs, rows, L, H = 1560, 15 * 1560, cfg.num_layers, cfg.num_heads
    kv = [{"k": torch.zeros(1, rows, H, 128, dtype=torch.bfloat16, device=DEV), "v": torch.zeros(1, rows, H, 128, dtype=torch.bfloat16, device=DEV),
           "global_end_index": torch.tensor([0], device=DEV), "local_end_index": torch.tensor([0], device=DEV),
           "attention_vis_index": []} for _ in range(L)]
    cross = [{"k": torch.zeros(1, cfg.text_len, H, 128, dtype=torch.bfloat16, device=DEV),
              "v": torch.zeros(1, cfg.text_len, H, 128, dtype=torch.bfloat16, device=DEV), "is_init": False} for _ in range(L)]
    for i, call in enumerate(fix["calls"]):
        if i == 3:
            for blk in kv:
                for val in (31200, 29640):
                    if val in blk["attention_vis_index"]:
                        blk["attention_vis_index"].remove(val)
        if i == 4:
            for blk in kv:
                for val in (31200, 29640):
                    if val not in blk["attention_vis_index"]:
                        blk["attention_vis_index"].append(val)
        frames = call["frames"]
        before = kv[1]["k"][0].clone()
        t = torch.full((1, len(frames)), call["t"], device=DEV)
        flow = model(noise[:, frames].permute(0, 2, 1, 3, 4), t=t, context=prompt, seq_len=32760, kv_cache=kv,
                     crossattn_cache=cross, current_start=[f * fs for f in frames], cache_start=[f * fs for f in frames])
        changed = (kv[1]["k"][0] != before).flatten(1).any(dim=1)
        assert sorted(set((changed.nonzero().flatten() // fs).tolist())) == call["changed_slots"]
        assert sorted(kv[0]["attention_vis_index"]) == call["vis"]
        assert (int(kv[0]["global_end_index"]), int(kv[0]["local_end_index"])) == (0, 0)
        got = flow[0].permute(1, 0, 2, 3)[:, :, ::4, ::4].float().cpu()
        ref = call["flow_sub"].float()
        err = (got - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        print(f"fps call {i} frames {frames}: max_abs={err:.4g} cos={cos:.6f}")
        assert err <= 0.0625 and cos >= 0.9999
    for name, got in (("kv_k_layer1_sub", kv[1]["k"][0, ::97]), ("kv_v_layer0_sub", kv[0]["v"][0, ::97])):
        ref = fix[name].float()
        err = (got.float().cpu() - ref).abs().max().item()
        print(name, err)
        assert err <= 0.125


def test_fps_pipeline_runs_the_macro_from_micro_schedule():
    """t2v schedule with 3 sampling steps on a 2-block model at 16x24 latents: stage order, forwards per stage
    (2 per step + 2 context), anchors payload, visibility lists at the end, finite output, determinism."""
    from mmpl_b200.pipeline import CausalFPSInferencePipeline
    from mmpl_b200.wan_wrapper import WanFPSWrapper
    cfg = O.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
    gen = WanFPSWrapper(model=_model(cfg, O.make_weights(cfg, 1)), timestep_shift=5.0)
    prompt = torch.randn(1, cfg.text_len, cfg.text_dim, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(DEV)

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": prompt if text_prompts[0] != "neg" else -prompt}

    class VAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="neg",
                                 independent_first_frame=False, sampling_steps=3, model_kwargs={})
    torch.manual_seed(0)
    anchors = []
    pipe = CausalFPSInferencePipeline(args, DEV, generator=gen, text_encoder=Text(), vae=VAE(), device_cond=DEV,
                                      device_uncond=DEV, anchor_sink=anchors.append)
    calls = []
    h = gen.register_forward_hook(lambda m, a, kw, out: calls.append(list(kw["current_start"])), with_kwargs=True)
    noise = torch.randn(1, 21, 16, 16, 24, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).to(DEV)
    torch.manual_seed(5)
    _, lat = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    h.remove()
    fs = 8 * 12
    stages = [[0, 1], [2, 3, 10, 11, 12, 19, 20], [4, 5, 6, 7, 8, 9], [13, 14, 15, 16, 17, 18]]
    expect = []
    for st in stages:
        expect += [[f * fs for f in st]] * (2 * 3 + 2)
    assert calls == expect
    assert len(anchors) == 1 and anchors[0].shape == (1, 8, 16, 16, 24)
    assert torch.equal(anchors[0][:, 0], lat[:, 0]) and torch.equal(anchors[0][:, 1:], lat[:, [2, 3, 10, 11, 12, 19, 20]])
    assert torch.isfinite(lat.float()).all() and lat.float().abs().mean() > 0
    assert sorted(pipe.kv_cache_pos[0]["attention_vis_index"]) == sorted(f * fs for f in list(range(13)) + [19, 20])
    # same seeds -> same result (no hidden state between runs; caches are reset in place)
    torch.manual_seed(5)
    _, lat2 = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    assert torch.equal(lat, lat2)
    # i2v variant: first frame given, hand-off payload = frames 0, 19, 20
    args_i = types.SimpleNamespace(**{**vars(args), "i2v": True})
    anchors_i = []
    pipe_i = CausalFPSInferencePipeline(args_i, DEV, generator=gen, text_encoder=Text(), vae=VAE(), device_cond=DEV,
                                        device_uncond=DEV, anchor_sink=anchors_i.append)
    first = torch.randn(1, 1, 16, 16, 24, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16).to(DEV)
    _, lat_i = pipe_i.inference(noise=noise, text_prompts=["p"], initial_latent=first, return_latents=True)
    assert torch.equal(lat_i[:, :1], first) and anchors_i[0].shape == (1, 3, 16, 16, 24)
    assert torch.equal(anchors_i[0], lat_i[:, [0, 19, 20]]) and torch.isfinite(lat_i.float()).all()


def test_cfg_diffusion_pipeline_on_gpu():
    """CausalDiffusionInferencePipeline mirror end to end on the contiguous-cache model: indices of both caches
    advance by one chunk per chunk, output finite and reproducible."""
    from mmpl_b200.causal_model import CausalWanModel
    from mmpl_b200.pipeline import CausalDiffusionInferencePipeline
    from mmpl_b200.wan_wrapper import WanDiffusionWrapper
    cfg = O.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
    m = CausalWanModel(text_len=cfg.text_len, dim=cfg.dim, ffn_dim=cfg.ffn_dim, text_dim=cfg.text_dim,
                       num_heads=cfg.num_heads, num_layers=cfg.num_layers)
    m.load_state_dict(O.make_weights(cfg, 4))
    gen = WanDiffusionWrapper(model=m.to(DEV, torch.bfloat16).eval(), timestep_shift=5.0)
    prompt = torch.randn(1, cfg.text_len, cfg.text_dim, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(DEV)

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": prompt if text_prompts[0] != "neg" else -prompt}

    class VAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=5.0, negative_prompt="neg",
                                 independent_first_frame=False, num_frame_per_block=3, sampling_steps=3, model_kwargs={})
    pipe = CausalDiffusionInferencePipeline(args, DEV, generator=gen, text_encoder=Text(), vae=VAE())
    noise = torch.randn(1, 6, 16, 16, 24, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).to(DEV)
    _, lat = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    fs = 8 * 12
    for cache in (pipe.kv_cache_pos, pipe.kv_cache_neg):
        assert int(cache[0]["global_end_index"]) == int(cache[-1]["local_end_index"]) == 6 * fs
    assert torch.isfinite(lat.float()).all() and lat.float().abs().mean() > 0
    _, lat2 = pipe.inference(noise=noise, text_prompts=["p"], return_latents=True)
    assert torch.equal(lat, lat2)
