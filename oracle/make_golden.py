"""ORACLE tooling — generates tests/golden/*.pt by running the UNMODIFIED reference (Tele-AI/MMPL) on CPU
through oracle/ref_shim.py. Run in the build container only:  python -m oracle.make_golden [tiny] [cfg1] [fps] [cfg2]

Every fixture stores what a parity test needs and nothing that can be regenerated from a seed:
  inputs   : seeds / shapes (weights come from oracle.make_weights(cfg, seed), inputs from `synth_inputs`),
             plus sha256 digests of the generated weights and inputs to detect generator drift;
  trace    : per generator call (current_start, timestep, global_end_index, local_end_index after the call);
  eps      : the noise tensors `torch.randn_like` returned inside the denoising loop (so a CUDA run can replay them);
  outputs  : per-call x0 predictions and the final latents, bf16.
"""
from __future__ import annotations

import hashlib
import sys
import time
from pathlib import Path

import torch

from oracle import causal_wan_oracle as O
from oracle import ref_shim

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"

TINY = O.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)


def digest(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()


def synth_inputs(cfg: O.WanConfig, frames: int, lat_h: int, lat_w: int, noise_seed: int = 0, prompt_seed: int = 1):
    """Synthetic latents and prompt embeddings: randn, bf16, batch 1."""
    g0 = torch.Generator().manual_seed(noise_seed)
    g1 = torch.Generator().manual_seed(prompt_seed)
    noise = torch.randn(1, frames, cfg.in_dim, lat_h, lat_w, generator=g0).to(torch.bfloat16)
    prompt = torch.randn(1, cfg.text_len, cfg.text_dim, generator=g1).to(torch.bfloat16)
    return noise, prompt


def run_reference_causal(cfg, weights, noise, prompt, cache_rows, steps=(1000, 750, 500, 250), nfpb=3, rng_seed=1234):
    ref = ref_shim.load()
    fs = (noise.shape[3] // 2) * (noise.shape[4] // 2)
    pipe = ref_shim.build_reference_pipeline(ref, cfg, weights, prompt, denoising_step_list=steps,
                                             num_frame_per_block=nfpb, frame_seq_length=fs, cache_rows=cache_rows)
    trace, x0s, eps = [], [], []

    def hook(mod, args, kwargs, out):
        kv = kwargs["kv_cache"]
        trace.append(dict(current_start=int(kwargs["current_start"]), timestep=float(kwargs["timestep"].flatten()[0]),
                          global_end=int(kv[0]["global_end_index"].item()), local_end=int(kv[0]["local_end_index"].item())))
        x0s.append(out[1][0].clone())

    handle = pipe.generator.register_forward_hook(hook, with_kwargs=True)
    orig_randn_like = torch.randn_like

    def rec_randn_like(x, *a, **k):
        e = orig_randn_like(x, *a, **k)
        eps.append(e.clone())
        return e

    torch.manual_seed(rng_seed)
    torch.randn_like = rec_randn_like
    import contextlib
    import io
    try:
        t0 = time.perf_counter()
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            _, latents = pipe.inference(noise=noise, text_prompts=["synthetic"], return_latents=True)
        dt = time.perf_counter() - t0
    finally:
        torch.randn_like = orig_randn_like
        handle.remove()
    kv = pipe.kv_cache1
    return dict(trace=trace, x0=x0s, eps=eps, latents=latents[0], seconds=dt,
                kv_k0=kv[0]["k"][0].clone(), kv_v0=kv[0]["v"][0].clone(),
                kv_k_last=kv[-1]["k"][0].clone(), cross_k0=pipe.crossattn_cache[0]["k"][0].clone())


def make_tiny():
    cfg = TINY
    w = O.make_weights(cfg, seed=0)
    noise, prompt = synth_inputs(cfg, frames=6, lat_h=8, lat_w=12)
    r = run_reference_causal(cfg, w, noise, prompt, cache_rows=160)
    fix = dict(kind="causal_inference", cfg=cfg.__dict__, weight_seed=0, frames=6, lat_h=8, lat_w=12, cache_rows=160,
               steps=(1000, 750, 500, 250), nfpb=3, rng_seed=1234,
               weights_sha=digest([w[k] for k in sorted(w)]), inputs_sha=digest([noise, prompt]),
               trace=r["trace"], eps=r["eps"], x0=r["x0"], latents=r["latents"],
               kv_k0=r["kv_k0"], kv_v0=r["kv_v0"], kv_k_last=r["kv_k_last"], cross_k0=r["cross_k0"],
               ref_seconds=r["seconds"])
    torch.save(fix, GOLDEN / "causal_tiny.pt")
    print("tiny:", len(r["trace"]), "calls,", f"{r['seconds']:.2f}s", r["trace"][:3])


def make_cfg1():
    """BASELINE.json configs[0]: Wan-1.3B dims, 1 chunk of 3 latent frames at 30x52, 4 denoise steps + context pass."""
    cfg = O.WAN_1_3B
    t0 = time.perf_counter()
    w = O.make_weights(cfg, seed=0)
    print(f"weights in {time.perf_counter() - t0:.1f}s")
    noise, prompt = synth_inputs(cfg, frames=3, lat_h=30, lat_w=52)
    r = run_reference_causal(cfg, w, noise, prompt, cache_rows=3 * 390)
    probe = [w["blocks.0.self_attn.q.weight"], w["blocks.29.ffn.2.weight"], w["head.head.weight"], w["patch_embedding.weight"]]
    fix = dict(kind="causal_inference", cfg=cfg.__dict__, weight_seed=0, frames=3, lat_h=30, lat_w=52, cache_rows=3 * 390,
               steps=(1000, 750, 500, 250), nfpb=3, rng_seed=1234,
               weights_probe_sha=digest(probe), inputs_sha=digest([noise, prompt]),
               trace=r["trace"], eps=r["eps"], x0=r["x0"], latents=r["latents"], ref_seconds=r["seconds"])
    torch.save(fix, GOLDEN / "causal_cfg1.pt")
    print("cfg1:", len(r["trace"]), "calls,", f"{r['seconds']:.2f}s", r["trace"])


def make_fps():
    """CausalFPSWanModel (wan/modules/causal_fps_model.py) driven through the four t2v stages of the MMPL schedule,
    one forward per stage (plus a repeat of the anchor stage), with the pipeline's attention_vis_index edits between
    stages (pipeline/casual_fps_inference.py:297-325). Tiny width (dim 256, 2 blocks) but the full 60x104 latent, because
    the reference hard-codes 1560 tokens per frame in this branch. Outputs are stored sub-sampled (every 4th latent
    row/column) together with the rows of the cache that changed."""
    import importlib
    ref_shim.load()
    fps_mod = importlib.import_module("wan.modules.causal_fps_model")
    attn_mod = importlib.import_module("wan.modules.attention")
    cfg = TINY
    w = O.make_weights(cfg, seed=7)
    model = fps_mod.CausalFPSWanModel(model_type="t2v", patch_size=cfg.patch_size, text_len=cfg.text_len, in_dim=cfg.in_dim,
                                      dim=cfg.dim, ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim, text_dim=cfg.text_dim,
                                      out_dim=cfg.out_dim, num_heads=cfg.num_heads, num_layers=cfg.num_layers, eps=cfg.eps)
    model.load_state_dict(w, strict=True)
    model = model.to(torch.bfloat16).eval().requires_grad_(False)
    fs, rows = 1560, 15 * 1560
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(1, 21, 16, 60, 104, generator=g).to(torch.bfloat16)
    prompt = torch.randn(1, cfg.text_len, cfg.text_dim, generator=g).to(torch.bfloat16)
    kv = [{"k": torch.zeros(1, rows, cfg.num_heads, 128, dtype=torch.bfloat16), "v": torch.zeros(1, rows, cfg.num_heads, 128, dtype=torch.bfloat16),
           "global_end_index": torch.tensor([0]), "local_end_index": torch.tensor([0]), "attention_vis_index": []}
          for _ in range(cfg.num_layers)]
    cross = [{"k": torch.zeros(1, cfg.text_len, cfg.num_heads, 128, dtype=torch.bfloat16),
              "v": torch.zeros(1, cfg.text_len, cfg.num_heads, 128, dtype=torch.bfloat16), "is_init": False} for _ in range(cfg.num_layers)]
    stages = [[0, 1], [2, 3, 10, 11, 12, 19, 20], [2, 3, 10, 11, 12, 19, 20], [4, 5, 6, 7, 8, 9], [13, 14, 15, 16, 17, 18]]
    tvals = [999.0, 640.0, 0.0, 320.0, 87.0]
    calls = []
    for si, (frames, tv) in enumerate(zip(stages, tvals)):
        if si == 3:
            for blk in kv:
                for val in (31200, 29640):
                    if val in blk["attention_vis_index"]:
                        blk["attention_vis_index"].remove(val)
        if si == 4:
            for blk in kv:
                for val in (31200, 29640):
                    if val not in blk["attention_vis_index"]:
                        blk["attention_vis_index"].append(val)
        before = kv[1]["k"][0].clone()
        x = noise[:, frames]
        t = torch.full((1, len(frames)), tv)
        t0 = time.perf_counter()
        with torch.no_grad():
            flow = model(x.permute(0, 2, 1, 3, 4), t=t, context=prompt, seq_len=32760, kv_cache=kv, crossattn_cache=cross,
                         current_start=[f * fs for f in frames], cache_start=[f * fs for f in frames])
        changed = (kv[1]["k"][0] != before).flatten(1).any(dim=1)
        changed_slots = sorted(set((changed.nonzero().flatten() // fs).tolist()))
        calls.append(dict(frames=frames, t=tv, vis=sorted(kv[0]["attention_vis_index"]), changed_slots=changed_slots,
                          flow_sub=flow[0].permute(1, 0, 2, 3)[:, :, ::4, ::4].clone(),
                          flow_absmean=flow.float().abs().mean().item(),
                          end_indices=(int(kv[0]["global_end_index"]), int(kv[0]["local_end_index"]))))
        print(f"fps stage {si}: frames {frames} vis {calls[-1]['vis']} changed slots {changed_slots} {time.perf_counter() - t0:.1f}s")
    fix = dict(kind="fps_model", cfg=cfg.__dict__, weight_seed=7, input_seed=11, calls=calls,
               inputs_sha=digest([noise, prompt]),
               kv_k_layer1_sub=kv[1]["k"][0, ::97].clone(), kv_v_layer0_sub=kv[0]["v"][0, ::97].clone())
    torch.save(fix, GOLDEN / "fps_model_tiny.pt")


def make_cfg2():
    """BASELINE.json configs[1] (the benchmarked workload): Wan-1.3B dims, 30 blocks, 21 latent frames at 60x104 in seven
    3-frame chunks, 4 denoise steps + context pass per chunk = 35 forwards, KV 4680 -> 32760. ~20 minutes on 8 cores.
    Stored spatially sub-sampled (every 2nd latent row / column) to keep the fixture small: per chunk the x0 of the first
    (t=1000) and last (t=625) denoising call, and the final latents. The torch.randn_like draws are NOT stored (12.6 MB):
    they are the CPU generator's stream after torch.manual_seed(rng_seed), regenerated by the test and checked against
    `eps_sha`."""
    cfg = O.WAN_1_3B
    w = O.make_weights(cfg, seed=0)
    noise, prompt = synth_inputs(cfg, frames=21, lat_h=60, lat_w=104)
    r = run_reference_causal(cfg, w, noise, prompt, cache_rows=32760)
    probe = [w["blocks.0.self_attn.q.weight"], w["blocks.29.ffn.2.weight"], w["head.head.weight"], w["patch_embedding.weight"]]
    sub = lambda t: t[..., ::2, ::2].clone()
    keep = [i for i in range(len(r["x0"])) if i % 5 in (0, 3)]
    fix = dict(kind="causal_inference", cfg=cfg.__dict__, weight_seed=0, frames=21, lat_h=60, lat_w=104, cache_rows=32760,
               steps=(1000, 750, 500, 250), nfpb=3, rng_seed=1234, sub=2,
               weights_probe_sha=digest(probe), inputs_sha=digest([noise, prompt]), eps_sha=digest(r["eps"]),
               eps_shapes=[tuple(e.shape) for e in r["eps"]],
               trace=r["trace"], x0_calls=keep, x0_sub=[sub(r["x0"][i]) for i in keep], latents_sub=sub(r["latents"]),
               latents_sha=digest([r["latents"]]), ref_seconds=r["seconds"])
    torch.save(fix, GOLDEN / "causal_cfg2.pt")
    print("cfg2:", len(r["trace"]), "calls,", f"{r['seconds']:.1f}s", r["trace"][-1])


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["tiny", "cfg1"]
    if "tiny" in which:
        make_tiny()
    if "cfg1" in which:
        make_cfg1()
    if "fps" in which:
        make_fps()
    if "cfg2" in which:
        make_cfg2()
