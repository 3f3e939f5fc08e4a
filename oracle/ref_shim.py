"""ORACLE tooling — imports the UNMODIFIED reference (Tele-AI/MMPL, /root/reference/MMPL_t2v) on CPU so that
oracle/make_golden.py can record golden vectors from it. Works only in the build container (the reference
tree does not exist on the GPU box) and is never imported by the product or by GPU tests.

Nothing under /root/reference is modified; the shims below live in sys.modules only:
  * namespace stubs for `wan`, `wan.modules`, `pipeline`, `demo_utils` (their __init__ pull absent packages),
  * a 3-symbol `diffusers` stub (ConfigMixin, register_to_config, ModelMixin) and an empty `ftfy`,
  * `demo_utils.memory` stub; torch.cuda.current_device patched during import (evaluated at import time),
  * CPU only: flash-attn flags forced off and wan.modules.model.flash_attention rebound to the SDPA
    `attention()` (cross-attention calls flash_attention directly, which asserts CUDA).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

REF_ROOT = Path("/root/reference/MMPL_t2v")


def available() -> bool:
    return (REF_ROOT / "wan" / "modules" / "causal_model.py").exists()


def _ns(name: str, path: Path) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [str(path)]
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference classes needed for the hot path."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    if "mmpl_ref_loaded" in sys.modules:
        return sys.modules["mmpl_ref_loaded"]
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    _ns("wan", REF_ROOT / "wan")
    _ns("wan.modules", REF_ROOT / "wan" / "modules")
    _ns("wan.utils", REF_ROOT / "wan" / "utils")
    _ns("pipeline", REF_ROOT / "pipeline")
    _ns("demo_utils", REF_ROOT / "demo_utils")

    # diffusers stub
    d = types.ModuleType("diffusers")
    d.__path__ = []
    cu = types.ModuleType("diffusers.configuration_utils")

    class ConfigMixin:
        pass

    def register_to_config(fn):
        return fn

    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    dm = types.ModuleType("diffusers.models")
    dm.__path__ = []
    mu = types.ModuleType("diffusers.models.modeling_utils")

    class ModelMixin(torch.nn.Module):
        pass

    mu.ModelMixin = ModelMixin
    sys.modules.update({"diffusers": d, "diffusers.configuration_utils": cu, "diffusers.models": dm,
                        "diffusers.models.modeling_utils": mu})
    sys.modules.setdefault("ftfy", types.ModuleType("ftfy"))

    mem = types.ModuleType("demo_utils.memory")
    mem.gpu = torch.device("cpu")
    mem.get_cuda_free_memory_gb = lambda *a, **k: 0.0
    mem.DynamicSwapInstaller = object
    mem.move_model_to_device_with_memory_preservation = lambda *a, **k: None
    sys.modules["demo_utils.memory"] = mem

    orig_cur = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0
    try:
        import importlib
        attention = importlib.import_module("wan.modules.attention")
        attention.FLASH_ATTN_2_AVAILABLE = False
        attention.FLASH_ATTN_3_AVAILABLE = False
        model = importlib.import_module("wan.modules.model")
        model.flash_attention = lambda q, k, v, k_lens=None, **kw: attention.attention(q, k, v)
        causal_model = importlib.import_module("wan.modules.causal_model")
        scheduler = importlib.import_module("utils.scheduler")
        wan_wrapper = importlib.import_module("utils.wan_wrapper")
        causal_inference = importlib.import_module("pipeline.causal_inference")
    finally:
        torch.cuda.current_device = orig_cur

    ns = types.ModuleType("mmpl_ref_loaded")
    ns.attention = attention
    ns.model = model
    ns.causal_model = causal_model
    ns.scheduler = scheduler
    ns.wan_wrapper = wan_wrapper
    ns.causal_inference = causal_inference
    sys.modules["mmpl_ref_loaded"] = ns
    return ns


def build_reference_pipeline(ref, cfg, weights: dict, prompt_embeds: torch.Tensor, *, shift=5.0,
                             denoising_step_list=(1000, 750, 500, 250), num_frame_per_block=3, frame_seq_length=1560,
                             cache_rows=32760):
    """CausalInferencePipeline around a CausalWanModel with `cfg` dims and `weights` (state dict), with the fake
    text encoder / VAE and the overrides listed in SURVEY.md §8c item 5."""
    CausalWanModel = ref.causal_model.CausalWanModel
    model = CausalWanModel(model_type="t2v", patch_size=cfg.patch_size, text_len=cfg.text_len, in_dim=cfg.in_dim,
                           dim=cfg.dim, ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim, text_dim=cfg.text_dim,
                           out_dim=cfg.out_dim, num_heads=cfg.num_heads, num_layers=cfg.num_layers, eps=cfg.eps)
    missing, unexpected = model.load_state_dict(weights, strict=True), None
    model = model.to(torch.bfloat16).eval().requires_grad_(False)

    W = ref.wan_wrapper.WanDiffusionWrapper
    gen = W.__new__(W)
    torch.nn.Module.__init__(gen)
    gen.model = model
    gen.uniform_timestep = False
    gen.scheduler = ref.scheduler.FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
    gen.scheduler.set_timesteps(1000, training=True)
    gen.seq_len = 32760
    gen.post_init()

    class FakeText(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": prompt_embeds}

    class FakeVAE(torch.nn.Module):
        def decode_to_pixel(self, latents, use_cache=False):
            return latents

    args = types.SimpleNamespace(denoising_step_list=list(denoising_step_list), warp_denoising_step=True,
                                 independent_first_frame=False, context_noise=0,
                                 num_frame_per_block=num_frame_per_block, model_kwargs={})
    pipe = ref.causal_inference.CausalInferencePipeline(args, torch.device("cpu"), generator=gen,
                                                        text_encoder=FakeText(), vae=FakeVAE())
    pipe.num_transformer_blocks = cfg.num_layers
    pipe.frame_seq_length = frame_seq_length
    # caches with this model's head count (the reference hard-codes 12 heads, causal_inference.py:292)
    B = prompt_embeds.shape[0]
    pipe.kv_cache1 = [{
        "k": torch.zeros([B, cache_rows, cfg.num_heads, 128], dtype=torch.bfloat16),
        "v": torch.zeros([B, cache_rows, cfg.num_heads, 128], dtype=torch.bfloat16),
        "global_end_index": torch.tensor([0], dtype=torch.long),
        "local_end_index": torch.tensor([0], dtype=torch.long),
    } for _ in range(cfg.num_layers)]
    pipe.crossattn_cache = [{
        "k": torch.zeros([B, cfg.text_len, cfg.num_heads, 128], dtype=torch.bfloat16),
        "v": torch.zeros([B, cfg.text_len, cfg.num_heads, 128], dtype=torch.bfloat16),
        "is_init": False,
    } for _ in range(cfg.num_layers)]
    return pipe


def load_unipc():
    """The reference's FlowUniPCMultistepScheduler (wan/utils/fm_solvers_unipc.py) with a minimal diffusers stub:
    SchedulerMixin / ConfigMixin whose register_to_config stores the constructor arguments in `self.config`."""
    load()
    import functools
    import importlib
    import inspect

    cu = sys.modules["diffusers.configuration_utils"]

    class ConfigMixin:
        def register_to_config(self, **kw):
            if not hasattr(self, "config"):
                self.config = types.SimpleNamespace()
            for k, v in kw.items():
                setattr(self.config, k, v)

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            self.config = types.SimpleNamespace(**{k: v for k, v in bound.arguments.items() if k != "self"})
            init(self, *args, **kwargs)
        return wrapper

    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    sch = types.ModuleType("diffusers.schedulers")
    sch.__path__ = []
    su = types.ModuleType("diffusers.schedulers.scheduling_utils")

    class SchedulerMixin:
        pass

    class SchedulerOutput:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    class KarrasDiffusionSchedulers(list):
        pass

    su.SchedulerMixin, su.SchedulerOutput = SchedulerMixin, SchedulerOutput
    su.KarrasDiffusionSchedulers = [types.SimpleNamespace(name="stub")]
    ut = types.ModuleType("diffusers.utils")
    ut.deprecate = lambda *a, **k: None
    ut.is_scipy_available = lambda: False
    sys.modules.update({"diffusers.schedulers": sch, "diffusers.schedulers.scheduling_utils": su, "diffusers.utils": ut})
    mod = importlib.import_module("wan.utils.fm_solvers_unipc")
    return mod.FlowUniPCMultistepScheduler
