from .causal_diffusion_inference import CausalDiffusionInferencePipeline
from .causal_fps_inference import CausalFPSInferencePipeline
from .causal_inference import CausalInferencePipeline

__all__ = ["CausalInferencePipeline", "CausalFPSInferencePipeline", "CausalDiffusionInferencePipeline"]
