"""mmpl_b200 — B200-native (sm_100a) chunk-wise causal denoising hot path of Tele-AI/MMPL.

Only what the path needs: `csrc/` (hand-written CUDA kernels + the C ABI of include/mmpl_b200.h),
`ops` (per-kernel wrappers), and the host-side mirror of the reference interfaces
(`attention`, `causal_model`, `wan_wrapper`, `scheduler`, `pipeline`).
"""
__version__ = "0.1.0"
