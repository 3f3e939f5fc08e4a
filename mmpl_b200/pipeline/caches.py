"""KV-cache and cross-attention-cache dicts in the reference layout:

  self-attention  list[num_layers] of {"k", "v": [B, rows, heads, 128] bf16, "global_end_index", "local_end_index": int64[1]}
                  (pipeline/causal_inference.py:278-297); the frame-slot model adds "attention_vis_index": list[int]
                  (pipeline/casual_fps_inference.py:453-482)
  cross-attention list[num_layers] of {"k", "v": [B, text_len, heads, 128] bf16, "is_init": bool}   (:299-312)

The model writes the tensors in place and replaces nothing but `is_init`, the index tensors' values and the visibility
list; a pipeline re-entering `inference()` keeps the allocations and calls `rewind()`.
"""
from __future__ import annotations

from typing import List, Optional

import torch

CONTIGUOUS_ROWS = 32760   # 21 frames x 1560 tokens (reference literal)
MMPL_SLOTS = 15           # frames 0..12 plus the far anchors 19, 20 (reference literal 32760 - 6*1560 rows)


def _zero_index(device):
    return torch.tensor([0], dtype=torch.long, device=device)


def new_kv_cache(model, batch_size: int, rows: int, dtype, device, visibility: bool = False) -> List[dict]:
    heads, head_dim = model.num_heads, model.dim // model.num_heads
    cache = []
    for _ in range(model.num_layers):
        entry = {name: torch.zeros([batch_size, rows, heads, head_dim], dtype=dtype, device=device) for name in ("k", "v")}
        entry["global_end_index"], entry["local_end_index"] = _zero_index(device), _zero_index(device)
        if visibility:
            entry["attention_vis_index"] = []
        cache.append(entry)
    return cache


def new_cross_cache(model, batch_size: int, dtype, device) -> List[dict]:
    heads, head_dim = model.num_heads, model.dim // model.num_heads
    return [{"k": torch.zeros([batch_size, model.text_len, heads, head_dim], dtype=dtype, device=device),
             "v": torch.zeros([batch_size, model.text_len, heads, head_dim], dtype=dtype, device=device),
             "is_init": False} for _ in range(model.num_layers)]


def rewind(kv_cache: Optional[List[dict]], cross_cache: Optional[List[dict]], device) -> None:
    """Start of a new rollout on existing allocations (causal_inference.py:123-132, casual_fps_inference.py:225-241): the
    text K/V must be recomputed, the write position returns to 0 (fresh index tensors, as the reference assigns them),
    nothing is visible. K/V rows are not cleared: every row is written before it is read."""
    for entry in cross_cache or []:
        entry["is_init"] = False
    for entry in kv_cache or []:
        entry["global_end_index"], entry["local_end_index"] = _zero_index(device), _zero_index(device)
        if "attention_vis_index" in entry:
            entry["attention_vis_index"] = []


def batch_of(kv_cache: Optional[List[dict]]) -> Optional[int]:
    return None if not kv_cache else kv_cache[0]["k"].shape[0]
