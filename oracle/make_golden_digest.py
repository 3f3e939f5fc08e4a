"""ORACLE tooling: the digest used by oracle/make_golden.py for inputs / draws, importable without the reference tree."""
import hashlib

import torch


def digest(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()
