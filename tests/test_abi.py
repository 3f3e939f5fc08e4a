"""The C-ABI library loads without a GPU, exports every symbol include/mmpl_b200.h declares, the ctypes binding covers
them all, and the product path fails loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest
import torch

from mmpl_b200 import _build, _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mmpl_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mmpl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads_without_gpu():
    path = _build.build_library()
    assert path.exists()
    lib = _lib.load()
    assert lib.mmpl_abi_version() == _lib.ABI_VERSION
    # no link-time dependency on the driver library: it must load on a box without libcuda.so.1
    out = subprocess.run(["ldd", str(path)], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libcudart" not in out


def test_abi_version_and_build_id_guard_against_stale_libraries():
    """The binding's ABI_VERSION is the header's MMPL_ABI_VERSION, the library carries the content hash of the sources it
    was built from, and the loader compares both (a stale library is rebuilt or refused, never loaded: mmpl_b200/_lib.py)."""
    assert int(re.search(r"#define MMPL_ABI_VERSION (\d+)", HEADER.read_text()).group(1)) == _lib.ABI_VERSION
    lib = _lib.load()
    assert lib.mmpl_build_id().decode() == _build.source_id() == _build.built_id()
    assert not _build.is_stale()


def test_sampler_and_unipc_entry_refuse_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    lib = _lib.load()
    fake = C.c_void_p(0x1000)
    k = _lib.UniPCCoeffs(5.0, 0.9, 0, 0, 0, 0, 1, 0, 0, 1, 0.5, 0.1, 0.1, 1, 0)
    assert lib.mmpl_unipc_cfg_step(fake, fake, fake, fake, fake, fake, fake, fake, fake, 64, C.byref(k), None) == -2
    from mmpl_b200.unipc import FusedUniPC, unipc_table
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedUniPC(unipc_table(3, 5.0, 5.0), torch.zeros(1, 2, 16, 4, 4, dtype=torch.bfloat16))
    from mmpl_b200.scheduler import FlowMatchScheduler
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlowMatchScheduler(shift=5.0).add_noise(torch.zeros(1, 4, 2, 2, dtype=torch.bfloat16), torch.zeros(1, 4, 2, 2, dtype=torch.bfloat16), torch.zeros(1))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 20
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_build.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mmpl_[a-z0-9_]+)", nm))
    assert set(names) <= exported, f"declared but not exported: {sorted(set(names) - exported)}"
    assert set(names) == set(_lib.SIGNATURES), f"binding/header mismatch: {sorted(set(names) ^ set(_lib.SIGNATURES))}"
    assert exported == set(names), f"exported but not declared: {sorted(exported - set(names))}"


def test_sass_contains_tcgen05_and_tma():
    """Evidence that the hot kernels are Blackwell-native: UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st),
    UTMALDG (TMA) in the sm_100a SASS, and no legacy HMMA path."""
    sass = subprocess.run(["cuobjdump", "-sass", str(_build.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert not re.search(r"\bHMMA\b", sass)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    cfg = _lib.ModelConfig(256, 512, 2, 1, 256, 64, 32, 16, 16, 1e-6, 128)
    ctx = C.c_void_p()
    rc = lib.mmpl_ctx_create(C.byref(cfg), C.byref(ctx))
    assert rc == -2 and b"sm_100" in lib.mmpl_last_error()
    with pytest.raises(_lib.MmplError):
        _lib.check(rc)
    # the VAE segment-connect entry points refuse as well (arch check before any pointer is touched)
    fake = C.c_void_p(0x1000)
    assert lib.mmpl_conv3d_cl(fake, fake, None, fake, None, 1, 4, 4, 8, 8, 3, 3, 3, 0, None) == -2
    assert lib.mmpl_vae_norm_act(fake, fake, 16, 8, fake, 1, None) == -2
    assert lib.mmpl_vae_upsample2x(fake, fake, 1, 4, 4, 8, None) == -2
    assert lib.mmpl_vae_pick_odd(fake, fake, 1, 4, 4, 8, None) == -2
    assert lib.mmpl_softmax_rows(fake, 8, fake, 8, 1, 8, 1.0, None) == -2
    assert lib.mmpl_anchor_broadcast(fake, fake, 16, 0, None) == -2
    from mmpl_b200 import ops
    from mmpl_b200.attention import flash_attention
    from mmpl_b200.causal_model import CausalWanModel
    x = torch.zeros(4, 256, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="no CPU path"):
        ops.linear(x, x)
    with pytest.raises(AssertionError):
        flash_attention(torch.zeros(1, 4, 2, 128), torch.zeros(1, 4, 2, 128), torch.zeros(1, 4, 2, 128))
    m = CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=64, text_len=32).to(torch.bfloat16)
    kv = [{"k": torch.zeros(1, 64, 2, 128), "v": torch.zeros(1, 64, 2, 128), "global_end_index": torch.tensor([0]),
           "local_end_index": torch.tensor([0])}]
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16, 1, 8, 8, dtype=torch.bfloat16), t=torch.zeros(1, 1), context=torch.zeros(1, 32, 64),
          seq_len=1, kv_cache=kv, crossattn_cache=[{"is_init": False}], current_start=0)


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may touch oracle/."""
    for py in list((ROOT / "mmpl_b200").rglob("*.py")) + list((ROOT / "tools").glob("*.py")):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", py.read_text(), flags=re.M), py
