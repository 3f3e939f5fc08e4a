"""bench.py's host-side pieces that can be checked without a GPU: the algorithmic FLOP counts behind `roofline` / `chain`
(SURVEY.md §8d figures), the committed ncu traffic record still describing the attention kernel that is in the tree, the
chain layouts per world size, and the reference arm end to end on a tiny workload (one JSON line on stdout, contract keys)."""
import hashlib
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_algorithmic_flops_match_the_survey():
    import bench
    from tools import chain_bench
    dims = bench.WORKLOADS["cfg2"][0]
    per_forward = [bench.forward_flops(dims, 4680, (c + 1) * 4680, 3) / 1e12 for c in range(7)]
    assert [round(v, 1) for v in per_forward] == [16.2, 20.2, 24.3, 28.3, 32.3, 36.4, 40.4]
    assert abs(sum(per_forward) * 5 - 990.7) < 0.5
    d14 = dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)
    assert abs(chain_bench.segment_flops(d14, 50) / 1e12 - 136421) < 150                           # per segment
    later = chain_bench.segment_flops(d14, 50, first=False)
    assert later < chain_bench.segment_flops(d14, 50) and later > 0.9 * chain_bench.segment_flops(d14, 50)


def test_ncu_traffic_record_matches_the_attention_sources():
    rec = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
    h = hashlib.sha256()
    for name in rec["sources"]:
        h.update((ROOT / name).read_bytes())
    assert h.hexdigest()[:16] == rec["sources_sha"], \
        "the attention kernel changed after the ncu capture: re-capture (tools/ncu_traffic.py) or bench.py reports traffic null"
    assert (ROOT / rec["capture"]).exists() and rec["dram_bytes"] > rec["algorithmic_bytes"] > 0
    import bench
    traffic, info = bench.ncu_traffic("cfg2")
    assert traffic == rec["dram_bytes"] and info["traffic_source"] == rec["capture"]
    assert bench.ncu_traffic("cfg1")[0] is None


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoised_latent_frames_per_s" and d["higher_is_better"] is True
    assert d["config"]["same_config"] is True and d["steps"] == 2 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_chain_layouts_per_world_size():
    """The layouts bench.py's `chain` record measures at each N (tools/chain_bench.py), as run on 1 / 2 / 4 / 8 B200."""
    from tools.chain_bench import chain_layouts
    L = lambda n: [(v["chains"], v["slots"], v["lanes"]) for v in chain_layouts(n)]
    assert L(1) == [(1, 1, 1)]
    assert L(2) == [(1, 1, 2)]
    assert L(4) == [(2, 1, 2), (1, 2, 2)]
    assert L(8) == [(2, 2, 2), (1, 4, 2), (4, 1, 2)]
    assert all(c * s * l == n for n in (1, 2, 4, 8) for c, s, l in L(n))
