#!/usr/bin/env python
"""Writes profiles/ncu_traffic.json: the DRAM traffic per launch of the dominant kernel (the tcgen05 self-attention at
L_kv = 32760), read from an `ncu --set full` capture, together with the hash of the kernel's sources at capture time.
bench.py reports `roofline.traffic` from this file, and only while the sources still hash to the same value.

Capture (on the GPU box; one GPU, nothing timed under the profiler is ever reported as a bench value):
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash_attn_kernel -c 4 \\
        -o gpurun_out/attn_full python tools/profile_forward.py --chunk 6 --forwards 2 --layers 2
then, here:
    python tools/ncu_traffic.py gpurun_out/attn_full.ncu-rep profiles/r02_ncu_attention_full.txt
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["mmpl_b200/csrc/attention_tcgen05.cu", "mmpl_b200/csrc/attention_dispatch.cu", "mmpl_b200/csrc/ptx.cuh"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def sources_sha() -> str:
    h = hashlib.sha256()
    for name in SOURCES:
        with open(os.path.join(ROOT, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def main(rep: str, summary_name: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def val(row, name, table):
        i = col[name]
        return float(row[i].replace(",", "")) * table[units[i]]

    attn = [r for r in data if "flash_attn_kernel" in r[col["Kernel Name"]]]
    best = max(attn, key=lambda r: val(r, "gpu__time_duration.sum", TIME))
    rd, wr = val(best, "dram__bytes_read.sum", SCALE), val(best, "dram__bytes_write.sum", SCALE)
    S, H, L = 4680, 12, 32760
    algorithmic = (2 * L + 2 * S) * H * 128 * 2   # K, V rows once + Q in, O out
    rec = {
        "kernel": best[col["Kernel Name"]][:80], "launch": f"self-attention, S={S}, L_kv={L}, {H} heads (cfg2 last chunk)",
        "dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes": algorithmic,
        "duration_us_under_ncu": val(best, "gpu__time_duration.sum", TIME),
        "tensor_pipe_active_pct": float(best[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
        if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
        "capture": f"profiles/{summary_name}", "sources": SOURCES, "sources_sha": sources_sha(),
    }
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], os.path.basename(sys.argv[2]))
