#!/usr/bin/env python
"""Measured integer-pipe peaks of this GPU (mg_microbench) -> profiles/INT_PEAKS.json: the denominators of the scan and
rows kernels' rooflines, next to the driver's MEASURED_PEAKS.json (which only has HBM and bf16 numbers).
Run on the GPU box: python tools/int_peaks.py [out.json]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minorg_b200 import _lib  # noqa: E402

NAMES = {0: "lop3", 1: "lds32", 2: "lds64_broadcast4", 3: "lds128_broadcast4", 4: "ldc_indexed", 5: "shfl",
         6: "shfl_plus_lds", 7: "imad", 8: "imad_plus_lop3"}


def main():
    name, sms = _lib.init(0)
    out = {"gpu_name": name, "sms": sms, "unit": "G lane-instructions/s", "how": "mg_microbench (csrc/api.cu), best of 3 after a warm-up launch, CUDA events",
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for k, nm in NAMES.items():
        out[nm] = _lib.microbench(k)
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True).stdout.strip().split(",")
        out["sm_mhz_after"], out["sm_max_mhz"] = float(q[0]), float(q[1])
    except Exception:
        pass
    f_hz = out.get("sm_max_mhz", 1965.0) * 1e6
    out["per_sm_per_clk_at_max_clock"] = {nm: out[nm] * 1e9 / sms / f_hz for nm in NAMES.values()}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "INT_PEAKS.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
