#!/usr/bin/env python
"""Gapped screen (<=4 mm / <=1 gap) at a chosen size: python tools/bench_gapped_scale.py <genome_bp> <n_grna>
(MG_TRACE=1 prints survivors, slices and the time of the two passes).  Used for the scale check in DESIGN.md."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minorg_b200 import _lib
from minorg_b200.ot_regex import make_criteria
_lib.init(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
gen = torch.Generator(device=dev); gen.manual_seed(7)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
nb = int(float(sys.argv[1])); n = int(sys.argv[2])
ascii_ = torch.take(lut, torch.randint(0, 4, (nb,), dtype=torch.int64, device=dev, generator=gen))
g = _lib.Genome.from_device_ascii(ascii_.data_ptr(), np.array([0, nb // 2, nb], dtype=np.int64), ["a", "b"], stream=stream)
rng = np.random.Generator(np.random.PCG64(3))
grnas = [bytes(r).decode() for r in np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, 20))]]
q = _lib.Queries(grnas, stream=stream)
crit = make_criteria(grna_length=20, ot_mismatch=5, ot_gap=2)
counts = torch.zeros(2 * n, dtype=torch.int32, device=dev)
for rep in range(2):
    counts.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.screen(g, q, crit, None, counts.data_ptr(), stream=stream)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"nb": nb, "n": n, "seconds": dt, "tcell_per_s": 2.0 * n * nb / dt / 1e12}))
