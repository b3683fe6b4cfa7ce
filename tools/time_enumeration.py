"""Timing of the enumeration side (device PAM scan + host object build + writers) on an 8 Mb target: python tools/time_enumeration.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorg_b200 import _lib
from minorg_b200.generate_grna import find_multi_gRNA
from minorg_b200.pam import PAM
_lib.init(0)
rng = np.random.Generator(np.random.PCG64(1))
n = 8_000_000
seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].tobytes().decode()
open("/tmp/t.fa", "w").write(">big\n" + seq + "\n")
t0 = time.time(); p, m = _lib.enumerate_sites(seq, PAM(".NGG", 20).program(), 20); t1 = time.time()
print("device enumerate: %.3f s, sites %d + %d" % (t1 - t0, len(p), len(m)))
t0 = time.time(); h = find_multi_gRNA("/tmp/t.fa", ".NGG", 20); t1 = time.time()
print("find_multi_gRNA total: %.2f s, %d seqs" % (t1 - t0, len(h)))
t0 = time.time(); h.write_fasta("/tmp/g.fa", write_all=True); h.write_mapping("/tmp/g.map", write_all=True, write_checks=True); print("write: %.2f s" % (time.time() - t0))
