"""Multi-GPU path of the MINORg-shaped class (needs >= 2 GPUs; skipped on a one-GPU box): ranks share the subject
files (or split one genome by window range) and all-reduce the per-gRNA counts over NCCL; the result must equal
the single-process oracle result."""
import gzip
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import c_oracle, minorg_oracle as O

pytestmark = pytest.mark.gpu
CFG1 = os.path.join(os.path.dirname(__file__), "golden", "config1")


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["files", "windows", "windows_gapped"])
def test_two_ranks_match_oracle(tmp_path, mode):
    d = tmp_path / "data"
    d.mkdir()
    for f in ("subset_ref_TAIR10.fasta", "subset_9654.fasta", "subset_9655.fasta"):
        with gzip.open(os.path.join(CFG1, f + ".gz"), "rb") as i, open(d / f, "wb") as o:
            shutil.copyfileobj(i, o)
    shutil.copy(os.path.join(CFG1, "targets.fasta"), d / "targets.fasta")
    out = tmp_path / "out"
    out.mkdir()
    worker = os.path.join(os.path.dirname(__file__), "mgpu_class_worker.py")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29577", worker, str(d), str(out), mode],
                   check=True, timeout=600)
    got = json.load(open(out / ("fails_%s.json" % mode)))
    want_entries = O.find_multi_grna(str(d / "targets.fasta"), ".NGG", 20)
    seqs = [s for s, _ in want_entries]
    targets = O.read_fasta(str(d / "targets.fasta"))
    fail = np.zeros(len(seqs), dtype=bool)
    subjects = ["subset_ref_TAIR10.fasta"] + (["subset_9654.fasta", "subset_9655.fasta"] if mode == "files" else [])
    for f in subjects:
        contigs = O.read_fasta(str(d / f))
        names = [c[0] for c in contigs]
        masks = [(names.index(c), s, e) for _, c, s, e in O.find_masks(targets, contigs)]
        if mode == "windows_gapped":        # ot_mismatch=2, ot_gap=2 -> M = 1, Gp = 1
            c = c_oracle.screen_nopattern(contigs, masks, seqs, 1, 1, threads=16)
            fail |= (c[:len(seqs)] + c[len(seqs):]) > 0
        else:
            fail |= c_oracle.screen_hamming(contigs, masks, seqs, 1, threads=8) > 0
    assert got == [s for s, fl in zip(seqs, fail) if fl]
    assert 0 < len(got) < len(seqs)
