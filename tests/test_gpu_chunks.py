"""Chunk + halo genomes (mg_genome_from_*_chunk): the shards of one subject, each holding only its own bases, evaluate
every cell exactly once - counts of all shards add up to the whole genome's for every prefilter (ungapped scan, edit
distance rows, exact pieces, PAM proximity), the mask finder reports every occurrence once, and a window range request
is clipped to the owned range.  Runs on one GPU; the multi-GPU runs of bench.py use the same calls, one shard per rank."""
import random

import numpy as np
import pytest

from oracle import c_oracle, minorg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from minorg_b200 import _lib
    assert _lib.device_count() > 0, "no CUDA device"
    _lib.init(0)
    return _lib


def _contig(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def _world(seed, glen=20, n_grna=64):
    rng = random.Random(seed)
    contigs = [("a", _contig(rng, 60000, "ACGT" * 12 + "N")), ("b", _contig(rng, 37)), ("c", _contig(rng, 45011)),
               ("d", ""), ("e", _contig(rng, 9000, "ACGTacgt"))]
    grnas = []
    for i in range(n_grna):
        _, seq = contigs[(0, 2, 4)[i % 3]]
        p = rng.randrange(0, len(seq) - glen - 2)
        g = list(seq[p:p + glen].upper().replace("N", "T"))
        for _ in range(i % 4):
            g[rng.randrange(glen)] = rng.choice("ACGT")
        if i % 5 == 0:
            k = rng.randrange(2, glen - 2)
            g = g[:k] + g[k + 1:] + [rng.choice("ACGT")]
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if i % 4 == 1 else g)
    return contigs, grnas


def _arrays(contigs):
    blob = np.frombuffer("".join(s for _, s in contigs).encode(), dtype=np.uint8)
    offsets = np.zeros(len(contigs) + 1, dtype=np.int64)
    np.cumsum([len(s) for _, s in contigs], out=offsets[1:])
    return blob, offsets


def _counts(L, genome, q, crit, pam, wb=0, we=-1):
    import torch
    c = torch.zeros(2 * q.n, dtype=torch.int32, device="cuda")
    L.screen(genome, q, crit, pam, c.data_ptr(), win_begin=wb, win_end=we, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return c.cpu().numpy().astype(np.int64)


def _cuts(glen, parts, rng):
    """parts contiguous ranges covering [0, glen) with ragged, unaligned cut points (one empty range included)."""
    pts = sorted(rng.randrange(1, glen) for _ in range(parts - 2))
    pts = [0] + pts + [pts[-1]] + [glen] if pts else [0, glen]
    return list(zip(pts[:-1], pts[1:]))


@pytest.mark.parametrize("source", ["host", "device"])
def test_chunks_add_up_to_the_whole(L, source):
    import torch
    from minorg_b200.ot_regex import make_criteria
    from minorg_b200.pam import PAM
    contigs, grnas = _world(5)
    names = [n for n, _ in contigs]
    blob, offsets = _arrays(contigs)
    masks = [(0, 1000, 1500), (2, 0, 300), (2, 20000, 22000), (4, 8000, 9000)]
    whole = L.Genome.from_contigs(contigs)
    whole.set_masks(masks)
    glen = whole.global_length
    assert L.layout_global_length(offsets) == glen
    q = L.Queries(grnas)
    rng = random.Random(9)
    dev = torch.from_numpy(blob.copy()).cuda()
    chunks = []
    for a, b in _cuts(glen, 6, rng):
        if source == "host":
            g = L.Genome.from_host_array_chunk(blob, offsets, names, a, b, halo=256)
        else:
            g = L.Genome.from_device_ascii_chunk(dev.data_ptr(), offsets, names, a, b, halo=300)
        assert g.owned_range[:2] == (a, b) and g.global_length == glen and g.n_contigs == len(contigs)
        g.set_masks(masks)
        chunks.append(g)
    assert sum(g.device_bytes for g in chunks) < 3 * whole.device_bytes
    pam = PAM(".NGG", 20).program()
    try:
        for kw, use_pam in ((dict(ot_mismatch=4, ot_gap=0), False), (dict(ot_mismatch=3, ot_gap=2), False),
                            (dict(ot_mismatch=2, ot_gap=3), False), (dict(ot_mismatch=4, ot_gap=2, ot_pamless=False), True),
                            (dict(ot_pattern="0mg-10,1mg-11-", ot_gap=2), False), (dict(ot_pattern="1mg1-", ot_gap=2), False),
                            (dict(ot_pattern="(0mg5,1mg6-)|(0mg6,1m7-)", ot_gap=0), False)):
            crit = make_criteria(grna_length=20, **kw)
            want = _counts(L, whole, q, crit, pam if use_pam else None)
            got = sum(_counts(L, g, q, crit, pam if use_pam else None) for g in chunks)
            assert np.array_equal(got, want), (kw, np.nonzero(got != want)[0][:10])
            assert want.sum() > 0, kw
        # a window range request on a chunk is clipped to what the chunk owns
        crit = make_criteria(ot_mismatch=4, ot_gap=0, grna_length=20)
        a, b = chunks[1].owned_range[:2]
        assert np.array_equal(_counts(L, chunks[1], q, crit, None, 0, glen), _counts(L, whole, q, crit, None, a, b))
        # the ungapped whole-genome counts are the oracle's
        n = len(grnas)
        w = _counts(L, whole, q, crit, None)
        assert np.array_equal(w[:n] + w[n:], c_oracle.screen_hamming(contigs, masks, grnas, 3, threads=8).astype(np.int64))
    finally:
        for g in chunks:
            g.free()
        q.free()
        whole.free()


def test_chunk_mask_finder_reports_each_occurrence_once(L):
    contigs, _ = _world(6)
    names = [n for n, _ in contigs]
    blob, offsets = _arrays(contigs)
    a, c = contigs[0][1].upper().replace("N", "A"), contigs[2][1]
    contigs[0] = ("a", a)
    blob, offsets = _arrays(contigs)
    seqs = [a[100:700], c[44000:45011], O.reverse_complement(c[5:405]), a[59000:60000], "ACGTACGTTTGACCA" * 20]
    whole = L.Genome.from_contigs(contigs)
    want = sorted(whole.find_exact(seqs))
    assert len(want) >= 4
    glen = whole.global_length
    # cuts through the middle of occurrences: the shard holding the FIRST base reports it
    g0 = whole.contig_global_start(0)
    g2 = whole.contig_global_start(2)
    pts = [0, g0 + 400, g0 + 59500, g2 + 200, g2 + 44500, glen]
    got = []
    for x, y in zip(pts[:-1], pts[1:]):
        g = L.Genome.from_host_array_chunk(blob, offsets, names, x, y, halo=1100)
        got += g.find_exact(seqs)
        g.free()
    assert sorted(got) == want
    # a halo shorter than a mask sequence is refused, not silently wrong
    g = L.Genome.from_host_array_chunk(blob, offsets, names, 0, glen // 2, halo=256)
    with pytest.raises(L.MinorgB200Error):
        g.find_exact(seqs)
    g.free()
    with pytest.raises(L.MinorgB200Error):
        L.Genome.from_host_array_chunk(blob, offsets, names, 0, glen // 2, halo=16)
    whole.free()


def test_subject_loader_reads_gzip_and_rejects_non_fasta(L, tmp_path):
    """mg_genome_from_fasta: gzip subjects are inflated (the repo's own config-1 genomes ship as .fasta.gz); a file
    without a single record is an error, not an empty genome that would pass every gRNA."""
    import gzip
    import os
    cfg = os.path.join(os.path.dirname(__file__), "golden", "config1", "subset_9655.fasta.gz")
    plain = tmp_path / "subset_9655.fasta"
    with gzip.open(cfg, "rb") as i:
        plain.write_bytes(i.read())
    a, b = L.Genome.from_fasta(cfg), L.Genome.from_fasta(str(plain))
    assert a.names == b.names and a.total_bases == b.total_bases == 149576
    n = a.contig_length(0)
    assert a.decode(0, 0, min(n, 500)) == b.decode(0, 0, min(n, 500))
    a.free()
    b.free()
    junk = tmp_path / "notes.txt"
    junk.write_text("this is not\na FASTA file\n")
    with pytest.raises(L.MinorgB200Error):
        L.Genome.from_fasta(str(junk))
