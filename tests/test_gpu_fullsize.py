"""Size-independent properties at BASELINE.json's full genome size (one 135 Mb synthetic genome, 5 contigs):
planted near-matches are found exactly once each, masks remove exactly the masked ones, window shards add up,
and a gRNA and its reverse complement get the same count.  (The oracle cannot run at this size.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GENOME = 135_000_000
CONTIGS = 5
L = 20


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


@pytest.fixture(scope="module")
def world():
    import torch
    from minorg_b200 import _lib
    _lib.init(0)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2026)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_ = torch.take(lut, torch.randint(0, 4, (GENOME,), dtype=torch.uint8, device=dev, generator=gen).to(torch.int64))
    rng = np.random.Generator(np.random.PCG64(5))
    grnas = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=L)) for _ in range(64)]
    clen = GENOME // CONTIGS
    planted = []          # (gRNA index, contig, start, n_mismatch, strand)
    for gi, g in enumerate(grnas[:48]):
        for rep in range(3):
            mm = rep                      # 0, 1, 2 mismatches
            seq = list(g)
            for p in rng.choice(L, size=mm, replace=False):
                seq[p] = "ACGT"[("ACGT".index(seq[p]) + 1 + int(rng.integers(0, 3))) % 4]
            seq = "".join(seq)
            strand = 1 if (gi + rep) % 2 == 0 else -1
            text = seq if strand == 1 else _rc(seq)
            c = int(rng.integers(0, CONTIGS))
            start = int(rng.integers(1000, clen - 1000)) // 64 * 64 + gi      # plants never overlap (64-aligned + gi)
            if gi < 4 and rep < 2:                                            # a few sit exactly at a contig start / end
                c, start = gi, (0 if rep == 0 else clen - L)
            a = c * clen + start
            ascii_[a:a + L] = torch.from_numpy(np.frombuffer(text.encode(), dtype=np.uint8).copy()).to(dev)
            planted.append((gi, c, start, mm, strand))
    offsets = np.arange(CONTIGS + 1, dtype=np.int64) * clen
    stream = torch.cuda.current_stream().cuda_stream
    genome = _lib.Genome.from_device_ascii(ascii_.data_ptr(), offsets, ["c%d" % i for i in range(CONTIGS)], stream=stream)
    torch.cuda.synchronize()
    del ascii_
    yield dict(lib=_lib, torch=torch, genome=genome, grnas=grnas, planted=planted, stream=stream)
    genome.free()


def _screen(w, grnas, M, win=(0, -1), masks=None):
    from minorg_b200.ot_regex import make_criteria
    torch, L_ = w["torch"], w["lib"]
    w["genome"].set_masks(masks or [])
    q = L_.Queries(grnas, stream=w["stream"])
    counts = torch.zeros(2 * len(grnas), dtype=torch.int32, device="cuda")
    L_.screen(w["genome"], q, make_criteria(ot_mismatch=M + 1), None, counts.data_ptr(), win_begin=win[0], win_end=win[1],
              stream=w["stream"])
    c = counts.cpu().numpy().astype(np.int64)
    q.free()
    return c[:len(grnas)] + c[len(grnas):]


def test_planted_hits_found(world):
    for M in (0, 1, 2):
        want = np.zeros(len(world["grnas"]), dtype=np.int64)
        for gi, _, _, mm, _ in world["planted"]:
            want[gi] += mm <= M
        got = _screen(world, world["grnas"], M)
        assert (got >= want).all(), M
        # chance hits over 2 x 135 Mb of random sequence, 64 gRNA: 1 / 61 / 1771 of 4^20 20-mers lie within 0 / 1 / 2
        # mismatches => expected 0.016 / 1.0 / 28 extra hits in total
        assert (got - want).sum() <= (1, 6, 60)[M], (M, got - want)
        if M == 2:
            assert (got - want).sum() >= 8


def test_masks_remove_exactly_the_masked_plants(world):
    M = 2
    base = _screen(world, world["grnas"], M)
    masked = [p for i, p in enumerate(world["planted"]) if i % 3 == 0]
    got = _screen(world, world["grnas"], M, masks=[(c, s, s + L) for _, c, s, _, _ in masked])
    drop = np.zeros_like(base)
    for gi, *_ in masked:
        drop[gi] += 1
    assert np.array_equal(base - got, drop)
    # a mask one base too short on either side does not cover the hit (MINORg.py:1994-2000: span inside ONE interval)
    got2 = _screen(world, world["grnas"], M, masks=[(c, s + 1, s + L) for _, c, s, _, _ in masked] +
                   [(c, s, s + L - 1) for _, c, s, _, _ in masked])
    assert np.array_equal(got2, base)


def test_window_shards_add_up_and_strand_symmetry(world):
    M = 2
    whole = _screen(world, world["grnas"], M)
    glen = world["genome"].global_length
    cuts = [0, glen // 8 + 5, glen // 2, glen // 2 + 1, glen]
    parts = sum(_screen(world, world["grnas"], M, win=(a, b)) for a, b in zip(cuts[:-1], cuts[1:]))
    assert np.array_equal(whole, parts)
    assert np.array_equal(_screen(world, [_rc(g) for g in world["grnas"]], M), whole)


def _screen_crit(w, grnas, kw, win=(0, -1), masks=None):
    from minorg_b200.ot_regex import make_criteria
    torch, L_ = w["torch"], w["lib"]
    w["genome"].set_masks(masks or [])
    q = L_.Queries(grnas, stream=w["stream"])
    counts = torch.zeros(2 * len(grnas), dtype=torch.int32, device="cuda")
    L_.screen(w["genome"], q, make_criteria(grna_length=L, **kw), None, counts.data_ptr(), win_begin=win[0], win_end=win[1],
              stream=w["stream"])
    c = counts.cpu().numpy().astype(np.int64)
    q.free()
    return c


def test_gapped_screen_full_size_properties(world):
    """Gapped criterion (<=1 mismatch / <=1 gap: edit-distance rows kernel + closed-form verifier) over the whole 135 Mb:
    * monotone: every gRNA counts at least its ungapped <=1-mismatch hits (those alignments are in the gapped universe);
    * a gRNA with one base deleted or inserted in the middle finds the exact plant of the original gRNA through one gap;
      searched WITHOUT gaps (ot_gap=1) it does not;
    * window shards add up; a gRNA and its reverse complement swap their '+' and '-' verdicts."""
    grnas = world["grnas"]
    n = len(grnas)
    ung = _screen(world, grnas, 1)
    gap = _screen_crit(world, grnas, dict(ot_mismatch=2, ot_gap=2))
    assert ((gap[:n] + gap[n:]) >= ung).all()
    # edited copies of the first 48 gRNA (each has an exact plant): drop base 9 and append a base / insert a base at 9
    rng = np.random.Generator(np.random.PCG64(11))
    dele = [g[:9] + g[10:] + "ACGT"[int(rng.integers(0, 4))] for g in grnas[:48]]
    ins = [g[:9] + "ACGT"[int(rng.integers(0, 4))] + g[9:19] for g in grnas[:48]]
    # deleted base: 19 positions align with one gap, the appended base is a mismatch or trimmed => needs M = 1, Gp = 1;
    # inserted base: all 20 positions are consumed (19 pairs + the insertion) => M = 0, Gp = 1 is enough
    for edited, mm in ((dele, 2), (ins, 1)):
        with_gap = _screen_crit(world, edited, dict(ot_mismatch=mm, ot_gap=2))
        assert ((with_gap[:48] + with_gap[48:]) >= 1).all()
        no_gap = _screen_crit(world, edited, dict(ot_mismatch=mm, ot_gap=1))
        assert (no_gap[:48] + no_gap[48:]).sum() <= 6          # chance only (expected 0.7 at <=1 mismatch)
    glen = world["genome"].global_length
    cuts = [0, glen // 8 + 5, glen // 2, glen // 2 + 1, glen]
    parts = sum(_screen_crit(world, grnas, dict(ot_mismatch=2, ot_gap=2), win=(a, b)) for a, b in zip(cuts[:-1], cuts[1:]))
    assert np.array_equal(gap, parts)
    rc = _screen_crit(world, [_rc(g) for g in grnas], dict(ot_mismatch=2, ot_gap=2))
    # counts are distinct virtual 3' ends, which a reversed alignment set need not preserve; the verdicts are symmetric
    assert np.array_equal(rc[:n] > 0, gap[n:] > 0) and np.array_equal(rc[n:] > 0, gap[:n] > 0)


def test_human_scale_genome_beyond_2_31(tmp_path):
    """BASELINE configs[4] shape: one 3.1 Gbp genome (24 contigs).  Global coordinates exceed 2^31, so every position
    computation on the path must be 64-bit: hits planted near the END of the genome are found, in every window
    shard of an 8-way split, and masks near the end apply."""
    import torch
    from minorg_b200 import _lib
    from minorg_b200.distributed import shard_windows
    from minorg_b200.ot_regex import make_criteria
    _lib.init(0)
    dev = torch.device("cuda", 0)
    total, n_contigs = 3_100_000_000, 24
    clen = total // n_contigs
    total = clen * n_contigs
    gen = torch.Generator(device=dev)
    gen.manual_seed(30)
    codes = torch.randint(0, 4, (total,), dtype=torch.uint8, device=dev, generator=gen)
    codes += 65                                  # 'A'..'D'
    codes[codes == 66] = ord("C")
    codes[codes == 68] = ord("T")
    codes[codes == 67 + 0] = ord("C")
    ascii_ = codes                               # A, C, C, T mix is fine: any fixed i.i.d. text works here
    rng = np.random.Generator(np.random.PCG64(31))
    grnas = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=L)) for _ in range(256)]
    grnas = [g.replace("G", "A") for g in grnas]                      # text has no G: keep the gRNA alphabet matched
    plants = []
    for gi in range(64):
        c = n_contigs - 1 - (gi % 3)                                   # last contigs: global position > 2.7e9
        start = clen - 5000 - 100 * gi
        text = grnas[gi] if gi % 2 == 0 else _rc(grnas[gi])
        a = c * clen + start
        ascii_[a:a + L] = torch.from_numpy(np.frombuffer(text.encode(), dtype=np.uint8).copy()).to(dev)
        plants.append((gi, c, start))
    offsets = np.arange(n_contigs + 1, dtype=np.int64) * clen
    stream = torch.cuda.current_stream().cuda_stream
    genome = _lib.Genome.from_device_ascii(ascii_.data_ptr(), offsets, ["chr%d" % i for i in range(n_contigs)], stream=stream)
    torch.cuda.synchronize()
    del ascii_, codes
    assert genome.global_length > 2 ** 31 and genome.total_bases == total
    assert genome.decode(n_contigs - 1, clen - 5000, clen - 5000 + L) == grnas[0]
    q = _lib.Queries(grnas, stream=stream)
    crit = make_criteria(ot_mismatch=1)
    n = len(grnas)

    def run(win=(0, -1)):
        counts = torch.zeros(2 * n, dtype=torch.int32, device="cuda")
        _lib.screen(genome, q, crit, None, counts.data_ptr(), win_begin=win[0], win_end=win[1], stream=stream)
        c = counts.cpu().numpy().astype(np.int64)
        return c[:n] + c[n:]
    whole = run()
    assert (whole[:64] >= 1).all()
    parts = sum(run(shard_windows(genome.global_length, L, r, 8)) for r in range(8))
    assert np.array_equal(parts, whole)
    genome.set_masks([(c, s, s + L) for gi, c, s in plants[:32]])
    masked = run()
    assert np.array_equal(whole[:32] - masked[:32], np.ones(32, dtype=np.int64)) and np.array_equal(whole[32:], masked[32:])
    q.free()
    # the gapped path (edit-distance rows kernel, survivor queue, verifier) with 64-bit coordinates: gRNA with one base
    # inserted find their plants beyond 2^31 through one gap, the masks (still set) remove exactly the masked ones, and
    # the 8 window shards add up
    edited = [g[:9] + "C" + g[9:19] for g in grnas[:64]]
    qg = _lib.Queries(edited, stream=stream)
    critg = make_criteria(ot_mismatch=1, ot_gap=2, grna_length=L)

    def run_gapped(win=(0, -1)):
        counts = torch.zeros(2 * 64, dtype=torch.int32, device="cuda")
        _lib.screen(genome, qg, critg, None, counts.data_ptr(), win_begin=win[0], win_end=win[1], stream=stream)
        c = counts.cpu().numpy().astype(np.int64)
        return c[:64] + c[64:]
    gm = run_gapped()
    genome.set_masks([])
    gw = run_gapped()
    assert (gw >= 1).all()
    assert ((gw[:32] - gm[:32]) >= 1).all() and np.array_equal(gw[32:], gm[32:])
    assert np.array_equal(sum(run_gapped(shard_windows(genome.global_length, L, r, 8)) for r in range(8)), gw)
    qg.free()
    genome.free()
