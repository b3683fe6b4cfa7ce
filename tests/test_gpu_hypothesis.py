"""Property-based GPU parity tests (hypothesis; run on the B200 box: pytest -m gpu): the CUDA path through the C ABI
against the CPU oracle on generated small genomes - N runs, lower case, contigs shorter than a gRNA, empty contigs,
overlapping masks, gRNA cut from the text with edits, either strand.  Derandomised, so a failure reproduces."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import c_oracle, minorg_oracle as O

pytestmark = pytest.mark.gpu

SETTINGS = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


@pytest.fixture(scope="module")
def L():
    from minorg_b200 import _lib
    assert _lib.device_count() > 0, "no CUDA device"
    _lib.init(0)
    return _lib


def _contig(max_len):
    return st.text(alphabet="ACGTACGTACGTacgtN", min_size=0, max_size=max_len)


@st.composite
def worlds(draw, glen_range=(4, 32), max_contig=400, max_grna=24):
    glen = draw(st.integers(*glen_range))
    contigs = [("c%d" % i, s) for i, s in enumerate(draw(st.lists(_contig(max_contig), min_size=1, max_size=4)))]
    grnas = []
    for _ in range(draw(st.integers(1, max_grna))):
        ci = draw(st.integers(0, len(contigs) - 1))
        seq = contigs[ci][1]
        if len(seq) >= glen and draw(st.booleans()):
            p = draw(st.integers(0, len(seq) - glen))
            g = list(seq[p:p + glen].upper().replace("N", "A"))
            for _ in range(draw(st.integers(0, 3))):                      # substitutions
                g[draw(st.integers(0, glen - 1))] = draw(st.sampled_from("ACGT"))
            if glen > 4 and draw(st.booleans()):                          # one indel, length kept
                k = draw(st.integers(1, glen - 2))
                g = g[:k] + g[k + 1:] + [draw(st.sampled_from("ACGT"))] if draw(st.booleans()) else \
                    g[:k] + [draw(st.sampled_from("ACGT"))] + g[k:glen - 1]
            g = "".join(g)
            grnas.append(O.reverse_complement(g) if draw(st.booleans()) else g)
        else:
            grnas.append(draw(st.text(alphabet="ACGTN", min_size=glen, max_size=glen)))
    masks = []
    for _ in range(draw(st.integers(0, 4))):
        ci = draw(st.integers(0, len(contigs) - 1))
        n = len(contigs[ci][1])
        if n:
            a = draw(st.integers(0, n - 1))
            masks.append((ci, a, draw(st.integers(a + 1, min(n, a + glen + 6)))))
    return glen, contigs, grnas, masks


def _screen(L, contigs, masks, grnas, glen, pam=".NGG", **kw):
    from minorg_b200.ot_regex import make_criteria
    from minorg_b200.pam import PAM
    blob = np.frombuffer("".join(s for _, s in contigs).encode() or b"\0", dtype=np.uint8)
    offsets = np.zeros(len(contigs) + 1, dtype=np.int64)
    np.cumsum([len(s) for _, s in contigs], out=offsets[1:])
    gb = np.frombuffer("".join(grnas).encode(), dtype=np.uint8)
    crit = make_criteria(grna_length=glen, **kw)
    return L.screen_host(blob, offsets, masks, gb, len(grnas), glen, crit, PAM(pam, glen).program())


@settings(max_examples=400, **SETTINGS)
@given(worlds(), st.integers(0, 7))
def test_ungapped_counts_equal_c_oracle(L, world, M):
    glen, contigs, grnas, masks = world
    M = min(M, glen - 1)
    got = _screen(L, contigs, masks, grnas, glen, ot_mismatch=M + 1, ot_gap=0)
    want = c_oracle.screen_hamming(contigs, masks, grnas, M, threads=1)
    assert np.array_equal(got, want)


@settings(max_examples=300, **SETTINGS)
@given(worlds(glen_range=(5, 32), max_contig=250, max_grna=12), st.integers(0, 3), st.integers(1, 3))
def test_gapped_counts_equal_c_oracle(L, world, M, Gp):
    """No-pattern criterion with gaps and trimming: per-gRNA anchor counts (both orientations summed by the host call)
    against the exhaustive C oracle; covers the closed-form one-gap verifier and the banded-DP verifier (2-3 gaps)."""
    glen, contigs, grnas, masks = world
    got = _screen(L, contigs, masks, grnas, glen, ot_mismatch=M + 1, ot_gap=Gp + 1)
    c = c_oracle.screen_nopattern(contigs, masks, grnas, M, Gp, threads=1).astype(np.int64)
    n = len(grnas)
    assert np.array_equal(np.asarray(got, dtype=np.int64), c[:n] + c[n:])


@settings(max_examples=150, **SETTINGS)
@given(worlds(glen_range=(8, 12), max_contig=50, max_grna=5),
       st.sampled_from([None, "1mg1-", "0mg-5,2mg1-", "1m1-|0g6", "1mi1-,0d-4", "0mg-6,1mg-7-", "2m1-,0g1-"]),
       st.integers(1, 3), st.integers(0, 3), st.booleans(), st.booleans(), st.booleans(),
       st.sampled_from([".NGG", "TTV.", ".NG", "R{1,2}T"]))
def test_general_verdicts_equal_python_oracle(L, world, pattern, ot_mismatch, ot_gap, uam, uag, pamless, pam):
    """Patterns, PAM proximity, unaligned-as-mismatch/gap switches: verdicts against the Python restatement of the
    reference's predicates over the exhaustive alignment universe."""
    glen, contigs, grnas, masks = world
    names = [c[0] for c in contigs]
    crit = O.Criteria(ot_mismatch, ot_gap, pattern, uam, uag, pamless, pam, glen)
    want, _ = O.screen_exhaustive(grnas, contigs, [("m%d" % i, names[c], a, b) for i, (c, a, b) in enumerate(masks)], crit)
    got = _screen(L, contigs, masks, grnas, glen, pam=pam, ot_mismatch=ot_mismatch, ot_gap=ot_gap, ot_pattern=pattern,
                  ot_unaligned_as_mismatch=uam, ot_unaligned_as_gap=uag, ot_pamless=pamless)
    assert [x > 0 for x in np.asarray(got).tolist()] == want
