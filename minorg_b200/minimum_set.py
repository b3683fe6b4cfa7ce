"""Minimum-set step - the default path of the reference's `MINORg.minimumset` (MINORg.py:2530-2650):
mutually exclusive gRNA sets that each cover every target, by the LAR set-cover heuristic
(minimum_set.py:871-912) with ties broken by closeness to the 5' end, then by least overlap with targets that
are already covered, then by first occurrence (minimum_set.py:719-782, MINORg.py:2606-2608).

Host-side combinatorics on <= 10^3 items - no kernel here.  The non-redundancy-first
variant (prioritise_nr, minweight_sc.py) is minimum_set_nr.py.  Pinned by tests/golden/minimumset.json, produced by
executing the reference's make_get_minimum_set."""


def _targets_of(hits):
    return {h.target_id for h in hits}


def _closest_to_5prime(candidates):
    """all_best_pos: mean start of the candidate's hits on not-yet-covered targets (mean distance from the
    3' end for antisense targets)."""
    score = {}
    for seq, hits in candidates.items():
        score[seq] = sum((h.target_len - h.range[1]) if h.target.sense == "-" else h.range[0] for h in hits) / len(hits)
    best = min(score.values())
    return {seq: candidates[seq] for seq, v in score.items() if v == best}


def _least_redundant(candidates, all_hits, covered):
    """all_best_nr: fewest distinct already-covered targets among ALL hits of the candidate."""
    score = {seq: len({h.target_id for h in all_hits[seq] if h.target_id in covered}) for seq in candidates}
    best = min(score.values())
    return {seq: candidates[seq] for seq, v in score.items() if v == best}


def set_cover_lar(coverage):
    """coverage: {seq: [gRNAHit]} in candidate order -> list of chosen sequences (selection order)."""
    remaining = {seq: list(hits) for seq, hits in coverage.items()}
    covered, chosen = set(), {}
    while any(remaining.values()):
        width = {seq: len(_targets_of(hits)) for seq, hits in remaining.items()}
        widest = max(width.values())
        tied = {seq: hits for seq, hits in remaining.items() if width[seq] == widest}
        tied = _least_redundant(_closest_to_5prime(tied), coverage, covered)
        seq = next(iter(tied))
        new = _targets_of(tied[seq])
        covered |= new
        chosen[seq] = new
        remaining = {s: [h for h in hits if h.target_id not in covered] for s, hits in remaining.items()}
    # drop a sequence whose targets are all covered by the others (cannot trigger: the recorded coverages are
    # disjoint by construction - kept because the reference has the step, minimum_set.py:907-911)
    for seq in list(chosen):
        others = set().union(*[v for k, v in chosen.items() if k != seq]) if len(chosen) > 1 else set()
        if chosen[seq] <= others:
            del chosen[seq]
    return list(chosen)


def review_answer(grna_hits, picked, answer):
    """What the manual check does with an answer (minimum_set.py:618-636): 'x' accepts (None); the id or the sequence of a
    listed gRNA rejects it (its index in `picked`); anything else is invalid (-1).  (The reference looks an id up exactly and
    a sequence in upper case - and raises ValueError on a sequence typed in lower case; here case does not matter.)"""
    if answer.upper() == "X":
        return None
    ids = [grna_hits.get_gRNAseq_by_seq(s).id for s in picked]
    if answer in ids:
        return ids.index(answer)
    upper = [str(s).upper() for s in picked]
    if answer.upper() in upper:
        return upper.index(answer.upper())
    print("Invalid input.")
    return -1


def listed_for_review(grna_hits, picked):
    """(id, sequence) of a drawn set in the order the reference prints it: by gRNA id (minimum_set.py:616)."""
    return sorted(((grna_hits.get_gRNAseq_by_seq(s).id, s) for s in picked), key=lambda x: x[0])


def minimum_sets(grna_hits, seqs, n_sets, targets=None, review=None):
    """Up to n_sets mutually exclusive covers drawn from `seqs` (candidate order = order of `seqs`).
    Stops early when the remaining candidates can no longer cover every target (minimum_set.py:848-855).
    review (manual mode, minimum_set.py:603-637): called with (set number, [(id, sequence)]) for every drawn set; it returns
    'x' to accept, or the id / sequence of a gRNA to reject - the set is then drawn again without it (the other gRNA of the
    rejected draw go back to the END of the candidate order, as in the reference, :850-851)."""
    coverage = {s: list(grna_hits.get_hits(s)) for s in seqs}
    targets = set(targets) if targets else {h.target_id for hits in coverage.values() for h in hits}
    taken = {}
    out = []
    while len(out) < n_sets:
        give_back = []
        while True:
            for s in give_back:
                coverage[s] = taken.pop(s)
            reachable = {h.target_id for hits in coverage.values() for h in hits}
            picked = set_cover_lar(coverage) if coverage and reachable >= targets else []
            for s in picked:
                taken[s] = coverage.pop(s)
            if not picked or review is None:
                break
            verdict = review_answer(grna_hits, picked, review(len(out) + 1, listed_for_review(grna_hits, picked)))
            if verdict is None:
                break
            if verdict >= 0:
                del picked[verdict]
            give_back = picked
        if not picked:
            break
        out.append(picked)
    return out
