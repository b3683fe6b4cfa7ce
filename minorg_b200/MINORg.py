"""`MINORg` - the slice of the reference's orchestrator class that sits on the hot path, with the same method
names, arguments, attributes and output files (reference: minorg/MINORg.py, class MINORg):

    grna()                      MINORg.py:1917-1927   PAM enumeration over self.target  -> self.grna_hits
    filter_background(...)      MINORg.py:2146-2220   mask on-targets, screen every genome, set 'background'
    _mask_ontargets(*fnames)    MINORg.py:1933-1967   -> self.masked {filename: tuple(Masked)}
    _offtarget_hits(subject)    MINORg.py:2123-2144   -> set of failing gRNA ids
    write_mask_report(fout)     MINORg.py:1969-1992

Everything the reference delegates to `regex` and to `blastn` subprocesses runs on the GPU through
libminorg_b200.so; there is no CPU fallback.  Target discovery (`seq()`), feature/GC filters and the
minimum-set step are outside this path: `target` must be given as a FASTA file."""
import os

import numpy as np

from . import _lib
from .fasta import fasta_to_dict, read_fasta, reverse_complement
from . import distributed as _dist
from .generate_grna import find_multi_gRNA
from .minimum_set import minimum_sets
from .minimum_set_nr import minimum_sets_nr
from .ot_regex import MINORgError, OffTargetExpression, make_criteria
from .pam import PAM

_MAX_QUERIES = 1 << 20      # gRNA per query handle (the seed-and-join index addresses 2^20; larger sets are screened in parts)


class Masked:
    """A masked (on-target) interval: 0-based, end-exclusive (filter_grna.py:44-65)."""

    def __init__(self, masked, molecule, start, end):
        self.masked, self.molecule, self.start, self.end = masked, molecule, int(start), int(end)

    def __eq__(self, other):
        return isinstance(other, Masked) and (self.masked, self.molecule, self.start, self.end) == \
            (other.masked, other.molecule, other.start, other.end)

    def __hash__(self):
        return hash((self.masked, self.molecule, self.start, self.end))

    def __repr__(self):
        return "Masked(%s, %s, %d, %d)" % (self.masked, self.molecule, self.start, self.end)

    def within(self, r):
        a, b = int(min(r)), int(max(r))
        return a >= self.start and b <= self.end


def _manual_check_prompt(set_num, listed):
    """The reference's terminal dialogue for one drawn gRNA set (minimum_set.py:43-64)."""
    print("\n\tID\tsequence (Set %d)" % set_num)
    for gid, seq in listed:
        print("\t%s\t%s" % (gid, seq))
    return input("Hit 'x' to continue if you are satisfied with these sequences."
                 " Otherwise, enter the sequence ID (case-sensitive) or sequence of"
                 " an undesirable gRNA and hit the return key to update this list: ")


def _fname(x):
    return getattr(x, "filename", None) or getattr(x, "fasta", None) or str(x)


def _find_identical_host(seqs, records):
    """Mask sequences with non-ACGT characters: case-sensitive exact string search of the sequence and its
    reverse complement, where 'N' matches only 'N' - the reference does this on the host too
    (fasta.find_identical_in_fasta, fasta.py:45-95); it is string matching, not part of the device path."""
    out = []
    for sid, seq in seqs:
        for needle in {seq, reverse_complement(seq)}:
            for cid, text in records:
                pos = text.find(needle)
                while pos >= 0:
                    out.append(Masked(sid, cid, pos, pos + len(needle)))
                    pos = text.find(needle, pos + 1)
    return out


class MINORg:
    def __init__(self, directory=None, prefix="minorg", thread=1, keep_tmp=False, device=None, **kwargs):
        self.directory = os.path.abspath(directory or os.getcwd())
        self.prefix = prefix
        self.thread = thread                  # kept for API parity; the GPU path does not fork workers
        self.keep_tmp = keep_tmp
        self.auto_update_files = True
        # hot-path attributes and their reference defaults (parse_config.py:460-513, MINORg.py:686-707)
        self.target = None
        self.mask = []
        self.query = {}
        self.background = {}
        self.reference = {}
        self.genes = None
        self.ref_gene = None
        self.screen_reference = True
        self.query_reference = False
        self.ot_mismatch = 1
        self.ot_gap = 0
        self.ot_pamless = True
        self._ot_unaligned_as_mismatch = True
        self._ot_unaligned_as_gap = False
        self._ot_pattern = None
        self._pam = None
        self._length = 20
        self.pam = ".NGG"
        self.accept_invalid = False
        self.grna_hits = None
        self.grna_fasta = None
        self.grna_map = None
        self.pass_fasta = None
        self.pass_map = None
        self.final_fasta = None
        self.final_map = None
        self.sets = 1
        self.prioritise_nr = False                  # parse_config.py:659-662 (--prioritise-nr/--prioritise-pos)
        self.gc_min, self.gc_max = 0.3, 0.7        # parse_config.py:537-540
        self.exclude = None
        self.masked = {}
        self._genomes = {}
        self._genome_halo = {}
        self._window_mode = None      # (rank, world) while one subject file is split over the ranks by chunk + halo
        # Not a reference attribute.  False (default): every cell of every subject is evaluated and the per-gRNA COUNTS are
        # exact (what bench.py and the parity tests measure).  True: filter_background only needs pass/fail, so each
        # subject is screened in window slices and a gRNA that already has an off-target leaves the query set - same
        # verdicts, the counts are lower bounds.  Most useful for the gapped criteria, where most gRNA fail early.
        self.screen_verdicts_only = False
        dev = int(os.environ.get("MINORG_B200_DEVICE", "0")) if device is None else int(device)
        self.device_name, self.device_sms = _lib.init(dev)
        for k, v in kwargs.items():
            setattr(self, k, v)

    # ---- attributes with the reference's setter behaviour (MINORg.py:969-1012) -----------------
    @property
    def pam(self):
        return self._pam

    @pam.setter
    def pam(self, pam):
        self._pam = PAM(pam=pam, gRNA_length=self._length)

    @property
    def length(self):
        return self._length

    @length.setter
    def length(self, length):
        self._length = length
        self._pam = PAM(pam=self._pam.raw_pam if self._pam is not None else ".NGG", gRNA_length=length)

    @property
    def ot_pattern(self):
        return self._ot_pattern

    @ot_pattern.setter
    def ot_pattern(self, pattern):
        if pattern is None or isinstance(pattern, OffTargetExpression):
            self._ot_pattern = pattern
        else:
            self._ot_pattern = OffTargetExpression(pattern, unaligned_as_mismatch=self._ot_unaligned_as_mismatch,
                                                   unaligned_as_gap=self._ot_unaligned_as_gap)

    @property
    def ot_unaligned_as_mismatch(self):
        return self._ot_unaligned_as_mismatch

    @ot_unaligned_as_mismatch.setter
    def ot_unaligned_as_mismatch(self, v):
        self._ot_unaligned_as_mismatch = bool(v)
        if self._ot_pattern is not None:      # (the reference raises AttributeError here: typo at MINORg.py:1003)
            self.ot_pattern = self._ot_pattern.pattern

    @property
    def ot_unaligned_as_gap(self):
        return self._ot_unaligned_as_gap

    @ot_unaligned_as_gap.setter
    def ot_unaligned_as_gap(self, v):
        self._ot_unaligned_as_gap = bool(v)
        if self._ot_pattern is not None:
            self.ot_pattern = self._ot_pattern.pattern

    # ---- files ---------------------------------------------------------------------------------
    def results_fname(self, name):
        d = os.path.join(self.directory, self.prefix)
        os.makedirs(d, exist_ok=True)
        return os.path.join(d, "%s_%s" % (self.prefix, name))

    def add_query(self, fname, alias=None):
        self.query[alias or "query_%d" % (len(self.query) + 1)] = str(fname)

    def add_background(self, fname, alias=None):
        self.background[alias or "bg_%d" % (len(self.background) + 1)] = str(fname)

    def write_all_grna_fasta(self, fout=None):
        fout = fout or self.grna_fasta or self.results_fname("gRNA_all.fasta")
        self.grna_fasta = fout
        self.grna_hits.write_fasta(fout, write_all=True)

    def write_all_grna_eqv(self, fout=None):
        self.grna_eqv = fout or self.results_fname("gRNA_all.eqv")
        self.grna_hits.write_equivalents(self.grna_eqv, write_all=True)

    def write_all_grna_map(self, fout=None, write_checks=True):
        fout = fout or self.grna_map or self.results_fname("gRNA_all.map")
        self.grna_map = fout
        self.grna_hits.write_mapping(fout, write_all=True, write_checks=write_checks)

    def valid_grna(self):
        """Sequences whose set sequence-level checks all pass (MINORg.py:1177-1233, sequence checks only)."""
        names = [n for n in ("background", "GC", "exclude") if any(s.check(n) is not None for s in self.grna_hits.flatten_gRNAseqs())]
        return self.grna_hits.filter_seqs(*names, accept_invalid=self.accept_invalid) if names else self.grna_hits.seqs

    def write_pass_grna_files(self, fasta=None, map=None):
        seqs = self.valid_grna()
        self.pass_map = map or self.pass_map or self.results_fname("gRNA_pass.map")
        self.pass_fasta = fasta or self.pass_fasta or self.results_fname("gRNA_pass.fasta")
        if seqs:
            self.grna_hits.write_mapping(self.pass_map, sets=[seqs], write_checks=False)
            self.grna_hits.write_fasta(self.pass_fasta, seqs=seqs)
            self.grna_hits.write_equivalents(self.results_fname("gRNA_pass.eqv"), seqs=seqs)
        else:
            open(self.pass_map, "w").close()
            open(self.pass_fasta, "w").close()

    # ---- sub-command: gRNA ---------------------------------------------------------------------
    def _generate_grna(self, fname, pam=None, length=None):
        return find_multi_gRNA(fname, pam=self.pam if pam is None else pam,
                               gRNA_len=self.length if length is None else length)

    def grna(self):
        """Generate all possible gRNA from self.target based on self.pam and self.length."""
        if not self.target:
            raise MINORgError("MINORg.grna requires self.target (FASTA of target sequences)")
        self._check_limits([self.length])
        self.grna_hits = self._generate_grna(self.target, pam=self.pam, length=self.length)
        self.write_all_grna_fasta()
        if self.auto_update_files:
            self.write_all_grna_map()
            self.write_all_grna_eqv()

    # ---- sub-command: filter (background) ------------------------------------------------------
    def _genome(self, fname, halo=256):
        """Packed genome of a FASTA file, resident in HBM (cached per file): (handle, contig names).
        In window mode (one subject split over the ranks of a torch.distributed job, self._window_mode = (rank, world))
        the handle is this rank's CHUNK of the file: the bases of its window range plus `halo` on both sides."""
        fname = _fname(fname)
        halo = max(256, int(halo))
        cached = self._genomes.get(fname)
        if cached is not None and self._window_mode is not None and self._genome_halo.get(fname, 0) < halo:
            cached[0].free()                            # a longer mask sequence needs a wider halo
            cached = None
        if cached is None:
            if self._window_mode is None:
                genome = _lib.Genome.from_fasta(fname)  # native reader (host threads; plain or gzip) + H2D + 2-bit pack
            else:
                rank, world = self._window_mode
                blob, offsets, names = _lib.read_fasta_native(fname)
                begin, end = _dist.shard_windows(_lib.layout_global_length(offsets), 1, rank, world)
                genome = _lib.Genome.from_host_array_chunk(blob, offsets, names, begin, end, halo=halo)
                self._genome_halo[fname] = halo
            if genome.n_contigs == 0:                   # an empty subject would pass every gRNA: refuse it
                raise MINORgError("%s holds no sequence records" % fname)
            cached = self._genomes[fname] = (genome, genome.names)
        return cached

    @staticmethod
    def _ranks():
        """(rank, world) of the torch.distributed job this process belongs to, (0, 1) outside one."""
        import sys
        if "torch" not in sys.modules:          # a process group can only exist if torch is already loaded;
            return 0, 1                         # a single-GPU run never pays for importing it
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        return 0, 1

    def _subject_groups(self):
        ref = {a: _fname(r) for a, r in self.reference.items()}
        return ref, {a: _fname(f) for a, f in self.query.items()}, {a: _fname(f) for a, f in self.background.items()}

    def _mask_ontargets(self, *mask_fnames):
        """All intervals of every subject file identical (end to end, either strand) to a sequence of the
        given FASTA files."""
        self.masked = {}
        seqs = []
        for mf in mask_fnames:
            if mf:
                seqs.extend(fasta_to_dict(mf).items())
        # filter_grna.py:71: U/u count as standard bases and go down the BLAST path, where U pairs with T
        standard = set("ACGTUacgtu")
        plain = [(i, s.replace("U", "T").replace("u", "t")) for i, s in seqs if s and set(s) <= standard]
        other = [(i, s) for i, s in seqs if s and not set(s) <= standard]
        halo = 64 + max([len(s) for _, s in plain] or [0])   # a chunk reports the occurrences that START in its range
        for group in self._subject_groups():
            for fname in group.values():
                if fname in self.masked:
                    continue
                genome, names = self._genome(fname, halo=halo)
                found = set()
                if plain:
                    for si, ci, start, end, _strand in genome.find_exact([s for _, s in plain]):
                        found.add(Masked(plain[si][0], names[ci], start, end))
                if self._window_mode is not None:       # every rank needs every mask: an interval may span a cut
                    import torch.distributed as dist
                    parts = [None] * self._window_mode[1]
                    dist.all_gather_object(parts, sorted((m.masked, m.molecule, m.start, m.end) for m in found))
                    found = {Masked(*t) for part in parts for t in part}
                if other:
                    found.update(_find_identical_host(other, read_fasta(fname)))
                self.masked[fname] = tuple(found)

    def write_mask_report(self, fout):
        ref, query, bg = self._subject_groups()
        inv = {}
        for alias, fname in list(ref.items()) + list(query.items()) + list(bg.items()):
            inv[fname] = alias                                   # later aliases win, like the dict merge at :1980-1983
        with open(fout, "w") as f:
            f.write("#alias\tfname\n")
            for fname, alias in inv.items():
                f.write("#%s\t%s\n" % (alias, fname))
            f.write("\t".join(["source", "molecule", "start", "end", "masked"]) + "\n")
            for fname, masked in self.masked.items():
                for m in masked:
                    f.write("\t".join([str(inv[fname]), m.molecule, str(m.start), str(m.end), m.masked]) + "\n")

    def _check_limits(self, lengths):
        """The device kernels hold a gRNA in one 64-bit 2-bit word and the gap search in MG_MAX_GAPS diagonals; the
        reference has no such limits (generate_grna.py:23), so exceeding them must fail HERE, by name, not deep in C."""
        too_long = sorted({n for n in lengths if n > _lib.MG_MAX_GRNA_LEN or n < 1})
        if too_long:
            raise MINORgError("gRNA length(s) %s outside 1..%d: not supported by the device screen (MG_MAX_GRNA_LEN)"
                              % (too_long, _lib.MG_MAX_GRNA_LEN))
        if max(1, int(self.ot_gap)) - 1 > _lib.MG_MAX_GAPS:
            raise MINORgError("ot_gap=%s allows more than %d gap characters: not supported by the device screen (MG_MAX_GAPS)"
                              % (self.ot_gap, _lib.MG_MAX_GAPS))

    def _criteria(self, length):
        self._check_limits([length])
        return make_criteria(ot_mismatch=self.ot_mismatch, ot_gap=self.ot_gap, ot_pattern=self.ot_pattern,
                             ot_unaligned_as_mismatch=self.ot_unaligned_as_mismatch,
                             ot_unaligned_as_gap=self.ot_unaligned_as_gap, ot_pamless=self.ot_pamless,
                             grna_length=length)

    def offtarget_counts(self, fasta_subject, seqs, windows=None):
        """Per-sequence count of problematic, unmasked alignments in one subject file (device screen).
        windows=(rank, world) restricts the scan to that rank's share of the window starts (one genome split
        over several GPUs; the caller sums the shares, e.g. with distributed.allreduce_counts)."""
        fname = _fname(fasta_subject)
        genome, names = self._genome(fname)
        index = {n: i for i, n in enumerate(names)}
        genome.set_masks([(index[m.molecule], m.start, m.end) for m in self.masked.get(fname, ()) if m.molecule in index])
        out = np.zeros(len(seqs), dtype=np.int64)
        by_len = {}
        for i, s in enumerate(seqs):
            by_len.setdefault(len(s), []).append(i)
        stream = 0
        for length, idx in by_len.items():
            pam = PAM(self.pam, length).program() if not self.ot_pamless else None
            own_b, own_e = genome.owned_range[:2]
            if windows is None or (own_b, own_e) != (0, genome.global_length):
                wb, we = own_b, own_e                   # whole genome, or this rank's chunk (mg_screen clips to it anyway)
            else:
                wb, we = _dist.shard_windows(genome.global_length, length, *windows)   # whole genome resident: range only
            we = min(we, max(0, genome.global_length - length + 1))
            if we <= wb:
                continue
            crit = self._criteria(length)
            # exact counts: one pass.  verdicts only: geometric window slices, survivors only carry on
            n_slices = 12 if self.screen_verdicts_only else 1
            cuts = [wb + ((we - wb) >> (n_slices - 1 - k)) // 32 * 32 for k in range(n_slices - 1)] + [we]
            alive, begin = list(idx), wb
            for end in cuts:
                if end <= begin or not alive:
                    continue
                c = np.zeros(len(alive), dtype=np.int64)
                for j in range(0, len(alive), _MAX_QUERIES):      # one query handle indexes at most 2^20 gRNA for the seed-and-join screen
                    part = alive[j:j + _MAX_QUERIES]
                    q = _lib.Queries([seqs[i] for i in part], stream=stream)
                    cp = _lib.screen_counts(genome, q, crit, pam, win_begin=begin, win_end=end, stream=stream).astype(np.int64)
                    c[j:j + len(part)] = cp[:len(part)] + cp[len(part):]
                    q.free()
                out[alive] += c
                if self.screen_verdicts_only:
                    alive = [i for i, x in zip(alive, c) if x == 0]
                begin = end
        return out

    def _grna_fasta_dict(self):
        """The records of self.grna_fasta, read once per (file, modification time, size): filter_background looks at them per
        subject file, and a 10^6-gRNA FASTA takes seconds to parse."""
        st = os.stat(self.grna_fasta)
        key = (self.grna_fasta, st.st_mtime_ns, st.st_size)
        if getattr(self, "_grna_fasta_cache", (None, None))[0] != key:
            self._grna_fasta_cache = (key, fasta_to_dict(self.grna_fasta))
        return self._grna_fasta_cache[1]

    def _offtarget_hits(self, fasta_subject, keep_output=False, fout=None, thread=None):
        """ids of the gRNA of self.grna_fasta with a problematic off-target hit in `fasta_subject`."""
        if self.grna_fasta is None:
            self.write_all_grna_fasta()
        ids_seqs = list(self._grna_fasta_dict().items())
        if not ids_seqs:
            return set()
        counts = self.offtarget_counts(fasta_subject, [s for _, s in ids_seqs])
        return {ids_seqs[i][0] for i in np.nonzero(counts)[0].tolist()}

    def filter_background(self, *other_mask_fnames, keep_blast_output=None, mask_reference=True):
        """Set the 'background' check of every candidate gRNA: mask on-targets in all subject files, screen
        all gRNA against reference (if screen_reference or query_reference), query and background files."""
        if not self.grna_hits:
            raise MINORgError("MINORg.filter_background requires gRNA (generated with self.grna())")
        if self.grna_fasta is None:
            self.write_all_grna_fasta()
        self._check_limits({len(s) for s in self._grna_fasta_dict().values()})     # before any device work
        ref, query, bg = self._subject_groups()
        files = []
        for enabled, group in (((self.screen_reference or self.query_reference) and ref, ref), (query, query), (bg, bg)):
            if enabled:
                files += [f for f in group.values() if f not in files]      # a file is screened once (:2188)
        rank, world = self._ranks()
        # fewer subject files than ranks: every file is cut into one chunk (+ halo) per rank instead of dealing files
        all_files = {f for group in (ref, query, bg) for f in group.values()}
        window_mode = (rank, world) if (world > 1 and len(files) < world and all_files == set(files)) else None
        if window_mode != self._window_mode:
            self._free_genomes()
            self._window_mode = window_mode
        print("Masking on-targets")
        if mask_reference and self.genes and not self.ref_gene:
            # the reference derives the gene sequences from its annotation here (MINORg.py:2168-2171); this build has no
            # GFF handling, so the caller must supply them - silently skipping the masks would fail valid gRNA
            import warnings
            warnings.warn("mask_reference=True with self.genes set but no self.ref_gene FASTA: the reference genes are "
                          "NOT masked (supply their sequences via ref_gene= or mask=)", RuntimeWarning, stacklevel=2)
        self._mask_ontargets(self.target, *([self.ref_gene] if (mask_reference and self.genes and self.ref_gene) else []),
                             *self.mask, *other_mask_fnames)
        if self._ranks()[0] == 0:
            self.write_mask_report(self.results_fname("masked.txt"))
        print("Finding off-targets")
        if world == 1:
            excl = set()
            for fname in files:
                excl |= self._offtarget_hits(fname)
        else:
            # one process per GPU: subject files are dealt to the ranks (a single large genome is split by window
            # range instead) and the per-gRNA counts are combined with ONE all-reduce over NCCL
            import torch
            ids_seqs = list(self._grna_fasta_dict().items())
            total = np.zeros(len(ids_seqs), dtype=np.int64)
            if self._window_mode is None and len(files) >= world:
                for fname in _dist.shard_files(files, rank, world):
                    total += self.offtarget_counts(fname, [s for _, s in ids_seqs])
            else:
                for fname in files:
                    total += self.offtarget_counts(fname, [s for _, s in ids_seqs], windows=(rank, world))
            t = torch.from_numpy(total).to("cuda")
            _dist.allreduce_counts(t)
            excl = {ids_seqs[i][0] for i in np.nonzero(t.cpu().numpy())[0].tolist()}
        grna_seqs = self._grna_fasta_dict()
        failed = {str(grna_seqs[i]).upper() for i in excl}
        passed = {str(s).upper() for s in grna_seqs.values()} - failed
        self.grna_hits.set_seqs_check("background", True, passed)
        self.grna_hits.set_seqs_check("background", False, failed)
        if self.auto_update_files and rank == 0:
            self.write_all_grna_map()
            self.write_pass_grna_files()

    # ---- cheap per-sequence filters next to the path ---------------------------
    def filter_gc(self, gc_min=None, gc_max=None):
        """Set the 'GC' check: gc_min <= (G+C)/len <= gc_max, '-' ignored (MINORg.py:2222-2238, grna.py:1284-1298,
        functions.py:490-493)."""
        lo = self.gc_min if gc_min is None else gc_min
        hi = self.gc_max if gc_max is None else gc_max
        for s in self.grna_hits.flatten_gRNAseqs():
            seq = s.seq.upper().replace("-", "")
            s.set_check("GC", lo <= (seq.count("G") + seq.count("C")) / len(seq) <= hi)
        if self.auto_update_files:
            self.write_all_grna_map()
            self.write_pass_grna_files()

    def filter_exclude(self):
        """Set the 'exclude' check from the FASTA file at self.exclude (MINORg.py:2449-2463)."""
        if self.exclude is not None:
            banned = {str(v).upper() for v in fasta_to_dict(self.exclude).values()}
            for s in self.grna_hits.flatten_gRNAseqs():
                s.set_check("exclude", s.seq not in banned)
            if self.auto_update_files:
                self.write_all_grna_map()
                self.write_pass_grna_files()

    def filter(self, background_check=True, feature_check=False, gc_check=True):
        """MINORg.py:2465-2486 without the feature check (needs GFF + MAFFT: out of scope)."""
        if background_check:
            self.filter_background()
        if gc_check:
            self.filter_gc()

    # ---- sub-command: minimumset -----------------------------------------------------------------
    def minimumset(self, sets=None, manual=None, fasta=None, fout_fasta=None, fout_map=None, fout_eqv=None,
                   report_full_path=True, **_unused_check_switches):
        """Generate `sets` mutually exclusive minimum sets of passing gRNA that each cover all targets and write
        <prefix>_gRNA_final.fasta / .map (MINORg.py:2530-2650, non-interactive paths: LAR set cover with ties -> closest
        to 5' -> least redundant -> first; or, with self.prioritise_nr, the non-redundancy-first covers of
        minimum_set_nr.py).  manual=True shows every drawn set on the terminal and lets the user reject gRNA by id or
        sequence before it is accepted (minimum_set.py:43-64, 603-637; the reference's default unless --auto, here opt-in).
        Also writes <prefix>_gRNA_pass.eqv (coverage groups of the passing gRNA).  `fasta` renaming is not provided."""
        if not self.grna_hits:
            raise MINORgError("MINORg.minimumset requires gRNA. You may read a mapping file using"
                              " self.grna_hits.parse_from_mapping('<path to file>') or use self.grna() to generate gRNA.")
        sets = self.sets if sets is None else sets
        targets = {h.target_id for h in self.grna_hits.flatten_hits()}
        self.filter_exclude()
        valid = list(self.valid_grna())
        grna_sets = (minimum_sets_nr if self.prioritise_nr else minimum_sets)(self.grna_hits, valid, sets, targets,
                                                                             review=_manual_check_prompt if manual else None)
        fout_fasta = fout_fasta or self.final_fasta or self.results_fname("gRNA_final.fasta")
        fout_map = fout_map or self.final_map or self.results_fname("gRNA_final.map")
        if grna_sets:
            self.grna_hits.assign_seqid(assign_all=False)
            self.grna_hits.write_fasta(fout_fasta, seqs=[s for group in grna_sets for s in group])
            print("\nFinal gRNA sequence(s) have been written to %s" % (fout_fasta if report_full_path else os.path.basename(fout_fasta)))
            self.final_fasta = fout_fasta
            self.grna_hits.write_mapping(fout_map, sets=grna_sets)
            print("Final gRNA sequence ID(s), gRNA sequence(s), and target(s) have been written to %s"
                  % (fout_map if report_full_path else os.path.basename(fout_map)))
            self.final_map = fout_map
            fout_eqv = fout_eqv or self.results_fname("gRNA_pass.eqv")           # MINORg.py:2645-2647
            self.grna_hits.write_equivalents(fout_eqv, seqs=valid)
            print("Groupings of passing gRNA with equivalent coverage have been written to %s"
                  % (fout_eqv if report_full_path else os.path.basename(fout_eqv)))
        print("\n%d mutually exclusive gRNA set(s) requested. %d set(s) found." % (sets, len(grna_sets)))
        return grna_sets

    def _free_genomes(self):
        for g, _ in self._genomes.values():
            g.free()
        self._genomes = {}
        self._genome_halo = {}

    def close(self):
        self._free_genomes()
