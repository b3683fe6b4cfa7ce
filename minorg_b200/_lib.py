"""ctypes binding of libminorg_b200.so (include/minorg_b200.h).  No fallback: if the shared library is
missing or a call fails, this module raises - the product never computes on the CPU."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINORG_B200_LIB") or os.path.join(_HERE, "libminorg_b200.so")   # override: A/B builds

MG_MAX_GRNA_LEN = 32
MG_MAX_PAM_LEN = 16
MG_MAX_PAM_ALTS = 32
MG_MAX_PATTERN_UNITS = 32
MG_MAX_PATTERN_OPS = 96
MG_MAX_GAPS = 3
MG_MAX_PIECES = 8
MG_OP_AND, MG_OP_OR = -1, -2
MG_CH_MISMATCH, MG_CH_INS, MG_CH_DEL = 1, 2, 4
MG_SLICE_NONE = -(2 ** 31)
MG_ERR_OVERFLOW = -4

EXPORTS = [
    "mg_last_error", "mg_version", "mg_device_count", "mg_init",
    "mg_genome_from_ascii", "mg_genome_from_device_ascii", "mg_genome_n_contigs", "mg_genome_contig_length",
    "mg_genome_total_bases", "mg_genome_device_bytes", "mg_genome_decode", "mg_genome_free", "mg_genome_set_masks",
    "mg_mask_find", "mg_enumerate_sites", "mg_queries_create", "mg_queries_count", "mg_queries_length",
    "mg_queries_free", "mg_screen", "mg_screen_host", "mg_genome_contig_global_start", "mg_genome_global_length",
    "mg_microbench", "mg_last_screen_stats", "mg_last_seedjoin_stats",
    "mg_fasta_read", "mg_fasta_n_records", "mg_fasta_name", "mg_fasta_offsets", "mg_fasta_sequence", "mg_fasta_free",
    "mg_genome_from_fasta", "mg_screen_to_host",
    "mg_layout_global_length", "mg_genome_from_ascii_chunk", "mg_genome_from_device_ascii_chunk", "mg_genome_owned_range",
]


class MinorgB200Error(RuntimeError):
    pass


class PamSide(C.Structure):
    _fields_ = [("n_alts", C.c_int32),
                ("len", C.c_int32 * MG_MAX_PAM_ALTS),
                ("allow", ((C.c_uint32 * 8) * MG_MAX_PAM_LEN) * MG_MAX_PAM_ALTS),
                ("allow4", (C.c_uint8 * MG_MAX_PAM_LEN) * MG_MAX_PAM_ALTS)]


class Pam(C.Structure):
    _fields_ = [("five", PamSide), ("three", PamSide), ("five_maxlen", C.c_int32), ("three_maxlen", C.c_int32)]


class PatternUnit(C.Structure):
    _fields_ = [("n", C.c_int32), ("chars", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("rvs", C.c_int32),
                ("add_unaligned_m", C.c_int32), ("add_unaligned_g", C.c_int32)]


class Criteria(C.Structure):
    _fields_ = [("max_mismatch", C.c_int32), ("max_gap", C.c_int32), ("pamless", C.c_int32), ("n_units", C.c_int32),
                ("n_ops", C.c_int32), ("units", PatternUnit * MG_MAX_PATTERN_UNITS), ("ops", C.c_int32 * MG_MAX_PATTERN_OPS),
                ("prefilter_edits", C.c_int32), ("n_pieces", C.c_int32), ("piece_start", C.c_int32 * MG_MAX_PIECES),
                ("piece_len", C.c_int32 * MG_MAX_PIECES)]


_lib = None


def lib():
    """The loaded shared library (built by `python -m minorg_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MinorgB200Error("%s is missing: build it with `python -m minorg_b200.build` (needs nvcc). "
                              "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64p, i32p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    L.mg_last_error.restype = C.c_char_p
    L.mg_init.argtypes = [C.c_int, C.c_char_p, C.c_int, i32p]
    L.mg_genome_from_ascii.argtypes = [vp, i64p, C.c_int, vp, C.POINTER(vp)]
    L.mg_genome_from_device_ascii.argtypes = [vp, i64p, C.c_int, vp, C.POINTER(vp)]
    L.mg_layout_global_length.argtypes = [i64p, C.c_int]
    L.mg_layout_global_length.restype = C.c_int64
    L.mg_genome_from_ascii_chunk.argtypes = [vp, i64p, C.c_int, C.c_int64, C.c_int64, C.c_int64, vp, C.POINTER(vp)]
    L.mg_genome_from_device_ascii_chunk.argtypes = [vp, i64p, C.c_int, C.c_int64, C.c_int64, C.c_int64, vp, C.POINTER(vp)]
    L.mg_genome_owned_range.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.mg_genome_n_contigs.argtypes = [vp]
    L.mg_genome_contig_length.argtypes = [vp, C.c_int]
    L.mg_genome_contig_length.restype = C.c_int64
    L.mg_genome_total_bases.argtypes = [vp]
    L.mg_genome_total_bases.restype = C.c_int64
    L.mg_genome_device_bytes.argtypes = [vp]
    L.mg_genome_device_bytes.restype = C.c_int64
    L.mg_genome_contig_global_start.argtypes = [vp, C.c_int]
    L.mg_genome_contig_global_start.restype = C.c_int64
    L.mg_genome_global_length.argtypes = [vp]
    L.mg_genome_global_length.restype = C.c_int64
    L.mg_genome_decode.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, vp]
    L.mg_genome_free.argtypes = [vp]
    L.mg_genome_free.restype = None
    L.mg_genome_set_masks.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    L.mg_mask_find.argtypes = [vp, vp, i64p, C.c_int, C.c_int64, vp, vp, vp, vp, vp, i64p, vp]
    L.mg_enumerate_sites.argtypes = [vp, C.c_int64, C.POINTER(Pam), C.c_int, C.c_int64, vp, i64p, vp, i64p, vp]
    L.mg_queries_create.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.mg_queries_count.argtypes = [vp]
    L.mg_queries_length.argtypes = [vp]
    L.mg_queries_free.argtypes = [vp]
    L.mg_queries_free.restype = None
    L.mg_screen.argtypes = [vp, vp, C.POINTER(Criteria), C.POINTER(Pam), C.c_int64, C.c_int64, vp, vp]
    L.mg_screen_to_host.argtypes = [vp, vp, C.POINTER(Criteria), C.POINTER(Pam), C.c_int64, C.c_int64, vp, vp]
    L.mg_screen_host.argtypes = [vp, i64p, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.POINTER(Criteria),
                                 C.POINTER(Pam), vp, vp]
    L.mg_microbench.argtypes = [C.c_int, C.POINTER(C.c_double), vp]
    L.mg_last_screen_stats.argtypes = [C.POINTER(C.c_uint64)]
    L.mg_last_seedjoin_stats.argtypes = [C.POINTER(C.c_uint64)]
    L.mg_fasta_read.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.mg_fasta_n_records.argtypes = [vp]
    L.mg_fasta_name.argtypes = [vp, C.c_int]
    L.mg_fasta_name.restype = C.c_char_p
    L.mg_fasta_offsets.argtypes = [vp]
    L.mg_fasta_offsets.restype = i64p
    L.mg_fasta_sequence.argtypes = [vp]
    L.mg_fasta_sequence.restype = vp
    L.mg_fasta_free.argtypes = [vp]
    L.mg_fasta_free.restype = None
    L.mg_genome_from_fasta.argtypes = [C.c_char_p, C.c_int, vp, C.POINTER(vp), C.POINTER(vp)]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().mg_last_error()
        raise MinorgB200Error("%s failed (%d): %s" % (what or "libminorg_b200 call", rc, (msg or b"").decode("utf-8", "replace")))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def read_fasta_native(path, threads=0):
    """Native FASTA reader (mg_fasta_read; host threads, no GPU involved) -> (uint8 array of the concatenated sequence
    bytes, int64 offsets[n+1], names): same result as minorg_b200.fasta.read_fasta_arrays."""
    f = C.c_void_p()
    check(lib().mg_fasta_read(os.fsencode(path), int(threads), C.byref(f)), "mg_fasta_read")
    try:
        n = lib().mg_fasta_n_records(f)
        offsets = np.ctypeslib.as_array(lib().mg_fasta_offsets(f), shape=(n + 1,)).copy()
        names = [lib().mg_fasta_name(f, i).decode("latin-1") for i in range(n)]
        total = int(offsets[-1])
        if total:
            buf = (C.c_uint8 * total).from_address(lib().mg_fasta_sequence(f))
            blob = np.frombuffer(buf, dtype=np.uint8).copy()
        else:
            blob = np.zeros(1, dtype=np.uint8)
    finally:
        lib().mg_fasta_free(f)
    return blob, offsets, names


def layout_global_length(offsets):
    """Global length (bases, incl. padding) of the device layout of contigs with these offsets - what
    distributed.shard_windows needs to cut a subject into chunks before any rank has packed anything."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    r = int(lib().mg_layout_global_length(_i64(offsets), len(offsets) - 1))
    if r < 0:
        raise MinorgB200Error("invalid contig offsets")
    return r


def init(device=0):
    name = C.create_string_buffer(256)
    sms = C.c_int32(0)
    check(lib().mg_init(int(device), name, 256, C.byref(sms)), "mg_init")
    return name.value.decode(), int(sms.value)


def device_count():
    return int(lib().mg_device_count())


class Genome:
    """2-bit packed genome resident in HBM (handle owned by the library)."""

    def __init__(self, handle, names):
        self._h = C.c_void_p(handle)
        self.names = list(names)

    @classmethod
    def from_contigs(cls, contigs, stream=0):
        """contigs: [(name, sequence as str/bytes), ...]"""
        names = [c[0] for c in contigs]
        bufs = [c[1].encode("ascii", "replace") if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
        offsets = np.zeros(len(bufs) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in bufs], out=offsets[1:])
        blob = b"".join(bufs)
        arr = np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(1, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().mg_genome_from_ascii(_ptr(arr), _i64(offsets), len(bufs), C.c_void_p(stream), C.byref(h)),
              "mg_genome_from_ascii")
        return cls(h.value, names)

    @classmethod
    def from_host_array(cls, arr, offsets, names, stream=0):
        """arr: uint8 numpy array (ASCII, e.g. pinned), offsets: int64[n+1]."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        check(lib().mg_genome_from_ascii(_ptr(arr), _i64(offsets), len(offsets) - 1, C.c_void_p(stream), C.byref(h)),
              "mg_genome_from_ascii")
        return cls(h.value, names)

    @classmethod
    def from_host_array_chunk(cls, arr, offsets, names, chunk_begin, chunk_end, halo=256, stream=0):
        """One rank's chunk of a subject (mg_genome_from_ascii_chunk): arr/offsets describe the WHOLE subject, only the
        bases of global positions [chunk_begin - halo, chunk_end + halo) cross PCIe; coordinates stay global."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        check(lib().mg_genome_from_ascii_chunk(_ptr(arr), _i64(offsets), len(offsets) - 1, int(chunk_begin), int(chunk_end),
                                               int(halo), C.c_void_p(stream), C.byref(h)), "mg_genome_from_ascii_chunk")
        return cls(h.value, names)

    @classmethod
    def from_device_ascii_chunk(cls, dptr, offsets, names, chunk_begin, chunk_end, halo=256, stream=0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        check(lib().mg_genome_from_device_ascii_chunk(C.c_void_p(dptr), _i64(offsets), len(offsets) - 1, int(chunk_begin),
                                                      int(chunk_end), int(halo), C.c_void_p(stream), C.byref(h)),
              "mg_genome_from_device_ascii_chunk")
        return cls(h.value, names)

    @property
    def owned_range(self):
        """(own_begin, own_end, stored_begin, stored_end) in global coordinates."""
        v = [C.c_int64(0) for _ in range(4)]
        check(lib().mg_genome_owned_range(self.handle, *[C.byref(x) for x in v]), "mg_genome_owned_range")
        return tuple(int(x.value) for x in v)

    @classmethod
    def from_fasta(cls, path, threads=0, stream=0):
        """Plain FASTA file -> resident genome through the native reader (mg_genome_from_fasta); names from the file."""
        h, f = C.c_void_p(), C.c_void_p()
        check(lib().mg_genome_from_fasta(os.fsencode(path), int(threads), C.c_void_p(stream), C.byref(h), C.byref(f)),
              "mg_genome_from_fasta")
        try:
            names = [lib().mg_fasta_name(f, i).decode("latin-1") for i in range(lib().mg_fasta_n_records(f))]
        finally:
            lib().mg_fasta_free(f)
        return cls(h.value, names)

    @classmethod
    def from_device_ascii(cls, dptr, offsets, names, stream=0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        check(lib().mg_genome_from_device_ascii(C.c_void_p(dptr), _i64(offsets), len(offsets) - 1, C.c_void_p(stream),
                                                C.byref(h)), "mg_genome_from_device_ascii")
        return cls(h.value, names)

    @property
    def handle(self):
        if not self._h:
            raise MinorgB200Error("genome handle already freed")
        return self._h

    @property
    def n_contigs(self):
        return int(lib().mg_genome_n_contigs(self.handle))

    def contig_length(self, i):
        return int(lib().mg_genome_contig_length(self.handle, i))

    @property
    def total_bases(self):
        return int(lib().mg_genome_total_bases(self.handle))

    @property
    def device_bytes(self):
        return int(lib().mg_genome_device_bytes(self.handle))

    @property
    def global_length(self):
        return int(lib().mg_genome_global_length(self.handle))

    def contig_global_start(self, i):
        return int(lib().mg_genome_contig_global_start(self.handle, i))

    def decode(self, contig, start, end):
        out = np.zeros(max(1, end - start), dtype=np.uint8)
        check(lib().mg_genome_decode(self.handle, contig, start, end, _ptr(out)), "mg_genome_decode")
        return out[:end - start].tobytes().decode()

    def set_masks(self, masks, stream=0):
        """masks: iterable of (contig index, start, end), 0-based end-exclusive."""
        masks = list(masks)
        c = np.array([m[0] for m in masks], dtype=np.int32)
        s = np.array([m[1] for m in masks], dtype=np.int64)
        e = np.array([m[2] for m in masks], dtype=np.int64)
        check(lib().mg_genome_set_masks(self.handle, _ptr(c), _ptr(s), _ptr(e), len(masks), C.c_void_p(stream)),
              "mg_genome_set_masks")

    def find_exact(self, seqs, stream=0):
        """All end-to-end exact occurrences (either strand) of each sequence: -> list of
        (sequence index, contig index, start, end, strand)."""
        bufs = [s.encode("ascii", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
        offsets = np.zeros(len(bufs) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in bufs], out=offsets[1:])
        blob = np.frombuffer(b"".join(bufs) or b"\0", dtype=np.uint8)
        cap = 1 << 16
        while True:
            o_seq = np.zeros(cap, dtype=np.int32)
            o_ctg = np.zeros(cap, dtype=np.int32)
            o_s = np.zeros(cap, dtype=np.int64)
            o_e = np.zeros(cap, dtype=np.int64)
            o_str = np.zeros(cap, dtype=np.int8)
            n = C.c_int64(0)
            rc = lib().mg_mask_find(self.handle, _ptr(blob), _i64(offsets), len(bufs), cap, _ptr(o_seq), _ptr(o_ctg),
                                    _ptr(o_s), _ptr(o_e), _ptr(o_str), C.byref(n), C.c_void_p(stream))
            if rc == MG_ERR_OVERFLOW:
                cap = int(n.value) + 16
                continue
            check(rc, "mg_mask_find")
            k = int(n.value)
            return list(zip(o_seq[:k].tolist(), o_ctg[:k].tolist(), o_s[:k].tolist(), o_e[:k].tolist(), o_str[:k].tolist()))

    def free(self):
        if self._h:
            lib().mg_genome_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Queries:
    """gRNA set (both orientations + bit-sliced tables) resident in HBM."""

    def __init__(self, seqs, stream=0):
        seqs = [s.encode("ascii", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
        self.n = len(seqs)
        self.length = len(seqs[0]) if seqs else 1
        if any(len(s) != self.length for s in seqs):
            raise MinorgB200Error("all gRNA of one screen must have the same length")
        blob = np.frombuffer(b"".join(seqs) or b"\0", dtype=np.uint8)
        h = C.c_void_p()
        check(lib().mg_queries_create(_ptr(blob), self.n, self.length, C.c_void_p(stream), C.byref(h)), "mg_queries_create")
        self._h = h

    @property
    def handle(self):
        return self._h

    def free(self):
        if self._h:
            lib().mg_queries_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def screen(genome, queries, criteria, pam, d_counts_ptr, win_begin=0, win_end=-1, stream=0):
    """Adds per-query problematic-alignment counts into the DEVICE array d_counts_ptr (2N uint32)."""
    check(lib().mg_screen(genome.handle, queries.handle, C.byref(criteria), C.byref(pam) if pam is not None else None,
                          int(win_begin), int(win_end), C.c_void_p(d_counts_ptr), C.c_void_p(stream)), "mg_screen")


def screen_counts(genome, queries, criteria, pam, win_begin=0, win_end=-1, stream=0):
    """mg_screen_to_host: per-query counts of one screen as a host uint32[2N] array (no device-array package needed)."""
    out = np.zeros(max(1, 2 * queries.n), dtype=np.uint32)
    check(lib().mg_screen_to_host(genome.handle, queries.handle, C.byref(criteria), C.byref(pam) if pam is not None else None,
                                  int(win_begin), int(win_end), _ptr(out), C.c_void_p(stream)), "mg_screen_to_host")
    return out[:2 * queries.n]


def screen_host(arr, offsets, masks, grna_blob, n, length, criteria, pam, stream=0):
    """Host-buffer end-to-end call (H2D + pack + screen + D2H): returns uint32[n] counts."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    masks = list(masks)
    mc = np.array([m[0] for m in masks], dtype=np.int32)
    ms = np.array([m[1] for m in masks], dtype=np.int64)
    me = np.array([m[2] for m in masks], dtype=np.int64)
    out = np.zeros(max(1, n), dtype=np.uint32)
    check(lib().mg_screen_host(_ptr(arr), _i64(offsets), len(offsets) - 1, _ptr(mc), _ptr(ms), _ptr(me), len(masks),
                               _ptr(grna_blob), n, length, C.byref(criteria), C.byref(pam) if pam is not None else None,
                               _ptr(out), C.c_void_p(stream)), "mg_screen_host")
    return out[:n]


def enumerate_sites(seq, pam, grna_len, stream=0):
    """K2: ascending PAM-site starts on the '+' strand and on the reverse-complemented record."""
    buf = seq.encode("ascii", "replace") if isinstance(seq, str) else bytes(seq)
    arr = np.frombuffer(buf or b"\0", dtype=np.uint8)
    cap = max(16, 2 * len(buf) // 4 + 16)
    while True:
        plus = np.zeros(cap, dtype=np.int64)
        minus = np.zeros(cap, dtype=np.int64)
        n_p, n_m = C.c_int64(0), C.c_int64(0)
        rc = lib().mg_enumerate_sites(_ptr(arr), len(buf), C.byref(pam), int(grna_len), cap, _ptr(plus), C.byref(n_p),
                                      _ptr(minus), C.byref(n_m), C.c_void_p(stream))
        if rc == MG_ERR_OVERFLOW:
            cap = max(int(n_p.value), int(n_m.value)) + 16
            continue
        check(rc, "mg_enumerate_sites")
        return plus[:n_p.value].copy(), minus[:n_m.value].copy()


def microbench(which, stream=0):
    v = C.c_double(0)
    check(lib().mg_microbench(int(which), C.byref(v), C.c_void_p(stream)), "mg_microbench")
    return float(v.value)


def last_screen_stats():
    """(steps, second-stage steps) of the most recent bit-sliced scan launch (see mg_last_screen_stats)."""
    out = (C.c_uint64 * 2)()
    check(lib().mg_last_screen_stats(out), "mg_last_screen_stats")
    return int(out[0]), int(out[1])


def last_seedjoin_stats():
    """(candidates compared with the text, candidates that reached the exact evaluation) of the most recent seed-and-join
    screen on this device (see mg_last_seedjoin_stats); (0, 0) if that path did not run."""
    out = (C.c_uint64 * 2)()
    check(lib().mg_last_seedjoin_stats(out), "mg_last_seedjoin_stats")
    return int(out[0]), int(out[1])
