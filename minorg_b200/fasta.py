"""FASTA reading/writing on the host with the semantics of the Biopython calls the reference makes
(Bio.SeqIO.parse/write 'fasta', fasta.py:11-43): id = first header token, case preserved, 60-column output."""
import numpy as np

_COMPLEMENT = bytes.maketrans(b"ACGTMRWSYKVHDBNXacgtmrwsykvhdbnx", b"TGCAKYWSRMBDHVNXtgcakywsrmbdhvnx")


def reverse_complement(seq):
    """IUPAC-aware, case preserving (Bio.Seq.reverse_complement); str in -> str out, bytes in -> bytes out."""
    if isinstance(seq, str):
        return seq.encode("latin-1").translate(_COMPLEMENT)[::-1].decode("latin-1")
    return bytes(seq).translate(_COMPLEMENT)[::-1]


def _open(path):
    """Binary handle on a FASTA file, transparently inflating gzip (magic 1f 8b)."""
    fh = open(path, "rb")
    magic = fh.read(2)
    fh.seek(0)
    if magic == b"\x1f\x8b":
        import gzip
        return gzip.open(fh, "rb")
    return fh


def read_fasta(path):
    """-> [(id, sequence str)] in file order (duplicate ids are kept, like SeqIO.parse)."""
    records, name, parts = [], None, []
    with _open(path) as fh:
        for raw in fh:
            if raw[:1] == b">":
                if name is not None:
                    records.append((name, b"".join(parts).decode("latin-1")))
                fields = raw[1:].split()
                name, parts = (fields[0].decode("latin-1") if fields else ""), []
            elif name is not None:
                parts.append(b"".join(raw.split()))
    if name is not None:
        records.append((name, b"".join(parts).decode("latin-1")))
    return records


def fasta_to_dict(path):
    """fasta.py:11-19: later duplicates overwrite earlier ones."""
    return dict(read_fasta(path))


def dict_to_fasta(d, path):
    with open(path, "w") as fh:
        for name, seq in d.items():
            seq = str(seq)
            fh.write(">%s\n" % name)
            for i in range(0, len(seq), 60):
                fh.write(seq[i:i + 60] + "\n")


def concat_records(records):
    """[(id, seq)] -> (uint8 array of the concatenated ASCII, int64 offsets[n+1], names)."""
    bufs = [s.encode("latin-1") for _, s in records]
    offsets = np.zeros(len(bufs) + 1, dtype=np.int64)
    if bufs:
        np.cumsum([len(b) for b in bufs], out=offsets[1:])
    blob = np.frombuffer(b"".join(bufs) or b"\0", dtype=np.uint8)
    return blob, offsets, [n for n, _ in records]


def read_fasta_arrays(path):
    """Whole-genome reader for the device path: -> (uint8 array of the concatenated sequence bytes, int64
    offsets[n+1], names).  Same record semantics as read_fasta (id = first header token, whitespace removed,
    case preserved) but vectorised with numpy: a 1 Gbp FASTA never becomes a Python str."""
    with _open(path) as fh:
        raw = np.frombuffer(fh.read(), dtype=np.uint8)
    if raw.size == 0:
        return np.zeros(1, dtype=np.uint8), np.zeros(1, dtype=np.int64), []
    line_starts = np.concatenate(([0], np.flatnonzero(raw == 10) + 1))
    line_starts = line_starts[line_starts < raw.size]
    header_lines = line_starts[raw[line_starts] == ord(">")]
    names, pieces, lengths = [], [], []
    ends = np.concatenate((header_lines[1:], [raw.size]))
    for h, nxt in zip(header_lines.tolist(), ends.tolist()):
        eol = h + int(np.argmax(raw[h:nxt] == 10)) if (raw[h:nxt] == 10).any() else nxt
        fields = raw[h + 1:eol].tobytes().split()
        names.append(fields[0].decode("latin-1") if fields else "")
        body = raw[eol:nxt]
        seq = body[(body != 10) & (body != 13) & (body != 32) & (body != 9)]
        pieces.append(seq)
        lengths.append(seq.size)
    offsets = np.zeros(len(names) + 1, dtype=np.int64)
    if names:
        np.cumsum(lengths, out=offsets[1:])
    blob = np.concatenate(pieces) if pieces and offsets[-1] > 0 else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(blob), offsets, names
