"""FASTA reading/writing on the host with the semantics of the Biopython calls the reference makes
(Bio.SeqIO.parse/write 'fasta', fasta.py:11-43): id = first header token, case preserved, 60-column output."""
import numpy as np

_COMPLEMENT = bytes.maketrans(b"ACGTMRWSYKVHDBNXacgtmrwsykvhdbnx", b"TGCAKYWSRMBDHVNXtgcakywsrmbdhvnx")


def reverse_complement(seq):
    """IUPAC-aware, case preserving (Bio.Seq.reverse_complement); str in -> str out, bytes in -> bytes out."""
    if isinstance(seq, str):
        return seq.encode("latin-1").translate(_COMPLEMENT)[::-1].decode("latin-1")
    return bytes(seq).translate(_COMPLEMENT)[::-1]


def read_fasta(path):
    """-> [(id, sequence str)] in file order (duplicate ids are kept, like SeqIO.parse)."""
    records, name, parts = [], None, []
    with open(path, "rb") as fh:
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
