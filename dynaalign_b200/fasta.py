"""Minimal FASTA reader for the sequence vectors the similarity entry points take.

The reference ships its examples as data/*.rda (see rda.py) and has no FASTA code of its own; this is the other
common on-disk form of the same input.  Plain or gzip-compressed files, multi-line records, '*' terminators and
blanks removed, residues upper-cased (the NW alphabet `ARNDCQEGHILKMFPSTWYVBZX*` is upper case; lower-case residues
would be rejected by similarityNW exactly as the reference rejects them).
"""
import gzip


def read_fasta(path, upper=True, strip_terminator=True):
    """-> (names, sequences) in file order."""
    opener = gzip.open if str(path).endswith(".gz") else open
    names, seqs, cur = [], [], None
    with opener(path, "rt") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith(";"):
                continue
            if line.startswith(">"):
                if cur is not None:
                    seqs.append("".join(cur))
                names.append(line[1:].split()[0] if len(line) > 1 else "")
                cur = []
            elif cur is None:
                raise ValueError("FASTA: sequence data before the first '>' header")
            else:
                cur.append("".join(line.split()))
    if cur is not None:
        seqs.append("".join(cur))
    out = []
    for s in seqs:
        if upper:
            s = s.upper()
        if strip_terminator and s.endswith("*"):
            s = s[:-1]
        out.append(s)
    return names, out
