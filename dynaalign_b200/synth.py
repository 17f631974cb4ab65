"""Deterministic synthetic inputs for the BASELINE.json configurations.

C4: N peptides of length 16, i.i.d. uniform over the 20 standard residues (PCG64 seed 20240615),
    or the *clustered* variant (parents x children with 2 substitutions, seed 20240616).
C5: N proteins, lengths clip(round(Normal(330,10)),300,360) (seed 20240617); families of
    ``family`` children per parent with 5% substitutions and 1% single-residue indels (seed 20240618),
    or a pure-uniform variant.
All generators return ``list[bytes]``.
"""
import numpy as np

RESIDUES20 = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)


def peptides_uniform(n=100_000, length=16, seed=20240615):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 20, size=(n, length))
    arr = RESIDUES20[codes]
    return [arr[i].tobytes() for i in range(n)]


def peptides_clustered(n=100_000, length=16, children=100, n_sub=2, seed=20240616):
    rng = np.random.default_rng(seed)
    parents = (n + children - 1) // children
    p = RESIDUES20[rng.integers(0, 20, size=(parents, length))]
    arr = np.repeat(p, children, axis=0)[:n].copy()
    for s in range(n_sub):
        pos = rng.integers(0, length, size=n)
        arr[np.arange(n), pos] = RESIDUES20[rng.integers(0, 20, size=n)]
    return [arr[i].tobytes() for i in range(n)]


def protein_lengths(n=20_000, mean=330.0, sd=10.0, lo=300, hi=360, seed=20240617):
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(mean, sd, size=n)), lo, hi).astype(np.int64)


def proteins_uniform(n=20_000, seed=20240618, **kw):
    lens = protein_lengths(n, **kw)
    rng = np.random.default_rng(seed)
    return [RESIDUES20[rng.integers(0, 20, size=int(L))].tobytes() for L in lens]


def proteins_families(n=20_000, family=100, p_sub=0.05, p_indel=0.01, seed=20240618, **kw):
    """HA1-like families: each parent is uniform random at the target length; each child copies the
    parent with per-residue substitution / insertion / deletion, then is trimmed or padded with
    random residues to its own target length."""
    lens = protein_lengths(n, **kw)
    rng = np.random.default_rng(seed)
    out = []
    parent = None
    for i in range(n):
        L = int(lens[i])
        if i % family == 0:
            parent = RESIDUES20[rng.integers(0, 20, size=int(lens[i]) + 8)]
        r = rng.random(parent.shape[0])
        child = []
        for pos in range(parent.shape[0]):
            x = r[pos]
            if x < p_indel / 2:
                continue                                       # deletion
            if x < p_indel:
                child.append(RESIDUES20[rng.integers(0, 20)])  # insertion before this residue
            if x > 1.0 - p_sub:
                child.append(RESIDUES20[rng.integers(0, 20)])  # substitution
            else:
                child.append(parent[pos])
        child = np.asarray(child, dtype=np.uint8)
        if child.shape[0] >= L:
            child = child[:L]
        else:
            child = np.concatenate([child, RESIDUES20[rng.integers(0, 20, size=L - child.shape[0])]])
        out.append(child.tobytes())
    return out
