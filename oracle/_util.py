"""TEST INFRASTRUCTURE ONLY: helpers shared by the oracle bindings."""
import ctypes as C

import numpy as np


def flatten(sequences):
    """list[str|bytes] -> (uint8 residues, int64 offsets[n+1]); the layout the C ABIs take."""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in sequences]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    residues = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if offsets[-1] else np.zeros(1, np.uint8)
    return residues, offsets


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def fnv1a64(buf: bytes) -> int:
    """FNV-1a 64 over raw bytes (the fingerprint used in SURVEY.md Appendix C)."""
    h = 1469598103934665603
    for chunk_start in range(0, len(buf), 1 << 20):
        for b in buf[chunk_start:chunk_start + (1 << 20)]:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h
