"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/libdynaoracle.so (oracle/dyna_oracle.c).

The plain-C restatement of the reference's algorithms.  It travels to the GPU box (built by
``__graft_entry__.build()``) and is what ``-m gpu`` parity tests, ``smoke()`` and ``bench.py``'s
``cpu_baseline`` compare against / time.  Never imported by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ._util import flatten, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libdynaoracle.so")
_lib = None


class OracleError(RuntimeError):
    pass


def build(force=False):
    if force or not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(os.path.join(_HERE, "dyna_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_murmur3_32.restype = C.c_uint32
        L.orc_murmur3_32.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def murmur3_32(key: bytes, seed: int) -> int:
    return int(lib().orc_murmur3_32(key, C.c_uint64(len(key)), C.c_uint32(seed)))


def hashfamily_seeds(seed: int, n_hash: int) -> np.ndarray:
    out = np.zeros(n_hash, dtype=np.uint32)
    lib().orc_hashfamily_seeds(C.c_uint32(seed), C.c_int(n_hash), ptr(out, C.c_uint32))
    return out


def mh_signatures(sequences, k, seeds) -> np.ndarray:
    res, off = flatten(sequences)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    n, n_hash = len(sequences), len(seeds)
    out = np.zeros((n, n_hash), dtype=np.uint32)
    _check(lib().orc_mh_signatures(ptr(res, C.c_uint8), ptr(off, C.c_int64), C.c_int64(n), C.c_int(k),
                                   ptr(seeds, C.c_uint32), C.c_int(n_hash), ptr(out, C.c_uint32)))
    return out


def tri_strict_size(n, row_begin=0, row_end=None):
    row_end = n if row_end is None else row_end
    f = lambda r: r * n - r * (r + 1) // 2
    return f(row_end) - f(row_begin)


def tri_diag_size(n, row_begin=0, row_end=None):
    row_end = n if row_end is None else row_end
    f = lambda r: r * n - r * (r - 1) // 2
    return f(row_end) - f(row_begin)


def mh_match_counts(sig, row_begin=0, row_end=None) -> np.ndarray:
    sig = np.ascontiguousarray(sig, dtype=np.uint32)
    n, n_hash = sig.shape
    row_end = n if row_end is None else row_end
    out = np.zeros(max(tri_strict_size(n, row_begin, row_end), 1), dtype=np.uint16)
    lib().orc_mh_match_counts(ptr(sig, C.c_uint32), C.c_int64(n), C.c_int(n_hash), C.c_int64(row_begin),
                              C.c_int64(row_end), ptr(out, C.c_uint16))
    return out[:tri_strict_size(n, row_begin, row_end)]


def similarityMH(sequences, k=4, n_hash=50, seed=42) -> np.ndarray:
    res, off = flatten(sequences)
    n = len(sequences)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    _check(lib().orc_similarityMH(ptr(res, C.c_uint8), ptr(off, C.c_int64), C.c_int64(n), C.c_int(k), C.c_int(n_hash),
                                  C.c_uint32(seed), ptr(out, C.c_double)))
    return out


def substitution_matrix(name) -> np.ndarray:
    out = np.zeros((24, 24), dtype=np.int8)
    _check(lib().orc_substitution_matrix(name.encode(), ptr(out, C.c_int8)))
    return out


def _as_bytes(s):
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def nw_pair(a, b, sub=None, matrixName="BLOSUM62", gapOpen=10, gapExt=4, forward=False):
    """(matches, alignment_length) for the ordered pair (a on rows)."""
    if sub is None:
        sub = substitution_matrix(matrixName)
    sub = np.ascontiguousarray(sub, dtype=np.int8)
    ab, bb = _as_bytes(a), _as_bytes(b)
    mt, ln = C.c_int32(0), C.c_int32(0)
    fn = lib().orc_nw_pair_forward if forward else lib().orc_nw_pair
    _check(fn(ab, C.c_int64(len(ab)), bb, C.c_int64(len(bb)), ptr(sub, C.c_int8), C.c_int(gapOpen), C.c_int(gapExt),
              C.byref(mt), C.byref(ln)))
    return mt.value, ln.value


def nw_pair_stats(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4, row_begin=0, row_end=None):
    """(matches uint32[], length uint32[]) over the packed upper triangle incl. diagonal, rows [row_begin,row_end)."""
    res, off = flatten(sequences)
    n = len(sequences)
    row_end = n if row_end is None else row_end
    sz = tri_diag_size(n, row_begin, row_end)
    mt = np.zeros(max(sz, 1), dtype=np.uint32)
    ln = np.zeros(max(sz, 1), dtype=np.uint32)
    _check(lib().orc_nw_pair_stats(ptr(res, C.c_uint8), ptr(off, C.c_int64), C.c_int64(n), matrixName.encode(),
                                   C.c_int(gapOpen), C.c_int(gapExt), C.c_int64(row_begin), C.c_int64(row_end),
                                   ptr(mt, C.c_uint32), ptr(ln, C.c_uint32)))
    return mt[:sz], ln[:sz]


def similarityNW(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4) -> np.ndarray:
    res, off = flatten(sequences)
    n = len(sequences)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    _check(lib().orc_similarityNW(ptr(res, C.c_uint8), ptr(off, C.c_int64), C.c_int64(n), matrixName.encode(),
                                  C.c_int(gapOpen), C.c_int(gapExt), ptr(out, C.c_double)))
    return out


def mh_signatures_linear(ranks, rank_offsets, a, b, m) -> np.ndarray:
    ranks = np.ascontiguousarray(ranks, dtype=np.int32)
    rank_offsets = np.ascontiguousarray(rank_offsets, dtype=np.int64)
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    n, n_hash = len(rank_offsets) - 1, len(a)
    out = np.zeros((n, n_hash), dtype=np.uint32)
    if len(ranks) == 0:
        ranks = np.zeros(1, np.int32)
    _check(lib().orc_mh_signatures_linear(ptr(ranks, C.c_int32), ptr(rank_offsets, C.c_int64), C.c_int64(n),
                                          ptr(a, C.c_int64), ptr(b, C.c_int64), C.c_int64(m), C.c_int(n_hash),
                                          ptr(out, C.c_uint32)))
    return out


def mh_distance_matrix(sig) -> np.ndarray:
    sig = np.ascontiguousarray(sig, dtype=np.uint32)
    n, n_hash = sig.shape
    out = np.zeros((n, n), dtype=np.float64, order="F")
    lib().orc_mh_distance_matrix(ptr(sig, C.c_uint32), C.c_int64(n), C.c_int(n_hash), ptr(out, C.c_double))
    return out
