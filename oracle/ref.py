"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libdynaref.so.

That library is the reference's src/minHash.cpp + src/pairwiseSeqAlign.cpp compiled unmodified
(see oracle/ref_wrapper.cpp, oracle/Makefile).  ``available()`` is False when it was never built
(no /root/reference and no prebuilt copy shipped).
"""
import ctypes as C
import os

import numpy as np

from ._util import flatten, ptr

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libdynaref.so")
_lib = None


class RefError(RuntimeError):
    pass


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RefError("oracle/_ref/libdynaref.so not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_murmur3_32.restype = C.c_uint32
        L.ref_murmur3_32.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32]
        L.ref_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RefError(lib().ref_last_error().decode())


def set_threads(t: int):
    lib().ref_set_threads(C.c_int(t))


def max_threads() -> int:
    return lib().ref_max_threads()


def similarityMH(sequences, k=4, n_hash=50, seed=42):
    """Reference similarityMH (src/minHash.cpp:119) with std::random_device replaced by ``seed``."""
    res, off = flatten(sequences)
    n = len(sequences)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    ok = C.c_int(0)
    _check(lib().ref_similarityMH(ptr(res, C.c_char), ptr(off, C.c_int64), C.c_int64(n), C.c_int(k), C.c_int(n_hash),
                                  C.c_uint(seed), ptr(out, C.c_double), C.byref(ok)))
    assert ok.value == 1, "reference dimnames are not 1..n"
    return out


def similarityNW(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4):
    """Reference similarityNW (src/pairwiseSeqAlign.cpp:331)."""
    res, off = flatten(sequences)
    n = len(sequences)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    ok = C.c_int(0)
    _check(lib().ref_similarityNW(ptr(res, C.c_char), ptr(off, C.c_int64), C.c_int64(n), matrixName.encode(),
                                  C.c_int(gapOpen), C.c_int(gapExt), ptr(out, C.c_double), C.byref(ok)))
    if n:
        assert ok.value == 1, "reference dimnames are not 1..n"
    return out


def calculate_similarity(a, b, matrixName="BLOSUM62", gapOpen=10, gapExt=4) -> float:
    """Reference calculate_similarity (src/pairwiseSeqAlign.cpp:209), a on rows."""
    ab = a.encode("latin-1") if isinstance(a, str) else bytes(a)
    bb = b.encode("latin-1") if isinstance(b, str) else bytes(b)
    out = C.c_double(0)
    _check(lib().ref_calculate_similarity(ab, C.c_int64(len(ab)), bb, C.c_int64(len(bb)), matrixName.encode(),
                                          C.c_int(gapOpen), C.c_int(gapExt), C.byref(out)))
    return out.value


def substitution_matrix(name) -> np.ndarray:
    out = np.zeros((24, 24), dtype=np.int32)
    _check(lib().ref_substitution_matrix(name.encode(), ptr(out, C.c_int)))
    return out


def aa_index_table() -> np.ndarray:
    out = np.zeros(256, dtype=np.int32)
    lib().ref_aa_index_table(ptr(out, C.c_int))
    return out


def murmur3_32(key: bytes, seed: int) -> int:
    return int(lib().ref_murmur3_32(key, C.c_uint64(len(key)), C.c_uint32(seed)))


def hashfamily_hash(seed: int, n_hash: int, kmer: bytes) -> np.ndarray:
    out = np.zeros(n_hash, dtype=np.uint32)
    _check(lib().ref_hashfamily_hash(C.c_uint(seed), C.c_int(n_hash), kmer, C.c_int64(len(kmer)), ptr(out, C.c_uint32)))
    return out


def mh_signatures(sequences, k, n_hash, seed=42) -> np.ndarray:
    """uint32[n, n_hash] built with the reference's HashFamily/generate_kmers (src/minHash.cpp:140-157)."""
    res, off = flatten(sequences)
    n = len(sequences)
    out = np.zeros((n, n_hash), dtype=np.uint32)
    _check(lib().ref_mh_signatures(ptr(res, C.c_char), ptr(off, C.c_int64), C.c_int64(n), C.c_int(k), C.c_int(n_hash),
                                   C.c_uint(seed), ptr(out, C.c_uint32)))
    return out
