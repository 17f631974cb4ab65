"""ctypes binding of libdynaalign_b200.so (the C ABI declared in include/dynaalign_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C dynaalign_b200/csrc``.
There is no Python or CPU fallback: if the library is missing, loading fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynaalign_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
MH_SIMILARITY, MH_DISTANCE = 0, 1


class DynaAlignError(RuntimeError):
    """Raised for every non-zero return code; ``str(e)`` is the library's message (for argument errors,
    the reference's Rcpp::stop text)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_lib = None

_u8p, _i8p = C.POINTER(C.c_uint8), C.POINTER(C.c_int8)
_u16p, _u32p = C.POINTER(C.c_uint16), C.POINTER(C.c_uint32)
_i32p, _i64p, _f64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

_SIGS = {
    "dyna_last_error": (C.c_char_p, []),
    "dyna_version": (C.c_int, []),
    "dyna_device_count": (C.c_int, []),
    "dyna_set_device": (C.c_int, [C.c_int]),
    "dyna_release_cached_memory": (C.c_int, [C.c_int]),
    "dyna_partition_rows": (C.c_int, [C.c_int64, _i64p, C.c_int, C.c_int, _i64p]),
    "dyna_hashfamily_seeds": (C.c_int, [C.c_uint32, C.c_int, _u32p]),
    "dyna_random_seed": (C.c_uint32, []),
    "dyna_mh_signatures_murmur3": (C.c_int, [_u8p, _i64p, C.c_int64, C.c_int, _u32p, C.c_int, _u32p]),
    "dyna_mh_match_counts": (C.c_int, [_u32p, C.c_int64, C.c_int, C.c_int64, C.c_int64, _u16p]),
    "dyna_mh_match_matrix": (C.c_int, [_u32p, C.c_int64, C.c_int, C.c_int, _f64p, C.c_int]),
    "dyna_similarityMH": (C.c_int, [_u8p, _i64p, C.c_int64, C.c_int, C.c_int, _u32p, _f64p, C.c_int]),
    "dyna_mh_signatures_linear": (C.c_int, [_i32p, _i64p, C.c_int64, _i64p, _i64p, C.c_int64, C.c_int, _u32p]),
    "dyna_minhash_vocab_ranks": (C.c_int, [_u8p, _i64p, C.c_int64, C.c_int, C.POINTER(C.c_uint64), C.c_int64, _i64p, _i32p, _i64p]),
    "dyna_substitution_matrix": (C.c_int, [C.c_char_p, _i8p]),
    "dyna_aa_index_table": (None, [_i8p]),
    "dyna_nw_pair_stats": (C.c_int, [_u8p, _i64p, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _u32p]),
    "dyna_nw_pair_stats8": (C.c_int, [_u8p, _i64p, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, _u8p, _u8p]),
    "dyna_similarityNW": (C.c_int, [_u8p, _i64p, C.c_int64, C.c_char_p, C.c_int, C.c_int, _f64p, C.c_int]),
    "dyna_mh_plan_create": (C.c_void_p, [C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int]),
    "dyna_mh_plan_upload_sequences": (C.c_int, [C.c_void_p, _u8p, _i64p, C.c_int, _u32p, C.c_void_p]),
    "dyna_mh_plan_upload_signatures": (C.c_int, [C.c_void_p, _u32p, C.c_void_p]),
    "dyna_mh_plan_run_signatures": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dyna_mh_plan_run_match": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dyna_mh_plan_run_signatures_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "dyna_mh_plan_code_rows": (C.c_int, [C.c_void_p]),
    "dyna_mh_plan_code_row_bytes": (C.c_int64, [C.c_void_p]),
    "dyna_mh_plan_codes_device_ptr": (C.c_void_p, [C.c_void_p]),
    "dyna_mh_plan_overflow_device_ptr": (C.c_void_p, [C.c_void_p]),
    "dyna_mh_plan_run_match_fetch": (C.c_int, [C.c_void_p, _u16p, C.c_void_p]),
    "dyna_mh_plan_fetch_signatures": (C.c_int, [C.c_void_p, _u32p, C.c_void_p]),
    "dyna_mh_plan_fetch_counts": (C.c_int, [C.c_void_p, _u16p, C.c_void_p]),
    "dyna_mh_plan_create_subset": (C.c_void_p, [C.c_void_p, _i64p, C.c_int64, C.c_int64, C.c_int64]),
    "dyna_mh_plan_count_histogram": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "dyna_quantile_type7_counts": (C.c_int, [C.POINTER(C.c_uint64), C.c_int, C.c_double, _f64p, C.POINTER(C.c_int)]),
    "dyna_mh_plan_threshold_edges": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _i32p, _i32p, _u16p, _i64p, C.c_void_p]),
    "dyna_mh_plan_run_match_fetch8": (C.c_int, [C.c_void_p, _u8p, C.c_int64, _i64p, _u16p, _i64p, C.c_void_p]),
    "dyna_mh_plan_run_match_sparse": (C.c_int, [C.c_void_p, C.c_int64, _i64p, C.POINTER(C.c_int), C.c_void_p]),
    "dyna_mh_plan_checksum": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "dyna_mh_plan_pairs": (C.c_int64, [C.c_void_p]),
    "dyna_mh_plan_launches": (C.c_int, [C.c_void_p]),
    "dyna_mh_plan_counts_device_ptr": (C.c_void_p, [C.c_void_p]),
    "dyna_mh_plan_destroy": (None, [C.c_void_p]),
    "dyna_nw_plan_create": (C.c_void_p, [_u8p, _i64p, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int]),
    "dyna_nw_plan_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dyna_nw_plan_fetch": (C.c_int, [C.c_void_p, _u32p, _u32p, C.c_void_p]),
    "dyna_nw_plan_fetch_packed8": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_void_p]),
    "dyna_nw_plan_checksum": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "dyna_nw_plan_max_len": (C.c_int, [C.c_void_p]),
    "dyna_nw_plan_stat_histogram": (C.c_int, [C.c_void_p, _i32p, C.c_int64, C.POINTER(C.c_uint64), C.c_void_p]),
    "dyna_nw_plan_fetch_diagonal": (C.c_int, [C.c_void_p, _i32p, C.c_int64, _u32p, _u32p, C.c_void_p]),
    "dyna_quantile_type7_identities": (C.c_int, [C.POINTER(C.c_uint64), C.c_int64, C.c_int64, C.c_double, _f64p]),
    "dyna_nw_plan_threshold_edges": (C.c_int, [C.c_void_p, _i32p, C.c_int64, C.c_double, C.c_int64, _i32p, _i32p, _u32p, _u32p,
                                              _i64p, C.c_void_p]),
    "dyna_nw_plan_layout": (C.c_void_p, [_u8p, _i64p, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64]),
    "dyna_nw_plan_unit_count": (C.c_int64, [C.c_void_p]),
    "dyna_nw_plan_export_units": (C.c_int, [C.c_void_p, _i32p]),
    "dyna_nw_plan_pairs": (C.c_int64, [C.c_void_p]),
    "dyna_nw_plan_cells": (C.c_int64, [C.c_void_p]),
    "dyna_nw_plan_launches": (C.c_int, [C.c_void_p]),
    "dyna_nw_plan_destroy": (None, [C.c_void_p]),
    "dyna_probe_int_issue": (C.c_int, [C.c_int, _f64p, _f64p, C.c_void_p]),
}

EXPORTS = tuple(sorted(_SIGS))


def lib():
    """The loaded library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "dynaalign_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C dynaalign_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().dyna_last_error().decode("latin-1")


def check(rc):
    if rc != OK:
        raise DynaAlignError(rc, last_error())


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def flatten(sequences):
    """list[str|bytes] -> (uint8 residues, int64 offsets[n+1]).  str is encoded latin-1 (raw bytes, as the
    reference hashes / compares raw chars)."""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in sequences]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offsets[1:])
    total = int(offsets[-1])
    residues = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if total else np.zeros(1, np.uint8)
    return residues, offsets
