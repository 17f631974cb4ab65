"""CPU: the C-ABI library loads, exports every symbol include/dynaalign_b200.h declares, reproduces the reference's
argument errors before touching a device, and its host-side logic (seed stream, tables, partitioning) is right.
No compute entry point is exercised here (that needs a GPU and is covered by the -m gpu tests)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import dynaalign_b200 as da
from conftest import ALPHABET24, GOLDEN, ROOT, TABLES
from dynaalign_b200 import _lib
from oracle import port


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dynaalign_b200.h")).read()
    return sorted(set(re.findall(r"DYNA_API[^;(]*?\b(dyna_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 35
    L = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    assert set(syms) == set(_lib.EXPORTS), "ctypes binding and header disagree"
    assert _lib.lib().dyna_version() >= 100


def test_product_does_not_reference_the_oracle():
    # the product path must never route through the CPU checker
    pkg = os.path.join(ROOT, "dynaalign_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".R")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "dynaoracle" not in txt \
                    and "libdynaref" not in txt, os.path.join(dirpath, f)


def test_reference_error_strings_before_any_device_work(golden):
    e = golden["errors"]
    cases = [(lambda: da.similarityMH([], 4, 50), "mh_empty"), (lambda: da.similarityMH(["AAAA"], 0, 50), "mh_k0"),
             (lambda: da.similarityMH(["AAAA"], 4, 0), "mh_nhash0"), (lambda: da.similarityNW(["AA"], "BLOSUM63"), "nw_badname"),
             (lambda: da.similarityNW(["JA", "AA"]), "nw_bad_seq1"), (lambda: da.similarityNW(["AJ", "AA"]), "nw_bad_seq2_self"),
             (lambda: da.similarityNW(["AA", "AAb"]), "nw_bad_seq2_other"),
             (lambda: da.similarityNW(["", "AA", "Ao"]), "nw_empty_first_skips")]
    for fn, key in cases:
        with pytest.raises(da.DynaAlignError) as ei:
            fn()
        assert str(ei.value) == e[key] and ei.value.code == _lib.ERR_INVALID
    # name is checked before residues, as in the reference (getSubstitutionMatrix runs first)
    with pytest.raises(da.DynaAlignError, match="Invalid substitution matrix name: nope"):
        da.similarityNW(["JJ"], "nope")
    # n == 0: the reference returns a 0x0 matrix without error
    assert da.similarityNW([]).shape == (0, 0)
    with pytest.raises(da.DynaAlignError, match="Invalid substitution matrix name"):
        da.similarityNW([], "BLOSUM1")


def test_no_cpu_fallback_without_device():
    if _lib.lib().dyna_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(da.DynaAlignError) as ei:
        da.similarityNW(["AA", "AC"])
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(da.DynaAlignError) as ei:
        da.similarityMH(["AAAA", "ACAA"], 2, 8, seed=1)
    assert ei.value.code == _lib.ERR_CUDA


def test_seed_stream_matches_mt19937():
    for seed in (0, 1, 42, 12345, 2 ** 32 - 1):
        assert (da.hashfamily_seeds(seed, 700) == port.hashfamily_seeds(seed, 700)).all()


def test_tables_match_golden():
    with open(os.path.join(GOLDEN, "blosum_tables.json")) as f:
        g = json.load(f)
    for nm in TABLES:
        assert da.substitution_matrix(nm).tolist() == g["tables"][nm]
    t = np.zeros(256, dtype=np.int8)
    _lib.lib().dyna_aa_index_table(_lib.ptr(t, C.c_int8))
    for c in range(256):
        assert t[c] == (ALPHABET24.index(chr(c)) if chr(c) in ALPHABET24 else -1)


@pytest.mark.parametrize("n,shards", [(1, 1), (2, 2), (10, 3), (1000, 8), (100000, 8), (7, 16)])
def test_partition_rows_uniform(n, shards):
    b = da.partition_rows(n, shards)
    assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
    if n >= 1000:
        pairs = np.array([da.api.tri_strict_size(n, int(b[s]), int(b[s + 1])) for s in range(shards)], dtype=np.float64)
        assert pairs.max() / pairs.mean() < 1.02


def test_partition_rows_weighted_balances_cells():
    rng = np.random.default_rng(0)
    lens = rng.integers(300, 361, size=5000)
    b = da.partition_rows(5000, 8, weights=lens, include_diagonal=True)
    suffix = np.concatenate([np.cumsum(lens[::-1])[::-1], [0]])
    work = lens * suffix[:-1]
    shares = np.array([work[b[s]:b[s + 1]].sum() for s in range(8)], dtype=np.float64)
    assert shares.sum() == work.sum()
    assert shares.max() / shares.mean() < 1.02


def test_rda_reader_on_reference_datasets():
    path = "/root/reference/data/evp_peparray.rda"
    if not os.path.exists(path):
        pytest.skip("reference datasets not present")
    from dynaalign_b200.rda import load_sequences
    seqs = load_sequences(path, "PROBE_SEQUENCE")
    with open(os.path.join(GOLDEN, "evp_probe_sequences.txt")) as f:
        assert seqs == [ln.strip() for ln in f if ln.strip()]
    h3 = load_sequences("/root/reference/data/h3n2sample.rda", "sequence")
    assert len(h3) == 8103 and len(h3[0]) == 566


def test_quantile_type7_from_histogram_matches_r_definition():
    from oracle.quantile_r import quantile_type7
    rng = np.random.default_rng(9)
    for n_hash in (1, 7, 50, 500):
        for _ in range(20):
            counts = rng.integers(0, n_hash + 1, size=int(rng.integers(1, 400)))
            hist = np.bincount(counts, minlength=n_hash + 1).astype(np.uint64)
            for p in (0.0, 0.1, 0.5, 0.8, 0.999, 1.0, float(rng.random())):
                thr, mc = da.quantile_type7_counts(hist, n_hash, p)
                assert thr == quantile_type7(counts / n_hash, p)
                kept = counts / n_hash >= thr
                assert ((counts >= mc) == kept).all()
    with pytest.raises(da.DynaAlignError):
        da.quantile_type7_counts(np.zeros(5, np.uint64), 4, 0.5)


def test_vocab_argument_errors_before_device():
    import ctypes as C2
    from dynaalign_b200._lib import flatten, lib, ptr
    res, off = flatten(["ACDE", "AC", "ACDEF"])
    V = C2.c_int64(0)
    rc = lib().dyna_minhash_vocab_ranks(ptr(res, C2.c_uint8), ptr(off, C2.c_int64), 3, 3, None, 0, C2.byref(V), None, None)
    assert rc == _lib.ERR_INVALID and _lib.last_error() == "'k' must be a positive integer between 1 and 2"
    rc = lib().dyna_minhash_vocab_ranks(ptr(res, C2.c_uint8), ptr(off, C2.c_int64), 3, 0, None, 0, C2.byref(V), None, None)
    assert rc == _lib.ERR_INVALID and _lib.last_error() == "'k' must be a positive integer between 1 and 4"


def test_fasta_reader(tmp_path):
    import gzip

    from dynaalign_b200.fasta import read_fasta
    text = ">sp|P1 first protein\nARND\ncqeg\n\n>P2\nHILK MFPS*\n;comment\n>empty\n>P4\nTWYV\n"
    p = tmp_path / "a.fasta"
    p.write_text(text)
    names, seqs = read_fasta(p)
    assert names == ["sp|P1", "P2", "empty", "P4"]
    assert seqs == ["ARNDCQEG", "HILKMFPS", "", "TWYV"]
    assert read_fasta(p, upper=False, strip_terminator=False)[1][:2] == ["ARNDcqeg", "HILKMFPS*"]
    gz = tmp_path / "a.fa.gz"
    with gzip.open(gz, "wt") as f:
        f.write(text)
    assert read_fasta(gz) == (names, seqs)
    bad = tmp_path / "bad.fa"
    bad.write_text("ARND\n>x\nAA\n")
    import pytest
    with pytest.raises(ValueError):
        read_fasta(bad)


def test_offsets_must_be_non_decreasing():
    # foreign callers hand in offsets[]: a decreasing or negative entry is an argument error, raised before any device work
    L = _lib.lib()
    res = np.frombuffer(b"ARNDARND", dtype=np.uint8).copy()
    out = np.zeros(9, dtype=np.float64)
    for bad in ([0, 5, 3, 8], [-1, 2, 4, 8], [0, 9, 8, 8]):
        off = np.array(bad, dtype=np.int64)
        rc = L.dyna_similarityNW(_lib.ptr(res, C.c_uint8), _lib.ptr(off, C.c_int64), 3, b"BLOSUM62", 10, 4, _lib.ptr(out, C.c_double), 1)
        assert rc == _lib.ERR_INVALID and "offsets" in _lib.last_error()
        seeds = np.zeros(4, dtype=np.uint32)
        rc = L.dyna_similarityMH(_lib.ptr(res, C.c_uint8), _lib.ptr(off, C.c_int64), 3, 2, 4, _lib.ptr(seeds, C.c_uint32),
                                 _lib.ptr(out, C.c_double), 1)
        assert rc == _lib.ERR_INVALID and "offsets" in _lib.last_error()
        assert not L.dyna_nw_plan_create(_lib.ptr(res, C.c_uint8), _lib.ptr(off, C.c_int64), 3, b"BLOSUM62", 10, 4, 0, 3, 0)
        assert "offsets" in _lib.last_error()


def test_checksum_restatement_is_position_weighted():
    import dynaalign_b200 as da
    v = np.array([3, 1, 4, 1, 5, 9, 2, 6], dtype=np.uint32)
    whole = da.checksum(v)
    assert whole == (da.checksum(v[:3]) + da.checksum(v[3:], 3)) % (1 << 64)   # additive over slabs at their global offsets
    assert whole != (da.checksum(v[:3]) + da.checksum(v[3:], 4)) % (1 << 64)   # ... and only there
    assert da.checksum(v) != da.checksum(v[::-1].copy())                        # order matters
    w = da.checksum_weights(0, 2)
    assert int(w[0]) == (0x9E3779B97F4A7C15 ^ (0x9E3779B97F4A7C15 >> 31))
