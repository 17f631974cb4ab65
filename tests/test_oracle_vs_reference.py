"""CPU: pins the oracle (oracle/dyna_oracle.c) against the reference.

Two anchors: (1) the committed fixtures in tests/golden/ that were generated from the reference's own C++
(tests/golden/make_golden.py), always checked; (2) the compiled reference itself (oracle/_ref), checked whenever
it is present (build container, and on the GPU box because the prebuilt .so travels)."""
import json
import os

import numpy as np
import pytest

from conftest import ALPHABET24, GOLDEN, TABLES, fingerprint, random_seqs, same_matrix
from oracle import port, ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def test_murmur3_known_answers(golden):
    for v in golden["murmur3"]:
        assert port.murmur3_32(v["key"].encode(), v["seed"]) == v["hash"]
    # canonical MurmurHash3_x86_32 vectors
    assert port.murmur3_32(b"", 0) == 0
    assert port.murmur3_32(b"", 1) == 0x514E28B7
    assert port.murmur3_32(b"abc", 0) == 0xB3DD93FA
    assert port.murmur3_32(b"Hello, world!", 1234) == 0xFAF6CDB3
    assert port.murmur3_32(b"aaaa", 0x9747B28C) == 0x5A97808A


def test_hashfamily_seed_stream(golden):
    assert port.hashfamily_seeds(42, 8).tolist() == golden["seeds_42_first8"]
    s = port.hashfamily_seeds(12345, 3)
    got = [port.murmur3_32(b"ABCD", int(x)) for x in s]
    assert got == golden["hashfamily_12345"]["hashes"]


def test_tables_match_golden():
    with open(os.path.join(GOLDEN, "blosum_tables.json")) as f:
        g = json.load(f)
    assert g["alphabet"] == ALPHABET24
    for nm in TABLES:
        assert port.substitution_matrix(nm).tolist() == g["tables"][nm]


def test_nw_known_answers(golden):
    g = golden["nw_pep4"]
    assert same_matrix(port.similarityNW(g["sequences"]), np.array(g["matrix"]))
    o = golden["nw_order"]
    m, l = port.nw_pair(o["a"], o["b"])
    assert m / l == o["ab"]
    m, l = port.nw_pair(o["b"], o["a"])
    assert m / l == o["ba"]
    assert o["ab"] != o["ba"]  # order sensitivity is a property of the reference
    assert np.isnan(port.similarityNW([""])[0, 0])
    assert port.nw_pair("", "AA") == (0, 2) and port.nw_pair("AA", "") == (0, 2)


def test_nw_errors(golden):
    e = golden["errors"]
    for seqs, name, key in [(["AA"], "BLOSUM63", "nw_badname"), (["JA", "AA"], "BLOSUM62", "nw_bad_seq1"),
                            (["AJ", "AA"], "BLOSUM62", "nw_bad_seq2_self"), (["AA", "AAb"], "BLOSUM62", "nw_bad_seq2_other"),
                            (["", "AA", "Ao"], "BLOSUM62", "nw_empty_first_skips")]:
        with pytest.raises(port.OracleError) as ei:
            port.similarityNW(seqs, name)
        assert str(ei.value) == e[key]


def test_mh_errors(golden):
    e = golden["errors"]
    for args, key in [(([], 4, 50), "mh_empty"), ((["AAAA"], 0, 50), "mh_k0"), ((["AAAA"], 4, 0), "mh_nhash0")]:
        with pytest.raises(port.OracleError) as ei:
            port.similarityMH(*args)
        assert str(ei.value) == e[key]


def test_dataset_fingerprints(golden, evp, h3n2):
    assert fingerprint(port.similarityMH(evp, 2, 50, 42)) == golden["mh_evp_k2_h50_seed42"]["fnv1a64"]
    assert fingerprint(port.similarityNW(evp)) == golden["nw_evp_blosum62_10_4"]["fnv1a64"]
    assert fingerprint(port.similarityNW(h3n2[:24])) == golden["nw_h3n2_24"]["fnv1a64"]
    assert same_matrix(port.similarityNW(h3n2[:24]), np.array(golden["nw_h3n2_24"]["matrix"]))
    assert fingerprint(port.similarityMH(h3n2, 4, 500, 42)) == golden["mh_h3n2_1000_k4_h500_seed42"]["fnv1a64"]
    for key, fp in golden.items():
        if key.startswith("nw_evp40_"):
            _, _, nm, go, ge = key.split("_")
            assert fingerprint(port.similarityNW(evp[:40], nm, int(go), int(ge))) == fp["fnv1a64"], key


def test_evp_signatures_golden(evp):
    sig = np.load(os.path.join(GOLDEN, "mh_evp_signatures_k2_h50_seed42.npz"))["sig"]
    assert (port.mh_signatures(evp, 2, port.hashfamily_seeds(42, 50)) == sig).all()


def test_protein_pairs_sample(golden, h3n2):
    for s in golden["nw_h3n2_1000_sample"][:12]:
        m, l = port.nw_pair(h3n2[s["i"]], h3n2[s["j"]])
        assert m / l == s["sim"]
        assert port.nw_pair(h3n2[s["i"]], h3n2[s["j"]], forward=True) == (m, l)


def test_forward_formulation_equals_traceback():
    rng = np.random.default_rng(11)
    for it in range(1500):
        alpha = ALPHABET24 if it % 2 else "ACDE"
        a, b = random_seqs(rng, 2, 0, 40, alpha)
        nm = TABLES[it % 6]
        go, ge = int(rng.integers(0, 13)), int(rng.integers(0, 6))
        assert port.nw_pair(a, b, matrixName=nm, gapOpen=go, gapExt=ge) == \
            port.nw_pair(a, b, matrixName=nm, gapOpen=go, gapExt=ge, forward=True)


def test_match_counts_slabs():
    rng = np.random.default_rng(5)
    sig = rng.integers(0, 4, size=(37, 9), dtype=np.uint32)
    full = port.mh_match_counts(sig)
    parts = [port.mh_match_counts(sig, a, b) for a, b in [(0, 5), (5, 20), (20, 37)]]
    assert (np.concatenate(parts) == full).all()
    brute = [int((sig[i] == sig[j]).sum()) for i in range(37) for j in range(i + 1, 37)]
    assert full.tolist() == brute


# ------------------------------------------------------------------ against the compiled reference itself
@needs_ref
def test_ref_murmur_and_seeds():
    rng = np.random.default_rng(1)
    for _ in range(500):
        key = bytes(rng.integers(0, 256, int(rng.integers(0, 24)), dtype=np.uint8))
        seed = int(rng.integers(0, 2 ** 32))
        assert ref.murmur3_32(key, seed) == port.murmur3_32(key, seed)
    for seed in [0, 1, 42, 2 ** 32 - 1]:
        s = port.hashfamily_seeds(seed, 700)  # crosses the 624-word MT refill
        r = ref.hashfamily_hash(seed, 700, b"WXYZ")
        assert (r == np.array([port.murmur3_32(b"WXYZ", int(x)) for x in s], dtype=np.uint32)).all()


@needs_ref
def test_ref_tables_and_alphabet():
    for nm in TABLES:
        assert (ref.substitution_matrix(nm) == port.substitution_matrix(nm)).all()
    t = ref.aa_index_table()
    assert "".join(chr(c) for c in np.argsort(np.where(t >= 0, t, 999))[:24]) == ALPHABET24


@needs_ref
def test_ref_nw_random_pairs():
    rng = np.random.default_rng(2)
    for it in range(3000):
        alpha = ALPHABET24 if it % 2 else "ACDE"
        a, b = random_seqs(rng, 2, 0, 40, alpha)
        nm = TABLES[it % 6]
        go, ge = int(rng.integers(0, 13)), int(rng.integers(0, 6))
        r = ref.calculate_similarity(a, b, nm, go, ge)
        m, l = port.nw_pair(a, b, matrixName=nm, gapOpen=go, gapExt=ge)
        with np.errstate(all="ignore"):
            p = np.float64(m) / np.float64(l)
        assert (np.isnan(r) and np.isnan(p)) or r == p


@needs_ref
def test_ref_similarity_matrices():
    rng = np.random.default_rng(3)
    seqs = random_seqs(rng, 30, 0, 30, "ACDEFGHIKLMNPQRSTVWY") + ["", "A"]
    assert same_matrix(ref.similarityNW(seqs, "BLOSUM80", 7, 2), port.similarityNW(seqs, "BLOSUM80", 7, 2))
    for k, nh in [(1, 7), (2, 50), (3, 33), (4, 500), (5, 64), (9, 20)]:
        assert same_matrix(ref.similarityMH(seqs, k, nh, 7), port.similarityMH(seqs, k, nh, 7))
        assert (ref.mh_signatures(seqs, k, nh, 7) == port.mh_signatures(seqs, k, port.hashfamily_seeds(7, nh))).all()
