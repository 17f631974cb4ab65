"""CPU: the reference's own testthat suite (tests/testthat/test-minHash.R), restated against the numpy
restatement of R/minHash.R (oracle/minhash_r.py) and against the product's host-side mirror where no GPU is
needed.  Line numbers refer to tests/testthat/test-minHash.R."""
import numpy as np
import pytest

import dynaalign_b200 as da
from oracle import minhash_r as R
from oracle import port


@pytest.mark.parametrize("impl", [R, da])
def test_shingle(impl):  # :2-14
    err = R.RError if impl is R else da.DynaAlignError
    assert impl.shingle("ABCDEF", 3) == ["ABC", "BCD", "CDE", "DEF"]
    with pytest.raises(err, match="Input 'x' must be a single character string"):
        impl.shingle(123, 3)
    with pytest.raises(err, match="'k' must be a positive integer between 1 and 6"):
        impl.shingle("ABCDEF", 0)
    with pytest.raises(err, match="'k' must be a positive integer between 1 and 6"):
        impl.shingle("ABCDEF", 7)
    assert impl.shingle("AB", 2) == ["AB"]
    assert len(impl.shingle("ABCDEF", 1)) == 6


@pytest.mark.parametrize("impl", [R, da])
def test_create_vocab(impl):  # :17-30
    vocab = impl.create_vocab(["ACDEGHHIKLLL", "ACDEGHHIKLMN"], k=3)
    assert vocab == sorted(vocab)
    assert len(vocab) == len(set(vocab))
    assert all(len(v) == 3 for v in vocab)


@pytest.mark.parametrize("impl", [R, da])
def test_create_char_matrix(impl):  # :33-44
    seqs = ["ACDEGHHIKLLL", "ACDEGHHIKLMN"]
    vocab = impl.create_vocab(seqs, k=3)
    cm = impl.create_char_matrix(seqs, vocab, k=3)
    assert cm.shape == (len(vocab), len(seqs))
    assert set(np.unique(cm)) <= {0, 1}
    assert (np.asarray(R.create_char_matrix(seqs, vocab, 3)) == cm).all()


@pytest.mark.parametrize("impl", [R, da])
def test_create_hash_parameters(impl):  # :47-60
    err = R.RError if impl is R else da.DynaAlignError
    hp = impl.create_hash_parameters(10, 100, np.random.default_rng(0))
    assert len(hp["a"]) == 10 and len(hp["b"]) == 10
    assert ((hp["a"] > 0) & (hp["a"] <= 100)).all()
    assert ((hp["b"] >= 0) & (hp["b"] <= 100)).all()
    with pytest.raises(err, match="Number of hash functions must be positive"):
        impl.create_hash_parameters(0, 100)
    with pytest.raises(err, match="Maximum value must be at least 2"):
        impl.create_hash_parameters(5, 1)


@pytest.mark.parametrize("impl", [R, da])
def test_apply_hash(impl):  # :63-72 (13 is the implied value)
    r = impl.apply_hash(5, 2, 3, 100)
    assert 0 <= r < 100 and r == 13
    assert impl.apply_hash(5, 2, 3, 100) == impl.apply_hash(5, 2, 3, 100)


def test_compute_signature_matrix_oracle():  # :75-89
    seqs = ["ACDEGHHIKLLL", "ACDEGHHIKLMN"]
    vocab = R.create_vocab(seqs, 3)
    cm = R.create_char_matrix(seqs, vocab, 3)
    hp = R.create_hash_parameters(10, len(vocab), np.random.default_rng(1))
    sig = R.compute_signature_matrix(cm, hp, len(vocab))
    assert sig.shape == (10, 2) and sig.dtype == np.float64
    # the C oracle (rank lists instead of the dense matrix) gives the same values
    rk, off = R.shingle_ranks(seqs, vocab, 3)
    assert (port.mh_signatures_linear(rk, off, hp["a"], hp["b"], len(vocab)).T.astype(float) == sig).all()


def test_compute_distance_matrix_oracle():  # :92-106, mock matrix(c(1,2,3, 1,2,4, 2,3,5), 3, 3) is column-major
    sig = np.array([[1, 2, 3], [1, 2, 4], [2, 3, 5]], dtype=float).T
    d = R.compute_distance_matrix(sig)
    assert (d == d.T).all() and (np.diag(d) == 0).all() and ((d >= 0) & (d <= 1)).all()
    assert d[0, 1] == 1 - 2 / 3 and d[0, 2] == 1.0 and d[1, 2] == 1.0
    assert (port.mh_distance_matrix(sig.T.astype(np.uint32)) == d).all()


def test_minhash_end_to_end_oracle():  # :109-122
    seqs = ["ACDEGHHIKLLL", "ACDEGHHIKLMN", "XXXXXYYYYYYZZ"]
    r = R.minhash(seqs, k=3, n_hash=100, rng=np.random.default_rng(2))
    assert {"vocabulary", "char_matrix", "sig_matrix", "dist_matrix"} <= set(r)
    assert r["char_matrix"].shape == (len(r["vocabulary"]), 3)
    assert r["sig_matrix"].shape == (100, 3)
    assert r["dist_matrix"].shape == (3, 3)


def test_long_double_mean_domain():
    # R's mean() divides in long double; identical to double division for n_hash <= 2050
    for n in (50, 500, 2050):
        m = np.arange(0, n + 1)
        assert ((m / n) == (np.longdouble(1) * m / np.longdouble(n)).astype(np.float64)).all()
    assert float(np.longdouble(115) / np.longdouble(2051)) != 115 / 2051
