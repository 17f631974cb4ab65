"""The R package's Rcpp shims (rpkg/src/dyna_shims.cpp), compiled against the stub Rcpp.h (R is not installed in
this image) and driven through tests/shim/libshimharness.so: the marshalling a maintainer would ship -- string
flattening, error propagation through Rcpp::stop, dimnames, column-major layout -- checked end to end."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, fingerprint, same_matrix
from oracle import minhash_r as R
from oracle import port
from oracle._util import flatten

_H = None


def harness():
    global _H
    if _H is None:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "shim")], stdout=subprocess.DEVNULL)
        _H = C.CDLL(os.path.join(ROOT, "tests", "shim", "libshimharness.so"))
        _H.shim_last_error.restype = C.c_char_p
    return _H


def call_mh(seqs, k, n_hash):
    res, off = flatten(seqs)
    n = len(seqs)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    ok = C.c_int(0)
    rc = harness().shim_similarityMH(res.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(n),
                                     k, n_hash, out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ok))
    if rc:
        raise RuntimeError(harness().shim_last_error().decode())
    return out, ok.value


def call_nw(seqs, name="BLOSUM62", go=10, ge=4):
    res, off = flatten(seqs)
    n = len(seqs)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    ok = C.c_int(0)
    rc = harness().shim_similarityNW(res.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(n),
                                     name.encode(), go, ge, out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ok))
    if rc:
        raise RuntimeError(harness().shim_last_error().decode())
    return out, ok.value


def test_shims_raise_reference_errors(golden):
    e = golden["errors"]
    for fn, key in [(lambda: call_mh([], 4, 50), "mh_empty"), (lambda: call_mh(["AAAA"], 0, 50), "mh_k0"),
                    (lambda: call_mh(["AAAA"], 4, 0), "mh_nhash0"), (lambda: call_nw(["AA"], "BLOSUM63"), "nw_badname"),
                    (lambda: call_nw(["JA", "AA"]), "nw_bad_seq1"), (lambda: call_nw(["AJ", "AA"]), "nw_bad_seq2_self"),
                    (lambda: call_nw(["", "AA", "Ao"]), "nw_empty_first_skips")]:
        with pytest.raises(RuntimeError) as ei:
            fn()
        assert str(ei.value) == e[key]
    out, ok = call_nw([])  # 0 x 0 matrix, no error, like the reference
    assert out.shape == (0, 0)


@pytest.mark.gpu
def test_shim_similarityMH_matches_reference(monkeypatch, golden, evp):
    monkeypatch.setenv("DYNAALIGN_SEED", "42")
    m, ok = call_mh(evp, 2, 50)
    assert ok == 1 and fingerprint(m) == golden["mh_evp_k2_h50_seed42"]["fnv1a64"]
    monkeypatch.delenv("DYNAALIGN_SEED")
    m, ok = call_mh(evp[:40], 2, 50)  # random_device seed: valid but not reproducible, as in the reference
    assert ok == 1 and (np.diag(m) == 1).all() and (m == m.T).all()


@pytest.mark.gpu
def test_shim_similarityNW_matches_reference(golden, h3n2):
    m, ok = call_nw(h3n2[:24])
    assert ok == 1 and fingerprint(m) == golden["nw_h3n2_24"]["fnv1a64"]
    seqs = ["", "ARND", "", "AR"]
    m, ok = call_nw(seqs, "BLOSUM80", 3, 1)
    assert ok == 1 and same_matrix(m, port.similarityNW(seqs, "BLOSUM80", 3, 1))


@pytest.mark.gpu
def test_shim_r_pipeline_halves():
    seqs = ["ACDEGHHIKLLL", "ACDEGHHIKLMN", "XXXXXYYYYYYZZ", "ACDEGHHIKLLL"]
    vocab = R.create_vocab(seqs, 3)
    hp = R.create_hash_parameters(100, len(vocab), np.random.default_rng(3))
    cm = R.create_char_matrix(seqs, vocab, 3)
    want_sig = R.compute_signature_matrix(cm, hp, len(vocab))
    ranks, off = R.shingle_ranks(seqs, vocab, 3)
    a, b = hp["a"].astype(np.float64), hp["b"].astype(np.float64)
    out = np.zeros((100, 4), dtype=np.float64, order="F")
    ranks32, offd = ranks.astype(np.int32), off.astype(np.float64)
    rc = harness().shim_mh_signatures_linear(ranks32.ctypes.data_as(C.POINTER(C.c_int)), C.c_int64(len(ranks32)),
                                             offd.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(offd)),
                                             a.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)),
                                             C.c_double(len(vocab)), 100, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0 and (out == want_sig).all()
    codes = np.asfortranarray(np.stack([np.unique(want_sig[h], return_inverse=True)[1] for h in range(100)]).astype(np.int32))
    d = np.zeros((4, 4), dtype=np.float64, order="F")
    rc = harness().shim_mh_distance_matrix(codes.ctypes.data_as(C.POINTER(C.c_int)), 100, 4, d.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0 and same_matrix(d, R.compute_distance_matrix(want_sig))


def call_mh_edges(seqs, k, n_hash, p, cap):
    res, off = flatten(seqs)
    edges = np.zeros((max(cap, 1), 3), dtype=np.float64, order="F")
    ne, thr = C.c_int64(0), C.c_double(0)
    rc = harness().shim_similarityMH_edges(res.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                           C.c_int64(len(seqs)), k, n_hash, C.c_double(p), C.c_int64(cap),
                                           edges.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ne), C.byref(thr))
    if rc:
        raise RuntimeError(harness().shim_last_error().decode())
    return thr.value, edges[:ne.value]


def test_shim_edges_raise_reference_errors(golden):
    e = golden["errors"]
    for fn, key in [(lambda: call_mh_edges([], 4, 50, 0.8, 1), "mh_empty"), (lambda: call_mh_edges(["AAAA"], 0, 50, 0.8, 1), "mh_k0"),
                    (lambda: call_mh_edges(["AAAA"], 4, 0, 0.8, 1), "mh_nhash0")]:
        with pytest.raises(RuntimeError) as ei:
            fn()
        assert str(ei.value) == e[key]


@pytest.mark.gpu
def test_shim_similarityMH_edges_is_the_thresholded_reference_matrix(monkeypatch, evp):
    # R/clusterbreak.R:219-221 on the reference's own matrix: threshold <- quantile(upper.tri, p); sim[sim < threshold] <- 0
    from oracle.quantile_r import quantile_type7
    monkeypatch.setenv("DYNAALIGN_SEED", "42")
    full = port.similarityMH(evp, 2, 50, 42)
    n = len(evp)
    for p in (0.8, 0.99):
        want_thr = quantile_type7(full[np.triu_indices(n, 1)], p)
        dense = full.copy()
        dense[dense < want_thr] = 0.0
        wi, wj = np.nonzero(np.triu(dense, 1))
        thr, edges = call_mh_edges(evp, 2, 50, p, n * n // 2)
        assert thr == want_thr
        assert (edges[:, 0] == wi + 1).all() and (edges[:, 1] == wj + 1).all() and (edges[:, 2] == dense[wi, wj]).all()


def call_nw_edges(seqs, p, cap, name="BLOSUM62", go=10, ge=4):
    res, off = flatten(seqs)
    edges = np.zeros((max(cap, 1), 3), dtype=np.float64, order="F")
    self_w = np.zeros(max(len(seqs), 1), dtype=np.float64)
    ne, thr = C.c_int64(0), C.c_double(0)
    rc = harness().shim_similarityNW_edges(res.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                           C.c_int64(len(seqs)), name.encode(), go, ge, C.c_double(p), C.c_int64(cap),
                                           edges.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ne), C.byref(thr),
                                           self_w.ctypes.data_as(C.POINTER(C.c_double)))
    if rc:
        raise RuntimeError(harness().shim_last_error().decode())
    return thr.value, edges[:ne.value], self_w[:len(seqs)]


def test_shim_nw_edges_raise_reference_errors(golden):
    e = golden["errors"]
    for fn, key in [(lambda: call_nw_edges(["AA"], 0.8, 1, "BLOSUM63"), "nw_badname"), (lambda: call_nw_edges(["JA", "AA"], 0.8, 1), "nw_bad_seq1")]:
        with pytest.raises(RuntimeError) as ei:
            fn()
        assert str(ei.value) == e[key]


@pytest.mark.gpu
def test_shim_similarityNW_edges_is_the_thresholded_reference_matrix(h3n2):
    # R/clusterbreak.R:219-221 on the reference's own similarityNW matrix
    from oracle.quantile_r import quantile_type7
    seqs = h3n2[:40] + ["ARNDARND", "ARNDCRND", "WWWW"]
    full = port.similarityNW(seqs)
    n = len(seqs)
    for p in (0.5, 0.93):
        want_thr = quantile_type7(full[np.triu_indices(n, 1)], p)
        dense = full.copy()
        dense[dense < want_thr] = 0.0
        wi, wj = np.nonzero(np.triu(dense, 1))
        thr, edges, self_w = call_nw_edges(seqs, p, n * n // 2)
        assert thr == want_thr
        assert (edges[:, 0] == wi + 1).all() and (edges[:, 1] == wj + 1).all() and (edges[:, 2] == dense[wi, wj]).all()
        assert (self_w == np.diag(full)).all()
    thr, edges, self_w = call_nw_edges(["ARND"], 0.8, 4)  # 1 x 1: no pairs, no threshold, one self-loop
    assert len(edges) == 0 and self_w[0] == 1.0
