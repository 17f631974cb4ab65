import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        from dynaalign_b200 import _lib
        return _lib.lib().dyna_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are only meaningful where a device exists; never silently pass them on a CPU box
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def evp():
    with open(os.path.join(GOLDEN, "evp_probe_sequences.txt")) as f:
        return [ln.strip() for ln in f if ln.strip()]


@pytest.fixture(scope="session")
def h3n2():
    with gzip.open(os.path.join(GOLDEN, "h3n2sample_first1000.json.gz"), "rt") as f:
        d = json.load(f)
    return [d["unique"][i] for i in d["index"]]


def fingerprint(m):
    from oracle._util import fnv1a64
    m = np.asfortranarray(m, dtype=np.float64)
    return "%016x" % fnv1a64(m.tobytes(order="F"))


def same_matrix(a, b):
    """bit-identical doubles (NaN == NaN)"""
    a = np.asfortranarray(a, dtype=np.float64)
    b = np.asfortranarray(b, dtype=np.float64)
    return a.shape == b.shape and a.tobytes(order="F") == b.tobytes(order="F")


ALPHABET24 = "ARNDCQEGHILKMFPSTWYVBZX*"
TABLES = ["BLOSUM45", "BLOSUM50", "BLOSUM62", "BLOSUM80", "BLOSUM90", "BLOSUM100"]


def random_seqs(rng, n, lo, hi, alphabet=ALPHABET24):
    al = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return [al[rng.integers(0, len(al), size=int(rng.integers(lo, hi + 1)))].tobytes().decode() for _ in range(n)]
