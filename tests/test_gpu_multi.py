"""GPU, >= 2 devices: one process driving several GPUs through the host entry points (row blocks, one host thread
per device, host-side scatter) gives bit-identical matrices to the single-GPU run."""
import pytest

import dynaalign_b200 as da
from dynaalign_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _need2():
    if _lib.lib().dyna_device_count() < 2:
        pytest.skip("needs 2 CUDA devices")


def test_similarityNW_two_gpus_identical():
    _need2()
    seqs = [s.decode() for s in synth.proteins_families(260)]
    a = da.similarityNW(seqs, n_gpus=1)
    b = da.similarityNW(seqs, n_gpus=2)
    assert a.tobytes(order="F") == b.tobytes(order="F")


def test_similarityMH_two_gpus_identical():
    _need2()
    peps = [s.decode() for s in synth.peptides_clustered(3000, children=20)]
    a = da.similarityMH(peps, 4, 100, seed=42, n_gpus=1)
    b = da.similarityMH(peps, 4, 100, seed=42, n_gpus=2)
    assert a.tobytes(order="F") == b.tobytes(order="F")
