"""Development helper (GPU box): randomized differential test of the MinHash match kernels (u32 and relabelled 16-bit
paths, row slabs, hash counts around the 16-row stage boundaries) against the oracle port.
python tools/fuzz_mh.py [rounds] [seed]"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynaalign_b200 as da  # noqa: E402
from oracle import port  # noqa: E402  (checker only)


def dataset(rng):
    n = int(rng.choice([2, 3, 127, 128, 129, 255, 257, 1000, int(rng.integers(2, 9000))]))
    n_hash = int(rng.choice([1, 2, 15, 16, 17, 31, 32, 33, 50, 63, 65, 100, 500, int(rng.integers(1, 600))]))
    n_hash = min(n_hash, max(1, int(4e9 // (n * n))))  # keep the oracle quick
    alphabet = int(rng.choice([2, 3, 50, 70000, 2 ** 32 - 1]))
    sig = rng.integers(0, alphabet, size=(n, n_hash), dtype=np.uint64).astype(np.uint32)
    if rng.random() < 0.3:
        sig[rng.integers(0, n, n // 3)] = sig[0]  # duplicates: counts up to n_hash
    a = int(rng.integers(0, n))
    b = int(rng.integers(a, n + 1))
    if rng.random() < 0.5:
        a, b = 0, n
    return sig, a, b, int(rng.integers(0, 2))


def oracle_counts(args):
    sig, a, b, _ = args
    return port.mh_match_counts(sig, a, b)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    sets = [dataset(rng) for _ in range(rounds)]
    with ProcessPoolExecutor(max_workers=min(rounds, os.cpu_count() or 1)) as ex:
        wants = list(ex.map(oracle_counts, sets))
    bad = 0
    for i, ((sig, a, b, pack), want) in enumerate(zip(sets, wants)):
        os.environ["DYNA_MH_PACK16"] = str(pack)
        got = da.mh_match_counts(sig, a, b)
        ok = len(got) == len(want) and (got == want).all()
        bad += 0 if ok else 1
        print("set %2d: n=%5d n_hash=%3d rows [%d,%d) pack16=%d pairs=%d max=%d %s" % (i, sig.shape[0], sig.shape[1], a, b, pack, len(want),
                                                                                      int(want.max()) if len(want) else 0, "OK" if ok else "MISMATCH"))
    print("fuzz:", "all OK" if bad == 0 else "%d sets FAILED" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
