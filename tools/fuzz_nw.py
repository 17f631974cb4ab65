"""Development helper (GPU box): randomized differential test of the NW kernels against the oracle port, aimed at the
shape boundaries of the packed kernels (strip heights, lane-31 rotation on/off, unit sizes of 64 columns, sorted
pairing with ties, odd unit tails, empty and 1-residue sequences).  python tools/fuzz_nw.py [rounds] [seed]"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynaalign_b200 as da  # noqa: E402
from oracle import port  # noqa: E402  (checker only)

AL = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVBZX*", dtype=np.uint8)
TABLES = ["BLOSUM45", "BLOSUM50", "BLOSUM62", "BLOSUM80", "BLOSUM90", "BLOSUM100"]


def dataset(rng):
    kind = rng.integers(0, 7)
    if kind == 0:  # around one strip height R: lengths 31R-3 .. 32R+3 (rotation boundary and next height)
        R = int(rng.integers(2, 21))
        lens = rng.integers(max(1, 31 * R - 3), 32 * R + 4, size=int(rng.integers(60, 140)))
    elif kind == 1:  # wide mix incl. empties and tiny ones
        lens = rng.integers(0, 700, size=int(rng.integers(40, 90)))
        lens[rng.integers(0, len(lens), 3)] = [0, 1, 2]
    elif kind == 2:  # many equal lengths (ties in the sorted pairing), unit tails of 63/64/65/66 columns
        base = int(rng.integers(33, 400))
        lens = np.full(int(rng.choice([63, 64, 65, 66, 127, 129])), base)
        lens[rng.integers(0, len(lens), 5)] += rng.integers(-3, 4, 5)
    elif kind == 3:  # short probes with a few long columns
        lens = rng.integers(1, 33, size=int(rng.integers(80, 200)))
        lens[rng.integers(0, len(lens), 4)] = rng.integers(300, 1500, 4)
    elif kind == 5:  # cooperative two-warp kernel: rows 385..768 around one strip height (64R-3 .. 64R+3), units of 128 columns
        R = int(rng.integers(7, 13))
        n = int(rng.choice([40, 127, 128, 129, 131]))
        lens = rng.integers(max(385, 64 * (R - 1) - 3), min(768, 64 * R + 3) + 1, size=n)
        lens[rng.integers(0, n, 6)] = [0, 1, 385, 768, 1024, 1030][: 6]
    elif kind == 6:  # cooperative kernel: HA-like family (equal lengths, ties in the pairing) with a few outliers
        base = int(rng.integers(500, 640))
        lens = np.full(int(rng.choice([64, 130, 200])), base)
        lens[rng.integers(0, len(lens), 8)] += rng.integers(-40, 40, 8)
    else:  # long rows (multi-pass) with mixed columns
        lens = np.concatenate([rng.integers(650, 1300, size=12), rng.integers(1, 600, size=20)])
    lens = np.maximum(lens, 0)
    fam = AL[rng.integers(0, 20, size=int(lens.max()) + 8)]
    seqs = []
    for L in lens.tolist():
        if rng.random() < 0.5:  # related sequences exercise gaps and ties
            s = fam[:L].copy()
            mut = rng.random(L) < 0.08
            s[mut] = AL[rng.integers(0, 24, size=int(mut.sum()))]
            if L > 4 and rng.random() < 0.5:
                cut = int(rng.integers(1, L - 1))
                s = np.concatenate([s[:cut], s[cut + 1:], AL[rng.integers(0, 20, size=1)]])
        else:
            s = AL[rng.integers(0, 24, size=L)]
        seqs.append(s.tobytes().decode())
    rng.shuffle(seqs)
    table = TABLES[int(rng.integers(0, 6))]
    go, ge = [(10, 4), (10, 4), (3, 1), (0, 0), (12, 2), (1, 7), (25, 3)][int(rng.integers(0, 7))]
    return seqs, table, go, ge


def oracle_stats(args):
    seqs, table, go, ge = args
    return port.nw_pair_stats(seqs, table, go, ge)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    sets = [dataset(rng) for _ in range(rounds)]
    with ProcessPoolExecutor(max_workers=min(rounds, os.cpu_count() or 1)) as ex:
        wants = list(ex.map(oracle_stats, sets))
    bad = 0
    for i, ((seqs, table, go, ge), (wm, wl)) in enumerate(zip(sets, wants)):
        gm, gl = da.nw_pair_stats(seqs, table, go, ge)
        ok = (gm == wm).all() and (gl == wl).all()
        os.environ["DYNA_NW_ROWS2"] = "1"  # the two-rows kernel is size-gated by default: force it as well
        gm, gl = da.nw_pair_stats(seqs, table, go, ge)
        del os.environ["DYNA_NW_ROWS2"]
        ok = ok and (gm == wm).all() and (gl == wl).all()
        bad += 0 if ok else 1
        lens = [len(s) for s in seqs]
        print("set %2d: n=%3d len %d..%d %s go=%d ge=%d pairs=%d %s" % (i, len(seqs), min(lens), max(lens), table, go, ge, len(wm),
                                                                      "OK" if ok else "MISMATCH at %s" % np.nonzero((gm != wm) | (gl != wl))[0][:5]))
    print("fuzz:", "all OK" if bad == 0 else "%d sets FAILED" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
