"""Generate the committed golden fixtures from the reference itself.

Run in the build container (where /root/reference exists and oracle/_ref has been built):

    python tests/golden/make_golden.py            # everything except the long C2 run
    python tests/golden/make_golden.py --c2       # also h3n2sample[1:1000] NW stats, every distinct pair checked
                                                  # against the compiled reference (minutes, all cores)

Inputs come from the reference's data/*.rda (read with dynaalign_b200.rda); outputs come from
oracle/_ref/libdynaref.so (the reference's src/*.cpp compiled unmodified).  The fixtures are small
test vectors; no reference source code is copied.
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dynaalign_b200.rda import load_rda  # noqa: E402
from oracle import port, ref  # noqa: E402
from oracle._util import fnv1a64  # noqa: E402

REF = "/root/reference"


def fp(m):
    m = np.asfortranarray(m, dtype=np.float64)
    return {"n": int(m.shape[0]), "fnv1a64": "%016x" % fnv1a64(m.tobytes(order="F")),
            "sum": float(np.sum(m.ravel(order="F"))) if not np.isnan(m).any() else None,
            "x12": float(m[0, 1]), "x1n": float(m[0, -1]), "x23": float(m[1, 2])}


def main():
    evp = load_rda(f"{REF}/data/evp_peparray.rda")["evp_peparray"]["PROBE_SEQUENCE"]
    h3 = load_rda(f"{REF}/data/h3n2sample.rda")["h3n2sample"]["sequence"][:1000]
    uniq = sorted(set(h3))
    idx = [uniq.index(s) for s in h3]
    with open(os.path.join(HERE, "evp_probe_sequences.txt"), "w") as f:
        f.write("\n".join(evp) + "\n")
    with gzip.open(os.path.join(HERE, "h3n2sample_first1000.json.gz"), "wt", compresslevel=9) as f:
        json.dump({"unique": uniq, "index": idx}, f, separators=(",", ":"))

    g = {}
    # --- known answers straight from the reference
    pep4 = ["RRAVELQTVAFP", "PPPSYETVMAAA", "TPPPSYETVMAA", "TPPASYHTVMAA"]
    g["nw_pep4"] = {"sequences": pep4, "matrix": ref.similarityNW(pep4).tolist()}
    a, b = "DCHFSPIG", "PNIWFPHLAWNAKFIPN"
    g["nw_order"] = {"a": a, "b": b, "ab": ref.calculate_similarity(a, b), "ba": ref.calculate_similarity(b, a)}
    g["murmur3"] = [{"key": k, "seed": s, "hash": ref.murmur3_32(k.encode(), s)} for k, s in
                    [("", 0), ("", 1), ("abc", 0), ("Hello, world!", 1234), ("aaaa", 0x9747b28c), ("ABCD", 0), ("AB", 42),
                     ("A", 7), ("ABC", 99), ("ABCDE", 5), ("ABCDEFGH", 6), ("ACDEFGHIKLMNP", 77)]]
    g["hashfamily_12345"] = {"kmer": "ABCD", "n_hash": 3, "hashes": ref.hashfamily_hash(12345, 3, b"ABCD").tolist()}
    g["seeds_42_first8"] = port.hashfamily_seeds(42, 8).tolist()  # validated against ref below
    assert all(ref.hashfamily_hash(42, 8, b"WXYZ")[i] == port.murmur3_32(b"WXYZ", int(s))
               for i, s in enumerate(g["seeds_42_first8"]))
    g["errors"] = {}
    for name, fn in [("mh_empty", lambda: ref.similarityMH([], 4, 50)), ("mh_k0", lambda: ref.similarityMH(["AAAA"], 0, 50)),
                     ("mh_nhash0", lambda: ref.similarityMH(["AAAA"], 4, 0)),
                     ("nw_badname", lambda: ref.similarityNW(["AA"], "BLOSUM63")),
                     ("nw_bad_seq1", lambda: ref.similarityNW(["JA", "AA"])),
                     ("nw_bad_seq2_self", lambda: ref.similarityNW(["AJ", "AA"])),
                     ("nw_bad_seq2_other", lambda: ref.similarityNW(["AA", "AAb"])),
                     ("nw_empty_first_skips", lambda: ref.similarityNW(["", "AA", "Ao"]))]:
        try:
            fn()
            g["errors"][name] = None
        except ref.RefError as e:
            g["errors"][name] = str(e)
    g["nw_edge"] = {"empty_empty_isnan": bool(np.isnan(ref.calculate_similarity("", ""))),
                    "empty_AA": ref.calculate_similarity("", "AA"), "AA_empty": ref.calculate_similarity("AA", ""),
                    "n0_shape": list(ref.similarityNW([]).shape)}
    # --- dataset fingerprints
    g["mh_evp_k2_h50_seed42"] = fp(ref.similarityMH(evp, 2, 50, 42))
    g["nw_evp_blosum62_10_4"] = fp(ref.similarityNW(evp))
    g["nw_h3n2_24"] = fp(ref.similarityNW(h3[:24]))
    g["nw_h3n2_24"]["matrix"] = ref.similarityNW(h3[:24]).tolist()
    g["mh_h3n2_1000_k4_h500_seed42"] = fp(ref.similarityMH(h3, 4, 500, 42))
    # --- other tables / parameters on a small real subset
    sub = evp[:40]
    for nm, go, ge in [("BLOSUM45", 12, 3), ("BLOSUM50", 0, 0), ("BLOSUM80", 7, 1), ("BLOSUM90", 3, 5), ("BLOSUM100", 11, 2)]:
        g[f"nw_evp40_{nm}_{go}_{ge}"] = fp(ref.similarityNW(sub, nm, go, ge))
    # --- a sample of C2 pairs through the reference's calculate_similarity (pins the port at protein length)
    rng = np.random.default_rng(7)
    samp = []
    for _ in range(40):
        i, j = sorted(int(x) for x in rng.integers(0, 1000, 2))
        samp.append({"i": i, "j": j, "sim": ref.calculate_similarity(h3[i], h3[j])})
    g["nw_h3n2_1000_sample"] = samp
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    # signatures for evp (uint32) as a compact array
    np.savez_compressed(os.path.join(HERE, "mh_evp_signatures_k2_h50_seed42.npz"), sig=ref.mh_signatures(evp, 2, 50, 42))
    print("golden.json written")

    if "--c2" in sys.argv:
        # Full config-2 NW (matches, length).  The integers come from the C port (the reference only returns the
        # double matches/length); BEFORE they are stored, every distinct ordered (row sequence, column sequence)
        # combination that occurs among the 500,500 pairs is run through the compiled reference's own
        # calculate_similarity (src/pairwiseSeqAlign.cpp:209-313) and the port's matches/length must reproduce that
        # double exactly.  h3n2sample[1:1000] holds 449 distinct sequences, so this is ~1.0e5 reference alignments
        # (a few CPU-minutes on all cores) instead of 5.0e5.
        import multiprocessing as mp
        mt, ln = port.nw_pair_stats(h3)
        n = len(h3)
        slot = {}
        for i in range(n):
            base = i * n - i * (i - 1) // 2 - i
            ui = idx[i]
            for j in range(i, n):
                slot.setdefault((ui, idx[j]), base + j)
        combos = sorted(slot)
        with mp.Pool(os.cpu_count()) as pool:
            sims = pool.starmap(_ref_sim, [(uniq[u], uniq[v]) for u, v in combos], chunksize=64)
        bad = 0
        for (u, v), sim in zip(combos, sims):
            k = slot[(u, v)]
            got = float(mt[k]) / float(ln[k]) if ln[k] else float("nan")
            if not (got == sim or (np.isnan(got) and np.isnan(sim))):
                bad += 1
        assert bad == 0, "%d of %d distinct pairs differ from the compiled reference" % (bad, len(combos))
        # every occurrence of a combination must carry the same integers (the port is deterministic per pair)
        for i in range(0, n, 37):
            base = i * n - i * (i - 1) // 2 - i
            for j in range(i, n, 11):
                k = slot[(idx[i], idx[j])]
                assert mt[base + j] == mt[k] and ln[base + j] == ln[k]
        np.savez_compressed(os.path.join(HERE, "nw_h3n2_1000_stats.npz"), matches=mt.astype(np.uint16), length=ln.astype(np.uint16))
        with open(os.path.join(HERE, "nw_h3n2_1000_stats.provenance.json"), "w") as f:
            json.dump({"distinct_sequences": len(uniq), "distinct_ordered_pairs_checked_against_compiled_reference": len(combos),
                       "mismatches": bad, "pairs_total": int(mt.size),
                       "check": "float(matches)/float(length) == ref.calculate_similarity(row, column), bit-exact doubles"}, f, indent=1)
        print("C2 stats written; all %d distinct ordered pairs checked against the compiled reference" % len(combos))


def _ref_sim(a, b):
    return ref.calculate_similarity(a, b)


if __name__ == "__main__":
    main()
