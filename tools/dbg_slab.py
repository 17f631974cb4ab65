import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    import dynaalign_b200 as da
    from oracle import port
    a, b, n, nh = map(int, sys.argv[1:5])
    rng = np.random.default_rng(8)
    sig = rng.integers(0, 5, size=(n, nh), dtype=np.uint32)
    try:
        got = da.mh_match_counts(sig, a, b)
        print(a, b, n, nh, os.environ.get("DYNA_MH_MATCH"), "OK" if (got == port.mh_match_counts(sig, a, b)).all() else "MISMATCH")
    except Exception as e:
        print(a, b, n, nh, os.environ.get("DYNA_MH_MATCH"), "ERR", str(e)[-80:])
else:
    for mode in ("tma", "ldg"):
        for (a, b, n, nh) in [(0, 95, 900, 100), (95, 203, 900, 100), (128, 256, 900, 100), (4, 100, 900, 100), (1, 2, 900, 100),
                              (2, 100, 900, 100), (64, 100, 900, 100), (0, 900, 900, 100), (0, 95, 900, 500), (96, 200, 900, 16)]:
            env = dict(os.environ, DYNA_MH_MATCH=mode)
            subprocess.run([sys.executable, __file__, str(a), str(b), str(n), str(nh)], env=env)
