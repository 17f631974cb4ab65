import gzip, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynaalign_b200 as da
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with gzip.open(os.path.join(ROOT, "tests/golden/h3n2sample_first1000.json.gz"), "rt") as f:
    d = json.load(f)
h3 = [d["unique"][i] for i in d["index"]]
evp = [l.strip() for l in open(os.path.join(ROOT, "tests/golden/evp_probe_sequences.txt")) if l.strip()]
for name, fn in [("MH h3n2 k4 h500", lambda: da.similarityMH(h3, 4, 500, seed=42)), ("MH evp k2 h50", lambda: da.similarityMH(evp, 2, 50, seed=42)),
                 ("NW evp", lambda: da.similarityNW(evp))]:
    fn()
    t0 = time.perf_counter(); fn(); print(name, "%.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
