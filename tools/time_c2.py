"""Development helper (GPU box): DYNA_TIMING phases of the R-facing config-2 call."""
import gzip, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynaalign_b200 as da
with gzip.open(os.path.join(ROOT, "tests/golden/h3n2sample_first1000.json.gz"), "rt") as f:
    d = json.load(f)
h3 = [d["unique"][i] for i in d["index"]]
da.similarityNW(h3)
os.environ["DYNA_TIMING"] = "1"
for _ in range(2):
    t0 = time.perf_counter(); da.similarityNW(h3); print("wall %.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
