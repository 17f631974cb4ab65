"""TEST INFRASTRUCTURE ONLY: R's quantile(x, probs, type = 7) (stats::quantile.default) restated for one
probability -- the threshold clusterbreak derives from the similarity matrix (R/clusterbreak.R:219)."""
import math

import numpy as np


def quantile_type7(x, prob):
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = len(x)
    index = 1 + max(n - 1, 0) * prob
    lo, hi = math.floor(index), math.ceil(index)
    qs = x[lo - 1]
    if index > lo and x[hi - 1] != qs:
        h = index - lo
        qs = (1 - h) * qs + h * x[hi - 1]
    return float(qs)
