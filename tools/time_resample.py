"""Device time of k_weights (exact left folds) and k_resample_indices by population size (run on a GPU box)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slamrs_b200.slam import debug_resample
rng = np.random.default_rng(0)
rows = []
for n in (1024, 8192, 65536, 262144):
    for name, w in (("uniform", rng.random(n)), ("lognormal10", np.exp(rng.normal(-200, 10, n))),
                    ("peaked", np.concatenate([[1.0], np.exp(rng.normal(-38, 2, n - 1))]))):
        r = debug_resample(w, 0.37, timing=True)
        rows.append(dict(n=n, weights=name, k_weights_us=r["us"][0], k_resample_indices_us=r["us"][1], rounds=r["fold_rounds"],
                         heads=r["fold_heads"], fallback=r["fold_fallback"]))
        print(rows[-1], flush=True)
json.dump(rows, open(os.path.join("gpurun_out", "time_resample.json"), "w"), indent=1)
