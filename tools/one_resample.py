import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slamrs_b200.slam import debug_resample
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
for _ in range(3):
    r = debug_resample(np.exp(rng.normal(-200, 5, n)), 0.37)
print(r["fold_rounds"], r["fold_heads"], r["fold_fallback"])
