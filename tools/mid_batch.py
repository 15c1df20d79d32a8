"""nlZ / nlZ+grad time for mid-size batches at cfg3 size (look-ahead on/off comparison)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import benign_hyp, synth_data  # noqa: E402
from gpyreg_b200 import Engine  # noqa: E402
from gpyreg_b200.spec import ModelSpec  # noqa: E402

eng = Engine(0)
spec = ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2)
N = 5000
X, y = synth_data(N, spec.D, 0)
eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
eng.set_data(X, y, None)
hyp = benign_hyp(spec, 64, y, 1)
out = []
for B in (4, 8, 9, 16, 32):
    for grad in (False, True):
        ts = []
        for i in range(5):
            rows = hyp[(i % 2) * 32:(i % 2) * 32 + B]
            t0 = time.perf_counter()
            eng.nlz_batch(rows, want_grad=grad)
            ts.append(time.perf_counter() - t0)
        out.append(f"B={B} {'grad' if grad else 'nlz'} {1e3 * np.median(ts[1:]):.2f} ms")
print("; ".join(out))
