"""nlZ and nlZ+grad time of a batch at small N (few tile columns) for the current GPB_LEFT / GPB_OUTER_BLOCK setting.
usage: [GPB_LEFT=0|1] python tools/left_small.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import benign_hyp, synth_data  # noqa: E402
from gpyreg_b200 import Engine  # noqa: E402
from gpyreg_b200.spec import ModelSpec  # noqa: E402

eng = Engine(0)
spec = ModelSpec(D=6, cov_kind=1, degree=5, ard=True, mean_kind=2)
out = []
for N, B in ((500, 64), (1000, 32), (1000, 128), (1500, 64)):
    X, y = synth_data(N, spec.D, 0)
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, None)
    hyp = benign_hyp(spec, 2 * B, y, 1)
    for grad in (False, True):
        ts = []
        for i in range(7):
            rows = hyp[(i % 2) * B:(i % 2) * B + B]
            t0 = time.perf_counter()
            eng.nlz_batch(rows, want_grad=grad)
            ts.append(time.perf_counter() - t0)
        out.append(f"N={N} B={B} {'grad' if grad else 'nlz'} {1e3 * np.median(ts[2:]):.2f}")
print(os.environ.get("GPB_LEFT", "default"), "; ".join(out))
