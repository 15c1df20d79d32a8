"""B=1 latency of one nlZ / nlZ+grad evaluation (a slice sampler's unit of work).
LIB=<path> picks another build of the library for A/B runs."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyreg_b200._lib as _lib  # noqa: E402

if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
from bench import benign_hyp, synth_data  # noqa: E402
from gpyreg_b200 import Engine  # noqa: E402
from gpyreg_b200.spec import ModelSpec  # noqa: E402

eng = Engine(0)
spec = ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2)
for N in [int(v) for v in os.environ.get("NS", "1000,2000,5000").split(",")]:
    X, y = synth_data(N, spec.D, 0)
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, None)
    hyp = benign_hyp(spec, 8, y, 1)
    res = {}
    for grad in (False, True):
        for B in (1, 2, 3):
            ts = []
            for i in range(12):
                rows = hyp[(i % 2) * B:(i % 2) * B + B] if B < 3 else hyp[(i % 2) * 4:(i % 2) * 4 + B]       # alternate rows: no factor-cache hits
                t0 = time.perf_counter()
                eng.nlz_batch(rows, want_grad=grad)
                ts.append(time.perf_counter() - t0)
            res[(grad, B)] = 1e3 * float(np.median(ts[2:]))
    print(f"N={N}: nlZ B=1 {res[(False,1)]:.3f} ms, B=2 {res[(False,2)]:.3f} ms, B=3 {res[(False,3)]:.3f} ms; "
          f"nlZ+grad B=1 {res[(True,1)]:.3f} ms, B=2 {res[(True,2)]:.3f} ms, B=3 {res[(True,3)]:.3f} ms", flush=True)
