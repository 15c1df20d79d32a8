"""One cached (solve-only) B=3 nlZ evaluation at cfg3 size, for an ncu launch list / timing."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import benign_hyp, synth_data  # noqa: E402
from gpyreg_b200 import Engine  # noqa: E402
from gpyreg_b200.spec import ModelSpec  # noqa: E402

eng = Engine(0)
spec = ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2)
N = int(os.environ.get("N", "5000"))
X, y = synth_data(N, spec.D, 0)
eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
eng.set_data(X, y, None)
hyp = benign_hyp(spec, 1, y, 1).repeat(3, axis=0)
eng.nlz_batch(hyp)
ts = []
for it in range(int(os.environ.get("REPS", "1"))):
    hyp[:, -1] += 0.01 * (1 + it)          # a mean hyperparameter moves: factor cache hit
    hyp[1, -2] += 0.01
    t0 = time.perf_counter()
    out = eng.nlz_batch(hyp)
    ts.append(time.perf_counter() - t0)
print(out[0], eng.cache_stats(), "ms per cached call:", [round(1e3 * t, 3) for t in ts])
