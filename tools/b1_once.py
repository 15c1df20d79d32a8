"""One B=1 nlZ-only evaluation at cfg3 size (for an ncu launch list)."""
import os
import sys

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
B = int(os.environ.get("B", "1"))
hyp = benign_hyp(spec, max(B, 2), y, 1)
print(eng.nlz_batch(hyp[:B], want_grad=bool(int(os.environ.get("GRAD", "0"))))[0])
