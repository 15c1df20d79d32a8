"""Latency of small predict calls (acquisition-function style use): N=1000, D=6."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpyreg_b200 import Engine
from gpyreg_b200.spec import ModelSpec
from bench import synth_data, benign_hyp
N, D = 1000, 6
spec = ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2)
X, y = synth_data(N, D, 0)
eng = Engine(0)
eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
eng.set_data(X, y, None)
for Ns in (1, 8, 64):
    post = eng.posterior_batch(benign_hyp(spec, Ns, y, 1))
    for M in (1, 16, 256, 4096):
        Xs = np.random.default_rng(2).uniform(-3, 3, (M, D))
        eng.predict(post, Xs)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            eng.predict(post, Xs)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"Ns": Ns, "M": M, "ms_per_call": round(dt * 1e3, 3), "launches_per_call": (eng.launch_count() - l0) // reps}))
    post.free()
