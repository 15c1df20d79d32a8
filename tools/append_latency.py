"""Rank-one append vs full rebuild: wall time per update(), N training points, Ns samples."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gpyreg_b200 as g  # noqa: E402
from bench import benign_hyp, synth_data  # noqa: E402
from gpyreg_b200.spec import ModelSpec  # noqa: E402


def main():
    for N, Ns, D in [(500, 8, 4), (1000, 16, 6), (2000, 16, 6), (2000, 64, 6)]:
        X, y = synth_data(N + 40, D, seed=0)
        spec = ModelSpec(D=D, cov_kind=0, degree=0, ard=True, mean_kind=1, noise_params=(1, 0, 0))
        hyp = benign_hyp(spec, Ns, y, seed=1)
        gp = g.GP(D, g.covariance_functions.SquaredExponential(), g.mean_functions.ConstantMean(),
                  g.noise_functions.GaussianNoise(constant_add=True))
        gp.update(X_new=X[:N], y_new=y[:N].reshape(-1, 1), hyp=hyp)
        gp.predict(X[:2])
        t_app = []
        for i in range(N, N + 20):
            t0 = time.perf_counter()
            gp.update(X_new=X[i:i + 1], y_new=y[i:i + 1].reshape(-1, 1))
            t_app.append(time.perf_counter() - t0)
        t_full = []
        for _ in range(3):
            t0 = time.perf_counter()
            gp.update(hyp=hyp)
            gp.predict(X[:2])           # the rebuild leaves W = L^-1 to the first predict
            t_full.append(time.perf_counter() - t0)
        print(f"N={N} Ns={Ns}: append median {1e3*np.median(t_app):.3f} ms  max {1e3*max(t_app):.3f} ms; "
              f"full rebuild+W {1e3*np.median(t_full):.3f} ms", flush=True)


main()
