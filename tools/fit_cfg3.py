"""config 3 (PyVBMC-shaped fit): Matern-5 ARD + NegativeQuadratic, N=5000, D=10, whole GP.fit."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpyreg_b200 as g
from gpyreg_b200.covariance_functions import Matern
from bench import synth_data
N, D = 5000, 10
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
init_N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
X, y = synth_data(N, D, 0)
np.random.seed(0)
gp = g.GP(D, Matern(5), g.mean_functions.NegativeQuadratic(), g.noise_functions.GaussianNoise(constant_add=True))
t0 = time.perf_counter()
hyp, opt, res = gp.fit(X=X, y=y, options={"n_samples": ns, "init_N": init_N})
dt = time.perf_counter() - t0
Xs = np.random.default_rng(2).uniform(-3, 3, (20000, D))
t1 = time.perf_counter()
mu, s2 = gp.predict(Xs)
dp = time.perf_counter() - t1
ytrue = np.sin(Xs.sum(1)) - 0.125 * (Xs ** 2).sum(1)
print(json.dumps({"N": N, "D": D, "n_samples": ns, "init_N": init_N, "fit_s": round(dt, 1),
                  "opt_nlZ": float(opt.fun), "nfev_opt": int(opt.nfev),
                  "slice_evals": None if res is None else int(len(res["f_vals"])),
                  "predict_20000_s": round(dp, 3), "rmse_vs_truth": float(np.sqrt(np.mean((mu[:, 0] - ytrue) ** 2))),
                  "launches": gp.engine.launch_count()}))
