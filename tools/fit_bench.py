"""Wall-clock of a whole GP.fit on a PyVBMC-shaped problem (host drivers + GPU path)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpyreg_b200 as g
from gpyreg_b200.covariance_functions import Matern
from bench import synth_data
N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 6
X, y = synth_data(N, D, 0)
for chains in (1, 8):
    np.random.seed(0)
    gp = g.GP(D, Matern(5), g.mean_functions.NegativeQuadratic(), g.noise_functions.GaussianNoise(constant_add=True))
    t0 = time.perf_counter()
    hyp, opt, res = gp.fit(X=X, y=y, options={"n_samples": 16, "n_chains": chains})
    dt = time.perf_counter() - t0
    lp = np.mean([gp.log_posterior(h) for h in hyp])
    print(json.dumps({"N": N, "D": D, "n_chains": chains, "fit_s": round(dt, 2), "mean_log_post": float(lp),
                      "launches": gp.engine.launch_count()}))
