"""Where a cfg3-shaped GP.fit spends its time: calls into the engine grouped by (batch, gradient)."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import gpyreg_b200 as g  # noqa: E402
from bench import synth_data  # noqa: E402
from gpyreg_b200.covariance_functions import Matern  # noqa: E402

N, D = 5000, 10
X, y = synth_data(N, D, 0)
np.random.seed(0)
gp = g.GP(D, Matern(5), g.mean_functions.NegativeQuadratic(), g.noise_functions.GaussianNoise(constant_add=True))
eng = gp.engine
acc = collections.defaultdict(lambda: [0, 0.0, 0, 0])
inner = eng.nlz_batch


def timed(hyp, want_grad=False):
    h0, m0 = eng.cache_stats()
    t0 = time.perf_counter()
    out = inner(hyp, want_grad=want_grad)
    dt = time.perf_counter() - t0
    h1, m1 = eng.cache_stats()
    a = acc[(np.atleast_2d(hyp).shape[0], bool(want_grad))]
    a[0] += 1
    a[1] += dt
    a[2] += h1 - h0
    a[3] += m1 - m0
    return out


eng.nlz_batch = timed
t0 = time.perf_counter()
opts = {"n_samples": 8, "init_N": 1024}
if os.environ.get("SPEC"):
    v = [int(t) for t in os.environ["SPEC"].split(",")]
    opts["speculate"] = v[0] if len(v) == 1 else v
hyp_s, _, res = gp.fit(X=X, y=y, options=opts)
print("speculate", opts.get("speculate", "default"), "sample checksum", float(np.sum(hyp_s)))
total = time.perf_counter() - t0
print(f"fit {total:.2f} s")
spent = 0.0
for (B, grad), (n, t, hits, miss) in sorted(acc.items()):
    spent += t
    print(f"B={B:5d} grad={int(grad)}: {n:5d} calls, {t:7.3f} s, {1e3 * t / n:8.3f} ms/call, cache hits {hits}, misses {miss}")
print(f"engine calls {spent:.2f} s, host side {total - spent:.2f} s")
