"""Debug aid: GP.predict_full against the reference (baseline/_ref) over sizes and noise models."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import gpyreg as ref
import gpyreg_b200 as ours
from bench import benign_hyp, synth_data
from gpyreg_b200.spec import ModelSpec


def make(gpr, D, npar):
    return gpr.GP(D=D, covariance=gpr.covariance_functions.SquaredExponential(), mean=gpr.mean_functions.NegativeQuadratic(),
                  noise=gpr.noise_functions.GaussianNoise(npar[0] == 1, npar[1] >= 1, npar[1] == 2, npar[2] == 1))


for N, M, D, npar in [(60, 11, 2, (1, 0, 0)), (60, 200, 2, (1, 0, 0)), (300, 11, 2, (1, 0, 0)), (300, 200, 2, (1, 0, 0)),
                      (500, 1000, 1, (1, 0, 0)), (500, 1000, 1, (1, 2, 1)), (300, 200, 2, (1, 2, 1)), (60, 11, 2, (1, 2, 1))]:
    spec = ModelSpec(D=D, cov_kind=0, ard=True, mean_kind=2, noise_params=npar)
    X, y = synth_data(N, D, 0)
    s2 = np.full((N, 1), 0.01) if npar[1] else None
    hyp = benign_hyp(spec, 3, y, 1)
    gr, go = make(ref, D, npar), make(ours, D, npar)
    gr.update(X_new=X, y_new=y, s2_new=s2, hyp=hyp)
    go.update(X_new=X, y_new=y, s2_new=s2, hyp=hyp)
    Xs = np.random.default_rng(3).uniform(-3, 3, (M, D))
    for an in (False, True):
        kw = dict(s2_star=0.01) if npar[1] else {}
        mr, cr = gr.predict_full(Xs, add_noise=an, **kw)
        mo, co = go.predict_full(Xs, add_noise=an, **kw)
        pr = gr.predict(Xs, add_noise=an, separate_samples=True, **kw)
        po = go.predict(Xs, add_noise=an, separate_samples=True, **kw)
        print(f"N={N} M={M} D={D} noise={npar} add_noise={an}: mu diff {np.max(np.abs(mr - mo)):.2e} cov diff {np.max(np.abs(cr - co)):.2e} "
              f"(|cov| {np.max(np.abs(cr)):.2e}); predict s2 diff {np.max(np.abs(pr[1] - po[1])):.2e}; "
              f"diag(cov) vs predict s2 (ours) {np.max(np.abs(np.einsum('iis->is', co) - po[1])):.2e}", flush=True)
