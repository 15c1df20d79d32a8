"""Debug aid: the body of the reference's test_quadrature_with_noise under both implementations."""
import os, sys, types
import numpy as np
import scipy.stats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import gpyreg as ref
import gpyreg_b200 as ours


def body(gpr, iso):
    np.random.seed(12345)
    N, D, s2c = 500, 1, 0.01
    X = np.reshape(np.linspace(-15, 15, N), (-1, 1))
    s2 = np.full(X.shape, s2c)
    mu_N = 1000
    x_star = np.reshape(np.linspace(-15, 15, mu_N), (-1, 1))
    y = np.sin(X) + np.sqrt(s2) * scipy.stats.norm.ppf(np.random.random_sample(X.shape))
    y[y < 0] = -(np.abs(3 * y[y < 0]) ** 2)
    cov = (gpr.isotropic_covariance_functions.SquaredExponentialIsotropic() if iso
           else gpr.covariance_functions.SquaredExponential())
    gp = gpr.GP(D=D, covariance=cov, mean=gpr.mean_functions.NegativeQuadratic(),
                noise=gpr.noise_functions.GaussianNoise(constant_add=True, user_provided_add=True,
                                                        scale_user_provided=True,
                                                        rectified_linear_output_dependent_add=True))
    hyp, opt, _ = gp.fit(X=X, y=y, s2=s2, options={"n_samples": 10})
    print("  opt fun", opt.fun, "hyp[0]", np.round(hyp[0], 4))
    print("  sn2_mult", [p.sn2_mult for p in gp.posteriors], "L_chol", [p.L_chol for p in gp.posteriors])
    f_mu, f_cov = gp.predict_full(x_star, s2_star=s2c, add_noise=True)
    F_predict = sum(f_mu[i, 0] * scipy.stats.norm.pdf(x_star[i], scale=0.11) for i in range(mu_N)) * 30 / mu_N
    pdf_tmp = np.reshape(scipy.stats.norm.pdf(x_star, scale=0.1), (-1, 1))
    tmp = np.dot(pdf_tmp, pdf_tmp.T)
    F_predict_var = np.sum(np.sum(f_cov[:, :, 0] * tmp)) * (30 / mu_N) ** 2
    F_bayes, F_bayes_var = gp.quad(0, 0.1, compute_var=True)
    Fs, Fvs = gp.quad(0, 0.1, compute_var=True, separate_samples=True)
    print("  F_predict", F_predict, "F_predict_var", F_predict_var, "F_bayes", F_bayes, "F_bayes_var", F_bayes_var)
    print("  per-sample F", np.round(Fs, 4), "\n  per-sample var", np.round(Fvs, 5))
    return gp, hyp


for iso in (False, True):
    print("reference iso=%s" % iso)
    gr, hr = body(ref, iso)
    print("ours iso=%s" % iso)
    go, ho = body(ours, iso)
    print("  max |hyp diff|", np.max(np.abs(hr - ho)))
    # same hyp through both: quad and predict_full
    go.update(hyp=hr)
    Fo = go.quad(0, 0.1, compute_var=True, separate_samples=True)
    Fr = gr.quad(0, 0.1, compute_var=True, separate_samples=True)
    print("  same hyp: quad F diff", np.max(np.abs(Fo[0] - Fr[0])), "var diff", np.max(np.abs(Fo[1] - Fr[1])), "ref var", Fr[1])
