"""config 4: one large GP (RationalQuadratic ARD, D=8) on one B200: nlZ + gradient timing and an
independent cross-check of nlZ (cuSOLVER float64 Cholesky through torch on a K built by the
plugin kernel).  usage: python tools/cfg4.py [N]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gpyreg_b200 import Engine
from gpyreg_b200.spec import ModelSpec
from bench import synth_data, benign_hyp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
spec = ModelSpec(D=8, cov_kind=2, ard=True, mean_kind=1)
X, y = synth_data(N, spec.D, 0)
hyp = benign_hyp(spec, 1, y, 1)
eng = Engine(0)
eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
eng.set_data(X, y, None)
out = {"N": N}
for grad in (False, True):
    warm = hyp.copy()
    warm[0, 0] += 1e-3                      # another length scale: the timed call cannot hit the factor cache
    eng.nlz_batch(warm, want_grad=grad)
    t0 = time.perf_counter()
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=grad)
    dt = time.perf_counter() - t0
    flops = N ** 3 * (1.0 if grad else 1 / 3)
    out["grad" if grad else "nlz"] = {"s": dt, "tflops_alg": flops / dt / 1e12, "phases_ms": eng.last_timings(),
                                      "nlZ": float(nlz[0]), "status": int(status[0]), "mult": float(mult[0])}
    if grad:
        out["dnlZ"] = dnlz[0].tolist()
# independent check: K from the plugin kernel, float64 Cholesky by cuSOLVER (torch)
if N <= 33000:
    h = hyp[0]
    K = eng.cov(spec.cov_kind, spec.degree, True, h[:spec.cov_n], X)
    sn2 = np.exp(2 * h[spec.cov_n])
    Kt = torch.from_numpy(K).cuda()
    del K
    Kt.diagonal().add_(sn2)
    r = torch.from_numpy(y[:, 0] - h[spec.cov_n + 1]).cuda()
    L = torch.linalg.cholesky(Kt)
    z = torch.linalg.solve_triangular(L, r[:, None], upper=False)[:, 0]
    ref = float(0.5 * (z @ z) + torch.log(L.diagonal()).sum() + 0.5 * N * np.log(2 * np.pi))
    out["nlZ_crosscheck"] = ref
    out["nlZ_rel_diff"] = abs(ref - out["nlz"]["nlZ"]) / abs(ref)
print(json.dumps(out))
