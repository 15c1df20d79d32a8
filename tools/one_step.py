"""One warm nlZ+gradient (or predict) step of a bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`: the launch list / DRAM-traffic table of exactly one step.
usage: python tools/one_step.py [cfg3|cfg2|cfg4|cfg5] [batch] [grad=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import benign_hyp, synth_data, workload
from gpyreg_b200 import Engine

wl = workload(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["B"]
grad = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) != 0
spec, N = wl["spec"], wl["N"]
X, y = synth_data(N, spec.D, 0)
eng = Engine(0)
eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
eng.set_data(X, y, None)
hyp = benign_hyp(spec, 2 * B, y, 1)
if wl["kind"] == "nlz":
    eng.nlz_batch(hyp[B:], want_grad=grad)               # warm: workspace, attributes, clocks
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    out = eng.nlz_batch(hyp[:B], want_grad=grad)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("nlZ[0]", out[0][0], "phases", eng.last_timings())
else:
    post = eng.posterior_batch(hyp[:B])
    Xs = np.random.default_rng(2).uniform(-3, 3, (wl["M"], spec.D))
    eng.predict(post, Xs[:8192])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    mu, s2 = eng.predict(post, Xs)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("mu[0]", mu[0, 0])
