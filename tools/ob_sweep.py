import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpyreg_b200 import Engine
from gpyreg_b200.spec import ModelSpec
from bench import synth_data, benign_hyp
spec = ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2)
X, y = synth_data(5000, 10, 0)
ref = None
for ob in (1, 2, 4, 8):
    os.environ["GPB_OUTER_BLOCK"] = str(ob)
    e = Engine(0)
    e.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    e.set_data(X, y, None)
    for B in (1, 8, 64):
        hyp = benign_hyp(spec, B, y, 1)
        e.nlz_batch(hyp, want_grad=True)
        r = e.nlz_batch(hyp, want_grad=True)
        tm = e.last_timings()
        if B == 8:
            if ref is None: ref = r
            else: print("  max rel nlZ diff vs OB=1:", np.max(np.abs(r[0]-ref[0])/np.abs(ref[0])), "grad:", np.max(np.abs(r[1]-ref[1]))/np.max(np.abs(ref[1])))
        print(f"OB={ob} B={B}: factor {tm['factor']:.2f} ms  inverse {tm['inverse']:.2f}  total {tm['total']:.2f}  -> {B/tm['total']*1e3:.1f} evals/s")
    e.close()
