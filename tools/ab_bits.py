"""Bit-level A/B of two builds of the library: nlZ and gradient of a few problems as hex, computed in one process
per build (GPYREG_B200_LIB picks the build).  usage: python tools/ab_bits.py <other.so>"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from bench import benign_hyp, synth_data
from gpyreg_b200 import Engine
from gpyreg_b200.spec import ModelSpec
out = {}
eng = Engine(0)
for N, D, B in ((100, 2, 2), (900, 5, 3), (2000, 6, 2), (5000, 10, 1)):
    spec = ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2)
    X, y = synth_data(N, D, seed=0)
    hyp = benign_hyp(spec, B, y, seed=1)
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, None)
    nlz, dnlz, _, st = eng.nlz_batch(hyp, want_grad=True)
    nlz0 = eng.nlz_batch(hyp)[0]
    out[str(N)] = [v.hex() for v in nlz] + [v.hex() for v in nlz0] + [v.hex() for v in dnlz.ravel()] + [int(s) for s in st]
print(json.dumps(out))
"""


def run(lib):
    env = dict(os.environ)
    if lib:
        env["GPYREG_B200_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        print(r.stderr[-2000:])
        raise SystemExit(1)
    return json.loads(r.stdout.strip().splitlines()[-1])


a, b = run(None), run(sys.argv[1])
for k in a:
    same = a[k] == b[k]
    ndiff = sum(1 for u, v in zip(a[k], b[k]) if u != v)
    print(f"N={k}: {'bit-identical' if same else 'DIFFERENT in %d of %d values' % (ndiff, len(a[k]))}")
