import os, sys
os.environ["GPB_DIAG_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpyreg_b200 import Engine
e = Engine(0)
rng = np.random.default_rng(0)
for n in (128, 128, 128):
    G = rng.standard_normal((n, n)); A = G @ G.T / n + np.eye(n)
    L, info = e.debug_potrf(A)
    print("info", info, "err", np.abs(L @ L.T - A).max())
