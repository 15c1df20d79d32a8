"""Latency of the diagonal-tile kernel in a dependent chain (GPB_DIAG_DBG=1 adds the phase clock stamps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpyreg_b200 import Engine

e = Engine(0)
for rhs in (True, False):
    print("diag_kernel, forward solve %s: %.2f us per launch" % ("on" if rhs else "off", e.debug_diag_bench(400, rhs)), flush=True)
