"""A/B of kernel builds: gradient-phase time of one warm step per library in build/var/ (built by hand with nvcc -D... -o build/var/lib_<tag>.so gpyreg_b200/csrc/api.cu)
(each build loaded in its own process through GPYREG_B200_LIB).  usage: python tools/grad_variants.py"""
import glob
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [("default", "")] + [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(ROOT, "build", "var", "lib_v*.so")))]
for wl, B in (("cfg3", 64), ("cfg2", 512)):
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["GPYREG_B200_LIB"] = path
        best = None
        for _ in range(1):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "one_step.py"), wl, str(B)], env=env,
                                 capture_output=True, text=True).stdout
            line = [l for l in out.splitlines() if "phases" in l]
            if not line:
                print(wl, name, "FAILED", out[-300:])
                break
            ph = line[0].split("phases", 1)[1].strip()
            g = eval(ph, {"np": np})
            gms = g["gradient"] if isinstance(g, dict) else g[4]
            best = gms if best is None else min(best, gms)
        print(f"{wl} B={B} {name}: gradient {best:.3f} ms   ({line[0][:60] if line else ''})", flush=True)
