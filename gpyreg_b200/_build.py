"""Build the CUDA library in-tree: gpyreg_b200/libgpyreg_b200.so (sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "api.cu")
LIB = os.path.join(HERE, "libgpyreg_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    d = os.path.join(HERE, "csrc")
    out = [os.path.join(d, f) for f in sorted(os.listdir(d))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "gpyreg_b200.h"))
    return out


def _source_hash():
    import hashlib
    h = hashlib.sha256()
    for path in _sources():
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale():
    """The library is stale when the sources it was built from changed (content hash kept
    next to the .so; file times do not survive the copy to the GPU box)."""
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".srchash"):
        return True
    with open(LIB + ".srchash") as f:
        return f.read().strip() != _source_hash()


def build_library(force=False, verbose=False):
    """Compile csrc/api.cu with nvcc. Returns the path of the shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libgpyreg_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB + ".tmp", SRC]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    with open(LIB + ".srchash", "w") as f:
        f.write(_source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
