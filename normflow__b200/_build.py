"""In-tree build of libnormflow_b200.so with nvcc for sm_100a.

    python -m normflow__b200._build          # build if sources are newer than the library

nvcc cross-compiles without a GPU.  The library lands in normflow__b200/lib/ (git-ignored,
shipped to the GPU box with the snapshot).
"""

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libnormflow_b200.so")
SOURCES = ["nfk_kernels.cu", "nfk_conv.cu", "nfk_fused.cu", "nfk_fused_tc.cu", "nfk_psd.cu", "nfk_knots.cu", "nfk_wgrad_tc.cu", "nfk_convnd_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libnormflow_b200.so")
    return exe


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "normflow_b200.h"))
    return max(os.path.getmtime(d) for d in deps) > os.path.getmtime(LIB)


def build(force=False, verbose=False):
    """Compile every .cu to an object (in parallel) and link the shared library."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        extra = os.environ.get("NFK_NVCC_EXTRA", "").split()      # tuning experiments, e.g. -DNFK_TC_COMPUTE_WARPS=12
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs]   # static cudart (nvcc default)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libnormflow_b200.so failed")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
