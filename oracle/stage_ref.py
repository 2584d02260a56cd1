"""Stage the UNMODIFIED reference into oracle/_ref/  --  TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

    python -m oracle.stage_ref            # needs /root/reference (the build container)

jkomijani/normflow_ is pure Python (no compiled sources), so "building" the real reference
for the CPU arm is a file copy: /root/reference/src  -> oracle/_ref/normflow_ref/  (an
importable package: the reference uses relative imports only, src/__init__.py:4-13) and
/root/reference/examples/*.py -> oracle/_ref/examples/.  oracle/_ref/ is git-ignored (no
reference source enters the history) but not gpurun-ignored, so the staged copy travels to
the GPU box like the built .so.  Consumers: `bench.py --impl reference` / `cpu_baseline`
(kind "reference": the reference's own `posterior.sample__`, `fit.step`, `mcmc.sample` on the
host cores) and tests/test_reference_examples.py (runs the byte-identical example scripts
against the `normflow` alias package).  Nothing under normflow__b200/ imports it.
"""

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DEST = os.path.join(HERE, "_ref")


def stage(force=False):
    """Copy the reference package and examples; returns the destination or None when
    /root/reference is absent (the GPU box: the staged copy of the snapshot is used as is)."""
    src = os.path.join(REF_ROOT, "src")
    if not os.path.isdir(src):
        return DEST if os.path.isdir(os.path.join(DEST, "normflow_ref")) else None
    pkg = os.path.join(DEST, "normflow_ref")
    if force or not os.path.isdir(pkg):
        if os.path.isdir(pkg):
            shutil.rmtree(pkg)
        shutil.copytree(src, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    ex = os.path.join(DEST, "examples")
    os.makedirs(ex, exist_ok=True)
    for name in os.listdir(os.path.join(REF_ROOT, "examples")):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(REF_ROOT, "examples", name), os.path.join(ex, name))
    return DEST


def import_reference():
    """Import the staged reference as `normflow_ref` (CPU, its default float64).  The caller must
    hide the GPUs first (CUDA_VISIBLE_DEVICES="") if the reference is to run on the host cores: it
    makes CUDA the default device at import when one is visible (src/device/__init__.py:7-13)."""
    import numpy
    if not hasattr(numpy, "product"):
        numpy.product = numpy.prod          # numpy >= 2 dropped the alias the reference still calls
    if not os.path.isdir(os.path.join(DEST, "normflow_ref")):
        raise ImportError("oracle/_ref/normflow_ref is missing: run `python -m oracle.stage_ref` "
                          "in the build container")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import normflow_ref
    return normflow_ref


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
