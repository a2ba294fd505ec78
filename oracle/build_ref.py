"""Recipe for ``oracle/_ref``: a runnable copy of the reference's own Python implementation of the path, so that the
GPU box (which has no /root/reference) can time the REFERENCE ITSELF beside the CUDA engine (``bench.py --impl
reference``, ``cpu_baseline.kind = "reference"``).

The reference is pure Python (no native code), so "building" it is copying the four directories it
imports at run time -- vihds/, models/, specs/, data/ (+ LICENSE) -- from /root/reference into oracle/_ref/vi-hds/.
Nothing is modified; the shims the reference needs on this image (munch, seaborn/matplotlib stubs, the ragged-safe
merge_observations, the restated torchdiffeq fixed-grid steppers, the relay monkeypatch) live in oracle/ref_harness.py
and are applied at import time.  oracle/_ref/ is git-ignored (never part of the history) but not gpurun-ignored.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box: uses the travelled copy)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VIHDS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "vi-hds")
PARTS = ["vihds", "models", "specs", "data", "LICENSE", "requirements.txt"]


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "vihds")):
        if verbose:
            print("oracle/build_ref.py: %s not present; keeping %s as it is (%s)" % (
                SRC, DST, "present" if os.path.isdir(os.path.join(DST, "vihds")) else "ABSENT"))
        return os.path.isdir(os.path.join(DST, "vihds"))
    os.makedirs(DST, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.exists(s):
            shutil.copy2(s, d)
    if verbose:
        print("oracle/build_ref.py: %s -> %s" % (SRC, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
