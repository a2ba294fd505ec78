"""Kept for ``make_golden.py``: the harness that imports the unmodified reference lives in oracle/ref_harness.py (the
reference arm of bench.py uses it too, through the vendored copy built by oracle/build_ref.py)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle"))
from ref_harness import *  # noqa: E402,F401,F403
from ref_harness import REFERENCE_ROOT, build_reference, import_reference, install_shims  # noqa: E402,F401
