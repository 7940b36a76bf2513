"""Helpers to import the UNMODIFIED reference: thin alias of baseline/ref_loader.py (vendored ``baseline/_ref`` on the
GPU box, ``/root/reference`` in the build container).  Test infrastructure only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline.ref_loader import (activate, available, build_reference_model, load_config, ref_root)  # noqa: E402,F401

REF_ROOT = ref_root()
