"""Moved to baseline/ref_loader.py (kept as an alias for the fixture generators)."""
from baseline.ref_loader import *  # noqa: F401,F403
from baseline.ref_loader import REF_ROOT, available, install_stubs, load_reference, reference_args  # noqa: F401
