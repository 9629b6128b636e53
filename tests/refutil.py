"""Locate the unmodified reference (baseline/_ref) for drop-in tests.  Never reads /root/reference."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(ROOT, "baseline", "stubs")


def import_pygrank():
    """Returns the pygrank module from baseline/_ref, or None when it was not installed."""
    if "pygrank" in sys.modules:
        return sys.modules["pygrank"]
    if not os.path.isdir(os.path.join(REF, "pygrank")):
        return None
    os.environ.setdefault("pygrankBackend", "numpy")   # avoids writing ~/.pygrank/config.json
    if "PYGRANK_KEEP_HOME" not in os.environ:
        os.environ["HOME"] = tempfile.mkdtemp()
    for p in (STUBS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pygrank
    return pygrank
