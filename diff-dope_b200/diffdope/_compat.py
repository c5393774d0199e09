"""Activate the in-repo stand-ins for hydra / omegaconf / icecream when (and only when) the real
packages are not importable."""
import importlib.util
import os
import sys

_COMPAT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")


def ensure():
    missing = [m for m in ("hydra", "omegaconf", "icecream") if importlib.util.find_spec(m) is None]
    if missing and os.path.isdir(_COMPAT_DIR) and _COMPAT_DIR not in sys.path:
        sys.path.append(_COMPAT_DIR)  # appended: a real installation always wins
    return missing
