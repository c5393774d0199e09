"""Test harness: run a Python script UNMODIFIED while absolute path prefixes it hard-codes are served from other directories.
Used for the reference's examples/run_bop_scene.py, whose BOP paths are its author's home directory (run_bop_scene.py:19-25).
PATH_MAP (environment) = JSON list of [prefix, replacement]; applied to builtins.open, cv2.imread and os.path.exists."""
import builtins
import json
import os
import runpy
import sys

import cv2

MAP = [(a.rstrip("/"), b) for a, b in json.loads(os.environ["PATH_MAP"])]


def remap(p):
    if isinstance(p, (str, os.PathLike)):
        s = os.fspath(p)
        for a, b in MAP:
            if s.startswith(a):
                return b + s[len(a):]
    return p


_open, _imread, _exists = builtins.open, cv2.imread, os.path.exists
builtins.open = lambda f, *a, **k: _open(remap(f), *a, **k)
cv2.imread = lambda f, *a, **k: _imread(remap(f), *a, **k)
os.path.exists = lambda f: _exists(remap(f))

script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
