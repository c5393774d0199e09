"""hydra.utils stand-in: original cwd helpers (no chdir happens, so these are the cwd)."""
import os


def get_original_cwd():
    return os.getcwd()


def to_absolute_path(path):
    return path if os.path.isabs(path) else os.path.join(os.getcwd(), path)
