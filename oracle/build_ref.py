"""ORACLE (test infrastructure): build / load `oracle/_ref/renderutils_plugin.so`, the reference's
own native plugin compiled from its unmodified sources where they lie
(`/root/reference/diffdope/c_src/{mesh.cu,common.cpp,torch_bindings.cpp}`, the same three files
and flags `diffdope/ops.py:56-91` passes to torch.utils.cpp_extension.load), for sm_100a.

It is CUDA code (the reference has no CPU path), so it can only *run* on the GPU box, where it is
the live reference for `xfm_points` / `xfm_vectors`. The rest of the
reference's hot path is nvdiffrast, which is not available (see oracle/nvdr.py).
Outputs only into oracle/_ref/ (git-ignored, shipped to the GPU box); no reference source is copied.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/diffdope/c_src"
NAME = "renderutils_plugin"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Compile if needed (needs /root/reference). Returns the path of the .so."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources not present and oracle/_ref not prebuilt")
    import torch.utils.cpp_extension as ext

    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    ext.load(
        name=NAME,
        sources=[os.path.join(SRC, f) for f in ("mesh.cu", "common.cpp", "torch_bindings.cpp")],
        extra_cflags=["-DNVDR_TORCH"],
        extra_cuda_cflags=["-DNVDR_TORCH", "-lineinfo"],
        build_directory=OUT,
        with_cuda=True,
        verbose=verbose,
        is_python_module=False,
    )
    return so_path()


def load():
    """Import the prebuilt plugin (GPU box). Returns the module or None if it is not there."""
    if not os.path.exists(so_path()):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(NAME, so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
