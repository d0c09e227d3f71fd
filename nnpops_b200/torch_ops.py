"""Build / load libNNPOpsPyTorch.so: the reference's torch custom-op surface (six namespaces) on top of the B200 C ABI.

    python nnpops_b200/torch_ops.py            # build in-tree (needs g++ and the torch headers; no GPU)
    from nnpops_b200 import torch_ops; torch_ops.load()   # torch.ops.load_library(...)

The library name and the registered names are the reference's (src/pytorch/__init__.py:14), so pointing the reference's
`NNPOps` package at this file makes its Python wrappers run on the B200 kernels unchanged.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "torch", "nnpops_torch_ops.cpp")
LIB = os.path.join(HERE, "libNNPOpsPyTorch.so")
CORE = os.path.join(HERE, "libnnpops_b200.so")


def build(force=False):
    deps = [SRC, os.path.join(HERE, "..", "include", "nnpops_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    import torch
    from torch.utils import cpp_extension as ext
    inc = []
    for p in ext.include_paths() + ["/usr/local/cuda/include"]:
        inc += ["-isystem", p]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", abi, "-DTORCH_API_INCLUDE_EXTENSION_H", SRC, "-o", LIB] + inc +
           ["-L" + tlib, "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda", "-L" + HERE, "-lnnpops_b200",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libNNPOpsPyTorch.so failed:\n" + r.stderr[-4000:])
    return LIB


_loaded = False


def load():
    """Register the op namespaces with torch (idempotent).  Fails loudly when the library has not been built."""
    global _loaded
    if _loaded:
        return
    import torch
    if not os.path.exists(LIB):
        raise ImportError("nnpops_b200: %s is missing; build it with `python nnpops_b200/torch_ops.py`" % LIB)
    torch.ops.load_library(LIB)
    _loaded = True


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
