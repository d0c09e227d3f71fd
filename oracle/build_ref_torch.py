"""TEST INFRASTRUCTURE: compile the UNMODIFIED reference torch extension (CPU sources of /root/reference/src/setup.py:36-47) from where
the sources lie into oracle/_ref/libNNPOpsPyTorch_refcpu.so (git-ignored, travels with gpurun).  Used only to produce TorchScript
archives SAVED BY THE REFERENCE (scripts/make_ref_archives.py -> tests/golden/ref_saved/) -- the fixtures of the pickle
byte-compatibility test -- never loaded next to libNNPOpsPyTorch.so (same op namespaces).

    python oracle/build_ref_torch.py        # ~1 min, no GPU; needs /root/reference"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref", "libNNPOpsPyTorch_refcpu.so")
SOURCES = ["ani/CpuANISymmetryFunctions.cpp", "pytorch/BatchedNN.cpp", "pytorch/CFConv.cpp", "pytorch/CFConvNeighbors.cpp",
           "pytorch/SymmetryFunctions.cpp", "pytorch/neighbors/getNeighborPairsCPU.cpp", "pytorch/neighbors/neighbors.cpp",
           "pytorch/pme/pmeCPU.cpp", "pytorch/pme/pme.cpp", "schnet/CpuCFConv.cpp"]
INCLUDES = ["ani", "pytorch", "pytorch/common", "pytorch/neighbors", "pytorch/pme", "schnet"]


def build(force=False):
    if os.path.exists(OUT) and not force:
        return OUT
    if not os.path.isdir(SRC):
        return None
    from torch.utils import cpp_extension as ext
    bdir = os.path.join(HERE, "_ref", "torch_build")
    os.makedirs(bdir, exist_ok=True)
    ext.load(name="NNPOpsPyTorchRefCpu", sources=[os.path.join(SRC, s) for s in SOURCES],
             extra_include_paths=[os.path.join(SRC, i) for i in INCLUDES], build_directory=bdir, is_python_module=False, verbose=False)
    shutil.copy(os.path.join(bdir, "NNPOpsPyTorchRefCpu.so"), OUT)
    shutil.rmtree(bdir, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
