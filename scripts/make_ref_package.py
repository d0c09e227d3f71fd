"""Assemble the reference's own Python package, unchanged, on top of this library -- the acceptance test of the drop-in boundary
("copy-free reuse of the reference *.py wrappers").

    python scripts/make_ref_package.py      # needs /root/reference; writes the git-ignored baseline/_ref/

baseline/_ref/NNPOps/       the reference's src/pytorch/{*.py, neighbors/*.py, pme/*.py}, byte for byte, except that the ONE line of
                            __init__.py that names the shared library (src/pytorch/__init__.py:14) points at
                            nnpops_b200/libNNPOpsPyTorch.so (the swap INTEGRATION.md describes)
baseline/_ref/ref_tests/    the reference's own pytest files for the paths in scope (TestNeighbors.py, TestPme.py, TestCFConv.py,
                            TestCFConvNeighbors.py), byte for byte
Nothing here is tracked by git or imported by the product; tests/test_reference_package_gpu.py runs it on the GPU box (the directory
travels with gpurun)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/pytorch"
OUT = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(REF):
        print("no reference tree at %s: nothing to do" % REF)
        return 0
    pkg = os.path.join(OUT, "NNPOps")
    tests = os.path.join(OUT, "ref_tests")
    for d in (pkg, tests):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    for name in sorted(os.listdir(REF)):
        if name.endswith(".py") and not name.startswith(("Test", "Benchmark")):
            shutil.copy(os.path.join(REF, name), os.path.join(pkg, name))
    for sub in ("neighbors", "pme"):
        os.makedirs(os.path.join(pkg, sub))
        for name in sorted(os.listdir(os.path.join(REF, sub))):
            if name.endswith(".py") and not name.startswith("Test"):
                shutil.copy(os.path.join(REF, sub, name), os.path.join(pkg, sub, name))
    init = open(os.path.join(pkg, "__init__.py")).read()
    old = "torch.ops.load_library(os.path.join(os.path.dirname(__file__), 'libNNPOpsPyTorch.so'))"
    assert init.count(old) == 1, "the reference's load_library line has changed"
    new = ("torch.ops.load_library(os.path.join(os.path.dirname(__file__), '..', '..', '..', 'nnpops_b200', 'libNNPOpsPyTorch.so'))"
           "  # the one-line swap (INTEGRATION.md section 1)")
    open(os.path.join(pkg, "__init__.py"), "w").write(init.replace(old, new))
    for rel in ("neighbors/TestNeighbors.py", "pme/TestPme.py", "TestCFConv.py", "TestCFConvNeighbors.py"):
        shutil.copy(os.path.join(REF, rel), os.path.join(tests, os.path.basename(rel)))
    shutil.copytree(os.path.join(REF, "molecules"), os.path.join(tests, "molecules"))
    print("assembled", pkg, "and", tests)
    return 0


if __name__ == "__main__":
    sys.exit(main())
