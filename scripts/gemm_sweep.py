"""Per-shape timing of the tcgen05 GEMM (the 12 GEMMs of one ANI-2x water evaluation), resident-B vs streaming."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.zeros(1, device="cuda")
from nnpops_b200._lib import lib, check
lib.nnpops_debug_gemm_bench.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_double)]
H, O = 33334, 16666
shapes = [("H L0 fwd", H, 2048, 1024, 1, 1), ("H L1 fwd", H, 192, 256, 8, 1), ("H L2 fwd+final", H, 192, 192, 8, 3),
          ("O L0 fwd", O, 1536, 1024, 1, 1), ("O L1 fwd", O, 192, 192, 8, 1), ("O L2 fwd+final", O, 128, 192, 8, 3),
          ("H dZ1", H, 192, 192, 8, 2), ("H dZ0", H, 256, 192, 8, 2), ("H dX", H, 1024, 2048, 1, 0),
          ("O dZ1", O, 192, 128, 8, 2), ("O dZ0", O, 192, 192, 8, 2), ("O dX", O, 1024, 1536, 1, 0)]
tot = {0: 0.0, 1: 0.0}
for name, m, n, k, b, mode in shapes:
    row = []
    for streaming in (0, 1):
        ms = C.c_double(0)
        check(lib.nnpops_debug_gemm_bench(m, n, k, b, mode, streaming, 20, C.byref(ms)))
        row.append(ms.value); tot[streaming] += ms.value
    fl = 2.0 * m * n * k * b
    print("%-16s M=%6d N=%5d K=%5d b=%d mode=%d  resident %.1f us (%.0f TF alg)  streaming %.1f us (%.0f TF alg)" %
          (name, m, n, k, b, mode, row[0] * 1e3, fl / row[0] / 1e9, row[1] * 1e3, fl / row[1] / 1e9))
print("total resident-where-possible %.1f us, all streaming %.1f us" % (tot[0] * 1e3, tot[1] * 1e3))
