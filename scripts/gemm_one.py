import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.zeros(1, device="cuda")
from nnpops_b200._lib import lib, check
lib.nnpops_debug_gemm_bench.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_double)]
m, n, k, b, mode = [int(x) for x in sys.argv[1:6]]
ms = C.c_double(0)
check(lib.nnpops_debug_gemm_bench(m, n, k, b, mode, 0, 3, C.byref(ms)))
print(ms.value)
