"""Development aid: decompose the time of one tcgen05 GEMM shape by switching parts of the kernel off (NNPOPS_GEMM_DBG bits:
1 no epilogue stores, 2 no activation loads, 4 no epilogue math, 8 no A loads, 16 no MMA issue).  The switches exist only in a library built
with NNPOPS_BUILD_DEFINES=-DNNPOPS_GEMM_DEBUG python -m nnpops_b200.build --force."""
import ctypes as C, os, sys, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    torch.zeros(1, device="cuda")
    from nnpops_b200._lib import lib, check
    lib.nnpops_debug_gemm_bench.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_double)]
    H = 33334
    shapes = [("H L0f", H, 2048, 128, 1, 1), ("H L1f", H, 192, 256, 8, 1), ("H L2f", H, 192, 192, 8, 3), ("H dZ1", H, 192, 192, 8, 2),
              ("H dZ0", H, 256, 192, 8, 2), ("H dX", H, 128, 2048, 1, 0)]
    out = []
    for name, m, n, k, b, mode in shapes:
        ms = C.c_double(0)
        check(lib.nnpops_debug_gemm_bench(m, n, k, b, mode, 1, 20, C.byref(ms)))
        out.append("%s %.1f" % (name, ms.value * 1e3))
    print("dbg=%-3s" % os.environ.get("NNPOPS_GEMM_DBG", "0"), " | ".join(out), flush=True)
else:
    for dbg in [int(x) for x in os.environ.get("DBGS", "0,7,24,15,23").split(",")]:
        env = dict(os.environ, NNPOPS_GEMM_DBG=str(dbg))
        subprocess.run([sys.executable, __file__, "x"], env=env)
