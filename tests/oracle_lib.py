"""ctypes access to the oracle (oracle/liboracle{32,64}.so) and, when present, the compiled
reference CPU classes (oracle/_ref/libnnpops_ref.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by nnpops_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_d = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C restatements (and oracle/_ref when /root/reference exists)."""
    need = force or not all(os.path.exists(os.path.join(ODIR, n)) for n in ("liboracle32.so", "liboracle64.so"))
    if need:
        subprocess.check_call(["make", "-s", "-C", ODIR, "oracle"])
    if os.path.isdir("/root/reference/src/ani") and (force or not os.path.exists(os.path.join(ODIR, "_ref", "libnnpops_ref.so"))):
        subprocess.check_call(["make", "-s", "-C", ODIR, "ref"])
    if os.path.isdir("/root/reference/src/ani") and (force or not os.path.exists(os.path.join(ODIR, "_ref", "libnnpops_ref_cuda.so"))):
        subprocess.check_call(["make", "-s", "-C", ODIR, "refcuda"])   # the reference's own CUDA kernels for sm_100a (GPU comparator)


_libs = {}


def lib(bits=32):
    if bits not in _libs:
        build()
        _libs[bits] = C.CDLL(os.path.join(ODIR, "liboracle%d.so" % bits))
    return _libs[bits]


def ref_lib():
    """The compiled reference (None when it was never built, e.g. fresh checkout without /root/reference)."""
    if "ref" not in _libs:
        build()
        p = os.path.join(ODIR, "_ref", "libnnpops_ref.so")
        _libs["ref"] = C.CDLL(p) if os.path.exists(p) else None
    return _libs["ref"]


def ref_cuda_lib():
    """The reference's CUDA classes compiled for sm_100a (None when never built).  GPU comparator only."""
    if "refcuda" not in _libs:
        build()
        p = os.path.join(ODIR, "_ref", "libnnpops_ref_cuda.so")
        l = C.CDLL(p) if os.path.exists(p) else None
        if l is not None:
            l.refcuda_ani_create.restype = C.c_void_p
            l.refcuda_ani_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            l.refcuda_ani_destroy.argtypes = [C.c_void_p]
            l.refcuda_ani_forward.argtypes = [C.c_void_p] * 5
            l.refcuda_ani_backward.argtypes = [C.c_void_p] * 4
            if hasattr(l, "refcuda_cfconv_create"):
                l.refcuda_cfconv_neighbors_create.restype = C.c_void_p
                l.refcuda_cfconv_neighbors_create.argtypes = [C.c_int, C.c_float, C.c_int]
                l.refcuda_cfconv_neighbors_destroy.argtypes = [C.c_void_p]
                l.refcuda_cfconv_neighbors_build.argtypes = [C.c_void_p] * 3
                l.refcuda_cfconv_neighbors_count.argtypes = [C.c_void_p]
                l.refcuda_cfconv_create.restype = C.c_void_p
                l.refcuda_cfconv_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 4
                l.refcuda_cfconv_destroy.argtypes = [C.c_void_p]
                l.refcuda_cfconv_compute.argtypes = [C.c_void_p] * 6
                l.refcuda_cfconv_backprop.argtypes = [C.c_void_p] * 8
        _libs["refcuda"] = l
    return _libs["refcuda"]


class RefCudaANI:
    """The reference's CudaANISymmetryFunctions on torch CUDA tensors (device pointers; src/ani/CudaANISymmetryFunctions.h:62-99)."""

    def __init__(self, species, n_species, rcr, rca, radial_fn, angular_fn, periodic, torchani=True):
        self.l = ref_cuda_lib()
        species = np.ascontiguousarray(species, np.int32)
        rf = np.ascontiguousarray(radial_fn, np.float32).reshape(-1, 2)
        af = np.ascontiguousarray(angular_fn, np.float32).reshape(-1, 4)
        self.n, self.ns, self.nr, self.na = len(species), n_species, len(rf), len(af)
        self.h = self.l.refcuda_ani_create(self.n, n_species, rcr, rca, int(periodic), species.ctypes.data, self.nr, rf.ctypes.data, self.na,
                                           af.ctypes.data, int(torchani))
        if not self.h:
            raise RuntimeError("reference CudaANISymmetryFunctions could not be constructed")

    def forward(self, pos, box):
        import torch
        radial = torch.empty((self.n, self.ns * self.nr), dtype=torch.float32, device=pos.device)
        angular = torch.empty((self.n, self.ns * (self.ns + 1) // 2 * self.na), dtype=torch.float32, device=pos.device)
        rc = self.l.refcuda_ani_forward(self.h, pos.data_ptr(), box.data_ptr() if box is not None else None, radial.data_ptr(), angular.data_ptr())
        assert rc == 0
        return radial, angular

    def backward(self, radial_grad, angular_grad):
        import torch
        out = torch.empty((self.n, 3), dtype=torch.float32, device=radial_grad.device)
        rc = self.l.refcuda_ani_backward(self.h, radial_grad.data_ptr(), angular_grad.data_ptr(), out.data_ptr())
        assert rc == 0
        return out

    def close(self):
        if self.h:
            self.l.refcuda_ani_destroy(self.h)
            self.h = None


class RefCudaCFConv:
    """The reference's CudaCFConvNeighbors + CudaCFConv on torch CUDA tensors (device pointers; src/schnet/CudaCFConv.h:37-190).
    w1: [width][numGaussians] in the reference's storage order, w2: [width][width] (host arrays)."""

    def __init__(self, n, width, n_gaussians, cutoff, periodic, gaussian_width, activation, w1, b1, w2, b2):
        self.l = ref_cuda_lib()
        self.n, self.w = n, width
        self.nb = self.l.refcuda_cfconv_neighbors_create(n, cutoff, int(periodic))
        arrs = [np.ascontiguousarray(a, np.float32) for a in (w1, b1, w2, b2)]
        self.h = self.l.refcuda_cfconv_create(n, width, n_gaussians, cutoff, int(periodic), gaussian_width, 0 if activation == "ssp" else 1,
                                              *[a.ctypes.data for a in arrs])
        if not self.nb or not self.h:
            raise RuntimeError("reference CudaCFConv could not be constructed")

    def build(self, pos, box):
        assert self.l.refcuda_cfconv_neighbors_build(self.nb, pos.data_ptr(), box.data_ptr() if box is not None else None) == 0

    def num_pairs(self):
        return self.l.refcuda_cfconv_neighbors_count(self.nb)

    def compute(self, pos, box, x, out=None):
        import torch
        out = torch.empty_like(x) if out is None else out
        assert self.l.refcuda_cfconv_compute(self.h, self.nb, pos.data_ptr(), box.data_ptr() if box is not None else None, x.data_ptr(), out.data_ptr()) == 0
        return out

    def backprop(self, pos, box, x, out_grad, in_grad=None, pos_grad=None):
        import torch
        in_grad = torch.empty_like(x) if in_grad is None else in_grad
        pos_grad = torch.empty_like(pos) if pos_grad is None else pos_grad
        assert self.l.refcuda_cfconv_backprop(self.h, self.nb, pos.data_ptr(), box.data_ptr() if box is not None else None, x.data_ptr(),
                                              out_grad.data_ptr(), in_grad.data_ptr(), pos_grad.data_ptr()) == 0
        return in_grad, pos_grad

    def close(self):
        if self.h:
            self.l.refcuda_cfconv_destroy(self.h)
            self.l.refcuda_cfconv_neighbors_destroy(self.nb)
            self.h = self.nb = None


def _opt(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype).ctypes.data_as(C.c_void_p)


def fn_tables(EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ):
    """Expand the TorchANI constant lists into the reference's function tables, in the order of
    /root/reference/src/pytorch/SymmetryFunctions.cpp:110-120."""
    radial = np.array([[e, s] for e in EtaR for s in ShfR], np.float32)
    angular = np.array([[e, s, z, t] for e in EtaA for z in Zeta for s in ShfA for t in ShfZ], np.float32)
    return radial, angular


def ani_forward(pos, species, n_species, rcr, rca, radial_fn, angular_fn, box=None, torchani=True, bits=32, impl="oracle"):
    pos = np.ascontiguousarray(pos, np.float32)
    species = np.ascontiguousarray(species, np.int32)
    radial_fn = np.ascontiguousarray(radial_fn, np.float32).reshape(-1, 2)
    angular_fn = np.ascontiguousarray(angular_fn, np.float32).reshape(-1, 4)
    n = pos.shape[0]
    nr, na = len(radial_fn), len(angular_fn)
    npairs = n_species * (n_species + 1) // 2
    if impl == "ref":
        bits = 32
    dt = np.float32 if bits == 32 else np.float64
    radial = np.zeros((n, n_species * nr), dt)
    angular = np.zeros((n, npairs * na), dt)
    boxp = _opt(box, np.float32)
    if impl == "ref":
        ref_lib().ref_ani(n, n_species, C.c_float(rcr), C.c_float(rca), int(torchani), species.ctypes.data_as(C.c_void_p), nr,
                          radial_fn.ctypes.data_as(C.c_void_p), na, angular_fn.ctypes.data_as(C.c_void_p),
                          pos.ctypes.data_as(C.c_void_p), boxp, radial.ctypes.data_as(C.c_void_p), angular.ctypes.data_as(C.c_void_p),
                          None, None, None)
    else:
        lib(bits).oracle_ani_forward(n, n_species, C.c_float(rcr), C.c_float(rca), int(torchani), species.ctypes.data_as(C.c_void_p), nr,
                                     radial_fn.ctypes.data_as(C.c_void_p), na, angular_fn.ctypes.data_as(C.c_void_p),
                                     pos.ctypes.data_as(C.c_void_p), boxp, radial.ctypes.data_as(C.c_void_p),
                                     angular.ctypes.data_as(C.c_void_p))
    return radial, angular


def ani_backward(pos, species, n_species, rcr, rca, radial_fn, angular_fn, radial_grad, angular_grad, box=None, torchani=True,
                 bits=32, impl="oracle"):
    pos = np.ascontiguousarray(pos, np.float32)
    species = np.ascontiguousarray(species, np.int32)
    radial_fn = np.ascontiguousarray(radial_fn, np.float32).reshape(-1, 2)
    angular_fn = np.ascontiguousarray(angular_fn, np.float32).reshape(-1, 4)
    n = pos.shape[0]
    nr, na = len(radial_fn), len(angular_fn)
    if impl == "ref":
        bits = 32
    dt = np.float32 if bits == 32 else np.float64
    rg = np.ascontiguousarray(radial_grad, dt)
    ag = np.ascontiguousarray(angular_grad, dt)
    out = np.zeros((n, 3), dt)
    boxp = _opt(box, np.float32)
    if impl == "ref":
        radial = np.zeros_like(rg)
        angular = np.zeros_like(ag)
        ref_lib().ref_ani(n, n_species, C.c_float(rcr), C.c_float(rca), int(torchani), species.ctypes.data_as(C.c_void_p), nr,
                          radial_fn.ctypes.data_as(C.c_void_p), na, angular_fn.ctypes.data_as(C.c_void_p),
                          pos.ctypes.data_as(C.c_void_p), boxp, radial.ctypes.data_as(C.c_void_p), angular.ctypes.data_as(C.c_void_p),
                          rg.ctypes.data_as(C.c_void_p), ag.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    else:
        lib(bits).oracle_ani_backward(n, n_species, C.c_float(rcr), C.c_float(rca), int(torchani), species.ctypes.data_as(C.c_void_p), nr,
                                      radial_fn.ctypes.data_as(C.c_void_p), na, angular_fn.ctypes.data_as(C.c_void_p),
                                      pos.ctypes.data_as(C.c_void_p), boxp, rg.ctypes.data_as(C.c_void_p),
                                      ag.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def cfconv(pos, width, n_gauss, cutoff, sigma, activation, w1, b1, w2, b2, inp, box=None, out_grad=None, bits=32, impl="oracle"):
    """-> output [n, W] or (output, input_grad, pos_grad, n_pairs) when out_grad is given."""
    pos = np.ascontiguousarray(pos, np.float32)
    n = pos.shape[0]
    if impl == "ref":
        bits = 32
    dt = np.float32 if bits == 32 else np.float64
    w1, b1, w2, b2 = (np.ascontiguousarray(a, np.float32) for a in (w1, b1, w2, b2))
    inp = np.ascontiguousarray(inp, dt)
    out = np.zeros((n, width), dt)
    og = None if out_grad is None else np.ascontiguousarray(out_grad, dt)
    ig = np.zeros((n, width), dt)
    pg = np.zeros((n, 3), dt)
    boxp = _opt(box, np.float32)
    act = {"ssp": 0, "tanh": 1}[activation]
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    if impl == "ref":
        npairs = C.c_longlong(0)
        ref_lib().ref_cfconv(n, width, n_gauss, C.c_float(cutoff), C.c_float(sigma), act, p(w1), p(b1), p(w2), p(b2), p(pos), boxp, p(inp),
                             p(out), p(og), p(ig), p(pg), C.byref(npairs))
        pairs = npairs.value
    else:
        fn = lib(bits).oracle_cfconv
        fn.restype = C.c_longlong
        pairs = fn(n, width, n_gauss, C.c_float(cutoff), C.c_float(sigma), act, p(w1), p(b1), p(w2), p(b2), p(pos), boxp, p(inp), p(out),
                   p(og), p(ig), p(pg))
    if out_grad is None:
        return out
    return out, ig, pg, pairs
