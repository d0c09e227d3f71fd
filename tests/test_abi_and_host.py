"""CPU-side checks: the C-ABI library loads and exports every symbol that include/nnpops_b200.h declares (no compute calls
without a GPU), compute entry points fail loudly without a device, and host-side helpers behave."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nnpops_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"NNPOPS_API\s+[\w\s\*]+?\b(nnpops_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import nnpops_b200._lib as L
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L.lib, n), "missing export: " + n
    assert L.lib.nnpops_abi_version() == 1


def test_header_cites_reference_interfaces():
    text = open(HEADER).read()
    for ref in ("ANISymmetryFunctions.h", "SymmetryFunctions.cpp", "BatchedNN.cpp", "CFConv.h", "getNeighborPairsCUDA.cu", "pmeCUDA.cu"):
        assert ref in text


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    """No CPU fallback: creating any compute object without a CUDA device is an error, never a silent slow path."""
    import nnpops_b200._lib as L
    h = ctypes.c_void_p()
    sp = np.zeros(3, np.int32); rf = np.array([[1.0, 1.0]], np.float32); af = np.array([[1.0, 1.0, 1.0, 1.0]], np.float32)
    rc = L.lib.nnpops_ani_create(ctypes.byref(h), 3, 1, 5.0, 3.5, L.ptr(sp), 1, L.ptr(rf), 1, L.ptr(af), 1, 0, 0)
    assert rc != 0 and b"CUDA device" in L.lib.nnpops_last_error()
    from nnpops_b200.SymmetryFunctions import Holder
    hold = Holder(1, 5.0, 3.5, [1.0], [1.0], [1.0], [1.0], [1.0], [1.0], [0, 0, 0])
    with pytest.raises(RuntimeError):
        hold.forward(torch.zeros((3, 3)), None)
    from nnpops_b200.neighbors import getNeighborPairs
    with pytest.raises(RuntimeError):
        getNeighborPairs(torch.zeros((3, 3)), 1.0)


def test_pme_host_logic_without_gpu():
    """PME: argument validation as the reference raises it (pme.py:76-92, 151-160), shard ranges of the multi-GPU forms, and no CPU
    path for any of the compute methods -- including the fused direct-space form and the one-call form."""
    from nnpops_b200.pme import PME
    from nnpops_b200.pme.pme import shard_range
    excl = torch.full((4, 1), -1, dtype=torch.int32)
    for bad in ((0, 8, 8, 5, 1.0, 1.0), (8, 8, 8, 0, 1.0, 1.0), (8, 8, 8, 5, 0.0, 1.0), (8, 8, 8, 5, 1.0, -1.0)):
        with pytest.raises(ValueError):
            PME(*bad, excl)
    pme = PME(8, 8, 8, 5, 1.0, 1.0, excl)
    pos, q, box = torch.zeros((4, 3)), torch.zeros(4), torch.eye(3)
    with pytest.raises(ValueError):
        pme.compute_direct(pos, q, -1.0, box)
    with pytest.raises(ValueError):
        pme.compute_direct(torch.zeros((5, 3)), q, 1.0, box)          # lengths differ from the exclusion table
    with pytest.raises(ValueError):
        pme.energy_and_derivatives(pos, q, 0.0, box, world=1)
    for call in (lambda: pme.compute_direct(pos, q, 1.0, box), lambda: pme.compute_direct(pos, q, 1.0, box, max_num_pairs=10),
                 lambda: pme.compute_direct_sharded(pos, q, 1.0, box, emulate=(0, 2)), lambda: pme.compute_reciprocal(pos, q, box),
                 lambda: pme.energy_and_derivatives(pos, q, 1.0, box, world=1)):
        with pytest.raises(RuntimeError):
            call()                                                       # CPU tensors: "runs on CUDA devices only"
    # contiguous blocks that cover [0, n) exactly once, whatever the remainder
    for n, world in ((10, 3), (200000, 8), (5, 8), (0, 2)):
        blocks = [shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n and all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
        assert all(lo <= hi for lo, hi in blocks)


def test_function_tables_follow_reference_order():
    from nnpops_b200.SymmetryFunctions import function_tables
    r, a = function_tables([1, 2], [10, 20, 30], [5, 6], [7], [0.1, 0.2], [0.5, 1.5, 2.5])
    assert r.shape == (6, 2) and a.shape == (12, 4)
    assert r[1].tolist() == [1, 20] and r[3].tolist() == [2, 10]           # k = iEta * |ShfR| + iShfR
    assert np.allclose(a[4], [5, 0.2, 7, 1.5])                             # m = ((iEta*|Zeta| + iZeta)*|ShfA| + iShfA)*|ShfZ| + iShfZ
    assert np.allclose(a[6], [6, 0.1, 7, 0.5])


def test_pack_network_params_layout():
    from nnpops_b200.OptimizedTorchANI import pack_network_params
    rng = np.random.default_rng(0)
    nets = [[[(rng.standard_normal((4, 6)), rng.standard_normal(4)), (rng.standard_normal((1, 4)), rng.standard_normal(1))] for _ in range(2)]
            for _ in range(3)]
    dims, params = pack_network_params(nets)
    assert dims.tolist() == [[6, 4, 1]] * 3
    assert params.size == 3 * 2 * (4 * 6 + 4 + 4 + 1)
    assert np.allclose(params[:24], nets[0][0][0][0].astype(np.float32).ravel())


def test_pme_moduli_match_reference_construction():
    """bspline_moduli restates pme.py:94-129; compare with an independent fp64 evaluation."""
    from nnpops_b200.pme.pme import bspline_moduli
    import importlib.util
    spec = importlib.util.spec_from_file_location("o", os.path.join(ROOT, "oracle", "neighbors_pme_oracle.py"))
    o = importlib.util.module_from_spec(spec); spec.loader.exec_module(o)
    for order in (4, 5):
        got = bspline_moduli(order, (14, 15, 16))
        ref = o.pme_moduli(order, (14, 15, 16))
        for g, r in zip(got, ref):
            assert np.allclose(g.numpy(), r, rtol=2e-5, atol=1e-7)


def test_angular_backward_enumeration_is_collision_free():
    """The angular backward kernel (csrc/ani_angular_v2.cu) walks the n (n - 1) / 2 neighbour pairs of a centre in flat rotation order,
    step = min(32, 2 n - 2) pairs per warp iteration, and accumulates the two forces of a pair with plain (non-atomic) read-modify-writes
    into arrays indexed by (parity of d, slot).  That is only correct if, within one iteration, no two lanes share (parity, first slot)
    or (parity, second slot) -- and the enumeration must visit every unordered pair exactly once.  Exhaustive check for n <= 128."""
    for n in range(2, 129):
        total = n * (n - 1) // 2
        step = min(32, 2 * n - 2)
        seen = set()
        for q0 in range(0, total, step):
            first, second = set(), set()
            for lane in range(32):
                q = q0 + lane
                if lane >= step or q >= total:
                    continue
                d, a = divmod(q, n)
                b = a + d + 1
                if b >= n:
                    b -= n
                assert 0 <= b < n and a != b
                pair = frozenset((a, b))
                assert pair not in seen
                seen.add(pair)
                ka, kb = (d & 1, a), (d & 1, b)
                assert ka not in first and kb not in second, (n, q0, lane)
                first.add(ka)
                second.add(kb)
        assert len(seen) == total
