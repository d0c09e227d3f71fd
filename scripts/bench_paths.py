"""Timing of the other hot-path rows of SURVEY.md section 8 at the BASELINE.json config sizes (one JSON line per path):
getNeighborPairs + PME direct/reciprocal (config 5: 200 000 charges, 128^3 grid, cutoff 0.9 nm) and SchNet CFConv + neighbour list
(config 4: 100 000 atoms periodic, width 128, 50 Gaussians, 6 interaction blocks).  CUDA events on the current stream, warm-up
first; achieved HBM GB/s from the algorithmic bytes of SURVEY section 8d.  CPU reference timings (oracle/_ref, bounded samples)
are taken with --cpu."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from systems import lattice, cubic_box

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


def bench_pme():
    from nnpops_b200.neighbors import getNeighborPairs
    from nnpops_b200.pme.pme import PME, pme_direct, pme_reciprocal
    n = 200000
    pos_np, L = lattice(n, 0.2154, 0.3, 5005)
    rng = np.random.default_rng(5)
    q_np = rng.uniform(-0.5, 0.5, n).astype(np.float32); q_np -= q_np.mean()
    pos = torch.tensor(pos_np, device="cuda"); q = torch.tensor(q_np, device="cuda"); box = torch.tensor(cubic_box(L), device="cuda")
    cap = 33_000_000
    out = {}

    def nb():
        out["r"] = getNeighborPairs(pos, 0.9, cap, box)
    ms = timed(nb, 5, 2)
    found = int(out["r"][3].item())
    gbs = (24.0 * found + 24.0 * n) / (ms * 1e-3) / 1e9
    emit(path="getNeighborPairs", config="200000 atoms periodic, cutoff 0.9 nm, compact list", ms=round(ms, 4), pairs=found,
         achieved_gbs=round(gbs, 1), peak_gbs=PEAK, frac=round(gbs / PEAK, 4), bound="hbm", unit_bytes="24 B per output pair + 24 B per atom")
    neighbors, deltas, distances, _ = out["r"]
    excl = torch.zeros((n, 0), dtype=torch.int32, device="cuda")
    ms = timed(lambda: pme_direct(pos, q, neighbors, deltas, distances, excl, 2.92, 138.935), 5, 2)
    gbs = 24.0 * found / (ms * 1e-3) / 1e9
    emit(path="pme_direct", config="pair list of the row above (forward computes energy and both derivatives)", ms=round(ms, 4), pairs=found,
         achieved_gbs=round(gbs, 1), peak_gbs=PEAK, frac=round(gbs / PEAK, 4), bound="hbm", unit_bytes="24 B per pair")
    from nnpops_b200.pme.pme import pme_direct_fused
    ms = timed(lambda: pme_direct_fused(pos, q, box, excl, 0.9, 2.92, 138.935), 10, 3)
    emit(path="pme_direct_fused (cell list + centre-owned erfc sum: replaces the two rows above)", config="200000 charges, cutoff 0.9 nm",
         ms=round(ms, 4), pairs=found, bound="fp32 issue", flops_note="every pair evaluated from both ends: 2 x %d erfc terms + the candidate tests of 27 cells" % found)
    for world in (2, 8):
        ms = timed(lambda: pme_direct_fused(pos, q, box, excl, 0.9, 2.92, 138.935, (world // 2, world)), 10, 3)
        emit(path="pme_direct_fused, one slab of %d (what one GPU of %d does)" % (world, world), ms=round(ms, 4))
    pme = PME(128, 128, 128, 5, 2.92, 138.935, torch.zeros((n, 0), dtype=torch.int32))
    pr = pos.clone().requires_grad_(True)
    qr = q.clone().requires_grad_(True)

    def recip():
        e = pme.compute_reciprocal(pr, qr, box)
        e.backward()
        pr.grad = None; qr.grad = None
    ms = timed(recip, 10, 3)
    emit(path="pme_reciprocal fwd+bwd", config="200000 charges, 128^3 grid, order 5 (spread, rFFT, convolve, irFFT, gather)", ms=round(ms, 4),
         note="L2/latency bound: 25 M grid RMW + 25 M grid reads per evaluation; algorithmic HBM floor ~16 us")


def bench_cfconv(cutoff):
    from nnpops_b200.CFConv import CFConv
    from nnpops_b200.CFConvNeighbors import CFConvNeighbors
    n, W, G = 100000, 128, 50
    pos_np, L = lattice(n, 2.154, 0.3, 4004)
    rng = np.random.default_rng(9)
    pos = torch.tensor(pos_np, device="cuda", requires_grad=True)
    box = torch.tensor(cubic_box(L), device="cuda")
    convs = [CFConv(0.2, "ssp", torch.tensor(rng.normal(0, 0.1, (G, W)), dtype=torch.float32), torch.tensor(rng.normal(0, 0.1, W), dtype=torch.float32),
                    torch.tensor(rng.normal(0, 0.1, (W, W)), dtype=torch.float32), torch.tensor(rng.normal(0, 0.1, W), dtype=torch.float32)) for _ in range(6)]
    x = torch.tensor(rng.standard_normal((n, W)).astype(np.float32), device="cuda", requires_grad=True)
    nb = CFConvNeighbors(cutoff)
    ms_build = timed(lambda: nb.build(pos, box), 5, 2)
    pairs = nb.num_pairs()

    def step():
        nb.build(pos, box)
        y = x
        for c in convs:
            y = c(nb, pos, y)
        y.sum().backward()
        pos.grad = None; x.grad = None
    ms = timed(step, 3, 1)
    ref_flops = pairs * (46.9e3 + 93e3) * 6
    emit(path="CFConv x6 fwd+bwd + neighbour build", config="100000 atoms periodic, width 128, 50 Gaussians, sigma 0.2, cutoff %.0f A" % cutoff,
         ms_per_step=round(ms, 3), ms_neighbor_build=round(ms_build, 3), undirected_pairs=pairs,
         reference_algorithmic_tflop_per_step=round(ref_flops / 1e12, 2),
         equivalent_tflops=round(ref_flops / (ms * 1e-3) / 1e12, 1),
         note="the filter is tabulated (cubic Hermite), so the dense-layer flops of the reference formulation are not executed; "
              "the kernels are bound by L2 gathers of feature rows and table rows")


def cpu_rows():
    import oracle_lib as O
    rng = np.random.default_rng(1)
    n, W, G = 2000, 128, 50
    pos, L = lattice(n, 2.154, 0.3, 4004)
    w1 = rng.normal(0, 0.1, (W, G)); b1 = rng.normal(0, 0.1, W); w2 = rng.normal(0, 0.1, (W, W)); b2 = rng.normal(0, 0.1, W)
    x = rng.standard_normal((n, W)); go = np.ones((n, W))
    impl = "ref" if O.ref_lib() is not None else "oracle"
    t = time.perf_counter()
    _, _, _, pairs = O.cfconv(pos, W, G, 5.0, 0.2, "ssp", w1, b1, w2, b2, x, box=cubic_box(L), out_grad=go, impl=impl)
    dt = time.perf_counter() - t
    emit(path="CFConv CPU baseline", kind="reference" if impl == "ref" else "port", cores=1,
         sample="one layer fwd+bwd, %d atoms periodic, cutoff 5 A: %d pairs in %.2f s -> %.1f us per pair" % (n, pairs, dt, dt / pairs * 1e6))


if __name__ == "__main__":
    if "--cpu" in sys.argv:
        cpu_rows()
    else:
        bench_pme()
        if "--pme-only" in sys.argv:
            sys.exit(0)
        bench_cfconv(5.0)
        bench_cfconv(10.0)
