"""Multi-GPU check of the sharded PME path (run under torchrun, one rank per GPU): BASELINE config 5 -- 200 000 charges, 128^3 grid,
order 5 -- reciprocal energy / derivatives of pme_reciprocal_sharded against the single-GPU op on every rank, and timings."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from systems import lattice
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nnpops_nccl.%h.%p.log")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nnpops_b200.pme import PME
from nnpops_b200.pme.pme import pme_reciprocal, pme_reciprocal_sharded
n = 200000
pos, L = lattice(n, 0.2154, 0.3, 5005)
q = np.random.default_rng(5006).uniform(-0.5, 0.5, n).astype(np.float32); q -= q.mean()
box = torch.tensor(np.eye(3, dtype=np.float32) * L, device="cuda")
pme = PME(128, 128, 128, 5, 2.92, 138.935, torch.zeros((n, 0), dtype=torch.int32))
mod = [m.cuda() for m in pme.moduli]
def run(fn):
    p = torch.tensor(pos, device="cuda", requires_grad=True); c = torch.tensor(q, device="cuda", requires_grad=True)
    e = fn(p, c); e.backward(); return e.item(), p.grad, c.grad
single = lambda p, c: pme_reciprocal(p, c, box, 128, 128, 128, 5, 2.92, 138.935, *mod)
sharded = lambda p, c: pme_reciprocal_sharded(p, c, box, 128, 128, 128, 5, 2.92, 138.935, *mod)
e0, g0, c0 = run(single); e1, g1, c1 = run(sharded)
err = [abs(e1 - e0) / abs(e0), float((g1 - g0).abs().max() / g0.abs().max()), float((c1 - c0).abs().max() / c0.abs().max())]
def timeit(fn, k=20):
    for _ in range(3): run(fn)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(k): run(fn)
    t1.record(); torch.cuda.synchronize()
    t = torch.tensor([t0.elapsed_time(t1) / k], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.cpu()[0])
ts, tm = timeit(single), timeit(sharded)
if rank == 0:
    print(json.dumps({"check": "pme_reciprocal_sharded", "n_gpus": world, "atoms": n, "grid": 128, "rel_err_energy_posgrad_chargegrad": err,
                      "ms_single_gpu_fwd_bwd": round(ts, 4), "ms_sharded_fwd_bwd": round(tm, 4)}))
dist.destroy_process_group()
