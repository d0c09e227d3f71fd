#!/usr/bin/env python
"""Headline benchmark: ANI-2x energy+force evaluations/s on a 50 000-atom periodic water box (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mlp tcgen05|simt] [--atoms 50000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one energy+force evaluation (AEV forward -> ensemble MLP forward/backward -> AEV backward) of one conformer.  With N
GPUs every rank evaluates its own conformers (independent conformers, no collective on the data path: weak scaling); the
timed region is bracketed by barrier + synchronize and the MAX over ranks is reported.  Rank 0 prints ONE JSON line.

--impl reference: the reference's own CPU implementation of the path on the host cores: AEV forward+backward by the unmodified
reference class CpuANISymmetryFunctions (oracle/_ref, compiled from /root/reference) plus, for the network, the per-species ATen
nn.Linear/CELU stand-in BASELINE.md section 3 prescribes (the reference's BatchedLinear replicates the weights per atom: 541 GB at
this size); a bounded sample per step, converted to the metric's unit with a measured a*N^2 + b*N model for the AEV and linearly
for the network, and the model is checked against ONE real 50 000-atom AEV evaluation in the same run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "ANI-2x energy+force evals/sec on 50k-atom box"
UNIT = "evals/s"
DEFAULT_MLP = "tcgen05"   # "simt" selects the fp32 validation GEMM
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # derived SIMT peak at max clock


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """NVML samples of the SM clock and the clock-event (throttle) reasons, polled from a thread DURING the timed region
    (the recipe's nvidia-smi clocks line, read through pynvml so that a 60 ms region still gets tens of samples)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.stop_flag, self.thread, self.h = index, [], 0, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:   # noqa: BLE001
            self.err = repr(ex)

    def _poll(self):
        while not self.stop_flag:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.stop_flag = False
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max, "reasons": reasons,
                "samples": len(self.sm)}


class StdoutToStderr:
    """NCCL prints its version banner (and any NCCL_DEBUG output) on the process's stdout; the contract is ONE JSON line there.
    File descriptor 1 is pointed at stderr for the duration of the run and restored just before the result is printed."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None

    def __exit__(self, *exc):
        self.restore()
        return False


def shard_conformers(total, rank, world):
    """Conformer ids evaluated by `rank`: BASELINE config 3 deals conformer c to rank c mod world (independent conformers, no
    collective on the data path)."""
    return [c for c in range(total) if c % world == rank]


def max_over_ranks(ms, dist, device):
    """MAX over ranks of a per-rank duration (the slowest rank defines the job's throughput)."""
    import torch
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.cpu()[0])


def make_conformer(n_atoms, c):
    from systems import lattice, cubic_box
    pos, L = lattice(n_atoms, 2.154, 0.3, 3000 + c)
    return pos, cubic_box(L)


def run_ours(args):
    import torch
    from systems import ANI2X, ANI2X_HIDDEN, ANI2X_ENSEMBLE, water_species
    from mlp_ref import random_networks
    from nnpops_b200.OptimizedTorchANI import FusedANI
    from nnpops_b200._lib import lib

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nnpops_nccl.%h.%p.log")      # NCCL logs to stdout otherwise: keep it to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.atoms
    species = water_species(n)
    nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
    model = FusedANI(7, 5.2, ANI2X["Rca"], ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species,
                     nets, mlp_impl=args.mlp, device="cuda:%d" % local)
    # pool of conformers per rank: conformer c of BASELINE config 3 goes to rank c mod world
    pool = 4
    confs = [make_conformer(n, c) for c in shard_conformers(pool * world, rank, world)]
    d_pos = [torch.tensor(p, device=dev) for p, _ in confs]
    d_box = [torch.tensor(b, device=dev) for _, b in confs]
    h_pos = [torch.tensor(p).pin_memory() for p, _ in confs]
    h_box = [torch.tensor(b).pin_memory() for _, b in confs]
    h_e = torch.zeros(1).pin_memory(); h_g = torch.zeros((n, 3)).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def launches():
        import ctypes
        c = ctypes.c_ulonglong(0)
        lib.nnpops_launch_count(ctypes.byref(c))
        return c.value

    # ---- device-resident throughput ("value") with per-stage CUDA events
    for i in range(args.warmup):
        model.energy_and_gradient(d_pos[i % pool], d_box[i % pool])
    sync_all()
    work = model.work()
    assert model.overflowed() == 0, "neighbour rows overflowed"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    model.timing_begin(args.steps)
    l0 = launches()
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        model.energy_and_gradient(d_pos[i % pool], d_box[i % pool])
    t1.record()
    sync_all()
    l1 = launches()
    stages, nrec = model.timing_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(t0.elapsed_time(t1), dist, dev)

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory; H2D + D2H inside the timed region)
    for i in range(min(args.warmup, 3)):
        model.energy_and_gradient_host(h_pos[i % pool].numpy(), h_box[i % pool].numpy(), h_e.numpy(), h_g.numpy())
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        model.energy_and_gradient_host(h_pos[i % pool].numpy(), h_box[i % pool].numpy(), h_e.numpy(), h_g.numpy())
    e1.record()
    sync_all()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), dist, dev)

    # ---- sustained leg: the same device-resident step for >= 3 s, so that the clocks under a real load are on record
    sustained = None
    if args.sustain > 0:
        per = max(ms_total / args.steps, 1e-3)
        reps = max(int(args.sustain * 1e3 / per), args.steps)
        s_sampler = ClockSampler(local)
        if rank == 0:
            s_sampler.start()
        sync_all()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(reps):
            model.energy_and_gradient(d_pos[i % pool], d_box[i % pool])
        s1.record()
        sync_all()
        s_ms = max_over_ranks(s0.elapsed_time(s1), dist, dev)
        sustained = {"value": round(world * reps / (s_ms / 1e3), 4), "unit": UNIT, "steps": reps, "seconds": round(s_ms / 1e3, 3),
                     "clocks": s_sampler.stop() if rank == 0 else None}

    # ---- MD-like leg: the callers' steady state is a time-stepping loop (small displacements per step), where a Verlet skin lets the
    # neighbour search reuse its candidate rows; the headline above rotates through unrelated conformers and rebuilds every step
    md = None
    if world == 1 and args.md_steps > 0:
        rng = np.random.default_rng(77)
        frames = [confs[0][0]]
        for _ in range(args.md_steps - 1):
            frames.append(frames[-1] + rng.normal(0.0, 0.01, frames[-1].shape).astype(np.float32))     # 0.01 A per coordinate per step
        h_frames = [torch.tensor(f).pin_memory() for f in frames]
        d_frames = [torch.tensor(f, device=dev) for f in frames]

        def md_run(skin):
            """the trajectory through the host-buffer entry point (H2D of the positions, CUDA-graph replay of the step, D2H of energy and
            forces: what a host-side integrator sees), then once more kernel by kernel for the per-stage events"""
            model.set_skin(skin)
            for f in h_frames[:3]:
                model.energy_and_gradient_host(f.numpy(), h_box[0].numpy(), h_e.numpy(), h_g.numpy())
            sync_all()
            r0 = model.skin_stats()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for f in h_frames:
                model.energy_and_gradient_host(f.numpy(), h_box[0].numpy(), h_e.numpy(), h_g.numpy())
            m1.record()
            sync_all()
            r1 = model.skin_stats()
            model.timing_begin(args.md_steps)
            for f in d_frames:
                model.energy_and_gradient(f, d_box[0])
            st, _ = model.timing_end()
            return args.md_steps / (m0.elapsed_time(m1) / 1e3), int(r1[0] - r0[0]), int(r1[1] - r0[1]), st["cells+rows"]

        v0, _, _, c0 = md_run(0.0)
        v1, rb, ru, c1 = md_run(0.4)
        model.set_skin(0.0)
        md = {"value": round(v1, 4), "unit": UNIT, "steps": args.md_steps, "skin_angstrom": 0.4, "value_without_skin": round(v0, 4),
              "api": "nnpops_ani_model_energy_grad_host (pinned host buffers, CUDA-graph replay)",
              "displacement": "Gaussian, sigma 0.01 A per coordinate per step, cumulative", "rebuild_steps": rb, "reuse_steps": ru,
              "cells_rows_ms_amortised_eager": round(c1, 4), "cells_rows_ms_rebuild_every_step_eager": round(c0, 4)}

    # ---- N > 1: the strong-scaling curve that matters -- ONE box over all GPUs by spatial decomposition with ghost halos
    box = None
    if world > 1 and not args.no_box:
        del model
        torch.cuda.empty_cache()
        box = measure_box(args, dist, dev, nets, rank, world, local)

    # ---- the PME row of the path (BASELINE config 5), on every N: one box of 200 000 charges, direct space sharded by slabs
    pme = None
    if not args.no_pme:
        try:
            pme = measure_pme(args, dist, dev, rank, world)
        except Exception as exc:   # noqa: BLE001  (an extra leg must not cost the bench line)
            pme = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * args.steps / (ms_total / 1e3)
    # ---- roofline of the dominant kernel and of the AEV kernels (algorithmic work: SURVEY.md section 8d)
    mlp_flops = 2.0 * work["mlp_flops_forward"]                      # forward + input-gradient backward
    mlp_ms = stages["mlp_fwd"] + stages["mlp_bwd"]
    tri, prs = work["triples"], work["radial_pairs"]
    kernels = []

    def kern(name, ms_, work_units, unit_work, bound, peak, unit):
        ach = work_units * unit_work / (ms_ * 1e-3) / (1e12 if unit == "TFLOP/s" else 1e9) if ms_ > 0 else 0.0
        kernels.append({"name": name, "ms": round(ms_, 4), "bound": bound, "achieved": round(ach, 3), "peak": round(peak, 1), "unit": unit,
                        "frac": round(ach / peak, 4)})
        return ach

    # the radial kernels run on an auxiliary stream concurrently with the angular ones: each pair is timed as one region on the
    # launching stream and rated against the sum of its algorithmic flops
    kern("angular forward chain (ani_angular_geo + seg_hist + seg_scatter + ani_angular_fwd_seg_kernel) || ani_radial_fwd_v2_kernel", stages["radial_fwd"] + stages["angular_fwd"], 1,
         tri * 146.0 + prs * 134.0, "fp32", FP32_PEAK_TFLOPS, "TFLOP/s")
    kern("ani_angular_bwd_v2_kernel || ani_radial_bwd_v2_kernel (+ grad_compact)", stages["radial_bwd"] + stages["angular_bwd"], 1,
         tri * 370.0 + prs * 212.0, "fp32", FP32_PEAK_TFLOPS, "TFLOP/s")
    kern("cell_list+ani_rows_kernel", stages["cells+rows"], n, 64.0 + 4.0 * 2 * prs / max(n, 1), "hbm", pk["hbm_gbs"], "GB/s")
    mlp_ach = mlp_flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    # DRAM traffic / tensor-pipe activity of the dominant kernel: a STATIC ncu --set full capture of the same command (profiles/), never
    # taken in this run (a number measured under a profiler is not a bench value)
    traffic, tensor_active, prof = None, None, ""
    import glob
    fused = bool(work.get("mlp_fused", False))
    profs = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_summary.json")))
    for cand in reversed(profs):
        tot = json.load(open(cand)).get("chain_step_totals" if fused else "gemm_step_totals")
        if tot and args.mlp == "tcgen05" and n == 50000:
            traffic, tensor_active, prof = tot.get("dram_bytes"), tot.get("tensor_pipe_active_pct"), cand
            break
    # burst vs sustained peak: the regime the clocks of THIS run show (the burst figure holds while the SM clock stays near its maximum)
    burst = clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"]
    mlp_peak = pk["bf16_tflops"] if burst else pk["bf16_tflops_sustained"]
    exec_flops = 2.0 * work["mlp_flops_forward_executed"]
    exec_ach = exec_flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    if args.mlp != "tcgen05":
        gemm_name = "gemm_tn_simt_kernel (fp32 validation GEMMs)"
    elif fused:
        gemm_name = "mlp_chain_kernel (ONE launch per evaluation: the six GEMMs of every 128-atom tile and ensemble member, forward + backward, activations on chip)"
    else:
        gemm_name = "gemm_tcgen05_kernel (twelve per-layer GEMM launches per evaluation, forward + backward)"
    roofline = {"kernel": gemm_name, "bound": "tensor",
                "achieved": round(mlp_ach, 3), "peak": mlp_peak, "unit": "TFLOP/s", "frac": round(mlp_ach / mlp_peak, 4), "traffic": traffic,
                "frac_executed": round(exec_ach / mlp_peak, 4),
                "frac_tensor_pipe": round(3.0 * exec_ach / mlp_peak, 4),
                "peak_source": "MEASURED_PEAKS.json %s (%s; SM clock median %s of %s MHz in the timed region)" %
                               ("bf16_tflops, burst" if burst else "bf16_tflops_sustained", pk["source"], clocks and clocks.get("sm_mhz"), clocks and clocks.get("sm_max_mhz")),
                "algorithmic_flops_per_step": mlp_flops, "ms_per_step": round(mlp_ms, 4), "executed_flops_per_step": exec_flops,
                "active_features": "%d of %d AEV columns (the others belong to species absent from the system and are identically zero: "
                                   "the first layer skips those 0*w products, results unchanged)" % (work["active_features"], work["aev_length"]),
                "traffic_note": "static capture, not measured in this run: dram__bytes_read+write of the kernel in %s; tensor pipe active %s %% "
                                "in the same capture; algorithmic bytes per evaluation = weights once + X in + dX out" % (os.path.relpath(prof, ROOT) if prof else None, tensor_active),
                "note": "achieved counts the ALGORITHMIC fp32 flops of SURVEY 8d (2*M*N*K un-padded on the full 1008-column AEV, forward + input-gradient "
                        "backward); frac_executed counts only the flops issued after dropping the structurally-zero AEV columns (lead with this one); "
                        "every product runs as 3 fp16 tensor-core MMAs (hi*hi, hi*lo, lo*hi): frac_tensor_pipe = 3 x frac_executed is the share of the "
                        "tensor pipe's measured bf16 rate in use"}
    fp32_measured = None
    try:
        import ctypes
        tf = ctypes.c_double(0.0)
        lib.nnpops_debug_fma_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
        if lib.nnpops_debug_fma_peak(ctypes.byref(tf)) == 0:
            fp32_measured = round(tf.value, 2)
    except Exception:   # noqa: BLE001
        pass
    for k in kernels:
        if k["bound"] == "fp32" and fp32_measured:
            k["peak_measured_fma"] = fp32_measured
            k["frac_of_measured_fma"] = round(k["achieved"] / fp32_measured, 4)
    out = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: ANI-2x, %d-atom periodic water box (jittered lattice, H/H/O), Rcr 5.2 A, Rca 3.5 A, "
                               "8-member ensemble with random ANI-2x-shaped weights, one conformer per step per GPU, "
                               "conformers split across GPUs with no collective" % n,
                   "atoms": n, "conformers_per_rank_pool": pool, "mlp_impl": args.mlp,
                   "l2": "per-step working set (MLP activations and gradients, > 1 GB) exceeds the 126 MB L2; positions rotate through %d conformers" % pool,
                   "triples_per_step": tri, "radial_pairs_per_step": prs},
        "e2e": {"value": round(world * args.steps / (e2e_ms / 1e3), 4), "unit": UNIT, "h2d_bytes_per_step": n * 12 + 36,
                "d2h_bytes_per_step": n * 12 + 4, "api": "nnpops_ani_model_energy_grad_host (C ABI, pinned host buffers)"},
        "gpu_launches": int(l1 - l0),
        "roofline": roofline,
        "kernels": kernels,
        "stage_ms": {k: round(v, 4) for k, v in stages.items()},
        "clocks": clocks,
        "sustained": sustained,
    }
    if box is not None:
        out["box"] = box
    if md is not None:
        out["md"] = md
    if pme is not None:
        out["pme"] = pme
    if fp32_measured:
        out["fp32_fma_peak_measured_tflops"] = fp32_measured
    if world == 1 and not args.no_cpu_baseline:
        out["forces_check"] = forces_check(nets, args.mlp, local)
        out["forces_max_abs_delta"] = out["forces_check"]["forces_max_abs_delta"]
        out["forces_rel"] = out["forces_check"]["forces_rel"]
        out["cpu_baseline"] = cpu_baseline(n, budget_s=20.0)
        try:
            out["gpu_comparator"] = gpu_comparator(local)
        except Exception as exc:   # noqa: BLE001  (a missing or failing comparator must not cost the bench line)
            out["gpu_comparator"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    args.quiet.restore()
    print(json.dumps(out), flush=True)
    if dist is not None:
        os.dup2(2, 1)   # the teardown may log again
        dist.destroy_process_group()


def gpu_comparator(local):
    """The reference's OWN CUDA kernels (src/ani/CudaANISymmetryFunctions.cu:186-670, compiled unmodified for sm_100a into
    oracle/_ref/libnnpops_ref_cuda.so by oracle/Makefile) next to ours on the same B200, same inputs, device-resident, CUDA events:
    AEV forward + backward (position gradient of a fixed upstream gradient) at the two BASELINE sizes its N x N neighbour table
    allows.  The reference has no MLP on the GPU that fits these sizes (its BatchedLinear replicates the weights per atom), so this
    compares the AEV half of the path only.  A reported comparison, never part of the headline value."""
    import torch
    import oracle_lib as O
    from systems import ANI2X, cubic_box, lattice, protein_species, water_species
    from nnpops_b200.SymmetryFunctions import Holder
    from nnpops_b200._lib import lib as L, ptr, current_stream
    if O.ref_cuda_lib() is None:
        return {"unavailable": "oracle/_ref/libnnpops_ref_cuda.so was not built (needs /root/reference at build time)"}
    dev = torch.device("cuda", local)
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    res = []
    for name, n, periodic in (("5000-atom non-periodic protein-like system (BASELINE configs[1] shape)", 5000, False),
                              ("40000-atom periodic water box (largest the reference's N x N table takes; configs[2] density)", 40000, True)):
        if periodic:
            pos, edge = lattice(n, 2.154, 0.3, 4000)
            species, box = water_species(n), cubic_box(edge)
        else:
            pos, _ = lattice(n, 2.0, 0.3, 11)
            species, box = protein_species(n), None
        p = torch.tensor(pos, device=dev)
        b = torch.tensor(box, device=dev) if box is not None else None
        ref = O.RefCudaANI(species, 7, 5.2, 3.5, rfn, afn, periodic)
        ours = Holder(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species.tolist())
        r1, a1 = ours.forward(p, b)          # creates the object, grows the rows if needed
        r0, a0 = ref.forward(p, b)
        g = torch.Generator(device=dev).manual_seed(3)
        rg = torch.randn(r0.shape, device=dev, generator=g)
        ag = torch.randn(a0.shape, device=dev, generator=g)
        d0 = ref.backward(rg, ag)
        d1 = ours.backward([rg, ag])
        torch.cuda.synchronize(dev)
        err = float((d1 - d0).abs().max() / d0.abs().max())
        err_a = float((a1 - a0).abs().max() / a0.abs().max())
        out3 = torch.empty((n, 3), dtype=torch.float32, device=dev)
        stream = current_stream(dev)

        def run_ref():
            ref.l.refcuda_ani_forward(ref.h, p.data_ptr(), b.data_ptr() if b is not None else None, r0.data_ptr(), a0.data_ptr())
            ref.l.refcuda_ani_backward(ref.h, rg.data_ptr(), ag.data_ptr(), out3.data_ptr())

        def run_ours():
            L.nnpops_ani_forward(ours._h, ptr(p), ptr(b), ptr(r1), ptr(a1), stream)
            L.nnpops_ani_backward(ours._h, ptr(rg), ptr(ag), ptr(out3), stream)

        times = {}
        for key, fn, reps in (("reference_cuda_ms", run_ref, 5 if n <= 5000 else 2), ("ours_ms", run_ours, 20)):
            fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            times[key] = e0.elapsed_time(e1) / reps
        ref.close()
        run = {"system": name, "atoms": n, "reference_cuda_ms": round(times["reference_cuda_ms"], 4), "ours_ms": round(times["ours_ms"], 4),
               "speedup": round(times["reference_cuda_ms"] / times["ours_ms"], 2), "angular_aev_rel_diff": err_a, "position_grad_rel_diff": err}
        if err_a > 1e-4:
            run["note"] = ("on this input the reference's angular CUDA kernel (sm_100a build of the unmodified source) deviates from the reference's "
                           "own CPU implementation; ours agrees with the CPU reference to < 1e-5 (tests/test_reference_cuda_gpu.py) -- the timing "
                           "comparison stands, the result comparison does not")
        res.append(run)
    out = {"what": "AEV forward + backward, reference CudaANISymmetryFunctions (unmodified source, nvcc -gencode arch=compute_100a,code=sm_100a) "
                   "vs this library, same GPU, same inputs, device-resident, CUDA events (the reference launches on the legacy default "
                   "stream, ours on torch's current stream)", "runs": res}
    try:
        out["cfconv"] = cfconv_comparator(dev)
    except Exception as exc:   # noqa: BLE001
        out["cfconv"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    return out


def cfconv_comparator(dev):
    """SchNet CFConv on the reference's own CUDA classes (CudaCFConvNeighbors + CudaCFConv, unmodified, sm_100a) next to ours: neighbour
    build + one layer forward + backprop on a 10 000-atom periodic box (width 128, 50 Gaussians, cutoff 5 A; the reference's N x N pair
    table in managed memory stops at about 16 000 atoms -- BASELINE config 4's 100 000 atoms are out of its reach)."""
    import numpy as np
    import torch
    import oracle_lib as O
    from systems import cubic_box, lattice
    from nnpops_b200.CFConv import CFConv
    from nnpops_b200.CFConvNeighbors import CFConvNeighbors
    l = O.ref_cuda_lib()
    if l is None or not hasattr(l, "refcuda_cfconv_create"):
        return {"unavailable": "oracle/_ref/libnnpops_ref_cuda.so has no CFConv classes (needs /root/reference at build time)"}
    n, W, Gn, cutoff, sigma = 10000, 128, 50, 5.0, 0.2
    rng = np.random.default_rng(17)
    pos, edge = lattice(n, 2.154, 0.3, 4004)
    w1 = rng.normal(0, 0.1, (W, Gn)).astype(np.float32); b1 = rng.normal(0, 0.1, W).astype(np.float32)
    w2 = rng.normal(0, 0.1, (W, W)).astype(np.float32); b2 = rng.normal(0, 0.1, W).astype(np.float32)
    x = torch.tensor(rng.standard_normal((n, W)).astype(np.float32), device=dev)
    go = torch.tensor(rng.standard_normal((n, W)).astype(np.float32), device=dev)
    p = torch.tensor(pos, device=dev); b = torch.tensor(cubic_box(edge), device=dev)
    ref = O.RefCudaCFConv(n, W, Gn, cutoff, True, sigma, "ssp", w1, b1, w2, b2)
    y0, ig0, pg0 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(p)

    def run_ref():
        ref.build(p, b)
        ref.compute(p, b, x, y0)
        ref.backprop(p, b, x, go, ig0, pg0)

    nb = CFConvNeighbors(cutoff)
    conv = CFConv(sigma, "ssp", torch.tensor(w1.reshape(Gn, W)), torch.tensor(b1), torch.tensor(w2), torch.tensor(b2))
    pr = p.clone().requires_grad_(True); xr = x.clone().requires_grad_(True)
    keep = {}

    def run_ours():
        pr.grad = None; xr.grad = None
        nb.build(pr, b)
        keep["y"] = conv(nb, pr, xr)
        keep["y"].backward(go)

    times = {}
    for key, fn, reps in (("reference_cuda_ms", run_ref, 3), ("ours_ms", run_ours, 10)):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        times[key] = e0.elapsed_time(e1) / reps
    res = {"what": "SchNet CFConv: neighbour build + one layer forward + backprop, reference CudaCFConvNeighbors + CudaCFConv (unmodified source, "
                   "sm_100a) vs this library through its torch module (autograd included), same GPU, same inputs",
           "system": "10000-atom periodic box, width 128, 50 Gaussians, cutoff 5 A", "pairs": ref.num_pairs(), "pairs_ours": nb.num_pairs(),
           "reference_cuda_ms": round(times["reference_cuda_ms"], 4), "ours_ms": round(times["ours_ms"], 4),
           "speedup": round(times["reference_cuda_ms"] / times["ours_ms"], 2),
           "output_rel_diff": float((keep["y"].detach() - y0).abs().max() / y0.abs().max()),
           "input_grad_rel_diff": float((xr.grad - ig0).abs().max() / ig0.abs().max()),
           "position_grad_rel_diff": float((pr.grad - pg0).abs().max() / pg0.abs().max())}
    ref.close()
    return res


def forces_check(nets, mlp_impl, local, n_atoms=6000):
    """The second half of the BASELINE metric, `forces max|delta|`: energy and forces of the bench model (same networks, same
    density, Rcr 5.2) on a %d-atom periodic water box against the oracle chain -- fp64 C restatement of the reference AEV ->
    fp64 ATen networks -> fp64 AEV backward (the O(N^2) oracle is too slow for the 50 000-atom box inside a bench run; the whole
    50 000-atom AEV + dE/dx is compared with the compiled reference in tests/test_baseline_sizes_gpu.py)."""
    import torch
    import oracle_lib as O
    from mlp_ref import mlp_energy_and_grad
    from systems import ANI2X, lattice, cubic_box, water_species, rel_err
    from nnpops_b200.OptimizedTorchANI import FusedANI
    pos, L = lattice(n_atoms, 2.154, 0.3, 7777)
    species = water_species(n_atoms)
    box = cubic_box(L)
    dev = torch.device("cuda", local)
    m = FusedANI(7, 5.2, ANI2X["Rca"], ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species, nets,
                 mlp_impl=mlp_impl, device="cuda:%d" % local)
    e, g = m.energy_and_gradient(torch.tensor(pos, device=dev), torch.tensor(box, device=dev))
    e = float(e.cpu()[0]); g = g.cpu().numpy().astype(np.float64)
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    r0, a0 = O.ani_forward(pos, species, 7, 5.2, 3.5, rfn, afn, box=box, bits=64)
    e0, dA = mlp_energy_and_grad(np.concatenate([r0, a0], axis=1), species, nets, torch.float64)
    g0 = O.ani_backward(pos, species, 7, 5.2, 3.5, rfn, afn, dA[:, :112], dA[:, 112:], box=box, bits=64)
    return {"system": "%d-atom periodic water box, same networks / density / cutoffs as the bench system" % n_atoms,
            "oracle": "fp64 restatement of the reference AEV (oracle/ani_oracle.c) -> fp64 ATen MLP -> fp64 AEV backward",
            "forces_max_abs_delta": float(np.abs(g - g0).max()), "forces_max_abs": float(np.abs(g0).max()), "forces_rel": rel_err(g, g0),
            "energy_rel": abs(e - e0) / abs(e0), "tolerance": 1e-5}


def measure_box(args, dist, dev, nets, rank, world, local):
    """ONE box evaluated cooperatively by all ranks: spatial domain decomposition into bricks with ghost-atom halos (SURVEY 8e
    variant i, nnpops_b200/halo.py) -- strong scaling.  Every rank keeps the positions of ITS atoms on its device; a step = forward
    halo (grouped NCCL send/recv of ghost positions) -> brick-local fused model -> reverse halo (ghost gradient rows back to the
    owners) -> all-reduce of the energy.  Device-event time, MAX over ranks."""
    import torch
    from systems import ANI2X, water_species
    from nnpops_b200.halo import HaloBoxANI
    n = args.atoms
    species = water_species(n)
    pool = 2
    models, pos_own, boxes = [], [], []
    for c in range(pool):                                   # the SAME conformers on every rank; one halo plan + local model per conformer
        pos, box = make_conformer(n, c)
        m = HaloBoxANI(7, 5.2, ANI2X["Rca"], ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species,
                       nets, pos, box, mlp_impl=args.mlp, device="cuda:%d" % local)
        pos_own.append(torch.tensor(pos[m.plan.owned[rank]], device=dev))
        boxes.append(torch.tensor(box, device=dev))
        if not args.no_graph:
            m.capture(boxes[-1])                             # whole step (halo exchange + local model + reverse halo + all-reduce) as one CUDA graph
        models.append(m)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        e, g = models[i % pool].energy_and_gradient(pos_own[i % pool], boxes[i % pool])
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        e, g = models[i % pool].energy_and_gradient(pos_own[i % pool], boxes[i % pool])
    t1.record()
    sync_all()
    ms_total = max_over_ranks(t0.elapsed_time(t1), dist, dev)
    # what the exchange costs: the brick-local model alone (same local positions, no halo phases, no all-reduce), kernel by kernel
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    l0.record()
    for i in range(args.steps):
        models[i % pool].local.energy_and_gradient(models[i % pool].local_pos, boxes[i % pool])
    l1.record()
    sync_all()
    ms_local = max_over_ranks(l0.elapsed_time(l1), dist, dev)
    counts = torch.tensor([models[0].n_owned, models[0].n_ghost, models[0].halo_bytes_forward + models[0].halo_bytes_reverse], dtype=torch.float64, device=dev)
    mx = counts.clone()
    if dist is not None:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return {"value": round(args.steps / (ms_total / 1e3), 4), "unit": UNIT, "ms_per_step": round(ms_total / args.steps, 4), "scaling": "strong",
            "decomposition": "bricks %s with ghost halos of Rcr = 5.2 A, one brick per GPU" % (models[0].plan.grid,),
            "atoms_owned_max": int(mx[0]), "ghost_atoms_max": int(mx[1]), "halo_bytes_per_step_per_gpu_max": int(mx[2]),
            "collectives": "2 all-to-all phases of ncclSend/ncclRecv (ghost positions out, ghost gradient rows back) + 1 scalar all-reduce per step",
            "cuda_graph": not args.no_graph,
            "local_model_ms_per_step": round(ms_local / args.steps, 4),
            "local_model_note": "the brick-local model alone, launched kernel by kernel without the halo phases and the all-reduce (MAX over ranks): "
                                "the difference to ms_per_step is what the exchange and the collectives cost -- or, when negative, what the CUDA graph saves",
            "energy": float(e.cpu()[0])}


def measure_pme(args, dist, dev, rank, world):
    """BASELINE config 5 as an extra key of the bench line: PME energy + dE/dx + dE/dq of ONE periodic box of 200 000 point charges
    (cubic, 128^3 grid, order 5, alpha 2.92 / nm, direct cutoff 0.9 nm) through nnpops_b200.pme.PME, strong scaling over the ranks.
    Direct space: the fused cell-list kernel, rank r owning slab r of the cell-sorted atoms as centres (no halo, no atomics), the
    per-atom derivatives assembled by ONE all-reduce of [atoms, 4] floats + one scalar all-reduce of the energy (NCCL).  Reciprocal
    space: replicated on every rank (the 128^3 grid is too small to pay for a grid exchange: DESIGN.md section 7).  Device-event
    time, MAX over ranks.  At N = 1 the reference's two-step direct path (getNeighborPairs + pme_direct) is timed beside it."""
    import numpy as np
    import torch
    from systems import lattice, cubic_box
    from nnpops_b200.neighbors import getNeighborPairs
    from nnpops_b200.pme import PME
    from nnpops_b200.pme.pme import pme_direct
    n, cutoff, alpha, coulomb = 200000, 0.9, 2.92, 138.935
    pos_np, L = lattice(n, 0.2154, 0.3, 5005)
    rng = np.random.default_rng(5005)
    q_np = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    q_np -= q_np.mean(dtype=np.float64).astype(np.float32)
    pos = torch.tensor(pos_np, device=dev, requires_grad=True)
    q = torch.tensor(q_np, device=dev, requires_grad=True)
    box = torch.tensor(cubic_box(L), device=dev)
    pme = PME(128, 128, 128, 5, alpha, coulomb, torch.zeros((n, 0), dtype=torch.int32))
    packed = torch.zeros((n, 4), dtype=torch.float32, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def direct():
        pos.grad = None; q.grad = None
        e = pme.compute_direct_sharded(pos, q, cutoff, box) if world > 1 else pme.compute_direct(pos, q, cutoff, box)
        e.backward()
        if world > 1:   # every rank holds the derivatives of its own centres: assemble them
            packed[:, :3] = pos.grad; packed[:, 3] = q.grad
            dist.all_reduce(packed)
        return e

    def recip():
        pos.grad = None; q.grad = None
        e = pme.compute_reciprocal(pos, q, box)
        e.backward()
        return e

    def timed(fn, iters):
        for _ in range(3):
            fn()
        sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            out = fn()
        t1.record()
        sync_all()
        return max_over_ranks(t0.elapsed_time(t1), dist, dev) / iters, out

    pd, qd = pos.detach(), q.detach()

    def one_call():   # PME.energy_and_derivatives: no autograd; at N > 1 direct by slabs, reciprocal by atom blocks + grid all-reduce
        return pme.energy_and_derivatives(pd, qd, cutoff, box, world=world)[0]

    iters = 10
    d_ms, e_d = timed(direct, iters)
    r_ms, e_r = timed(recip, iters)
    a_ms, _ = timed(lambda: (direct(), recip()), iters)
    f_ms, e_f = timed(one_call, iters)
    t_ms = min(a_ms, f_ms)
    out = {"workload": "BASELINE configs[4]: PME, one periodic box of 200000 charges, 128^3 grid, order 5, direct cutoff 0.9 nm; energy + dE/dx + dE/dq",
           "value": round(1e3 / t_ms, 2), "unit": "evals/s", "ms_per_step": round(t_ms, 4), "scaling": "strong",
           "api": "PME.energy_and_derivatives (one call)" if f_ms <= a_ms else "PME.compute_direct(_sharded) + PME.compute_reciprocal through autograd",
           "one_call_ms": round(f_ms, 4), "autograd_ms": round(a_ms, 4), "autograd_direct_ms": round(d_ms, 4), "autograd_reciprocal_ms": round(r_ms, 4),
           "direct": "fused cell-list kernel (no pair list, centre-owned, no atomics)" + (
               "; rank r owns slab r of the cell-sorted atoms" if world > 1 else ""),
           "reciprocal": ("one call: rank r spreads its block of atoms, ncclAllReduce of the 128^3 charge grid (8.4 MB), every rank solves the grid and "
                          "interpolates its own atoms; autograd path: replicated on every rank") if world > 1 else "spread, rFFT (cuFFT), convolution, irFFT, gather",
           "collectives": ("one call: grid all-reduce + ONE ncclAllReduce of the packed [200000, 4] derivatives (direct + reciprocal, 3.2 MB) + 1 scalar; "
                           "autograd path: all-reduce of the direct derivatives + 1 scalar") if world > 1 else "none",
           "energy_direct": float(e_d.detach().cpu()), "energy_reciprocal": float(e_r.detach().cpu()), "energy_one_call": float(e_f.detach().cpu())}
    if world == 1:
        cap = 33_000_000
        excl = torch.zeros((n, 0), dtype=torch.int32, device=dev)
        keep = {}

        def nb():
            keep["r"] = getNeighborPairs(pd, cutoff, cap, box)
        nb_ms, _ = timed(nb, 5)
        nbrs, deltas, dists, found = keep["r"]
        ld_ms, e_l = timed(lambda: pme_direct(pd, qd, nbrs, deltas, dists, excl, alpha, coulomb), 5)
        out["reference_two_step_direct"] = {"getNeighborPairs_ms": round(nb_ms, 4), "pme_direct_ms": round(ld_ms, 4), "pairs": int(found.item()),
                                            "energy_direct": float(e_l.cpu())}
    return out


def run_box(args):
    """--mode box: only the strong-scaling measurement of ONE box (see measure_box), as its own JSON line."""
    import torch
    from systems import ANI2X_HIDDEN, ANI2X_ENSEMBLE
    from mlp_ref import random_networks
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nnpops_nccl.%h.%p.log")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    box = measure_box(args, dist, dev, nets, rank, world, local)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        args.quiet.restore()
        print(json.dumps({
            "metric": METRIC, "value": box["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": box["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "ONE %d-atom periodic water box per step over %d GPU(s): %s" % (args.atoms, world, box["decomposition"]),
                       "atoms": args.atoms, "mode": "box", "mlp_impl": args.mlp},
            "box": box, "clocks": clocks}), flush=True)
    if dist is not None:
        os.dup2(2, 1)
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm.  AEV forward+backward: the reference's CpuANISymmetryFunctions through oracle/_ref (kind
# "reference") or, when that was never built, the C restatement (kind "port").  Network: the reference's BatchedLinear replicates
# the weights per atom (BatchedNN.py:71-83; 541 GB at 50 000 atoms), so the per-species ATen nn.Linear / CELU chain of
# BASELINE.md section 3 stands in for it (same arithmetic, species-grouped), forward + input-gradient backward, one thread.
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_aev(n_atoms, seed):
    """One AEV forward+backward of an n-atom periodic box on one core; returns seconds."""
    import oracle_lib as O
    from systems import ANI2X, lattice, cubic_box, water_species
    pos, L = lattice(n_atoms, 2.154, 0.3, seed)
    species = water_species(n_atoms)
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    rg = np.ones((n_atoms, 112), np.float32); ag = np.ones((n_atoms, 896), np.float32)
    impl = "ref" if O.ref_lib() is not None else "oracle"
    t = time.perf_counter()
    O.ani_backward(pos, species, 7, 5.2, 3.5, rfn, afn, rg, ag, box=cubic_box(L), impl=impl)   # ref_ani runs forward + backward
    if impl == "oracle":
        O.ani_forward(pos, species, 7, 5.2, 3.5, rfn, afn, box=cubic_box(L))
    return time.perf_counter() - t


_NETS = {}


def _cpu_mlp(n_atoms):
    """The ensemble of per-species networks on n_atoms water atoms (2/3 H, 1/3 O), forward + backward to the AEV, ATen fp32 on
    ONE thread; returns seconds.  Random ANI-2x-shaped weights (same generator as the GPU arm)."""
    import torch
    from systems import ANI2X_HIDDEN, ANI2X_ENSEMBLE
    from mlp_ref import random_networks
    torch.set_num_threads(1)
    if "nets" not in _NETS:
        nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
        _NETS["nets"] = {s: [[(torch.tensor(W), torch.tensor(bb)) for W, bb in member] for member in nets[s]] for s in (0, 3)}
    x = {0: torch.randn(n_atoms - n_atoms // 3, 1008), 3: torch.randn(n_atoms // 3, 1008)}
    t = time.perf_counter()
    for s, xs in x.items():
        xs = xs.requires_grad_(True)
        tot = 0.0
        for member in _NETS["nets"][s]:
            h = xs
            for l, (W, bb) in enumerate(member):
                h = torch.nn.functional.linear(h, W, bb)
                if l < len(member) - 1:
                    h = torch.nn.functional.celu(h, 0.1)
            tot = tot + h.sum()
        (tot / len(_NETS["nets"][s])).backward()
    return time.perf_counter() - t


def _cpu_eval(n_atoms, seed):
    """(seconds AEV, seconds network) of one sample."""
    return _cpu_aev(n_atoms, seed), _cpu_mlp(n_atoms)


def _scale_model(n_full):
    """Fit t_aev(N) = a N^2 + b N from two small sizes (the reference scans all pairs) and return (a, b, kind)."""
    import oracle_lib as O
    n1, n2 = 4000, 12000
    t1, t2 = _cpu_aev(n1, 9001), _cpu_aev(n2, 9002)
    a = (t2 / n2 - t1 / n1) / (n2 - n1)
    b = t1 / n1 - a * n1
    if a <= 0 or b <= 0:   # degenerate fit: fall back to pure N^2
        a, b = t2 / (n2 * n2), 0.0
    return a, b, ("reference" if O.ref_lib() is not None else "port")


def cpu_baseline(n_full, budget_s):
    a, b, kind = _scale_model(n_full)
    n_s = 6000
    t_aev, t_mlp = _cpu_eval(n_s, 9003)
    t_full = t_aev * (a * n_full ** 2 + b * n_full) / (a * n_s ** 2 + b * n_s) + t_mlp * n_full / n_s
    return {"value": round(1.0 / t_full, 6), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "one %d-atom periodic box of the same density on one core: reference CpuANISymmetryFunctions AEV forward+backward %.2f s "
                      "(scaled to %d atoms with the measured a*N^2+b*N model, a=%.3e, b=%.3e) + per-species ATen nn.Linear/CELU stand-in for the "
                      "network, forward+backward %.2f s (scaled linearly; the reference's BatchedLinear needs 541 GB of replicated weights at "
                      "this size) -> %.1f s per evaluation" % (n_s, t_aev, n_full, a, b, t_mlp, t_full)}


def _worker(q_in, q_out):
    while True:
        job = q_in.get()
        if job is None:
            return
        q_out.put(_cpu_eval(*job))


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle_lib as O
    O.build()
    cores = os.cpu_count() or 1
    n_s = 6000
    a, b, kind = _scale_model(args.atoms)
    scale_aev = (a * args.atoms ** 2 + b * args.atoms) / (a * n_s ** 2 + b * n_s)
    scale_mlp = args.atoms / n_s
    ctx = mp.get_context("fork")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(q_in, q_out), daemon=True) for _ in range(cores)]
    for p in procs:
        p.start()

    def step(i):
        for c in range(cores):
            q_in.put((n_s, 9100 + i * cores + c))
        return [q_out.get() for _ in range(cores)]

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    times = []
    for i in range(args.steps):
        times += step(args.warmup + i)
    dt = time.perf_counter() - t0
    for _ in procs:
        q_in.put(None)
    # every core worked through `steps` samples; a sample stands for (its AEV seconds x scale_aev + its network seconds x scale_mlp)
    # of a full evaluation.  Full-evaluation seconds per core-sample, averaged; `cores` evaluations proceed in parallel.
    full_s = float(np.mean([ta * scale_aev + tm * scale_mlp for ta, tm in times]))
    wall_factor = dt / max(float(np.sum([ta + tm for ta, tm in times])) / cores, 1e-9)   # queueing / memory-bandwidth overhead of the parallel run
    value = cores / (full_s * wall_factor)
    # one REAL 50 000-atom AEV evaluation on one core: the error of the a*N^2 + b*N model is stated, not assumed
    model_check = None
    if args.atoms >= 20000 and not args.no_model_check:
        t_real = _cpu_aev(args.atoms, 3000)
        t_model = a * args.atoms ** 2 + b * args.atoms
        model_check = {"real_seconds": round(t_real, 2), "model_seconds": round(t_model, 2), "model_over_real": round(t_model / t_real, 3),
                       "note": "one real %d-atom AEV forward+backward on one otherwise idle core vs the fitted model" % args.atoms}
    sample = ("per step and core: one %d-atom periodic box -- reference CpuANISymmetryFunctions (oracle/_ref, unmodified reference C++) AEV "
              "forward+backward, scaled to %d atoms by the measured a*N^2+b*N cost model (factor %.1f), plus the per-species ATen nn.Linear/CELU "
              "stand-in for the network (BASELINE.md section 3; the reference's own BatchedLinear needs 541 GB of replicated weights here), "
              "forward+backward, scaled linearly (factor %.2f); %d cores, one sample each, in parallel; mean AEV %.2f s + network %.2f s per sample"
              % (n_s, args.atoms, scale_aev, scale_mlp, cores, float(np.mean([t[0] for t in times])), float(np.mean([t[1] for t in times]))))
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BASELINE configs[2]: ANI-2x, %d-atom periodic water box, Rcr 5.2 A, Rca 3.5 A" % args.atoms, "atoms": args.atoms},
           "cpu_baseline": {"value": round(value, 6), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "model_check": model_check,
           "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mlp", default=os.environ.get("NNPOPS_MLP", DEFAULT_MLP), choices=["tcgen05", "simt"])
    ap.add_argument("--atoms", type=int, default=50000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain", type=float, default=3.0, help="seconds of the extra sustained leg (0 = skip)")
    ap.add_argument("--md-steps", type=int, default=64, help="steps of the extra MD-like leg with a Verlet skin (0 = skip; N = 1 only)")
    ap.add_argument("--no-graph", action="store_true", help="box mode: launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-pme", action="store_true", help="skip the extra PME leg (BASELINE config 5, 200 000 charges)")
    ap.add_argument("--no-box", action="store_true", help="N > 1: skip the extra one-box strong-scaling measurement")
    ap.add_argument("--no-model-check", action="store_true", help="--impl reference: skip the one real full-size AEV evaluation (~80 s)")
    ap.add_argument("--mode", default="conformers", choices=["conformers", "box"],
                    help="conformers: independent conformers per GPU, no collective (BASELINE config 3, the default; with N > 1 the line "
                         "also carries a \"box\" key); box: only ONE box over the GPUs by bricks with ghost halos (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    with StdoutToStderr() as quiet:
        args.quiet = quiet
        if args.mode == "box":
            run_box(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
