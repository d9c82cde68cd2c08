#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ANIm path (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (see below)

Workload ("step" = one pass of the hot path over the whole pair list):
  N = 1 : ANIm on 50 synthetic ~5 Mb bacterial genomes (5 clades x 10 strains, seed 20260925),
          all 2,450 directed comparisons.
  N > 1 : weak scaling -- N islands of 50 genomes (island k seeded 20260925+k); the global pair list
          (N x 2,450 directed pairs, no cross-island pairs) is block-partitioned over the ranks,
          rank r computes its block, and ONE all-gather of the 48-byte result structs assembles the
          full result array on every rank (inside the timed region).
Metric: directed genome-pair comparisons per second (BASELINE.json `metric`).
  value : genomes already packed in HBM when the timed region starts; each step rebuilds the seed
          indexes and runs every comparison.  Timed on the device (CUDA events on the engine's stream,
          max over ranks).
  e2e   : the same job through the public API with HOST buffers: per step clear the context, copy
          every genome from pinned host memory (H2D), pack, index, compare, and read the results
          back (D2H).
CPU arm (`--impl reference`, also the `cpu_baseline` object): pyani's own CPU path for ANIm is the
external nucmer/delta-filter binaries, which are not available here, so the arm times the
in-repo CPU oracle (oracle/, "port") on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260925  # 20260924 + config index 1
N_CLADES, PER_CLADE, GENOME_LEN = 5, 10, 5_000_000
DPX_OPS_PER_CELL = 4  # E, F, H max-plus updates + 1 add per affine cell
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of this workload
# (profiles/r01_v30_prof_v30.txt: one batch per step, first captured launch of each kernel; seed_extend from the v22 capture
# scaled to one launch per step); None where no capture exists
NCU_TRAFFIC_BYTES_PER_LAUNCH = {"gap_dp_kernel": 21.246920e9 + 1.759511e9, "ext_dp_kernel": 216.553e6 + 35.053e6,
                                "seed_probe_kernel": 23.121308e9 + 432.529e6, "seed_extend_kernel": 2 * (7.349801e9 + 489.665536e6)}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_island(k: int, n_clades=N_CLADES, per_clade=PER_CLADE, length=GENOME_LEN):
    from pyani_b200 import synth

    return synth.make_set(n_clades, per_clade, length, seed=SEED + k)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        # only samples under load say anything about a clock lock
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample_pairs(n_clades=N_CLADES, per_clade=PER_CLADE):
    """Bounded sample of the workload: 4 strains of clade 0 + 2 of clade 1 + 2 of clade 2, all 56 directed
    pairs among them (within- and between-clade mix close to the full set's 18 % / 82 %)."""
    idx = [0, 1, 2, 3, per_clade, per_clade + 1, 2 * per_clade, 2 * per_clade + 1]
    pairs = [(a, b) for a in idx for b in idx if a != b]
    return idx, pairs


def run_cpu_sample(genomes, cores: int):
    """Oracle (CPU port) over the bounded sample on `cores` threads -> (comparisons/s, n_pairs, seconds)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor

    idx, pairs = cpu_sample_pairs()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        og = dict(zip(idx, ex.map(lambda i: oracle.OGenome(genomes[i].residues, genomes[i].contig_off), idx)))
        list(ex.map(lambda p: oracle.anim_pair(og[p[0]], og[p[1]])[0], pairs))
    dt = time.perf_counter() - t0
    return len(pairs) / dt, len(pairs), dt


def reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    genomes = make_island(0)
    for _ in range(args.warmup):
        pass  # the CPU arm has no caches worth warming; keep the sample bounded
    vals, secs = [], []
    for _ in range(max(1, args.steps)):
        v, n_pairs, dt = run_cpu_sample(genomes, cores)
        vals.append(v); secs.append(dt)
    value = float(np.mean(vals))
    sample = f"{n_pairs} directed pairs among 8 of the 50 genomes (4+2+2 strains of 3 clades), oracle index build included, per step"
    line = {
        "impl": "reference", "metric": "genome-pair comparisons/sec", "value": value, "unit": "directed comparisons/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "directed comparisons/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "directed comparisons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "pyani's ANIm CPU path is nucmer+delta-filter (absent here, SURVEY.md 8c); this arm times the in-repo CPU oracle port",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int):
    return {"workload": f"ANIm, {n_gpus} island(s) x 50 synthetic ~5 Mb bacterial genomes (5 clades x 10 strains), "
                        f"{n_gpus * 2450} directed comparisons (BASELINE.json configs[1])",
            "genomes_per_island": N_CLADES * PER_CLADE, "genome_len": GENOME_LEN, "seed": SEED,
            "parallelism": f"pairs block-partitioned over {n_gpus} GPU(s), one all-gather of result structs" if n_gpus > 1 else "single GPU",
            "l2": "inputs larger than L2 (packed genomes 95 MB + seed indexes 2.2 GB per island)"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny workload for smoke-testing the harness")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ncu", action="store_true", help="profiling run: honour --warmup as given, skip the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" and not args.small and not args.ncu else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch

    import pyani_b200
    from pyani_b200 import dist as pdist
    from pyani_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: island `rank` (the block partition of the global pair list is exactly one island per rank)
    global N_CLADES, PER_CLADE, GENOME_LEN
    if args.small:
        N_CLADES, PER_CLADE, GENOME_LEN = 2, 3, 300_000
    t0 = time.perf_counter()
    genomes = make_island(rank, N_CLADES, PER_CLADE, GENOME_LEN)
    n_g = len(genomes)
    local_pairs = synth.all_directed_pairs(n_g)
    n_local = len(local_pairs)
    lengths = [len(g.residues) for g in genomes]
    # global pair list = islands back to back; equal-cost islands -> rank r's block is island r
    blocks = [(r * n_local, (r + 1) * n_local) for r in range(world)]
    log(f"[rank {rank}] generated {n_g} genomes ({sum(lengths)/1e6:.1f} Mb) in {time.perf_counter()-t0:.1f}s")
    # pinned host copies for the e2e leg
    pinned = []
    for g in genomes:
        t = torch.empty(len(g.residues), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = g.residues
        pinned.append(t)
    h2d_bytes = int(sum(lengths) + sum(8 * len(g.contig_off) for g in genomes) + local_pairs.nbytes)
    d2h_bytes = int(n_local * 48)

    eng = pyani_b200.Engine(local_rank)

    def stage_all():
        eng.clear()
        return [eng.add_genome(p.numpy(), g.contig_off, g.name) for p, g in zip(pinned, genomes)]

    def gather(local):
        if world == 1:
            return local
        return pdist.all_gather_results(local, blocks, device=dev)

    # ---- resident leg: genomes packed in HBM, each step = index build + all comparisons (+ all-gather)
    stage_all()
    stage_tot, counters_last, step_ms = {}, {}, []
    gather_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for it in range(args.warmup):
        eng.reset_indexes()
        gather(eng.anim(local_pairs))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    results = None
    for it in range(args.steps):
        eng.reset_indexes()
        res = eng.anim(local_pairs)
        ms, cnt = eng.last_profile()
        gather_ev[0].record()
        results = gather(res)
        gather_ev[1].record()
        torch.cuda.synchronize()
        step_ms.append(ms["total"] + (gather_ev[0].elapsed_time(gather_ev[1]) if world > 1 else 0.0))
        for k, v in ms.items():
            stage_tot[k] = stage_tot.get(k, 0.0) + v
        counters_last = cnt
    barrier()
    clocks = sampler.stop()
    mean_ms = float(np.mean(step_ms))
    if world > 1:
        t = torch.tensor([mean_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mean_ms = float(t.item())
    value = world * n_local / (mean_ms * 1e-3)

    # ---- e2e leg: host buffers in, results out, every step
    e2e_s = []
    for it in range(0 if args.ncu else max(1, min(args.steps, 3)) + 1):
        barrier()
        t1 = time.perf_counter()
        stage_all()
        gather(eng.anim(local_pairs))
        torch.cuda.synchronize()
        if it > 0:  # first pass re-warms allocator state after the resident leg
            e2e_s.append(time.perf_counter() - t1)
    e2e_ms = float(np.mean(e2e_s)) * 1e3 if e2e_s else float("nan")
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * n_local / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0, per-launch averages over the timed steps)
    K = args.steps
    hbm_peak, peak_src = peaks()
    dpx_peak = eng.dpx_peak_gops()
    cells_gap, cells_ext = counters_last["gap_cells"], counters_last["ext_cells"]
    probes, hits = counters_last["probes"], counters_last["seed_hits"]
    kern = {
        "seed_probe_kernel": {"ms": stage_tot["seed"] / K, "launches": counters_last["seed_launches"]},
        "seed_extend_kernel": {"ms": stage_tot.get("seed_extend", 0.0) / K, "launches": counters_last["seed_launches"]},
        "gap_dp_kernel": {"ms": stage_tot["gap_dp"] / K, "launches": counters_last["gap_dp_launches"]},
        "ext_dp_kernel": {"ms": stage_tot["ext_dp_kernel"] / K, "launches": counters_last["ext_dp_launches"]},
    }
    n_anchors = counters_last["anchors"]

    def roofline_of(name):
        ms_k, launches = kern[name]["ms"], max(1, kern[name]["launches"])
        if name.startswith("seed"):
            # algorithmic bytes: probe = 4 B k-mer streamed + 4 B presence-filter word + (for the ~10 % that
            # pass it) one 8 B table slot, 8 B written per hit; extend = 8 B hit + 4 B posflag + 2 x 16 B packed sequence
            # per hit, 16 B per anchor written
            alg = probes * 8 + hits * 16 if name == "seed_probe_kernel" else hits * 44 + n_anchors * 16
            ach = alg / (ms_k * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(name) if not args.small else None,
                    "peak_source": peak_src, "per_launch_ms": ms_k / launches, "algorithmic_bytes_per_launch": alg / launches}
        cells = cells_gap if name == "gap_dp_kernel" else cells_ext
        ach = cells * DPX_OPS_PER_CELL / (ms_k * 1e-3) / 1e9
        return {"kernel": name, "bound": "alu", "achieved": ach, "peak": dpx_peak, "unit": "Gop/s (DPX-class int32 max-plus ops)",
                "frac": ach / dpx_peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(name) if not args.small else None,
                "peak_source": "measured live: b200ani_dpx_peak (__viaddmax_s32 issue rate)",
                "per_launch_ms": ms_k / launches, "cells_per_launch": cells / launches, "ops_per_cell": DPX_OPS_PER_CELL,
                "gcups": cells / (ms_k * 1e-3) / 1e9}

    dom = max(kern, key=lambda k: kern[k]["ms"])
    roof = roofline_of(dom)
    roof["all_kernels"] = {k: {kk: vv for kk, vv in roofline_of(k).items() if kk in ("bound", "achieved", "peak", "frac", "per_launch_ms", "gcups", "traffic")} for k in kern}
    roof["hbm_peak_gbs"] = hbm_peak
    roof["dpx_peak_gops"] = dpx_peak

    line = {
        "metric": "genome-pair comparisons/sec", "value": value, "unit": "directed comparisons/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "directed comparisons/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms},
        "gpu_launches": int(counters_last["launches"] * K),
        "clocks": clocks,
        "roofline": roof,
        "dp_cell_updates_per_s": (cells_gap + cells_ext) / (mean_ms * 1e-3),
        "unordered_pairs_per_s": value / 2,
        "stage_ms_per_step": {k: v / K for k, v in stage_tot.items()},
        "kernel_ms_per_step": {k: v["ms"] for k, v in kern.items()},
        "counters_per_step": counters_last,
        "results_check": {"aligned_pairs": int((results["status"] == 0).sum()), "pairs": int(len(results)),
                          "mean_identity": float(np.mean(results["sum_weighted"][results["status"] == 0] / results["sum_aligned"][results["status"] == 0]))},
    }
    if world == 1 and not args.no_cpu and not args.small:
        cores = os.cpu_count() or 1
        v, n_pairs, dt = run_cpu_sample(genomes, cores)
        line["cpu_baseline"] = {"value": v, "unit": "directed comparisons/s", "cores": cores, "kind": "port",
                                "sample": f"{n_pairs} directed pairs among 8 of the 50 genomes (4+2+2 strains of 3 clades), oracle index build included; {dt:.1f} s wall"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
