#!/usr/bin/env python
"""bench.py -- benchmarks of the B200-native pairwise-comparison path (BASELINE.json configs[1..4]).

  python bench.py --gpus N --steps K --warmup W [--workload anim|anim50|anib|tetra|mixed]
  python bench.py --impl reference --gpus N --steps K --warmup W [--workload ...]     # the CPU arm

Workloads (one synthetic, phylogenetically structured genome set per workload, THE SAME set for every N; "step" = one
pass of the hot path over the whole set):
  anim   (default) BASELINE configs[3] -- ANIm on 500 x ~5 Mb genomes (50 clades x 10 strains), all 249,500 directed
         comparisons.  N = 1: everything on one GPU.  N > 1: STRONG scaling -- every rank stages all genomes, the
         pair list is dealt over the ranks (pyani_b200.dist.deal) and ONE all-gather of the 48-byte result structs
         assembles the full result on every rank, inside the timed region.
  anim50 BASELINE configs[1] -- ANIm on 50 x ~5 Mb genomes (5 clades x 10), 2,450 directed comparisons.
  anib   BASELINE configs[2] -- ANIb (1020-nt fragments) on 200 x ~5 Mb genomes (20 clades x 10), 39,800 comparisons.
  tetra  TETRA half of BASELINE configs[4] -- counts, Z-scores and the N x N correlation matrix of 1000 mixed-length
         genomes (0.5-12 Mb, log-uniform per clade); also reports the pack kernel (staging leg).
  mixed  BASELINE configs[4] -- ANIm (999,000 directed comparisons) + TETRA on the same 1000 genomes.
Metric: directed genome-pair comparisons per second (BASELINE.json `metric`); TETRA: unordered correlations per second.
  value : genomes already packed in HBM when the timed region starts; each step rebuilds the seed indexes and runs every
          comparison (+ the all-gather).  Timed on the device (CUDA events on the engine's stream), max over ranks.
  e2e   : the same job through the public API with HOST buffers: per step clear the context, copy every genome from pinned
          host memory (H2D), pack, index, compare, and read the results back (D2H).
CPU arm (`--impl reference`, and the `cpu_baseline` object of the GPU arm at N = 1): pyani's own CPU path for ANIm / ANIb
is the external nucmer / BLAST+ binaries, which are not available here, so the arm times the in-repo CPU
oracle ("port") with one pair per host thread over a bounded sample that keeps every core busy (>= 2 x cores pairs, same
within-/between-clade mix as the full pair list).  TETRA: the reference's own tetra.py (pip-installed copy under
baseline/_ref, "reference") on a few genomes, single thread like the reference.
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

BASE_SEED = 20260924  # + config index
WORKLOADS = {
    #          kind     clades per  length                 config
    "anim":   ("anim",  50,   10,  5_000_000,             3),
    "anim50": ("anim",  5,    10,  5_000_000,             1),
    "anib":   ("anib",  20,   10,  5_000_000,             2),
    "tetra":  ("tetra", 100,  10,  (500_000, 12_000_000), 4),
    "mixed":  ("mixed", 100,  10,  (500_000, 12_000_000), 4),
}
SMALL = {"anim": (3, 3, 300_000), "anim50": (2, 3, 300_000), "anib": (2, 3, 300_000), "tetra": (4, 3, (100_000, 400_000)),
         "mixed": (3, 3, (200_000, 400_000))}
DPX_OPS_PER_CELL = 4  # E, F, H max-plus updates + 1 add per affine cell


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_params(args):
    kind, n_clades, per_clade, length, cfg = WORKLOADS[args.workload]
    if args.small:
        n_clades, per_clade, length = SMALL[args.workload]
    if args.genomes:  # explicit override (profiling runs): keep 10 strains per clade
        n_clades = max(1, args.genomes // per_clade)
    return kind, n_clades, per_clade, length, cfg, BASE_SEED + cfg


def make_genomes(args, rank, world, only_clades=None):
    from pyani_b200 import synth

    kind, n_clades, per_clade, length, cfg, seed = workload_params(args)
    return synth.make_set_cached(n_clades, per_clade, length, seed, rank=rank, world=world, only_clades=only_clades)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the committed `ncu --set full` captures, keyed by
    workload then kernel (profiles/ncu_traffic.json, written by profiles/summarise.py from the captures); None if absent."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    return json.load(open(p)) if p.exists() else {}


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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(n_clades: int, per_clade: int, cores: int, min_pairs: int = 0, max_pairs: int = 0):
    """Genomes and directed pairs of the bounded CPU sample: `g` genomes spread as evenly as possible over the first `nc`
    clades, ALL directed pairs among them -- at least `min_pairs` (default 2 x cores: every host thread stays busy) and, if it can
    be had, at most `max_pairs` (default 6 x cores) pairs with the within-clade share closest to the full list's
    (per_clade-1)/(n_clades*per_clade-1)."""
    min_pairs = min_pairs or 2 * cores
    max_pairs = max_pairs or 6 * cores
    full_within = (per_clade - 1) / max(1, n_clades * per_clade - 1)
    n_all = n_clades * per_clade
    best = None
    for nc in range(2, n_clades + 1):
        for g in range(nc, nc * per_clade + 1):
            pairs = g * (g - 1)
            if pairs < min_pairs and g < n_all:
                continue
            q, r = divmod(g, nc)  # r clades hold q+1 strains, nc-r hold q
            within = r * (q + 1) * q + (nc - r) * q * (q - 1)
            cand = (pairs > max_pairs and g < n_all, abs(within / pairs - full_within), pairs, nc, g)
            if best is None or cand < best:
                best = cand
            if pairs > max_pairs:
                break
    _, _, _, nc, g = best
    q, r = divmod(g, nc)
    idx = [c * per_clade + s for c in range(nc) for s in range(q + (1 if c < r else 0))]
    pairs = [(a, b) for a in idx for b in idx if a != b]
    within = sum(1 for a, b in pairs if a // per_clade == b // per_clade)
    return idx, pairs, within, full_within


def run_cpu_sample(args, kind: str, cores: int):
    """Oracle (CPU port) over the bounded sample, one comparison per thread -> dict(value, n_pairs, seconds, ...)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor

    _, n_clades, per_clade, _, _, _ = workload_params(args)
    # ANIb: one comparison of two 5 Mb genomes keeps a core busy for ~40 s (the word stage is scalar here), so the sample is one
    # pair per thread (a single round) instead of >= 2 per thread
    idx, pairs, within, full_within = cpu_sample(n_clades, per_clade, cores, *((int(0.8 * cores), int(1.2 * cores)) if kind == "anib" else ()))
    genomes = make_genomes(args, 0, 1, only_clades=sorted({i // per_clade for i in idx}))
    by_id = {i: genomes[i] for i in idx}
    def make(res, off):
        g = oracle.OGenome(res, off, anim_index=(kind == "anim"))
        if kind != "anim":
            g.anib_handle  # build the ANIb seed index here, not inside the first comparison that needs it
        return g

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        og = list(ex.map(lambda i: make(np.asarray(by_id[i].residues), by_id[i].contig_off), idx))
    t1 = time.perf_counter()
    # the comparisons themselves run in C on `cores` POSIX threads pulling pairs from a shared counter: no interpreter in the loop
    local = {g: k for k, g in enumerate(idx)}
    lp = np.array([[local[a], local[b]] for a, b in pairs], dtype=np.int32)
    out, used = (oracle.anim_pairs_mt if kind == "anim" else oracle.anib_pairs_mt)(og, lp, cores)
    assert int((out[:, 0] > 0).sum()) == len(pairs), "the CPU arm must align every sampled pair"
    dt = time.perf_counter() - t0
    return {"value": len(pairs) / dt, "n_pairs": len(pairs), "seconds": dt, "index_seconds": t1 - t0, "cores_busy": min(used, len(pairs)),
            "sample": f"{len(pairs)} directed pairs = all pairs among {len(idx)} of the genomes ({within} within-clade = "
                      f"{100 * within / len(pairs):.1f} %, full list {100 * full_within:.1f} %), one comparison at a time per host thread ({used} POSIX threads in C), "
                      f"oracle index build of the {len(idx)} genomes included ({t1 - t0:.1f} s of {dt:.1f} s)"}


def run_cpu_tetra(args):
    """The reference's own tetra.py (baseline/_ref, pip-installed from /root/reference; Biopython stubbed by oracle/refstubs)
    on the first genomes of the set, single thread like the reference -> dict; falls back to the oracle port."""
    import tempfile

    genomes = make_genomes(args, 0, 1, only_clades=[0])[:3]
    total = sum(len(g.residues) for g in genomes)
    ref_dir = ROOT / "baseline" / "_ref"
    kind = "port"
    t0 = time.perf_counter()
    if (ref_dir / "pyani" / "tetra.py").exists():
        sys.path[:0] = [str(ROOT / "oracle" / "refstubs"), str(ref_dir)]
        from pyani import tetra as ref_tetra  # the unmodified reference

        kind = "reference"
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for g in genomes:
                p = Path(td) / f"{g.name}.fna"
                with open(p, "wb") as fh:
                    for c in range(len(g.contig_off) - 1):
                        fh.write(b">r%d\n" % c + bytes(np.asarray(g.residues[int(g.contig_off[c]):int(g.contig_off[c + 1])])) + b"\n")
                paths.append(p)
            t0 = time.perf_counter()
            z = {p.stem: ref_tetra.calculate_tetra_zscore(p) for p in paths}
            ref_tetra.calculate_correlations(z)
    else:
        import oracle

        zs = [oracle.tetra_zscores(*oracle.tetra_counts(np.asarray(g.residues), g.contig_off))[0] for g in genomes]
        oracle.tetra_corr(np.array(zs))
    dt = time.perf_counter() - t0
    return {"kind": kind, "seconds": dt, "bases": total, "genomes": len(genomes),
            "sample": f"{len(genomes)} genomes ({total / 1e6:.1f} Mb): Z-scores + correlations, 1 thread, {dt:.1f} s"}


def reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    kind, n_clades, per_clade, length, cfg, seed = workload_params(args)
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    if kind in ("anim", "anib", "mixed"):
        k = "anib" if kind == "anib" else "anim"
        runs = []
        budget_s = 240.0  # the whole arm ends within a few minutes: stop repeating the sample when the budget is spent
        t_all = time.perf_counter()
        for it in range(args.warmup + steps):
            r = run_cpu_sample(args, k, cores)
            if it >= args.warmup or it == args.warmup + steps - 1:
                runs.append(r)
            if time.perf_counter() - t_all > budget_s and runs:
                break
        value = float(np.mean([r["value"] for r in runs]))
        ms = 1e3 * float(np.mean([r["seconds"] for r in runs]))
        cb = {"value": value, "unit": "directed comparisons/s", "cores": cores, "cores_busy": runs[-1]["cores_busy"], "kind": "port",
              "sample": runs[-1]["sample"] + f"; {len(runs)} timed pass(es)"}
        unit = "directed comparisons/s"
    else:
        r = run_cpu_tetra(args)
        n = n_clades * per_clade
        # extrapolated to the workload: counting scales with bases, correlations are negligible next to it in the reference
        value = (n * (n - 1) / 2) / (r["seconds"] / r["bases"] * expected_bases(args))
        ms = r["seconds"] * 1e3
        cb = {"value": value, "unit": "unordered correlations/s", "cores": 1, "cores_busy": 1, "kind": r["kind"],
              "sample": r["sample"] + "; value = workload correlations / (measured s per base x workload bases)"}
        unit = "unordered correlations/s"
    line = {
        "impl": "reference", "metric": metric_name(kind), "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype_of(kind), "data": "synthetic",
        "config": workload_config(args, args.gpus), "cpu_baseline": cb,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "pyani's ANIm/ANIb CPU path is nucmer / BLAST+ (absent here, SURVEY.md 8c); this arm times the in-repo CPU oracle "
                "port, which is NOT nucmer (minimizer hashing instead of a suffix tree: expect it to be several times faster than "
                "nucmer per pair)" if kind != "tetra" else "the reference's own tetra.py, single thread (it has no parallelism)",
    }
    print(json.dumps(line), flush=True)


def expected_bases(args):
    from pyani_b200 import synth

    kind, n_clades, per_clade, length, cfg, seed = workload_params(args)
    if isinstance(length, (tuple, list)):
        return float(sum(synth.clade_lengths(seed, n_clades, length, True)) * per_clade) * 1.05
    return float(n_clades * per_clade * length) * 1.05


def metric_name(kind):
    return "TETRA correlations/sec" if kind == "tetra" else "genome-pair comparisons/sec"


def dtype_of(kind):
    return "f64" if kind == "tetra" else "int32"


def workload_config(args, n_gpus: int):
    kind, n_clades, per_clade, length, cfg, seed = workload_params(args)
    n = n_clades * per_clade
    what = {"anim": f"ANIm, all {n * (n - 1)} directed comparisons", "anib": f"ANIb (1020-nt fragments), all {n * (n - 1)} directed comparisons",
            "tetra": f"TETRA, Z-scores + {n * (n - 1) // 2} correlations", "mixed": f"ANIm ({n * (n - 1)} directed comparisons) + TETRA"}[kind]
    ln = f"{length[0] / 1e6:g}-{length[1] / 1e6:g} Mb (log-uniform per clade)" if isinstance(length, (tuple, list)) else f"~{length / 1e6:g} Mb"
    return {"workload": f"{what} among {n} synthetic {ln} bacterial genomes ({n_clades} clades x {per_clade} strains, between-clade "
                        f"divergence 15-25 %, within-clade 0.1-5 %), BASELINE.json configs[{cfg}]" + (" [--small harness check]" if args.small else ""),
            "genomes": n, "seed": seed,
            "parallelism": (f"strong scaling: genomes replicated on {n_gpus} GPUs, seed indexes built once across the GPUs and shared by one all-gather, "
                            f"unordered genome pairs dealt round-robin in walk order, one all-gather of result structs"
                            if n_gpus > 1 else "single GPU"),
            "l2": "inputs larger than L2 (packed genomes + seed indexes: ~50 MB per genome)"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="anim", choices=sorted(WORKLOADS))
    ap.add_argument("--genomes", type=int, default=0, help="override the number of genomes (profiling runs only; the line says so)")
    ap.add_argument("--small", action="store_true", help="tiny workload for smoke-testing the harness")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ncu", action="store_true", help="profiling run: honour --warmup as given, skip the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" and not args.small and not args.ncu else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    kind, n_clades, per_clade, length, cfg, seed = workload_params(args)

    # ---- the genome set: generated once per machine (ranks share the work), before CUDA is initialised (the pool forks)
    t0 = time.perf_counter()
    genomes = make_genomes(args, rank, world)
    n_g = len(genomes)
    lengths = [len(g.residues) for g in genomes]
    log(f"[rank {rank}] {n_g} genomes ({sum(lengths) / 1e6:.1f} Mb) ready in {time.perf_counter() - t0:.1f}s")

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

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pairs = synth.all_directed_pairs(n_g)
    n_pairs = len(pairs)
    shares = pdist.deal(pairs, world)
    my_pairs = np.ascontiguousarray(pairs[shares[rank]])
    # pinned host copies for the e2e leg (and the first staging)
    pinned = []
    for g in genomes:
        t = torch.empty(len(g.residues), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = g.residues
        pinned.append(t)
    h2d_bytes = int(sum(lengths) + sum(8 * len(g.contig_off) for g in genomes) + my_pairs.nbytes)

    eng = pyani_b200.Engine(local_rank)

    def stage_all():
        eng.clear()
        return [eng.add_genome(p.numpy(), g.contig_off, g.name) for p, g in zip(pinned, genomes)]

    res_dtype = pyani_b200.engine.ANIB_DTYPE if kind == "anib" else pyani_b200.engine.ANIM_DTYPE

    def gather(local):
        if world == 1:
            return local
        return pdist.all_gather_results(local, shares, device=dev, dtype=res_dtype)

    gids = stage_all()
    pack_ms = eng.last_profile()[0].get("pack", 0.0)

    share = world > 1 and os.environ.get("B200ANI_SHARE_INDEX", "1") != "0"
    keep = {}

    def shared_index_ms():
        """N > 1: the seed indexes are built once across the ranks (every rank its block of the genomes), moved by one
        all-gather over NVLink and adopted (pyani_b200.dist.share_indexes) instead of being rebuilt on every rank."""
        if not share:
            return 0.0
        torch.cuda.synchronize()
        t = time.perf_counter()
        keep["slab"] = pdist.share_indexes(eng, gids, rank, world, dev)
        torch.cuda.synchronize()
        return (time.perf_counter() - t) * 1e3

    def step():
        """One pass of the hot path; returns (device ms of the engine call(s), results, stage ms, counters)."""
        eng.reset_indexes()
        if kind == "anim":
            sh_ms = shared_index_ms()
            res = eng.anim(my_pairs)
            ms, cnt = eng.last_profile()
            ms = dict(ms)
            ms["index_share"] = sh_ms
            return ms["total"] + sh_ms, res, ms, cnt
        if kind == "anib":
            res = eng.anib(my_pairs)
            ms, cnt = eng.last_profile()
            return ms["total"], res, ms, cnt
        if kind == "tetra":
            z, obs = eng.tetra_zscores(gids)
            ms, cnt = eng.last_profile()
            corr = eng.tetra_corr(z)
            ms2, _ = eng.last_profile()
            ms = dict(ms)
            ms["tetra_corr"] = ms2.get("tetra_corr", 0.0)
            return ms.get("tetra_total", 0.0) + ms2.get("tetra_corr_total", 0.0), corr, ms, cnt
        # mixed: ANIm share + TETRA (rank 0 only does the tiny TETRA part)
        res = eng.anim(my_pairs)
        ms, cnt = eng.last_profile()
        tot = ms["total"]
        if rank == 0:
            z, obs = eng.tetra_zscores(gids)
            tot += eng.last_profile()[0].get("tetra_total", 0.0)
            eng.tetra_corr(z)
            tot += eng.last_profile()[0].get("tetra_corr_total", 0.0)
        return tot, res, ms, cnt

    # ---- resident leg
    stage_tot, counters_last, step_ms, gather_ms = {}, {}, [], []
    gather_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for it in range(args.warmup):
        _, res, _, _ = step()
        if kind != "tetra":
            gather(res)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    results = None
    t_wall = time.perf_counter()
    for it in range(args.steps):
        dev_ms, res, ms, cnt = step()
        g_ms = 0.0
        if kind != "tetra":
            gather_ev[0].record()
            results = gather(res)
            gather_ev[1].record()
            torch.cuda.synchronize()
            g_ms = gather_ev[0].elapsed_time(gather_ev[1]) if world > 1 else 0.0
        else:
            results = res
        step_ms.append(dev_ms + g_ms)
        gather_ms.append(g_ms)
        for k, v in ms.items():
            stage_tot[k] = stage_tot.get(k, 0.0) + v
        counters_last = cnt
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3 / max(1, args.steps)
    clocks = sampler.stop()
    my_ms = float(np.mean(step_ms))
    mean_ms = max_over_ranks(my_ms)
    wall_ms = max_over_ranks(wall_ms)
    n_units = n_g * (n_g - 1) // 2 if kind == "tetra" else n_pairs
    value = n_units / (mean_ms * 1e-3)

    # ---- e2e leg: host buffers in, results out, every step
    e2e_s = []
    n_e2e = 0 if args.ncu else (2 if mean_ms > 5000 else max(1, min(args.steps, 3)))
    for it in range(n_e2e + (1 if n_e2e else 0)):
        barrier()
        t1 = time.perf_counter()
        gids = stage_all()
        _, res, _, _ = step()
        if kind != "tetra":
            gather(res)
        torch.cuda.synchronize()
        if it > 0:  # first pass re-warms allocator state after the resident leg
            e2e_s.append(time.perf_counter() - t1)
    e2e_ms = max_over_ranks(float(np.mean(e2e_s)) * 1e3) if e2e_s else float("nan")
    e2e_value = n_units / (e2e_ms * 1e-3)
    d2h_bytes = int(len(my_pairs) * res_dtype.itemsize) if kind != "tetra" else int(n_g * 256 * 9 + n_g * n_g * 8)

    # per-rank stage times (attribution of the scaling residual)
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "pairs": int(len(my_pairs)), "ms_per_step": my_ms, "gather_ms": float(np.mean(gather_ms)),
                "stage_ms": {k: v / args.steps for k, v in stage_tot.items()},
                "dp_cells": int(counters_last.get("gap_cells", 0) + counters_last.get("ext_cells", 0))}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0, per-launch averages over the timed steps)
    K = args.steps
    hbm_peak, peak_src = peaks()
    traffic = ncu_traffic().get(args.workload, {}) if not (args.small or args.genomes) else {}
    line = {
        "metric": metric_name(kind), "value": value, "unit": "unordered correlations/s" if kind == "tetra" else "directed comparisons/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": dtype_of(kind), "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "unordered correlations/s" if kind == "tetra" else "directed comparisons/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms},
        "gpu_launches": int(counters_last.get("launches", 0) * K),
        "clocks": clocks, "wall_ms_per_step": wall_ms,
    }
    if args.genomes:
        line["config"]["workload"] += f" [--genomes {args.genomes} override: NOT the BASELINE size]"
    if kind == "tetra":
        bases = float(sum(lengths) + sum(len(g.contig_off) - 1 for g in genomes))
        cnt_ms = stage_tot.get("tetra_count", 0.0) / K
        alg = bases * (1 / 4 + 1 / 8)
        ach = alg / (cnt_ms * 1e-3) / 1e9 if cnt_ms else 0.0
        pack_alg = bases * (1 + 1 / 4 + 1 / 8)
        line["roofline"] = {"kernel": "tetra_count_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                            "traffic": traffic.get("tetra_count_kernel"), "peak_source": peak_src, "per_launch_ms": cnt_ms,
                            "algorithmic_bytes_per_launch": alg,
                            "note": "one shared-memory histogram update per base: in practice bound by shared-atomic throughput, not HBM",
                            "all_kernels": {"pack_kernel": {"bound": "hbm", "achieved": pack_alg / (pack_ms * 1e-3) / 1e9 if pack_ms else None, "peak": hbm_peak,
                                                            "frac": pack_alg / (pack_ms * 1e-3) / 1e9 / hbm_peak if pack_ms else None,
                                                            "ms_total": pack_ms, "algorithmic_bytes": pack_alg, "traffic": traffic.get("pack_kernel")}}}
        line["stage_ms_per_step"] = {k: v / K for k, v in stage_tot.items()}
        line["results_check"] = {"corr_shape": list(results.shape), "corr_min": float(np.min(results)), "diag_ok": bool(np.all(np.diag(results) == 1.0))}
    else:
        dpx_peak = eng.dpx_peak_gops()
        cells_gap, cells_ext = counters_last["gap_cells"], counters_last["ext_cells"]
        probes, hits, n_anchors = counters_last["probes"], counters_last["seed_hits"], counters_last["anchors"]
        kern = {
            "seed_probe_kernel": {"ms": stage_tot["seed"] / K, "launches": counters_last["seed_launches"]},
            "seed_extend_kernel": {"ms": stage_tot.get("seed_extend", 0.0) / K, "launches": counters_last["seed_launches"]},
            "gap_dp_kernel": {"ms": stage_tot["gap_dp"] / K, "launches": counters_last["gap_dp_launches"]},
            "ext_dp_kernel": {"ms": stage_tot["ext_dp_kernel"] / K, "launches": counters_last["ext_dp_launches"]},
        }
        if kind == "anib":  # the 11-mer word stage: one launch per batch like the seed probe
            kern["word_probe_kernel"] = {"ms": stage_tot.get("words", 0.0) / K, "launches": counters_last["seed_launches"]}
        word_hits = counters_last.get("word_hits", 0)

        def roofline_of(name):
            ms_k, launches = kern[name]["ms"], max(1, kern[name]["launches"])
            if name == "word_probe_kernel":
                # algorithmic bytes: both strands of every probing query once (2 bit + 1 bit per base), 8 B of group
                # starts per position, 16 B per occurrence looked at; the bound is L2 random access, reported against HBM peak
                mean_len = float(np.mean(lengths))
                qbytes = (len(my_pairs) / 2) * 2 * (mean_len / 4 + mean_len / 8 + 8 * mean_len)  # one pass per unordered pair
                alg = qbytes + word_hits * 16
                ach = alg / (ms_k * 1e-3) / 1e9 if ms_k else 0.0
                return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic.get(name), "peak_source": peak_src, "per_launch_ms": ms_k / launches,
                        "algorithmic_bytes_per_launch": alg / launches, "occurrences_per_s": word_hits / (ms_k * 1e-3) if ms_k else 0.0}
            if name.startswith("seed"):
                # algorithmic bytes: probe = 4 B k-mer streamed + 4 B presence-filter word + 8 B written per hit (+ 8 B
                # table slot per hit); extend = 8 B hit + 4 B posflag + 2 x 16 B packed sequence per hit, 16 B per anchor written
                alg = probes * 8 + hits * 16 if name == "seed_probe_kernel" else hits * 44 + n_anchors * 16
                ach = alg / (ms_k * 1e-3) / 1e9 if ms_k else 0.0
                return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic.get(name), "peak_source": peak_src, "per_launch_ms": ms_k / launches,
                        "algorithmic_bytes_per_launch": alg / launches}
            cells = cells_gap if name == "gap_dp_kernel" else cells_ext
            ach = cells * DPX_OPS_PER_CELL / (ms_k * 1e-3) / 1e9 if ms_k else 0.0
            return {"kernel": name, "bound": "alu", "achieved": ach, "peak": dpx_peak, "unit": "Gop/s (DPX-class int32 max-plus ops)",
                    "frac": ach / dpx_peak, "traffic": traffic.get(name),
                    "peak_source": "measured live: b200ani_dpx_peak (__viaddmax_s32 issue rate)",
                    "per_launch_ms": ms_k / launches, "cells_per_launch": cells / launches, "ops_per_cell": DPX_OPS_PER_CELL,
                    "gcups": cells / (ms_k * 1e-3) / 1e9 if ms_k else 0.0}

        dom = max(kern, key=lambda k: kern[k]["ms"])
        roof = roofline_of(dom)
        roof["all_kernels"] = {k: {kk: vv for kk, vv in roofline_of(k).items() if kk in ("bound", "achieved", "peak", "frac", "per_launch_ms", "gcups", "traffic")} for k in kern}
        roof["hbm_peak_gbs"] = hbm_peak
        roof["dpx_peak_gops"] = dpx_peak
        roof["dp_share_of_step"] = (kern["gap_dp_kernel"]["ms"] + kern["ext_dp_kernel"]["ms"]) / my_ms
        line["roofline"] = roof
        line["dp_cell_updates_per_s"] = (cells_gap + cells_ext) * (n_pairs / max(1, len(my_pairs))) / (mean_ms * 1e-3)
        line["unordered_pairs_per_s"] = value / 2
        line["stage_ms_per_step"] = {k: v / K for k, v in stage_tot.items()}
        line["kernel_ms_per_step"] = {k: v["ms"] for k, v in kern.items()}
        line["counters_per_step_rank0"] = counters_last
        ok = results["status"] == 0
        if kind == "anib":
            line["results_check"] = {"aligned_pairs": int(ok.sum()), "pairs": int(len(results)),
                                     "mean_identity": float(np.mean(results["pid_milli_sum"][ok] / np.maximum(results["n_pass"][ok], 1) / 1e5))}
        else:
            line["results_check"] = {"aligned_pairs": int(ok.sum()), "pairs": int(len(results)),
                                     "mean_identity": float(np.mean(results["sum_weighted"][ok] / results["sum_aligned"][ok])),
                                     "mean_coverage": float(np.mean(results["ref_aln_len"][ok] / np.array(lengths)[pairs[ok, 0]]))}
    if per_rank:
        line["per_rank"] = per_rank
    if world == 1 and not args.no_cpu and not args.small and not args.ncu:
        cores = os.cpu_count() or 1
        if kind == "tetra":
            r = run_cpu_tetra(args)
            line["cpu_baseline"] = {"value": (n_g * (n_g - 1) / 2) / (r["seconds"] / r["bases"] * float(sum(lengths))), "unit": "unordered correlations/s",
                                    "cores": 1, "kind": r["kind"], "sample": r["sample"] + "; value extrapolated to the workload's bases"}
        else:
            r = run_cpu_sample(args, "anib" if kind == "anib" else "anim", cores)
            line["cpu_baseline"] = {"value": r["value"], "unit": "directed comparisons/s", "cores": cores, "cores_busy": r["cores_busy"], "kind": "port",
                                    "sample": r["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
