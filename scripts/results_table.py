#!/usr/bin/env python
"""Bench lines under profiles/ (r02_*.json) -> the markdown table of profiles/README.md."""
import json
import sys
from pathlib import Path

P = Path(__file__).resolve().parents[1] / "profiles"


def load(name):
    f = P / name
    if not f.exists():
        return None
    for l in open(f):
        if l.startswith("{"):
            return json.loads(l)


def main():
    rows = []
    ref = load("r02_reference_n1.json")
    n1 = load("r02_bench_n1.json")
    for tag, name in (("1", "r02_bench_n1.json"), ("2", "r02_bench_n2.json"), ("4", "r02_bench_n4.json"), ("8", "r02_bench_n8.json")):
        d = load(name)
        if not d:
            continue
        eff = d["value"] / (int(tag) * n1["value"]) if n1 else float("nan")
        r = d["roofline"]["all_kernels"]
        rows.append(f"| ANIm 500 x ~5 Mb (configs[3]), N = {tag} | {d['value']:,.0f} comparisons/s | {d['ms_per_step']:,.0f} ms | {d['e2e']['value']:,.0f} /s | "
                    f"{eff:.3f} | ext DP {r['ext_dp_kernel']['frac']:.3f}, link DP {r['gap_dp_kernel']['frac']:.3f}, probe {r['seed_probe_kernel']['frac']:.3f} | `{name}` |")
    for label, name in (("ANIm 50 x ~5 Mb (configs[1])", "r02_bench_anim50.json"), ("ANIb 200 x ~5 Mb (configs[2])", "r02_bench_anib.json"),
                        ("ANIm + TETRA, 1000 x 0.5-12 Mb (configs[4]), N = 8", "r02_mixed_n8.json")):
        d = load(name)
        if not d:
            continue
        r = d.get("roofline", {}).get("all_kernels", {})
        rf = ", ".join(f"{k.replace('_kernel', '')} {v['frac']:.3f}" for k, v in r.items() if v.get("frac") is not None)
        rows.append(f"| {label} | {d['value']:,.0f} comparisons/s | {d['ms_per_step']:,.1f} ms | {d['e2e']['value']:,.0f} /s | | {rf} | `{name}` |")
    t = load("r02_bench_tetra.json")
    if t:
        rows.append(f"| TETRA 1000 x 0.5-12 Mb (configs[4], TETRA half) | {t['value']:,.0f} correlations/s | {t['ms_per_step']:.1f} ms | {t['e2e']['value']:,.0f} /s | | "
                    f"count {t['roofline']['frac']:.3f} of HBM | `r02_bench_tetra.json` |")
    print("| workload | value (resident) | per step | e2e (host buffers) | scaling efficiency | roofline fractions | file |")
    print("|---|---|---|---|---|---|---|")
    print("\n".join(rows))
    for name in ("r02_reference_n1.json", "r02_bench_n1.json", "r02_bench_anib.json", "r02_bench_tetra.json"):
        d = load(name)
        if d and d.get("cpu_baseline"):
            c = d["cpu_baseline"]
            print(f"\nCPU arm ({name}): {c['value']:.1f} {c['unit']} on {c.get('cores_busy', c['cores'])} of {c['cores']} host threads busy, kind `{c['kind']}`: {c['sample']}")


if __name__ == "__main__":
    main()
