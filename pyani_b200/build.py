"""Build the engine's shared library in-tree with nvcc for sm_100a (no JIT cache, no torch)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libb200ani.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "--fmad=false",              # TETRA must reproduce CPython's unfused fp64 arithmetic bit for bit
    "-Xptxas", "-v",
    "-shared",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "b200ani.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -> pyani_b200/libb200ani.so. Raises if nvcc is missing or compilation fails."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: cannot build libb200ani.so (and there is no CPU fallback)")
    objs = []
    build_dir = HERE / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = build_dir / (src.stem + ".o")
        objs.append(obj)
        deps = [src] + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "b200ani.h"]
        if not force and obj.exists() and all(d.stat().st_mtime <= obj.stat().st_mtime for d in deps):
            continue
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src.name, out))
    (build_dir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    tmp = LIB.with_suffix(".so.tmp")  # link next to the target and swap it in atomically: a snapshot never sees a half-written library
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp)] + [str(o) for o in objs] + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_lib(force="--force" in sys.argv, verbose=True))
