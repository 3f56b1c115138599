"""Build liblanczos_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension:
the library is a plain C-ABI shared object loaded with ctypes).

    python -m lanczos_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblanczos_b200.so")
SYNTH_LIB = os.path.join(HERE, "liblz_synth.so")          # synthetic benchmark inputs (include/lz_synth.h)
SYNTH_SOURCES = ["synth_rgg.cu"]
SOURCES = ["capi.cu", "stencil.cu", "stencil27.cu", "vecops.cu", "reorth.cu", "spmv.cu", "sellw.cu", "fused.cu", "lanczos.cu", "potential.cu", "kba.cu", "small.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build liblanczos_b200.so")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "lanczos_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_synth(force: bool = False) -> str:
    """liblz_synth.so: device-side generator of the config-4 random geometric graph (bench/tests)."""
    srcs = [os.path.join(CSRC, s) for s in SYNTH_SOURCES]
    deps = srcs + [os.path.join(os.path.dirname(HERE), "include", "lz_synth.h")]
    if not force and os.path.exists(SYNTH_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(SYNTH_LIB) for d in deps):
        return SYNTH_LIB
    res = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-o", SYNTH_LIB] + srcs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return SYNTH_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    build_synth(force)
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB


def build_variant(out: str, defines: list[str]) -> str:
    """Kernel-tuning aid: another build of the same library with extra -D macros, loaded through
    LANCZOS_B200_LIB (see _capi.py).  python -m lanczos_b200.build --variant out.so -DLZ_KBA_MINBLOCKS=3"""
    cmd = [_nvcc()] + NVCC_FLAGS + list(defines) + ["-o", out] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
