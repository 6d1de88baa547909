"""In-tree build of ``libwicca_b200.so`` (nvcc, sm_100a only).

The shared library is written to ``wicca_b200/lib/`` so that it travels with the
source tree; there is no JIT cache and no pip install step.  nvcc cross-compiles
without a GPU, so this also runs in a CPU-only container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIBNAME = "libwicca_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libwicca_b200.so cannot be built (there is no CPU fallback)")


def lib_path() -> Path:
    return LIBDIR / LIBNAME


def _newest_source_mtime() -> float:
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    srcs.append(PKG.parent / "include" / "wicca_b200.h")
    return max(p.stat().st_mtime for p in srcs if p.exists())


def is_stale() -> bool:
    lib = lib_path()
    return (not lib.exists()) or lib.stat().st_mtime < _newest_source_mtime()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` and link the shared library.  Returns its path.
    ``WICCA_DEV=1`` in the environment adds ``-DWICCA_DEV`` (kernel knobs for ``tools/``; never for a release)."""
    lib = lib_path()
    if not force and not is_stale():
        return lib
    nvcc = _nvcc()
    OBJDIR.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    hdr_mtime = max(p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG.parent / "include" / "wicca_b200.h"])

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if (not force) and obj.exists() and obj.stat().st_mtime >= max(src.stat().st_mtime, hdr_mtime):
            return obj
        dev = ["-DWICCA_DEV"] if os.environ.get("WICCA_DEV") == "1" else []
        cmd = [nvcc, *NVCC_FLAGS, *dev, "-Xptxas", "-v", "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(" ".join(cmd))
            print(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{res.stderr}")
        (OBJDIR / (src.stem + ".ptxas.log")).write_text(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    tmp = lib.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *map(str, objs),
           "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined", "-lpthread", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stderr}")
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
