"""Builds libtuatara_b200.so (CUDA kernels + C ABI) and the pytuatara extension in-tree.

nvcc cross-compiles for sm_100a without a GPU.  Objects are cached under
tuatara_b200/_build/ (git-ignored) keyed on source mtime; the shared objects land in
tuatara_b200/lib/ so they travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
INC = ROOT.parent / "include"
OBJ = ROOT / "_build"
LIB = ROOT / "lib"
LIBNAME = "libtuatara_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", f"-I{INC}", f"-I{CSRC}",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall"]


def _sources():
    return sorted(p for p in CSRC.iterdir() if p.suffix in (".cu", ".cpp"))


def _compile(src: Path, newest_header: float, verbose: bool):
    obj = OBJ / (src.name + ".o")
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, newest_header):
        return obj
    cmd = [NVCC, *ARCH, *COMMON, "-c", str(src), "-o", str(obj)]
    if src.suffix == ".cu" and verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        print(r.stderr, file=sys.stderr)
    return obj


def build(verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIB.mkdir(exist_ok=True)
    headers = [p for p in list(CSRC.iterdir()) + list(INC.iterdir()) if p.suffix in (".h", ".cuh")]
    newest_header = max([p.stat().st_mtime for p in headers] + [Path(__file__).stat().st_mtime])
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, newest_header, verbose), srcs))
    so = LIB / LIBNAME
    if not so.exists() or any(o.stat().st_mtime > so.stat().st_mtime for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(so), *map(str, objs), "-Xcompiler", "-fPIC", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_pybind(so)
    return so


def build_pybind(so: Path) -> Path | None:
    """pytuatara: the reference's Python module (bindings/python.cpp:54-58) on top of the C++ API."""
    src = ROOT / "bindings" / "python.cpp"
    if not src.exists():
        return None
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = LIB / f"pytuatara{ext}"
    if out.exists() and out.stat().st_mtime > max(src.stat().st_mtime, so.stat().st_mtime,
                                                  (INC / "tuatara.h").stat().st_mtime):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{INC}", f"-I{pybind11.get_include()}",
           f"-I{sysconfig.get_paths()['include']}", str(src), "-o", str(out),
           f"-L{LIB}", "-ltuatara_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"pytuatara build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
