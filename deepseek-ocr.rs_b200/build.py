"""Build libdsocr.so (sm_100a only) in-tree with nvcc.  No torch dependency: the library links
against libcudart only and exposes the C ABI declared in include/dsocr.h."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIBDIR = HERE / "lib"
OBJDIR = HERE / "build"
LIB = LIBDIR / "libdsocr.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math=false",
    "-I", str(HERE.parent / "include"), "-I", str(CSRC),
]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]


def _deps_hash(src: Path) -> str:
    h = hashlib.sha1()
    h.update(" ".join(FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((HERE.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path) -> Path:
    obj = OBJDIR / (src.name + ".o")
    stamp = OBJDIR / (src.name + ".sha1")
    digest = _deps_hash(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return obj
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return obj


def build(verbose: bool = True) -> Path:
    OBJDIR.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    srcs = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB} ({LIB.stat().st_size / 1e6:.1f} MB) from {len(srcs)} sources")
    return LIB


if __name__ == "__main__":
    build()
