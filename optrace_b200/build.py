"""In-tree build of the CUDA engine (nvcc, sm_100a).  Used by __graft_entry__.build() and by userfunc.py for
scene-specialised variants that inline user callables as device functions."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import pathlib
import subprocess

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
SOURCES = ["otb_api.cu", "otb_trace.cu", "otb_render.cu", "otb_detect.cu", "otb_gen.cu", "otb_image.cu", "otb_focus.cu", "otb_spectrum.cu", "otb_tiles.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",           # op-for-op parity with the reference's numpy arithmetic
              "-Xcompiler", "-fPIC", f"-I{ROOT / 'include'}", f"-I{CSRC}"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and pathlib.Path(c).exists():
            return c
    return "nvcc"


_digest_cache = None


def source_digest() -> str:
    """hash of the engine sources + flags (computed once per process)"""
    global _digest_cache
    if _digest_cache is None:
        _digest_cache = _source_digest()
    return _digest_cache


def _source_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "otb.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS[:8]).encode())
    return h.hexdigest()


def build_library(out_path=None, extra_flags=(), force=False, objdir=None, sources=None, extra_objects=()) -> pathlib.Path:
    """Compiles the engine sources in parallel and links a shared library."""
    out_path = pathlib.Path(out_path) if out_path else CSRC / "libotb.so"
    base_build = objdir is None
    # variant builds get a private object directory per process: several ranks (torchrun) or threads may miss the
    # cache for the same variant at the same time; the finished library is moved into place atomically
    objdir = CSRC / "build" if base_build else pathlib.Path(f"{objdir}.{os.getpid()}")
    objdir.mkdir(parents=True, exist_ok=True)
    stamp = objdir / (out_path.name + ".digest")
    digest = source_digest() + "|" + " ".join(extra_flags)
    if not force and out_path.exists() and stamp.exists() and stamp.read_text() == digest:
        return out_path
    cc = nvcc()
    srcs = list(sources or SOURCES)

    # variant builds (user callables, scene specialisation) carry no line info: every variant travels to the GPU box
    flags = NVCC_FLAGS if base_build else [f for f in NVCC_FLAGS if f != "-lineinfo"]

    def compile_one(src):
        obj = objdir / (src.replace(".cu", ".o"))
        r = subprocess.run([cc, *flags, *extra_flags, "-c", str(CSRC / src), "-o", str(obj)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp_out = out_path.with_name(f".{out_path.name}.{os.getpid()}.tmp")
    r = subprocess.run([cc, "-shared", "-o", str(tmp_out), *[str(o) for o in objs], *[str(o) for o in extra_objects],
                        "-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_out, out_path)
    if base_build:
        stamp.write_text(digest)
    else:
        import shutil
        shutil.rmtree(objdir, ignore_errors=True)
    return out_path
