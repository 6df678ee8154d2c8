"""nvcc driver: builds ``libdynoed_cuda.so`` (compiled-in models) and per-model libraries for
user-defined systems ("emit the RHS into the kernel template, compile once").

sm_100a only; extensions are built in-tree so the built ``.so`` travels with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LAST_BUILD: dict = {}      # library path -> "compiled by nvcc in N s" | "reused (...)", for build()'s report
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_library(out_path: str, models_dir: str | None = None, force: bool = False,
                  verbose: bool = False) -> str:
    """Compile csrc/dynoed_api.cu against ``<models_dir>/models.cuh`` into ``out_path``."""
    models_dir = models_dir or os.path.join(CSRC, "generated")
    srcs = [os.path.join(CSRC, "dynoed_api.cu")]
    srcs += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    srcs += [os.path.join(_HERE, "..", "include", "dynoed.h")]
    srcs += [os.path.join(models_dir, f) for f in sorted(os.listdir(models_dir)) if f.endswith(".cuh")]
    if not force and _newer(out_path, srcs):
        LAST_BUILD[os.path.abspath(out_path)] = "reused (up to date with its sources)"
        return out_path
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += [f"-D{d}" for d in os.environ.get("DYNOED_NVCC_DEFINES", "").split() if d]
    if os.path.abspath(models_dir) != os.path.abspath(os.path.join(CSRC, "generated")):
        cmd += [f'-DDYNOED_MODELS_HEADER="{os.path.join(os.path.abspath(models_dir), "models.cuh")}"']
    # nvcc writes to a private temporary next to the target; os.replace() publishes it atomically, so
    # concurrent ranks (torchrun) never dlopen a half-written library nor write the same file
    tmp_path = f"{out_path}.{os.getpid()}.tmp"
    cmd += ["-o", tmp_path, srcs[0]]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp_path):
            os.unlink(tmp_path)
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp_path, out_path)
    LAST_BUILD[os.path.abspath(out_path)] = f"compiled by nvcc in {time.perf_counter() - t0:.0f} s"
    if verbose:
        print(r.stderr)
    return out_path


def build_default(force: bool = False, regenerate: bool = True, verbose: bool = False) -> str:
    """(Re-)emit the compiled-in models and build ``libdynoed_cuda.so`` in-tree."""
    gen = os.path.join(CSRC, "generated")
    if regenerate:
        from .emit import emit_builtin
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            emit_builtin(tmp)
            for f in os.listdir(tmp):       # only touch files whose content changed
                new = open(os.path.join(tmp, f)).read()
                dst = os.path.join(gen, f)
                if not os.path.exists(dst) or open(dst).read() != new:
                    with open(dst, "w") as fh:
                        fh.write(new)
    return build_library(os.path.join(_HERE, "libdynoed_cuda.so"), gen, force=force, verbose=verbose)


def build_user_model(src: str, info: dict, cache_dir: str | None = None) -> str:
    """Compile a library that contains exactly one emitted model; returns the .so path."""
    cache_dir = cache_dir or os.path.join(_HERE, "_user_models")
    h = hashlib.sha256(src.encode())
    # the kernel template and the ABI are part of the key: a cached library built against older
    # headers must never be loaded by a newer host
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(_HERE, "..", "include", "dynoed.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("DYNOED_NVCC_DEFINES", "").encode())
    key = h.hexdigest()[:16]
    d = os.path.join(cache_dir, key)
    out = os.path.join(d, f"libdynoed_{key}.so")
    if os.path.exists(out):             # published atomically by build_library: complete if present
        return out
    os.makedirs(d, exist_ok=True)
    name = info["name"]
    with open(os.path.join(d, f"model_{name}.cuh"), "w") as fh:
        fh.write(src)
    with open(os.path.join(d, "models.cuh"), "w") as fh:
        fh.write(f'#pragma once\n#include "model_{name}.cuh"\n'
                 f"#define DYNOED_FOR_EACH_MODEL(X) X(Model_{name})\n")
    return build_library(out, d, force=True)
