"""Build libader_b200.so in-tree with nvcc for sm_100a (no torch headers: the ABI is plain C)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libader_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libader_b200 needs the CUDA 12.9 toolkit")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "ader_b200.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant / defines: a second library beside the product one (e.g. variant="tl", defines=["-DADER_TC_TIMELINE"] for
    the phase-stamp debug build, loaded with ADER_B200_LIB=<path>); the default call builds the product library."""
    lib_path = LIB_PATH if not variant else os.path.join(LIB_DIR, "libader_b200_%s.so" % variant)
    if not variant and not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + (".o" if not variant else ".%s.o" % variant))
        cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], *os.environ.get("ADER_B200_DEFINES", "").split(),
               *defines, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return lib_path


if __name__ == "__main__":
    if "--timeline" in sys.argv:
        print(build(force=True, verbose="-v" in sys.argv, variant="tl", defines=["-DADER_TC_TIMELINE"]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
