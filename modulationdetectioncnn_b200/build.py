"""Build libmdc.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmdc.so")
SOURCES = ["mdc_api.cu", "q612.cu", "tiny_f32.cu", "fwht.cu", "sdr.cu", "vt_f32.cu", "vt_tensor.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math=false",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = os.path.join(HERE, "build", "flags.txt")


def _flag_stamp() -> str:
    return " ".join(NVCC_FLAGS + os.environ.get("MDC_NVCC_EXTRA", "").split())


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    # a library built with other flags (e.g. -DMDC_VT_ABLATE for timing experiments) is never reused
    if not os.path.exists(STAMP) or open(STAMP).read() != _flag_stamp():
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "mdc.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("MDC_NVCC_EXTRA", "").split()          # e.g. -DMDC_VT_ABLATE for timing experiments
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
           "-Xcompiler", "-fPIC"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as fh:
        fh.write(_flag_stamp())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
