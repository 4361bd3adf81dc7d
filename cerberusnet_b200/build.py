"""Builds libcerberus_costvolume.so in-tree with nvcc for sm_100a (no torch headers involved)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libcerberus_costvolume.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["costvolume_fwd.cu", "costvolume_fwd_tc.cu", "costvolume_bwd.cu", "costvolume_bwd_tc.cu", "costvolume_splat.cu", "costvolume_api.cu", "costvolume_sampler.cu", "photometric.cu"]
HEADERS = ["costvolume_common.cuh", "costvolume_launch.h",
           os.path.join("..", "..", "include", "cerberus_costvolume.h"),
           os.path.join("..", "..", "include", "cerberus_trt_plugin.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    return os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> str:
    if not force and not _stale():
        return LIB_PATH
    # one builder at a time (torchrun starts every rank at once): the others wait on the lock and then find
    # the library up to date; the link goes to a temporary name and is renamed into place atomically
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            return _build_locked(verbose, extra_flags)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, extra_flags: list[str] | None) -> str:
    print(f"[cerberusnet_b200] building {LIB_NAME} with nvcc (sm_100a) ...", file=sys.stderr, flush=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + ["-ccbin", ccbin] + (extra_flags or [])

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout[-6000:]}")
        if verbose and r.stdout.strip():
            print(r.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [_nvcc(), "-shared", "-ccbin", ccbin, "-o", tmp, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout[-4000:]}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True,
                        extra_flags=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else None))
