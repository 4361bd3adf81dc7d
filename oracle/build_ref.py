#!/usr/bin/env python
"""Compile the reference's own CUDA correlation op, UNMODIFIED, from where it lies.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Sources (never copied): <ref>/nnet_training/correlation_package/correlation_cuda.cpp and
correlation_cuda_kernel.cu.  The reference's setup.py is not used (it pins compute_86,
/usr/local/cuda-11.1 and -std=c++14, setup.py:10-15,25); nvcc / g++ are driven directly for
sm_100a.  Output: oracle/_ref/correlation_ref.so, which registers
``torch.ops.cerberus.correlation`` / ``correlation_backward`` when loaded with
``torch.ops.load_library`` (correlation_cuda.cpp:45-48).  oracle/_ref/ is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "correlation_ref.so")


def build(ref_root: str = "/root/reference", force: bool = False, verbose: bool = True) -> str | None:
    src_dir = os.path.join(ref_root, "nnet_training", "correlation_package")
    cpp = os.path.join(src_dir, "correlation_cuda.cpp")
    cu = os.path.join(src_dir, "correlation_cuda_kernel.cu")
    if not (os.path.exists(cpp) and os.path.exists(cu)):
        if verbose:
            print(f"[build_ref] reference sources not found under {ref_root}; keeping any prebuilt {OUT_SO}")
        return OUT_SO if os.path.exists(OUT_SO) else None
    if os.path.exists(OUT_SO) and not force:
        newest = max(os.path.getmtime(cpp), os.path.getmtime(cu))
        if os.path.getmtime(OUT_SO) >= newest:
            return OUT_SO
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{src_dir}"]
    import sysconfig
    inc.append(f"-I{sysconfig.get_paths()['include']}")
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    obj_cu = os.path.join(OUT_DIR, "correlation_cuda_kernel.o")
    obj_cpp = os.path.join(OUT_DIR, "correlation_cuda.o")
    cmds = [
        [nvcc, "-ccbin", gxx, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", abi, *inc, "-c", cu, "-o", obj_cu],
        [gxx, "-std=c++17", "-O2", "-fPIC", abi, *inc, "-c", cpp, "-o", obj_cpp],
    ]
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for c in cmds]
    for c, pr in zip(cmds, procs):
        out, _ = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out.decode(errors="replace")[-4000:])
            raise RuntimeError(f"[build_ref] failed: {' '.join(c[:3])} ...")
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    link = [gxx, "-shared", "-o", OUT_SO, obj_cu, obj_cpp, f"-L{tlib}", f"-L{cuda_lib}", "-ltorch", "-ltorch_cpu",
            "-ltorch_cuda", "-lc10", "-lc10_cuda", "-lcudart", f"-Wl,-rpath,{tlib}"]
    subprocess.check_call(link)
    for o in (obj_cu, obj_cpp):
        os.remove(o)
    if verbose:
        print(f"[build_ref] built {OUT_SO}")
    return OUT_SO


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    build(a.ref, a.force)
