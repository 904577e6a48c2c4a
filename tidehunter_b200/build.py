"""Builds the native pieces in-tree (so the .so files travel to the GPU box with the snapshot):

  tidehunter_b200/libth_gpu.so   CUDA kernels + C ABI (include/th_gpu.h), sm_100a only
  tidehunter_b200/libth_host.so  host C layer (host/th_host.c) linked against libth_gpu.so
  host/tidehunter-b200           command line front end with TideHunter's flags
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tidehunter_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(ROOT, "host")
GPU_SO = os.path.join(PKG, "libth_gpu.so")
HOST_SO = os.path.join(PKG, "libth_host.so")
CLI = os.path.join(HOST, "tidehunter-b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=False):
    cu_src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "th_gpu.h")]
    if force or _newer(GPU_SO, cu_src):
        extra = os.environ.get("TH_NVCC_FLAGS", "").split()   # tuning experiments only (e.g. -DPOA_MIN_BLOCKS=6)
        out = _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"] + extra + ["-o", GPU_SO, os.path.join(CSRC, "th_api.cu")])
        if verbose:
            print(out)
    host_src = [os.path.join(HOST, "th_host.c"), os.path.join(HOST, "th_host.h"), os.path.join(ROOT, "include", "th_gpu.h"), GPU_SO]
    if force or _newer(HOST_SO, host_src):
        _run(["gcc", "-std=gnu99", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-shared", "-o", HOST_SO,
              os.path.join(HOST, "th_host.c"), "-L" + PKG, "-lth_gpu", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"])
    if force or _newer(CLI, [os.path.join(HOST, "th_main.c"), os.path.join(HOST, "th_reader.h"), HOST_SO]):
        _run(["gcc", "-std=gnu99", "-O2", "-ffp-contract=off", "-Wall", "-o", CLI, os.path.join(HOST, "th_main.c"),
              "-L" + PKG, "-lth_host", "-lth_gpu", "-Wl,-rpath,$ORIGIN/../tidehunter_b200", "-lz", "-lm", "-lpthread"])
    return GPU_SO, HOST_SO, CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
