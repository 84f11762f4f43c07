#!/usr/bin/env python3
"""Builds the in-tree native libraries.

  text-based-image-style-transfer_b200/libnst_b200.so      all CUDA kernels + the C ABI (sm_100a only)
  text-based-image-style-transfer_b200/liblbfgs_ctl_host.so host build of the L-BFGS controller (CPU unit tests)
  text-based-image-style-transfer_b200/libnst_b200_instr.so  (--instrument only) the same sources with -DNST_INSTRUMENT: phase
      stamps, wait counters, launch spans and the wrong-result timing experiments used by tools/conv_phases.py,
      conv_timeline.py and ctl_phases.py; never loaded by the product, the tests or bench.py
  oracle/_ref is not needed: the reference is Python (see oracle/README.md)

nvcc cross-compiles without a GPU; the .so files are git-ignored but travel to the GPU box.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "text-based-image-style-transfer_b200")
CSRC = os.path.join(PKG, "csrc")
CU = ["api.cu", "conv_tc.cu", "conv1_tc.cu", "gram.cu", "pixel.cu", "lbfgs.cu", "loss_fn.cu", "mask.cu", "video.cu", "depth.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False, instrument=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "nst_b200.h")]
    if instrument:
        _build_cuda(nvcc, deps, os.path.join(PKG, "build_instr"), os.path.join(PKG, "libnst_b200_instr.so"), ["-DNST_INSTRUMENT"],
                    force, verbose)
    lib = os.path.join(PKG, "libnst_b200.so")
    _build_cuda(nvcc, deps, os.path.join(PKG, "build"), lib, [], force, verbose)
    host = os.path.join(PKG, "liblbfgs_ctl_host.so")
    hsrc = os.path.join(CSRC, "lbfgs_ctl_host.cpp")
    if force or _newer(host, [hsrc, os.path.join(CSRC, "lbfgs_ctl.h")]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", host, hsrc])
    return lib, host


def _build_cuda(nvcc, deps, objdir, lib, defines, force, verbose):
    os.makedirs(objdir, exist_ok=True)
    if force or _newer(lib, deps):
        objs = []
        procs = []
        for f in CU:
            o = os.path.join(objdir, f.replace(".cu", ".o"))
            objs.append(o)
            cmd = [nvcc] + NVCC_FLAGS + defines + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for cmd, pr in procs:
            out, _ = pr.communicate()
            if verbose or pr.returncode != 0:
                sys.stderr.write(out)
            if pr.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, instrument="--instrument" in sys.argv))
