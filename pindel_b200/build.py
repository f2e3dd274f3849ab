"""Build every native artefact in-tree (nvcc cross-compiles sm_100a without a GPU).

  pindel_b200/libpindel_pg.so   the product: CUDA kernels + C ABI (include/pindel_pg.h)
  oracle/libpg_oracle.so        test infrastructure: C restatement of the reference algorithm (libsv_oracle.so: the SV drivers)
  oracle/_ref/*                 test infrastructure: the reference itself, when /root/reference is present
  tests/emul/libpg_emul.so      test infrastructure: CPU lane emulator of the warp program
  tests/emul/libsv_emul.so      test infrastructure: the SV-classification program compiled for the host (librp_emul.so: read-pair clustering)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpindel_pg.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--compiler-bindir",
              "/usr/bin/g++", "-shared", "-Xcompiler", "-fPIC"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "pindel_pg.h")]
    if force or _stale(LIB, srcs):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "pindel_pg.cu")]
        subprocess.check_call(cmd)
    return LIB


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    if os.path.exists("/root/reference/src/pindel.cpp"):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"])


def build_emulator():
    sys.path.insert(0, ROOT)
    from tests.emul import bam_emul, bgzf_emul, boxsort_emul, emul, rp_emul, sr_emul, sv_emul
    emul.build()
    sv_emul.build()
    rp_emul.build()
    bgzf_emul.build()
    bam_emul.build()
    boxsort_emul.build()
    sr_emul.build()


def build_all(force=False):
    build_cuda(force)
    build_oracle()
    build_emulator()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB)
