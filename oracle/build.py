"""TEST INFRASTRUCTURE ONLY (oracle).  Build recipes for the CPU checkers.

* ``build_oracle()``  — compiles oracle/oracle_native.cpp (our CPU restatement) into oracle/_build/liboracle.so.
* ``build_ref()``     — when /root/reference is present (the build container), compiles the *unmodified* reference
  sources where they lie (knn_.cxx + nanoflann.hpp; grid_subsampling.cpp + cloud.cpp) behind the two extern "C"
  shims of this directory into oracle/_ref/.  Outputs only go to oracle/_ref/ (git-ignored, NOT gpurun-ignored, so
  the prebuilt .so files travel to the GPU box, where /root/reference does not exist).  No reference source is copied.

The reference's own build system is not run: its Cython pyx names a non-existent ``knn.cxx`` (knn.pyx:2) and its
subsampling wrapper needs the numpy-1.x C API (wrapper.cpp:2,104-106) — see DESIGN.md §Oracle.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"   # $CXX=/opt/gcc/bin/g++ cannot link -fopenmp here


def _run(cmd):
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def build_oracle(force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(HERE, "oracle_native.cpp")
    if force or _stale(out, [src]):
        _run([CXX, "-O2", "-std=c++14", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", out])
    return out


def ref_available() -> bool:
    return os.path.isdir(os.path.join(REF, "utils", "nearest_neighbors"))


def build_ref(force: bool = False) -> dict:
    """Returns {'knn': path|None, 'subsample': path|None}; builds only if the reference tree is mounted."""
    out_dir = os.path.join(HERE, "_ref")
    knn_so = os.path.join(out_dir, "libref_knn.so")
    sub_so = os.path.join(out_dir, "libref_subsample.so")
    if ref_available():
        os.makedirs(out_dir, exist_ok=True)
        nn = os.path.join(REF, "utils", "nearest_neighbors")
        cw = os.path.join(REF, "utils", "cpp_wrappers")
        knn_srcs = [os.path.join(HERE, "ref_knn_shim.cpp"), os.path.join(nn, "knn_.cxx")]
        if force or _stale(knn_so, knn_srcs):
            # flags of utils/nearest_neighbors/setup.py:13 (+ -O2: distutils' default optimisation level)
            _run([CXX, "-O2", "-std=c++11", "-fopenmp", "-shared", "-fPIC", "-w", "-I", nn, *knn_srcs, "-o", knn_so])
        sub_srcs = [os.path.join(HERE, "ref_subsample_shim.cpp"),
                    os.path.join(cw, "cpp_subsampling", "grid_subsampling", "grid_subsampling.cpp"),
                    os.path.join(cw, "cpp_utils", "cloud", "cloud.cpp")]
        if force or _stale(sub_so, sub_srcs):
            # flags of utils/cpp_wrappers/cpp_subsampling/setup.py:18-19
            _run([CXX, "-O2", "-std=c++11", "-D_GLIBCXX_USE_CXX11_ABI=0", "-shared", "-fPIC", "-w",
                  "-I", os.path.join(cw, "cpp_subsampling"), *sub_srcs, "-o", sub_so])
    return {"knn": knn_so if os.path.exists(knn_so) else None,
            "subsample": sub_so if os.path.exists(sub_so) else None}


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
