"""In-tree build of the native extension (distributed_vgg_f_b200/_C*.so) for sm_100a.

Every .cu under csrc/ is compiled with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo``
(no other arch, no PTX for older parts), bindings.cpp with g++ against torch's headers, and the
objects are linked into one shared library next to the Python package so that it travels with
the repo snapshot to the GPU box.  Incremental: an object is rebuilt only when its source or any
header under csrc/include is newer.  Usage: ``python build_native.py [-j N] [--force] [--verbose]``.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import glob
import os
import subprocess
import sys
import sysconfig
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
BUILD = os.path.join(ROOT, "build", "obj")
PKG = os.path.join(ROOT, "distributed_vgg_f_b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def ext_path() -> str:
    return os.path.join(PKG, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def _newest_header() -> float:
    hs = glob.glob(os.path.join(CSRC, "include", "*"))
    return max(os.path.getmtime(h) for h in hs) if hs else 0.0


def _stale(src: str, obj: str, hdr_time: float) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return os.path.getmtime(src) > t or hdr_time > t


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return r.stdout + r.stderr


def build(jobs: int = 8, force: bool = False, verbose: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(BUILD, exist_ok=True)
    hdr_time = _newest_header()
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cu_srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cpp_srcs = sorted(glob.glob(os.path.join(CSRC, "*.cpp")))
    inc = ["-I" + os.path.join(CSRC, "include")]
    torch_inc = []
    for p in ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames \
            else ce.include_paths(cuda=True):
        torch_inc += ["-isystem", p]
    py_inc = ["-isystem", sysconfig.get_paths()["include"]]

    tasks = []
    for src in cu_srcs:
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        if force or _stale(src, obj, hdr_time):
            cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
                   "-Xcompiler", "-fPIC", *inc, "-c", src, "-o", obj]
            tasks.append(cmd)
    for src in cpp_srcs:
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        if force or _stale(src, obj, hdr_time):
            cmd = ["g++", "-O2", "-fPIC", "-std=c++17", "-DTORCH_EXTENSION_NAME=_C",
                   "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
                   "-Wno-deprecated-declarations", *inc, *torch_inc, *py_inc,
                   "-isystem", "/usr/local/cuda/include", "-c", src, "-o", obj]
            tasks.append(cmd)
    t0 = time.time()
    if tasks:
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            list(ex.map(lambda c: _run(c, verbose), tasks))
    objs = [os.path.join(BUILD, os.path.basename(s) + ".o") for s in cu_srcs + cpp_srcs]
    out = ext_path()
    if tasks or not os.path.exists(out):
        libdirs = ce.library_paths(device_type="cuda") if "device_type" in ce.library_paths.__code__.co_varnames \
            else ce.library_paths(cuda=True)
        link = ["g++", "-shared", "-o", out, *objs]
        for d in libdirs:
            link += ["-L" + d, "-Wl,-rpath," + d]
        link += ["-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                 "-ltorch_python", "-lcudart", "-lz"]
        _run(link, verbose)
    if verbose or tasks:
        print("[build_native] %d object(s) rebuilt in %.1fs -> %s" % (len(tasks), time.time() - t0,
                                                                       os.path.relpath(out, ROOT)))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    build(a.j, a.force, a.verbose)
