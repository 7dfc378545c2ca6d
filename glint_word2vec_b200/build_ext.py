"""In-tree build of the native extensions.

``_C.so``    CUDA kernels for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)
             + torch bindings (csrc/bindings.cpp)
``_host.so`` host-side text processing (csrc/host/textproc.cpp, pybind11 only)

The objects are built straight into the package directory so the ``.so`` files
travel with a snapshot of the repository (they are git-ignored).  nvcc
cross-compiles without a GPU.  Usage: ``python -m glint_word2vec_b200.build_ext [--force]``.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")

# csrc/legacy/ (the superseded round-1 kernel generations) is deliberately not part of the build
CU_SOURCES = ["sgns_pairs.cu", "pairgen.cu", "prep_kernels.cu", "infer_kernels.cu", "nn_tc.cu", "nn_select.cu", "serve_fused.cu", "umma_probe.cu", "sgns_tile.cu"]
CPP_SOURCES = ["bindings.cpp", "bindings_tile.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return os.path.join(cuda_home, "bin", "nvcc")


def _newer(src_list, out) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build command failed: " + " ".join(cmd))
    return r.stdout + r.stderr


def build_cuda(force=False, verbose=False) -> str:
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(PKG, "_C.so")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    objs, jobs = [], []
    for cu in CU_SOURCES:
        src = os.path.join(CSRC, cu)
        obj = os.path.join(BUILD, cu.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            jobs.append([_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-c", src, "-o", obj])
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    for cpp in CPP_SOURCES:
        bsrc = os.path.join(CSRC, cpp)
        bobj = os.path.join(BUILD, cpp.replace(".cpp", ".o"))
        objs.append(bobj)
        if force or _newer([bsrc] + headers, bobj):
            jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
                         "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H", "-I", CSRC] + inc +
                        ["-c", bsrc, "-o", bobj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or force or not os.path.exists(out):
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        link = ["g++", "-shared", "-o", out] + objs + [
            f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            f"-L{cuda_home}/lib64", "-lcudart", f"-Wl,-rpath,{tlib}", f"-Wl,-rpath,{cuda_home}/lib64"]
        _run(link, verbose)
    return out


def build_host(force=False, verbose=False) -> str:
    import pybind11
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(CSRC, "host", "textproc.cpp")
    out = os.path.join(PKG, "_host.so")
    if force or _newer([src], out):
        cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", pybind11.get_include(),
               "-I", sysconfig.get_paths()["include"], src, "-o", out]
        _run(cmd, verbose)
    return out


def build_all(force=False, verbose=False):
    return [build_host(force, verbose), build_cuda(force, verbose)]


if __name__ == "__main__":
    outs = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    for o in outs:
        print("built", o)
