"""In-tree build of libkfunca_b200.so (C ABI + sm_100a kernels) and the pybind11 module _kfunca.

Generates a ninja file under kfunca_b200/_build/ and runs it: nvcc cross-compiles for sm_100a
(`-gencode arch=compute_100a,code=sm_100a -lineinfo`) without a GPU.  Outputs land next to this file so
they travel with the repo snapshot to the GPU box.  Usage: `python -m kfunca_b200.build [-v] [--clean]`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))

HOST_SRCS = ["runtime.cpp", "tensor.cpp", "plan.cpp", "ops.cpp", "api.cpp", "dist.cpp"]
LIB_NAME = "libkfunca_b200.so"
EXT_NAME = "_kfunca" + sysconfig.get_config_var("EXT_SUFFIX")

NVCC_FLAGS = (
    "-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr "
    "-Xcompiler -fPIC -Xcompiler -fvisibility=hidden -diag-suppress 20058"
)
CXX_FLAGS = f"-O2 -std=c++17 -fPIC -fvisibility=hidden -I{CUDA_HOME}/include -Wall -Wno-unused-function"


def cuda_sources() -> list[str]:
    return sorted(p.name for p in (CSRC / "kernels").glob("*.cu"))


def write_ninja() -> Path:
    import pybind11

    BUILD.mkdir(exist_ok=True)
    py_inc = sysconfig.get_paths()["include"]
    pb_inc = pybind11.get_include()
    lines = [
        f"nvcc = {CUDA_HOME}/bin/nvcc",
        "cxx = g++",
        f"nvflags = {NVCC_FLAGS}",
        f"cxxflags = {CXX_FLAGS}",
        "rule cu",
        "  command = $nvcc $nvflags -MD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "  description = NVCC $in",
        "rule cc",
        "  command = $cxx $cxxflags -MD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "  description = CXX $in",
        "rule link_lib",
        "  command = $nvcc -shared -o $out $in -cudart static -ldl -Xlinker -soname=" + LIB_NAME,
        "  description = LINK $out",
        "rule ext",
        f"  command = $cxx -O2 -std=c++17 -fPIC -shared -fvisibility=hidden -I{py_inc} -I{pb_inc} $in -o $out "
        f"-L{PKG} -lkfunca_b200 -Wl,-rpath,'$$ORIGIN'",
        "  description = PYBIND $out",
    ]
    objs = []
    for s in cuda_sources():
        o = BUILD / (s + ".o")
        lines.append(f"build {o}: cu {CSRC / 'kernels' / s}")
        objs.append(str(o))
    for s in HOST_SRCS:
        o = BUILD / (s + ".o")
        lines.append(f"build {o}: cc {CSRC / s}")
        objs.append(str(o))
    lines.append(f"build {PKG / LIB_NAME}: link_lib {' '.join(objs)}")
    lines.append(f"build {PKG / EXT_NAME}: ext {CSRC / 'pybind_module.cpp'} | {PKG / LIB_NAME}")
    lines.append(f"default {PKG / EXT_NAME}")
    path = BUILD / "build.ninja"
    path.write_text("\n".join(lines) + "\n")
    return path


def build(verbose: bool = False) -> None:
    ninja = shutil.which("ninja")
    if ninja is None:
        try:
            import ninja as _n  # the pip package ships the binary

            ninja = str(Path(_n.BIN_DIR) / "ninja")
        except Exception as e:  # pragma: no cover
            raise RuntimeError("ninja not found") from e
    nf = write_ninja()
    cmd = [ninja, "-f", str(nf), "-C", str(BUILD)] + (["-v"] if verbose else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        sys.stdout.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("kfunca_b200 build failed")


def is_built() -> bool:
    return (PKG / LIB_NAME).exists() and (PKG / EXT_NAME).exists()


if __name__ == "__main__":
    if "--clean" in sys.argv:
        shutil.rmtree(BUILD, ignore_errors=True)
        for f in (PKG / LIB_NAME, PKG / EXT_NAME):
            f.unlink(missing_ok=True)
    build(verbose="-v" in sys.argv)
    print("built", PKG / LIB_NAME, PKG / EXT_NAME)
