"""Builds raytracingrenderer_b200/librtb200.so in-tree with nvcc for sm_100a.

-fmad=false: the hit decisions must round like the reference's g++ -ffp-contract=off build
(SURVEY F9); -lineinfo: ncu source pages; no --use_fast_math anywhere.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librtb200.so")

SOURCES = ["rtb_api.cu"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".h")))  # every source and header of the library

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "--shared", "-cudart", "shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.isfile(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(ROOT, "include", "rtb.h"), os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


HOST_LIB = os.path.join(HERE, "librtb200_host.so")
HOST_DEPS = ["rtb_host.cpp", "rtb_scene.hpp", "rtb_flatten.hpp", "rtb_jpeg.hpp", "rtb_standalone.hpp"]
CLI = os.path.join(HERE, "rtb_render")


def build_host(force=False):
    """Stand-alone host layer (scene loader, reference-order BVH builder, flattener): plain g++."""
    hdir = os.path.join(HERE, "host")
    deps = [os.path.join(hdir, d) for d in HOST_DEPS] + [os.path.join(ROOT, "include", "rtb.h")]
    if not force and os.path.isfile(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-Wall",
                           os.path.join(hdir, "rtb_host.cpp"), "-o", HOST_LIB, "-lz", "-lpthread"])
    return HOST_LIB


def build_cli(force=False):
    """tools/rtb_render.cpp: the headless Main.cpp counterpart built only from this repository
    (host/standalone shims + host/Renderer.h + librtb200.so)."""
    hdir = os.path.join(HERE, "host")
    src = os.path.join(ROOT, "tools", "rtb_render.cpp")
    deps = [src, LIB, os.path.join(hdir, "Renderer.h"), os.path.join(hdir, "rtb_standalone.hpp"),
            os.path.join(hdir, "rtb_scene.hpp"), os.path.join(hdir, "rtb_flatten.hpp"), os.path.join(hdir, "rtb_jpeg.hpp")]
    if not force and os.path.isfile(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wno-unused",
                           "-I", os.path.join(hdir, "standalone"), "-I", hdir, "-I", os.path.join(ROOT, "include"),
                           src, "-o", CLI, "-L", HERE, "-lrtb200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return CLI


def build(force=False, verbose=False, extra=()):
    build_host(force)
    if not force and up_to_date():
        build_cli(force)
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-I", os.path.join(ROOT, "include")]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    build_cli(True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
