"""Builds tests/mock/_build/libmodelxdigest_mock.so: the HOST sources of libmodelxdigest.so compiled with g++ against
a synchronous stand-in for the CUDA runtime, with CPU kernel launchers that hash through the oracle.
TEST INFRASTRUCTURE ONLY (see tests/mock/include/cuda_runtime.h); the product never loads it."""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock")
OUT = os.path.join(MOCK, "_build", "libmodelxdigest_mock.so")
CSRC = os.path.join(ROOT, "modelx_b200", "csrc")
HOST_SOURCES = ["mxd_api.cu", "mxd_lockstep.cu", "mxd_hasher.cu", os.path.join("host", "client_host.cpp"),
                os.path.join("host", "stage_copy.cpp")]


def build(sanitize: str = "") -> str:
    """sanitize: "" | "thread" | "address" | "address,undefined" -> a separately named library"""
    tag = sanitize.replace(",", "_")
    out = OUT if not sanitize else OUT.replace(".so", f"_{tag}.so")
    srcs = [os.path.join(CSRC, f) for f in HOST_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in ("kernels.h", "mxd_core.h")] + \
        [os.path.join(MOCK, "mock_kernels.cpp"), os.path.join(MOCK, "include", "cuda_runtime.h"),
         os.path.join(ROOT, "include", "modelx_digest.h"), os.path.join(ROOT, "include", "modelx_client.h"),
         os.path.join(ROOT, "oracle", "sha256_ref.c"), os.path.join(ROOT, "oracle", "modelx_ref.c")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    objdir = os.path.join(os.path.dirname(out), "obj" + ("_" + tag if sanitize else ""))
    os.makedirs(objdir, exist_ok=True)
    san = [f"-fsanitize={sanitize}", "-fno-omit-frame-pointer"] if sanitize else []
    objs = []
    for c in ("sha256_ref.c", "modelx_ref.c"):
        o = os.path.join(objdir, c + ".o")
        subprocess.run(["gcc", "-O2", "-g", "-fPIC", "-std=gnu11", *san, "-c", os.path.join(ROOT, "oracle", c), "-o", o], check=True)
        objs.append(o)
    cxx = ["g++", "-O1", "-g", "-fPIC", "-std=c++17", "-Wall", "-DMXD_MOCK_CUDA", *san, "-I", os.path.join(MOCK, "include")]
    for s in srcs + [os.path.join(MOCK, "mock_kernels.cpp")]:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        subprocess.run([*cxx, "-x", "c++", "-c", s, "-o", o], check=True)
        objs.append(o)
    tmp = out + f".tmp{os.getpid()}"
    subprocess.run(["g++", "-shared", *san, "-o", tmp, *objs, "-lpthread", "-lz"], check=True)
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build())
