"""Builds liboctozk.so (and the JNI shim libraries) in-tree with nvcc for sm_100a.

    python -m octopuszk_b200.build [--force]

The build container has no GPU; nvcc cross-compiles.  Outputs land in octopuszk_b200/lib/ (git-ignored, but they
travel to the GPU box with the gpurun snapshot)."""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(ROOT, "build", "obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
]

CU_SOURCES = ["capi.cu", "ntt.cu", "msm.cu", "msm_g1.cu", "msm_g2.cu", "msm_ba_g1.cu", "fixed_base.cu", "stage.cu"]
HEADERS = ["common.h", "consts.cuh", "chains.cuh", "curve.cuh", "fp256.cuh", "ptx_arith.cuh", "msm_impl.cuh", "fixed_impl.cuh", os.path.join(ROOT, "include", "octozk.h")]


# headers only some translation units include (a change must not rebuild the slow msm_g2.cu)
EXTRA_DEPS = {"msm.cu": ["msm_ba_impl.cuh"], "msm_ba_g1.cu": ["msm_ba_impl.cuh"]}


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    sources = [s for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    jobs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not _newer(obj, [src] + hdrs + [os.path.join(CSRC, h) for h in EXTRA_DEPS.get(s, [])]):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + NVCC_FLAGS + extra + ["-c", src, "-o", obj])
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out:
                print(out)
    lib = os.path.join(LIBDIR, "liboctozk.so")
    if force or jobs or not _newer(lib, objs):
        _run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    # JNI shims (plain C++, link against liboctozk.so)
    jni_dir = os.path.join(CSRC, "jni")
    shim_src = os.path.join(jni_dir, "jni_shim.cc")
    if os.path.exists(shim_src):
        for name, macro in (("libAlgebraMSMVariableBaseMSM.so", "OZK_SHIM_VARMSM"),
                            ("libAlgebraMSMFixedBaseMSM.so", "OZK_SHIM_FIXEDMSM"),
                            ("libAlgebraFFTAuxiliary.so", "OZK_SHIM_FFT")):
            out = os.path.join(LIBDIR, name)
            deps = [shim_src, os.path.join(jni_dir, "jni_min.h"), os.path.join(ROOT, "include", "octozk.h")]
            if force or not _newer(out, deps):
                _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-D" + macro, "-I", os.path.join(ROOT, "include"),
                      "-I", jni_dir, shim_src, "-o", out, "-L", LIBDIR, "-loctozk", "-Wl,-rpath,$ORIGIN"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
