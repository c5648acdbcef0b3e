"""Build libariadne_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libariadne_b200.so")
SOURCES = ["context.cu", "blas1.cu", "stencil.cu", "krylov.cu", "newton.cu", "user.cu", "precond.cu", "sweep.cu"]
HEADERS = ["ak_internal.h", "common.cuh", "sweep.h", os.path.join("..", "..", "include", "ariadne_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
    "-fmad=true",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libariadne_b200.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that nvcc should not pick up as host compiler
    objs = []
    log = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if pr.returncode != 0:
            failed = True
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see newtonkrylov.jl_b200/build/build.log")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-ccbin", "/usr/bin/g++", "-Xlinker", "-ldl", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libariadne_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
