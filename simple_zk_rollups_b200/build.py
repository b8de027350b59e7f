"""In-tree build of libzkr.so (CUDA, sm_100a only) and nothing else.

    python -m simple_zk_rollups_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A/B of compile-time variants on one GPU box: ZKR_BUILD_VARIANT=name ZKR_BUILD_DEFINES="-DX=1 ..." builds libzkr_<name>.so
# beside the product library (own object directory); ZKR_LIB=<path> makes _lib.py load it.  Never part of build().
_VARIANT = os.environ.get("ZKR_BUILD_VARIANT", "")
OBJ = os.path.join(CSRC, "build" + ("_" + _VARIANT if _VARIANT else ""))
LIB = os.path.join(HERE, "libzkr%s.so" % ("_" + _VARIANT if _VARIANT else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v" if os.environ.get("ZKR_PTXAS_V") else "-O3",
] + (os.environ.get("ZKR_BUILD_DEFINES", "").split() if _VARIANT else [])


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    nvcc = _nvcc()
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def run(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        p = subprocess.run(cmd, capture_output=True, text=True)
        return job, p

    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for (s, o), p in ex.map(run, jobs):
            if verbose or p.returncode != 0:
                sys.stderr.write("[nvcc] %s\n%s%s" % (os.path.basename(s), p.stdout, p.stderr))
            if p.returncode != 0:
                failed = True
    if failed:
        raise RuntimeError("libzkr build failed")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("libzkr link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
