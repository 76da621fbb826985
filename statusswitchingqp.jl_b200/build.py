"""In-tree build of libssqp_b200.so (sm_100a).  `python build.py` or ssqp_b200.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SSQP_TIMELINE=1: developer build with the per-section timeline of a Phase-2 trip, kept apart from the product library
# (libssqp_b200_tl.so; load it with SSQP_LIB=<path>)
_TL = bool(os.environ.get("SSQP_TIMELINE"))
OBJ = os.path.join(HERE, "build_tl" if _TL else "build")
LIB = os.path.join(HERE, "libssqp_b200_tl.so" if _TL else "libssqp_b200.so")
NTS = (128, 256, 512)           # CTA widths of the solve kernel; each also in a 256-bit-loads-only flavour (one TU each)
EXTRA_DEFS = ["-DSSQP_TIMELINE"] if os.environ.get("SSQP_TIMELINE") else []     # developer build: per-section timeline
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


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
    hdrs = [os.path.join(CSRC, "ssqp_kernel.cuh"), os.path.join(CSRC, "ssqp_helpers.cuh"),
            os.path.join(HERE, "..", "include", "ssqp_b200.h")]
    jobs = []
    objs = []
    o = os.path.join(OBJ, "ssqp_capi.o")
    objs.append(o)
    src = os.path.join(CSRC, "ssqp_capi.cu")
    if force or _newer(o, [src] + hdrs):
        jobs.append([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", o])
    src = os.path.join(CSRC, "ssqp_inst.cu")
    for nt in NTS:
        for vw4 in (0, 1):
            o = os.path.join(OBJ, "ssqp_inst_%d_%d.o" % (nt, vw4))
            objs.append(o)
            # developer shortcut (perf iterations on the benchmark flavour): SSQP_BUILD_ONLY=512_1 recompiles that TU only
            # and links the other, possibly stale, objects — always finish with a full build
            only = os.environ.get("SSQP_BUILD_ONLY")
            if only and only != "%d_%d" % (nt, vw4) and os.path.exists(o):
                continue
            if force or _newer(o, [src] + hdrs):
                jobs.append([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + EXTRA_DEFS +
                            ["-DSSQP_NT=%d" % nt] + (["-DSSQP_ONLY_VW4"] if vw4 else []) + ["-c", src, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc:
                    sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
                if rc:
                    raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if jobs or force or _newer(LIB, objs):
        cmd = [_nvcc()] + ["-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        cmd, rc, out = run(cmd)
        if rc:
            sys.stderr.write(out)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
