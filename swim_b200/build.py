"""Build libswim_b200.so (CUDA sm_100a kernels + the C ABI of include/swim.h) in-tree with nvcc.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repo snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
SO = os.path.join(HERE, "libswim_b200.so")
INCLUDE = os.path.join(HERE, "..", "include")

SOURCES = ["swim_sim.cu", "swim_scalar.cu", "swim_dist.cu", "swim_export.cu", "swim_topology.cpp", "swim_codec.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-fopenmp,-Wall", "-I", INCLUDE]


def _deps():
    out = [os.path.join(INCLUDE, "swim.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".h", ".cuh")):
            out.append(os.path.join(CSRC, f))
    return out


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    global NVCC_FLAGS
    wpb = os.environ.get("SWIM_WPB")  # experiment: CTA size of the per-round kernels (8, 16 or 32 warps)
    if wpb and not any(f.startswith("-DSWIM_WARPS_PER_BLOCK") for f in NVCC_FLAGS):
        NVCC_FLAGS = NVCC_FLAGS + ["-DSWIM_WARPS_PER_BLOCK=" + wpb]
        force = True
    os.makedirs(OBJ, exist_ok=True)
    deps = _deps()
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        op = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + deps):
            cmd = [nvcc] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cu") else []) + ["-c", sp, "-o", op]
            if verbose and src.endswith(".cu"):
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(SO, objs):
        run([nvcc, "-shared", "-o", SO] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
