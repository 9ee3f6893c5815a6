"""Build tests/emu/libswim_emu.so: the library's CUDA sources (swim_b200/csrc) compiled with g++ against the SIMT
emulator of tests/emu/include/cuda_runtime.h. TEST INFRASTRUCTURE — exports the same C ABI as libswim_b200.so so the
parity tests can run the device code on the CPU; the product never loads it."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "swim_b200", "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libswim_emu.so")
SOURCES = ["swim_sim.cu", "swim_scalar.cu", "swim_dist.cu", "swim_export.cu", "swim_topology.cpp", "swim_codec.cpp"]
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-fvisibility=hidden", "-fopenmp", "-pthread", "-DSWIM_EMU",
         "-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "include"), "-Wno-unknown-pragmas"]


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, extra=()):
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(ROOT, "include", "swim.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    jobs, objs = [], []
    for src in SOURCES + ["emu_core.cpp"]:
        sp = os.path.join(HERE if src == "emu_core.cpp" else CSRC, src)
        op = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + deps):
            jobs.append(["g++"] + FLAGS + list(extra) + ["-x", "c++", "-c", sp, "-o", op])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("emulator build failed: " + os.path.basename(cmd[-3]))

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(SO, objs):
        run(["g++", "-shared", "-o", SO] + objs + ["-fopenmp", "-pthread", "-ldl"] + list(extra))
    # the in-process NCCL stand-in the staged exchange binds through SWIM_NCCL_LIB in the emulated tests
    fake = os.path.join(os.path.dirname(SO), "libfake_nccl.so")
    src = os.path.join(HERE, "fake_nccl.cpp")
    if force or _stale(fake, [src] + deps):
        run(["g++"] + FLAGS + list(extra) + ["-shared", "-o", fake, src])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
