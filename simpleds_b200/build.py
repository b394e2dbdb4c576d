"""Build libsds.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsds.so")
SOURCES = [
    "sds_api.cu", "sds_fft_generic.cu", "sds_fft_pow2.cu", "sds_fft_tma.cu", "sds_fft_bluestein.cu", "sds_elementwise.cu",
    "sds_cross.cu", "sds_thermal.cu", "sds_engine.cu", "sds_engine_f32.cu", "sds_engine_blue.cu", "sds_engine_uv2.cu",
    "sds_estimators.cu", "sds_cosmo.cu", "sds_ingest.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "sds.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    """Compile every CUDA source into simpleds_b200/libsds.so; returns its path."""
    if not force and not needs_build():
        return LIB
    flags = list(NVCC_FLAGS) + os.environ.get("SDS_NVCC_FLAGS", "").split()
    objs, procs = [], []
    for src in SOURCES:  # one nvcc per translation unit, all at once (the engines take ~20 s each)
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    cmd = [_nvcc(), *flags, "-shared", "-o", LIB, *objs]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
