"""Builds libpcm_b200.so (the C-ABI library of include/pcm_b200.h) in-tree with nvcc
for sm_100a.  nvcc cross-compiles without a GPU, so this runs in the build container;
the resulting .so travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("PCM_LIB_PATH") or os.path.join(PKG, "libpcm_b200.so")
SOURCES = ["pcm_api.cu", "pcm_host_simd.cpp", "pcm_felzenszwalb.cpp"]
DEPS = ["pcm_api.cu", "pcm_kernels.cuh", "pcm_device.cuh", "pcm_host.h", "pcm_host_simd.cpp", "pcm_felzenszwalb.cpp", "pcm_quickshift.cuh", "pcm_forest_fit.cuh", "pcm_prior.cuh", os.path.join("..", "..", "include", "pcm_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-Xcompiler", "-ffp-contract=off", "-shared",
    "-Xptxas", "-v",
]


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile if sources are newer than the library.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("PCM_BUILD_DEFS", "").split()      # e.g. "-DPCM_NTHREADS=512" (tuning experiments)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (%d): %s" % (res.returncode, " ".join(cmd)))
    if LIB.endswith("libpcm_b200.so"):       # registers / spills / shared memory per kernel (ptxas -v), without timings
        with open(os.path.join(PKG, "csrc", "ptxas_info.txt"), "w") as f:
            f.write("".join(l for l in res.stderr.splitlines(True) if "Compile time" not in l))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
