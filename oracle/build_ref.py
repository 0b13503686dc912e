"""Recipe for oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

The PC masker path of the reference is Python; the only C/C++ of the reference that touches a
stage next to the path is the vendored Felzenszwalb-Huttenlocher segmentation under
/root/reference/prim/src/FelzenSegment (header-only, no build system needed).  It is compiled
here FROM WHERE IT LIES (the sources are not copied) together with oracle/felzen_ref_shim.cpp:

    g++ -O2 -shared -fPIC -I/root/reference/prim/src/FelzenSegment oracle/felzen_ref_shim.cpp \
        -o oracle/_ref/libfelzen_ref.so

oracle/_ref/ is git-ignored and travels to the GPU box with the snapshot.  When /root/reference is
absent (GPU box) an already built library is used as it is; without either, build() returns None
and the tests that need it skip.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PCM_REFERENCE_ROOT", "/root/reference")
FELZEN_INC = os.path.join(REF_SRC, "prim", "src", "FelzenSegment")
SHIM = os.path.join(HERE, "felzen_ref_shim.cpp")
LIB = os.path.join(HERE, "_ref", "libfelzen_ref.so")


def build(force=False):
    """Returns the path of libfelzen_ref.so, or None when it neither exists nor can be built."""
    have_src = os.path.isfile(os.path.join(FELZEN_INC, "segment-graph.h"))
    if not have_src:
        return LIB if os.path.isfile(LIB) else None
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SHIM):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++14", "-w", "-ffp-contract=off", "-shared", "-fPIC", "-I", FELZEN_INC, SHIM, "-o", LIB]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
