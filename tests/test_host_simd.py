"""csrc/pcm_host_simd.cpp (strided gather / scatter of the interleaved mask channel) against
numpy slicing, compiled stand-alone with g++ -- CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np

from helpers import PKG

DRIVER = r'''
#include <cstdint>
namespace pcm {
void gather_strided(const uint8_t* src, int64_t stride, uint8_t* dst, int n);
void scatter_strided(const uint8_t* src, uint8_t* dst, int64_t stride, int n);
}
extern "C" void t_gather(const uint8_t* s, int64_t st, uint8_t* d, int n) { pcm::gather_strided(s, st, d, n); }
extern "C" void t_scatter(const uint8_t* s, uint8_t* d, int64_t st, int n) { pcm::scatter_strided(s, d, st, n); }
'''


def test_gather_scatter_strided(tmp_path):
    drv = tmp_path / "drv.cpp"
    drv.write_text(DRIVER)
    so = tmp_path / "libsimd_test.so"
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", str(so), str(drv),
                    os.path.join(PKG, "csrc", "pcm_host_simd.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.t_gather.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]      # addresses are 64-bit
    lib.t_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.t_gather.restype = lib.t_scatter.restype = None
    rng = np.random.default_rng(5)
    for stride in (1, 3, 4):
        for n in (0, 1, 15, 16, 17, 18, 33, 48, 100, 1920):
            # exactly-sized buffers: an over-read / over-write of even one byte lands in the guard
            src = rng.integers(0, 256, max((n - 1) * stride + 1, 1) + 64, dtype=np.uint8)
            src[(n - 1) * stride + 1 if n else 0:] = 0xAB
            dst = np.full(n + 64, 0xCD, np.uint8)
            lib.t_gather(src.ctypes.data, stride, dst.ctypes.data, n)
            assert np.array_equal(dst[:n], src[:n * stride:stride][:n]), (stride, n)
            assert (dst[n:] == 0xCD).all(), "gather wrote past the end"
            plane = rng.integers(0, 256, n + 8, dtype=np.uint8)
            img = rng.integers(0, 256, max((n - 1) * stride + 1, 1) + 64, dtype=np.uint8)
            want = img.copy()
            if n:
                want[:(n - 1) * stride + 1:stride] = plane[:n]
            lib.t_scatter(plane.ctypes.data, img.ctypes.data, stride, n)
            assert np.array_equal(img, want), (stride, n)
