"""H2D / D2H rates of cudaMallocHost-pinned vs cudaHostRegister-ed (numpy) memory, flat vs 2-D copies."""
import ctypes, time, sys
import numpy as np
import torch

rt = ctypes.CDLL("libcudart.so.12")
N = 1080 * 1920 * 3
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
pinned = torch.empty(N, dtype=torch.uint8).pin_memory()
arr = np.random.randint(0, 255, N, dtype=np.uint8)
assert rt.cudaHostRegister(ctypes.c_void_p(arr.ctypes.data), ctypes.c_size_t(N), 1) == 0
st = torch.cuda.Stream()
H2D, D2H = 1, 2


def timeit(fn, n=30):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); st.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort(); return ts[n // 2]


def flat(src, nbytes, kind=H2D):
    return lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(src), ctypes.c_size_t(nbytes), kind, ctypes.c_void_p(st.cuda_stream))


def two_d(src, width, rows):
    return lambda: rt.cudaMemcpy2DAsync(ctypes.c_void_p(dev.data_ptr()), ctypes.c_size_t(width), ctypes.c_void_p(src), ctypes.c_size_t(width),
                                        ctypes.c_size_t(width), ctypes.c_size_t(rows), H2D, ctypes.c_void_p(st.cuda_stream))


for name, fn in (("pinned flat 6.2MB", flat(pinned.data_ptr(), N)), ("registered flat 6.2MB", flat(arr.ctypes.data, N)),
                 ("registered 2D 1080x5760", two_d(arr.ctypes.data, 5760, 1080)),
                 ("registered flat 2.07MB", flat(arr.ctypes.data, N // 3)), ("registered flat 1.5MB", flat(arr.ctypes.data, N // 4))):
    t = timeit(fn)
    print("%-28s %7.1f us" % (name, t * 1e6))
