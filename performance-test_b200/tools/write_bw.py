"""Measured HBM bandwidth of a pure write stream (cudaMemsetAsync of 4 GB) and of a device-to-device
copy (2 GB -> 2 GB), CUDA events, for the write-dominated assembly kernels' roofline.
    python performance-test_b200/tools/write_bw.py"""
import ctypes as C
import json

rt = C.CDLL("libcudart.so")
vp = C.c_void_p


def chk(rc):
    assert rc == 0, rc


n = 4 << 30
a, b = vp(), vp()
chk(rt.cudaMalloc(C.byref(a), C.c_size_t(n)))
chk(rt.cudaMalloc(C.byref(b), C.c_size_t(n)))
e0, e1 = vp(), vp()
chk(rt.cudaEventCreate(C.byref(e0)))
chk(rt.cudaEventCreate(C.byref(e1)))
out = {}
for name, fn, nbytes in (("memset_4GB", lambda: rt.cudaMemsetAsync(a, 0, C.c_size_t(n), None), n),
                         ("copy_2GB_to_2GB", lambda: rt.cudaMemcpyAsync(b, a, C.c_size_t(n // 2), 3, None), n)):
    for _ in range(3):
        chk(fn())
    chk(rt.cudaDeviceSynchronize())
    chk(rt.cudaEventRecord(e0, None))
    for _ in range(10):
        chk(fn())
    chk(rt.cudaEventRecord(e1, None))
    chk(rt.cudaEventSynchronize(e1))
    ms = C.c_float()
    chk(rt.cudaEventElapsedTime(C.byref(ms), e0, e1))
    out[name] = {"ms": ms.value / 10, "GBps": nbytes / (ms.value / 10 * 1e-3) / 1e9}
print(json.dumps(out))
