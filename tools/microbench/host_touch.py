"""Host first-touch cost of the result array of a 2048x518x518 video (page faults), plain vs huge-page advice vs threads."""
import mmap
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

n = 2048 * 518 * 518 * 4


def touch(a):
    t = time.perf_counter()
    a[::1024].fill(0)
    return time.perf_counter() - t


a = np.empty(n // 4, np.float32)
print("np.empty first touch, 1 thread: %.3f s" % touch(a))
print("second touch: %.3f s" % touch(a))
m = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
m.madvise(mmap.MADV_HUGEPAGE)
b = np.frombuffer(m, dtype=np.float32)
print("mmap + MADV_HUGEPAGE first touch: %.3f s" % touch(b))
c = np.empty(n // 4, np.float32)
T = 6
cuts = np.linspace(0, c.size, T + 1).astype(np.int64)
pool = ThreadPoolExecutor(T)
t = time.perf_counter()
list(pool.map(lambda i: touch(c[cuts[i]:cuts[i + 1]]), range(T)))
print("np.empty first touch, %d threads: %.3f s" % (T, time.perf_counter() - t))
src = np.ones((32, 518, 518), np.float32)
d = np.empty((2048, 518, 518), np.float32)
t = time.perf_counter()
for k in range(32):
    np.copyto(d[k * 32:(k + 1) * 32], src)
print("copy into untouched: %.1f GB/s" % (32 * src.nbytes / (time.perf_counter() - t) / 1e9))
t = time.perf_counter()
for k in range(32):
    np.copyto(d[k * 32:(k + 1) * 32], src)
print("copy into touched: %.1f GB/s" % (32 * src.nbytes / (time.perf_counter() - t) / 1e9))
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
