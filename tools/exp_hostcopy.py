"""Host memcpy rates on the GPU box: pageable<->pageable, pageable->pinned, pinned->pageable."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import portablert_b200 as prt
from portablert_b200.backend import pinned_empty

b = prt.CUDABackend(device=0)
b.init()
n = 64 << 20
page_a = np.ones(n, np.uint8); page_b = np.ones(n, np.uint8)
pin_a = pinned_empty((n,), np.uint8); pin_a[:] = 1
pin_b = pinned_empty((n,), np.uint8); pin_b[:] = 1


def rate(dst, src, label):
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); dst[:] = src; ts.append(time.perf_counter() - t0)
    print("%-24s %.2f ms  %.1f GB/s" % (label, np.median(ts) * 1e3, n / np.median(ts) / 1e9))


rate(page_b, page_a, "pageable -> pageable")
rate(pin_a, page_a, "pageable -> pinned")
rate(page_b, pin_a, "pinned -> pageable")
rate(pin_b, pin_a, "pinned -> pinned")
small = 6 << 20
for _ in range(2):
    t0 = time.perf_counter()
    for k in range(0, n, small):
        pin_a[k:k + small] = page_a[k:k + small]
    print("pageable -> pinned in 6 MB pieces: %.1f GB/s" % (n / (time.perf_counter() - t0) / 1e9))
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
