#!/usr/bin/env python
"""What the host side of the R-shaped calls is made of on this box: H2D / D2H of 100 MB (the
near.obs result of simulation.R: 250 000 rows x 50 columns of fp64) from pageable and pinned
memory, first touch of fresh pages, and host memcpy with 1..8 threads.  Prints one table; the
numbers decide how rfsi_near_obs_f64 / rfsi_forest_predict stage their buffers.
python tools/host_copy_bench.py"""
import ctypes
import threading
import time

import numpy as np
import torch

N = 100_000_000 // 8
dev = torch.device("cuda", 0)
d = torch.empty(N, dtype=torch.float64, device=dev)
d.normal_()
pin = torch.empty(N, dtype=torch.float64).pin_memory()
libc = ctypes.CDLL(None)
libc.memcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]


def t(fn, reps=5):
    fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


def row(name, sec):
    print(f"| {name:62s} | {1e3 * sec:7.2f} ms | {0.1 / sec:6.1f} GB/s |")


print("| 100 MB | time | rate |\n|---|---|---|")
row("D2H into pinned memory", t(lambda: pin.copy_(d, non_blocking=True)))
row("H2D from pinned memory", t(lambda: d.copy_(pin, non_blocking=True)))
touched = torch.empty(N, dtype=torch.float64); touched.zero_()
row("D2H into pageable memory (pages already mapped)", t(lambda: touched.copy_(d)))
row("H2D from pageable memory", t(lambda: d.copy_(touched)))


def fresh_d2h():
    x = torch.empty(N, dtype=torch.float64)      # fresh pages every call, like a new R vector / np.empty
    x.copy_(d)


row("D2H into a FRESH pageable array (allocation + first touch)", t(fresh_d2h))


def first_touch():
    x = np.empty(N)
    x[::512] = 0.0                                 # one write per 4 KB page


row("first touch of 100 MB of fresh pages, one thread", t(first_touch))
src = pin.numpy()
for nt in (1, 2, 4, 8):
    def par_copy(fresh=True):
        dst = np.empty(N) if fresh else par_copy.dst
        step = N // nt
        th = [threading.Thread(target=libc.memcpy, args=(dst.ctypes.data + 8 * i * step, src.ctypes.data + 8 * i * step, 8 * step))
              for i in range(nt)]
        [x.start() for x in th]; [x.join() for x in th]
    par_copy.dst = np.zeros(N)
    row(f"memcpy pinned -> FRESH pageable array, {nt} thread(s)", t(lambda: par_copy(True)))
    row(f"memcpy pinned -> mapped pageable array, {nt} thread(s)", t(lambda: par_copy(False)))
