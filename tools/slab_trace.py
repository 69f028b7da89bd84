#!/usr/bin/env python
"""Diagnostic (torchrun, >= 2 GPUs): per-launch timeline of the slab psi tile kernels.
    torchrun --nproc-per-node 2 tools/slab_trace.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svirl_b200 import _lib  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = bench.workload("cfg2")
wl = dict(wl, Ny=wl["Ny"] * world)
gl = bench.make_solver(wl, device_id=local, slab="auto")
gl.solve.td(dt=0.1, Nt=30)
NL = 64
gl.par.set_option("trace", NL)
gl.solve.td(dt=0.1, Nt=8)
buf = (C.c_ulonglong * (4 * NL))()
n = C.c_int()
_lib.call("svl_debug_trace", gl.par.ctx, buf, NL, C.byref(n))
a = np.array(buf[:4 * n.value], dtype=np.uint64).reshape(-1, 4).astype(np.int64)
t0 = a[0, 0]
for r in range(world):
    dist.barrier()
    if r == rank:
        print("rank %d: launch  start_us  dur_us  gap_us  maxwait_us  waitend-start_us" % rank)
        for k in range(n.value):
            gap = (a[k, 0] - a[k - 1, 1]) / 1e3 if k else 0.0
            print("   %3d %10.1f %8.1f %8.1f %8.1f %8.1f" % (k, (a[k, 0] - t0) / 1e3, (a[k, 1] - a[k, 0]) / 1e3, gap,
                                                      a[k, 2] / 1e3, (a[k, 3] - a[k, 0]) / 1e3 if a[k, 3] else 0.0))
        sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
