#!/usr/bin/env python
"""Diagnostic (1 GPU or torchrun): cfg2 weak-scaling workload timed in chunks of 5 steps after a 5-step warm-up, to see
how the step time settles (sweep-count prediction, clocks, rank skew).  Prints per chunk: ms/step (max over ranks),
sweeps, launches, replays, gate hits/misses."""
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

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = bench.workload("cfg2")
wl = dict(wl, Ny=wl["Ny"] * world)
gl = bench.make_solver(wl, device_id=local, slab="auto" if world > 1 else None)
par, td = gl.par, None
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    par.set_option(k, int(v))
gl.solve.td(dt=0.1, Nt=5)
td = gl.solve._td
for chunk in range(10):
    par.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s0, l0, r0, h0, m0 = td.sweeps_order_parameter, par.stat("launches"), par.stat("replays"), par.stat("spec_hit"), par.stat("spec_miss")
    _lib.call("svl_event_record", par.ctx, 0)
    gl.solve.td(dt=0.1, Nt=5)
    _lib.call("svl_event_record", par.ctx, 1)
    ms = C.c_double()
    _lib.call("svl_event_elapsed_ms", par.ctx, 0, 1, C.byref(ms))
    t = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("chunk %d: %.4f ms/step  sweeps %d launches %d replays %d gate %d/%d" % (
            chunk, float(t.item()) / 5, td.sweeps_order_parameter - s0, par.stat("launches") - l0,
            par.stat("replays") - r0, par.stat("spec_hit") - h0, par.stat("spec_miss") - m0), flush=True)
if world > 1:
    dist.destroy_process_group()
