#!/usr/bin/env python
"""Hardware run of the scale driver (BASELINE configs[4]): weak scaling 65536 x 8192 nodes per GPU
(65536^2 on 8 GPUs), a few TDGL steps, then the FULL vortex detector (GPU winding pass + host triangulation with
sub-cell positions) and a checkpoint round trip.  torchrun --nproc-per-node N tools/scale_run.py [steps] [Nx] [rows_per_gpu]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svirl_b200.scale import ScaleTD  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
Nx = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
Ny = rows * world
t0 = time.perf_counter()
st = ScaleTD(Nx, Ny, 0.5, 0.5, np.float64, np.inf, 1.0, 0.1, 1.0, 1234, 1.0, device_id=local, distributed=world > 1)
t_build = time.perf_counter() - t0
st.td(0.1, 3)
st.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
st.td(0.1, steps)
st.synchronize()
if world > 1:
    dist.barrier()
t_td = time.perf_counter() - t0
t0 = time.perf_counter()
npos, nneg = st.vortex_count()
t_count = time.perf_counter() - t0
t0 = time.perf_counter()
vx, vy, vv = st.vortices()
t_vort = time.perf_counter() - t0
ok_sorted = bool(np.all(np.diff(np.floor(vy / 0.5)) >= -1))
t = torch.tensor([float(npos), float(nneg), float(vx.size), float((vv > 0).sum()), t_vort, t_count], device="cuda", dtype=torch.float64)
tm = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"grid": [Nx, Ny], "n_gpus": world, "build_s": t_build, "steps": steps, "ms_per_step": 1e3 * t_td / steps,
                      "cell_steps_per_s": Nx * Ny * steps / t_td, "sweeps_psi": int(st.sweeps[0]),
                      "winding_cells": {"positive": int(t[0]), "negative": int(t[1]), "seconds_max": float(tm[5])},
                      "vortices_triangulated": {"count": int(t[2]), "positive": int(t[3]), "seconds_max": float(tm[4]),
                                                "positions_inside_grid": bool(vx.size == 0 or (vx.min() >= -0.5 and vx.max() <= 0.5 * Nx))}}))
st.close()
if world > 1:
    dist.destroy_process_group()
