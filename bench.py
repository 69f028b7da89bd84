#!/usr/bin/env python
"""Benchmark of the TDGL hot path (BASELINE.json metric: TDGL cell-steps/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]

Workload at N=1 (BASELINE.json configs[1], SURVEY.md section 8d "cfg2"): infinite-kappa TDGL
(psi only), 2048^2 nodes, fp32, dx=dy=0.5, dt=0.1, H=0.1, eps=1, seed 1234, material tiling =
square lattice (period 16) of circular holes of radius 2.  A "step" is one TDGL time step = one
complete psi Jacobi solve (about 20 sweeps) over the whole grid.

One JSON line on stdout (rank 0).  `value` = nodes * steps / device time with the fields resident
in HBM; `e2e` = the same through the C ABI with HOST buffers (pinned psi up, one step, psi down,
every step); `roofline` = algorithmic bytes of the psi sweep kernel / its launch time vs the
measured HBM copy peak; `cpu_baseline` = the reference's own kernel source compiled for the host
cores (oracle/_ref, kind "reference") or the NumPy port, on a bounded sample.  Beside the contract keys:

  parity        GPU psi after the cpu_baseline leg's steps vs the reference-kernel psi of that leg (same seeded state)
  gpu_baseline  "B-ref": the UNMODIFIED reference package with its own CUDA kernels on this GPU (baseline/bref.py)
  also          the other BASELINE configs on this GPU: cfg3 (8192^2 fp64 kappa=2 TDGL), cfg4s (8192^2 fp64 CG
                iterations, the second half of BASELINE's metric), cfg1 (README 129^2), each with its roofline and B-ref
  strong        BASELINE configs[4]: 32768^2 fp64 kappa=inf split over the N GPUs of this run (strong scaling)
  slab_bitwise  N > 1: a 300 x 401 run on the N slabs equals the single-GPU run bit for bit (checked before timing)

--impl reference times that CPU arm alone (the reference has no CPU path and pyCUDA cannot be
installed offline).  With --gpus N > 1 (launched under torchrun, one rank per GPU)
the grid grows to Nx x (Ny*N) and is decomposed into N row slabs (weak scaling: every GPU keeps the
N=1 workload); halo rows move by direct peer stores over NVLink and the only collective is the MAX
of the per-sweep residuals.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tdgl_cell_steps_per_s"
UNIT = "cell-steps/s"


# ------------------------------------------------------------------------------------ workload
def workload(name):
    if name == "cfg2":
        return dict(name="cfg2: TDGL kappa=inf 2048^2 fp32, hole-lattice tiling", Nx=2048, Ny=2048, dtype=np.float32,
                    kappa=np.inf, sigma=1.0, H=0.1, tiling=True, eps_field=False)
    if name == "cfg3":
        return dict(name="cfg3: TDGL kappa=2 8192^2 fp64, disordered eps", Nx=8192, Ny=8192, dtype=np.float64,
                    kappa=2.0, sigma=10.0, H=0.1, tiling=False, eps_field=True)
    if name == "cfg4":
        return dict(name="cfg4: CG minimiser 16384^2 fp64 kappa=2 (20 TD steps, then CG iterations)", Nx=16384, Ny=16384,
                    dtype=np.float64, kappa=2.0, sigma=10.0, H=0.1, tiling=False, eps_field=False, cg=True)
    if name == "cfg4s":
        return dict(name="cfg4s: CG minimiser 8192^2 fp64 kappa=2", Nx=8192, Ny=8192, dtype=np.float64, kappa=2.0,
                    sigma=10.0, H=0.1, tiling=False, eps_field=False, cg=True)
    if name == "cfg4inf":
        return dict(name="cfg4inf: CG minimiser 8192^2 fp32 kappa=inf tiled", Nx=8192, Ny=8192, dtype=np.float32,
                    kappa=np.inf, sigma=1.0, H=0.1, tiling=True, eps_field=False, cg=True)
    if name == "cfg5":        # BASELINE configs[4], strong scaling: the grid is fixed, rows are split over the GPUs
        return dict(name="cfg5: TDGL kappa=inf 32768^2 fp64, strong scaling, slab-local fields + vortex count", Nx=32768,
                    Ny=32768, dtype=np.float64, kappa=np.inf, sigma=1.0, H=0.1, tiling=False, eps_field=False, scale="strong")
    if name == "cfg5k2":
        return dict(name="cfg5k2: TDGL kappa=2 32768^2 fp64, strong scaling", Nx=32768, Ny=32768, dtype=np.float64,
                    kappa=2.0, sigma=10.0, H=0.1, tiling=False, eps_field=False, scale="strong")
    if name == "cfg5w":       # weak scaling: 65536 x 8192 nodes per GPU
        return dict(name="cfg5w: TDGL kappa=inf 65536 x 8192 per GPU fp64, weak scaling", Nx=65536, Ny=8192,
                    dtype=np.float64, kappa=np.inf, sigma=1.0, H=0.1, tiling=False, eps_field=False, scale="weak")
    if name == "cfg5mini":    # the same driver at a size that runs anywhere
        return dict(name="cfg5mini: TDGL kappa=inf 4096^2 fp64 through the scale driver", Nx=4096, Ny=4096,
                    dtype=np.float64, kappa=np.inf, sigma=1.0, H=0.1, tiling=False, eps_field=False, scale="strong")
    if name == "cfg1":
        return dict(name="cfg1: README 129^2 fp64 kappa=5", Nx=129, Ny=129, dtype=np.float64, kappa=5.0, sigma=200.0,
                    H=0.1, tiling=False, eps_field=False)
    if name == "small":
        return dict(name="small: 512^2 fp32 kappa=inf tiled", Nx=512, Ny=512, dtype=np.float32, kappa=np.inf,
                    sigma=1.0, H=0.1, tiling=True, eps_field=False)
    raise SystemExit("unknown workload " + name)


def hole_tiling(Nx, Ny, dx=0.5, dy=0.5):
    x = (np.arange(Nx - 1) + 0.5) * dx
    y = (np.arange(Ny - 1) + 0.5) * dy
    fx = (np.mod(x, 16.0) - 8.0) ** 2
    fy = (np.mod(y, 16.0) - 8.0) ** 2
    return ~((fx[:, None] + fy[None, :]) < 4.0)


def make_solver(wl, device_id=0, slab=None):
    from svirl_b200 import GLSolver
    kw = dict(Nx=wl["Nx"], Ny=wl["Ny"], dx=0.5, dy=0.5, dtype=wl["dtype"], gl_parameter=wl["kappa"],
              normal_conductivity=wl["sigma"], homogeneous_external_field=wl["H"], random_seed=wl.get("seed", 1234),
              device_id=device_id, slab=slab)
    if wl["tiling"]:
        kw["material_tiling"] = hole_tiling(wl["Nx"], wl["Ny"])
    if wl["eps_field"]:
        kw["linear_coefficient"] = (0.7 + 0.3 * np.random.RandomState(4321).rand(wl["Nx"], wl["Ny"])).astype(wl["dtype"])
    return GLSolver(**kw)


def bytes_per_node_sweep(wl):
    """psi node-sweep 8R+1(+R); A node-sweep 9R+1."""
    R = np.dtype(wl["dtype"]).itemsize
    psi = 8 * R + 1 + (R if wl["eps_field"] else 0)
    return psi, 9 * R + 1


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock and throttle reasons sampled while the timed region runs: NVML in-process (a sample
    every ~2 ms, so even a 20 ms region is covered), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None
        self.armed = False
        self.wake, self.wake2 = threading.Event(), threading.Event()      # start() / stop()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap the torch index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        R = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        if not hasattr(self, "_mx"):
            try:
                self._mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                self._mx = None
        try:
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append([sm, self._mx, [k for k, v in R.items() if bits & v]])
        except Exception:
            pass

    def _run_nvml(self):
        # An NVML query holds a driver lock for ~1 ms, and a launch-bound timed region feels it (measured: one query per
        # millisecond turned 0.41 ms/step into 1.7).  So: first query 20 ms into the region, then one every 25 ms; a region
        # shorter than that gets its one sample from stop(), taken the moment the last timed step has finished.
        # The thread sleeps on an Event (no periodic wake-ups while the ranks are being timed: a box has ~2 host threads
        # per GPU rank).
        self.wake.wait()
        if self.stop_flag:
            return
        if self.wake2.wait(0.020):
            return
        while not self.stop_flag:
            self._sample_nvml()
            if self.wake2.wait(0.025):
                return

    def _run_smi(self):
        while not self.stop_flag:
            if not self.armed:
                time.sleep(0.001)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                s = [x.strip() for x in out.strip().split(",")]
                names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                self.samples.append([float(s[0]), float(s[1]),
                                     [k for k, v in zip(names, s[3:7]) if v.lower().startswith("active")]])
            except Exception:
                pass
            time.sleep(0.05)

    def prepare(self):
        """Everything slow (NVML first calls, thread start) BEFORE the barrier that precedes the timed region: done after
        it, it would skew the ranks' start times by milliseconds, which the first timed step then absorbs."""
        if self.th is not None:
            return
        if self.nvml:
            self._sample_nvml()
            self.samples = []
        self.th = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self.th.start()

    def start(self):
        self.prepare()
        self.armed = True
        self.wake.set()

    def stop(self):
        """Call it the moment the timed steps have finished (before the closing barrier): a region of a few ms gets its
        clock sample here -- SM clocks do not drop within microseconds of the last kernel -- see _run_nvml."""
        self.armed = False
        self.stop_flag = True
        self.wake.set()
        self.wake2.set()
        if self.nvml and not self.samples:
            self._sample_nvml()
        if self.th:
            self.th.join(timeout=10)
        sm = [s[0] for s in self.samples]
        mx = next((s[1] for s in self.samples if s[1]), None)
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def aligned_start(world):
    """After the barrier: all ranks of the box leave at the same instant of the shared monotonic clock (rank 0 names it).
    A NCCL barrier releases the ranks up to a few hundred microseconds apart; on row slabs every rank waits for its
    neighbours, so that skew would be charged to the first timed step (it is 5 % of a 20-step cfg2 region).  The instant
    is agreed on with two small collectives; a rank that hears of it too late makes everybody try again with more margin."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    clock = lambda: time.clock_gettime(time.CLOCK_MONOTONIC)       # noqa: E731
    t = torch.zeros(1, dtype=torch.float64, device="cuda")
    for margin in (0.01, 0.05, 0.25):
        if dist.get_rank() == 0:
            t[0] = clock() + margin
        dist.broadcast(t, 0)
        target = float(t.item())
        ok = torch.tensor([1.0 if clock() < target - 0.004 else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0.0 and clock() < target:
            while clock() < target:
                pass
            return


# ------------------------------------------------------------------------------------ CPU reference arm
def _ref_lib(wl):
    """oracle/_ref: the reference's kernel sources compiled for host cores (built by oracle/build_ref.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    rvl = 5 if np.isinf(wl["kappa"]) else 17
    path = os.path.join(build_ref.OUT, build_ref.prebuilt_name(wl["dtype"], wl["Nx"], wl["Ny"], 0.5, 0.5, rvl))
    if not os.path.exists(path) and os.path.isdir(build_ref.REF_CUDA):
        path = build_ref.prebuild(wl["dtype"], wl["Nx"], wl["Ny"], 0.5, 0.5, rvl)
    return path if os.path.exists(path) else None


def cpu_reference_steps(wl, nsteps, warmup=0):
    """Time `nsteps` TDGL steps of the psi equation on the host cores.  Uses the reference kernel
    (oracle/_ref .so) driven with the reference's launch pattern (svirl/solvers/td.py:157-218) when
    it was built, else the NumPy port.  Returns (cell_steps_per_s, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import glnumpy as O
    Nx, Ny, dt_ = wl["Nx"], wl["Ny"], wl["dtype"]
    band = None
    if (warmup + nsteps) > 60 and Ny > 512:
        # bounded sample: a 512-row band of the same workload per step (same cost per node), so that
        # a long --steps run still ends within a few minutes
        band = 512
        Ny = band
        wl = dict(wl, Ny=band)
    g = O.Grid(Nx, Ny, 0.5, 0.5, dt_)
    psi = O.initial_psi(g, 1.0, 1234)
    a, b = O.initial_A(g, wl["H"])
    mt = hole_tiling(Nx, Ny) if wl["tiling"] else None
    eps = (0.7 + 0.3 * np.random.RandomState(4321).rand(Nx, Ny)).astype(dt_) if wl["eps_field"] else 1.0
    N = Nx * Ny
    so = _ref_lib(wl)
    sweeps = 0
    if so is None:
        t0 = time.perf_counter()
        for s in range(warmup + nsteps):
            if s == warmup:
                t0 = time.perf_counter()
                sweeps = 0
            psi, n = O.td_psi_solve(g, 0.1, eps, mt, a, b, psi)
            sweeps += n
        el = time.perf_counter() - t0
        return N * nsteps / el, dict(kind="port", cores=1, sweeps=sweeps, seconds=el, band=band)
    lib = C.CDLL(so)
    try:       # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(os.cpu_count() or 1))
    except OSError:
        pass
    fn = lib.simt_launch_iterate_order_parameter_jacobi_step
    real = C.c_float if dt_ is np.float32 else C.c_double
    cplx = np.complex64 if dt_ is np.float32 else np.complex128
    fn.argtypes = [C.c_int] * 4 + [real, real, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    real, C.c_uint32, C.c_uint32, real, C.c_void_p]
    fn.restype = None
    flat = lambda x: np.ascontiguousarray(np.reshape(x.T, x.size))
    cur = flat(psi).astype(cplx)
    nxt = np.zeros_like(cur)
    rhs = np.zeros_like(cur)
    ab = np.concatenate([flat(a), flat(b)]).astype(dt_)
    mtf = flat(mt).astype(np.bool_) if mt is not None else None
    epsf = flat(eps).astype(dt_) if wl["eps_field"] else None
    r2 = np.zeros(1, dtype=np.int32)
    grid = (N + 127) // 128
    P = lambda x: x.ctypes.data if x is not None else None
    t0 = time.perf_counter()
    for s in range(warmup + nsteps):
        if s == warmup:
            t0 = time.perf_counter()
            sweeps = 0
        rhs[:] = cur
        for j in range(1024):
            r2[0] = 0
            fn(grid, 1, 128, 0, 0.1, 0.0 if wl["eps_field"] else 1.0, P(epsf), P(mtf), P(ab), P(rhs), P(cur), P(nxt),
               0.0, j, 1, 1e-6, P(r2))
            cur, nxt = nxt, cur
            sweeps += 1
            if 1.0e-4 * float(r2[0]) < 1.0:
                break
    el = time.perf_counter() - t0
    return N * nsteps / el, dict(kind="reference", cores=os.cpu_count(), sweeps=sweeps, seconds=el, band=band,
                                 psi=cur if (warmup == 0 and band is None) else None)


# ------------------------------------------------------------------------------------ helpers
def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f).get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback 6650 (B200_PROFILING.md)"


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json names
    the kernel instantiation and the command each number was captured on); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def solver_kwargs(wl):
    kw = dict(Nx=wl["Nx"], Ny=wl["Ny"], dx=0.5, dy=0.5, dtype=wl["dtype"], gl_parameter=wl["kappa"],
              normal_conductivity=wl["sigma"], homogeneous_external_field=wl["H"], random_seed=1234)
    if wl["tiling"]:
        kw["material_tiling"] = hole_tiling(wl["Nx"], wl["Ny"])
    if wl["eps_field"]:
        kw["linear_coefficient"] = (0.7 + 0.3 * np.random.RandomState(4321).rand(wl["Nx"], wl["Ny"])).astype(wl["dtype"])
    return kw


def bref_td(wl, steps, warmup=3):
    """B-ref: the reference's own solver + kernels on this GPU, same seeded workload (TDGL steps)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import bref
    if not bref.available():
        return {"unavailable": "baseline/_ref not installed (baseline/install_ref.py)"}
    gl = bref.make_solver(**solver_kwargs(wl))
    el, sp, sa = bref.time_td(gl, steps, warmup)
    N = wl["Nx"] * wl["Ny"]
    out = {"kind": "reference-kernels-sm100", "value": N * steps / el, "unit": UNIT, "ms_per_step": 1e3 * el / steps,
           "steps": steps, "warmup": warmup, "sweeps_psi": int(sp), "sweeps_A": int(sa),
           "how": "unmodified reference package (baseline/_ref) on baseline/gpu_pycuda: block 128, one thread per node, per "
                  "sweep fill + launch + blocking 4-byte read-back (svirl/solvers/td.py:164-202, 274-311); wall clock"}
    del gl
    return out


def bref_cg(wl, iters, td_steps=20):
    """B-ref CG iterations (svirl/solvers/cg.py:477-544 as written, SciPy line search) after `td_steps` TDGL steps."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import bref
    if not bref.available():
        return {"unavailable": "baseline/_ref not installed (baseline/install_ref.py)"}
    gl = bref.make_solver(**solver_kwargs(wl))
    gl.solve.td(dt=0.1, Nt=td_steps)
    with np.errstate(all="ignore"):
        el, E = bref.time_cg(gl, iters, warmup=0)
    fin = int(np.sum(np.isfinite(np.array(E, dtype=np.float64))))
    out = {"kind": "reference-kernels-sm100", "value": iters / el, "unit": "iterations/s", "ms_per_iter": 1e3 * el / iters,
           "iterations": iters, "energies": E, "finite_iterations": fin,
           "how": "unmodified reference package on baseline/gpu_pycuda; the first `iterations` CG iterations after the TDGL "
                  "steps (no warm-up: its SciPy BFGS line search runs away after 3 iterations at this size, "
                  "profiles/r02_cfg4_adjudicate_8192.json)"}
    del gl
    return out


# ------------------------------------------------------------------------------------ CG mode
def measure_cg(wl, gl, par, steps, warmup, line_search=None, td_steps=20):
    """Hot path 2: K timed CG iterations through gl.solve.cg() (metric: CG iterations / s) after `td_steps` TDGL steps
    (skipped when td_steps = 0) and W warm-up iterations.  The host line search is inside the timed region because it
    is part of an iteration.  line_search: None / "reference" = SciPy BFGS / polyroots exactly as the reference calls
    them (redone on the normalised polynomial only when BFGS runs away; rescues are counted), "native" = the
    library's own search (opt-in)."""
    from svirl_b200 import _lib
    N = wl["Nx"] * wl["Ny"]
    if td_steps:
        gl.solve.td(dt=0.1, Nt=td_steps)
    gl.cfg.cg_line_search = line_search or "reference"
    gl.cfg.convergence_rtol = 0.0
    gl.solve._init_cg()
    gl.solve._cg._CG__convergence_rtol = -1.0          # never stop early: time exactly K iterations
    gl.solve._cg.line_search_rescues = 0
    gl.solve.cg(n_iter=max(warmup, 3))
    par.synchronize()
    l0 = par.stat("launches")
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    _lib.call("svl_event_record", par.ctx, 0)
    gl.solve.cg(n_iter=steps)
    _lib.call("svl_event_record", par.ctx, 1)
    ms = C.c_double()
    _lib.call("svl_event_elapsed_ms", par.ctx, 0, 1, C.byref(ms))
    par.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = par.stat("launches") - l0
    R = np.dtype(wl["dtype"]).itemsize
    finite = not np.isinf(wl["kappa"])
    per_iter = (36 * R + 2) if finite else (22 * R + 2)          # fused lower bounds
    peak, peak_src = hbm_peak()
    achieved = per_iter * N * steps / (ms.value * 1e-3) / 1e9
    E = gl.solve._cg.cg_energies
    return {"metric": "cg_iters_per_s", "value": steps / (ms.value * 1e-3), "unit": "iterations/s", "n_gpus": 1,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms.value / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if wl["dtype"] is np.float32 else "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "Nx": wl["Nx"], "Ny": wl["Ny"], "seed": 1234,
                       "line_search": gl.cfg.cg_line_search, "l2": "working set exceeds the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("cg_pass_a_%s" % wl.get("key", "")), "bytes_per_node_iter": per_iter,
                         "peak_source": peak_src, "kernel": "k_cgp_a + k_cgp_b (two passes per iteration)",
                         "note": "algorithmic bytes = fused lower bound of SURVEY 8d (36R+2 finite kappa, 22R+2 kappa=inf); "
                                 "the iteration time includes the host line search"},
            "clocks": clocks, "gpu_launches": int(launches),
            "wall_ms_per_iter": 1e3 * wall / steps, "energy_first_last": [float(E[0]), float(E[-1])],
            "energies": [float(e) for e in E],
            "energy_decreasing": bool(np.all(np.diff(np.array(E, dtype=np.float64)) < 0)),
            "line_search_rescues": int(gl.solve._cg.line_search_rescues),
            "e2e": {"value": steps / wall, "unit": "iterations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 8 * (17 if finite else 5) + 8,
                    "api": "gl.solve.cg(): per iteration the 5/17 coefficients and the energy come back to the host"}}


def cg_cpu_baseline(wl):
    """NumPy port of the same iteration on a bounded sample: a 768^2 sub-problem with the same parameters (the cost
    per node does not depend on the grid size), scaled to this grid."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import glnumpy as O
    N = wl["Nx"] * wl["Ny"]
    ns = 768
    og = O.Grid(ns, ns, 0.5, 0.5, wl["dtype"])
    op = O.initial_psi(og, 1.0, 1234)
    oa, ob = O.initial_A(og, wl["H"])
    omt = hole_tiling(ns, ns) if wl["tiling"] else None
    op, oa, ob, _ = O.td_run(og, 0.1, 3, 1.0, omt, wl["kappa"], wl["sigma"], wl["H"], op, oa, ob, rand_t=1234)
    tc = time.perf_counter()
    nit = 3
    O.cg_run(og, nit, wl["kappa"], 1.0, wl["H"], omt, op, np.zeros_like(oa), np.zeros_like(ob), oa, ob, rtol=-1.0)
    tc = time.perf_counter() - tc
    return {"value": nit / tc * (ns * ns) / N, "unit": "iterations/s", "cores": 1, "kind": "port",
            "sample": "%d NumPy-oracle CG iterations on a 768^2 grid with the same parameters (%.1f s), "
                      "scaled by the node ratio to %dx%d" % (nit, tc, wl["Nx"], wl["Ny"])}


def bench_cg(args, wl, gl, par, N):
    """--workload cfg4 / cfg4s / cfg4inf.  Under torchrun the grid is split into row slabs (strong scaling): the sums of an
    iteration go over all ranks in rank order, the host line search is replicated."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    line = measure_cg(wl, gl, par, args.steps, args.warmup, args.line_search)
    line["n_gpus"] = world
    line["scaling"] = "strong" if world > 1 else "weak"
    if world > 1:
        line["roofline"]["frac"] /= world            # the bytes of an iteration are spread over the ranks
        line["roofline"]["achieved"] /= world
        line["roofline"]["note"] += "; per GPU (the grid is split over %d row slabs; three-pass iteration on slabs)" % world
    if rank == 0:
        line["cpu_baseline"] = None if (args.no_cpu_baseline or world > 1) else cg_cpu_baseline(wl)
        print(json.dumps(line))


def measure_td(wl, gl, par, steps, warmup):
    """K timed TDGL steps on one GPU (device time on the library's stream) -> the contract keys of a line."""
    from svirl_b200 import _lib
    N = wl["Nx"] * wl["Ny"]
    gl.solve.td(dt=0.1, Nt=max(warmup, 3))
    td = gl.solve._td
    par.synchronize()
    s0 = (td.sweeps_order_parameter, td.sweeps_vector_potential)
    l0 = par.stat("launches")
    sampler = ClockSampler(0)
    sampler.start()
    _lib.call("svl_event_record", par.ctx, 0)
    gl.solve.td(dt=0.1, Nt=steps)
    _lib.call("svl_event_record", par.ctx, 1)
    ms = C.c_double()
    _lib.call("svl_event_elapsed_ms", par.ctx, 0, 1, C.byref(ms))
    par.synchronize()
    clocks = sampler.stop()
    sw_psi, sw_A = td.sweeps_order_parameter - s0[0], td.sweeps_vector_potential - s0[1]
    bpsi, bA = bytes_per_node_sweep(wl)
    peak, peak_src = hbm_peak()
    achieved = (sw_psi * bpsi + sw_A * bA) * N / (ms.value * 1e-3) / 1e9
    return {"metric": METRIC, "value": N * steps / (ms.value * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms.value / steps,
            "dtype": "f32" if wl["dtype"] is np.float32 else "f64",
            "config": {"workload": wl["name"], "Nx": wl["Nx"], "Ny": wl["Ny"], "dt": 0.1, "seed": 1234},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "bytes_per_node_sweep": [bpsi, bA], "sweeps_psi": int(sw_psi),
                         "sweeps_A": int(sw_A), "traffic": None},
            "clocks": clocks, "gpu_launches": int(par.stat("launches") - l0), "replays": par.stat("replays")}


def also_blocks(args):
    """The other BASELINE configs on this GPU (N = 1 only), each beside the reference's own kernels (B-ref).  A failure in
    one block is recorded in that block and does not touch the main line."""
    out = {}

    def guarded(name, fn):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:                       # noqa: BLE001 -- the main line must survive
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        out[name]["seconds"] = time.perf_counter() - t0

    def cfg3():
        wl = workload("cfg3")
        gl = make_solver(wl)
        r = measure_td(wl, gl, gl.par, 5, 3)
        gl.par.close()
        del gl
        r["gpu_baseline"] = bref_td(wl, 3, 2)
        return r

    def cfg4s():
        wl = dict(workload("cfg4s"), key="cfg4s")
        gl = make_solver(wl)
        r = measure_cg(wl, gl, gl.par, 10, 3, "reference")
        nat = measure_cg(wl, gl, gl.par, 10, 3, "native", td_steps=0)
        r["native_line_search"] = {k: nat[k] for k in ("value", "ms_per_step", "wall_ms_per_iter", "energy_decreasing")}
        r["native_line_search"]["roofline_frac"] = nat["roofline"]["frac"]
        r["native_line_search"]["note"] = ("cfg.cg_line_search = 'native' (opt-in, SURVEY row f3): the library's trust-region "
                                           "Newton search instead of SciPy BFGS; same minimiser to ~1e-8")
        gl.par.close()
        del gl
        r["gpu_baseline"] = bref_cg(wl, 3)
        return r

    def cfg4():
        # BASELINE configs[3] at its full size (16384^2): the kernels take 4x longer than at 8192^2 while the host line
        # search does not, so this is the CG iteration's roofline fraction at the size the target is quoted on
        wl = dict(workload("cfg4"), key="cfg4")
        gl = make_solver(wl)
        r = measure_cg(wl, gl, gl.par, 6, 2, "reference")
        gl.par.close()
        del gl
        return r

    def cfg1():
        wl = workload("cfg1")
        gl = make_solver(wl)
        r = measure_td(wl, gl, gl.par, 200, 50)
        gl.par.close()
        del gl
        r["gpu_baseline"] = bref_td(wl, 200, 50)
        return r

    for name, fn in (("cfg3", cfg3), ("cfg4s", cfg4s), ("cfg4", cfg4), ("cfg1", cfg1)):
        if name in args.also.split(","):
            guarded(name, fn)
    return out


# ------------------------------------------------------------------------------------ scale mode (cfg5)
def measure_scale(wl, rank, world, local, steps, warmup, opts=()):
    """BASELINE configs[4]: grids whose fields do not fit full-size host arrays, through
    svirl_b200.scale.ScaleTD (slab-local seeded fields, row slabs over NVLink, GPU vortex count).  Needs the default
    process group when world > 1.  -> the line (dict) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from svirl_b200 import _lib
    from svirl_b200.scale import ScaleTD
    Nx, Ny = wl["Nx"], wl["Ny"] * (world if wl["scale"] == "weak" else 1)
    N = Nx * Ny
    t_build = time.perf_counter()
    st = ScaleTD(Nx, Ny, 0.5, 0.5, wl["dtype"], wl["kappa"], wl["sigma"], wl["H"], 1.0, 1234, 1.0, device_id=local,
                 distributed=world > 1)
    t_build = time.perf_counter() - t_build
    for kv in opts:
        k, v = kv.split("=")
        _lib.call("svl_set_option", st._ctx, k.encode(), int(v))

    def barrier():
        st.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st.td(0.1, max(warmup, 3))
    sampler = ClockSampler(local)
    sampler.prepare()
    barrier()
    aligned_start(world)
    s0 = (st.sweeps[0], st.sweeps[1])
    l0 = st.stat("launches")
    sampler.start()
    t0 = time.perf_counter()
    _lib.call("svl_event_record", st._ctx, 0)
    st.td(0.1, steps)
    _lib.call("svl_event_record", st._ctx, 1)
    ms = C.c_double()
    _lib.call("svl_event_elapsed_ms", st._ctx, 0, 1, C.byref(ms))
    npos, nneg = st.vortex_count()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    t = torch.tensor([ms.value, wall, float(npos), float(nneg)], device="cuda", dtype=torch.float64)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t_ms, wall = float(tm[0]), float(tm[1])
    else:
        t_ms = ms.value
    npos, nneg = int(t[2].item()), int(t[3].item())
    line = None
    if rank == 0:
        sw_psi, sw_A = st.sweeps[0] - s0[0], st.sweeps[1] - s0[1]
        bpsi, bA = bytes_per_node_sweep(wl)
        peak, peak_src = hbm_peak()
        achieved = (sw_psi * bpsi + sw_A * bA) * (N // world) / (t_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": N * steps / (t_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": max(warmup, 3), "ms_per_step": t_ms / steps, "higher_is_better": True,
                "scaling": wl["scale"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["name"], "Nx": Nx, "Ny": Ny, "dt": 0.1, "seed": 1234,
                           "l2": "working set exceeds the 126 MB L2", "rows_per_gpu": Ny // world},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": "psi Jacobi sweep", "bytes_per_node_sweep": bpsi,
                             "sweeps_psi": int(sw_psi), "sweeps_A": int(sw_A), "peak_source": peak_src},
                "cpu_baseline": None, "clocks": clocks, "gpu_launches": int(st.stat("launches") - l0),
                "vortices": {"positive": npos, "negative": nneg, "how": "GPU winding pass, reference thresholds"},
                "build_s": t_build,
                "e2e": {"value": N * steps / wall, "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": int(16 * (npos + nneg) / max(steps, 1)),
                        "api": "ScaleTD.td(dt, Nt) + ScaleTD.vortex_count(): fields stay on the GPUs, the vortex list comes back"}}
    st.close()
    return line


def bench_scale(args, wl, rank, world, local):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    line = measure_scale(wl, rank, world, local, args.steps, args.warmup, args.opt)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--psi-kernel", type=int, default=None)
    ap.add_argument("--psi-k", type=int, default=None)
    ap.add_argument("--a-kernel", type=int, default=None)
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (svl_set_option)")
    ap.add_argument("--ny-mult", type=int, default=1, help="multiply Ny (to run an N-GPU weak-scaling grid on fewer GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-instances", type=int, default=3,
                    help="independent solver instances in flight in the end-to-end (host buffer) measurement; 1 = a single one")
    ap.add_argument("--also", default="cfg3,cfg4s,cfg4,cfg1", help="extra single-GPU configs reported in `also` (default line only)")
    ap.add_argument("--no-extras", action="store_true", help="main line only: no parity / gpu_baseline / also / strong blocks")
    ap.add_argument("--line-search", default=None, choices=[None, "reference", "normalized", "native"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload)
    if args.ny_mult > 1:
        wl = dict(wl, Ny=wl["Ny"] * args.ny_mult, name=wl["name"] + " (Ny x%d)" % args.ny_mult)
    if world > 1 and args.impl == "ours" and not wl.get("scale") and not wl.get("cg"):
        wl = dict(wl, Ny=wl["Ny"] * world, name=wl["name"] + " x%d row slabs (Ny=%d)" % (world, wl["Ny"] * world))
    N = wl["Nx"] * wl["Ny"]
    cfgd = {"workload": wl["name"], "Nx": wl["Nx"], "Ny": wl["Ny"], "dt": 0.1, "seed": 1234,
            "l2": "working set (3 psi planes + a,b + flags) exceeds the 126 MB L2" if N >= 2048 * 2048 else
                  "working set fits in L2 (small-grid latency case)"}

    if args.impl == "reference":
        if rank != 0:
            return
        v, info = cpu_reference_steps(wl, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"] / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if wl["dtype"] is np.float32 else "f64", "data": "synthetic", "config": cfgd,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                 "sample": "%d TDGL steps on %s (%d Jacobi sweeps)"
                                           % (args.steps, "a %d-row band of the grid per step" % info["band"]
                                              if info.get("band") else "the full grid", info["sweeps"])},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if wl.get("scale"):
        return bench_scale(args, wl, rank, world, local)

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    from svirl_b200 import _lib
    extras = (not args.no_extras) and args.workload == "cfg2" and args.ny_mult == 1
    slab_bitwise = None
    if world > 1 and extras:
        # correctness of what is about to be timed: a 300 x 401 grid on these N slabs equals the single-GPU run bit for bit
        # (psi-only fp32 and finite-kappa fp64, with holes; tests/slab_gpu_check.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import slab_gpu_check
        try:
            slab_bitwise = bool(slab_gpu_check.check(local, verbose=False))
        except Exception as e:                   # noqa: BLE001
            slab_bitwise = "error: %s" % (str(e)[:200],)
    gl = make_solver(wl, device_id=local, slab="auto" if world > 1 else None)
    par = gl.par

    def apply_opts(par_):
        if args.psi_kernel is not None:
            par_.set_option("psi_kernel", args.psi_kernel)
        if args.psi_k is not None:
            par_.set_option("psi_k", args.psi_k)
        if args.a_kernel is not None:
            par_.set_option("a_kernel", args.a_kernel)
        for kv in args.opt:
            k, v = kv.split("=")
            par_.set_option(k, int(v))

    apply_opts(par)
    td_kw = dict(dt=0.1)
    if wl.get("cg"):
        return bench_cg(args, wl, gl, par, N)

    def barrier():
        par.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also settles the sweep-count prediction), then K timed steps
    gl.solve.td(Nt=args.warmup, **td_kw)
    td = gl.solve._td
    sampler = ClockSampler(local)
    sampler.prepare()
    barrier()
    aligned_start(world)
    s0 = (td.sweeps_order_parameter, td.sweeps_vector_potential)
    l0 = par.stat("launches")
    sampler.start()
    _lib.call("svl_event_record", par.ctx, 0)
    gl.solve.td(Nt=args.steps, **td_kw)
    _lib.call("svl_event_record", par.ctx, 1)
    ms = C.c_double()
    _lib.call("svl_event_elapsed_ms", par.ctx, 0, 1, C.byref(ms))
    clocks = sampler.stop()
    barrier()
    launches = par.stat("launches") - l0
    sw_psi = td.sweeps_order_parameter - s0[0]
    sw_A = td.sweeps_vector_potential - s0[1]
    t_ms = ms.value
    if world > 1:
        t = torch.tensor([t_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
    value = N * args.steps / (t_ms * 1e-3)          # N = nodes of the whole (all-rank) grid

    # ---- end to end through the C ABI with HOST buffers: psi up, one step, psi down, every step
    cdt = np.complex64 if wl["dtype"] is np.float32 else np.complex128
    j0, j1 = gl.cfg.slab if gl.cfg.slab is not None else (0, wl["Ny"])
    nloc = wl["Nx"] * (j1 - j0)                        # this rank's nodes
    pin_in = torch.empty(nloc * 2, dtype=torch.float32 if wl["dtype"] is np.float32 else torch.float64).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    h_in, h_out = pin_in.numpy().view(cdt), pin_out.numpy().view(cdt)
    psi_h = gl.vars.order_parameter_h().handle
    _lib.call("svl_d2h_rows", par.ctx, h_in.ctypes.data_as(C.c_void_p), psi_h, 0, int(j0), int(j1))
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _lib.call("svl_h2d_rows", par.ctx, psi_h, 0, int(j0), int(j1), h_in.ctypes.data_as(C.c_void_p))
        if gl.slab_comm is not None:
            _lib.call("svl_slab_exchange", par.ctx, psi_h)
        gl.solve.td(Nt=1, **td_kw)
        _lib.call("svl_d2h_rows", par.ctx, h_out.ctypes.data_as(C.c_void_p), psi_h, 0, int(j0), int(j1))
        h_in, h_out = h_out, h_in
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = N * e2e_steps / e2e_s
    replays = par.stat("replays")
    e2e_blk = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(N * np.dtype(cdt).itemsize),
               "d2h_bytes_per_step": int(N * np.dtype(cdt).itemsize), "steps": e2e_steps,
               "api": "svl_h2d_rows + svl_td_run(1 step) + svl_d2h_rows on pinned host buffers"}

    # ---- the same end-to-end loop for an ENSEMBLE of independent solver instances on this GPU, one host thread and
    # one library context / stream each (svirl_b200.parallel.pipeline): upload, sweeps and download of different
    # instances overlap; every instance-step still uploads its psi, runs one step and downloads the result
    if world == 1 and args.e2e_instances > 1:
        try:
            from svirl_b200.parallel.pipeline import HostStepPipeline
            M = args.e2e_instances
            sols = [gl] + [make_solver(dict(wl, seed=wl.get("seed", 1234) + k), device_id=local) for k in range(1, M)]
            for g_ in sols[1:]:
                apply_opts(g_.par)
                g_.solve.td(Nt=args.warmup, **td_kw)             # settles each instance's sweep prediction
            pins = [(torch.empty_like(pin_in).pin_memory(), torch.empty_like(pin_in).pin_memory()) for _ in range(M)]
            hin = [p_[0].numpy().view(cdt) for p_ in pins]
            hout = [p_[1].numpy().view(cdt) for p_ in pins]
            for g_, h_ in zip(sols, hin):
                _lib.call("svl_d2h_rows", g_.par.ctx, h_.ctypes.data_as(C.c_void_p), g_.vars.order_parameter_h().handle,
                          0, 0, int(wl["Ny"]))
            pipe = HostStepPipeline(sols)
            pipe.run(hin, hout, 2, **td_kw)                      # warm-up of the threads / pinned pages
            torch.cuda.synchronize()
            pipe_s = pipe.run(hin, hout, e2e_steps, **td_kw)
            torch.cuda.synchronize()
            e2e_blk = dict(e2e_blk, value=N * M * e2e_steps / pipe_s, steps=M * e2e_steps, instances=M,
                           single_instance_value=e2e_val,
                           api="svirl_b200.parallel.pipeline.HostStepPipeline over %d independent solver instances "
                               "(one host thread + context + stream each); per instance-step: svl_h2d_rows (psi up), "
                               "svl_td_run(1 step), svl_d2h_rows (psi down) on pinned host buffers; "
                               "`single_instance_value` is the same loop with one instance" % M)
            for g_ in sols[1:]:
                g_.par.close()
            del sols, pins, hin, hout, pipe
        except Exception as e:                   # noqa: BLE001 -- keep the single-instance number
            e2e_blk["ensemble_error"] = "%s: %s" % (type(e).__name__, str(e)[:300])

    # ---- BASELINE configs[4] on the same N GPUs: 32768^2 fp64 kappa=inf, STRONG scaling (all ranks take part)
    strong = None
    if extras:
        del pin_in, pin_out, h_in, h_out
        gl.par.close()
        del gl
        try:
            strong = measure_scale(workload("cfg5"), rank, world, local, 10, 3)
        except Exception as e:                   # noqa: BLE001 -- the main line must survive
            strong = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (psi Jacobi sweep)
    peak, peak_src = hbm_peak()
    bpsi, bA = bytes_per_node_sweep(wl)
    alg_bytes = (sw_psi * bpsi + sw_A * bA) * (N // world)      # per GPU: the roofline is a per-device quantity
    achieved = alg_bytes / (t_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel: read from the committed ncu capture that names the kernel
    # instantiation and command it was taken on (profiles/ncu_traffic.json); None when this run is not that command
    tr = ncu_traffic("cfg2_psi_tile") if (args.workload == "cfg2" and world == 1 and args.psi_kernel in (None, 2)
                                          and args.psi_k in (None, 4)) else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": tr["dram_bytes_per_launch"] if tr else None,
            "traffic_note": ("%s; algorithmic bytes of the same launch: %d" % (tr["what"], 4 * bpsi * (N // world))) if tr else None,
            "peak_source": peak_src,
            "kernel": "psi Jacobi sweep", "bytes_per_node_sweep": bpsi, "sweeps_psi": int(sw_psi),
            "sweeps_A": int(sw_A), "launches": int(launches),
            "avg_launch_us": 1e3 * t_ms / max(launches, 1)}

    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        nst = 25 if N <= 2048 * 2048 else 2            # ~10 s of host work at cfg2
        v, info = cpu_reference_steps(wl, nst, 0)
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
               "sample": "%d full-grid TDGL steps from the seeded initial state (%d Jacobi sweeps, %.1f s)"
                         % (nst, info["sweeps"], info["seconds"])}
        if info.get("psi") is not None:
            # parity at the benchmark's own size: the same `nst` steps from the same seeded state on the GPU
            # (a fresh solver) against the psi the reference kernel just produced on the host
            try:
                g2 = make_solver(wl, device_id=local)
                apply_opts(g2.par)                   # the kernel variant that was timed
                g2.solve.td(Nt=nst, **td_kw)
                got = g2.flatten_array(g2.vars.order_parameter)
                ref = info["psi"]
                err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))
                tol = 1e-4 if wl["dtype"] is np.float32 else 1e-10
                parity = {"max_rel": err, "tol": tol, "ok": bool(err < tol), "steps": nst, "sweeps_ref": int(info["sweeps"]),
                          "sweeps_gpu": int(g2.solve._td.sweeps_order_parameter),
                          "against": "reference kernel source on the host cores (oracle/_ref), psi after %d steps" % nst}
                g2.par.close()
                del g2
            except Exception as e:               # noqa: BLE001
                parity = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    gpu_baseline, also = None, None
    if extras and world == 1:
        try:
            gpu_baseline = bref_td(wl, max(3, min(args.steps, 20)), 3)
        except Exception as e:                   # noqa: BLE001
            gpu_baseline = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        also = also_blocks(args)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if wl["dtype"] is np.float32 else "f64", "data": "synthetic", "config": cfgd,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "e2e": e2e_blk,
            "gpu_launches": int(launches), "replays": replays,
            "psi_kernel": int(args.psi_kernel) if args.psi_kernel is not None else None,
            "parity": parity, "gpu_baseline": gpu_baseline, "slab_bitwise": slab_bitwise, "also": also, "strong": strong}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
