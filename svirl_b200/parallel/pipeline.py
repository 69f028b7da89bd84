"""Host-buffer pipelining of independent solver instances on one GPU.

A TDGL step fed from and read back to HOST memory is three resources in a row: the H2D copy engine,
the SMs, the D2H copy engine.  One solver instance keeps two of them idle at any time.  An ensemble
of M instances (independent simulations: different seeds, fields, disorder realisations -- the usual
way svirl is used for statistics) is driven here by M host threads, one library context and stream
each; ctypes releases the GIL for the duration of every library call, so the upload of one instance,
the sweeps of another and the download of a third overlap.  Every instance still performs its own
full upload, step and download per step.

The reference has nothing of the kind (its GPUArray copies are synchronous on the default stream,
svirl/storage/arrays.py:520-587); this sits above the same C-ABI calls svl_h2d_rows / svl_td_run /
svl_d2h_rows that GArray.push / TD._run / GArray.pull make."""
import ctypes as C
import threading
import time

from svirl_b200 import _lib


class HostStepPipeline(object):

    def __init__(self, solvers):
        assert len(solvers) >= 1
        self.solvers = list(solvers)
        self.errors = []

    def _worker(self, k, h_in, h_out, nsteps, td_kw, start):
        gl = self.solvers[k]
        try:
            ctx = gl.par.ctx
            psi_h = gl.vars.order_parameter_h().handle
            j0, j1 = gl.cfg.slab if gl.cfg.slab is not None else (0, int(gl.cfg.Ny))
            a, b = h_in, h_out
            start.wait()
            for _ in range(nsteps):
                _lib.call("svl_h2d_rows", ctx, psi_h, 0, int(j0), int(j1), a.ctypes.data_as(C.c_void_p))
                gl.solve.td(Nt=1, **td_kw)
                _lib.call("svl_d2h_rows", ctx, b.ctypes.data_as(C.c_void_p), psi_h, 0, int(j0), int(j1))
                a, b = b, a
        except Exception as e:                     # noqa: BLE001 -- reported by run()
            self.errors.append((k, e))
            try:
                start.abort()
            except Exception:                      # noqa: BLE001
                pass

    def run(self, host_in, host_out, nsteps, **td_kw):
        """Every instance k does `nsteps` times: psi <- host_in[k] (upload), one td() step, psi -> host_out[k]
        (download); the two host buffers of an instance swap roles after every step.  Buffers should be
        pinned.  Returns the wall-clock seconds from the common start to the moment all instances are done
        (every call of the loop is synchronous on its instance's stream, so the streams are idle then)."""
        M = len(self.solvers)
        assert len(host_in) == M and len(host_out) == M
        self.errors = []
        start = threading.Barrier(M + 1)
        th = [threading.Thread(target=self._worker, args=(k, host_in[k], host_out[k], int(nsteps), td_kw, start))
              for k in range(M)]
        for t in th:
            t.start()
        try:
            start.wait()
        except threading.BrokenBarrierError:
            pass
        t0 = time.perf_counter()
        for t in th:
            t.join()
        dt_s = time.perf_counter() - t0
        if self.errors:
            raise self.errors[0][1]
        return dt_s
