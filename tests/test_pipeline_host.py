"""CPU: host logic of svirl_b200.parallel.pipeline.HostStepPipeline with stand-in solvers (no GPU, no library calls):
every instance performs upload, step, download `nsteps` times in that order on its own thread, the two host buffers of
an instance swap roles after every step, instances run concurrently, and an error in one instance surfaces in run()."""
import threading
import time

import numpy as np
import pytest

import svirl_b200.parallel.pipeline as pipeline


class _FakeHandle(object):
    def __init__(self, owner):
        self.handle = owner


class _FakeSolver(object):
    """Keeps 'device' psi in self.dev; td() adds 1 to it.  The pipeline reaches the library through _lib.call, which
    the test replaces by `fake_call`."""

    def __init__(self, k, n, fail_at=None):
        self.k, self.dev, self.log, self.fail_at, self.steps = k, np.zeros(n), [], fail_at, 0
        self.par = type("P", (), {"ctx": self})()
        self.vars = type("V", (), {"order_parameter_h": lambda s_: _FakeHandle(self)})()
        self.cfg = type("Cfg", (), {"slab": None, "Ny": 4})()
        self.solve = type("S", (), {"td": self._td})()
        self.threads = set()

    def _td(self, Nt=1, **kw):
        assert Nt == 1 and kw == {"dt": 0.1}
        self.threads.add(threading.get_ident())
        self.steps += 1
        if self.fail_at is not None and self.steps == self.fail_at:
            raise RuntimeError("instance %d failed" % self.k)
        time.sleep(0.002)
        self.dev += 1.0
        self.log.append("td")


def _as_array(p, n):
    import ctypes as C
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))


def fake_call(name, ctx, *args):
    sol = ctx
    if name == "svl_h2d_rows":          # (handle, part, r0, r1, host pointer)
        sol.dev[:] = _as_array(args[4], sol.dev.size)
        sol.log.append("h2d")
    elif name == "svl_d2h_rows":        # (host pointer, handle, part, r0, r1)
        _as_array(args[0], sol.dev.size)[:] = sol.dev
        sol.log.append("d2h")
    else:
        raise AssertionError(name)


def test_pipeline_order_buffers_and_concurrency(monkeypatch):
    monkeypatch.setattr(pipeline._lib, "call", fake_call)
    n, M, nsteps = 16, 3, 5
    sols = [_FakeSolver(k, n) for k in range(M)]
    hin = [np.full(n, 10.0 * k) for k in range(M)]
    hout = [np.full(n, -1.0) for _ in range(M)]
    secs = pipeline.HostStepPipeline(sols).run(hin, hout, nsteps, dt=0.1)
    assert secs > 0
    for k, s in enumerate(sols):
        assert s.log == ["h2d", "td", "d2h"] * nsteps
        last = hout[k] if nsteps % 2 == 1 else hin[k]          # buffers swap roles every step
        assert np.all(last == 10.0 * k + nsteps) and np.all(s.dev == 10.0 * k + nsteps)
        assert len(s.threads) == 1
    assert len({next(iter(s.threads)) for s in sols}) == M     # one host thread per instance
    assert secs < 0.9 * M * nsteps * 0.002 + 0.05              # the instances overlapped (sequential would be M * nsteps * 2 ms)


def test_pipeline_reports_the_failure_of_an_instance(monkeypatch):
    monkeypatch.setattr(pipeline._lib, "call", fake_call)
    sols = [_FakeSolver(0, 4), _FakeSolver(1, 4, fail_at=2)]
    hin, hout = [np.zeros(4), np.zeros(4)], [np.zeros(4), np.zeros(4)]
    with pytest.raises(RuntimeError, match="instance 1 failed"):
        pipeline.HostStepPipeline(sols).run(hin, hout, 3, dt=0.1)
