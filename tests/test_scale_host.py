"""CPU: slab-local generation of the seeded initial fields (svirl_b200/scale.py) equals the full-grid
construction of the oracle, which tests/test_oracle_golden.py pins to the reference's psi0 / a0 / b0."""
import numpy as np
import pytest

import glnumpy as O
import svirl_b200.scale as sc
from conftest import load_golden


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_slab_local_initial_fields_equal_full_grid(dtype, monkeypatch):
    monkeypatch.setattr(sc, "_CHUNK", 777)            # force many skip chunks
    Nx, Ny = 37, 29
    g = O.Grid(Nx, Ny, 0.5, 0.4, dtype)
    full = O.initial_psi(g, 1.0, 1234)
    a, b = O.initial_A(g, 0.1)
    for r0, r1 in ((0, Ny), (0, 1), (5, 13), (Ny - 9, Ny), (Ny - 1, Ny)):
        assert np.array_equal(sc.seeded_psi_rows(Nx, Ny, r0, r1, 1234, 1.0, dtype), full[:, r0:r1])
        aa, bb = sc.symmetric_gauge_rows(Nx, Ny, 0.5, 0.4, 0.1, r0, r1, dtype)
        assert np.array_equal(aa, a[:, r0:r1]) and np.array_equal(bb, b[:, r0:min(r1, Ny - 1)])


def test_slab_local_fields_equal_reference_fixture():
    d = load_golden("td_f64_k5")
    m = d["meta"]
    p = sc.seeded_psi_rows(m["Nx"], m["Ny"], 3, 17, m["random_seed"], 1.0, np.float64)
    assert np.array_equal(p, d["psi0"][:, 3:17])
    a, b = sc.symmetric_gauge_rows(m["Nx"], m["Ny"], m["dx"], m["dy"], m["homogeneous_external_field"], 3, 17, np.float64)
    assert np.array_equal(a, d["a0"][:, 3:17]) and np.array_equal(b, d["b0"][:, 3:17])
