"""CPU: slab-local generation of the seeded initial fields (svirl_b200/scale.py) equals the full-grid
construction of the oracle, which tests/test_oracle_golden.py pins to the reference's psi0 / a0 / b0."""
import numpy as np
import pytest

import glnumpy as O
import svirl_b200.scale as sc
from conftest import load_golden


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_slab_local_initial_fields_equal_full_grid(dtype):
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


def test_c_mersenne_twister_stream_equals_numpy_legacy():
    """svl_mt19937_doubles (host helper of the library) walks numpy's legacy RandomState stream bit for bit, with
    skips that end inside, at the end of and beyond a 624-word block."""
    for seed in (1234, 5, 0):
        ref = np.random.RandomState(seed).random_sample(5000)
        for skip, n in ((0, 5000), (1, 100), (311, 700), (312, 10), (313, 1), (2000, 3000), (4999, 1)):
            s = sc._MTStream(seed, skip)
            assert np.array_equal(s.draw(n), ref[skip:skip + n])
        s = sc._MTStream(seed)
        assert np.array_equal(np.concatenate([s.draw(7), s.draw(1000), s.draw(0), s.draw(3993)]), ref)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c_seeded_psi_equals_numpy_expression(dtype):
    """svl_seeded_psi (host helper) == the reference's numpy expression (vars.py:106) bit for bit on 2^20 draws,
    including the special values u2 = 0.5 (zero phase) and u1 = 0."""
    g = sc.SeededPsi(1024, 1024, 0, 99, 1.0, dtype)
    u1, u2 = g.draws(1024)
    u1, u2 = u1.copy(), u2.copy()
    u1[:3] = (0.0, 0.5, 1.0 - 2.0 ** -53)
    u2[:3] = (0.5, 0.0, 1.0 - 2.0 ** -53)
    want = g.transform(u1.copy(), u2.copy(), 1024)                    # (Nx, rows)
    got = g.transform_rows(u1.copy(), u2.copy(), 1024)                # (rows, Nx)
    assert got.dtype == want.dtype and np.array_equal(got.T, want)
    for level in (0.3,):
        g.level = level
        assert np.array_equal(g.transform_rows(u1.copy(), u2.copy(), 1024).T, g.transform(u1.copy(), u2.copy(), 1024))
