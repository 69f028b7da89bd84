"""CPU: the vectorised triangulation (svirl_b200/observables/triangulate.py) reproduces the reference's vortex
lists BIT FOR BIT on every fixture that stores detector output (all cells offered as candidates)."""
import numpy as np
import pytest

from conftest import load_golden
from svirl_b200.observables.triangulate import triangulate

CASES = ["td_f64_k5", "td_f64_k2_tiled_eps", "td_f64_kinf", "td_f64_k3_langevin", "td_f32_kinf_tiled",
         "td_f32_k2_tiled_eps", "td_f32_k3_langevin", "cfg1_td200", "cfg1_td1000", "td_f64_k2_fixed",
         "td_f32_kinf_fixed", "td_f64_k3_fixed_nolock"]


@pytest.mark.parametrize("name", CASES)
def test_vectorised_triangulation_is_bit_identical(name):
    d = load_golden(name)
    psi, a, b = d["psi1"], d["a1"], d["b1"]
    dtype = a.dtype.type
    meta = d.get("meta", {})
    if name.startswith("cfg1"):
        dx = dy = dtype(0.5)
        H = 0.1
    else:
        dx, dy, H = dtype(meta["dx"]), dtype(meta["dy"]), meta["homogeneous_external_field"]
    Nx, Ny = psi.shape
    a_ai, b_bi = (a + d["ai0"], b + d["bi0"]) if "ai0" in d else (a, b)
    cells = np.arange((Nx - 1) * (Ny - 1))
    vx, vy, vv = triangulate(cells, psi, a, b, a_ai, b_bi, H, dx, dy, np.int32(Nx - 1), 0, dtype)
    assert np.array_equal(vv, d["obs_vv"]) and np.array_equal(vx, d["obs_vx"]) and np.array_equal(vy, d["obs_vy"])
    # a row band with an offset gives the same vortices for the cells it contains
    r0, r1 = Ny // 3, 2 * Ny // 3
    band = cells[(cells // (Nx - 1) >= r0) & (cells // (Nx - 1) < r1)]
    bx, by, bv = triangulate(band, psi[:, r0:r1 + 1], a[:, r0:r1 + 1], b[:, r0:r1], a_ai[:, r0:r1 + 1], b_bi[:, r0:r1],
                             H, dx, dy, np.int32(Nx - 1), r0, dtype)
    # compare against the full-grid result restricted to the same cells
    fx, fy, fv = triangulate(band, psi, a, b, a_ai, b_bi, H, dx, dy, np.int32(Nx - 1), 0, dtype)
    assert np.array_equal(bx, fx) and np.array_equal(by, fy) and np.array_equal(bv, fv)
