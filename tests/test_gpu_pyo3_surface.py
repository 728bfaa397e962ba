"""The equilibrium names of the reference's PyO3 module served by the B200 path:
PyFusionKernel / PyEquilibriumResult / multigrid_vcycle - exercised the way the reference's
`_rust_compat.RustAcceleratedKernel` (`_rust_compat.py:83-212`) and `_rust_multigrid_solve` use them."""
from __future__ import annotations

import json

import numpy as np
import pytest

from conftest import golden, golden_cfg, rel_l2

pytestmark = pytest.mark.gpu


def test_pyfusionkernel_surface(tmp_path):
    from scpn_fusion_core_b200 import providers
    z = golden("solves")
    cfg = golden_cfg(z, "iter65")
    path = tmp_path / "cfg.json"
    path.write_text(json.dumps(cfg))
    rk = providers.PyFusionKernel(str(path))
    nr, nz = rk.grid_shape()
    assert (nr, nz) == (65, 65)
    R, Z = np.asarray(rk.get_r()), np.asarray(rk.get_z())
    assert R.shape == (nr,) and Z.shape == (nz,) and np.all(np.diff(R) > 0) and np.all(np.diff(Z) > 0)
    assert np.asarray(rk.get_psi()).shape == (nz, nr) and np.all(np.isfinite(rk.get_psi()))
    assert rk.solver_method() == "multigrid"
    res = rk.solve_equilibrium()
    meta, topo = z["iter65_meta"], z["iter65_topo"]
    assert abs(res.iterations - int(meta[0])) <= 1 and res.converged == bool(meta[1])
    assert rel_l2(rk.get_psi(), z["iter65_psi"]) <= 1e-9
    assert rel_l2(rk.get_j_phi(), z["iter65_jphi"]) <= 1e-8
    assert abs(res.axis_r - topo[0]) <= 1e-6 and abs(res.axis_z - topo[1]) <= 1e-6
    assert abs(res.x_point_r - topo[3]) <= 1e-6 and abs(res.x_point_z - topo[4]) <= 1e-6
    assert res.solve_time_ms > 0 and "EquilibriumResult(converged=" in repr(res)
    with pytest.raises(AttributeError):
        res.converged = False
    psi = rk.get_psi()
    psi[:] = 0.0                                   # accessors hand out copies, like into_pyarray(clone)
    assert np.any(rk.get_psi() != 0.0)
    rk.set_solver_method("picard_sor")
    assert rk.solver_method() == "sor"
    rk.set_solver_method("MG")
    assert rk.solver_method() == "multigrid"
    with pytest.raises(ValueError):
        rk.set_solver_method("newton")
    with pytest.raises(NotImplementedError):
        rk.calculate_thermodynamics(50.0)
    with pytest.raises(OSError):
        providers.PyFusionKernel(str(tmp_path / "missing.json"))


def test_rs_multigrid_vcycle_name():
    from scpn_fusion_core_b200 import providers
    z = golden("mg_solve")
    src, bc = z["g65_source"], z["g65_bc"]
    psi, res, cyc, conv = providers.multigrid_vcycle(src, bc, 4.0, 8.0, -4.0, 4.0, 65, 65, 1e-8, 60)
    np.testing.assert_array_equal(psi, z["g65_psi"])
    assert (res, cyc, float(conv)) == tuple(z["g65_meta"])
