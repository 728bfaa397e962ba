"""The gates of the reference's validation/benchmark_free_boundary.py, run through the B200 path
(tools/benchmark_free_boundary.py) with the reference's thresholds."""
from __future__ import annotations

import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_free_boundary_benchmark_gates():
    import benchmark_free_boundary as fb
    res = fb.run_free_boundary_benchmark()
    assert res["gate_summary"]["failed_gates"] == [], {g: res[g] for g in res["gate_summary"]["failed_gates"]}
    assert res["passes"] and res["gate_summary"]["gate_pass_count"] == len(fb.GATES)
    assert res["single_coil"]["error_rel"] < 1e-6
    assert res["x_point"]["pass"] and res["helmholtz"]["pass"]
