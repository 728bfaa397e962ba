"""The drop-in boundary proven from the REFERENCE's side (build container only: needs /root/reference).

1. The unmodified ``scpn_fusion.hpc.hpc_bridge.HPCBridge`` loads ``libgsb200.so`` through its own trust gate
   (``SCPN_SOLVER_LIB`` + the ``.sha256`` sidecar written by the Makefile; hpc_bridge.py:117-153,
   _hpc_native_trust.py:186-206), binds all six symbols with its own ``_setup_signatures`` (:190-250) and degrades
   exactly as documented when no GPU is present (``create_solver`` -> NULL).  A tampered digest is refused.
2. ``scpn_fusion_core_b200.providers.register`` plugs the GPU tier into the unmodified
   ``scpn_fusion.core._multi_compat`` registry (:240-300,411): with the tier available, ``dispatch`` resolves
   ``multigrid_solve`` / ``gs_rb_sor_smooth`` to this package's callables and ``dispatch_kernel_class`` to its
   ``FusionKernel``; with the tier unavailable the reference's NumPy tier is chosen (its fallback order is intact).

Each check runs in a fresh interpreter (the import shim of oracle/ref_shim.py mocks matplotlib in sys.modules).
On the GPU box the reference tree is absent and these tests skip; tests/test_gpu_abi.py exercises the same six
symbols functionally there.
"""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")
LIB = os.path.join(ROOT, "scpn_fusion_core_b200", "libgsb200.so")


def _run(code: str, env_extra: dict | None = None) -> str:
    env = dict(os.environ)
    env.pop("SCPN_SOLVER_LIB", None)
    env.pop("SCPN_SOLVER_LIB_SHA256", None)
    env.update(env_extra or {})
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "oracle"), env.get("PYTHONPATH", "")])
    pre = "import ref_shim; ref_shim.install()\n"
    p = subprocess.run([sys.executable, "-c", pre + code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    return p.stdout


def test_unmodified_hpc_bridge_loads_libgsb200_through_its_trust_gate():
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".sha256"):
        pytest.skip("libgsb200.so not built")
    out = _run(
        "import hashlib, os\n"
        "from scpn_fusion.hpc.hpc_bridge import HPCBridge\n"
        "b = HPCBridge()\n"
        "assert b.is_available(), b.load_error\n"
        "assert b.lib_path == os.environ['SCPN_SOLVER_LIB']\n"
        "assert b.lib_sha256 == hashlib.sha256(open(b.lib_path, 'rb').read()).hexdigest()\n"
        "assert b._has_converged_api and b._has_boundary_api and b._destroy_symbol == 'destroy_solver'\n"
        "for sym in ('create_solver', 'run_step', 'run_step_converged', 'set_boundary_dirichlet', 'destroy_solver', 'delete_solver'):\n"
        "    assert hasattr(b.lib, sym), sym\n"
        "import ctypes\n"
        "n = ctypes.CDLL(b.lib_path).gsb_device_count()\n"
        "b.initialize(33, 33, (2.0, 10.0), (-4.0, 4.0))\n"
        "# no CUDA device here -> create_solver returns NULL, exactly the documented degradation\n"
        "assert (b.solver_ptr is None) == (n == 0)\n"
        "b.close()\n"
        "print('bridge ok', n)\n",
        {"SCPN_SOLVER_LIB": LIB})
    assert "bridge ok" in out
    # a digest that does not match the file is refused by the reference's gate, before ctypes loads anything
    out = _run(
        "from scpn_fusion.hpc.hpc_bridge import HPCBridge\n"
        "b = HPCBridge()\n"
        "assert not b.is_available() and 'SHA-256' in (b.load_error or ''), b.load_error\n"
        "print('refused')\n",
        {"SCPN_SOLVER_LIB": LIB, "SCPN_SOLVER_LIB_SHA256": "0" * 64})
    assert "refused" in out


def test_gpu_tier_registers_into_the_unmodified_registry():
    out = _run(
        "from scpn_fusion.core import _multi_compat as multi\n"
        "from scpn_fusion.core import _multi_compat_providers  # registers the reference's own tiers\n"
        "import scpn_fusion_core_b200.providers as b200\n"
        "multi.is_available(multi.BackendTier.NUMPY)  # run the probes once\n"
        "numpy_mg = multi.dispatch('multigrid_solve')\n"
        "assert numpy_mg is not b200._gpu_multigrid_solve\n"
        "b200.register(multi)\n"
        "# tier registered but unavailable (no device here): the reference keeps its own fallback order\n"
        "multi._availability[multi.BackendTier.GPU] = False\n"
        "assert multi.dispatch('multigrid_solve') is numpy_mg\n"
        "# tier available: the GPU tier outranks NumPy for every kernel it provides\n"
        "multi._availability[multi.BackendTier.GPU] = True\n"
        "multi._dispatch_cache.clear()\n"
        "assert multi.dispatch('multigrid_solve') is b200._gpu_multigrid_solve\n"
        "assert multi.dispatch('gs_rb_sor_smooth') is b200._gpu_gs_rb_sor_smooth\n"
        "multi._class_dispatch_cache.clear()\n"
        "from scpn_fusion_core_b200 import FusionKernel\n"
        "assert multi.dispatch_kernel_class('equilibrium_kernel') is FusionKernel\n"
        "assert multi.dispatch_tier('multigrid_solve') == 'gpu'\n"
        "print('registry ok')\n")
    assert "registry ok" in out
